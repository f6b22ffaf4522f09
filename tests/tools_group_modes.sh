#!/bin/bash
# Group decode kernel (LEPB200_DEC_MODE=4) on the GPU box: parity, then the decode leg of bench.py per group size.
#   gpurun --timeout 1500 -- 'bash tests/tools_group_modes.sh > gpurun_out/group_modes.log 2>&1'
mkdir -p gpurun_out
if [ -z "$SKIP_PARITY" ]; then echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "group_kernel" 2>&1 | tail -3; fi
for lanes in ${LANES:-4 8 16 2}; do
  LEPB200_DEC_MODE=${MODE:-4} LEPB200_DEC_LANES=$lanes timeout 600 python bench.py --images ${IMAGES:-4096} --distinct 32 --no-e2e --no-cpu-baseline --steps 2 --warmup 2 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('mode ${MODE:-4} lanes $lanes  decode ms', round(d['decode']['ms_per_step'],1), ' MB/s', round(d['decode']['value'],1), ' round trip', d.get('roundtrip_pass_rate'))
except Exception as e: print('lanes $lanes: no result', e)"
done
if [ -n "$NCU_LANES" ]; then
  echo "== ncu full capture, lanes $NCU_LANES, ${NCU_IMAGES:-1024} images"
  LEPB200_DEC_MODE=${MODE:-4} LEPB200_DEC_LANES=$NCU_LANES timeout 900 ncu ${NCU_SET:---set full} --clock-control none --import-source on -k regex:lep_decode_g -s 1 -c 1 -o gpurun_out/dec_mode${MODE:-4}_g${NCU_LANES}_full \
    python bench.py --images ${NCU_IMAGES:-1024} --distinct 32 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
  ls -la gpurun_out | tail -3
fi
