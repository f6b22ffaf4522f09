#!/bin/bash
# parallel range coder: parity on the device, then kernel B time (serial form for comparison).
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
for m in 1 0; do
LEPB200_RC_MODE=$m timeout 600 python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('rc_mode $m  kernel A ms', round(e['roofline']['kernel_ms'],1), ' range coder ms', round(e['roofline']['rangecode_kernel_ms'],1), ' encode MB/s', round(e['value'],1), ' decode ms', round(k['ms_per_step'],1), ' round trip', d.get('roundtrip_pass_rate'))
except Exception as ex: print('no result', ex)"
done
LEPB200_RC_MODE=1 timeout 600 python bench.py --distinct 32 --images 1024 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['encode']; print('1024 images: kernel A ms', round(e['roofline']['kernel_ms'],1), ' range coder ms', round(e['roofline']['rangecode_kernel_ms'],1))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lep_range|lep_digit" -c 12 --csv python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 2>/dev/null | grep -E "lep_range|lep_digit" | cut -d, -f5,15 | head -12
