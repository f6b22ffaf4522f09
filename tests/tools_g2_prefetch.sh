#!/bin/bash
# Decode kernel v2/v3: build-time variants on the GPU box.
#   gpurun --timeout 1500 -- 'bash tests/tools_g2_prefetch.sh > gpurun_out/g2_variants.log 2>&1'
mkdir -p gpurun_out
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "group_kernel and 5" 2>&1 | tail -3
run() {  # $1 label, $2 lanes
  LEPB200_DEC_MODE=5 LEPB200_DEC_LANES=$2 timeout 600 python bench.py --images 4096 --distinct 32 --no-e2e --no-cpu-baseline --steps 2 --warmup 2 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1 lanes $2  decode ms', round(d['decode']['ms_per_step'],1), ' MB/s', round(d['decode']['value'],1), ' round trip', d.get('roundtrip_pass_rate'))
except Exception as e: print('$1 lanes $2: no result', e)"
}
for cfg in ${CFGS:-"2 3" "0 3" "1 3"}; do
  set -- $cfg
  LEPB200_G2_PREFETCH=$1 LEPB200_G2_PF_DIST=$2 python -m lepton_b200.build --force > /dev/null 2>&1 || { echo "build failed: $cfg"; continue; }
  run "prefetch=$1 dist=$2" 4
  if [ "$cfg" = "${NCU_CFG:-2 3}" ]; then
    run "prefetch=$1 dist=$2" 8
    run "prefetch=$1 dist=$2" 2
    LEPB200_DEC_MODE=5 LEPB200_DEC_LANES=4 timeout 900 ncu --section SourceCounters --section WarpStateStats --section MemoryWorkloadAnalysis --section SchedulerStats --section LaunchStats --section Occupancy --section SpeedOfLight --clock-control none --import-source on -k regex:lep_decode_g2 -s 1 -c 1 -o gpurun_out/dec_g3_pf$1_g4 \
      python bench.py --images 4096 --distinct 32 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
  fi
done
python -m lepton_b200.build --force > /dev/null 2>&1
ls -la gpurun_out | tail -4
