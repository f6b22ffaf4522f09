#!/bin/bash
# decode kernel with precomputed candidate addresses: parity on the device, then the decode leg.
mkdir -p gpurun_out
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "group_kernel or both_decode" 2>&1 | tail -3
for lanes in 4 8; do
LEPB200_DEC_LANES=$lanes timeout 600 python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('lanes $lanes  kernel A ms', round(e['roofline']['kernel_ms'],1), ' decode ms', round(k['ms_per_step'],1), k['roofline']['kernel'][:24], ' round trip', d.get('roundtrip_pass_rate'))
except Exception as ex: print('no result', ex)"
done
LEPB200_DEC_LANES=4 timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:lep_decode_g2 -s 1 -c 1 -o gpurun_out/dec_g4_cand \
  python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out | tail -3
