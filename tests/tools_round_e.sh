#!/bin/bash
# decode group kernel (candidate addresses prepared ahead) with and without line requests: prebuilt variants, no compiling on the box.
mkdir -p gpurun_out
for v in pf0 pf2 pf1; do
LEPB200_LIBRARY=$PWD/lepton_b200/variants/lib$v.so timeout 600 python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('$v  kernel A ms', round(e['roofline']['kernel_ms'],1), ' decode ms', round(k['ms_per_step'],1), k['roofline']['kernel'][:24], ' round trip', d.get('roundtrip_pass_rate'))
except Exception as ex: print('no result', ex)"
done
LEPB200_LIBRARY=$PWD/lepton_b200/variants/lib${NCU_VARIANT:-pf2}.so timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:lep_decode_g2 -s 1 -c 1 -o gpurun_out/dec_g4_cand_${NCU_VARIANT:-pf2} \
  python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out | tail -2
