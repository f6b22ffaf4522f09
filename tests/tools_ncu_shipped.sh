#!/bin/bash
# ncu captures of the shipped build at the bench size (4096 images), one kernel each: kernel A, kernel B, the decode
# kernel (device-resident legs), Huffman decode / encode (file-level legs).  Limited section set: ~10 replay passes.
#   gpurun --timeout 2400 -- 'bash tests/tools_ncu_shipped.sh > gpurun_out/ncu_shipped.log 2>&1'
mkdir -p gpurun_out
SEC="--section SourceCounters --section WarpStateStats --section MemoryWorkloadAnalysis --section SchedulerStats --section LaunchStats --section Occupancy --section SpeedOfLight"
echo "== kernel A, kernel B, decode kernel (2nd launch of each where there is one)"
timeout 1500 ncu $SEC --clock-control none --import-source on -k regex:"lep_encode_kernel|lep_rangecode_kernel|lep_decode_g2" -s 2 -c 3 -o gpurun_out/r02_shipped_abd \
  python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_abd.out 2>&1; tail -2 gpurun_out/ncu_abd.out | cut -c1-300
echo "== Huffman decode / encode kernels (file-level legs, 4 chunks of 1024 files on the way in, one chunk on the way back)"
timeout 1200 ncu $SEC --clock-control none --import-source on -k regex:"lep_huffdecode|lep_huffencode" -c 6 -o gpurun_out/r02_shipped_huff \
  python bench.py --distinct 32 --e2e-only --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_huff.out 2>&1; tail -2 gpurun_out/ncu_huff.out | cut -c1-300
echo "== launch list of one default step (timing of every kernel, cold cache, serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --distinct 32 --no-cpu-baseline --steps 1 --warmup 1 --e2e-steps 1 > /dev/null 2>&1
ls -la gpurun_out | tail -6
