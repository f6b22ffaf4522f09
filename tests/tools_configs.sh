#!/bin/bash
# BASELINE.json configs 3, 4, 5 through the file-level API on one GPU (per-GPU share of the 8-GPU workloads, reduced where noted).
#   gpurun --timeout 2400 -- 'bash tests/tools_configs.sh > gpurun_out/configs.log 2>&1'
mkdir -p gpurun_out
echo "== config 5: decode-only thumbnails, 131072 files in calls of 1024"; timeout 900 python bench.py --config 5 2>gpurun_out/cfg5.err | tail -1 > gpurun_out/r02_bench_config5.json; cut -c1-1800 gpurun_out/r02_bench_config5.json; tail -2 gpurun_out/cfg5.err
echo "== config 5, calls of 256"; timeout 600 python bench.py --config 5 --images 32768 --batch 256 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_config5_b256.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_config5_b256.json')); print({k: d.get(k) for k in ('value','images_per_s','latency')})"
echo "== config 4: 128 x 4K 4:4:4 progressive, encode"; timeout 900 python bench.py --config 4 --e2e-steps 2 2>gpurun_out/cfg4.err | tail -1 > gpurun_out/r02_bench_config4.json; cut -c1-1800 gpurun_out/r02_bench_config4.json; tail -2 gpurun_out/cfg4.err
echo "== config 3: mixed 256^2-4096^2, ${CFG3_IMAGES:-2048} files (share of one GPU: 8192), encode+decode"; timeout 1500 python bench.py --config 3 --images ${CFG3_IMAGES:-2048} --e2e-steps 2 2>gpurun_out/cfg3.err | tail -1 > gpurun_out/r02_bench_config3.json; cut -c1-1800 gpurun_out/r02_bench_config3.json; tail -2 gpurun_out/cfg3.err
