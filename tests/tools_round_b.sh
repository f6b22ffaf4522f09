#!/bin/bash
# e2e settings sweep + config 5 with several calls in flight.
#   gpurun --timeout 1500 -- 'bash tests/tools_round_b.sh > gpurun_out/round_b.log 2>&1'
mkdir -p gpurun_out
echo "== e2e sweep (chunks in flight, kernel A CTA cap)"; timeout 900 python tests/tools_e2e.py 4096 1,0 2,0 3,0 4,0 2,5 2>&1 | grep -v Warning
echo "== config 5, 4 calls of 1024 in flight"; timeout 600 python bench.py --config 5 --images 65536 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_config5_s4.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_config5_s4.json')); print({k: d.get(k) for k in ('value','images_per_s','latency')})"
echo "== config 5, 8 calls of 256 in flight"; timeout 600 python bench.py --config 5 --images 32768 --batch 256 --streams 8 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_config5_b256_s8.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_config5_b256_s8.json')); print({k: d.get(k) for k in ('value','images_per_s','latency')})"
