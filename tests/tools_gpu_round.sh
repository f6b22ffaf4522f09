#!/bin/bash
# One GPU call: the GPU test suite, then the default bench line (and the same with one chunk in flight for comparison).
#   gpurun --timeout 2400 -- 'bash tests/tools_gpu_round.sh > gpurun_out/gpu_round.log 2>&1'
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "== default bench"; timeout 1200 python bench.py 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json; cat gpurun_out/bench_default.json | cut -c1-6000; tail -3 gpurun_out/bench_default.err
if [ -n "$COMPARE_CHUNKS" ]; then
  echo "== bench, one chunk in flight"; LEPB200_CHUNKS_IN_FLIGHT=1 timeout 1200 python bench.py --distinct 32 --no-cpu-baseline --steps 2 2>/dev/null | tail -1 > gpurun_out/bench_chunks1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_chunks1.json')); print('e2e one chunk in flight:', json.dumps(d.get('e2e'))[:1500])"
fi
