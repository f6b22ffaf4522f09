#!/bin/bash
# Round m (last GPU minutes of round 2): the GPU suite on the final tree (device container assembly on by default), file-API
# A/B with the .lep bytes of every setting compared with the default's (kernel A capped at 5 CTAs per SM so that the other
# chunks' small kernels find room; host MuxWriter), then the default bench line.
#   gpurun --timeout 460 -- 'bash tests/tools_round_m.sh > gpurun_out/round_m.log 2>&1'
mkdir -p gpurun_out
date +%s > gpurun_out/m_t0
el() { echo "$(( $(date +%s) - $(cat gpurun_out/m_t0) )) s"; }
echo "== pytest -m gpu"; timeout 240 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== file API, 4096 files ($(el))"
timeout 150 python tests/tools_e2e2.py 4096 "" "LEPB200_ENC_CTA_CAP=5" "LEPB200_ENC_CTA_CAP=5,LEPB200_CHUNK_SPLIT=2" "LEPB200_ENC_CTA_CAP=4" "LEPB200_DEVICE_MUX=0" 2>&1 | grep -v "^\[trace\] *$" | cut -c1-200
echo "== default bench ($(el))"
timeout 200 python bench.py 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default_m.json; cut -c1-7000 gpurun_out/bench_default_m.json; tail -3 gpurun_out/bench_default.err
echo "== done ($(el))"
