#!/bin/bash
# Re-entry round: the GPU suite, the default bench line, the file-API timelines of the shipped defaults against the
# serial range coder / one-warp-per-image Huffman kernel, the launch list, and ncu captures of the new kernels.
#   gpurun --timeout 900 -- 'bash tests/tools_round_k.sh > gpurun_out/round_k.log 2>&1'
mkdir -p gpurun_out
date +%s > gpurun_out/k_t0
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== default bench ($(( $(date +%s) - $(cat gpurun_out/k_t0) )) s)"
timeout 480 python bench.py 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json; cut -c1-7000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== file API, 4096 files ($(( $(date +%s) - $(cat gpurun_out/k_t0) )) s)"
timeout 240 python tests/tools_e2e2.py 4096 "" "LEPB200_RC_MODE=0" "LEPB200_HUFF_PAR=0" 2>&1 | grep -v "^\[trace\] *$" | cut -c1-220
echo "== launch list of one default step ($(( $(date +%s) - $(cat gpurun_out/k_t0) )) s)"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_final.csv \
  python bench.py --distinct 32 --no-cpu-baseline --steps 1 --warmup 1 --e2e-steps 1 > /dev/null 2>&1
grep -c lep_ gpurun_out/r02_launches_bench_final.csv
echo "== ncu, new kernels, one chunk of 1024 files ($(( $(date +%s) - $(cat gpurun_out/k_t0) )) s)"
E2E_DECOMPRESS=0 LEPB200_CHUNKS_IN_FLIGHT=1 timeout 240 ncu --set full --import-source on --clock-control none -k regex:"huffpar_sync|huffpar_write|lep_rangepass|lep_rangepiece|lep_rangenorm" -c 8 -o gpurun_out/r02_huffpar_rangecoder -f python tests/tools_e2e2.py 1024 "" > gpurun_out/ncu_k.out 2>&1; tail -3 gpurun_out/ncu_k.out
ls -la gpurun_out/*.ncu-rep | tail -3
echo "== done ($(( $(date +%s) - $(cat gpurun_out/k_t0) )) s)"
