#!/bin/bash
# Round l (one call for everything): GPU suite; A/B of the cp.async token feed, the front turns, chunk sizes and the
# Huffman-encode parts through the file API (timelines); device legs with both feeds; default bench line; launch list;
# ncu of the range pass.
#   gpurun --timeout 840 -- 'bash tests/tools_round_l.sh > gpurun_out/round_l.log 2>&1'
mkdir -p gpurun_out
date +%s > gpurun_out/l_t0
el() { echo "$(( $(date +%s) - $(cat gpurun_out/l_t0) )) s"; }
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== file API, 4096 files ($(el))"
timeout 200 python tests/tools_e2e2.py 4096 "" "LEPB200_RC_FEED=0" "LEPB200_FRONT_TURNS=0" "LEPB200_CHUNK_SPLIT=2" "E2E_CHUNK_IMAGES=888" "LEPB200_HENC_PARTS=1" 2>&1 | grep -v "^\[trace\] *$" | cut -c1-200
echo "== device legs ($(el))"
for m in 1 0; do
LEPB200_TRACE=1 LEPB200_RC_FEED=$m timeout 100 python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/dev_$m.err | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('rc_feed $m  kernel A ms', round(e['roofline']['kernel_ms'],1), ' range coder ms', round(e['roofline']['rangecode_kernel_ms'],1), ' encode MB/s', round(e['value'],1), ' decode ms', round(k['ms_per_step'],1), ' round trip', d.get('roundtrip_pass_rate'))
except Exception as ex: print('no result', ex)"
grep "range coder" gpurun_out/dev_$m.err | tail -1
done
echo "== default bench ($(el))"
timeout 300 python bench.py 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json; cut -c1-7000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== ncu, range pass, one chunk of 1024 files ($(el))"
E2E_DECOMPRESS=0 LEPB200_CHUNKS_IN_FLIGHT=1 timeout 120 ncu --set full --import-source on --clock-control none -k regex:"lep_rangepass" -c 1 -o gpurun_out/r02_rangepass_async -f python tests/tools_e2e2.py 1024 "" > gpurun_out/ncu_l.out 2>&1; tail -2 gpurun_out/ncu_l.out
echo "== launch list of one default step ($(el))"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_l.csv \
  python bench.py --distinct 32 --no-cpu-baseline --steps 1 --warmup 1 --e2e-steps 1 > /dev/null 2>&1
grep -c lep_ gpurun_out/r02_launches_bench_l.csv
echo "== done ($(el))"
