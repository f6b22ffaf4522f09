#!/bin/bash
# range coder (table pass with the corrected address, deeper token prefetch in both forms): parity, kernel times, file API timelines.
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -4
echo "== device legs"
for m in 1 0; do
LEPB200_TRACE=1 LEPB200_RC_MODE=$m timeout 600 python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/dev_$m.err | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('rc_mode $m  kernel A ms', round(e['roofline']['kernel_ms'],1), ' range coder ms', round(e['roofline']['rangecode_kernel_ms'],1), ' encode MB/s', round(e['value'],1), ' decode ms', round(k['ms_per_step'],1), ' round trip', d.get('roundtrip_pass_rate'))
except Exception as ex: print('no result', ex)"
grep "range coder" gpurun_out/dev_$m.err | tail -1
done
echo "== file API, 4096 files"
timeout 900 python tests/tools_e2e2.py 4096 "" "LEPB200_RC_MODE=0" "LEPB200_CHUNKS_IN_FLIGHT=2" 2>&1 | grep -E "^==|huffman kernels|range coder|kernel A|front |back |containers|pack|decode|fetch|-- decompress" | cut -c1-200
