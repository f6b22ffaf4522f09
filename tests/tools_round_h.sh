#!/bin/bash
# sub-sequence Huffman decode + register-window range coder pieces: parity, then timelines and kernel times.
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refimages.py -q -x 2>&1 | tail -4
echo "== file API, 4096 files: timelines"
timeout 900 python tests/tools_e2e2.py 4096 "LEPB200_HUFF_PAR=0" "LEPB200_HUFF_PAR=1" "LEPB200_HUFF_PAR=1,LEPB200_HUFF_SUBSEQ_BITS=2048" "LEPB200_HUFF_PAR=1,LEPB200_HUFF_SUBSEQ_BITS=8192" "LEPB200_HUFF_PAR=1,LEPB200_CHUNKS_IN_FLIGHT=2" "LEPB200_HUFF_PAR=1,LEPB200_CHUNKS_IN_FLIGHT=1" "LEPB200_HUFF_PAR=1,LEPB200_RC_MODE=0" 2>&1 | grep -v "^\[trace\] *$" | cut -c1-200
echo "== device legs"
for m in 1 0; do
LEPB200_RC_MODE=$m timeout 600 python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('rc_mode $m  kernel A ms', round(e['roofline']['kernel_ms'],1), ' range coder ms', round(e['roofline']['rangecode_kernel_ms'],1), ' encode MB/s', round(e['value'],1), ' decode ms', round(k['ms_per_step'],1), ' round trip', d.get('roundtrip_pass_rate'))
except Exception as ex: print('no result', ex)"
done
echo "== decode of 8192 images in one launch (32768 segments)"
LEPB200_DEC_THREADS=32768 timeout 900 python bench.py --distinct 32 --images 8192 --no-e2e --no-cpu-baseline --steps 2 --warmup 2 2>gpurun_out/b8192.err | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('8192 images: encode ms', round(e['ms_per_step'],1), 'MB/s', round(e['value'],1), ' decode ms', round(k['ms_per_step'],1), 'MB/s', round(k['value'],1), ' round trip', d.get('roundtrip_pass_rate'))
except Exception as ex: print('no result', ex)"; tail -2 gpurun_out/b8192.err
