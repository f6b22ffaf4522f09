#!/bin/bash
# dense first-exponent-bit table: parity of all kernels on the device, then the device-resident legs.
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('dense first-bit table  kernel A ms', round(e['roofline']['kernel_ms'],1), ' B ms', round(e['roofline']['rangecode_kernel_ms'],1), ' decode ms', round(k['ms_per_step'],1), k['roofline']['kernel'][:24], ' round trip', d.get('roundtrip_pass_rate'))
except Exception as ex: print('no result', ex)"
LEPB200_DEC_MODE=1 timeout 600 python bench.py --distinct 32 --images 1024 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('1024 images, warp decode kernel: kernel A ms', round(e['roofline']['kernel_ms'],1), ' decode ms', round(k['ms_per_step'],1), ' (round 1: 510 ms)')
except Exception as ex: print('no result', ex)"
