#!/bin/bash
# occupancy re-tune of the encode / decode kernels (rebuilds on the GPU box)
for cfg in "7 5" "6 4" "6 6"; do
  set -- $cfg
  LEPB200_ENC_MINBLOCKS=$1 LEPB200_DEC_MINBLOCKS=$2 python -m lepton_b200.build --force >/dev/null 2>&1
  python bench.py --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('enc_minblocks $1 dec_minblocks $2', 'A ms', round(d['roofline']['kernel_ms'],1), 'B ms', round(d['roofline']['rangecode_kernel_ms'],1), 'decode ms', round(d['decode']['ms_per_step'],1))"
done
