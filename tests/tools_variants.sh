#!/bin/bash
# Diagnostic: rebuild with different occupancy targets and time the kernels (run on the GPU box).
for v in "5 5" "6 6" "7 8" "7 10"; do
  set -- $v
  LEPB200_ENC_MINBLOCKS=$1 LEPB200_DEC_MINBLOCKS=$2 python lepton_b200/build.py --force > /dev/null 2>&1
  python bench.py --images 1024 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > /tmp/v.json 2>/tmp/v.err
  python -c "
import json; d=json.load(open('/tmp/v.json')); print('enc_minblocks $1 dec_minblocks $2: A %.1f ms  B %.1f ms  decode %.1f ms' % (d['roofline']['kernel_ms'], d['roofline']['rangecode_kernel_ms'], d['decode']['ms_per_step']))"
done
python lepton_b200/build.py --force > /dev/null 2>&1
