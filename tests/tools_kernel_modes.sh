#!/bin/bash
# Decode kernel comparison on the GPU box (diagnostic): parity of the lock-step kernel first, then the decode leg of
# bench.py for the kernels (0 warp per segment, 1 thread per segment, 2 lock step, 3 lock step + warp kernel side by side), then a launch list and one full ncu capture of the lock-step kernel.
#   gpurun --timeout 1500 -- 'bash tests/tools_kernel_modes.sh'
# Everything runs under `timeout` so that a kernel that does not terminate costs one step, not the box.
mkdir -p gpurun_out
echo "== parity, LEPB200_DEC_MODE=2"
# decode modes 2 and 3, encode mode 1
LEPB200_TEST_LOCKSTEP=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "lockstep or thread_per_segment" 2>&1 | tail -3
echo "== decode leg, 1024 and 4096 images"
for images in 4096; do
  for cfg in "0 16384 50" "2 16384 50" "2 8192 50" "3 16384 50"; do
    set -- $cfg; mode=$1; thr=$2; split=$3
    for once in 1; do
      LEPB200_DEC_MODE=$mode LEPB200_DEC_THREADS=$thr LEPB200_DEC_SPLIT=$split timeout 600 python bench.py --images $images --no-e2e --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('images $images mode $mode threads $thr split $split  decode ms', round(d['decode']['ms_per_step'],1), ' MB/s', round(d['decode']['value'],1), ' round trip', d.get('roundtrip_pass_rate'))
except Exception as e: print('images $images mode $mode: no result', e)"
    done
  done
done
echo "== kernel A: warp per segment (0) against the lock-step thread-per-segment kernel (1)"
for images in 4096; do
  for mode in 0 1; do
    LEPB200_ENC_MODE=$mode timeout 600 python bench.py --images $images --no-e2e --no-cpu-baseline --no-decode --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('images $images enc mode $mode  value MB/s', round(d['value'],1), ' kernel A ms', round(d['roofline']['kernel_ms'],1), ' B ms', round(d['roofline']['rangecode_kernel_ms'],1))
except Exception as e: print('images $images enc mode $mode: no result', e)"
  done
done
echo "== build-time options on the default kernels: streaming cache hints, position-innermost model layout (rebuilds the library in place)"
for opt in "LEPB200_STREAM_HINTS=1" "LEPB200_MODEL_LAYOUT=1"; do
  env $opt python -m lepton_b200.build --force > /dev/null 2>&1 || { echo "build failed: $opt"; continue; }
  timeout 600 python bench.py --images 4096 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$opt  value MB/s', round(d['value'],1), ' kernel A ms', round(d['roofline']['kernel_ms'],1), ' decode ms', round(d['decode']['ms_per_step'],1), ' round trip', d.get('roundtrip_pass_rate'))
except Exception as e: print('$opt: no result', e)"
done
python -m lepton_b200.build --force > /dev/null 2>&1    # back to the default build
echo "== lock-step kernel: launch list + full capture (256 images)"
LEPB200_DEC_MODE=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_dec_mode2.csv \
  python bench.py --images 256 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
LEPB200_DEC_MODE=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:lep_decode_lockstep -c 1 -o gpurun_out/dec_lockstep_full \
  python bench.py --images 256 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out | tail -5
