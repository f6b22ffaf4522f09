#!/bin/bash
# kernel A with its per-component tables in shared memory against the previous build; decode kernel crossover by batch size.
#   gpurun --timeout 1500 -- 'bash tests/tools_round_c.sh > gpurun_out/round_c.log 2>&1'
mkdir -p gpurun_out
run() {  # label, env..., bench args
  label=$1; shift
  env "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['encode']; k=d['decode']
    print('$label  kernel A ms', round(e['roofline']['kernel_ms'],1), ' B ms', round(e['roofline']['rangecode_kernel_ms'],1), ' decode ms', round(k['ms_per_step'],1), k['roofline']['kernel'][:24], ' round trip', d.get('roundtrip_pass_rate'))
except Exception as ex: print('$label: no result', ex)"
}
B="timeout 600 python bench.py --distinct 32 --no-e2e --no-cpu-baseline --steps 3 --warmup 3"
run "base build (tables in global memory)" LEPB200_LIBRARY=$PWD/lepton_b200/variants/libbase.so $B
run "tables in shared memory            " $B
for n in 2048 3072; do
  run "images $n warp kernel " LEPB200_DEC_MODE=1 $B --images $n
  run "images $n group kernel" LEPB200_DEC_MODE=2 $B --images $n
done
