#!/bin/bash
# Everything that was written without a GPU at hand, in one call:
#   gpurun --timeout 2400 -- 'bash tests/tools_round2_first_call.sh > gpurun_out/first_call.log 2>&1'
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== reference CLI with the adapters plugged in (exit fix + row entry)"
LEPB200_TEST_PLUG=1 timeout 300 python -m pytest tests -m gpu -q -k adapters 2>&1 | tail -5
echo "== CLI quick check"; timeout 300 bash tests/tools_quick_cli_check.sh | grep -v "^TS_\|^TIMING\|^START\|^Read took"
echo "== kernel candidates and build-time options"; bash tests/tools_kernel_modes.sh
echo "== default bench line"; timeout 900 python bench.py 2>/dev/null | tail -1
