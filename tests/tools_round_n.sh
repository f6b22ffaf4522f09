#!/bin/bash
# Last GPU seconds of round 2: ncu --set full of the gather kernel (device container assembly) and of the Huffman encode
# kernel at one chunk of 1024 bench images.
mkdir -p gpurun_out
LEPB200_HENC_PARTS=1 LEPB200_CHUNKS_IN_FLIGHT=1 timeout 110 ncu --set full --import-source on --clock-control none -k regex:"lep_huffencode|lep_gather" -c 5 -o gpurun_out/r02_henc_gather -f python tests/tools_e2e2.py 1024 "" > gpurun_out/ncu_n.out 2>&1
tail -3 gpurun_out/ncu_n.out; ls -la gpurun_out/r02_henc_gather.ncu-rep
