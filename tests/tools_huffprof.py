"""Diagnostic: one GPU-Huffman chunk (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from lepton_b200 import LeptonB200FileCodec
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
distinct = bench.make_corpus(8)
jpegs = [distinct[i % 8] for i in range(n)]
fc = LeptonB200FileCodec(0, host_threads=8, chunk_images=n, gpu_huffman=True)
r = fc.compress(jpegs, copy=False)
print(fc.last_timing())
