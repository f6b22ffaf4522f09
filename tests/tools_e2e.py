"""Diagnostic (not a test): e2e file-level compress throughput for several co-scheduling / chunk settings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from lepton_b200 import LeptonB200FileCodec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
distinct = bench.make_corpus(16)
jpegs = [distinct[i % 16] for i in range(n)]
tot = sum(len(j) for j in jpegs)
handle = LeptonB200FileCodec.prepare(jpegs)
combos = [(0, 4, 1024), (5, 7, 1024), (6, 4, 1024), (5, 4, 1024), (4, 7, 1024), (5, 8, 1024), (5, 7, 512)]
if len(sys.argv) > 2:
    combos = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]
for cap, hw, chunk in combos:
    os.environ["LEPB200_ENC_CTA_CAP"] = str(cap)
    os.environ["LEPB200_HUFF_WARPS"] = str(hw)
    fc = LeptonB200FileCodec(0, host_threads=16, chunk_images=chunk, gpu_huffman=True)
    r = fc.compress(handle, copy=False)
    assert all(st == 0 for st, _ in r)
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        fc.compress(handle, copy=False)
        best = min(best, time.perf_counter() - t0)
    print("enc_cta_cap %d huff_warps %d chunk %4d  %.3f s  %.0f MB/s  stages %s" % (cap, hw, chunk, best, tot / best / 1e6, fc.last_timing()), flush=True)
    fc.close()
