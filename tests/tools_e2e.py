"""Diagnostic (not a test): e2e file-level compress throughput for several host-thread / chunk settings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from lepton_b200 import LeptonB200FileCodec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
distinct = bench.make_corpus(16)
jpegs = [distinct[i % 16] for i in range(n)]
tot = sum(len(j) for j in jpegs)
for threads, chunk, gh in [(16, 512, False), (16, 512, True), (16, 1024, True), (16, 256, True), (8, 512, True), (32, 512, True)]:
    fc = LeptonB200FileCodec(0, host_threads=threads, chunk_images=chunk, gpu_huffman=gh)
    r = fc.compress(jpegs, copy=False)
    assert all(st == 0 for st, _ in r)
    t0 = time.perf_counter()
    fc.compress(jpegs, copy=False)
    dt = time.perf_counter() - t0
    print("gpu_huffman %s threads %3d chunk %4d  %.3f s  %.0f MB/s  stages %s" % (gh, threads, chunk, dt, tot / dt / 1e6, fc.last_timing()), flush=True)
    fc.close()
