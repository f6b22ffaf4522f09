import os, sys, time
sys.path.insert(0, '/root/repo')
import bench
from lepton_b200 import LeptonB200FileCodec
n = 6000
distinct = bench.make_corpus(16)
jpegs = [distinct[i % 16] for i in range(n)]
fc = LeptonB200FileCodec(0, host_threads=16)
h = fc.prepare(jpegs)
t0 = time.perf_counter(); r = fc.compress(h, copy=True); t1 = time.perf_counter()
assert all(st == 0 for st, _ in r)
leps = [b for _, b in r]
t2 = time.perf_counter(); r = fc.compress(h, copy=False); t3 = time.perf_counter()
tot = sum(len(j) for j in jpegs)
print("compress 6000 (2 chunks): %.3f s  %.0f MB/s" % (t3 - t2, tot / (t3 - t2) / 1e6), fc.last_timing())
lh = fc.prepare(leps)
back = fc.decompress(lh, copy=True)
ok = sum(int(st == 0 and b == j) for (st, b), j in zip(back, jpegs))
t4 = time.perf_counter(); fc.decompress(lh, copy=False); t5 = time.perf_counter()
print("decompress: %.3f s  %.0f MB/s  roundtrip %d/%d gpu_recoded %d" % (t5 - t4, tot / (t5 - t4) / 1e6, ok, n, fc.last_gpu_recoded), fc.last_timing())
import torch
print("device mem allocated GB (free,total):", [x / 2**30 for x in torch.cuda.mem_get_info()])
