"""Diagnostic (not a test): file-level throughput of one 4096-file call for several settings of the chunk pipeline, in one
process (the settings are read when a codec is created).

    python tests/tools_e2e.py [files] [chunks_in_flight,enc_cta_cap ...]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from lepton_b200 import LeptonB200FileCodec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
distinct = bench.make_corpus(2, 32)
jpegs = [distinct[i % 32] for i in range(n)]
tot = sum(len(j) for j in jpegs)
handle = LeptonB200FileCodec.prepare(jpegs)
combos = [(1, 0), (2, 0), (3, 0), (4, 0), (4, 5), (2, 5)]
if len(sys.argv) > 2:
    combos = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]
leps = None
for w, cap in combos:
    os.environ["LEPB200_CHUNKS_IN_FLIGHT"] = str(w)
    os.environ["LEPB200_ENC_CTA_CAP"] = str(cap)
    fc = LeptonB200FileCodec(0, host_threads=16)
    r = fc.compress(handle, copy=leps is None)
    assert all(st == 0 for st, _ in r)
    if leps is None:
        leps = [b for _, b in r]
        lhandle = LeptonB200FileCodec.prepare(leps)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        fc.compress(handle, copy=False)
        best = min(best, time.perf_counter() - t0)
    fc.decompress(lhandle, copy=False)
    bd = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        fc.decompress(lhandle, copy=False)
        bd = min(bd, time.perf_counter() - t0)
    print("chunks in flight %d  enc_cta_cap %d   compress %.3f s  %.0f MB/s   decompress %.3f s  %.0f MB/s" % (w, cap, best, tot / best / 1e6, bd, tot / bd / 1e6), flush=True)
    fc.close()
