"""Diagnostic (not a test): one 4096-file compress / decompress through the file API for several environment settings, with
the stage timeline (LEPB200_TRACE) of one call each.

    python tests/tools_e2e2.py [files] "K=V,K=V" "K=V" ...        (E2E_CHUNK_IMAGES=n as one of the K=V: files per chunk)
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from lepton_b200 import LeptonB200FileCodec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
combos = sys.argv[2:] or [""]
distinct = bench.make_corpus(2, 32)
jpegs = [distinct[i % 32] for i in range(n)]
tot = sum(len(j) for j in jpegs)
handle = LeptonB200FileCodec.prepare(jpegs)
leps = None
for combo in combos:
    kv = dict(x.split("=") for x in combo.split(",") if x)
    for k, v in kv.items():
        os.environ[k] = v
    ci = int(os.environ.get("E2E_CHUNK_IMAGES", "0"))
    fc = LeptonB200FileCodec(0, host_threads=16, chunk_images=ci) if ci else LeptonB200FileCodec(0, host_threads=16)
    r = fc.compress(handle, copy=True)
    assert all(st == 0 for st, _ in r)
    if leps is None:
        leps = [b for _, b in r]
        lhandle = LeptonB200FileCodec.prepare(leps)
    else:                     # every setting must write the bytes of the first one
        assert [b for _, b in r] == leps, "setting %r changes the .lep bytes" % combo
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        fc.compress(handle, copy=False)
        best = min(best, time.perf_counter() - t0)
    bd = 1e9
    if os.environ.get("E2E_DECOMPRESS", "1") == "1":
        fc.decompress(lhandle, copy=False)
        for _ in range(2):
            t0 = time.perf_counter()
            fc.decompress(lhandle, copy=False)
            bd = min(bd, time.perf_counter() - t0)
    print("== %-60s compress %.3f s  %.0f MB/s   decompress %.3f s  %.0f MB/s" % (combo or "(default)", best, tot / best / 1e6, bd, tot / bd / 1e6), flush=True)
    sys.stderr.flush()
    os.environ["LEPB200_TRACE"] = "1"
    fc.compress(handle, copy=False)
    if os.environ.get("E2E_DECOMPRESS", "1") == "1":
        sys.stderr.write("[trace] -- decompress\n")
        fc.decompress(lhandle, copy=False)
    del os.environ["LEPB200_TRACE"]
    sys.stderr.flush()
    fc.close()
    for k in kv:
        del os.environ[k]
