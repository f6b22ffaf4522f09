"""Diagnostic (not a test): BASELINE config 5 in miniature -- decode-only .lep -> .jpg of 640x480 thumbnails.
Every image is one thread-segment, i.e. one serial chain: the batch latency is the chain latency, the throughput comes
from the number of chains in flight."""
import io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from PIL import Image
from lepton_b200 import LeptonB200FileCodec

rng = np.random.default_rng(5)
distinct = []
for k in range(32):
    y, x = np.mgrid[0:480, 0:640]
    a = np.clip((128 + 60 * np.sin(x / (11.0 + k)) + 50 * np.cos(y / (7.0 + k)))[..., None] + rng.normal(0, 14, (480, 640, 3)), 0, 255).astype(np.uint8)
    b = io.BytesIO(); Image.fromarray(a).save(b, "JPEG", quality=85, subsampling=2); distinct.append(b.getvalue())
fc = LeptonB200FileCodec(0, host_threads=16)
for n in (256, 2048, 16384):
    jpegs = [distinct[i % 32] for i in range(n)]
    tot = sum(len(j) for j in jpegs)
    leps = [l for _, l in fc.compress(fc.prepare(jpegs), copy=True)]
    h = fc.prepare(leps)
    back = fc.decompress(h, copy=True)
    ok = sum(int(st == 0 and b == j) for (st, b), j in zip(back, jpegs))
    t0 = time.perf_counter(); fc.decompress(h, copy=False); dt = time.perf_counter() - t0
    print("thumbnails %6d: decompress %.1f ms per batch (= per-image latency)  %.0f MB/s  %.0f images/s  roundtrip %d/%d  segments per image %d" %
          (n, dt * 1e3, tot / dt / 1e6, n / dt, ok, n, 1), fc.last_timing(), flush=True)
