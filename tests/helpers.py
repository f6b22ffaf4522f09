"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import hashlib
import json
import os

import numpy as np

import lepfmt
import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))


def golden_leps():
    """All committed reference-written .lep files (baseline and progressive)."""
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".lep"))


def load_lep(name):
    return lepfmt.parse_container(open(os.path.join(GOLDEN, name), "rb").read())


def geometry_of(lf):
    f = lf.frame
    tbcv, tbc = lepfmt.truncation(lf)
    q = [f.qtables[f.qidx[c]] for c in range(f.ncmp)]
    return oracle.make_geometry(f.ncmp, f.bch, f.bcv, f.mcuv, q, tbcv, tbc), tbcv, tbc


def segments_of(lf):
    hs = lf.handoffs
    return [(h.luma_y_start, h.luma_y_end, i == len(hs) - 1) for i, h in enumerate(hs)]


def oracle_decode_planes(lf):
    """Coefficient planes of a .lep, decoded by the ORACLE from the reference-written streams."""
    f = lf.frame
    g, _, _ = geometry_of(lf)
    streams = lepfmt.demux(lf.payload, lf.version)
    planes = [np.zeros((f.bch[c] * f.bcv[c], 64), dtype=np.int16) for c in range(f.ncmp)]
    for i, (y0, y1, last) in enumerate(segments_of(lf)):
        rc, _ = oracle.decode_segment(g, planes, y0, y1, last, streams[i])
        assert rc == 0, (i, rc)
    return planes, streams


def plane_hashes(planes):
    return [hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in planes]
