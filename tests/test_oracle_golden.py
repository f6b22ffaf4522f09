"""Pins the CPU oracle (oracle/lepton_oracle.c) against files written by the UNMODIFIED reference.

tests/golden/*.lep were written by oracle/_ref/lepton (the reference CLI compiled from /root/reference by
oracle/Makefile.ref); manifest.json holds sha256s of the coefficient planes from the reference's own -ujg dump.
 * oracle decode of the reference's segment streams must reproduce the reference's coefficient planes;
 * oracle encode of those planes must reproduce the reference's segment streams byte for byte.
The reference repository's own golden vector images/iphone16.lep (test_suite/test_16threads.sh) is checked in-container.
"""
import os

import numpy as np
import pytest

import lepfmt
import oracle
from helpers import GOLDEN, MANIFEST, geometry_of, golden_leps, load_lep, oracle_decode_planes, plane_hashes, segments_of


@pytest.mark.parametrize("name", golden_leps())
def test_oracle_roundtrips_reference_streams(name):
    lf = load_lep(name)
    planes, streams = oracle_decode_planes(lf)
    src = MANIFEST.get(name[:-4] + ".jpg") or MANIFEST.get(MANIFEST.get(name, {}).get("source", ""), {})
    if "plane_sha256" in src:
        assert plane_hashes(planes) == src["plane_sha256"], "oracle decode != reference -ujg coefficient dump"
    g, _, _ = geometry_of(lf)
    for i, (y0, y1, last) in enumerate(segments_of(lf)):
        rc, s, nd = oracle.encode_segment(g, planes, y0, y1, last)
        assert rc == 0
        assert s == streams[i], "oracle encode != reference stream for segment %d" % i
        assert nd > 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/images"), reason="reference tree only exists in the build container")
@pytest.mark.slow
def test_reference_repo_golden_vector_iphone16(tmp_path):
    """images/iphone16.lep (16 thread segments) is the reference's own golden vector: test_suite/test_16threads.sh
    expects it to decode to iphone.jpg (md5 8ea9fcf1b2c24877aa838dd6ac1df413).  The oracle must decode its 16
    streams to exactly the coefficient planes the reference dumps for iphone.jpg, and re-encode them bit-exactly."""
    import hashlib
    import subprocess
    from conftest import REF_LEPTON
    jpg = "/root/reference/images/iphone.jpg"
    assert hashlib.md5(open(jpg, "rb").read()).hexdigest() == "8ea9fcf1b2c24877aa838dd6ac1df413"
    lf = lepfmt.parse_container(open("/root/reference/images/iphone16.lep", "rb").read())
    assert lf.nseg == 16
    planes, streams = oracle_decode_planes(lf)
    ujg = str(tmp_path / "a.ujg")
    assert subprocess.run([REF_LEPTON, "-ujg", "-skipverify", jpg, ujg], capture_output=True).returncode == 0
    _, ref_planes = lepfmt.parse_ujg_planes(open(ujg, "rb").read())
    for a, b in zip(planes, ref_planes):
        assert np.array_equal(a, b)
    g, _, _ = geometry_of(lf)
    for i, (y0, y1, last) in enumerate(segments_of(lf)):
        rc, s, _ = oracle.encode_segment(g, planes, y0, y1, last)
        assert rc == 0 and s == streams[i]


def test_branch_update_matches_reference_semantics():
    """Branch::record_obs_and_update corner cases (src/vp8/model/branch.hh:82-100) through a tiny stream."""
    # An all-zero 1x1-block grayscale image exercises the identity priors; encode/decode must agree.
    q = [[16] * 64]
    g = oracle.make_geometry(1, [1], [1], 1, q)
    planes = [np.zeros((1, 64), dtype=np.int16)]
    rc, s, nd = oracle.encode_segment(g, planes, 0, 1, True)
    assert rc == 0 and nd == 6 + 3 + 3 + 1
    out = [np.ones((1, 64), dtype=np.int16)]
    rc, nd2 = oracle.decode_segment(g, out, 0, 1, True, s)
    assert rc == 0 and nd2 == nd and not out[0].any()


def test_out_of_range_coefficient_is_rejected():
    """COEFFICIENT_OUT_OF_RANGE (exit code 6): src/vp8/encoder/encoder.cc:124,265,343."""
    q = [[16] * 64]
    g = oracle.make_geometry(1, [1], [1], 1, q)
    planes = [np.zeros((1, 64), dtype=np.int16)]
    planes[0][0, 3] = 4096
    rc, _, _ = oracle.encode_segment(g, planes, 0, 1, True)
    assert rc == 6


@pytest.mark.skipif(not os.path.isdir("/root/reference/images"), reason="reference tree only exists in the build container")
@pytest.mark.slow
def test_oracle_vs_live_reference_on_a_large_fixture(tmp_path):
    import subprocess
    from conftest import REF_LEPTON
    jpg = "/root/reference/images/iphonecrop.jpg"
    lep, ujg = str(tmp_path / "a.lep"), str(tmp_path / "a.ujg")
    assert subprocess.run([REF_LEPTON, "-skipverify", jpg, lep], capture_output=True).returncode == 0
    assert subprocess.run([REF_LEPTON, "-ujg", "-skipverify", jpg, ujg], capture_output=True).returncode == 0
    lf = lepfmt.parse_container(open(lep, "rb").read())
    _, planes = lepfmt.parse_ujg_planes(open(ujg, "rb").read())
    g, _, _ = geometry_of(lf)
    streams = lepfmt.demux(lf.payload, lf.version)
    assert lf.nseg == 4
    for i, (y0, y1, last) in enumerate(segments_of(lf)):
        rc, s, _ = oracle.encode_segment(g, planes, y0, y1, last)
        assert rc == 0 and s == streams[i]
