"""GPU parity over EVERY file of the reference's images/ directory (BASELINE.json: "bit-exact .jpg<->.lep round-trip on
every file in images/"; /root/reference/Makefile.am:238-362 are the reference's own tests over these files).

tests/golden/_refimages/ is staged by __graft_entry__.build() in the build container (tests/golden/make_refimages.py) and
travels to the GPU box with the snapshot; expected.json holds what the UNMODIFIED reference CLI did with each file.  All
files go through the CUDA path by the file-level C ABI; comparison is by md5 of whole files (bit-exact)."""
import hashlib
import json
import os

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REFIMG = os.path.join(HERE, "golden", "_refimages")


def md5(b):
    return hashlib.md5(b).hexdigest()


@pytest.fixture(scope="module")
def expected():
    p = os.path.join(REFIMG, "expected.json")
    assert os.path.exists(p), ("tests/golden/_refimages/ is missing: run __graft_entry__.build() where /root/reference exists "
                               "(the staged images travel with the snapshot)")
    return json.load(open(p))


@pytest.fixture(scope="module")
def compressed(expected):
    """All JPEGs of images/ (good and expected-failure ones alike) in ONE batch through lepb200_compress_jpegs."""
    from lepton_b200 import LeptonB200FileCodec
    names = sorted(n for n in expected if n.endswith(".jpg"))
    jpegs = [open(os.path.join(REFIMG, n), "rb").read() for n in names]
    fc = LeptonB200FileCodec(0, host_threads=8)
    res = fc.compress(jpegs)
    launches = fc.kernel_launches
    fc.close()
    assert launches > 0
    return names, jpegs, res


def test_every_reference_image_compresses_to_the_reference_lep(expected, compressed):
    names, jpegs, res = compressed
    good = 0
    for n, j, (st, lep) in zip(names, jpegs, res):
        e = expected[n]
        assert md5(j) == e["jpg_md5"], n
        assert st == e["status_want"], (n, st, e["status_want"])
        if e["status_want"] == 0:
            assert len(lep) == e["lep_size"] and md5(lep) == e["lep_md5"], "%s: .lep differs from the reference CLI's" % n
            good += 1
        else:
            assert lep == b"", n
    assert good == 26          # iphone.jpg (BASELINE config 1), hq, slr*, iphonecity, iphonecrop, trunc included


def test_expected_failure_exit_codes(expected, compressed):
    """Makefile.am:302-304 (arithmetic: EXPECT_FAILURE) and :357-359 (badzerorun: EXPECT_FAILURE): the reference process
    that meets the error leaves with UNSUPPORTED_JPEG (42) / ASSERTION_FAILURE (1, the assert at jpgcoder.cc:4951)."""
    names, _, res = compressed
    st = {n: s for n, (s, _) in zip(names, res)}
    assert st["arithmetic.jpg"] == 42
    assert st["badzerorun.jpg"] == 1
    # what the live reference CLI said in the build container.  Its process status is not a stable witness (custom_exit
    # ends ONE thread with SYS_exit, memory.cc:246-247: 42 or 0 depending on which thread leaves last), the name it
    # writes first (memory.cc:238-245) and the empty output are
    assert expected["arithmetic.jpg"]["exit_name"] == "UNSUPPORTED_JPEG" and "lep_md5" not in expected["arithmetic.jpg"]
    assert expected["badzerorun.jpg"]["rc_skipverify"] != 0 and "lep_md5" not in expected["badzerorun.jpg"]


def test_every_reference_image_round_trips(expected, compressed):
    """.lep -> .jpg through lepb200_decompress_leps (GPU arithmetic decode, GPU or host Huffman re-encode): equal to the
    input for every file the reference round-trips, equal to the REFERENCE's (different) decoding for roundtripfail.jpg."""
    from lepton_b200 import LeptonB200FileCodec
    names, jpegs, res = compressed
    ok = [(n, j, lep) for n, j, (st, lep) in zip(names, jpegs, res) if st == 0]
    fc = LeptonB200FileCodec(0, host_threads=8)
    back = fc.decompress([lep for _, _, lep in ok])
    fc.close()
    for (n, j, _), (st, out) in zip(ok, back):
        assert st == 0, (n, st)
        assert md5(out) == expected[n]["back_md5"], "%s: restored JPEG differs from the reference's decoding" % n
        if n != "roundtripfail.jpg":
            assert out == j, n


def test_roundtripfail_with_verify_is_withheld(expected):
    """test_suite/test_roundtrip.sh territory: with validation on (the reference CLI's default) the file exits 41."""
    from lepton_b200 import LeptonB200FileCodec
    assert expected["roundtripfail.jpg"]["rc_verify"] == 41 or expected["roundtripfail.jpg"]["exit_name_verify"] == "ROUNDTRIP_FAILURE"
    fc = LeptonB200FileCodec(0, host_threads=4, verify=True)
    data = [open(os.path.join(REFIMG, n), "rb").read() for n in ("iphonecrop.jpg", "roundtripfail.jpg", "trunc.jpg")]
    res = fc.compress(data)
    fc.close()
    assert [st for st, _ in res] == [0, 41, 0]
    assert md5(res[0][1]) == expected["iphonecrop.jpg"]["lep_md5"] and md5(res[2][1]) == expected["trunc.jpg"]["lep_md5"]


def test_reference_golden_lep_vectors_decode_to_the_pinned_md5(expected):
    """The reference repository's own golden vectors: iphone16.lep (16 thread-segments, test_suite/test_16threads.sh) and
    gold-legacy.lep (test_suite/test_legacy.sh) must decode to the md5 those scripts pin; so must narrowrst.lep
    (test_suite/test_future_compat.sh), a version-4 container whose header blob is brotli-coded (read through the system's
    libbrotlidec; where that library is missing the file must be REFUSED with status 200 -- never produce bytes for it)."""
    import ctypes
    from lepton_b200 import lib
    from lepton_b200 import LeptonB200FileCodec
    names = ["iphone16.lep", "gold-legacy.lep", "narrowrst.lep"]
    fc = LeptonB200FileCodec(0, host_threads=4)
    back = fc.decompress([open(os.path.join(REFIMG, n), "rb").read() for n in names])
    fc.close()
    for n, (st, out) in zip(names[:2], back[:2]):
        assert st == 0, (n, st)
        assert md5(out) == expected[n]["decoded_md5"], n
    L = lib()
    L.lepb200_host_brotli_available.restype = ctypes.c_int
    if L.lepb200_host_brotli_available() == 1:
        assert back[2][0] == 0 and md5(back[2][1]) == expected["narrowrst.lep"]["decoded_md5"] == "07e9021d35114bd69f44f5bc1c3788e3"
    else:
        assert back[2][0] == 200 and back[2][1] == b""
