#!/usr/bin/env python
"""Stage EVERY file of the reference's images/ directory for the GPU box (tests/golden/_refimages/, git-ignored but
shipped with the snapshot like the built .so files), with what the UNMODIFIED reference CLI does to each of them.

Run by __graft_entry__.build() in the build container (needs /root/reference/images and oracle/_ref/lepton).  The images
are the reference's own test data (BASELINE.json: "bit-exact round-trip on every file in images/"); nothing here is
reference source code.  expected.json holds, per file:

  jpg_md5 / jpg_size       the input
  rc_skipverify            exit code of `lepton -unjailed -skipverify -allowprogressive in.jpg out.lep`
  rc_verify                exit code of the same without -skipverify (41 = ROUNDTRIP_FAILURE, roundtripfail.jpg)
  exit_name                the ExitCode name the reference wrote to stderr when it left through custom_exit with an error
                           (src/vp8/util/memory.cc:238-245).  The NUMBER the shell sees is not stable across kernels:
                           custom_exit ends with syscall(SYS_exit) (memory.cc:246-247), which ends one thread, so the
                           process status is that of whichever thread leaves last (42 in one build container, 0 with
                           an empty output file in another, for the same binary and input) -- the name is
  lep_md5 / lep_size       the .lep the reference wrote (only when rc_skipverify == 0 and the file is non-empty)
  back_md5                 md5 of what the reference decodes that .lep to (== jpg_md5 unless the file does not round-trip)
  status_want              the status the LIBRARY must report for the file: 0, or the ExitCode of the reference process
                           that meets the error (src/vp8/util/memory.hh:13-39): arithmetic.jpg 42 UNSUPPORTED_JPEG
                           (jpgcoder.cc:2911-2925 "image is coded arithm."), badzerorun.jpg 1 ASSERTION_FAILURE
                           (jpgcoder.cc:4951) -- Makefile.am:302-304,357-359 expect them to fail
and for the reference repository's golden .lep vectors (iphone16.lep, gold-legacy.lep, narrowrst.lep) the md5 its own
test scripts pin for the decoded JPEG (test_suite/test_16threads.sh, test_legacy.sh, test_future_compat.sh).
"""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_IMAGES = "/root/reference/images"
LEPTON = os.path.join(ROOT, "oracle", "_ref", "lepton")
OUT = os.path.join(HERE, "_refimages")

STATUS_WANT = {"arithmetic.jpg": 42, "badzerorun.jpg": 1}
GOLDEN_LEP_MD5 = {"iphone16.lep": "8ea9fcf1b2c24877aa838dd6ac1df413", "gold-legacy.lep": "9ffbfc24d1157d0b1ed7a9b53bef4c23",
                  "narrowrst.lep": "07e9021d35114bd69f44f5bc1c3788e3"}


def md5(b):
    return hashlib.md5(b).hexdigest()


def up_to_date():
    exp = os.path.join(OUT, "expected.json")
    if not os.path.exists(exp):
        return False
    if "exit_name" not in json.load(open(exp)).get("arithmetic.jpg", {}):
        return False
    have = set(os.listdir(OUT))
    return all(n in have for n in os.listdir(REF_IMAGES))


def main(force=False):
    if not (os.path.isdir(REF_IMAGES) and os.path.exists(LEPTON)):
        return False
    if not force and up_to_date():
        return True
    os.makedirs(OUT, exist_ok=True)
    expected = {}
    with tempfile.TemporaryDirectory() as td:
        for name in sorted(os.listdir(REF_IMAGES)):
            src = os.path.join(REF_IMAGES, name)
            dst = os.path.join(OUT, name)
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
            data = open(src, "rb").read()
            if name.endswith(".lep"):
                back = os.path.join(td, "g.jpg")
                rc = subprocess.run([LEPTON, "-unjailed", src, back], capture_output=True).returncode
                got = md5(open(back, "rb").read()) if rc == 0 else None
                assert got == GOLDEN_LEP_MD5[name], (name, rc, got)          # the reference still meets its own golden md5
                expected[name] = {"lep_md5": md5(data), "lep_size": len(data), "decoded_md5": GOLDEN_LEP_MD5[name]}
                continue
            e = {"jpg_md5": md5(data), "jpg_size": len(data)}
            lep = os.path.join(td, "o.lep")
            for key, flags in (("rc_skipverify", ["-skipverify"]), ("rc_verify", [])):
                if os.path.exists(lep):
                    os.unlink(lep)
                run = subprocess.run([LEPTON, "-unjailed", "-allowprogressive"] + flags + [src, lep], capture_output=True)
                e[key] = run.returncode
                names = re.findall(rb"^([A-Z][A-Z0-9_]{3,})$", run.stderr, re.M)
                e["exit_name" if key == "rc_skipverify" else "exit_name_verify"] = names[-1].decode() if names else None
                if key == "rc_skipverify" and e[key] == 0 and os.path.getsize(lep) > 0:
                    ld = open(lep, "rb").read()
                    e.update(lep_md5=md5(ld), lep_size=len(ld))
                    back = os.path.join(td, "b.jpg")
                    rc = subprocess.run([LEPTON, "-unjailed", lep, back], capture_output=True).returncode
                    e["back_md5"] = md5(open(back, "rb").read()) if rc == 0 else None
            e["status_want"] = STATUS_WANT.get(name, 0)
            assert (e["status_want"] == 0) == ("lep_md5" in e), (name, e)
            expected[name] = e
            print(name, e.get("lep_size"), e["rc_skipverify"], e["rc_verify"], e["exit_name"], flush=True)
    json.dump(expected, open(os.path.join(OUT, "expected.json"), "w"), indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    sys.exit(0 if main(force="--force" in sys.argv) else 1)
