#!/usr/bin/env python
"""Regenerate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference/images and oracle/_ref/lepton, built by
`make -C oracle`).  For every fixture JPEG it stores

  <name>.jpg   the reference's own test image (data fixture, copied verbatim)
  <name>.lep   what the reference CLI writes for it (`lepton -skipverify in.jpg out.lep`)

and records in manifest.json the reference's exit code, sizes, md5s, and a sha256 of each coefficient
plane taken from the reference's `-ujg` dump (so our own JPEG front end can be pinned without shipping
the multi-megabyte dumps).  The reference repo's own golden .lep vectors are exercised in-container by
tests/test_oracle_golden.py (iphone16.lep, test_suite/test_16threads.sh); narrowrst.lep (version 4, brotli
header) and gold-legacy.lep (pre-handoff legacy header) need the brotli / legacy container readers, which are
outside the hot path.

prog8x8.jpg is the one fixture that is not a reference image: a single 8x8 block of noise saved by Pillow as an
optimised progressive JPEG (quality 92), whose AC tables hold no end-of-band code; prog8x8.lep is the reference
CLI's output for it (make_prog8x8 below; kept as committed, not regenerated, because the bytes depend on the
Pillow/libjpeg build).
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lepfmt  # noqa: E402

REF_IMAGES = "/root/reference/images"
LEPTON = os.path.join(ROOT, "oracle", "_ref", "lepton")

# small fixtures only (a few MB in total); the large ones are exercised in-container by tests that read
# /root/reference/images directly when it exists.
FIXTURES = [
    "android.jpg", "androidcrop.jpg", "androidcropoptions.jpg", "androidtrail.jpg", "colorswap.jpg", "gray2sf.jpg",
    "grayscale.jpg", "iphonecrop2.jpg", "narrowrst.jpg", "nofsync.jpg", "trailingrst.jpg", "trailingrst2.jpg",
    "truncatedzerorun.jpg", "singlerowtrunc.jpg", "androidprogressive.jpg", "iphoneprogressive.jpg",
    "iphoneprogressive2.jpg",
]


# extra multi-segment variants of small colour images: (output stem, source, extra reference flags)
VARIANTS = [
    ("android_t4", "android.jpg", ["-minencodethreads=4"]),
    ("iphonecrop2_t8", "iphonecrop2.jpg", ["-minencodethreads=8"]),
    ("androidcrop_t2", "androidcrop.jpg", ["-minencodethreads=2"]),
]


def md5(b):
    return hashlib.md5(b).hexdigest()


def make_prog8x8():
    """One-block progressive JPEG with optimised tables + the reference's .lep for it (only when absent)."""
    dst = os.path.join(HERE, "prog8x8.jpg")
    if os.path.exists(dst):
        return
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(7)
    a = np.clip(128 + rng.normal(0, 40, (8, 8, 3)), 0, 255).astype(np.uint8)
    Image.fromarray(a).save(dst, "JPEG", quality=92, progressive=True, optimize=True, subsampling=0)
    rc = subprocess.run([LEPTON, "-skipverify", "-allowprogressive", dst, dst[:-4] + ".lep"], capture_output=True).returncode
    assert rc == 0, rc


def main():
    make_prog8x8()
    manifest = {}
    with tempfile.TemporaryDirectory() as td:
        for name in FIXTURES:
            src = os.path.join(REF_IMAGES, name)
            stem = name[:-4]
            dst_jpg = os.path.join(HERE, name)
            shutil.copyfile(src, dst_jpg)
            os.chmod(dst_jpg, 0o644)
            lep = os.path.join(td, stem + ".lep")
            ujg = os.path.join(td, stem + ".ujg")
            rc = subprocess.run([LEPTON, "-skipverify", "-allowprogressive", src, lep], capture_output=True).returncode
            entry = {"jpg_md5": md5(open(src, "rb").read()), "jpg_size": os.path.getsize(src), "encode_rc": rc}
            if rc == 0:
                data = open(lep, "rb").read()
                shutil.copyfile(lep, os.path.join(HERE, stem + ".lep"))
                entry.update(lep_md5=md5(data), lep_size=len(data))
                lf = lepfmt.parse_container(data)
                entry.update(nseg=lf.nseg, flag=chr(lf.flag), progressive=lf.frame.progressive,
                             splits=[h.luma_y_start for h in lf.handoffs], sections=sorted(lf.sections))
                rc2 = subprocess.run([LEPTON, "-ujg", "-skipverify", "-allowprogressive", src, ujg],
                                     capture_output=True).returncode
                if rc2 == 0:
                    _, planes = lepfmt.parse_ujg_planes(open(ujg, "rb").read())
                    entry["plane_sha256"] = [hashlib.sha256(p.tobytes()).hexdigest() for p in planes]
                    entry["plane_blocks"] = [int(p.shape[0]) for p in planes]
            manifest[name] = entry
            print(name, entry.get("nseg"), entry.get("lep_size"), rc)
        for stem, srcname, flags in VARIANTS:
            src = os.path.join(REF_IMAGES, srcname)
            lep = os.path.join(td, stem + ".lep")
            rc = subprocess.run([LEPTON, "-skipverify", "-allowprogressive"] + flags + [src, lep],
                                capture_output=True).returncode
            assert rc == 0, (stem, rc)
            data = open(lep, "rb").read()
            shutil.copyfile(lep, os.path.join(HERE, stem + ".lep"))
            lf = lepfmt.parse_container(data)
            manifest[stem + ".lep"] = {"source": srcname, "flags": flags, "lep_md5": md5(data), "lep_size": len(data),
                                       "nseg": lf.nseg, "splits": [h.luma_y_start for h in lf.handoffs]}
            print(stem, lf.nseg, len(data))
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
