"""CPU emulation of the thread-per-segment decode kernels (test infrastructure).

`lep_decode_thread.cu` and `lep_decode_lockstep.cu` give every lane its own segment and exchange nothing between lanes
but votes, so their source compiles as host C++ through `cuda_shim.h` and runs one lane at a time (emu_decode.cc).  That
pins the per-lane arithmetic -- bool decoder, token state machine, predictors, IDCT, block stores -- to the oracle without
a GPU; the GPU parity tests then only have to confirm the same source under real warps.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "lepton_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libemu_decode.so")
SOURCES = [os.path.join(HERE, "emu_decode.cc"), os.path.join(HERE, "cuda_shim.h"), os.path.join(HERE, "fake", "cuda_runtime.h"),
           os.path.join(CSRC, "lep_decode_thread.cu"), os.path.join(CSRC, "lep_decode_lockstep.cu"),
           os.path.join(CSRC, "lep_common.cuh"), os.path.join(CSRC, "lep_predict.cuh"), os.path.join(ROOT, "include", "lepton_b200.h")]

KERNEL_THREAD = 1
KERNEL_LOCKSTEP = 2

_LIB = None


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in SOURCES):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(HERE, "fake"),
                           "-Wno-unknown-pragmas", "-o", OUT, os.path.join(HERE, "emu_decode.cc")])
    return OUT


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.emu_decode_images.restype = ctypes.c_int
    return _LIB


def decode_images(kernel, images, streams):
    """Same contract as LeptonB200Codec.decode_images: decodes into images[i].planes, returns (status, ndecisions) per segment."""
    from lepton_b200.codec import _Image, _Stream
    n = sum(im.nseg for im in images)
    arr = (_Stream * n)()
    keep, k = [], 0
    for im, segs in zip(images, streams):
        assert len(segs) == im.nseg
        for s in segs:
            buf = np.frombuffer(bytes(s), dtype=np.uint8)
            keep.append(buf)
            arr[k].data = buf.ctypes.data if len(buf) else None
            arr[k].len = len(buf)
            k += 1
    cim = (_Image * len(images))(*[im.to_c() for im in images])
    st = (ctypes.c_int32 * n)()
    nd = (ctypes.c_uint64 * n)()
    rc = lib().emu_decode_images(int(kernel), cim, len(images), arr, st, nd)
    if rc != 0:
        raise RuntimeError("emu_decode_images failed with %d" % rc)
    return list(st), list(nd)
