"""Memory / undefined-behaviour check of the kernel sources on the CPU (diagnostic): the warp emulator build of the
encode and decode kernels compiled with AddressSanitizer + UBSan (out-of-bounds shared / global / local accesses,
misaligned vector loads, ...), driven through random geometries and truncated fixtures.  Run as

    g++ -O1 -g -std=c++17 -shared -fPIC -x c++ -I tests/emu/fake -fsanitize=address,undefined \
        -fno-sanitize=shift,signed-integer-overflow -fno-sanitize-recover=undefined -Wno-unknown-pragmas \
        -o /tmp/libemu_asan.so tests/emu/emu_kernels.cc
    LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
        ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python tests/tools_emu_sanitize.py

(shift / signed-overflow checks are off: the reference IDCT is defined on wrapping 32-bit arithmetic.)

The whole emulator suite runs on the sanitizer build too (tests/emu/emu.py honours LEPB200_EMU_LIB): every kernel --
kernel A, the three range coder forms, both decode kernels, the Huffman decode / sub-sequence / encode kernels, the
gather kernel -- with out-of-bounds accesses to global, shared and local memory trapped:

    LD_PRELOAD=... ASAN_OPTIONS=... LEPB200_EMU_LIB=/tmp/libemu_asan.so python -m pytest tests/test_emu_*.py -q -p no:cacheprovider

(round 2, final tree: all 141 emulator tests passed on the sanitizer build, no report; host parsers: tests/tools_fuzz_host.cc, 14 000 mutated inputs clean)."""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
for d in (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle"), HERE, os.path.join(HERE, "emu")):
    sys.path.insert(0, d)
import numpy as np
import emu
emu._LIB = ctypes.CDLL(os.environ.get("LEPB200_EMU_LIB", "/tmp/libemu_asan.so"))
emu._LIB.emu_decode_images.restype = ctypes.c_int
emu._LIB.emu_encode_images.restype = ctypes.c_int
from helpers import *
rng = np.random.default_rng(7)
cfgs=[dict(ncmp=3, mcuh=7, mcuv=6, sf=((2, 2), (1, 1), (1, 1)), nseg=3),
      dict(ncmp=1, mcuh=1, mcuv=9, sf=((1, 1),), nseg=2),
      dict(ncmp=3, mcuh=12, mcuv=8, sf=((2, 2), (1, 1), (1, 1)), nseg=8, density=0.9, amp=100, qscale=0.3),
      dict(ncmp=3, mcuh=6, mcuv=4, sf=((2, 1), (1, 1), (1, 1)), nseg=2)]
from lepton_b200 import CoefImage
for cfg in cfgs:
    img = random_coef_image(rng, **cfg)
    ref = oracle_encode_image(img)
    got = emu.encode_images([img])[0]
    assert [g[1] for g in got]==[r[1] for r in ref]
    for k in (0, emu.KERNEL_G2(4), emu.KERNEL_G2(8)):
        out = CoefImage(ncmp=img.ncmp, mcuv=img.mcuv, bch=img.bch, bcv=img.bcv, qtables_zigzag=img.qtables_zigzag,
                    planes=[np.full_like(p, -5) for p in img.planes], luma_y_start=img.luma_y_start)
        st,_ = emu.decode_images(k,[out],[[g[1] for g in got]])
        assert all(s==0 for s in st) and all(np.array_equal(a,b) for a,b in zip(out.planes,img.planes))
    print('ok', cfg['mcuh'])
for name in ['androidcrop_t2.lep','truncatedzerorun.lep','singlerowtrunc.lep','colorswap.lep']:
    lf = load_lep(name); planes, streams = oracle_decode_planes(lf)
    img = coef_image_from_lep(lf, planes)
    got = emu.encode_images([img])[0]
    assert [g[1] for g in got]==list(streams[:lf.nseg])
    for k in (0, emu.KERNEL_G2(4), emu.KERNEL_G2(8)):
        out = coef_image_from_lep(lf, [np.full_like(p, 9) for p in planes])
        st,_=emu.decode_images(k,[out],[streams[:lf.nseg]])
        assert all(s==0 for s in st) and all(np.array_equal(a,b) for a,b in zip(out.planes,planes))
    print('ok', name)
