"""Cache footprint of the adaptive model (diagnostic, CPU only; needs PIL for bench.synth_jpeg).

Builds the warp-emulator harness with LEPB200_EMU_TRACE, decodes one bench image (1920x1080 4:2:0 q85, 4 segments) with
the lock-step kernel and replays every segment's model accesses through an LRU of N 128-byte lines: how much L2 a
segment needs for its 1.58 MB model to behave as if resident.  Quoted in DESIGN.md section 4.

    python tests/tools_model_reuse.py
"""
import ctypes
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for d in (ROOT, os.path.join(ROOT, "oracle"), HERE, os.path.join(HERE, "emu")):
    sys.path.insert(0, d)

TRACE_CC = r'''
#include <list>
#include <map>
#include <unordered_map>
#include <vector>
#include <cstdint>
#include <cstdio>
static std::map<const void*, std::vector<uint32_t>> g_trace;
static inline void emu_trace_model_access(const uint16_t* model, uint32_t addr) { g_trace[model].push_back(addr >> 6); }   // 128-byte lines
#define LEPB200_EMU_TRACE 1
#include "emu_kernels.cc"
extern "C" void trace_report() {
    const size_t caps[] = {32, 64, 128, 256, 512, 1024};
    double hits[6] = {0, 0, 0, 0, 0, 0}, total = 0;
    size_t distinct = 0;
    int lanes = 0;
    for (auto& kv : g_trace) {
        const auto& tr = kv.second;
        if (tr.empty()) continue;
        ++lanes;
        std::list<uint32_t> lru;
        std::unordered_map<uint32_t, std::list<uint32_t>::iterator> pos;
        for (uint32_t line : tr) {
            auto it = pos.find(line);
            size_t dist = SIZE_MAX;
            if (it != pos.end()) { dist = 0; for (auto j = lru.begin(); j != it->second; ++j) ++dist; lru.erase(it->second); }
            lru.push_front(line);
            pos[line] = lru.begin();
            for (int c = 0; c < 6; ++c) if (dist < caps[c]) hits[c] += 1;
            total += 1;
        }
        distinct += pos.size();
    }
    printf("segments %d, model accesses %.0f, distinct 128-byte lines per segment %.0f (%.1f KB of a %.0f KB model)\n", lanes, total,
           (double)distinct / lanes, distinct / (double)lanes * 128 / 1024, lepb200::MODEL_BYTES / 1024.0);
    for (int c = 0; c < 6; ++c) printf("  LRU of %4zu lines (%5.1f KB per segment): hit rate %.4f\n", caps[c], caps[c] * 128 / 1024.0, hits[c] / total);
    g_trace.clear();
}
'''


def main():
    import numpy as np
    import bench
    import emu
    from lepton_b200 import CoefImage, HostJpeg
    with tempfile.TemporaryDirectory() as td:
        cc = os.path.join(td, "trace.cc")
        open(cc, "w").write(TRACE_CC)
        so = os.path.join(td, "libemu_trace.so")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(HERE, "emu", "fake"),
                               "-I", os.path.join(HERE, "emu"), "-Wno-unknown-pragmas", "-o", so, cc])
        emu._LIB = ctypes.CDLL(so)
        emu._LIB.emu_decode_images.restype = ctypes.c_int
        emu._LIB.emu_encode_images.restype = ctypes.c_int
        img = HostJpeg(bench.synth_jpeg(0)).coef_image()
        enc = emu.encode_images([img])[0]
        out = CoefImage(ncmp=img.ncmp, mcuv=img.mcuv, bch=img.bch, bcv=img.bcv, qtables_zigzag=img.qtables_zigzag,
                        planes=[np.zeros_like(np.asarray(p)) for p in img.planes], luma_y_start=img.luma_y_start)
        st, nd = emu.decode_images(emu.KERNEL_LOCKSTEP, [out], [[e[1] for e in enc]])
        assert all(s == 0 for s in st) and all(np.array_equal(a, np.asarray(b)) for a, b in zip(out.planes, img.planes))
        sys.stdout.flush()
        emu._LIB.trace_report()


if __name__ == "__main__":
    main()
