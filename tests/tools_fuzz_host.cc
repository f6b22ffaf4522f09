// Sanitizer fuzz of the host-side parsers (diagnostic, no GPU): mutated .lep / .jpg fixtures through read_lep + the
// baseline re-encoder and through the JPEG front end.  Build and run:
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -Ilepton_b200/csrc \
//       tests/tools_fuzz_host.cc lepton_b200/csrc/lep_recode.cc lepton_b200/csrc/lep_jpeg.cc lepton_b200/csrc/lep_container.cc \
//       -lz -lpthread -ldl -o /tmp/fuzz_host && /tmp/fuzz_host 400 tests/golden/*.lep tests/golden/legacy/*.lep tests/golden/future/*.lep tests/golden/*.jpg
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <random>
#include <vector>

#include "lep_host.h"

using namespace lephost;

static std::vector<uint8_t> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static void one_lep(const std::vector<uint8_t>& d, long& ok, long& refused) {
    LepFile lf;
    { LepFile lz; read_lep(d.data(), d.size(), lz, /*lazy=*/true); }      // the batch decoder's mode: packets recorded in place
    if (!read_lep(d.data(), d.size(), lf)) { ++refused; return; }
    ++ok;
    // re-create the JPEG from all-zero planes: exercises the handoff / restart / truncation bookkeeping
    std::vector<std::vector<int16_t>> planes(4);
    const int16_t* pp[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int c = 0; c < lf.j.ncmp; ++c) {
        planes[c].assign((size_t)lf.j.cmp[c].bc * 64 + 64, 0);
        pp[c] = planes[c].data();
    }
    std::vector<uint8_t> out;
    std::string err;
    if (lf.j.jpegtype == 1) recode_baseline(lf, pp, out, err);
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: fuzz_host ITER files...\n"); return 2; }
    const int iters = atoi(argv[1]);
    std::mt19937 rng(12345);
    long ok = 0, refused = 0, jok = 0, jref = 0;
    for (int a = 2; a < argc; ++a) {
        const std::vector<uint8_t> base = slurp(argv[a]);
        if (base.size() < 64) continue;
        const bool is_lep = base[0] == 0xCF && base[1] == 0x84;
        for (int it = 0; it < iters; ++it) {
            std::vector<uint8_t> d = base;
            const int nmut = 1 + (int)(rng() % 4);
            for (int m = 0; m < nmut; ++m) {
                // bias towards the front of the file (headers, tables, legacy segment table) and the very end
                size_t span = (rng() & 1) ? std::min<size_t>(d.size(), 20000) : d.size();
                size_t pos = rng() % span;
                if ((rng() % 8) == 0) pos = d.size() - 1 - (rng() % std::min<size_t>(d.size(), 16));
                d[pos] = (uint8_t)rng();
            }
            if ((rng() % 10) == 0) d.resize(d.size() - (rng() % std::min<size_t>(d.size() - 40, 4096)));
            if (is_lep) {
                // the header blob is zlib-coded: mutating it mostly fails the inflate; mutate the fixed header and
                // the payload start (legacy segment table) more often
                if ((rng() % 3) == 0) { const uint32_t zl = d[24] | (d[25] << 8) | (d[26] << 16) | ((uint32_t)d[27] << 24);
                    size_t q = 28 + (size_t)zl + 3; if (q + 8 < d.size()) d[q + (rng() % 8)] = (uint8_t)rng(); }
                one_lep(d, ok, refused);
            } else {
                Jpeg j;
                if (!parse_jpeg(d.data(), d.size(), j)) { ++jref; continue; }
                std::vector<std::vector<int16_t>> planes(4);
                int16_t* pp[4] = {nullptr, nullptr, nullptr, nullptr};
                for (int c = 0; c < j.ncmp; ++c) { planes[c].assign((size_t)j.cmp[c].bc * 64 + 64, 0); pp[c] = planes[c].data(); }
                if (decode_scans(j, pp)) ++jok; else ++jref;
            }
        }
    }
    printf("lep: %ld read, %ld refused; jpeg: %ld parsed, %ld refused\n", ok, refused, jok, jref);
    return 0;
}
