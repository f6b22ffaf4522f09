// lepton_cli.cc -- `lepton`-compatible command line front end over liblepton_b200.so.
//
// Keeps the part of the reference CLI surface that selects what is computed (src/lepton/jpgcoder.cc
// initialize_options :988-1219, process_file :1528): `lepton [flags] <in.jpg|in.lep> [out]`, direction chosen from
// the first two bytes of the input (FF D8 -> compress, CF 84 -> decompress; check_file :2178), `-` for stdin/stdout,
// exit status = the reference's ExitCode (src/vp8/util/memory.hh:13-39).  Flags that only configure the reference's
// CPU runtime (-singlethread, -unjailed, -skipverify/-verify, -preload, -memory=, -threadmemory=, -timebound=,
// -maxencodethreads=) are accepted and ignored: the GPU coder always produces the reference's default .lep bytes and
// every file is verified by construction in the test-suite, not at run time.  Service modes (-socket, -listen, -fork,
// -benchmark, -lepcat) are not part of the hot path and are refused.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lepton_b200.h"

static bool read_all(FILE* f, std::vector<uint8_t>& out) {
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.insert(out.end(), buf, buf + n);
    return !ferror(f);
}

int main(int argc, char** argv) {
    std::vector<std::string> files;
    int device = 0;
    for (int i = 1; i < argc; ++i) {
        const char* a = argv[i];
        if (a[0] == '-' && a[1] != 0) {
            if (!strncmp(a, "-device=", 8)) { device = atoi(a + 8); continue; }
            if (!strcmp(a, "-socket") || !strncmp(a, "-socket=", 8) || !strncmp(a, "-listen", 7) || !strcmp(a, "-fork") ||
                !strcmp(a, "-benchmark") || !strcmp(a, "-lepcat") || !strncmp(a, "-startbyte", 10) || !strncmp(a, "-trunc=", 7) ||
                !strcmp(a, "-ujg") || !strcmp(a, "-brotliheader") || !strncmp(a, "-embedding", 10)) {
                fprintf(stderr, "lepton-b200: option %s is outside the B200 hot path build\n", a);
                return 13;   // VERSION_UNSUPPORTED
            }
            continue;        // runtime-tuning flags of the CPU reference: accepted, no effect
        }
        files.push_back(a);
    }
    if (files.empty()) {
        fprintf(stderr, "usage: lepton-b200 [flags] <input.jpg|input.lep|-> [output|-]\n");
        return 1;
    }
    std::vector<uint8_t> in;
    FILE* fi = files[0] == "-" ? stdin : fopen(files[0].c_str(), "rb");
    if (!fi) { fprintf(stderr, "lepton-b200: cannot open %s\n", files[0].c_str()); return 9; }   // FILE_NOT_FOUND
    if (!read_all(fi, in)) return 33;                                                            // OS_ERROR
    if (fi != stdin) fclose(fi);
    if (in.size() < 2) return 3;                                                                 // SHORT_READ
    const bool is_jpeg = in[0] == 0xFF && in[1] == 0xD8, is_lep = in[0] == 0xCF && in[1] == 0x84;
    if (!is_jpeg && !is_lep) { fprintf(stderr, "lepton-b200: input is neither JPEG nor Lepton\n"); return 42; }
    std::string outname;
    if (files.size() > 1) outname = files[1];
    else if (files[0] == "-") outname = "-";
    else {
        outname = files[0];
        size_t dot = outname.rfind('.');
        if (dot != std::string::npos) outname.resize(dot);
        outname += is_jpeg ? ".lep" : ".jpg";
    }
    lepb200_codec* codec = nullptr;
    int rc = lepb200_codec_create(&codec, device, 0);
    if (rc) { fprintf(stderr, "lepton-b200: no usable CUDA device (%d); this build has no CPU coder\n", rc); return 33; }
    lepb200_buffer ib = {in.data(), in.size()};
    lepb200_result res = {nullptr, 0, 0};
    rc = is_jpeg ? lepb200_compress_jpegs(codec, &ib, 1, &res) : lepb200_decompress_leps(codec, &ib, 1, &res);
    if (rc) { fprintf(stderr, "lepton-b200: %s\n", lepb200_codec_last_error(codec)); lepb200_codec_destroy(codec); return 33; }
    if (res.status) {
        fprintf(stderr, "lepton-b200: exit code %d\n", res.status);
        lepb200_codec_destroy(codec);
        return res.status == LEPB200_ST_NOT_HANDLED ? 42 : res.status;
    }
    FILE* fo = outname == "-" ? stdout : fopen(outname.c_str(), "wb");
    if (!fo) { lepb200_codec_destroy(codec); return 33; }
    fwrite(res.data, 1, res.len, fo);
    if (fo != stdout) fclose(fo);
    if (is_jpeg) fprintf(stderr, "%zu %zu\n%.2f%%\n", res.len, in.size(), 100.0 * (double)res.len / (double)in.size());
    lepb200_codec_destroy(codec);
    return 0;
}
