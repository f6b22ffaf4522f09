// lep_recode.cc -- decode-side host halves: .lep container reader (read_ujpg, jpgcoder.cc:4117-4362; fixed header
// :2140-2176; MuxReader, src/io/MuxReader.hh:230-283; ThreadHandoff::deserialize, thread_handoff.cc:4-39) and the
// baseline JPEG re-creation from coefficient planes (recode_baseline_jpeg, src/lepton/recoder.cc:694-889;
// recode_one_mcu_row :316-410; encode_block_seq :245-313; 0xff stuffing :144-185).
//
// The reference re-encodes every thread-segment independently from its handoff (overhang bits + last DCs) and
// concatenates; encoding the scan front to back from the complete planes yields the same bytes, so that is what
// is done here (one host thread per file; files are processed in parallel by the caller).
#include <zlib.h>
#include <dlfcn.h>

#include <emmintrin.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "lep_host.h"

namespace lephost {

namespace {
inline uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }
bool lfail(LepFile& lf, int st, const char* msg) { lf.status = st; lf.error = msg; return false; }

// Header blobs of container versions 2 and 4 are brotli streams (write_ujpg, jpgcoder.cc:4032-4041; read_ujpg :4168-4175;
// the reference vendors the brotli sources).  Brotli needs the 122 KB static dictionary of RFC 7932, which is third-party
// DATA, so no decoder is written here: the system's libbrotlidec (the same kind of dependency as -lz for version 1) is
// loaded at run time through its stable streaming C API.  Without the library such files are refused (status 200) as before.
struct BrotliApi {
    void* (*create)(void*, void*, void*) = nullptr;
    int (*stream)(void*, size_t*, const uint8_t**, size_t*, uint8_t**, size_t*) = nullptr;
    void (*destroy)(void*) = nullptr;
    bool ok = false;
};
const BrotliApi& brotli_api() {
    static BrotliApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : {"libbrotlidec.so.1", "libbrotlidec.so"}) { h = dlopen(name, RTLD_NOW | RTLD_LOCAL); if (h) break; }
        if (!h) return;
        api.create = reinterpret_cast<void* (*)(void*, void*, void*)>(dlsym(h, "BrotliDecoderCreateInstance"));
        api.stream = reinterpret_cast<int (*)(void*, size_t*, const uint8_t**, size_t*, uint8_t**, size_t*)>(dlsym(h, "BrotliDecoderDecompressStream"));
        api.destroy = reinterpret_cast<void (*)(void*)>(dlsym(h, "BrotliDecoderDestroyInstance"));
        api.ok = api.create && api.stream && api.destroy;
    });
    return api;
}
// 0 ok, 1 no library, 2 corrupt stream, 3 larger than `cap`
int brotli_decompress(const uint8_t* in, size_t n, std::vector<uint8_t>& out, size_t cap) {
    const BrotliApi& b = brotli_api();
    if (!b.ok) return 1;
    void* st = b.create(nullptr, nullptr, nullptr);
    if (!st) return 1;
    out.resize(std::max<size_t>(4096, n * 6));
    size_t avail_in = n, have = 0;
    const uint8_t* next_in = in;
    int rc = 2;
    for (;;) {
        size_t avail_out = out.size() - have;
        uint8_t* next_out = out.data() + have;
        const int r = b.stream(st, &avail_in, &next_in, &avail_out, &next_out, nullptr);
        have = out.size() - avail_out;
        if (r == 1) { rc = 0; break; }                     // BROTLI_DECODER_RESULT_SUCCESS
        if (r != 3) { rc = 2; break; }                     // error, or more input wanted than the blob has
        if (out.size() >= cap) { rc = 3; break; }          // BROTLI_DECODER_RESULT_NEEDS_MORE_OUTPUT
        out.resize(std::min(cap, out.size() * 2));
    }
    b.destroy(st);
    out.resize(rc == 0 ? have : 0);
    return rc;
}

// AlignedBlock index of each zig-zag position (src/vp8/util/aligned_block.hh:56-65)
const uint8_t k_zigzag_to_aligned[64] = {
    49, 50, 57, 58, 0, 51, 52, 1, 2, 59, 60, 3, 4, 5, 53, 54, 6, 7, 8, 9, 61, 62, 10, 11,
    12, 13, 14, 55, 56, 15, 16, 17, 18, 19, 20, 63, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32,
    33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48};
}  // namespace

bool brotli_available() { return brotli_api().ok; }

bool read_lep(const uint8_t* d, size_t n, LepFile& lf, bool lazy) {
    if (n < 28 + 3 + 4 || d[0] != 0xCF || d[1] != 0x84) return lfail(lf, VERSION_UNSUPPORTED, "not a .lep file");
    lf.version = d[2]; lf.flag = d[3]; lf.nseg = d[4];
    // version 1: zlib header blob; 2 and 4: brotli header blob (and an EOF marker behind the mux packets); 3: brotli header AND the
    // ANS coder instead of the bool coder (makeDecoder(..., ujgversion == 3), jpgcoder.cc:1727) -- another codec, refused
    if (lf.version != 1 && lf.version != 2 && lf.version != 4) return lfail(lf, NOT_HANDLED, "container version not handled (3 = ANS coder)");
    if (lf.flag == 'Y') return lfail(lf, NOT_HANDLED, "-startbyte slices are not handled");
    lf.jpeg_size = rd32(d + 20);
    const uint32_t zlen = rd32(d + 24);
    if ((size_t)28 + zlen + 3 + 4 > n) return lfail(lf, SHORT_READ, "truncated .lep");
    // inflate the header blob
    std::vector<uint8_t> blob;
    if (lf.version != 1) {
        const int rc = brotli_decompress(d + 28, zlen, blob, size_t(256) << 20);
        if (rc == 1) return lfail(lf, NOT_HANDLED, "brotli header blob (container version 2 / 4) and no libbrotlidec on this system");
        if (rc == 3) return lfail(lf, 38 /*TOO_MUCH_MEMORY_NEEDED*/, "header blob too large");
        if (rc) return lfail(lf, ASSERTION_FAILURE, "Data not properly brotli coded");
    } else {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit(&zs) != Z_OK) return lfail(lf, ASSERTION_FAILURE, "inflateInit failed");
        zs.next_in = const_cast<uint8_t*>(d + 28); zs.avail_in = zlen;
        blob.resize(std::max<size_t>(4096, (size_t)zlen * 4));
        size_t have = 0;
        int ret;
        do {
            if (have == blob.size()) {
                // untrusted input: a few KB of zlib can expand to gigabytes.  The blob holds the JPEG header, garbage and
                // per-restart bookkeeping -- all bounded by the JPEG it describes (the reference bounds it through its
                // memory limit); 256 MB is far beyond any file the 4-byte size fields can describe sensibly
                if (blob.size() >= (size_t(256) << 20)) { inflateEnd(&zs); return lfail(lf, 38 /*TOO_MUCH_MEMORY_NEEDED, memory.hh:34*/, "header blob too large"); }
                blob.resize(blob.size() * 2);
            }
            zs.next_out = blob.data() + have; zs.avail_out = (uInt)(blob.size() - have);
            ret = inflate(&zs, Z_NO_FLUSH);
            have = blob.size() - zs.avail_out;
        } while (ret == Z_OK);
        inflateEnd(&zs);
        if (ret != Z_STREAM_END) return lfail(lf, ASSERTION_FAILURE, "Data not properly zlib coded");
        blob.resize(have);
    }
    size_t p = 0;
    auto need = [&](size_t k) { return p + k <= blob.size(); };
    if (!need(7) || memcmp(&blob[p], "HDR", 3)) return lfail(lf, UNSUPPORTED_JPEG, "HDR marker not found");
    const uint32_t hdrs = rd32(&blob[p + 3]);
    p += 7;
    if (!need(hdrs)) return lfail(lf, SHORT_READ, "short HDR");
    Jpeg& j = lf.j;
    j.hdr.assign(blob.begin() + p, blob.begin() + p + hdrs);
    p += hdrs;
    if (!parse_frame(j)) { lf.status = j.status; lf.error = j.error; return false; }
    if (!need(4)) return lfail(lf, SHORT_READ, "short pad section");
    if (!memcmp(&blob[p], "P0D", 3)) {
        j.padbit = (int8_t)blob[p + 3];
    } else if (!memcmp(&blob[p], "PAD", 3)) {
        // legacy: one pad bit that stands for all of them (jpgcoder.cc:4228-4242)
        const int8_t pb = (int8_t)blob[p + 3];
        if (!(pb == 0 || pb == 1 || pb == -1)) return lfail(lf, STREAM_INCONSISTENT, "Legacy Padbit must be 0, 1 or -1");
        j.padbit = pb == 1 ? 0x7f : pb;
    } else {
        return lfail(lf, UNSUPPORTED_JPEG, "PAD marker not found");
    }
    p += 4;
    j.grb.clear();
    bool have_grb = false;
    while (need(3)) {
        const uint8_t* m = &blob[p];
        if (!memcmp(m, "CRS", 3)) {
            if (!need(7)) return lfail(lf, SHORT_READ, "short CRS");
            uint32_t k = rd32(m + 3);
            if (!need(7 + 4 * (size_t)k)) return lfail(lf, SHORT_READ, "short CRS");
            j.rst_cnt.resize(k);
            for (uint32_t i = 0; i < k; ++i) j.rst_cnt[i] = rd32(m + 7 + 4 * i);
            lf.rst_cnt_set = true;
            p += 7 + 4 * (size_t)k;
        } else if (m[0] == 'H' && m[1] == 'H') {
            const int k = m[2];
            if (!need(3 + 16 * (size_t)k)) return lfail(lf, VERSION_UNSUPPORTED, "short handoff table");
            for (int i = 0; i < k; ++i) {
                const uint8_t* r = m + 3 + 16 * i;
                Handoff h;
                h.luma_y_start = (uint16_t)(r[0] | (r[1] << 8));
                h.segment_size = rd32(r + 2);
                h.overhang_byte = r[6]; h.num_overhang_bits = r[7];
                for (int q = 0; q < 4; ++q) h.last_dc[q] = (int16_t)(r[8 + 2 * q] | (r[9 + 2 * q] << 8));
                lf.handoffs.push_back(h);
            }
            for (size_t i = 1; i < lf.handoffs.size(); ++i) lf.handoffs[i - 1].luma_y_end = lf.handoffs[i].luma_y_start;
            p += 3 + 16 * (size_t)k;
        } else if (!memcmp(m, "FRS", 3)) {
            if (!need(7)) return lfail(lf, SHORT_READ, "short FRS");
            uint32_t k = rd32(m + 3);
            if (!need(7 + (size_t)k)) return lfail(lf, SHORT_READ, "short FRS");
            j.rst_err.assign(m + 7, m + 7 + k);
            p += 7 + (size_t)k;
        } else if (!memcmp(m, "GRB", 3)) {
            if (!need(7)) return lfail(lf, SHORT_READ, "short GRB");
            uint32_t k = rd32(m + 3);
            if (!need(7 + (size_t)k)) return lfail(lf, SHORT_READ, "short GRB");
            j.grb.assign(m + 7, m + 7 + k);
            have_grb = true;
            p += 7 + (size_t)k;
        } else if (!memcmp(m, "EEE", 3)) {
            if (!need(31)) return lfail(lf, SHORT_READ, "short EEE");
            lf.has_eee = true;
            for (int i = 0; i < 7; ++i) lf.eee[i] = rd32(m + 3 + 4 * i);
            // UncompressedComponents::set_truncation_bounds (uncompressed_components.hh:166-188)
            j.early_eof = true;
            j.max_cmp = (int)lf.eee[0]; j.max_bpos = (int)lf.eee[1]; j.max_sah = (int)lf.eee[2];
            for (int c = 0; c < j.ncmp; ++c) {
                const Component& k = j.cmp[c];
                j.max_dpos[c] = (int)lf.eee[3 + c];
                const long tbc = (long)lf.eee[3 + c] + 1;
                if (tbc > k.bc) return lfail(lf, STREAM_INCONSISTENT, "truncation bound beyond the component");
                int vs = (int)std::min<long>(tbc / k.bch + (tbc % k.bch ? 1 : 0), k.bcv);
                const int ratio = k.bcv / j.mcuv;
                while (vs % ratio != 0 && vs + 1 <= k.bcv) ++vs;
                j.trunc_bcv[c] = vs;
                j.trunc_bc[c] = (int)tbc;
            }
            p += 31;
        } else if (!memcmp(m, "PGR", 3) || !memcmp(m, "PGE", 3) || !memcmp(m, "SIZ", 3)) {
            return lfail(lf, NOT_HANDLED, "prefix garbage / embedded JPEG sections are not handled");
        } else {
            return lfail(lf, UNSUPPORTED_JPEG, "unknown data found in header blob");
        }
    }
    if (!have_grb) { j.grb = {0xFF, 0xD9}; }          // "if we don't have any garbage, assume FFD9 EOI" (jpgcoder.cc:4194)
    size_t q = 28 + (size_t)zlen;
    if (memcmp(d + q, "CMP", 3)) return lfail(lf, UNSUPPORTED_JPEG, "CMP marker missing");
    q += 3;
    const size_t end = n - 4;
    if (lf.handoffs.empty()) {
        // Legacy files (before the handoff table existed): the payload opens with the number of thread-segments and
        // the luma rows where they end (VP8ComponentDecoder::initialize_baseline_decoder, vp8_decoder.cc:337-369).
        // Their handoffs carry no Huffman state (ThreadHandoff::LEGACY_OVERHANG_BITS), so the scan can only be
        // re-created front to back -- which is what the host re-encoder does anyway.
        if (q + 1 > end) return lfail(lf, SHORT_READ, "legacy segment table missing");
        const int k = d[q];
        if (k == 0) return lfail(lf, THREADING_PARTIAL_MCU, "legacy segment count is zero");
        if (q + 1 + 2 * (size_t)(k - 1) > end) return lfail(lf, SHORT_READ, "short legacy segment table");
        const int luma_mul = j.cmp[0].bcv / j.mcuv;
        for (int i = 0; i < k; ++i) {
            Handoff h;
            h.num_overhang_bits = 0xff;
            h.luma_y_end = i + 1 < k ? (uint16_t)(d[q + 1 + 2 * i] | (d[q + 2 + 2 * i] << 8)) : (uint16_t)j.cmp[0].bcv;
            if (i + 1 < k && h.luma_y_end % luma_mul) return lfail(lf, THREADING_PARTIAL_MCU, "legacy split inside an MCU row");
            h.luma_y_start = i ? lf.handoffs[i - 1].luma_y_end : 0;
            lf.handoffs.push_back(h);
        }
        q += 1 + 2 * (size_t)(k - 1);
        lf.nseg = k;
        lf.legacy = true;
    }
    if ((int)lf.handoffs.size() != lf.nseg) return lfail(lf, VERSION_UNSUPPORTED, "handoff table inconsistent with the thread count");
    if (lf.nseg > 16) return lfail(lf, NOT_HANDLED, "more than 16 thread-segments");             // MAX_NUM_THREADS of this build
    // a corrupt handoff table must fail this file only, not the batch it travels in
    if (lf.handoffs[0].luma_y_start != 0) return lfail(lf, STREAM_INCONSISTENT, "first thread-segment does not start at row 0");
    for (int i = 0; i < lf.nseg; ++i) {
        if ((int)lf.handoffs[i].luma_y_start > j.cmp[0].bcv || (i + 1 < lf.nseg && lf.handoffs[i].luma_y_start > lf.handoffs[i + 1].luma_y_start))
            return lfail(lf, STREAM_INCONSISTENT, "thread-segment rows out of order or beyond the image");
    }
    // demux (src/io/MuxReader.hh:230-283); the last 4 bytes are the file-size trailer
    if (lazy) { lf.spans.assign(16, {}); lf.stream_len.assign(16, 0); }
    else lf.streams.assign(16, std::vector<uint8_t>());
    while (q + 3 <= end) {
        if (d[q] == 0xFF && d[q + 1] == 0xFE && d[q + 2] == 0xFF) break;      // MuxReader::getEofMarker (MuxReader.hh:131-139,240-243), written by versions > 1
        const uint8_t hd = d[q];
        const int sid = hd & 15, flags = (hd >> 4) & 3;
        size_t len, skip;
        if (flags == 0) { len = (size_t)d[q + 1] + 256 * (size_t)d[q + 2] + 1; skip = 3; }
        else { len = (size_t)1024 << (2 * flags); skip = 1; }
        if (q + skip + len > end) return lfail(lf, SHORT_READ, "mux packet runs past the end of the file");
        if (lazy) { lf.spans[sid].emplace_back(d + q + skip, (uint32_t)len); lf.stream_len[sid] += len; }
        else lf.streams[sid].insert(lf.streams[sid].end(), d + q + skip, d + q + skip + len);
        q += skip + len;
    }
    if (lazy) { lf.spans.resize(lf.nseg); lf.stream_len.resize(lf.nseg); }
    else lf.streams.resize(lf.nseg);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Huffman re-encoding of the (single, interleaved or single-component) baseline scan
// ------------------------------------------------------------------------------------------------
namespace {

inline int bitlen16(int v) { return v ? 32 - __builtin_clz((unsigned)v) : 0; }

// Bit writer of the baseline re-encoder (the hot loop of the .lep -> JPEG direction on the host): 64-bit accumulator,
// four bytes leave at a time unless one of them is 0xFF and needs its stuffed zero (recoder.cc:144-185).
struct FastWriter {
    std::vector<uint8_t>& out;
    uint8_t* p;
    uint8_t* lim;
    uint64_t acc = 0;
    int nbits = 0;
    explicit FastWriter(std::vector<uint8_t>& o, size_t expect) : out(o) {
        const size_t used = out.size();
        out.resize(used + expect + 4096);
        p = out.data() + used; lim = out.data() + out.size();
    }
    inline void room(size_t n) {
        if ((size_t)(lim - p) >= n) return;
        const size_t used = (size_t)(p - out.data());
        out.resize(out.size() * 2 + n + 4096);
        p = out.data() + used; lim = out.data() + out.size();
    }
    inline void raw(uint8_t b) { room(1); *p++ = b; }
    inline void raw_bytes(const uint8_t* src, size_t n) { room(n); memcpy(p, src, n); p += n; }
    inline void put(uint32_t v, int n) {          // n <= 32, v < 2^n
        if (n > 32) n = 32;                       // tables are validated by HuffTable::build(); never shift by >= 64
        acc = (acc << n) | v;
        nbits += n;
        if (nbits >= 32) {
            const uint32_t w = (uint32_t)(acc >> (nbits - 32));
            nbits -= 32;
            room(8);
            if (((w & 0x7f7f7f7fu) + 0x01010101u) & w & 0x80808080u) {          // some byte is 0xFF (exact test: b == 0xFF <=> (b & 0x7f) + 1 carries into a set top bit)
                for (int sh = 24; sh >= 0; sh -= 8) {
                    const uint8_t b = (uint8_t)(w >> sh);
                    *p++ = b;
                    if (b == 0xFF) *p++ = 0x00;
                }
            } else {
                p[0] = (uint8_t)(w >> 24); p[1] = (uint8_t)(w >> 16); p[2] = (uint8_t)(w >> 8); p[3] = (uint8_t)w;
                p += 4;
            }
        }
    }
    inline void flush_bytes() {
        room(16 + (size_t)(nbits / 8) * 2);
        while (nbits >= 8) {
            const uint8_t b = (uint8_t)(acc >> (nbits - 8));
            *p++ = b;
            if (b == 0xFF) *p++ = 0x00;
            nbits -= 8;
        }
    }
    // abitwriter::pad (bitops.hh:168-175): successive bits of `fill`, LSB first
    inline void pad(uint8_t fill) {
        int offset = 1;
        while (nbits & 7) { put((fill & offset) ? 1 : 0, 1); offset <<= 1; }
        flush_bytes();
    }
    inline void finish() { out.resize((size_t)(p - out.data())); }
};

// zig-zag-ordered bit mask of the non-zero coefficients of an AlignedBlock: SSE2 compare + a fixed bit permutation
struct ZzPerm {
    uint64_t t[8][256];
    ZzPerm() {
        int al2zz[64];
        for (int z = 0; z < 64; ++z) al2zz[k_zigzag_to_aligned[z]] = z;
        for (int byte = 0; byte < 8; ++byte)
            for (int v = 0; v < 256; ++v) {
                uint64_t m = 0;
                for (int b = 0; b < 8; ++b) if (v & (1 << b)) m |= 1ull << al2zz[byte * 8 + b];
                t[byte][v] = m;
            }
    }
};
const ZzPerm g_zzperm;

inline uint64_t nonzero_mask_zigzag(const int16_t* blk) {
    const __m128i zero = _mm_setzero_si128();
    uint64_t zmask = 0;                                   // bit a: coefficient a (aligned order) IS zero
    for (int i = 0; i < 4; ++i) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(blk + 16 * i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(blk + 16 * i + 8));
        const __m128i eq = _mm_packs_epi16(_mm_cmpeq_epi16(a, zero), _mm_cmpeq_epi16(b, zero));
        zmask |= (uint64_t)(uint32_t)_mm_movemask_epi8(eq) << (16 * i);
    }
    const uint64_t nz = ~zmask;
    uint64_t m = 0;
    for (int byte = 0; byte < 8; ++byte) m |= g_zzperm.t[byte][(nz >> (8 * byte)) & 255];
    return m;
}

// bit z set when |coefficient at zig-zag position z| >= thr (thr in 1..32768): the positions a progressive
// scan at successive-approximation bit `sal` sees as non-zero (thr = 1 << sal) or as already significant
// (thr = 2 << sal)
inline uint64_t magnitude_mask_zigzag(const int16_t* blk, int thr) {
    const __m128i zero = _mm_setzero_si128();
    const __m128i t1 = _mm_set1_epi16((short)(thr - 1));
    uint64_t zmask = 0;                                   // bit a: |coefficient a| < thr
    for (int i = 0; i < 4; ++i) {
        __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(blk + 16 * i));
        __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(blk + 16 * i + 8));
        a = _mm_max_epi16(a, _mm_sub_epi16(zero, a));     // |x| as u16 (-32768 stays 0x8000 = 32768)
        b = _mm_max_epi16(b, _mm_sub_epi16(zero, b));
        const __m128i eq = _mm_packs_epi16(_mm_cmpeq_epi16(_mm_subs_epu16(a, t1), zero),
                                           _mm_cmpeq_epi16(_mm_subs_epu16(b, t1), zero));
        zmask |= (uint64_t)(uint32_t)_mm_movemask_epi8(eq) << (16 * i);
    }
    const uint64_t nz = ~zmask;
    uint64_t m = 0;
    for (int byte = 0; byte < 8; ++byte) m |= g_zzperm.t[byte][(nz >> (8 * byte)) & 255];
    return m;
}

}  // namespace

bool recode_scans(const LepFile& lf, const int16_t* const planes[4], std::vector<uint8_t>& out, std::string& err);

bool gpu_recode_setup(const LepFile& lf, GpuRecodeSetup& out) {
    const Jpeg& j = lf.j;
    if (lf.flag != 'Z' || j.jpegtype != 1 || lf.has_eee || j.early_eof || j.ncmp < 1 || j.ncmp > 3) return false;
    const std::vector<uint8_t>& h = j.hdr;
    struct Raw { bool set = false; uint8_t bits[17]; uint8_t vals[256]; } dc[4], ac[4];
    size_t hpos = 0;
    int nsos = 0;
    out.rsti = 0; out.hpos = 0;
    while (hpos + 4 <= h.size()) {
        const uint8_t type = h[hpos + 1];
        const size_t len = 2 + be16(&h[hpos + 2]);
        if (hpos + len > h.size()) return false;
        const uint8_t* seg = &h[hpos];
        if (type == 0xC4) {
            if (nsos) return false;                    // tables after the scan: general path
            size_t p = 4;
            while (p < len) {
                const int tc = seg[p] >> 4, th = seg[p] & 15;
                if (tc >= 2 || th >= 4) return false;
                ++p;
                if (p + 16 > len) return false;
                Raw& t = tc ? ac[th] : dc[th];
                int total = 0;
                t.bits[0] = 0;
                for (int i = 0; i < 16; ++i) { t.bits[i + 1] = seg[p + i]; total += seg[p + i]; }
                if (total > 256 || p + 16 + total > len) return false;
                memset(t.vals, 0, 256);
                memcpy(t.vals, seg + p + 16, total);
                t.set = true;
                p += 16 + total;
            }
        } else if (type == 0xDD) {
            if (nsos) return false;
            if (len < 6) return false;
            out.rsti = be16(seg + 4);
        } else if (type == 0xDA) {
            if (++nsos > 1) return false;
            const int nc = seg[4];
            if (nc != j.ncmp || len < (size_t)(8 + 2 * nc)) return false;
            for (int i = 0; i < nc; ++i) {
                if (seg[5 + 2 * i] != j.cmp[i].jid) return false;       // scan order must be frame order
                const int td = seg[6 + 2 * i] >> 4, ta = seg[6 + 2 * i] & 15;
                if (td >= 4 || ta >= 4 || !dc[td].set || !ac[ta].set) return false;
                memcpy(out.dc_bits[i], dc[td].bits, 17); memcpy(out.dc_vals[i], dc[td].vals, 256);
                memcpy(out.ac_bits[i], ac[ta].bits, 17); memcpy(out.ac_vals[i], ac[ta].vals, 256);
            }
            out.hpos = hpos + len;
        }
        hpos += len;
    }
    if (nsos != 1) return false;
    for (int c = 0; c < j.ncmp; ++c) if (j.cmp[c].H < 1 || j.cmp[c].H > 2 || j.cmp[c].V < 1 || j.cmp[c].V > 2) return false;
    if (j.ncmp == 1 && (j.cmp[0].H != 1 || j.cmp[0].V != 1 || j.cmp[0].bch != j.cmp[0].nch || j.cmp[0].bcv != j.cmp[0].ncv)) return false;   // single-component scans walk nch x ncv blocks in raster order
    // every restart marker of the scan must be wanted (truncated originals limit them, recoder.cc:381-397)
    const unsigned nrst = out.rsti ? (unsigned)(j.mcuc - 1) / (unsigned)out.rsti : 0u;
    if (lf.rst_cnt_set && !j.rst_cnt.empty() && j.rst_cnt[0] < nrst) return false;
    const size_t trailing = j.rst_err.empty() ? 0 : 2 * (size_t)j.rst_err[0];
    const size_t fixed = 2 + h.size() + j.grb.size() + trailing;
    if ((size_t)lf.jpeg_size <= fixed) return false;
    out.scan_bytes = (uint32_t)(lf.jpeg_size - fixed);
    return true;
}

bool assemble_baseline(const LepFile& lf, const GpuRecodeSetup& gs, const uint8_t* scan, std::vector<uint8_t>& out, std::string& err) {
    const Jpeg& j = lf.j;
    const std::vector<uint8_t>& h = j.hdr;
    out.clear();
    out.reserve((size_t)lf.jpeg_size + 16);
    out.push_back(0xFF); out.push_back(0xD8);
    out.insert(out.end(), h.begin(), h.begin() + gs.hpos);
    out.insert(out.end(), scan, scan + gs.scan_bytes);
    if (!j.rst_err.empty()) {
        const unsigned cum = gs.rsti ? (unsigned)(j.mcuh * j.mcuv - 1) / gs.rsti : 0;
        for (unsigned i = 0; i < j.rst_err[0]; ++i) { out.push_back(0xFF); out.push_back((uint8_t)(0xD0 + ((cum + i) & 7))); }
    }
    out.insert(out.end(), h.begin() + gs.hpos, h.end());
    out.insert(out.end(), j.grb.begin(), j.grb.end());
    if (out.size() != lf.jpeg_size) { err = "re-created JPEG has the wrong size"; return false; }
    return true;
}

bool recode_baseline(const LepFile& lf, const int16_t* const planes[4], std::vector<uint8_t>& out, std::string& err) {
    const Jpeg& j = lf.j;
    if (lf.flag != 'Z' || j.jpegtype != 1) return recode_scans(lf, planes, out, err);     // multi-scan / progressive files
    const std::vector<uint8_t>& h = j.hdr;
    HuffTable dc_t[4], ac_t[4];
    int rsti = 0;
    size_t hpos = 0;
    int ncomp = 0, scmp[4] = {0, 0, 0, 0}, td[4] = {0}, ta[4] = {0};
    // handle_initial_segments (recoder.cc:412-461): everything up to and including the first SOS
    bool found = false;
    while (hpos + 4 <= h.size()) {
        const uint8_t type = h[hpos + 1];
        const size_t len = 2 + be16(&h[hpos + 2]);
        if (hpos + len > h.size()) { err = "truncated header segment"; return false; }
        const uint8_t* seg = &h[hpos];
        if (type == 0xC4) {
            size_t p = 4;
            while (p < len) {
                const int tc = seg[p] >> 4, th = seg[p] & 15;
                if (tc >= 2 || th >= 4) break;
                ++p;
                HuffTable& t = tc ? ac_t[th] : dc_t[th];
                int total = 0;
                t.bits[0] = 0;
                for (int i = 0; i < 16; ++i) { t.bits[i + 1] = seg[p + i]; total += seg[p + i]; }
                if (total > 256 || p + 16 + total > len) { err = "bad DHT"; return false; }
                memcpy(t.vals, seg + p + 16, total);
                if (!t.build()) { err = "bad huffman table"; return false; }
                t.set = true;
                p += 16 + total;
            }
        } else if (type == 0xDD) {
            if (len < 6) { err = "bad DRI"; return false; }
            rsti = be16(seg + 4);
        } else if (type == 0xDA) {
            if (len < 5) { err = "bad SOS"; return false; }
            ncomp = seg[4];
            if (ncomp < 1 || ncomp > j.ncmp || len < (size_t)(8 + 2 * ncomp)) { err = "bad SOS"; return false; }
            for (int i = 0; i < ncomp; ++i) {
                int c = 0;
                while (c < j.ncmp && j.cmp[c].jid != seg[5 + 2 * i]) ++c;
                if (c == j.ncmp) { err = "component id mismatch"; return false; }
                scmp[i] = c; td[c] = seg[6 + 2 * i] >> 4; ta[c] = seg[6 + 2 * i] & 15;
                // the .lep header is untrusted input: table selectors in range and the tables present (as recode_scans does)
                if (td[c] >= 4 || ta[c] >= 4 || !dc_t[td[c]].set || !ac_t[ta[c]].set) { err = "scan refers to a missing huffman table"; return false; }
            }
            hpos += len;
            found = true;
            break;
        }
        hpos += len;
    }
    if (!found) { err = "overran headers"; return false; }
    if (ncomp != j.ncmp) { err = "non-interleaved multi-scan baseline not handled"; return false; }
    out.clear();
    out.reserve((size_t)lf.jpeg_size + 64);
    out.push_back(0xFF); out.push_back(0xD8);
    out.insert(out.end(), h.begin(), h.begin() + hpos);

    FastWriter bw(out, (size_t)lf.jpeg_size);
    int lastdc[4] = {0, 0, 0, 0};
    const int mcuc = j.mcuc;
    unsigned rst_written = 0;
    const bool rst_limited = lf.rst_cnt_set && !j.rst_cnt.empty();
    int rstw = rsti;
    auto encode_block = [&](int c, int dpos) {
        const int16_t* blk = planes[c] + (size_t)dpos * 64;
        const HuffTable& dct = dc_t[td[c]];
        const HuffTable& act = ac_t[ta[c]];
        // DC (encode_block_seq, recoder.cc:245-262)
        const int16_t dc = blk[49];
        const int16_t diff = (int16_t)(dc - (int16_t)lastdc[c]);
        lastdc[c] = dc;
        int s = bitlen16(diff > 0 ? diff : -diff);
        int nb = diff > 0 ? diff : (diff - 1) + (1 << s);
        bw.put(((uint32_t)dct.ecode[s] << s) | (uint32_t)nb, dct.elen[s] + s);
        // AC: walk the non-zero coefficients in zig-zag order
        uint64_t m = nonzero_mask_zigzag(blk) >> 1;           // bit k: zig-zag position k + 1
        int prev = 0;
        while (m) {
            const int z = __builtin_ctzll(m) + 1;
            m &= m - 1;
            int run = z - prev - 1;
            prev = z;
            while (run >= 16) { bw.put(act.ecode[0xF0], act.elen[0xF0]); run -= 16; }
            const int v = blk[k_zigzag_to_aligned[z]];
            s = bitlen16(v > 0 ? v : -v);
            nb = v > 0 ? v : (v - 1) + (1 << s);
            const int hc = (run << 4) + s;
            bw.put(((uint32_t)act.ecode[hc] << s) | (uint32_t)nb, act.elen[hc] + s);
        }
        if (prev != 63) bw.put(act.ecode[0x00], act.elen[0x00]);
    };
    auto restart = [&]() {      // recoder.cc:381-397
        bw.pad((uint8_t)j.padbit);
        if (!rst_limited || rst_written < j.rst_cnt[0]) {
            bw.raw(0xFF);
            bw.raw((uint8_t)(0xD0 + (rst_written & 7)));
            rst_written++;
        }
        rstw = rsti;
        lastdc[0] = lastdc[1] = lastdc[2] = lastdc[3] = 0;
    };
    if (j.ncmp > 1) {
        for (int mcu = 0; mcu < mcuc; ++mcu) {
            const int my = mcu / j.mcuh, mx = mcu - my * j.mcuh;
            for (int ci = 0; ci < ncomp; ++ci) {
                const int c = scmp[ci];
                const Component& k = j.cmp[c];
                for (int sub = 0; sub < k.mbs; ++sub) {
                    const int sy = sub / k.H, sx = sub - sy * k.H;
                    encode_block(c, (my * k.V + sy) * k.bch + mx * k.H + sx);
                }
            }
            if (mcu + 1 < mcuc && rsti > 0 && --rstw == 0) restart();
        }
    } else {
        // single component: next_mcuposn order (jpgcoder.cc:5432-5456)
        const Component& k = j.cmp[0];
        int dpos = 0;
        while (true) {
            encode_block(0, dpos);
            dpos++;
            if (k.bch != k.nch && dpos % k.bch == k.nch) dpos += k.bch - k.nch;
            if (k.bcv != k.ncv && dpos / k.bch == k.ncv) dpos = k.bc;
            if (dpos >= k.bc) break;
            if (rsti > 0 && --rstw == 0) restart();
        }
    }
    bw.pad((uint8_t)j.padbit);
    // trailing bogus restart markers of the (only) scan (recoder.cc:839-848)
    if (!j.rst_err.empty()) {
        const unsigned cum = rsti ? (unsigned)(j.mcuh * j.mcuv - 1) / rsti : 0;
        for (unsigned i = 0; i < j.rst_err[0]; ++i) { bw.raw(0xFF); bw.raw((uint8_t)(0xD0 + ((cum + i) & 7))); }
    }
    bw.finish();
    out.insert(out.end(), h.begin() + hpos, h.end());      // header data after the first SOS, if any
    // everything before the garbage is bounded to (original size - garbage size): for truncated originals the scan is
    // cut exactly where the file ended (str_out->set_bound, recoder.cc:699-700, 880-886)
    if (lf.jpeg_size >= j.grb.size() && out.size() > lf.jpeg_size - j.grb.size()) out.resize(lf.jpeg_size - j.grb.size());
    out.insert(out.end(), j.grb.begin(), j.grb.end());
    if (out.size() != lf.jpeg_size) { err = "re-created JPEG has the wrong size"; return false; }
    return true;
}


// ------------------------------------------------------------------------------------------------
// General (multi-scan) re-encoder for containers flagged 'X': progressive JPEGs and sequential files whose scans do
// not interleave every component.  Restates recode_jpeg (jpgcoder.cc:3309-3724) + the block routines
// (encode_dc_prg_fs :4993, encode_ac_prg_fs :5078, encode_dc_prg_sa :5137, encode_ac_prg_sa :5247, encode_eobrun :5346,
// encode_crbits :5379) and the marker/stuffing pass merge_jpeg_streaming (:2562-2740), fused into one front-to-back walk.
// ------------------------------------------------------------------------------------------------
namespace {

struct ScanHdr { int ncomp = 0, cmp[4] = {0, 0, 0, 0}, from = 0, to = 0, sah = 0, sal = 0; };

inline int fdiv2(int v, int p) { return v < 0 ? -((-v) >> p) : (v >> p); }

struct ProgWriter {
    FastWriter& bw;
    std::vector<uint32_t> crwords;        // stored correction bits (abytewriter "storw"), 32 to a word, oldest first
    uint32_t crcur = 0;
    int crn = 0;
    unsigned eobrun = 0;
    bool bad = false;                      // coefficients the scan's tables cannot express (inconsistent .lep)
    explicit ProgWriter(FastWriter& b) : bw(b) {}
    inline void push_crbit(uint32_t b) {
        crcur = (crcur << 1) | b;
        if (++crn == 32) { crwords.push_back(crcur); crcur = 0; crn = 0; }
    }
    void flush_crbits() {
        for (uint32_t w : crwords) bw.put(w, 32);
        crwords.clear();
        if (crn) { bw.put(crcur, crn); crcur = 0; crn = 0; }
    }
    void flush_eobrun(const HuffTable& t) {
        if (eobrun == 0) return;
        if (t.max_eobrun <= 0) { bad = true; eobrun = 0; return; }      // the table has no end-of-band code at all
        while (eobrun > (unsigned)t.max_eobrun) {
            bw.put(t.ecode[0xE0], t.elen[0xE0]);
            bw.put(32767 - (1 << 14), 14);
            eobrun -= (unsigned)t.max_eobrun;
        }
        int s = bitlen16((int)eobrun);
        if (s) --s;
        bw.put(t.ecode[s << 4], t.elen[s << 4]);
        bw.put(eobrun - (1u << s), s);
        eobrun = 0;
    }
};

}  // namespace

bool recode_scans(const LepFile& lf, const int16_t* const planes[4], std::vector<uint8_t>& out, std::string& err) {
    const Jpeg& j = lf.j;
    const std::vector<uint8_t>& h = j.hdr;
    HuffTable dc_t[4], ac_t[4];
    int td[4] = {0, 0, 0, 0}, ta[4] = {0, 0, 0, 0};
    int rsti = 0;
    size_t hpos = 0;
    out.clear();
    out.reserve((size_t)lf.jpeg_size + 64);
    out.push_back(0xFF); out.push_back(0xD8);
    FastWriter bw(out, (size_t)lf.jpeg_size);
    ProgWriter pw(bw);
    bool rst_stuck = false;                // a refused marker is never retried (rpos stops advancing, :2640-2650)
    int scan = 0;
    while (true) {
        // ---- header segments up to and including the next SOS; DHT / DRI / SOS are interpreted on the way
        ScanHdr sc;
        uint8_t type = 0;
        const size_t seg_begin = hpos;
        while (type != 0xDA) {
            if (hpos + 3 >= h.size()) break;
            type = h[hpos + 1];
            const size_t len = 2 + be16(&h[hpos + 2]);
            if (hpos + len > h.size()) { err = "truncated header segment"; return false; }
            const uint8_t* seg = &h[hpos];
            if (type == 0xC4) {
                size_t p = 4;
                while (p < len) {
                    const int tc = seg[p] >> 4, th = seg[p] & 15;
                    if (tc >= 2 || th >= 4) break;
                    ++p;
                    HuffTable& t = tc ? ac_t[th] : dc_t[th];
                    int total = 0;
                    t.bits[0] = 0;
                    for (int i = 0; i < 16; ++i) { t.bits[i + 1] = seg[p + i]; total += seg[p + i]; }
                    if (total > 256 || p + 16 + total > len) { err = "bad DHT"; return false; }
                    memcpy(t.vals, seg + p + 16, total);
                    if (!t.build()) { err = "bad huffman table"; return false; }
                    t.set = true;
                    p += 16 + total;
                }
            } else if (type == 0xDD) {
                if (len >= 6) rsti = be16(seg + 4);
            } else if (type == 0xDA) {
                if (len < 5) { err = "bad SOS"; return false; }
                sc.ncomp = seg[4];
                if (sc.ncomp < 1 || sc.ncomp > j.ncmp || len < (size_t)(8 + 2 * sc.ncomp)) { err = "bad SOS"; return false; }
                for (int i = 0; i < sc.ncomp; ++i) {
                    int c = 0;
                    while (c < j.ncmp && j.cmp[c].jid != seg[5 + 2 * i]) ++c;
                    if (c == j.ncmp) { err = "component id mismatch"; return false; }
                    sc.cmp[i] = c; td[c] = seg[6 + 2 * i] >> 4; ta[c] = seg[6 + 2 * i] & 15;
                    if (td[c] >= 4 || ta[c] >= 4) { err = "huffman table number mismatch"; return false; }
                }
                const uint8_t* t = seg + 5 + 2 * sc.ncomp;
                sc.from = t[0]; sc.to = t[1]; sc.sah = t[2] >> 4; sc.sal = t[2] & 15;
                if (sc.from > sc.to || sc.to > 63 || sc.sah >= 12 || sc.sal >= 12) { err = "scan parameters out of range"; return false; }
            }
            hpos += len;
        }
        bw.raw_bytes(h.data() + seg_begin, std::min(hpos, h.size()) - seg_begin);
        if (type != 0xDA) break;
        ++scan;
        // ---- one scan
        unsigned cpos = 0, rst_this_scan = 0;
        auto rst_ok = [&]() {              // rst_cnt_ok (:2509-2517)
            if (rsti == 0 || rst_stuck) return false;
            if (!lf.rst_cnt_set) return true;
            return j.rst_cnt.size() > (size_t)scan - 1 && rst_this_scan < j.rst_cnt[scan - 1];
        };
        int cmp = sc.cmp[0], csc = 0, mcu = 0, sub = 0, dpos = 0;
        const bool inter = sc.ncomp > 1;
        auto coef = [&](int bpos) -> int { return planes[cmp][(size_t)dpos * 64 + k_zigzag_to_aligned[bpos]]; };
        auto advance = [&](int& rstw) -> int {
            if (inter) {                   // next_mcupos (recoder.cc:190-243)
                int sta = 0;
                if (++sub >= j.cmp[cmp].mbs) {
                    sub = 0;
                    if (++csc >= sc.ncomp) {
                        csc = 0; cmp = sc.cmp[0]; ++mcu;
                        if (mcu >= j.mcuc) sta = 2;
                        else if (rsti > 0 && --rstw == 0) sta = 1;
                    } else cmp = sc.cmp[csc];
                }
                const Component& k = j.cmp[cmp];
                if (k.V > 1) {
                    const int my = mcu / j.mcuh, mx = mcu - my * j.mcuh, sy = sub / k.H, sx = sub - sy * k.H;
                    dpos = (my * k.V + sy) * k.bch + mx * k.H + sx;
                } else if (k.H > 1) dpos = mcu * k.mbs + sub;
                else dpos = mcu;
                return sta;
            }
            const Component& k = j.cmp[cmp];          // next_mcuposn (jpgcoder.cc:5432-5456)
            dpos++;
            if (k.bch != k.nch && dpos % k.bch == k.nch) dpos += k.bch - k.nch;
            if (k.bcv != k.ncv && dpos / k.bch == k.ncv) dpos = k.bc;
            if (dpos >= k.bc) return 2;
            if (rsti > 0 && --rstw == 0) return 1;
            return 0;
        };
        for (int i = 0; i < sc.ncomp; ++i) {
            const int c = sc.cmp[i];
            const bool need_dc = j.jpegtype == 1 || ((inter || sc.to == 0) && sc.sah == 0);
            const bool need_ac = j.jpegtype == 1 || (!inter && sc.to > 0);
            if ((need_dc && !dc_t[td[c]].set) || (need_ac && !ac_t[ta[c]].set)) { err = "huffman table missing in scan"; return false; }
        }
        while (true) {
            int lastdc[4] = {0, 0, 0, 0};
            int sta = 0, rstw = rsti;
            pw.eobrun = 0;
            if (j.jpegtype == 1) {
                // ---- sequential scan (encode_block_seq, recoder.cc:245-313)
                while (sta == 0) {
                    const HuffTable& dct = dc_t[td[cmp]];
                    const HuffTable& act = ac_t[ta[cmp]];
                    const int16_t dc = (int16_t)coef(0);
                    const int16_t diff = (int16_t)(dc - (int16_t)lastdc[cmp]);
                    lastdc[cmp] = dc;
                    int s = bitlen16(diff > 0 ? diff : -diff);
                    int nb = diff > 0 ? diff : (diff - 1) + (1 << s);
                    bw.put(dct.ecode[s], dct.elen[s]);
                    bw.put((uint32_t)nb, s);
                    uint64_t m = nonzero_mask_zigzag(planes[cmp] + (size_t)dpos * 64) & ~1ull;
                    const int end = m ? 63 - __builtin_clzll(m) : 0;
                    int prev = 0;
                    while (m) {
                        const int bpos = __builtin_ctzll(m);
                        m &= m - 1;
                        int z = bpos - prev - 1;
                        prev = bpos;
                        const int v = coef(bpos);
                        while (z & 0xf0) { bw.put(act.ecode[0xF0], act.elen[0xF0]); z -= 16; }
                        s = bitlen16(v > 0 ? v : -v);
                        nb = v > 0 ? v : (v - 1) + (1 << s);
                        const int hc = ((z & 0xf) << 4) + s;
                        bw.put(act.ecode[hc], act.elen[hc]);
                        bw.put((uint32_t)nb, s);
                    }
                    if (end != 63) bw.put(act.ecode[0x00], act.elen[0x00]);
                    sta = advance(rstw);
                }
            } else if (inter || sc.to == 0) {
                if (sc.sah == 0) {
                    // ---- DC first stage
                    while (sta == 0) {
                        const HuffTable& dct = dc_t[td[cmp]];
                        const int tmp = coef(0) >> sc.sal;
                        const int16_t diff = (int16_t)(tmp - lastdc[cmp]);
                        lastdc[cmp] = tmp;
                        const int s = bitlen16(diff > 0 ? diff : -diff);
                        const int nb = diff > 0 ? diff : (diff - 1) + (1 << s);
                        bw.put(dct.ecode[s], dct.elen[s]);
                        bw.put((uint32_t)nb, s);
                        sta = advance(rstw);
                    }
                } else {
                    // ---- DC refinement bit
                    while (sta == 0) {
                        bw.put((uint32_t)((coef(0) >> sc.sal) & 1), 1);
                        sta = advance(rstw);
                    }
                }
            } else if (sc.sah == 0) {
                // ---- AC first stage
                const HuffTable& act = ac_t[ta[cmp]];
                const uint64_t band = (sc.to == 63 ? ~0ull : (1ull << (sc.to + 1)) - 1) & ~((1ull << sc.from) - 1);
                while (sta == 0) {
                    uint64_t m = magnitude_mask_zigzag(planes[cmp] + (size_t)dpos * 64, 1 << sc.sal) & band;
                    int prev = sc.from - 1;
                    if (m) pw.flush_eobrun(act);
                    while (m) {
                        const int bpos = __builtin_ctzll(m);
                        m &= m - 1;
                        int z = bpos - prev - 1;
                        prev = bpos;
                        const int tmp = fdiv2(coef(bpos), sc.sal);
                        while (z >= 16) { bw.put(act.ecode[0xF0], act.elen[0xF0]); z -= 16; }
                        const int a = tmp > 0 ? tmp : -tmp;
                        const int s = bitlen16(a);
                        const int nb = tmp > 0 ? tmp : (tmp - 1) + (1 << s);
                        const int hc = (z << 4) + s;
                        bw.put(act.ecode[hc], act.elen[hc]);
                        bw.put((uint32_t)nb, s);
                    }
                    if (prev < sc.to) {
                        ++pw.eobrun;
                        if (pw.eobrun == (unsigned)act.max_eobrun) pw.flush_eobrun(act);
                    }
                    sta = advance(rstw);
                }
                pw.flush_eobrun(act);
            } else {
                // ---- AC refinement
                const HuffTable& act = ac_t[ta[cmp]];
                const uint64_t band = (sc.to == 63 ? ~0ull : (1ull << (sc.to + 1)) - 1) & ~((1ull << sc.from) - 1);
                while (sta == 0) {
                    // sig: non-zero at this bit plane; old: significant before this scan (|v| >> sal >= 2);
                    // the remaining sig positions turn significant here (|v| >> sal == 1)
                    const int16_t* blkp = planes[cmp] + (size_t)dpos * 64;
                    uint64_t sig = magnitude_mask_zigzag(blkp, 1 << sc.sal) & band;
                    const uint64_t old = magnitude_mask_zigzag(blkp, 2 << sc.sal) & band;
                    const uint64_t fresh = sig & ~old;
                    const int eob = fresh ? 64 - __builtin_clzll(fresh) : sc.from;
                    if (eob > sc.from && pw.eobrun > 0) { pw.flush_eobrun(act); pw.flush_crbits(); }
                    int z = 0, prev = sc.from - 1;
                    while (sig) {
                        const int bpos = __builtin_ctzll(sig);
                        sig &= sig - 1;
                        const int v = coef(bpos);
                        if (bpos >= eob) {                 // behind the last new coefficient: correction bits only
                            pw.push_crbit((uint32_t)(((v < 0 ? -v : v) >> sc.sal) & 1));
                            continue;
                        }
                        z += bpos - prev - 1;
                        prev = bpos;
                        while (z >= 16) { bw.put(act.ecode[0xF0], act.elen[0xF0]); pw.flush_crbits(); z -= 16; }
                        if ((fresh >> bpos) & 1) {
                            const int hc = (z << 4) + 1;
                            bw.put(act.ecode[hc], act.elen[hc]);
                            bw.put(v > 0 ? 1u : 0u, 1);
                            pw.flush_crbits();
                            z = 0;
                        } else {
                            pw.push_crbit((uint32_t)(((v < 0 ? -v : v) >> sc.sal) & 1));
                        }
                    }
                    if (eob <= sc.to) {
                        ++pw.eobrun;
                        if (pw.eobrun == (unsigned)act.max_eobrun) { pw.flush_eobrun(act); pw.flush_crbits(); }
                    }
                    sta = advance(rstw);
                }
                pw.flush_eobrun(act);
                pw.flush_crbits();
            }
            bw.pad((uint8_t)j.padbit);
            if (sta == 2) break;
            // restart: marker after the padded byte when the scan's marker budget allows (:2640-2650)
            if (rsti > 0) {
                if (rst_ok()) {
                    bw.raw(0xFF);
                    bw.raw((uint8_t)(0xD0 + (cpos & 7)));
                    ++cpos; ++rst_this_scan;
                } else {
                    rst_stuck = true;
                }
            }
        }
        // bogus trailing restart markers of this scan (:2711-2719)
        if ((size_t)scan - 1 < j.rst_err.size())
            for (unsigned i = 0; i < j.rst_err[scan - 1]; ++i) { bw.raw(0xFF); bw.raw((uint8_t)(0xD0 + (cpos & 7))); ++cpos; }
    }
    bw.finish();
    if (scan == 0) { err = "no scan found"; return false; }
    if (pw.bad) { err = "coefficients not expressible with the scan's huffman tables"; return false; }
    if (lf.jpeg_size >= j.grb.size() && out.size() > lf.jpeg_size - j.grb.size()) out.resize(lf.jpeg_size - j.grb.size());
    out.insert(out.end(), j.grb.begin(), j.grb.end());
    if (out.size() != lf.jpeg_size) { err = "re-created JPEG has the wrong size"; return false; }
    return true;
}

}  // namespace lephost
