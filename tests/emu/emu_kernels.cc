// emu_kernels.cc -- runs the coder kernels on the CPU under the warp emulator of cuda_shim.h (test infrastructure).
//
// The kernel sources are compiled as host C++; this file builds the same job descriptors lep_capi.cu's build_batch /
// lepb200_decode_upload build for the device (ImageDesc, SegDesc, order, work counter, zeroed model and row pools, zeroed
// planes) with host addresses in place of device addresses and launches the kernel with its device launch shape.
// Entry points mirror lepb200_decode_images / lepb200_encode_images (include/lepton_b200.h).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "cuda_shim.h"
#include "../../lepton_b200/csrc/lep_encode.cu"
#include "../../lepton_b200/csrc/lep_decode.cu"
#include "../../lepton_b200/csrc/lep_decode_g2.cu"
#include "../../lepton_b200/csrc/lep_huffpar.cu"
#include "../../lepton_b200/csrc/lep_mux.cu"
#include "../../lepton_b200/csrc/lep_huffenc.cu"
#include "../../include/lepton_b200.h"

using namespace lepb200;

namespace {

const uint8_t k_zigzag[64] = {
    0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
const int k_icos_base_col0[8] = {8192, 11363, 10703, 9633, 8192, 6436, 4433, 2260};
const uint16_t k_freqmax[64] = {
    1024, 931, 985, 968, 1020, 968, 1020, 1020, 932, 858, 884, 840, 932, 838, 854, 854,
    985, 884, 871, 875, 985, 878, 871, 854, 967, 841, 876, 844, 967, 886, 870, 837,
    1020, 932, 985, 967, 1020, 969, 1020, 1020, 969, 838, 878, 886, 969, 838, 969, 838,
    1020, 854, 871, 870, 1010, 969, 1020, 1020, 1020, 854, 854, 838, 1020, 838, 1020, 838};

// ProbabilityTablesBase::set_quantization_table (src/vp8/model/model.hh:247-290), as fill_quant in lep_capi.cu
int fill_quant(ImageDesc& d, int c, const uint16_t zz[64]) {
    uint16_t* q = d.q[c];
    for (int i = 0; i < 64; ++i) q[i] = zz[k_zigzag[i]];
    for (int r = 0; r < 8; ++r) {
        for (int i = 0; i < 8; ++i) {
            d.icos_x[c][r * 8 + i] = k_icos_base_col0[i] * (int)q[i * 8 + r];
            d.icos_y[c][r * 8 + i] = k_icos_base_col0[i] * (int)q[r * 8 + i];
        }
        if (d.icos_x[c][r * 8] == 0 || d.icos_y[c][r * 8] == 0) return LEPB200_ST_UNSUPPORTED_JPEG_WITH_ZERO_IDCT_0;
    }
    for (int k = 0; k < 64; ++k) {
        uint16_t fm = (uint16_t)(k_freqmax[k] + q[k] - 1);
        if (q[k]) fm = (uint16_t)(fm / q[k]);
        int len = 0;
        for (uint32_t v = fm; v; v >>= 1) ++len;
        d.min_thr[c][k] = (uint8_t)(len > 7 ? len - 7 : 0);
    }
    return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct LaunchArgs {
    int kernel;
    const ImageDesc* images; SegDesc* segs; int nseg; const int* order; int* counter;
    uint16_t* models; uint8_t* rows; size_t row_stride;
};

void kernel_body(void* p) {
    const LaunchArgs& a = *static_cast<const LaunchArgs*>(p);
    if (a.kernel == 0) lep_decode_kernel(a.images, a.segs, a.nseg, a.order, a.counter, a.models, a.rows, a.row_stride);
    else if (a.kernel == 201) lep_decode_g2_kernel<1>(a.images, a.segs, 0, a.nseg, a.order, a.counter, a.models, a.rows, a.row_stride);
    else if (a.kernel == 202) lep_decode_g2_kernel<2>(a.images, a.segs, 0, a.nseg, a.order, a.counter, a.models, a.rows, a.row_stride);
    else if (a.kernel == 204) lep_decode_g2_kernel<4>(a.images, a.segs, 0, a.nseg, a.order, a.counter, a.models, a.rows, a.row_stride);
    else if (a.kernel == 208) lep_decode_g2_kernel<8>(a.images, a.segs, 0, a.nseg, a.order, a.counter, a.models, a.rows, a.row_stride);
    else if (a.kernel == 216) lep_decode_g2_kernel<16>(a.images, a.segs, 0, a.nseg, a.order, a.counter, a.models, a.rows, a.row_stride);
    else lep_decode_g2_kernel<32>(a.images, a.segs, 0, a.nseg, a.order, a.counter, a.models, a.rows, a.row_stride);
}

// launch shape of the group kernel: warps per CTA and thread-segments per warp for G lanes per segment
void group_shape(int G, int& warps, int& per_warp) {
    per_warp = 32 / G;
    warps = G >= 4 ? 4 : G;
}

}  // namespace

// kernel: 0 = lep_decode_kernel (warp per segment, persistent CTAs; `grid_cap` > 0 limits the CTAs so that warps take
// several segments from the queue), 200 + G = lep_decode_g2_kernel<G> (G lanes per segment).  Decodes into images[i].planes (zeroed first,
// like the device arena); per-segment status and decision counts come back like lepb200_decode_fetch reports them.
extern "C" int emu_decode_images(int kernel, int grid_cap, const lepb200_image* images, int nimages, const lepb200_stream* in,
                                 int32_t* status_out, uint64_t* ndecisions_out) {
    const int gl = kernel % 100;
    const bool group = kernel / 100 == 2 && (gl == 1 || gl == 2 || gl == 4 || gl == 8 || gl == 16 || gl == 32);
    if (nimages <= 0 || !images || !in || (kernel != 0 && !group)) return LEPB200_ERR_INVALID;
    std::vector<ImageDesc> descs(nimages);
    std::vector<SegDesc> segs;
    std::vector<size_t> seg_blocks;
    std::vector<std::vector<uint8_t>> padded;          // streams with the 16 readable bytes the device arena has behind them
    size_t row_stride = 0;
    for (int i = 0; i < nimages; ++i) {
        const lepb200_image& im = images[i];
        if (im.ncmp < 1 || im.ncmp > 3 || im.mcuv <= 0 || im.nseg < 1 || im.nseg > LEPB200_MAX_SEGMENTS) return LEPB200_ERR_INVALID;
        ImageDesc& d = descs[i];
        memset(&d, 0, sizeof(d));
        d.ncmp = im.ncmp; d.mcuv = im.mcuv;
        int qstatus = 0;
        size_t rs = 0;
        for (int c = 0; c < im.ncmp; ++c) {
            d.bch[c] = im.bch[c]; d.bcv[c] = im.bcv[c]; d.trunc_bcv[c] = im.trunc_bcv[c]; d.trunc_bc[c] = im.trunc_bc[c];
            d.mult[c] = im.bcv[c] / im.mcuv;
            const int qs = fill_quant(d, c, im.qtable_zigzag[c]);
            if (qs) qstatus = qs;
            d.plane[c] = (unsigned long long)(uintptr_t)im.planes[c];
            memset(im.planes[c], 0, (size_t)im.bch[c] * im.bcv[c] * 128);
            rs += (size_t)im.bch[c] * 16 + align_up((size_t)im.bch[c], 16);
        }
        row_stride = std::max(row_stride, align_up(rs, 256));
        for (int s = 0; s < im.nseg; ++s) {
            SegDesc sd;
            memset(&sd, 0, sizeof(sd));
            sd.image = i;
            sd.min_y = im.luma_y_start[s];
            sd.is_last = s + 1 == im.nseg;
            sd.max_y = sd.is_last ? im.bcv[0] : im.luma_y_start[s + 1];
            sd.status = qstatus;
            const lepb200_stream& st = in[segs.size()];
            padded.emplace_back((size_t)st.len + 16, 0xA5);      // the device arena is not cleared: whatever is behind a stream must not matter
            if (st.len) memcpy(padded.back().data(), st.data, (size_t)st.len);
            sd.cap = (uint32_t)st.len;
            size_t nb = 0;                                  // segment_blocks of lep_capi.cu: only the launch order depends on it
            const int v0 = std::max(im.bcv[0] / im.mcuv, 1);
            for (int c = 0; c < im.ncmp; ++c) {
                const int mult = im.bcv[c] / im.mcuv;
                long y0 = (long)(sd.min_y / v0) * mult, y1 = sd.is_last ? im.trunc_bcv[c] : (long)((sd.max_y + v0 - 1) / v0) * mult;
                y1 = std::min<long>(y1, im.trunc_bcv[c]);
                if (y1 > y0) nb += (size_t)(y1 - y0) * im.bch[c];
            }
            seg_blocks.push_back(nb);
            segs.push_back(sd);
        }
    }
    const int nseg = (int)segs.size();
    for (int s = 0; s < nseg; ++s) segs[s].stream = (unsigned long long)(uintptr_t)padded[s].data();
    std::vector<int> order(nseg);
    for (int i = 0; i < nseg; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return seg_blocks[a] > seg_blocks[b]; });

    LaunchArgs a;
    a.kernel = kernel; a.images = descs.data(); a.segs = segs.data(); a.nseg = nseg; a.order = order.data(); a.row_stride = row_stride;
    int counter = 0;
    a.counter = &counter;
    unsigned grid, block;
    size_t group_slots = 0;
    if (group) {
        // 100 + G: lep_decode_group_kernel<G>; grid_cap > 0 limits the CTAs so that groups take several segments from the queue
        int warps, per_warp;
        group_shape(kernel % 100, warps, per_warp);
        grid = (unsigned)((nseg + warps * per_warp - 1) / (warps * per_warp));
        if (grid_cap > 0) grid = std::min(grid, (unsigned)grid_cap);
        block = (unsigned)warps * 32;
        group_slots = (size_t)grid * warps * per_warp;
    } else {
        grid = (unsigned)((nseg + DEC_WARPS_PER_CTA - 1) / DEC_WARPS_PER_CTA);
        if (grid_cap > 0) grid = std::min(grid, (unsigned)grid_cap);
        block = DEC_WARPS_PER_CTA * 32;
    }
    const size_t slots = kernel == 0 ? (size_t)grid * DEC_WARPS_PER_CTA : (size_t)nseg;
    std::vector<uint16_t> models(slots * M_TOTAL, kernel == 0 ? 0x5a5a : 0);       // thread / group kernels: zero fill before the launch; the warp kernel clears its own
    std::vector<uint8_t> rows((group ? group_slots : slots) * row_stride, 0);
    a.models = models.data(); a.rows = rows.data();
    emu::launch(grid, block, kernel_body, &a);
    for (int s = 0; s < nseg; ++s) {
        if (status_out) status_out[s] = segs[s].status;
        if (ndecisions_out) ndecisions_out[s] = (uint64_t)segs[s].ndecisions_lo | ((uint64_t)segs[s].ndecisions_hi << 32);
    }
    return 0;
}

namespace {

struct EncArgs {
    int stage;                       // 0 count, 1 offsets, 2 symbolise (kernel A), 3 range coder (kernel B)
    const ImageDesc* images; SegDesc* segs; int nseg; const int* order; int* counter;
    uint16_t* models; uint8_t* rows; size_t row_stride; uint16_t* tokens; unsigned long long* total;
    unsigned long long* ck; uint32_t* digits;
};

void enc_body(void* p) {
    const EncArgs& a = *static_cast<const EncArgs*>(p);
    if (a.stage == 0) lep_count_kernel(a.images, a.segs, a.nseg);
    else if (a.stage == 1) lep_token_offsets_kernel(a.segs, a.nseg, a.total);
    else if (a.stage == 2) lep_encode_kernel(a.images, a.segs, a.nseg, a.order, a.counter, a.models, a.rows, a.row_stride, a.tokens);
    else if (a.stage == 3) lep_rangecode_kernel(a.segs, a.nseg, a.order, a.tokens);
    else if (a.stage == 5) lep_rangepass_kernel<true>(a.segs, a.nseg, a.order, a.tokens, a.ck);
    else if (a.stage == 4) lep_rangepass_kernel<false>(a.segs, a.nseg, a.order, a.tokens, a.ck);
    else if (a.stage == 6) lep_digit_offsets_kernel(a.segs, a.nseg, a.total);
    else if (a.stage == 7) lep_rangepiece_kernel(a.segs, a.nseg, a.tokens, a.ck, a.digits);
    else lep_rangenorm_kernel(a.segs, a.nseg, a.order, a.digits);
}

}  // namespace

// lepb200_encode_images on the emulator: count pre-pass -> token offsets -> kernel A (symbolise + model) -> kernel B
// (range coder), with the launch shapes of lep_capi.cu.  `out` must hold sum(nseg) entries; the bytes of every stream are
// copied into `arena` (capacity arena_cap) back to back and out[i].data points there.  grid_cap > 0 limits kernel A's
// CTAs (persistent warps then take several segments each).  kernel: 0 = parallel range coder, 1 = serial range coder,
// 2 = parallel range coder whose range pass feeds tokens through registers (LEPB200_RC_FEED=0).
extern "C" int emu_encode_images(int kernel, int grid_cap, const lepb200_image* images, int nimages, lepb200_stream* out, uint8_t* arena, size_t arena_cap) {
    if (nimages <= 0 || !images || !out || !arena) return LEPB200_ERR_INVALID;
    std::vector<ImageDesc> descs(nimages);
    std::vector<SegDesc> segs;
    std::vector<size_t> seg_blocks;
    size_t row_stride = 0, stream_total = 0;
    for (int i = 0; i < nimages; ++i) {
        const lepb200_image& im = images[i];
        if (im.ncmp < 1 || im.ncmp > 3 || im.mcuv <= 0 || im.nseg < 1 || im.nseg > LEPB200_MAX_SEGMENTS) return LEPB200_ERR_INVALID;
        ImageDesc& d = descs[i];
        memset(&d, 0, sizeof(d));
        d.ncmp = im.ncmp; d.mcuv = im.mcuv;
        int qstatus = 0;
        size_t rs = 0;
        for (int c = 0; c < im.ncmp; ++c) {
            d.bch[c] = im.bch[c]; d.bcv[c] = im.bcv[c]; d.trunc_bcv[c] = im.trunc_bcv[c]; d.trunc_bc[c] = im.trunc_bc[c];
            d.mult[c] = im.bcv[c] / im.mcuv;
            const int qs = fill_quant(d, c, im.qtable_zigzag[c]);
            if (qs) qstatus = qs;
            d.plane[c] = (unsigned long long)(uintptr_t)im.planes[c];
            rs += (size_t)im.bch[c] * 16 + align_up((size_t)im.bch[c], 16);
        }
        row_stride = std::max(row_stride, align_up(rs, 256));
        for (int s = 0; s < im.nseg; ++s) {
            SegDesc sd;
            memset(&sd, 0, sizeof(sd));
            sd.image = i;
            sd.min_y = im.luma_y_start[s];
            sd.is_last = s + 1 == im.nseg;
            sd.max_y = sd.is_last ? im.bcv[0] : im.luma_y_start[s + 1];
            sd.status = qstatus;
            size_t nb = 0;
            const int v0 = std::max(im.bcv[0] / im.mcuv, 1);
            for (int c = 0; c < im.ncmp; ++c) {
                const int mult = im.bcv[c] / im.mcuv;
                long y0 = (long)(sd.min_y / v0) * mult, y1 = sd.is_last ? im.trunc_bcv[c] : (long)((sd.max_y + v0 - 1) / v0) * mult;
                y1 = std::min<long>(y1, im.trunc_bcv[c]);
                if (y1 > y0) nb += (size_t)(y1 - y0) * im.bch[c];
            }
            seg_blocks.push_back(nb);
            const size_t cap = align_up(nb * 64 + 4096, 256);
            sd.stream = stream_total; sd.cap = (uint32_t)cap;
            stream_total += cap;
            segs.push_back(sd);
        }
    }
    const int nseg = (int)segs.size();
    std::vector<uint8_t> streams(stream_total + 256, 0);
    for (auto& sd : segs) sd.stream += (unsigned long long)(uintptr_t)streams.data();
    std::vector<int> order(nseg);
    for (int i = 0; i < nseg; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return seg_blocks[a] > seg_blocks[b]; });

    EncArgs a;
    a.images = descs.data(); a.segs = segs.data(); a.nseg = nseg; a.order = order.data(); a.row_stride = row_stride;
    int counter = 0;
    unsigned long long total_tokens = 0;
    a.counter = &counter; a.total = &total_tokens; a.models = nullptr; a.rows = nullptr; a.tokens = nullptr;
    a.stage = 0; emu::launch((unsigned)nseg, CNT_THREADS, enc_body, &a);
    a.stage = 1; emu::launch(1, 1024, enc_body, &a);
    std::vector<uint16_t> tokens((size_t)total_tokens + 128, 0);
    a.tokens = tokens.data();
    std::vector<uint16_t> models;
    std::vector<uint8_t> rows;
    {
        unsigned grid = (unsigned)((nseg + ENC_WARPS_PER_CTA - 1) / ENC_WARPS_PER_CTA);
        if (grid_cap > 0) grid = std::min(grid, (unsigned)grid_cap);
        models.assign((size_t)grid * ENC_WARPS_PER_CTA * M_TOTAL, 0x5a5a);      // the kernel clears its own
        rows.assign((size_t)grid * ENC_WARPS_PER_CTA * row_stride, 0);
        a.models = models.data(); a.rows = rows.data();
        a.stage = 2; emu::launch(grid, ENC_WARPS_PER_CTA * 32, enc_body, &a);
    }
    if (const char* dump = getenv("EMU_DUMP_TOKENS")) {          // diagnostics: the (probability, bit) tokens of every segment
        FILE* f = fopen(dump, "wb");
        for (int s = 0; f && s < nseg; ++s) { const uint32_t nt = segs[s].ntok; fwrite(&nt, 4, 1, f); fwrite(tokens.data() + segs[s].tokens, 2, nt, f); }
        if (f) fclose(f);
    }
    std::vector<unsigned long long> ck;
    std::vector<uint32_t> digits;
    if (kernel == 0 || kernel == 2) {   // parallel range coder: range-only pass, digit layout, pieces, carries (lep_capi.cu: rc_mode 1; 2 = register-ring token feed)
        ck.assign(((size_t)total_tokens >> 10) + 2 * (size_t)nseg + 8, 0);
        a.ck = ck.data();
        a.stage = kernel == 0 ? 5 : 4; emu::launch((unsigned)((nseg + RCT_THREADS - 1) / RCT_THREADS), RCT_THREADS, enc_body, &a);
        unsigned long long total_digits = 0;
        a.total = &total_digits;
        a.stage = 6; emu::launch(1, 1024, enc_body, &a);
        digits.assign((size_t)total_digits + 64, 0);
        a.digits = digits.data();
        a.stage = 7;
        for (int y = 0; y < nseg; ++y) { emu::g_block_idx.y = (unsigned)y; emu::launch(8, RCP_THREADS, enc_body, &a); }      // grid (8, nseg)
        emu::g_block_idx.y = 0;
        a.stage = 8; emu::launch((unsigned)((nseg + RCN_WARPS - 1) / RCN_WARPS), RCN_WARPS * 32, enc_body, &a);
    } else {                  // kernel 1: the serial range coder (rc_mode 0)
        a.stage = 3; emu::launch((unsigned)((nseg + RC_THREADS - 1) / RC_THREADS), RC_THREADS, enc_body, &a);
    }
    size_t used = 0;
    for (int s = 0; s < nseg; ++s) {
        const size_t n = segs[s].status == 0 ? segs[s].len : 0;
        if (used + n > arena_cap) return LEPB200_ERR_NOMEM;
        memcpy(arena + used, reinterpret_cast<const uint8_t*>(segs[s].stream), n);
        out[s].data = arena + used;
        out[s].len = n;
        out[s].status = segs[s].status;
        out[s].reserved = 0;
        out[s].ndecisions = (uint64_t)segs[s].ndecisions_lo | ((uint64_t)segs[s].ndecisions_hi << 32);
        used += n;
    }
    return 0;
}


// ---- baseline Huffman decode: lep_huffdecode_kernel alone (mode 0) or the sub-sequence kernels of lep_huffpar.cu followed
// by it (mode 1), with the job set-up and the iteration loop of lepb200_huffman_decode_to_device (lep_capi.cu).
// planes[3 * i + c] = host plane of component c of scan i (zeroed by the caller); scans[i] outputs are filled like the
// device call fills them.  info[0] = synchronisation iterations, info[1] = images the serial kernel had to redo.
namespace {
struct HuffArgs { HuffJob* jobs; int n; const HuffTableDev* tabs; int ntabs; HpArrays a; int iter; int last_iter; unsigned int* dirty; int which; };
void huff_body(void* p) {
    const HuffArgs& h = *static_cast<const HuffArgs*>(p);
    if (h.which == 0) lep_huffdecode_kernel(h.jobs, h.n, h.tabs, h.ntabs);
    else if (h.which == 1) lep_huffpar_sync_kernel(h.jobs, h.n, h.tabs, h.ntabs, h.a, h.iter, h.last_iter, h.dirty);
    else if (h.which == 2) lep_huffpar_prefix_kernel(h.jobs, h.n, h.a);
    else lep_huffpar_write_kernel(h.jobs, h.n, h.tabs, h.ntabs, h.a);
}
}  // namespace

extern "C" int emu_huffman_decode(int mode, int sub_bits, int iter_cap, lepb200_jpeg_scan* scans, int n, int16_t** planes, int* info) {
    std::vector<HuffJob> jobs(n);
    std::vector<HuffTableDev> tabs;
    std::vector<std::vector<HuffRow>> rows(n);
    std::vector<uint32_t> sub_base((size_t)n + 1, 0);
    uint32_t sub_total = 0;
    auto table_index = [&](const lepb200_hufftable& t, bool& ok) -> int {
        HuffTableDev d;
        if (!huff_build_table(t.bits, t.vals, d)) ok = false;
        for (size_t k = 0; k < tabs.size(); ++k) if (!memcmp(&tabs[k], &d, sizeof(d))) return (int)k;
        tabs.push_back(d);
        return (int)tabs.size() - 1;
    };
    std::vector<std::vector<uint8_t>> padded(n);
    for (int i = 0; i < n; ++i) {
        lepb200_jpeg_scan& sc = scans[i];
        HuffJob& jb = jobs[i];
        memset(&jb, 0, sizeof(jb));
        bool ok = sc.ncmp >= 1 && sc.ncmp <= 3 && sc.mcuh > 0 && sc.mcuv > 0 && sc.entropy;
        jb.ncmp = sc.ncmp; jb.mcuh = sc.mcuh; jb.mcuv = sc.mcuv; jb.rsti = sc.rsti; jb.nbytes = sc.nbytes;
        for (int c = 0; ok && c < sc.ncmp; ++c) {
            jb.H[c] = sc.H[c]; jb.V[c] = sc.V[c];
            ok = ok && sc.H[c] >= 1 && sc.H[c] <= 2 && sc.V[c] >= 1 && sc.V[c] <= 2;
            jb.bch[c] = sc.mcuh * sc.H[c]; jb.bcv[c] = sc.mcuv * sc.V[c];
            jb.nch[c] = sc.nch[c]; jb.ncv[c] = sc.ncv[c];
            jb.dc_tab[c] = table_index(sc.dc[c], ok); jb.ac_tab[c] = table_index(sc.ac[c], ok);
            jb.plane[c] = (unsigned long long)(uintptr_t)planes[3 * i + c];
        }
        jb.status = ok ? 0 : LEPB200_ST_NOT_HANDLED;
        padded[i].assign((size_t)sc.nbytes + 32, 0);                 // the kernels read whole words past the end
        if (sc.entropy) memcpy(padded[i].data(), sc.entropy, sc.nbytes);
        jb.huff = (unsigned long long)(uintptr_t)padded[i].data();
        rows[i].assign((size_t)sc.mcuv + 1, HuffRow());
        jb.rows = (unsigned long long)(uintptr_t)rows[i].data();
        const uint64_t bits = (uint64_t)sc.nbytes * 8;
        jb.sub_base = sub_total;
        if (mode == 1 && jb.status == 0 && sc.ncmp > 1 && sc.rsti == 0 && bits >= 4ull * (uint64_t)sub_bits)
            jb.nsub = (uint32_t)((bits + (uint64_t)sub_bits - 1) / (uint64_t)sub_bits);
        sub_total += jb.nsub;
        sub_base[i] = jb.sub_base;
    }
    sub_base[n] = sub_total;
    HuffArgs h;
    memset(&h, 0, sizeof(h));
    h.jobs = jobs.data(); h.n = n; h.tabs = tabs.data(); h.ntabs = (int)tabs.size();
    int iters = 0;
    std::vector<unsigned long long> ex(sub_total + 1);
    std::vector<uint32_t> epoch(sub_total + 1, 0), tok(sub_total + 1, 0);
    std::vector<uint4> cnt(sub_total + 1);
    std::vector<unsigned int> dirty(iter_cap + 4, 0);
    if (sub_total > 0) {
        h.a.exit = ex.data(); h.a.epoch = epoch.data(); h.a.cnt = cnt.data(); h.a.tok = tok.data();
        h.a.sub_base = sub_base.data(); h.a.total = sub_total; h.a.sub_bits = (uint32_t)sub_bits;
        h.dirty = dirty.data(); h.last_iter = iter_cap;
        const unsigned grid = (sub_total + HP_THREADS - 1) / HP_THREADS;
        h.which = 1;
        for (;;) {
            h.iter = iters;
            emu::launch(grid, HP_THREADS, huff_body, &h);
            ++iters;
            if (iters > iter_cap || dirty[iters] == 0) break;
        }
        h.which = 2; emu::launch((n + 127) / 128, 128, huff_body, &h);
        h.which = 3; emu::launch(grid, HP_THREADS, huff_body, &h);
    }
    int redo = 0;
    for (int i = 0; i < n; ++i) if (jobs[i].nsub && !(jobs[i].par_done && !jobs[i].par_redo)) ++redo;
    h.which = 0;
    emu::launch((n + 3) / 4, 4 * 32, huff_body, &h);
    for (int i = 0; i < n; ++i) {
        scans[i].status = jobs[i].status;
        scans[i].padbit = jobs[i].padbit;
        scans[i].end_bitpos = jobs[i].end_bitpos;
        scans[i].nrows = jobs[i].nrows;
        if (scans[i].rows && jobs[i].nrows > 0) memcpy(scans[i].rows, rows[i].data(), sizeof(HuffRow) * (size_t)std::min(jobs[i].nrows, scans[i].mcuv + 1));
    }
    if (info) { info[0] = iters; info[1] = redo; }
    return 0;
}


// ---- device container assembly: lep_gather_kernel over the pieces lepb200_encode_fetch_files builds (gather_file_pieces,
// lep_mux.cu) for `nfiles` files at once.  File f has header hdr[f] (hlen[f] bytes), nseg[f] streams (streams[] holds them file
// after file) and the MuxWriter plan plan[plan_first[f] .. plan_first[f + 1]) (from lepb200_host_mux_plan).  The files are
// written back to back (16-byte aligned starts, like the device buffer) into out; off[f] / len[f] say where.
namespace {
struct GatherArgs { const GatherPiece* pieces; uint32_t n; uint8_t* out; };
void gather_body(void* p) { const GatherArgs& a = *static_cast<const GatherArgs*>(p); lep_gather_kernel(a.pieces, a.n, a.out); }
}  // namespace

extern "C" int emu_mux_files(int nfiles, const uint8_t* const* hdr, const size_t* hlen, const int* nseg, const lepb200_stream* streams,
                             const lepb200_mux_packet* plan, const uint32_t* plan_first, int grid, uint8_t* out, size_t cap, size_t* off, size_t* len) {
    if (nfiles <= 0 || !hdr || !hlen || !nseg || !streams || !plan || !plan_first || !out) return LEPB200_ERR_INVALID;
    // stream arena: 256-byte aligned starts and 256 bytes of slack like d_streams; poisoned so that a byte taken from outside a
    // stream shows
    size_t arena = 256, nstreams = 0;
    for (int f = 0; f < nfiles; ++f) for (int k = 0; k < nseg[f]; ++k) arena += align_up((size_t)streams[nstreams++].len + 16, 256);
    std::vector<uint8_t> sbuf(arena + 512, 0xEE);
    uint8_t* sbase = reinterpret_cast<uint8_t*>(align_up((size_t)(uintptr_t)sbuf.data(), 256));
    size_t lit_total = 0;
    for (int f = 0; f < nfiles; ++f) lit_total += align_up(hlen[f] + 4, 16);
    std::vector<uint8_t> lit(lit_total + 512, 0xDD);
    uint8_t* lbase = reinterpret_cast<uint8_t*>(align_up((size_t)(uintptr_t)lit.data(), 256));
    std::vector<GatherPiece> pieces;
    size_t spos = 0, lpos = 0, total = 0, si = 0;
    for (int f = 0; f < nfiles; ++f) {
        unsigned long long saddr[LEPB200_MAX_SEGMENTS];
        if (nseg[f] < 1 || nseg[f] > LEPB200_MAX_SEGMENTS) return LEPB200_ERR_INVALID;
        for (int k = 0; k < nseg[f]; ++k, ++si) {
            if (streams[si].len) memcpy(sbase + spos, streams[si].data, (size_t)streams[si].len);
            saddr[k] = (unsigned long long)(uintptr_t)(sbase + spos);
            spos += align_up((size_t)streams[si].len + 16, 256);
        }
        memcpy(lbase + lpos, hdr[f], hlen[f]);
        total = align_up(total, 16);
        off[f] = total;
        len[f] = gather_file_pieces(plan + plan_first[f], plan_first[f + 1] - plan_first[f], saddr, (unsigned long long)(uintptr_t)(lbase + lpos),
                                    lbase + lpos, hlen[f], total, pieces);
        total += len[f];
        lpos += align_up(hlen[f] + 4, 16);
    }
    if (total > cap) return LEPB200_ERR_NOMEM;
    std::vector<uint8_t> dense(total + 512, 0xCC);
    uint8_t* dbase = reinterpret_cast<uint8_t*>(align_up((size_t)(uintptr_t)dense.data(), 256));
    GatherArgs a{pieces.data(), (uint32_t)pieces.size(), dbase};
    emu::launch((unsigned)std::max(1, grid), GATHER_WARPS * 32, gather_body, &a);
    memcpy(out, dbase, total);
    return 0;
}


// ---- baseline Huffman ENCODE for the way back (lep_huffenc.cu): the job set-up of lepb200_huffman_encode_resident (henc_launch,
// lep_capi.cu) with host addresses, one launch over all thread-segments.  planes[c] = coefficient plane of component c
// (AlignedBlock layout), bch[c] = blocks per row; out receives im->scan_bytes bytes; seg_status[k] / seg_produced[k] what each
// segment reported.
namespace {
struct HEncArgs { const HEncImage* images; HEncSeg* segs; int nseg; const HEncTable* tables; };
void henc_body(void* p) { const HEncArgs& a = *static_cast<const HEncArgs*>(p); lep_huffencode_kernel(a.images, a.segs, a.nseg, a.tables); }
bool emu_build_enc_table(const lepb200_hufftable& in, HEncTable& t) {      // build_enc_table of lep_capi.cu (build_huffcodes, jpgcoder.cc:5508-5540)
    memset(&t, 0, sizeof(t));
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < in.bits[len]; ++i, ++k, ++code) {
            if (k >= 256 || code >= (1 << len)) return false;
            t.code[in.vals[k]] = (uint16_t)code;
            t.len[in.vals[k]] = (uint8_t)len;
        }
        if (code > (1 << len)) return false;
        code <<= 1;
    }
    return true;
}
}  // namespace

extern "C" int emu_huffman_encode(const lepb200_henc_image* im, int ncmp, int mcuv, const int16_t* const* planes, const int* bch,
                                  uint8_t* out, int32_t* seg_status, uint32_t* seg_produced) {
    if (!im || !planes || !bch || !out || im->scan_bytes == 0 || im->nseg < 1 || im->nseg > LEPB200_MAX_SEGMENTS) return LEPB200_ERR_INVALID;
    HEncImage d;
    memset(&d, 0, sizeof(d));
    std::vector<HEncTable> tabs;
    d.ncmp = ncmp; d.mcuv = mcuv; d.rsti = im->rsti; d.padbit = im->padbit; d.scan_len = im->scan_bytes;
    for (int c = 0; c < ncmp; ++c) {
        d.H[c] = im->H[c]; d.V[c] = im->V[c]; d.bch[c] = bch[c];
        d.plane[c] = (unsigned long long)(uintptr_t)planes[c];
        HEncTable t;
        if (!emu_build_enc_table(im->dc[c], t)) return LEPB200_ERR_INVALID;
        d.dc_tab[c] = (int)tabs.size(); tabs.push_back(t);
        if (!emu_build_enc_table(im->ac[c], t)) return LEPB200_ERR_INVALID;
        d.ac_tab[c] = (int)tabs.size(); tabs.push_back(t);
    }
    d.mcuh = bch[0] / im->H[0];
    std::vector<uint8_t> obuf((size_t)im->scan_bytes + 256 + 512, 0xA5);
    uint8_t* obase = reinterpret_cast<uint8_t*>(align_up((size_t)(uintptr_t)obuf.data(), 256));
    d.out = (unsigned long long)(uintptr_t)obase;
    std::vector<HEncSeg> segs;
    uint32_t off = 0;
    for (int k = 0; k < im->nseg; ++k) {
        HEncSeg sg;
        memset(&sg, 0, sizeof(sg));
        sg.image = 0; sg.my0 = im->seg[k].mcu_row_start; sg.my1 = im->seg[k].mcu_row_end;
        for (int c = 0; c < 3; ++c) sg.lastdc[c] = im->seg[k].last_dc[c];
        sg.ov_bits = im->seg[k].overhang_bits; sg.ov_byte = im->seg[k].overhang_byte;
        sg.out_off = off; sg.expect = im->seg[k].expect_bytes; sg.is_last = k + 1 == im->nseg;
        off += im->seg[k].expect_bytes;
        segs.push_back(sg);
    }
    HEncArgs a{&d, segs.data(), (int)segs.size(), tabs.data()};
    emu::launch((unsigned)((segs.size() + HENC_WARPS - 1) / HENC_WARPS), HENC_WARPS * 32, henc_body, &a);
    memcpy(out, obase, im->scan_bytes);
    for (int k = 0; k < im->nseg; ++k) { if (seg_status) seg_status[k] = segs[k].status; if (seg_produced) seg_produced[k] = segs[k].produced; }
    return 0;
}
