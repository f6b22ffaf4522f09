// emu_decode.cc -- runs the thread-per-segment decode kernels on the CPU, one lane at a time (test infrastructure).
//
// The kernel sources are compiled as host C++ through cuda_shim.h; this file builds the same job descriptors
// lep_capi.cu's build_batch / lepb200_decode_upload build for the device (ImageDesc, SegDesc, order, zeroed model and row
// pools, zeroed planes) with host addresses in place of device addresses, then calls the kernel body once per lane.
// Entry point mirrors lepb200_decode_images (include/lepton_b200.h) plus a kernel selector.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "cuda_shim.h"
#include "../../lepton_b200/csrc/lep_decode_thread.cu"
#include "../../lepton_b200/csrc/lep_decode_lockstep.cu"
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

}  // namespace

// kernel: 1 = lep_decode_thread_kernel, 2 = lep_decode_lockstep_kernel.  Decodes into images[i].planes (zeroed first,
// like the device arena); per-segment status and decision counts come back like lepb200_decode_fetch reports them.
extern "C" int emu_decode_images(int kernel, const lepb200_image* images, int nimages, const lepb200_stream* in,
                                 int32_t* status_out, uint64_t* ndecisions_out) {
    if (nimages <= 0 || !images || !in || (kernel != 1 && kernel != 2)) return LEPB200_ERR_INVALID;
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
            padded.emplace_back((size_t)st.len + 16, 0);
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

    std::vector<uint16_t> models((size_t)nseg * M_TOTAL, 0);       // identity prior = zero fill
    std::vector<uint8_t> rows((size_t)nseg * row_stride, 0);
    const int lanes = 32;
    blockDim.x = lanes; gridDim.x = (unsigned)((nseg + lanes - 1) / lanes);
    // every lane of every launched warp runs, idle ones included (they fill their stripe of the shared tables and must
    // not touch any job)
    for (unsigned b = 0; b < gridDim.x; ++b)
        for (int pass = 0; pass < 2; ++pass)               // pass 0: all lanes with no work (shared-memory tables filled), pass 1: the real run
            for (int l = 0; l < lanes; ++l) {
                blockIdx.x = b; threadIdx.x = (unsigned)l;
                const int count = pass == 0 ? 0 : nseg;
                if (kernel == 1) lep_decode_thread_kernel(descs.data(), segs.data(), 0, count, order.data(), models.data(), rows.data(), row_stride);
                else lep_decode_lockstep_kernel(descs.data(), segs.data(), 0, count, order.data(), models.data(), rows.data(), row_stride);
            }
    for (int s = 0; s < nseg; ++s) {
        if (status_out) status_out[s] = segs[s].status;
        if (ndecisions_out) ndecisions_out[s] = (uint64_t)segs[s].ndecisions_lo | ((uint64_t)segs[s].ndecisions_hi << 32);
    }
    return 0;
}
