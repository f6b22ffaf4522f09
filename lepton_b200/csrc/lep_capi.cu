// lep_capi.cu -- context, device memory management and the C ABI declared in include/lepton_b200.h.
// Single translation unit: the kernels are included so that nvcc sees one module (no -rdc needed).
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/lepton_b200.h"
#include "lep_common.cuh"
#include "lep_predict.cuh"
#include "lep_encode.cu"
#include "lep_decode.cu"
#include "lep_decode_g2.cu"
#include "lep_huffpar.cu"
#include "lep_huffenc.cu"
#include "lep_mux.cu"
#include "lep_host.h"

using namespace lepb200;

namespace {

// zig-zag position of each raster index (reference src/vp8/model/jpeg_meta.hh:13-23)
const uint8_t k_zigzag[64] = {
    0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
// first column of icos_base_8192_scaled (src/vp8/model/jpeg_meta.hh:48-58): entries [i*8]
const int k_icos_base_col0[8] = {8192, 11363, 10703, 9633, 8192, 6436, 4433, 2260};
// src/vp8/model/model.hh:264-274
const uint16_t k_freqmax[64] = {
    1024, 931, 985, 968, 1020, 968, 1020, 1020, 932, 858, 884, 840, 932, 838, 854, 854,
    985, 884, 871, 875, 985, 878, 871, 854, 967, 841, 876, 844, 967, 886, 870, 837,
    1020, 932, 985, 967, 1020, 969, 1020, 1020, 969, 838, 878, 886, 969, 838, 969, 838,
    1020, 854, 871, 870, 1010, 969, 1020, 1020, 1020, 854, 854, 838, 1020, 838, 1020, 838};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 4096;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct lepb200_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
    std::string err;
    DevBuf d_planes, d_streams, d_tokens, d_dense, d_huff, d_hjobs, d_htabs, d_hrows, d_hpar, d_images, d_segs, d_order, d_counter, d_models, d_rows;
    DevBuf d_henc_out, d_henc_imgs, d_henc_segs, d_henc_tabs;
    DevBuf d_gather, d_lit;                // device container assembly: piece list, literal bytes (file headers + trailers)
    HostBuf h_lit;
    DevBuf d_rc_ck, d_rc_digits;           // parallel range coder: checkpoints of the range-only pass, deferred-carry digits
    HostBuf h_henc_out, h_henc_segs, h_henc_desc;
    cudaEvent_t status_ev = nullptr;       // parts mode: the D2H of the decode batch's segment records, queued ahead of the part copies
    bool status_queued = false;
    std::vector<size_t> henc_off;         // per image: offset of its scan bytes in the output buffers (SIZE_MAX = skipped)
    std::vector<int> henc_seg_first;      // per image: index of its first segment record
    int henc_nseg = 0;
    // parts of the last lepb200_huffman_encode_resident_parts call: image range, segment range, byte range of the output,
    // the event behind the part's D2H copies on copy_stream
    struct HEncPart { int i0, i1, s0, s1; size_t b0, b1; cudaEvent_t done; };
    std::vector<HEncPart> henc_parts;
    std::vector<cudaEvent_t> part_events;      // pool: [2k] = part k's kernel, [2k + 1] = part k's copies
    cudaStream_t copy_stream = nullptr;        // D2H copies that run under the kernels of `stream`
    HostBuf h_segs, h_dense, h_stage, h_hjobs, h_hpar;
    size_t resident_plane_total = 0;
    int resident_images = 0;
    std::vector<ImageDesc> images;
    std::vector<SegDesc> segs;
    std::vector<int> order;
    std::vector<size_t> seg_blocks;
    std::vector<size_t> plane_bytes;      // per image*3
    size_t row_stride = 0;
    int grid = 0;
    bool have_batch = false, launched = false, symbolised = false, is_encode = true;
    float last_ms = -1.f, last_ms_a = -1.f, last_ms_huff = -1.f;
    uint64_t launches = 0;
    uint64_t alg_bytes = 0;
    uint64_t coded_blocks = 0;
    int enc_cta_cap = 0;                  // 0 = as many encode CTAs per SM as fit
    int huff_par = 1;                     // 1: images with enough entropy bytes take the many-threads-per-image kernels (lep_huffpar.cu)
    int huff_sub_bits = 0;                // bits per sub-sequence (one thread each); 0 = by batch size (LEPB200_HUFF_SUBSEQ_BITS overrides)
    int huff_par_iters = 0;               // synchronisation iterations of the last batch (diagnostic)
    int huff_warps = 4;                   // images per CTA of the Huffman kernel
    int host_threads = 1;                 // host threads this context may use for staging copies
    int rc_feed = -1;                     // range pass token feed: 1 = cp.async ring in shared memory, 0 = register ring of plain loads, -1 = by
                                          // batch size (measured, profiles/r02_round_l.log: 4096 segments 26.3 ms against 40 ms, 16384
                                          // segments 37.5 against 35.2 ms -- with every SM streaming, the feed is no longer what bounds it); LEPB200_RC_FEED
    int rc_mode = 1;                      // range coder: 1 = range-only pass + parallel pieces + carry pass (lep_rangepass / piece / norm kernels),
                                          // 0 = one serial chain per segment (lep_rangecode_kernel); LEPB200_RC_MODE
    int dec_mode = 0;                     // decode kernel: 0 = by batch size (group kernel when at least dec_group_min segments are in the
                                          // batch, else one warp per segment), 1 = always one warp per segment (lep_decode.cu),
                                          // 2 = always the group kernel (lep_decode_g2.cu)
    int dec_group_min = 10240;            // the group kernel carries 8 serial chains per warp: fewer instructions per decision, but a
                                          // longer latency per chain -- it wins when the chains alone fill the machine (measured: 16 384
                                          // segments 1043 ms against 1620 ms; 4 096 segments 1000 ms against 510 ms)
    int dec_threads_max = 16384;          // group kernel: segments per launch (one zero-filled 1.58 MB model each)
    int dec_lanes = 4;                    // group kernel: lanes per thread-segment, 32 / dec_lanes segments per warp in lock step
    int dec_group_grid = 0;               // group kernel: CTAs of the current batch
    int dec_threads = 0;                  // group kernel: model slots of the current batch (0: the batch uses the warp kernel)
    bool tokens_known = false;            // token streams laid out on the host from caller-supplied bounds (no counting pre-pass)
    unsigned long long token_total = 0;
    bool stage_preuploaded = false;       // the caller pushed the staged scans itself (lepb200_huffman_stage_upload)
};

#define CK(call)                                                                            \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                  \
            return e_ == cudaErrorMemoryAllocation ? LEPB200_ERR_NOMEM : LEPB200_ERR_CUDA;  \
        }                                                                                   \
    } while (0)

namespace {

// ProbabilityTablesBase::set_quantization_table (src/vp8/model/model.hh:247-290)
int fill_quant(ImageDesc& d, int c, const uint16_t zz[64], bool check_zero) {
    uint16_t* q = d.q[c];
    for (int i = 0; i < 64; ++i) q[i] = zz[k_zigzag[i]];
    for (int r = 0; r < 8; ++r) {
        for (int i = 0; i < 8; ++i) {
            d.icos_x[c][r * 8 + i] = k_icos_base_col0[i] * (int)q[i * 8 + r];
            d.icos_y[c][r * 8 + i] = k_icos_base_col0[i] * (int)q[r * 8 + i];
        }
        if (d.icos_x[c][r * 8] == 0 || d.icos_y[c][r * 8] == 0) {
            if (check_zero) return LEPB200_ST_UNSUPPORTED_JPEG_WITH_ZERO_IDCT_0;
            // decode side (filetype == LEPTON, model.hh:257): the reference goes on and would divide by zero; refuse.
            return LEPB200_ST_UNSUPPORTED_JPEG_WITH_ZERO_IDCT_0;
        }
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

// Rows of component c coded by a segment [min_y, max_y) (row iteration of lepton_codec.hh:41-100).
size_t segment_blocks(const lepb200_image& im, int min_y, int max_y, bool last) {
    size_t n = 0;
    int v0 = im.bcv[0] / im.mcuv;
    for (int c = 0; c < im.ncmp; ++c) {
        int mult = im.bcv[c] / im.mcuv;
        long y0 = (long)(min_y / std::max(v0, 1)) * mult, y1 = last ? im.trunc_bcv[c] : (long)((max_y + v0 - 1) / std::max(v0, 1)) * mult;
        y1 = std::min<long>(y1, im.trunc_bcv[c]);
        if (y1 > y0) n += (size_t)(y1 - y0) * im.bch[c];
    }
    return n;
}

int validate_image(lepb200_ctx* ctx, const lepb200_image& im) {
    if (im.ncmp < 1 || im.ncmp > 3 || im.mcuv <= 0 || im.nseg < 1 || im.nseg > LEPB200_MAX_SEGMENTS) {
        ctx->err = "invalid image descriptor (ncmp/mcuv/nseg)";
        return LEPB200_ERR_INVALID;
    }
    for (int c = 0; c < im.ncmp; ++c) {
        if (im.bch[c] <= 0 || im.bcv[c] <= 0 || im.bcv[c] % im.mcuv || im.trunc_bcv[c] < 0 || im.trunc_bcv[c] > im.bcv[c] ||
            im.trunc_bc[c] < 0 || (long long)im.trunc_bc[c] > (long long)im.bch[c] * im.bcv[c] || !im.planes[c]) {
            ctx->err = "invalid component geometry";
            return LEPB200_ERR_INVALID;
        }
    }
    for (int s = 0; s + 1 < im.nseg; ++s)
        if (im.luma_y_start[s] > im.luma_y_start[s + 1]) { ctx->err = "segment starts must be non-decreasing"; return LEPB200_ERR_INVALID; }
    return 0;
}

// lep_decode_g2_kernel<G>: launch shape (warps per CTA, thread-segments per warp) and resident CTAs per SM
template <int G> int group_ctas_per_sm() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lep_decode_g2_kernel<G>, G2Cfg<G>::THREADS, 0) != cudaSuccess) n = 1;
    return std::max(n, 1);
}
void group_launch_shape(int lanes, int& warps, int& per_warp, int& ctas_per_sm) {
    per_warp = 32 / lanes;
    switch (lanes) {
    case 8: warps = G2Cfg<8>::WARPS; ctas_per_sm = group_ctas_per_sm<8>(); break;
    case 32: warps = G2Cfg<32>::WARPS; ctas_per_sm = group_ctas_per_sm<32>(); break;
    default: warps = G2Cfg<4>::WARPS; ctas_per_sm = group_ctas_per_sm<4>(); break;
    }
}
template <int G> void launch_group_kernel(int grid, cudaStream_t st, const ImageDesc* images, SegDesc* segs, int first, int count, const int* order,
                                          int* counter, uint16_t* models, uint8_t* rows, size_t row_stride) {
    lep_decode_g2_kernel<G><<<grid, G2Cfg<G>::THREADS, 0, st>>>(images, segs, first, count, order, counter, models, rows, row_stride);
}

// Common part of encode/decode upload: job tables, pools, plane arena layout.
int build_batch(lepb200_ctx* ctx, const lepb200_image* images, int nimages, bool encode, const lepb200_stream* in) {
    ctx->have_batch = false; ctx->launched = false; ctx->is_encode = encode;
    if (nimages <= 0 || !images) { ctx->err = "empty batch"; return LEPB200_ERR_INVALID; }
    ctx->images.assign(nimages, ImageDesc());
    ctx->segs.clear(); ctx->seg_blocks.clear(); ctx->plane_bytes.assign((size_t)nimages * 3, 0);
    size_t plane_total = 0, stream_total = 0, row_stride = 0;
    unsigned long long token_total = 0;   // token arena layout when every segment comes with a caller-supplied bound
    bool tokens_known = encode;
    int sidx = 0;
    for (int i = 0; i < nimages; ++i) {
        const lepb200_image& im = images[i];
        int v = validate_image(ctx, im);
        if (v) return v;
        ImageDesc& d = ctx->images[i];
        memset(&d, 0, sizeof(d));
        d.ncmp = im.ncmp; d.mcuv = im.mcuv;
        int qstatus = 0;
        size_t rs = 0;
        for (int c = 0; c < im.ncmp; ++c) {
            d.bch[c] = im.bch[c]; d.bcv[c] = im.bcv[c]; d.trunc_bcv[c] = im.trunc_bcv[c]; d.trunc_bc[c] = im.trunc_bc[c];
            d.mult[c] = im.bcv[c] / im.mcuv;
            int qs = fill_quant(d, c, im.qtable_zigzag[c], encode);
            if (qs) qstatus = qs;
            size_t pb = (size_t)im.bch[c] * im.bcv[c] * 128;
            ctx->plane_bytes[(size_t)i * 3 + c] = pb;
            d.plane[c] = plane_total;                     // offset for now; rebased below
            plane_total += align_up(pb, 256);
            rs += (size_t)im.bch[c] * 16 + align_up((size_t)im.bch[c], 16);
        }
        row_stride = std::max(row_stride, align_up(rs, 256));
        for (int s = 0; s < im.nseg; ++s, ++sidx) {
            SegDesc sd;
            memset(&sd, 0, sizeof(sd));
            sd.image = i;
            sd.min_y = im.luma_y_start[s];
            sd.is_last = s + 1 == im.nseg;
            sd.max_y = sd.is_last ? im.bcv[0] : im.luma_y_start[s + 1];
            sd.status = qstatus;
            size_t nb = segment_blocks(im, sd.min_y, sd.max_y, sd.is_last);
            ctx->seg_blocks.push_back(nb);
            if (encode) {
                size_t cap = align_up(nb * 64 + 4096, 256);   // 64 B/block is > 1.5x what q=100 photos need; overflow is reported, never silent
                sd.stream = stream_total; sd.cap = (uint32_t)cap;
                stream_total += cap;
                sd.tokens = 0; sd.tok_cap = 0;                    // assigned on the device by the counting pre-pass ...
                if (im.seg_token_bound[s]) {                      // ... unless the caller knows a bound (GPU Huffman decoder)
                    const unsigned long long tcap = ((unsigned long long)im.seg_token_bound[s] + 64 + 63) & ~63ull;
                    sd.tok_cap = tcap > 0xffffff00ull ? 0xffffff00u : (uint32_t)tcap;
                    sd.tokens = token_total;
                    token_total += sd.tok_cap;
                } else {
                    tokens_known = false;
                }
            } else {
                sd.stream = stream_total; sd.cap = (uint32_t)in[sidx].len;
                stream_total += align_up((size_t)in[sidx].len + 16, 16);
            }
            ctx->segs.push_back(sd);
        }
    }
    const int nseg = (int)ctx->segs.size();
    // largest segments first (longest-processing-time-first on the persistent warps)
    ctx->order.resize(nseg);
    for (int i = 0; i < nseg; ++i) ctx->order[i] = i;
    std::stable_sort(ctx->order.begin(), ctx->order.end(), [&](int a, int b) { return ctx->seg_blocks[a] > ctx->seg_blocks[b]; });

    // persistent grid: as many CTAs as can be resident, but no more warps than segments
    int per_sm = 0;
    if (encode) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lep_encode_kernel, ENC_WARPS_PER_CTA * 32, 0));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lep_decode_kernel, DEC_WARPS_PER_CTA * 32, 0));
    if (encode && ctx->enc_cta_cap > 0) per_sm = std::min(per_sm, ctx->enc_cta_cap);
    const int wpc = encode ? ENC_WARPS_PER_CTA : DEC_WARPS_PER_CTA;
    int grid = std::max(1, std::min(per_sm * ctx->sm_count, (nseg + wpc - 1) / wpc));
    ctx->grid = grid;
    ctx->row_stride = row_stride;
    ctx->tokens_known = tokens_known;
    ctx->token_total = token_total;

    CK(ctx->d_planes.reserve(plane_total));
    CK(ctx->d_streams.reserve(stream_total + 256));
    CK(ctx->d_images.reserve(sizeof(ImageDesc) * nimages));
    CK(ctx->d_segs.reserve(sizeof(SegDesc) * nseg));
    CK(ctx->d_order.reserve(sizeof(int) * nseg));
    CK(ctx->d_counter.reserve(256));
    ctx->dec_threads = 0;
    const bool use_group = !encode && (ctx->dec_mode == 2 || (ctx->dec_mode == 0 && nseg >= ctx->dec_group_min));
    if (use_group) {
        // group kernel: one zero-filled model per segment of a launch, one row buffer per resident group
        ctx->dec_threads = std::max(1, std::min(nseg, ctx->dec_threads_max));
        int warps = 0, per_warp = 0, gsm = 0;
        group_launch_shape(ctx->dec_lanes, warps, per_warp, gsm);
        const int per_cta = warps * per_warp;
        ctx->dec_group_grid = std::max(1, std::min(gsm * ctx->sm_count, (ctx->dec_threads + per_cta - 1) / per_cta));
        CK(ctx->d_models.reserve((size_t)ctx->dec_threads * MODEL_BYTES));
        CK(ctx->d_rows.reserve((size_t)ctx->dec_group_grid * per_cta * row_stride));
    } else {
        CK(ctx->d_models.reserve((size_t)grid * wpc * MODEL_BYTES));
        CK(ctx->d_rows.reserve((size_t)grid * wpc * row_stride));
    }
    for (int i = 0; i < nimages; ++i)
        for (int c = 0; c < ctx->images[i].ncmp; ++c) ctx->images[i].plane[c] += (unsigned long long)(uintptr_t)ctx->d_planes.p;
    for (auto& sd : ctx->segs) sd.stream += (unsigned long long)(uintptr_t)ctx->d_streams.p;
    CK(cudaMemcpyAsync(ctx->d_images.p, ctx->images.data(), sizeof(ImageDesc) * nimages, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_segs.p, ctx->segs.data(), sizeof(SegDesc) * nseg, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_order.p, ctx->order.data(), sizeof(int) * nseg, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

__global__ void lep_compact_kernel(const SegDesc* __restrict__ segs, const unsigned long long* __restrict__ dst_off, uint8_t* __restrict__ dense, int nseg) {
    // one CTA per segment: copy the produced stream bytes into the dense output buffer (16-byte body, byte tails)
    const int s = blockIdx.x;
    if (s >= nseg) return;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(segs[s].stream);
    uint8_t* dst = dense + dst_off[s];
    const uint32_t n = segs[s].status == 0 ? segs[s].len : 0;
    const uint32_t n16 = n / 16;           // src is 256-byte aligned, dst offsets are 16-byte aligned
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = s4[i];
    for (uint32_t i = n16 * 16 + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace

extern "C" {

static int encode_prepass(lepb200_ctx* ctx);

int lepb200_device_available(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

size_t lepb200_model_bytes(void) { return MODEL_BYTES; }

int lepb200_create(lepb200_ctx** out, int device) {
    if (!out) return LEPB200_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return LEPB200_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return LEPB200_ERR_INVALID;
    lepb200_ctx* ctx = new lepb200_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return LEPB200_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return LEPB200_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess ||
        cudaEventCreate(&ctx->ev1) != cudaSuccess || cudaEventCreate(&ctx->ev_mid) != cudaSuccess) {
        delete ctx;
        return LEPB200_ERR_CUDA;
    }
    if (cudaFuncSetAttribute(lep_rangepass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCT_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(lep_rangepass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCT_SMEM_TABLE_BYTES) != cudaSuccess) {
        delete ctx;
        return LEPB200_ERR_CUDA;
    }
    if (const char* e = getenv("LEPB200_ENC_CTA_CAP")) ctx->enc_cta_cap = atoi(e);          // tuning overrides
    if (const char* e = getenv("LEPB200_HUFF_WARPS")) ctx->huff_warps = atoi(e);
    if (const char* e = getenv("LEPB200_HUFF_PAR")) ctx->huff_par = atoi(e);
    if (const char* e = getenv("LEPB200_HUFF_SUBSEQ_BITS")) ctx->huff_sub_bits = std::max(256, std::min(1 << 20, atoi(e))) & ~31;
    if (const char* e = getenv("LEPB200_DEC_MODE")) ctx->dec_mode = atoi(e);
    if (const char* e = getenv("LEPB200_RC_MODE")) ctx->rc_mode = atoi(e);
    if (const char* e = getenv("LEPB200_RC_FEED")) ctx->rc_feed = atoi(e);
    if (const char* e = getenv("LEPB200_DEC_THREADS")) ctx->dec_threads_max = std::max(32, atoi(e));
    if (const char* e = getenv("LEPB200_DEC_GROUP_MIN")) ctx->dec_group_min = std::max(1, atoi(e));
    if (const char* e = getenv("LEPB200_DEC_LANES")) {
        const int g = atoi(e);
        if (g == 4 || g == 8 || g == 32) ctx->dec_lanes = g;
    }
    *out = ctx;
    return LEPB200_OK;
}

void lepb200_set_encode_ctas_per_sm(lepb200_ctx* ctx, int n) { if (ctx && !getenv("LEPB200_ENC_CTA_CAP")) ctx->enc_cta_cap = n; }
void lepb200_set_host_threads(lepb200_ctx* ctx, int n) { if (ctx && n > 0) ctx->host_threads = n; }
void lepb200_set_huffman_warps_per_cta(lepb200_ctx* ctx, int n) { if (ctx && !getenv("LEPB200_HUFF_WARPS")) ctx->huff_warps = n; }

void lepb200_destroy(lepb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf* b : {&ctx->d_henc_out, &ctx->d_henc_imgs, &ctx->d_henc_segs, &ctx->d_henc_tabs, &ctx->d_huff, &ctx->d_hjobs, &ctx->d_htabs, &ctx->d_hrows, &ctx->d_hpar, &ctx->d_planes, &ctx->d_streams, &ctx->d_tokens, &ctx->d_dense, &ctx->d_images, &ctx->d_segs, &ctx->d_order, &ctx->d_counter, &ctx->d_models, &ctx->d_rows, &ctx->d_rc_ck, &ctx->d_rc_digits, &ctx->d_gather, &ctx->d_lit})
        b->release();
    for (HostBuf* b : {&ctx->h_segs, &ctx->h_dense, &ctx->h_stage, &ctx->h_hjobs, &ctx->h_hpar}) b->release();
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->ev_mid);
    for (cudaEvent_t e : ctx->part_events) cudaEventDestroy(e);
    if (ctx->status_ev) cudaEventDestroy(ctx->status_ev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* lepb200_last_error(const lepb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int lepb200_sync(lepb200_ctx* ctx) {
    if (!ctx) return LEPB200_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return LEPB200_OK;
}
float lepb200_last_kernel_ms(lepb200_ctx* ctx) {
    if (!ctx) return -1.f;
    if (ctx->launched) {
        cudaSetDevice(ctx->device);
        float ms = -1.f;
        if (cudaEventSynchronize(ctx->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ms = ms;
        if (ctx->is_encode && cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev_mid) == cudaSuccess) ctx->last_ms_a = ms;
    }
    return ctx->last_ms;
}
float lepb200_last_symbolise_ms(lepb200_ctx* ctx) {
    if (!ctx) return -1.f;
    lepb200_last_kernel_ms(ctx);
    return ctx->last_ms_a;
}
float lepb200_last_huffman_ms(lepb200_ctx* ctx) { return ctx ? ctx->last_ms_huff : -1.f; }
int lepb200_last_huffman_iterations(lepb200_ctx* ctx) { return ctx ? ctx->huff_par_iters : 0; }
uint64_t lepb200_kernel_launches(const lepb200_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t lepb200_last_algorithmic_bytes(const lepb200_ctx* ctx) { return ctx ? ctx->alg_bytes : 0; }

void* lepb200_pinned_alloc(size_t bytes) {
    void* p = nullptr;
    return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
}
void lepb200_pinned_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------------------ encode
int lepb200_encode_upload(lepb200_ctx* ctx, const lepb200_image* images, int nimages) {
    if (!ctx) return LEPB200_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int r = build_batch(ctx, images, nimages, true, nullptr);
    if (r) return r;
    for (int i = 0; i < nimages; ++i)
        for (int c = 0; c < images[i].ncmp; ++c)
            CK(cudaMemcpyAsync(reinterpret_cast<void*>(ctx->images[i].plane[c]), images[i].planes[c], ctx->plane_bytes[(size_t)i * 3 + c],
                               cudaMemcpyHostToDevice, ctx->stream));
    return encode_prepass(ctx);
}

static int encode_prepass(lepb200_ctx* ctx) {
    if (ctx->tokens_known) {              // bounds came with the images (GPU Huffman decoder): nothing to count, nothing to wait for
        CK(ctx->d_tokens.reserve((size_t)ctx->token_total * 2 + 256));
        ctx->have_batch = true;
        return LEPB200_OK;
    }
    // pre-pass: per-segment token upper bounds -> exact-fit token arena (sizes depend on the data, so one sync here)
    const int nseg = (int)ctx->segs.size();
    lep_count_kernel<<<nseg, CNT_THREADS, 0, ctx->stream>>>(static_cast<const ImageDesc*>(ctx->d_images.p), static_cast<SegDesc*>(ctx->d_segs.p), nseg);
    CK(cudaGetLastError());
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ctx->d_counter.p) + 64);
    lep_token_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(static_cast<SegDesc*>(ctx->d_segs.p), nseg, d_total);
    CK(cudaGetLastError());
    ctx->launches += 2;
    unsigned long long total_tokens = 0;
    CK(cudaMemcpyAsync(&total_tokens, d_total, sizeof(total_tokens), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(ctx->d_tokens.reserve((size_t)total_tokens * 2 + 256));
    ctx->token_total = total_tokens;
    ctx->have_batch = true;
    return LEPB200_OK;
}

// ------------------------------------------------------------------------------------------------ GPU Huffman decode
static bool build_table_dev(const lepb200_hufftable& in, HuffTableDev& t) { return huff_build_table(in.bits, in.vals, t); }

uint8_t* lepb200_huffman_stage_reserve(lepb200_ctx* ctx, size_t bytes) {
    if (!ctx) return nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    if (ctx->h_stage.reserve(bytes + 256) != cudaSuccess) { ctx->err = "pinned staging allocation failed"; return nullptr; }
    if (ctx->d_huff.reserve(bytes + 512) != cudaSuccess) { ctx->err = "device staging allocation failed"; return nullptr; }
    ctx->stage_preuploaded = false;
    return static_cast<uint8_t*>(ctx->h_stage.p);
}

// Asynchronous H2D of one staged range (callable from several host threads while others are still parsing); once used,
// the following lepb200_huffman_decode_to_device does not copy the staging buffer again, so ALL scans must be pushed.
int lepb200_huffman_stage_upload(lepb200_ctx* ctx, size_t offset, size_t bytes) {
    if (!ctx || !ctx->h_stage.p || offset + bytes > ctx->h_stage.cap || offset + bytes > ctx->d_huff.cap) return LEPB200_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LEPB200_ERR_CUDA;
    ctx->stage_preuploaded = true;
    if (cudaMemcpyAsync(static_cast<uint8_t*>(ctx->d_huff.p) + offset, static_cast<uint8_t*>(ctx->h_stage.p) + offset, bytes,
                        cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return LEPB200_ERR_CUDA;
    return LEPB200_OK;
}

int lepb200_huffman_decode_to_device(lepb200_ctx* ctx, lepb200_jpeg_scan* scans, int n) {
    if (!ctx || !scans || n <= 0) return LEPB200_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->have_batch = false; ctx->launched = false; ctx->resident_images = 0;
    std::vector<HuffJob> jobs(n);
    std::vector<HuffTableDev> tabs;
    std::vector<std::pair<const lepb200_hufftable*, int>> seen;     // dedupe identical tables (most files share the standard ones)
    std::vector<char> tab_ok;                                       // build result per table: a cache hit on a table that failed to build fails too
    auto table_index = [&](const lepb200_hufftable& t, bool& ok) -> int {
        for (auto& s : seen) if (!memcmp(s.first, &t, sizeof(t))) { if (!tab_ok[s.second]) ok = false; return s.second; }
        HuffTableDev d;
        const bool built = build_table_dev(t, d);
        if (!built) ok = false;
        tabs.push_back(d);
        tab_ok.push_back(built ? 1 : 0);
        seen.emplace_back(&t, (int)tabs.size() - 1);
        return (int)tabs.size() - 1;
    };
    size_t plane_total = 0, huff_total = 0, rows_total = 0;
    uint32_t sub_total = 0;
    std::vector<uint32_t> sub_base((size_t)n + 1, 0);
    // in-place mode: the caller de-stuffed straight into this context's pinned staging buffer
    // (lepb200_huffman_stage_reserve), 16-byte aligned with >= 16 spare bytes after each scan -> no gather copy
    const uint8_t* stage0 = static_cast<const uint8_t*>(ctx->h_stage.p);
    int inside = 0, placeholders = 0;
    for (int i = 0; i < n; ++i)
        if (scans[i].entropy == nullptr) ++placeholders;
        else if (stage0 && scans[i].entropy >= stage0 && scans[i].entropy + scans[i].nbytes + 16 <= stage0 + ctx->h_stage.cap) ++inside;
    if (inside != 0 && inside != n - placeholders) { ctx->err = "huffman_decode_to_device: scans partly inside the staging buffer"; return LEPB200_ERR_INVALID; }
    const bool in_place = inside > 0 && inside == n - placeholders;
    for (int i = 0; i < n; ++i) if (in_place && scans[i].entropy && ((scans[i].entropy - stage0) & 15)) { ctx->err = "huffman_decode_to_device: staged scan not 16-byte aligned"; return LEPB200_ERR_INVALID; }
    // sub-sequence length of the many-threads-per-image kernels: longer sub-sequences need fewer synchronisation passes
    // (8192 bits: 3, 4096: 5, 2048: 9 on 1080p files), as long as the batch still gives every SM its 2048 threads
    int sub_bits = ctx->huff_sub_bits;
    if (sub_bits <= 0) {
        uint64_t bits = 0;
        for (int i = 0; i < n; ++i) if (scans[i].entropy && scans[i].ncmp > 1 && scans[i].rsti == 0) bits += (uint64_t)scans[i].nbytes * 8;
        sub_bits = 2048;
        for (int cand : {16384, 8192, 4096}) if (bits / (uint64_t)cand >= (uint64_t)ctx->sm_count * 2048) { sub_bits = cand; break; }
    }
    for (int i = 0; i < n; ++i) {
        lepb200_jpeg_scan& sc = scans[i];
        HuffJob& jb = jobs[i];
        memset(&jb, 0, sizeof(jb));
        const bool placeholder = sc.entropy == nullptr;
        bool ok = sc.ncmp >= 1 && sc.ncmp <= 3 && sc.mcuh > 0 && sc.mcuv > 0 && (placeholder || sc.rows);
        jb.ncmp = sc.ncmp; jb.mcuh = sc.mcuh; jb.mcuv = sc.mcuv; jb.rsti = sc.rsti; jb.nbytes = sc.nbytes;
        for (int c = 0; ok && c < sc.ncmp; ++c) {
            jb.H[c] = sc.H[c]; jb.V[c] = sc.V[c];
            ok = ok && sc.H[c] >= 1 && sc.H[c] <= 2 && sc.V[c] >= 1 && sc.V[c] <= 2;
            jb.bch[c] = sc.mcuh * sc.H[c]; jb.bcv[c] = sc.mcuv * sc.V[c];
            jb.nch[c] = sc.nch[c]; jb.ncv[c] = sc.ncv[c];
            if (!placeholder) { jb.dc_tab[c] = table_index(sc.dc[c], ok); jb.ac_tab[c] = table_index(sc.ac[c], ok); }
            jb.plane[c] = plane_total;
            plane_total += align_up((size_t)jb.bch[c] * jb.bcv[c] * 128, 256);
        }
        jb.status = ok ? (placeholder ? HUFF_JOB_SKIP : 0) : LEPB200_ST_NOT_HANDLED;
        if (placeholder) {
            jb.huff = 0; jb.nbytes = 0;
        } else if (in_place) {
            jb.huff = (size_t)(sc.entropy - stage0);
            huff_total = std::max(huff_total, align_up((size_t)jb.huff + sc.nbytes + 16, 16));
        } else {
            jb.huff = huff_total;
            huff_total += align_up((size_t)sc.nbytes + 16, 16);
        }
        jb.rows = rows_total;
        rows_total += (size_t)(sc.mcuv + 1) * sizeof(HuffRow);
        // interleaved scans without restart intervals and with enough data take the many-threads-per-image kernels
        const uint64_t bits = (uint64_t)sc.nbytes * 8;
        jb.sub_base = sub_total;
        if (ctx->huff_par && jb.status == 0 && sc.ncmp > 1 && sc.rsti == 0 && bits >= 4ull * (uint64_t)sub_bits && bits < (1ull << 32))
            jb.nsub = (uint32_t)((bits + (uint64_t)sub_bits - 1) / (uint64_t)sub_bits);
        sub_total += jb.nsub;
        sub_base[i] = jb.sub_base;
    }
    sub_base[n] = sub_total;
    CK(ctx->d_planes.reserve(plane_total + 256));
    CK(ctx->d_huff.reserve(huff_total + 256));
    CK(ctx->d_hrows.reserve(rows_total + 256));
    CK(ctx->d_htabs.reserve(sizeof(HuffTableDev) * std::max<size_t>(1, tabs.size())));
    if (!in_place) CK(ctx->h_stage.reserve(huff_total + 256));
    uint8_t* hs = static_cast<uint8_t*>(ctx->h_stage.p);
    if (in_place) {
        if (!ctx->stage_preuploaded)          // (a caller that uploads itself has zeroed the 16 bytes behind each scan)
            for (int i = 0; i < n; ++i) if (scans[i].entropy) memset(hs + jobs[i].huff + scans[i].nbytes, 0, 16);   // the decoder reads whole words past the end
    } else {
        // gather the de-stuffed scans into the pinned staging buffer (hundreds of MB per chunk): split over host threads
        const int nt = std::max(1, std::min(ctx->host_threads, n));
        auto copy_range = [&](int t) {
            for (int i = t; i < n; i += nt) {
                const HuffJob& jb = jobs[i];
                if (!scans[i].entropy) continue;
                memcpy(hs + jb.huff, scans[i].entropy, scans[i].nbytes);
                memset(hs + jb.huff + scans[i].nbytes, 0, align_up((size_t)scans[i].nbytes + 16, 16) - scans[i].nbytes);
            }
        };
        if (nt == 1) copy_range(0);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back(copy_range, t);
            for (auto& t : th) t.join();
        }
    }
    for (int i = 0; i < n; ++i) {
        HuffJob& jb = jobs[i];
        jb.huff += (unsigned long long)(uintptr_t)ctx->d_huff.p;
        jb.rows += (unsigned long long)(uintptr_t)ctx->d_hrows.p;
        for (int c = 0; c < jb.ncmp; ++c) jb.plane[c] += (unsigned long long)(uintptr_t)ctx->d_planes.p;
    }
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;          // one warp per image: no geometry grouping needed
    const std::vector<HuffJob>& sorted = jobs;
    CK(ctx->d_hjobs.reserve(align_up(sizeof(HuffJob) * n, 256)));
    CK(cudaMemsetAsync(ctx->d_planes.p, 0, plane_total, ctx->stream));
    if (!(in_place && ctx->stage_preuploaded)) CK(cudaMemcpyAsync(ctx->d_huff.p, hs, huff_total, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stage_preuploaded = false;
    CK(cudaMemcpyAsync(ctx->d_hjobs.p, sorted.data(), sizeof(HuffJob) * n, cudaMemcpyHostToDevice, ctx->stream));

    if (!tabs.empty()) CK(cudaMemcpyAsync(ctx->d_htabs.p, tabs.data(), sizeof(HuffTableDev) * tabs.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    // CTA width: 7 images per CTA puts a 1024-image chunk on one CTA per SM, whose registers fit next to the encode
    // kernel of the previous chunk when that one is capped (lepb200_set_encode_ctas_per_sm), so the two overlap
    ctx->huff_par_iters = 0;
    if (sub_total > 0) {
        // sub-sequence kernels (lep_huffpar.cu): arrays of 32 bytes per sub-sequence, then sub_base and the dirty counters
        constexpr int ITER_CAP = 62;
        const size_t o_exit = 0, o_cnt = align_up((size_t)sub_total * 8, 256), o_tok = o_cnt + align_up((size_t)sub_total * 16, 256),
                     o_epoch = o_tok + align_up((size_t)sub_total * 4, 256), o_dirty = o_epoch + align_up((size_t)sub_total * 4, 256),
                     o_base = o_dirty + 256, o_end = o_base + align_up(((size_t)n + 1) * 4, 256);
        CK(ctx->d_hpar.reserve(o_end));
        uint8_t* hp = static_cast<uint8_t*>(ctx->d_hpar.p);
        CK(cudaMemsetAsync(hp + o_epoch, 0, o_base - o_epoch, ctx->stream));
        CK(ctx->h_hpar.reserve(align_up(((size_t)n + 1) * 4, 256) + 256));
        uint32_t* h_base = static_cast<uint32_t*>(ctx->h_hpar.p);
        unsigned int* h_dirty = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(ctx->h_hpar.p) + align_up(((size_t)n + 1) * 4, 256));
        memcpy(h_base, sub_base.data(), ((size_t)n + 1) * 4);
        CK(cudaMemcpyAsync(hp + o_base, h_base, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        HpArrays a;
        a.exit = reinterpret_cast<unsigned long long*>(hp + o_exit); a.cnt = reinterpret_cast<uint4*>(hp + o_cnt);
        a.tok = reinterpret_cast<uint32_t*>(hp + o_tok); a.epoch = reinterpret_cast<uint32_t*>(hp + o_epoch);
        a.sub_base = reinterpret_cast<const uint32_t*>(hp + o_base); a.total = sub_total; a.sub_bits = (uint32_t)sub_bits;
        unsigned int* d_dirty = reinterpret_cast<unsigned int*>(hp + o_dirty);
        HuffJob* dj = static_cast<HuffJob*>(ctx->d_hjobs.p);
        const HuffTableDev* dt = static_cast<const HuffTableDev*>(ctx->d_htabs.p);
        const unsigned grid = (sub_total + HP_THREADS - 1) / HP_THREADS;
        int iter = 0;
        auto sync_iter = [&]() {
            lep_huffpar_sync_kernel<<<grid, HP_THREADS, 0, ctx->stream>>>(dj, n, dt, (int)tabs.size(), a, iter, ITER_CAP, d_dirty);
            ++iter; ctx->launches += 1;
        };
        sync_iter(); sync_iter(); sync_iter();
        for (;;) {              // until no sub-sequence has to run again (dirty[k] = sub-sequences that iteration k has to decode)
            CK(cudaMemcpyAsync(h_dirty, d_dirty, 64 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (iter > ITER_CAP || h_dirty[iter] == 0) break;
            sync_iter();
            if (iter <= ITER_CAP) sync_iter();
        }
        ctx->huff_par_iters = iter;
        lep_huffpar_prefix_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(dj, n, a);
        lep_huffpar_write_kernel<<<grid, HP_THREADS, 0, ctx->stream>>>(dj, n, dt, (int)tabs.size(), a);
        ctx->launches += 2;
        CK(cudaGetLastError());
    }
    // the serial walk: every image the kernels above did not take or did not finish cleanly
    const int hw = std::max(1, std::min(HUFF_MAX_WARPS, ctx->huff_warps));
    lep_huffdecode_kernel<<<(n + hw - 1) / hw, hw * 32, 0, ctx->stream>>>(
        static_cast<HuffJob*>(ctx->d_hjobs.p), n, static_cast<const HuffTableDev*>(ctx->d_htabs.p), (int)tabs.size());
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_mid, ctx->stream));
    ctx->launches += 1;
    CK(ctx->h_hjobs.reserve(sizeof(HuffJob) * n + rows_total));
    HuffJob* hj = static_cast<HuffJob*>(ctx->h_hjobs.p);
    uint8_t* hrows = reinterpret_cast<uint8_t*>(hj + n);
    CK(cudaMemcpyAsync(hj, ctx->d_hjobs.p, sizeof(HuffJob) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hrows, ctx->d_hrows.p, rows_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms_huff, ctx->ev0, ctx->ev_mid);
    for (int k = 0; k < n; ++k) {
        const int i = perm[k];
        scans[i].status = hj[k].status == HUFF_JOB_SKIP ? 0 : hj[k].status;
        scans[i].padbit = hj[k].padbit;
        scans[i].end_bitpos = hj[k].end_bitpos;
        scans[i].nrows = hj[k].nrows;
        const size_t off = (size_t)(jobs[i].rows - (unsigned long long)(uintptr_t)ctx->d_hrows.p);
        static_assert(sizeof(HuffRow) == sizeof(lepb200_huffrow), "row record layout");
        if (hj[k].nrows > 0) memcpy(scans[i].rows, hrows + off, sizeof(HuffRow) * (size_t)std::min(hj[k].nrows, scans[i].mcuv + 1));
    }
    ctx->resident_plane_total = plane_total;
    ctx->resident_images = n;
    return LEPB200_OK;
}

int lepb200_encode_upload_resident(lepb200_ctx* ctx, const lepb200_image* images, int nimages) {
    if (!ctx) return LEPB200_ERR_INVALID;
    if (ctx->resident_images != nimages) { ctx->err = "encode_upload_resident: no matching huffman_decode_to_device batch"; return LEPB200_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    // resident images carry no host planes (validate_image wants non-null pointers); placeholder images do
    std::vector<lepb200_image> tmp(images, images + nimages);
    for (auto& im : tmp) for (int c = 0; c < im.ncmp && c < 3; ++c) if (!im.planes[c]) im.planes[c] = reinterpret_cast<int16_t*>(uintptr_t(1));
    void* const planes_before = ctx->d_planes.p;
    int r = build_batch(ctx, tmp.data(), nimages, true, nullptr);
    if (r) return r;
    if (ctx->d_planes.p != planes_before) { ctx->err = "encode_upload_resident: plane arena moved (geometry mismatch)"; return LEPB200_ERR_INVALID; }
    for (int i = 0; i < nimages; ++i)
        for (int c = 0; c < images[i].ncmp && c < 3; ++c)
            if (images[i].planes[c])
                CK(cudaMemcpyAsync(reinterpret_cast<void*>(ctx->images[i].plane[c]), images[i].planes[c], ctx->plane_bytes[(size_t)i * 3 + c],
                                   cudaMemcpyHostToDevice, ctx->stream));
    ctx->resident_images = 0;
    return encode_prepass(ctx);
}

int lepb200_encode_launch_symbolise(lepb200_ctx* ctx) {
    if (!ctx || !ctx->have_batch || !ctx->is_encode) { if (ctx) ctx->err = "encode_launch without encode_upload"; return LEPB200_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const int nseg = (int)ctx->segs.size();
    CK(cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(int), ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    lep_encode_kernel<<<ctx->grid, ENC_WARPS_PER_CTA * 32, 0, ctx->stream>>>(
        static_cast<const ImageDesc*>(ctx->d_images.p), static_cast<SegDesc*>(ctx->d_segs.p), nseg, static_cast<const int*>(ctx->d_order.p),
        static_cast<int*>(ctx->d_counter.p), static_cast<uint16_t*>(ctx->d_models.p), static_cast<uint8_t*>(ctx->d_rows.p), ctx->row_stride,
        static_cast<uint16_t*>(ctx->d_tokens.p));
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_mid, ctx->stream));
    ctx->launches += 1;
    ctx->symbolised = true;
    return LEPB200_OK;
}

int lepb200_encode_launch_rangecode(lepb200_ctx* ctx) {
    if (!ctx || !ctx->symbolised || !ctx->is_encode) { if (ctx) ctx->err = "encode_launch_rangecode without encode_launch_symbolise"; return LEPB200_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const int nseg = (int)ctx->segs.size();
    if (ctx->rc_mode == 1) {
        // range-only pass -> digit layout (one small D2H + sync: the arena size depends on the data) -> parallel pieces -> carries
        SegDesc* ds = static_cast<SegDesc*>(ctx->d_segs.p);
        const int* dord = static_cast<const int*>(ctx->d_order.p);
        const uint16_t* dtok = static_cast<const uint16_t*>(ctx->d_tokens.p);
        CK(ctx->d_rc_ck.reserve((((size_t)ctx->token_total >> 10) + 2 * (size_t)nseg + 8) * sizeof(unsigned long long)));
        unsigned long long* dck = static_cast<unsigned long long*>(ctx->d_rc_ck.p);
        const bool trace = getenv("LEPB200_TRACE") != nullptr;           // per-kernel times on stderr (diagnostics; adds a sync)
        cudaEvent_t te[4] = {nullptr, nullptr, nullptr, nullptr};
        if (trace) { for (auto& e : te) cudaEventCreate(&e); cudaEventRecord(te[0], ctx->stream); }
        const bool async_feed = ctx->rc_feed < 0 ? nseg <= 8192 : ctx->rc_feed != 0;
        if (async_feed) lep_rangepass_kernel<true><<<(nseg + RCT_THREADS - 1) / RCT_THREADS, RCT_THREADS, RCT_SMEM_BYTES, ctx->stream>>>(ds, nseg, dord, dtok, dck);
        else lep_rangepass_kernel<false><<<(nseg + RCT_THREADS - 1) / RCT_THREADS, RCT_THREADS, RCT_SMEM_TABLE_BYTES, ctx->stream>>>(ds, nseg, dord, dtok, dck);
        CK(cudaGetLastError());
        unsigned long long* d_total = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ctx->d_counter.p) + 128);
        lep_digit_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(ds, nseg, d_total);
        CK(cudaGetLastError());
        unsigned long long total_digits = 0;
        CK(cudaMemcpyAsync(&total_digits, d_total, sizeof(total_digits), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (trace) cudaEventRecord(te[1], ctx->stream);
        CK(ctx->d_rc_digits.reserve((size_t)total_digits * 4 + 256));
        CK(cudaMemsetAsync(ctx->d_rc_digits.p, 0, (size_t)total_digits * 4, ctx->stream));
        uint32_t* ddig = static_cast<uint32_t*>(ctx->d_rc_digits.p);
        lep_rangepiece_kernel<<<dim3(8, (unsigned)nseg), RCP_THREADS, 0, ctx->stream>>>(ds, nseg, dtok, dck, ddig);
        CK(cudaGetLastError());
        if (trace) cudaEventRecord(te[2], ctx->stream);
        lep_rangenorm_kernel<<<(nseg + RCN_WARPS - 1) / RCN_WARPS, RCN_WARPS * 32, 0, ctx->stream>>>(ds, nseg, dord, ddig);
        CK(cudaGetLastError());
        if (trace) {
            cudaEventRecord(te[3], ctx->stream);
            cudaEventSynchronize(te[3]);
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, te[0], te[1]); cudaEventElapsedTime(&b, te[1], te[2]); cudaEventElapsedTime(&c, te[2], te[3]);
            fprintf(stderr, "[trace]   range coder: range pass + offsets %.1f ms, pieces (incl. digit zero fill) %.1f ms, carries %.1f ms, %d segments\n", a, b, c, nseg);
            for (auto& e : te) cudaEventDestroy(e);
        }
        ctx->launches += 3;
    } else {
        lep_rangecode_kernel<<<(nseg + RC_THREADS - 1) / RC_THREADS, RC_THREADS, 0, ctx->stream>>>(
            static_cast<SegDesc*>(ctx->d_segs.p), nseg, static_cast<const int*>(ctx->d_order.p), static_cast<const uint16_t*>(ctx->d_tokens.p));
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    ctx->symbolised = false;
    ctx->launched = true;
    return LEPB200_OK;
}

int lepb200_encode_launch(lepb200_ctx* ctx) {
    const int r = lepb200_encode_launch_symbolise(ctx);
    return r ? r : lepb200_encode_launch_rangecode(ctx);
}

int lepb200_encode_fetch(lepb200_ctx* ctx, lepb200_stream* out) {
    if (!ctx || !ctx->launched || !ctx->is_encode || !out) { if (ctx) ctx->err = "encode_fetch without encode_launch"; return LEPB200_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const int nseg = (int)ctx->segs.size();
    CK(ctx->h_segs.reserve(sizeof(SegDesc) * nseg + sizeof(unsigned long long) * nseg));
    SegDesc* hs = static_cast<SegDesc*>(ctx->h_segs.p);
    CK(cudaMemcpyAsync(hs, ctx->d_segs.p, sizeof(SegDesc) * nseg, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    // dense layout of the produced streams
    unsigned long long* offs = reinterpret_cast<unsigned long long*>(hs + nseg);
    size_t total = 0;
    uint64_t alg = 0;
    for (int s = 0; s < nseg; ++s) {
        offs[s] = total;
        size_t n = hs[s].status == 0 ? hs[s].len : 0;
        total += align_up(n, 16);
        alg += (uint64_t)ctx->seg_blocks[s] * 128 + n;
    }
    ctx->alg_bytes = alg;
    const size_t dense_bytes = align_up(total, 256);
    CK(ctx->d_dense.reserve(dense_bytes + sizeof(unsigned long long) * nseg));
    CK(ctx->h_dense.reserve(total + 16));
    unsigned long long* d_offs = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ctx->d_dense.p) + dense_bytes);
    CK(cudaMemcpyAsync(d_offs, offs, sizeof(unsigned long long) * nseg, cudaMemcpyHostToDevice, ctx->stream));
    lep_compact_kernel<<<nseg, 256, 0, ctx->stream>>>(static_cast<const SegDesc*>(ctx->d_segs.p), d_offs, static_cast<uint8_t*>(ctx->d_dense.p), nseg);
    CK(cudaGetLastError());
    ctx->launches += 1;
    if (total) CK(cudaMemcpyAsync(ctx->h_dense.p, ctx->d_dense.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < nseg; ++s) {
        out[s].data = static_cast<const uint8_t*>(ctx->h_dense.p) + offs[s];
        out[s].len = hs[s].status == 0 ? hs[s].len : 0;
        out[s].status = hs[s].status;
        out[s].reserved = 0;
        out[s].ndecisions = (uint64_t)hs[s].ndecisions_lo | ((uint64_t)hs[s].ndecisions_hi << 32);
    }
    return LEPB200_OK;
}

// Final .lep files from the device: headers[i] = everything in front of the mux packets
// of image i (fixed header + zlib'd JPEG header + "CMP", built on the host: lephost::build_lep_header); the MuxWriter
// schedule is planned on the host from the stream lengths the range coder reports (lephost::plan_mux, data-free), the
// bytes are moved by lep_gather_kernel, the LE32 size trailer (vp8_encoder.cc:603-614) is a literal.
int lepb200_encode_fetch_files(lepb200_ctx* ctx, const lepb200_buffer* headers, lepb200_result* files) {
    if (!ctx || !ctx->launched || !ctx->is_encode || !headers || !files) { if (ctx) ctx->err = "encode_fetch_files without encode_launch"; return LEPB200_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const int nseg = (int)ctx->segs.size(), nimg = (int)ctx->images.size();
    CK(ctx->h_segs.reserve(sizeof(SegDesc) * nseg));
    SegDesc* hs = static_cast<SegDesc*>(ctx->h_segs.p);
    CK(cudaMemcpyAsync(hs, ctx->d_segs.p, sizeof(SegDesc) * nseg, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    // literals: header + 4 trailer bytes per image, 16-byte aligned
    size_t lit_total = 0;
    std::vector<size_t> lit_off(nimg);
    for (int i = 0; i < nimg; ++i) { lit_off[i] = lit_total; lit_total += align_up(headers[i].len + 4, 16); }
    CK(ctx->h_lit.reserve(lit_total + 64));
    CK(ctx->d_lit.reserve(lit_total + 256));
    uint8_t* hl = static_cast<uint8_t*>(ctx->h_lit.p);
    const unsigned long long dl = (unsigned long long)(uintptr_t)ctx->d_lit.p;
    std::vector<GatherPiece> pieces;
    pieces.reserve((size_t)nseg * 32 + 2 * (size_t)nimg);
    std::vector<size_t> file_off(nimg, 0);
    std::vector<lephost::MuxPacket> plan;
    size_t total = 0;
    uint64_t alg = 0;
    int s0 = 0;
    for (int i = 0; i < nimg; ++i) {
        int s1 = s0;
        while (s1 < nseg && hs[s1].image == i) ++s1;
        int st = 0;
        size_t lens[LEPB200_MAX_SEGMENTS];
        const int ns = s1 - s0;
        for (int s = s0; s < s1; ++s) {
            if (hs[s].status && !st) st = hs[s].status;
            if (s - s0 < LEPB200_MAX_SEGMENTS) lens[s - s0] = hs[s].len;
            alg += (uint64_t)ctx->seg_blocks[s] * 128 + (hs[s].status == 0 ? hs[s].len : 0);
        }
        files[i].data = nullptr; files[i].len = 0; files[i].status = st;
        if (st == 0 && (ns < 1 || ns > LEPB200_MAX_SEGMENTS || !headers[i].data || headers[i].len == 0)) files[i].status = st = LEPB200_ST_NOT_HANDLED;
        if (st == 0) {
            lephost::plan_mux(lens, ns, plan);
            total = align_up(total, 16);
            file_off[i] = total;
            memcpy(hl + lit_off[i], headers[i].data, headers[i].len);
            unsigned long long saddr[LEPB200_MAX_SEGMENTS];
            for (int k = 0; k < ns; ++k) saddr[k] = hs[s0 + k].stream;
            const uint32_t fsz = gather_file_pieces(plan.data(), plan.size(), saddr, dl + lit_off[i], hl + lit_off[i], headers[i].len, total, pieces);
            files[i].len = fsz;
            total += fsz;
        }
        s0 = s1;
    }
    ctx->alg_bytes = alg;
    if (pieces.empty()) return LEPB200_OK;
    CK(ctx->d_dense.reserve(total + 256));
    CK(ctx->h_dense.reserve(total + 16));
    CK(ctx->d_gather.reserve(sizeof(GatherPiece) * pieces.size()));
    CK(cudaMemcpyAsync(ctx->d_lit.p, hl, lit_total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_gather.p, pieces.data(), sizeof(GatherPiece) * pieces.size(), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = (unsigned)std::min<size_t>((pieces.size() + GATHER_WARPS - 1) / GATHER_WARPS, (size_t)ctx->sm_count * 8);
    lep_gather_kernel<<<grid, GATHER_WARPS * 32, 0, ctx->stream>>>(static_cast<const GatherPiece*>(ctx->d_gather.p), (uint32_t)pieces.size(), static_cast<uint8_t*>(ctx->d_dense.p));
    CK(cudaGetLastError());
    ctx->launches += 1;
    CK(cudaMemcpyAsync(ctx->h_dense.p, ctx->d_dense.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < nimg; ++i) if (files[i].status == 0) files[i].data = static_cast<const uint8_t*>(ctx->h_dense.p) + file_off[i];
    return LEPB200_OK;
}

int lepb200_encode_images(lepb200_ctx* ctx, const lepb200_image* images, int nimages, lepb200_stream* out) {
    int r = lepb200_encode_upload(ctx, images, nimages);
    if (r) return r;
    r = lepb200_encode_launch(ctx);
    if (r) return r;
    return lepb200_encode_fetch(ctx, out);
}

// ------------------------------------------------------------------------------------------------ GPU Huffman encode
static bool build_enc_table(const lepb200_hufftable& in, HEncTable& t) {
    memset(&t, 0, sizeof(t));
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < in.bits[len]; ++i, ++k, ++code) {
            if (k >= 256 || code >= (1 << len)) return false;        // over-subscribed table
            t.code[in.vals[k]] = (uint16_t)code;
            t.len[in.vals[k]] = (uint8_t)len;
        }
        if (code > (1 << len)) return false;
        code <<= 1;
    }
    return true;
}

static int henc_launch(lepb200_ctx* ctx, lepb200_henc_image* imgs, int n, int nparts) {
    if (!ctx || !imgs || n <= 0) return LEPB200_ERR_INVALID;
    ctx->henc_parts.clear();
    if (!ctx->launched || ctx->is_encode || n != (int)ctx->images.size()) { ctx->err = "huffman_encode_resident: needs the decode batch just launched on this context"; return LEPB200_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    std::vector<HEncImage> hi(n);
    std::vector<HEncSeg> hs;
    std::vector<HEncTable> tabs;
    std::vector<const lepb200_hufftable*> seen;
    std::vector<char> tab_ok;                     // build result per table (a cache hit on a table that failed to build fails too)
    auto table_index = [&](const lepb200_hufftable& t, bool& ok) -> int {
        for (size_t q = 0; q < seen.size(); ++q) if (!memcmp(seen[q], &t, sizeof(t))) { if (!tab_ok[q]) ok = false; return (int)q; }
        HEncTable e;
        const bool built = build_enc_table(t, e);
        if (!built) ok = false;
        tabs.push_back(e);
        tab_ok.push_back(built ? 1 : 0);
        seen.push_back(&t);
        return (int)tabs.size() - 1;
    };
    ctx->henc_off.assign(n, SIZE_MAX);
    ctx->henc_seg_first.assign(n, -1);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        lepb200_henc_image& im = imgs[i];
        im.data = nullptr; im.status = 0;
        HEncImage& d = hi[i];
        memset(&d, 0, sizeof(d));
        if (im.scan_bytes == 0) continue;
        const ImageDesc& g = ctx->images[i];
        bool ok = im.nseg >= 1 && im.nseg <= LEPB200_MAX_SEGMENTS && g.ncmp >= 1 && g.ncmp <= 3;
        d.ncmp = g.ncmp; d.mcuv = g.mcuv; d.rsti = im.rsti; d.padbit = im.padbit; d.scan_len = im.scan_bytes;
        for (int c = 0; ok && c < g.ncmp; ++c) {
            ok = im.H[c] >= 1 && im.H[c] <= 2 && im.V[c] >= 1 && im.V[c] <= 2 && g.bch[c] % im.H[c] == 0 && g.bcv[c] == g.mcuv * im.V[c];
            d.H[c] = im.H[c]; d.V[c] = im.V[c]; d.bch[c] = g.bch[c]; d.plane[c] = g.plane[c];
            if (ok) { d.dc_tab[c] = table_index(im.dc[c], ok); d.ac_tab[c] = table_index(im.ac[c], ok); }
        }
        if (ok) {
            d.mcuh = g.bch[0] / im.H[0];
            for (int c = 1; ok && c < g.ncmp; ++c) ok = g.bch[c] / im.H[c] == d.mcuh;
        }
        if (!ok) { im.status = LEPB200_ST_NOT_HANDLED; im.scan_bytes = 0; continue; }
        ctx->henc_off[i] = total;
        ctx->henc_seg_first[i] = (int)hs.size();
        uint32_t off = 0;
        for (int k = 0; k < im.nseg; ++k) {
            HEncSeg sg;
            memset(&sg, 0, sizeof(sg));
            sg.image = i; sg.my0 = im.seg[k].mcu_row_start; sg.my1 = im.seg[k].mcu_row_end;
            for (int c = 0; c < 3; ++c) sg.lastdc[c] = im.seg[k].last_dc[c];
            sg.ov_bits = im.seg[k].overhang_bits; sg.ov_byte = im.seg[k].overhang_byte;
            sg.out_off = off; sg.expect = im.seg[k].expect_bytes; sg.is_last = k + 1 == im.nseg;
            off += im.seg[k].expect_bytes;
            hs.push_back(sg);
        }
        total += align_up((size_t)im.scan_bytes + 16, 256);
    }
    ctx->henc_nseg = (int)hs.size();
    if (hs.empty()) return LEPB200_OK;
    CK(ctx->d_henc_out.reserve(total + 256));
    CK(ctx->d_henc_imgs.reserve(sizeof(HEncImage) * n));
    CK(ctx->d_henc_segs.reserve(sizeof(HEncSeg) * hs.size()));
    CK(ctx->d_henc_tabs.reserve(sizeof(HEncTable) * std::max<size_t>(1, tabs.size())));
    for (int i = 0; i < n; ++i) if (ctx->henc_off[i] != SIZE_MAX) hi[i].out = (unsigned long long)(uintptr_t)ctx->d_henc_out.p + ctx->henc_off[i];
    // the records travel from pinned memory: a copy from pageable memory of this size would hold the calling thread
    // until the stream has reached it, i.e. for the whole decode kernel queued in front
    {
        const size_t b_img = align_up(sizeof(HEncImage) * (size_t)n, 256), b_seg = align_up(sizeof(HEncSeg) * hs.size(), 256), b_tab = sizeof(HEncTable) * tabs.size();
        CK(ctx->h_henc_desc.reserve(b_img + b_seg + b_tab + 256));
        uint8_t* hd = static_cast<uint8_t*>(ctx->h_henc_desc.p);
        memcpy(hd, hi.data(), sizeof(HEncImage) * (size_t)n);
        memcpy(hd + b_img, hs.data(), sizeof(HEncSeg) * hs.size());
        memcpy(hd + b_img + b_seg, tabs.data(), b_tab);
        CK(cudaMemcpyAsync(ctx->d_henc_imgs.p, hd, sizeof(HEncImage) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_henc_segs.p, hd + b_img, sizeof(HEncSeg) * hs.size(), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_henc_tabs.p, hd + b_img + b_seg, b_tab, cudaMemcpyHostToDevice, ctx->stream));
    }
    const int nseg = (int)hs.size();
    if (nparts <= 0) {                    // one launch, the caller fetches everything with lepb200_huffman_encode_fetch
        lep_huffencode_kernel<<<(nseg + HENC_WARPS - 1) / HENC_WARPS, HENC_WARPS * 32, 0, ctx->stream>>>(
            static_cast<const HEncImage*>(ctx->d_henc_imgs.p), static_cast<HEncSeg*>(ctx->d_henc_segs.p), nseg, static_cast<const HEncTable*>(ctx->d_henc_tabs.p));
        CK(cudaGetLastError());
        ctx->launches += 1;
        return LEPB200_OK;
    }
    // parts of consecutive images with about equal output bytes: one launch each, and behind each launch the D2H of its
    // scan bytes and segment records on the copy stream -- part k travels (and the host assembles its files) while
    // part k + 1 is encoded
    CK(ctx->h_henc_out.reserve(total + 256));
    CK(ctx->h_henc_segs.reserve(sizeof(HEncSeg) * hs.size()));
    while (ctx->part_events.size() < 2 * (size_t)nparts) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->part_events.push_back(e);
    }
    // the decode batch's segment records first: they must not queue behind the part copies on the copy stream
    {
        const int nseg_dec = (int)ctx->segs.size();
        CK(ctx->h_segs.reserve(sizeof(SegDesc) * nseg_dec));
        if (!ctx->status_ev) CK(cudaEventCreateWithFlags(&ctx->status_ev, cudaEventDisableTiming));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev1, 0));
        CK(cudaMemcpyAsync(ctx->h_segs.p, ctx->d_segs.p, sizeof(SegDesc) * nseg_dec, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CK(cudaEventRecord(ctx->status_ev, ctx->copy_stream));
        ctx->status_queued = true;
    }
    int i0 = 0;
    for (int k = 0; k < nparts && i0 < n; ++k) {
        const size_t want_end = k + 1 == nparts ? total : total / nparts * (k + 1);
        int i1 = i0;
        size_t bend = 0;
        auto end_of = [&](int i) { return ctx->henc_off[i] == SIZE_MAX ? (size_t)0 : ctx->henc_off[i] + align_up((size_t)imgs[i].scan_bytes + 16, 256); };
        while (i1 < n && (k + 1 == nparts || std::max(bend, end_of(i1)) <= want_end || i1 == i0)) { bend = std::max(bend, end_of(i1)); ++i1; }
        lepb200_ctx::HEncPart pt;
        pt.i0 = i0; pt.i1 = i1; pt.s0 = -1; pt.s1 = -1; pt.b0 = SIZE_MAX; pt.b1 = 0; pt.done = ctx->part_events[2 * k + 1];
        for (int i = i0; i < i1; ++i) {
            if (ctx->henc_off[i] == SIZE_MAX) continue;
            if (pt.s0 < 0) pt.s0 = ctx->henc_seg_first[i];
            pt.s1 = ctx->henc_seg_first[i] + imgs[i].nseg;
            pt.b0 = std::min(pt.b0, ctx->henc_off[i]);
            pt.b1 = std::max(pt.b1, ctx->henc_off[i] + (size_t)imgs[i].scan_bytes);
        }
        if (pt.s0 >= 0) {
            const int ns = pt.s1 - pt.s0;
            lep_huffencode_kernel<<<(ns + HENC_WARPS - 1) / HENC_WARPS, HENC_WARPS * 32, 0, ctx->stream>>>(
                static_cast<const HEncImage*>(ctx->d_henc_imgs.p), static_cast<HEncSeg*>(ctx->d_henc_segs.p) + pt.s0, ns, static_cast<const HEncTable*>(ctx->d_henc_tabs.p));
            CK(cudaGetLastError());
            ctx->launches += 1;
            CK(cudaEventRecord(ctx->part_events[2 * k], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->part_events[2 * k], 0));
            CK(cudaMemcpyAsync(static_cast<uint8_t*>(ctx->h_henc_out.p) + pt.b0, static_cast<const uint8_t*>(ctx->d_henc_out.p) + pt.b0, pt.b1 - pt.b0, cudaMemcpyDeviceToHost, ctx->copy_stream));
            CK(cudaMemcpyAsync(static_cast<HEncSeg*>(ctx->h_henc_segs.p) + pt.s0, static_cast<const HEncSeg*>(ctx->d_henc_segs.p) + pt.s0, sizeof(HEncSeg) * (size_t)ns, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        CK(cudaEventRecord(pt.done, ctx->copy_stream));
        ctx->henc_parts.push_back(pt);
        i0 = i1;
    }
    return LEPB200_OK;
}

int lepb200_huffman_encode_resident(lepb200_ctx* ctx, lepb200_henc_image* imgs, int n) { return henc_launch(ctx, imgs, n, 0); }

int lepb200_huffman_encode_resident_parts(lepb200_ctx* ctx, lepb200_henc_image* imgs, int n, int nparts) {
    return henc_launch(ctx, imgs, n, std::max(1, std::min(16, nparts)));
}

int lepb200_huffman_encode_parts(const lepb200_ctx* ctx) { return ctx ? (int)ctx->henc_parts.size() : 0; }

int lepb200_huffman_encode_wait_part(lepb200_ctx* ctx, lepb200_henc_image* imgs, int n, int part, int* first, int* last) {
    if (!ctx || !imgs || !first || !last || n != (int)ctx->henc_off.size() || part < 0 || part >= (int)ctx->henc_parts.size()) return LEPB200_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const lepb200_ctx::HEncPart& pt = ctx->henc_parts[part];
    CK(cudaEventSynchronize(pt.done));
    const HEncSeg* hs = static_cast<const HEncSeg*>(ctx->h_henc_segs.p);
    for (int i = pt.i0; i < pt.i1; ++i) {
        if (ctx->henc_off[i] == SIZE_MAX) continue;
        imgs[i].data = static_cast<const uint8_t*>(ctx->h_henc_out.p) + ctx->henc_off[i];
        int st = 0;
        for (int k = 0; k < imgs[i].nseg; ++k) if (hs[ctx->henc_seg_first[i] + k].status) st = 1;
        imgs[i].status = st;
    }
    *first = pt.i0; *last = pt.i1;
    return LEPB200_OK;
}

int lepb200_huffman_encode_fetch(lepb200_ctx* ctx, lepb200_henc_image* imgs, int n) {
    if (!ctx || !imgs || n != (int)ctx->henc_off.size()) return LEPB200_ERR_INVALID;
    if (ctx->henc_nseg == 0) return LEPB200_OK;
    CK(cudaSetDevice(ctx->device));
    size_t total = 0;
    for (int i = 0; i < n; ++i) if (ctx->henc_off[i] != SIZE_MAX) total = std::max(total, ctx->henc_off[i] + imgs[i].scan_bytes);
    CK(ctx->h_henc_out.reserve(total + 256));
    CK(ctx->h_henc_segs.reserve(sizeof(HEncSeg) * ctx->henc_nseg));
    CK(cudaMemcpyAsync(ctx->h_henc_out.p, ctx->d_henc_out.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_henc_segs.p, ctx->d_henc_segs.p, sizeof(HEncSeg) * ctx->henc_nseg, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const HEncSeg* hs = static_cast<const HEncSeg*>(ctx->h_henc_segs.p);
    for (int i = 0; i < n; ++i) {
        if (ctx->henc_off[i] == SIZE_MAX) continue;
        imgs[i].data = static_cast<const uint8_t*>(ctx->h_henc_out.p) + ctx->henc_off[i];
        int st = 0;
        for (int k = 0; k < imgs[i].nseg; ++k) if (hs[ctx->henc_seg_first[i] + k].status) st = 1;
        imgs[i].status = st;
    }
    return LEPB200_OK;
}

// ------------------------------------------------------------------------------------------------ decode
static int decode_upload_impl(lepb200_ctx* ctx, const lepb200_image* images, int nimages, const lepb200_stream* in,
                              const lepb200_buffer* spans, const uint32_t* span_first) {
    if (!ctx || !in) return LEPB200_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->d_tokens.release();              // the encoder's token arena (the largest buffer of that direction) is not needed on the way back
    int r = build_batch(ctx, images, nimages, false, in);
    if (r) return r;
    const int nseg = (int)ctx->segs.size();
    // planes start zeroed: blocks outside the coded range (truncated images) stay zero like the reference's calloc.
    // Issued first: the device clears them while the host packs the streams.
    size_t plane_total = 0;
    for (int i = 0; i < nimages; ++i)
        for (int c = 0; c < images[i].ncmp; ++c) plane_total = std::max(plane_total, (size_t)(ctx->images[i].plane[c] - (unsigned long long)(uintptr_t)ctx->d_planes.p) + ctx->plane_bytes[(size_t)i * 3 + c]);
    CK(cudaMemsetAsync(ctx->d_planes.p, 0, plane_total, ctx->stream));
    // pack the streams into the pinned staging buffer (over a gigabyte for a 4096-image batch: a single-threaded copy
    // loop took longer than the H2D itself) in a few slices of consecutive segments; every slice is packed by the
    // context's host threads and its H2D copy runs while the next one is packed
    const unsigned long long base = (unsigned long long)(uintptr_t)ctx->d_streams.p;
    size_t total = 0;
    for (int s = 0; s < nseg; ++s) total = std::max(total, (size_t)(ctx->segs[s].stream - base) + in[s].len);
    CK(ctx->h_stage.reserve(total + 16));
    uint8_t* hs = static_cast<uint8_t*>(ctx->h_stage.p);
    const int nslices = total > (size_t(64) << 20) ? 8 : 1;
    int s0 = 0;
    for (int k = 0; k < nslices && s0 < nseg; ++k) {
        // segments are laid out in order, so a slice is a contiguous byte range of the staging buffer
        const size_t want_end = total / nslices * (k + 1);
        int s1 = s0;
        while (s1 < nseg && (k + 1 == nslices || (size_t)(ctx->segs[s1].stream - base) < want_end)) ++s1;
        const int nt = std::max(1, std::min(ctx->host_threads, s1 - s0));
        auto pack = [&](int t) {
            for (int s = s0 + t; s < s1; s += nt) {
                uint8_t* dst = hs + (ctx->segs[s].stream - base);
                if (!spans) { if (in[s].len) memcpy(dst, in[s].data, in[s].len); continue; }
                size_t room = in[s].len;                       // the pieces of a stream as they lie in the caller's file
                for (uint32_t q = span_first[s]; q < span_first[s + 1] && room; ++q) {
                    const size_t n = std::min(room, spans[q].len);
                    memcpy(dst, spans[q].data, n);
                    dst += n; room -= n;
                }
            }
        };
        if (nt == 1) pack(0);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back(pack, t);
            for (auto& t : th) t.join();
        }
        const size_t b0 = (size_t)(ctx->segs[s0].stream - base);
        const size_t b1 = s1 < nseg ? (size_t)(ctx->segs[s1].stream - base) : total;
        if (b1 > b0) CK(cudaMemcpyAsync(static_cast<uint8_t*>(ctx->d_streams.p) + b0, hs + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->stream));
        s0 = s1;
    }
    ctx->have_batch = true;
    return LEPB200_OK;
}

int lepb200_decode_upload(lepb200_ctx* ctx, const lepb200_image* images, int nimages, const lepb200_stream* in) {
    return decode_upload_impl(ctx, images, nimages, in, nullptr, nullptr);
}

int lepb200_decode_upload_gather(lepb200_ctx* ctx, const lepb200_image* images, int nimages, const lepb200_stream* in,
                                 const lepb200_buffer* spans, const uint32_t* span_first) {
    if (!spans || !span_first) return LEPB200_ERR_INVALID;
    return decode_upload_impl(ctx, images, nimages, in, spans, span_first);
}

int lepb200_decode_launch(lepb200_ctx* ctx) {
    if (!ctx || !ctx->have_batch || ctx->is_encode) { if (ctx) ctx->err = "decode_launch without decode_upload"; return LEPB200_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const int nseg = (int)ctx->segs.size();
    CK(cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(int), ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (ctx->dec_threads > 0) {
        // group kernel (lep_decode_g2.cu): G lanes per segment, 32 / G segments per warp in lock step; the groups of a
        // launch share a queue of at most dec_threads segments (one zero-filled model each), largest first
        int warps = 0, per_warp = 0, gsm = 0;
        group_launch_shape(ctx->dec_lanes, warps, per_warp, gsm);
        const int per_cta = warps * per_warp;
        for (int first = 0; first < nseg; first += ctx->dec_threads) {
            const int count = std::min(ctx->dec_threads, nseg - first);
            const int grid = std::max(1, std::min(ctx->dec_group_grid, (count + per_cta - 1) / per_cta));
            CK(cudaMemsetAsync(ctx->d_models.p, 0, (size_t)count * MODEL_BYTES, ctx->stream));       // identity prior = zero fill
            CK(cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(int), ctx->stream));
            const ImageDesc* di = static_cast<const ImageDesc*>(ctx->d_images.p);
            SegDesc* ds = static_cast<SegDesc*>(ctx->d_segs.p);
            const int* dord = static_cast<const int*>(ctx->d_order.p);
            int* dcnt = static_cast<int*>(ctx->d_counter.p);
            uint16_t* dm = static_cast<uint16_t*>(ctx->d_models.p);
            uint8_t* dr = static_cast<uint8_t*>(ctx->d_rows.p);
            switch (ctx->dec_lanes) {
            case 8: launch_group_kernel<8>(grid, ctx->stream, di, ds, first, count, dord, dcnt, dm, dr, ctx->row_stride); break;
            case 32: launch_group_kernel<32>(grid, ctx->stream, di, ds, first, count, dord, dcnt, dm, dr, ctx->row_stride); break;
            default: launch_group_kernel<4>(grid, ctx->stream, di, ds, first, count, dord, dcnt, dm, dr, ctx->row_stride); break;
            }
            CK(cudaGetLastError());
            ctx->launches += 1;
        }
        ctx->launches -= 1;
    } else {
        lep_decode_kernel<<<ctx->grid, DEC_WARPS_PER_CTA * 32, 0, ctx->stream>>>(
            static_cast<const ImageDesc*>(ctx->d_images.p), static_cast<SegDesc*>(ctx->d_segs.p), nseg, static_cast<const int*>(ctx->d_order.p),
            static_cast<int*>(ctx->d_counter.p), static_cast<uint16_t*>(ctx->d_models.p), static_cast<uint8_t*>(ctx->d_rows.p), ctx->row_stride);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->status_queued = false;
    ctx->launches += 1;
    ctx->launched = true;
    return LEPB200_OK;
}

int lepb200_decode_fetch(lepb200_ctx* ctx, const lepb200_image* images, int nimages, int32_t* status_out) {
    if (!ctx || !ctx->launched || ctx->is_encode) { if (ctx) ctx->err = "decode_fetch without decode_launch"; return LEPB200_ERR_INVALID; }
    if (nimages != (int)ctx->images.size()) { ctx->err = "decode_fetch: batch size mismatch"; return LEPB200_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const int nseg = (int)ctx->segs.size();
    CK(ctx->h_segs.reserve(sizeof(SegDesc) * nseg));
    SegDesc* hs = static_cast<SegDesc*>(ctx->h_segs.p);
    CK(cudaMemcpyAsync(hs, ctx->d_segs.p, sizeof(SegDesc) * nseg, cudaMemcpyDeviceToHost, ctx->stream));
    for (int i = 0; i < nimages; ++i)
        for (int c = 0; c < images[i].ncmp; ++c)
            if (images[i].planes[c])        // NULL: the caller does not need this plane on the host (scan re-encoded on the device)
                CK(cudaMemcpyAsync(images[i].planes[c], reinterpret_cast<const void*>(ctx->images[i].plane[c]), ctx->plane_bytes[(size_t)i * 3 + c],
                                   cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    uint64_t alg = 0;
    for (int s = 0; s < nseg; ++s) {
        if (status_out) status_out[s] = hs[s].status;
        alg += (uint64_t)ctx->seg_blocks[s] * 128 + ctx->segs[s].cap;
    }
    ctx->alg_bytes = alg;
    return LEPB200_OK;
}

int lepb200_decode_fetch_status(lepb200_ctx* ctx, int32_t* status_out) {
    if (!ctx || !status_out || !ctx->launched || ctx->is_encode) { if (ctx) ctx->err = "decode_fetch_status without decode_launch"; return LEPB200_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const int nseg = (int)ctx->segs.size();
    CK(ctx->h_segs.reserve(sizeof(SegDesc) * nseg));
    SegDesc* hs = static_cast<SegDesc*>(ctx->h_segs.p);
    if (ctx->status_queued) {                                            // queued by lepb200_huffman_encode_resident_parts ahead of its copies
        CK(cudaEventSynchronize(ctx->status_ev));
    } else {
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev1, 0));          // the decode kernel, not what was queued behind it
        CK(cudaMemcpyAsync(hs, ctx->d_segs.p, sizeof(SegDesc) * nseg, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CK(cudaStreamSynchronize(ctx->copy_stream));
    }
    for (int s = 0; s < nseg; ++s) status_out[s] = hs[s].status;
    return LEPB200_OK;
}

int lepb200_decode_images(lepb200_ctx* ctx, const lepb200_image* images, int nimages, const lepb200_stream* in, int32_t* status_out) {
    int r = lepb200_decode_upload(ctx, images, nimages, in);
    if (r) return r;
    r = lepb200_decode_launch(ctx);
    if (r) return r;
    return lepb200_decode_fetch(ctx, images, nimages, status_out);
}

}  // extern "C"
