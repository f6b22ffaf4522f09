// lep_decode_thread.cu -- sm_100a decode kernel, ONE THREAD per Lepton thread-segment.
//
// Decoding is one serial dependency chain per segment: every context index depends on values decoded just before, and
// the bool decoder's state threads through all of them (parse_tokens, src/vp8/decoder/decoder.cc:167-318; vpx_reader,
// src/vp8/decoder/boolreader.hh:184-258,376-416).  A warp that walks one segment (lep_decode.cu) spends 32 lanes on that
// scalar chain; here every lane walks its OWN segment, so a warp instruction advances 32 chains.  The per-block
// predictors (aavrg priors, Lakhani edge predictor, 8x8 IDCT, DC estimate) are scalar per thread, on thread-local
// raster copies of the current / left / above / above-left blocks.  Segments are assigned in order of decreasing size,
// so the threads of a warp (usually the segments of one image) run loops of similar length.
//
// The adaptive model is the same 16-bit packed layout as the encoder's (lep_common.cuh), one 1.58 MB model per thread,
// zero-filled by a memset before the launch.
#include "lep_common.cuh"
#include "lep_predict.cuh"

namespace lepb200 {

constexpr int DECT_THREADS = 32;          // one warp per CTA: the few hundred warps spread over all SMs

struct TBool {                            // vpx_reader (boolreader.hh:184-258), per thread
    unsigned long long value;
    uint32_t range;
    int count;
    const uint8_t* p;
    const uint8_t* end;
};

__device__ __forceinline__ void tb_fill(TBool& r) {
    int shift = 64 - 8 - (r.count + 8);
    while (shift >= 0) {
        const unsigned long long byte = (r.p < r.end) ? (unsigned long long)__ldg(r.p) : 0ull;
        r.p++;
        r.value |= byte << shift;
        r.count += 8;
        shift -= 8;
    }
}

struct TDec {
    TBool br;
    uint16_t* model;
    const uint32_t* rcp;
    unsigned long long ndec;
};

// VPXBoolReader::get (vpx_bool_reader.hh:45-57) = vpx_read + Branch::record_obs_and_update
__device__ __forceinline__ uint32_t td_get(TDec& d, uint32_t addr) {
    const uint32_t w = d.model[addr];
    const uint32_t prob = branch_prob(w, d.rcp);
    TBool& r = d.br;
    const uint32_t split = (r.range * prob + (256 - prob)) >> 8;
    if (r.count < 0) tb_fill(r);
    const uint32_t top = (uint32_t)(r.value >> 56);               // value >= split << 56  <=>  top byte >= split
    const uint32_t bit = top >= split;
    const uint32_t range = bit ? r.range - split : split;
    if (bit) r.value -= (unsigned long long)split << 56;
    const int shift = __clz(range) - 24;
    r.range = range << shift;
    r.value <<= shift;
    r.count -= shift;
    const bool plain = (w & 0xffu) < 254u && (w >> 8) < 254u;       // no count about to saturate, not the special state
    d.model[addr] = (uint16_t)(plain ? w + (bit ? 0x100u : 1u) : branch_update(w, bit));
    d.ndec++;
    return bit;
}

// exponent unary + sign + residual bits of one coefficient (decoder.cc:212-240)
__device__ __forceinline__ int td_coef_plain(TDec& d, uint32_t exp_addr, uint32_t sign_addr, uint32_t res_addr) {
    int len = 0;
    while (len < 11) { if (!td_get(d, exp_addr + len)) break; ++len; }
    if (len == 0) return 0;
    const bool neg = !td_get(d, sign_addr);
    int val = 1 << (len - 1);
    for (int i = len - 2; i >= 0; --i) val |= (int)td_get(d, res_addr + i) << i;
    return neg ? -val : val;
}

// compute_lak (model.hh:1033-1071), scalar; same arithmetic as lak_pred in lep_predict.cuh
__device__ __forceinline__ int t_lak(const int16_t* cur, const int16_t* nb, const int32_t* __restrict__ icos, int first, int step) {
    uint32_t pred = (uint32_t)(int32_t)nb[first] * (uint32_t)icos[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        const int32_t t = (int32_t)cur[first + i * step] + ((i & 1) ? (int32_t)nb[first + i * step] : -(int32_t)nb[first + i * step]);
        pred -= (uint32_t)icos[i] * (uint32_t)t;
    }
    const int32_t p = (int32_t)pred;
    const int32_t t = (p + ((p >> 31) & 8191)) >> 13;
    return div_trunc_small(t, icos[0] >> 13);
}

__global__ void __launch_bounds__(DECT_THREADS)
lep_decode_thread_kernel(const ImageDesc* __restrict__ images, SegDesc* __restrict__ segs, int first, int count, const int* __restrict__ order,
                         uint16_t* __restrict__ model_pool, uint8_t* __restrict__ row_pool, size_t row_pool_stride) {
    __shared__ uint32_t s_rcp[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_rcp[i] = i < 2 ? 0u : (uint32_t)((0x100000000ull + i - 1) / i);
    __syncthreads();
    const int t = blockIdx.x * DECT_THREADS + threadIdx.x;
    if (t >= count) return;
    SegDesc& sd = segs[order[first + t]];
    const ImageDesc& g = images[sd.image];
    if (sd.status != ST_OK) return;               // rejected on the host (e.g. zero quantiser, model.hh:257-262)
    uint8_t* rowbuf = row_pool + (size_t)t * row_pool_stride;

    TDec d;
    d.model = model_pool + (size_t)t * M_TOTAL;   // zero-filled before the launch
    d.rcp = s_rcp; d.ndec = 0;
    d.br.value = 0; d.br.count = -8; d.br.range = 255;
    d.br.p = reinterpret_cast<const uint8_t*>(sd.stream); d.br.end = d.br.p + sd.cap;
    tb_fill(d.br);
    {   // marker bit at p = 128 (boolreader.cc:26-35); no model involved
        TBool& r = d.br;
        const uint32_t split = (r.range * 128u + 128u) >> 8;
        const uint32_t bit = (uint32_t)(r.value >> 56) >= split;
        const uint32_t range = bit ? r.range - split : split;
        if (bit) r.value -= (unsigned long long)split << 56;
        const int shift = __clz(range) - 24;
        r.range = range << shift; r.value <<= shift; r.count -= shift;
    }

    const int bw0 = g.bch[0], bw1 = g.ncmp > 1 ? g.bch[1] : 0, bw2 = g.ncmp > 2 ? g.bch[2] : 0;
    const size_t nz_base = (size_t)(bw0 + bw1 + bw2) * 16;
    const int nzs0 = (bw0 + 15) & ~15, nzs1 = (bw1 + 15) & ~15;

    // thread-local raster-order blocks: [0],[1] current / left (ping-pong), [2],[3] above / above-left (ping-pong)
    int16_t blk[4][64];
    int32_t tmp[64];
    int16_t pix[64];

    int status = ST_OK;
    uint32_t top_mask = 7u;
    uint32_t index = 0;
    for (;;) {
        const RowSpec rs = row_spec_from_index(index++, g);
        if (rs.done) break;
        if (rs.luma_y >= sd.max_y && !sd.is_last) break;
        if (rs.skip) continue;
        if (rs.luma_y < sd.min_y) continue;
        const int c = rs.component, y = rs.curr_y;
        const bool has_above = !((top_mask >> c) & 1u);
        top_mask &= ~(1u << c);
        const int ci = c == 0 ? 0 : 1;
        const int w = g.bch[c];
        int16_t* rowp = reinterpret_cast<int16_t*>(g.plane[c]) + (size_t)y * w * 64;
        const int16_t* abovep = rowp - (size_t)w * 64;
        const uint16_t* q = g.q[c];
        const int q0 = q[0];
        int16_t* redge = reinterpret_cast<int16_t*>(rowbuf + (size_t)(c == 0 ? 0 : (c == 1 ? bw0 : bw0 + bw1)) * 16);
        uint8_t* rnz = rowbuf + nz_base + (c == 0 ? 0 : (c == 1 ? nzs0 : nzs0 + nzs1));
        const int32_t* icx = g.icos_x[c];
        const int32_t* icy = g.icos_y[c];

        int pc = 0, pa = 2;                       // slots of the current and the above block
        int16_t left_edge[8];                     // right-column edge prediction of the left neighbour
        int nz_left = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) left_edge[i] = 0;
        for (int x = 0; x < w; ++x) {
            const bool has_left = x > 0;
            int16_t* rcur = blk[pc];
            const int16_t* rleft = blk[pc ^ 1];
            int16_t* rabove = blk[pa];
            const int16_t* raleft = blk[pa ^ 1];
            // above block: aligned order in memory -> raster copy (one 128-byte line per thread)
            if (has_above) {
                const uint4* src = reinterpret_cast<const uint4*>(abovep + (size_t)x * 64);
#pragma unroll
                for (int v4 = 0; v4 < 8; ++v4) {
                    const uint4 u = src[v4];
                    const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        rabove[c_aligned_to_raster[v4 * 8 + 2 * k]] = (int16_t)(wds[k] & 0xffff);
                        rabove[c_aligned_to_raster[v4 * 8 + 2 * k + 1]] = (int16_t)(wds[k] >> 16);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 64; ++i) rcur[i] = 0;

            // ---- (i) 7x7 non-zero count
            const int nz_above = has_above ? (int)rnz[x] : 0;
            int nz = 0;
            {
                int ctx = 0;
                if (has_above && !has_left) ctx = (nz_above + 1) / 2;
                else if (has_left && !has_above) ctx = (nz_left + 1) / 2;
                else if (has_left && has_above) ctx = (nz_above + nz_left + 2) / 4;
                const int bin = c_nonzero_to_bin[ctx];
                int prefix = 0;
                for (int idx = 5; idx >= 0; --idx) {
                    const uint32_t b = td_get(d, m_nz7(ci, bin, idx, prefix));
                    nz |= (int)b << idx;
                    prefix = (prefix << 1) | (int)b;
                }
            }
            if (nz > 49) { status = ST_STREAM_INCONSISTENT; break; }
            // ---- (ii) 7x7 coefficients (zig-zag order == aligned order 0..48)
            int eobx = 0, eoby = 0, left_nz = nz;
            for (int zz = 0; zz < 49 && left_nz > 0; ++zz) {
                const int coord = c_aligned_to_raster[zz];
                const int prior = aavrg16(rleft[coord], rabove[coord], raleft[coord], has_left, has_above);
                const int bin = c_nonzero_to_bin[left_nz];
                const int bsr = bitlen((uint32_t)min(iabs(prior), 1023));
                const int v = td_coef_plain(d, m_exp7(ci, bin, zz, bsr), m_sign(ci, 0, 0), m_resn(ci, coord, bin));
                if (v != 0) {
                    --left_nz;
                    eobx = max(eobx, coord & 7); eoby = max(eoby, coord >> 3);
                    rcur[coord] = (int16_t)v;
                }
            }
            // ---- (iii) edges: horizontal (raster 1..7) then vertical (raster 8..56)
            for (int vert = 0; vert < 2; ++vert) {
                const int eob = vert ? eoby : eobx;
                int ne = 0, prefix = 0;
                for (int i = 2; i >= 0; --i) {
                    const uint32_t b = td_get(d, m_nze(vert, ci, eob, (nz + 3) / 7, i, prefix));
                    ne |= (int)b << i;
                    prefix = (prefix << 1) | (int)b;
                }
                // the Lakhani predictions of one edge all use the block as it is BEFORE that edge is decoded only through
                // the 7x7 part and the other edge's row/column 0 entries are not involved (model.hh:1033-1071): predict on demand
                for (int ln = 0; ln < 7 && ne > 0; ++ln) {
                    const int kk = ln + 1;
                    const int coord = vert ? 8 * kk : kk;
                    const int zig15 = vert ? 7 + ln : ln;
                    int prior = 0;
                    if (!vert && has_above) prior = t_lak(rcur, rabove, icx + kk * 8, kk, 8);
                    if (vert && has_left) prior = t_lak(rcur, rleft, icy + kk * 8, 8 * kk, 1);
                    const int bsr = bitlen((uint32_t)min(iabs(prior), 1023));
                    const uint32_t ea = m_expx(ci, ne, zig15, bsr);
                    int len = 0;
                    while (len < 11) { if (!td_get(d, ea + len)) break; ++len; }
                    if (len) {
                        const int p16 = (int)(int16_t)prior;
                        const int sctx = p16 == 0 ? 0 : (p16 > 0 ? 1 : 2);
                        const bool neg = !td_get(d, m_sign(ci, sctx, bsr));
                        const int ne0 = ne;
                        --ne;
                        int val = 1 << (len - 1);
                        if (len > 1) {
                            const int min_thr = g.min_thr[c][coord];
                            int i = len - 2;
                            if (i >= min_thr) {
                                const int ctx_abs = iabs(prior) & 0xffff;
                                const uint32_t ta = m_thr(ci, min(ctx_abs >> min_thr, 255), min(len - min_thr, 7));
                                uint32_t so = 1;
                                for (; i >= min_thr; --i) {
                                    const uint32_t b = td_get(d, ta + so);
                                    val |= (int)b << i;
                                    so = min((so << 1) | b, 127u);
                                }
                            }
                            const uint32_t ra = m_resn(ci, coord, ne0);
                            for (; i >= 0; --i) val |= (int)td_get(d, ra + i) << i;
                        }
                        rcur[coord] = (int16_t)(neg ? -val : val);
                    }
                }
            }
            // ---- (iv) DC: pixels of the block without its DC, prediction from the neighbours' edge pixels
            {
                int32_t in[8], out[8];
                for (int r = 0; r < 8; ++r) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) in[k] = (int32_t)rcur[r * 8 + k] * (int32_t)q[r * 8 + k];
                    if (r == 0) in[0] = 0;
                    idct_row(in, out);
#pragma unroll
                    for (int k = 0; k < 8; ++k) tmp[r * 8 + k] = out[k];
                }
                for (int col = 0; col < 8; ++col) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) in[k] = tmp[k * 8 + col];
                    idct_col(in, out);
#pragma unroll
                    for (int k = 0; k < 8; ++k) pix[k * 8 + col] = (int16_t)out[k];
                }
            }
            int pred = 0, unc = 0, unc2 = 0;
            {
                // adv_predict_dc_pix (model.hh:678-784), 16-bit lane arithmetic of the SSE build
                int sl = 0, sa = 0, mnl = 32767, mxl = -32768, mna = 32767, mxa = -32768;
                if (has_left) {
                    for (int i = 0; i < 8; ++i) {
                        const int16_t p0 = pix[i * 8], p1 = pix[i * 8 + 1];
                        const int16_t delta = (int16_t)(p0 - p1);
                        const int est = (int16_t)((int16_t)((int16_t)left_edge[i] - half_rz16(delta)) - (int16_t)(p0 + 1024));
                        sl += est; mnl = min(mnl, est); mxl = max(mxl, est);
                    }
                }
                if (has_above) {
                    for (int i = 0; i < 8; ++i) {
                        const int16_t p0 = pix[i], p1 = pix[8 + i];
                        const int16_t delta = (int16_t)(p0 - p1);
                        const int est = (int16_t)((int16_t)((int16_t)redge[(size_t)x * 8 + i] - half_rz16(delta)) - (int16_t)(p0 + 1024));
                        sa += est; mna = min(mna, est); mxa = max(mxa, est);
                    }
                }
                int avgmed = 0;
                if (has_left || has_above) {
                    int a0, a1, mn_all, mx_all;
                    if (has_left && has_above) { a0 = sl; a1 = sa; mn_all = min(mnl, mna); mx_all = max(mxl, mxa); }
                    else if (has_left) { a0 = a1 = sl; mn_all = mnl; mx_all = mxl; }
                    else { a0 = a1 = sa; mn_all = mna; mx_all = mxa; }
                    avgmed = (a0 + a1) >> 1;
                    unc = (mx_all - mn_all) >> 3;
                    a0 -= avgmed; a1 -= avgmed;
                    int far_afield = a1;
                    if (iabs(a0) < iabs(a1)) far_afield = a0;
                    unc2 = far_afield >> 3;
                }
                pred = (div_trunc_small(avgmed, q0) + 4) >> 3;
            }
            int dc;
            {
                const int lm = min(bitlen((uint32_t)iabs(unc) & 0xffff), 11), lo16 = min(bitlen((uint32_t)iabs(unc2) & 0xffff), 16);
                const int sctx = unc2 >= 0 ? (unc2 == 0 ? 3 : 2) : 1;
                const int v = td_coef_plain(d, m_expdc(lm, lo16), m_sign(ci, 0, sctx), m_resdc(lm));
                dc = (int)(int16_t)adv_unpredict((int)(int16_t)v, true, pred);          // decoder.cc:305-309
            }
            rcur[0] = (int16_t)dc;
            // ---- (v) neighbour summaries (block_context.hh:44-78) + store the block in aligned order
            {
                const int16_t qdc = (int16_t)((uint32_t)q0 * (uint32_t)dc);
                for (int i = 0; i < 8; ++i) {
                    {   // right column -> the next block's left neighbour
                        const int16_t cur = pix[i * 8 + 7], prev = pix[i * 8 + 6];
                        const int16_t delta = (int16_t)(cur - prev);
                        left_edge[i] = (int16_t)(cur + half_rz16(delta) + 1024 + qdc);
                    }
                    {   // bottom row -> the block below
                        const int16_t cur = pix[56 + i], prev = pix[48 + i];
                        const int16_t delta = (int16_t)(cur - prev);
                        redge[(size_t)x * 8 + i] = (int16_t)(cur + half_rz16(delta) + 1024 + qdc);
                    }
                }
                rnz[x] = (uint8_t)nz;
                nz_left = nz;
                uint4* dst = reinterpret_cast<uint4*>(rowp + (size_t)x * 64);
#pragma unroll
                for (int v4 = 0; v4 < 8; ++v4) {
                    uint32_t wds[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t lo = (uint16_t)rcur[c_aligned_to_raster[v4 * 8 + 2 * k]];
                        const uint32_t hi = (uint16_t)rcur[c_aligned_to_raster[v4 * 8 + 2 * k + 1]];
                        wds[k] = lo | (hi << 16);
                    }
                    dst[v4] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                }
            }
            if (x + 1 < w && (uint32_t)((size_t)y * w + x + 1) >= (uint32_t)g.trunc_bc[c]) break;
            pc ^= 1; pa ^= 1;
        }
        if (status != ST_OK) break;
    }
    sd.status = status;
    sd.len = (uint32_t)(d.br.p - reinterpret_cast<const uint8_t*>(sd.stream));
    sd.ndecisions_lo = (uint32_t)d.ndec;
    sd.ndecisions_hi = (uint32_t)(d.ndec >> 32);
}

}  // namespace lepb200
