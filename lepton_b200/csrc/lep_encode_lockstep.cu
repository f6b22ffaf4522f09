// lep_encode_lockstep.cu -- sm_100a encode kernel A', one thread per Lepton thread-segment, lanes in LOCK STEP.
//
// Same job as lep_encode_kernel (kernel A: coefficient planes -> (probability, bit) tokens + adaptive model; the range
// coder stays lep_rangecode_kernel), organised like lep_decode_lockstep.cu: every lane walks its own segment, the token
// grammar (serialize_tokens, src/vp8/encoder/encoder.cc:194-402) is a per-lane state machine around ONE copy of the
// adaptive-model step (Branch::record_obs_and_update, src/vp8/model/branch.hh:82-100), and the lanes meet where the scalar
// predictors run (49 priors -> [step: count + 7x7] -> 14 Lakhani predictions -> [step: edge counts + edges] ->
// IDCT + DC estimate -> [step: DC] -> neighbour summaries).  Unlike the decoder nothing here waits on the coder: the
// branch of the NEXT decision does not depend on the model, so its count is requested one step ahead and the model
// latency hides behind the previous step's arithmetic.
//
// Candidate (LEPB200_ENC_MODE=1), pinned on the CPU warp emulator against the reference streams (tests/test_emu_encode.py).
#include "lep_common.cuh"
#include "lep_predict.cuh"

namespace lepb200 {

constexpr int ENCL_THREADS = 32;          // one warp per CTA

enum : int { ES_IDLE = 0, ES_COUNT, ES_EXP, ES_SIGN, ES_THR, ES_RES };
enum : int { EEV_NONE = 0, EEV_COUNT, EEV_COEF };

// Per-lane position in the token grammar: `addr`/`bit` is the pending decision.
struct EMicro {
    int st;
    uint32_t addr, bit;
    // count
    uint32_t cnt_base; int cnt_shift, cnt_idx, cnt_prefix, cnt_val;
    // coefficient: magnitude a (16-bit), exponent tlen <= 11
    uint32_t exp_base, sign_addr, res_base, thr_ctx, a;
    int len, tlen, ri, min_thr, neg;
    uint32_t so;
};

__device__ __forceinline__ void e_start_count(EMicro& m, int value, uint32_t base, int shift, int nbits) {
    m.st = ES_COUNT; m.cnt_base = base; m.cnt_shift = shift; m.cnt_idx = nbits - 1; m.cnt_prefix = 0; m.cnt_val = value;
    m.addr = base + ((uint32_t)(nbits - 1) << shift);
    m.bit = ((uint32_t)value >> (nbits - 1)) & 1u;
}
// returns false when the magnitude does not fit the grammar (COEFFICIENT_OUT_OF_RANGE, encoder.cc:124,265,343)
__device__ __forceinline__ bool e_start_coef(EMicro& m, int v, uint32_t exp_base, uint32_t sign_addr, uint32_t res_base, uint32_t thr_ctx, int min_thr) {
    m.st = ES_EXP; m.exp_base = exp_base; m.sign_addr = sign_addr; m.res_base = res_base; m.thr_ctx = thr_ctx; m.min_thr = min_thr;
    m.a = (uint32_t)iabs(v) & 0xffffu;
    const int raw = bitlen(m.a);
    m.tlen = min(raw, 11);
    m.neg = v < 0;
    m.len = 0; m.addr = exp_base; m.bit = m.tlen > 0;
    return raw <= 11;
}

// the pending decision has been coded: move to the next one (EEV_NONE), or report the end of the count / coefficient
__device__ __forceinline__ int e_advance(EMicro& m) {
    const uint32_t bit = m.bit;
    switch (m.st) {
    case ES_COUNT:
        m.cnt_prefix = (m.cnt_prefix << 1) | (int)bit;
        if (--m.cnt_idx < 0) { m.st = ES_IDLE; return EEV_COUNT; }
        m.addr = m.cnt_base + ((uint32_t)m.cnt_idx << m.cnt_shift) + (uint32_t)m.cnt_prefix;
        m.bit = ((uint32_t)m.cnt_val >> m.cnt_idx) & 1u;
        return EEV_NONE;
    case ES_EXP:
        if (bit && ++m.len < 11) { m.addr = m.exp_base + (uint32_t)m.len; m.bit = m.len < m.tlen; return EEV_NONE; }
        if (m.tlen == 0) break;
        m.st = ES_SIGN; m.addr = m.sign_addr; m.bit = !m.neg;
        return EEV_NONE;
    case ES_SIGN:
        m.ri = m.tlen - 2;
        if (m.ri < 0) break;
        if (m.ri >= m.min_thr) {
            m.st = ES_THR; m.so = 1;
            m.thr_ctx += (uint32_t)min(m.tlen - m.min_thr, 7) << 7;          // m_thr(ci, ctx, len - min_thr)
            m.addr = m.thr_ctx + 1;
        } else {
            m.st = ES_RES; m.addr = m.res_base + (uint32_t)m.ri;
        }
        m.bit = (m.a >> m.ri) & 1u;
        return EEV_NONE;
    case ES_THR:
        m.so = min((m.so << 1) | bit, 127u);
        if (--m.ri < 0) break;
        if (m.ri >= m.min_thr) m.addr = m.thr_ctx + m.so;
        else { m.st = ES_RES; m.addr = m.res_base + (uint32_t)m.ri; }
        m.bit = (m.a >> m.ri) & 1u;
        return EEV_NONE;
    case ES_RES:
        if (--m.ri < 0) break;
        m.addr = m.res_base + (uint32_t)m.ri;
        m.bit = (m.a >> m.ri) & 1u;
        return EEV_NONE;
    default:
        return EEV_NONE;
    }
    m.st = ES_IDLE;
    return EEV_COEF;
}

// Token writer + adaptive model of one lane.  `w` is the count word of the PENDING decision, requested one step early.
struct ELane {
    uint16_t* model;
    uint16_t* tokens;
    uint32_t ntok, tok_cap;
    uint32_t w;
};

// codes the pending decision (cur_addr, cur_bit) whose count word is already in L.w, after requesting the word of the
// decision that follows (next_addr; ~0u = none yet)
__device__ __forceinline__ void e_put(ELane& L, const uint32_t* rcp, uint32_t cur_addr, uint32_t cur_bit, uint32_t next_addr) {
    const uint32_t w = L.w;
    uint32_t wn = next_addr != ~0u ? (uint32_t)L.model[next_addr] : 0u;       // issued before anything below needs `w`
    const uint32_t pb = branch_prob(w, rcp) | (cur_bit << 8);
    const bool plain = (w & 0xffu) < 254u && (w >> 8) < 254u;
    const uint32_t neww = plain ? w + (cur_bit ? 0x100u : 1u) : branch_update(w, cur_bit);
    L.model[cur_addr] = (uint16_t)neww;
    if (L.ntok < L.tok_cap) L.tokens[L.ntok] = (uint16_t)pb;
    L.ntok++;
    if (next_addr == cur_addr) wn = neww & 0xffffu;                         // same branch twice in a row: forward the new count
    L.w = wn;
}

__device__ __forceinline__ int e_lak(const int16_t* cur, const int16_t* nb, const int32_t* __restrict__ icos, int first, int step) {
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

constexpr int E_NO_THR = 127;

__global__ void __launch_bounds__(ENCL_THREADS)
lep_encode_lockstep_kernel(const ImageDesc* __restrict__ images, SegDesc* __restrict__ segs, int first, int count, const int* __restrict__ order,
                           uint16_t* __restrict__ model_pool, uint8_t* __restrict__ row_pool, size_t row_pool_stride,
                           uint16_t* __restrict__ token_base) {
    __shared__ uint32_t s_rcp[512];
    __shared__ uint8_t s_a2r[64];
    __shared__ uint8_t s_nzbin[64];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_rcp[i] = i < 2 ? 0u : (uint32_t)((0x100000000ull + i - 1) / i);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { s_a2r[i] = c_aligned_to_raster[i]; s_nzbin[i] = i < 50 ? c_nonzero_to_bin[i] : 0; }
    __syncthreads();
    const int t = blockIdx.x * ENCL_THREADS + threadIdx.x;
    // no lane leaves before the end: the votes below are over the full warp
    SegDesc* sdp = t < count ? &segs[order[first + t]] : nullptr;
    bool alive = sdp != nullptr && sdp->status == ST_OK;
    const bool report = alive;
    const ImageDesc& g = images[alive ? sdp->image : 0];
    const int seg_min_y = alive ? sdp->min_y : 0, seg_max_y = alive ? sdp->max_y : 0;
    const bool seg_last = alive ? sdp->is_last != 0 : false;
    uint8_t* rowbuf = row_pool + (size_t)(t < count ? t : 0) * row_pool_stride;

    ELane L;
    L.model = model_pool + (size_t)(t < count ? t : 0) * M_TOTAL;          // zero-filled before the launch
    L.tokens = token_base + (alive ? sdp->tokens : 0ull);
    L.ntok = 0; L.tok_cap = alive ? sdp->tok_cap : 0u; L.w = 0;

    const int bw0 = g.bch[0], bw1 = g.ncmp > 1 ? g.bch[1] : 0, bw2 = g.ncmp > 2 ? g.bch[2] : 0;
    const size_t nz_base = (size_t)(bw0 + bw1 + bw2) * 16;
    const int nzs0 = (bw0 + 15) & ~15, nzs1 = (bw1 + 15) & ~15;

    // lane-local raster-order blocks: [0],[1] current / left (ping-pong), [2],[3] above / above-left (ping-pong)
    int16_t blk[4][64];
    int32_t tmp[64];
    int16_t pix[64];
    uint8_t pbsr[49];
    int32_t lak[14];
    int16_t left_edge[8];

    int status = ST_OK;
    uint32_t top_mask = 7u, index = 0;
    int c = 0, ci = 0, y = 0, w = 0, x = 0, q0 = 1, pc = 0, pa = 2, nz_left = 0;
    bool has_above = false;
    const int16_t* rowp = nullptr;
    const int16_t* abovep = nullptr;
    const uint16_t* q = g.q[0];
    int16_t* redge = nullptr;
    uint8_t* rnz = nullptr;
    const int32_t* icx = g.icos_x[0];
    const int32_t* icy = g.icos_y[0];

    bool need_row = true;
    for (;;) {
        // ---- (0) next row of this lane's segment (row iteration of lepton_codec.hh:41-100); a segment that met an
        //          out-of-range coefficient stops at the end of that row
        if (alive && need_row) {
            if (status != ST_OK) alive = false;
            while (alive) {
                const RowSpec rs = row_spec_from_index(index++, g);
                if (rs.done || (rs.luma_y >= seg_max_y && !seg_last)) { alive = false; break; }
                if (rs.skip || rs.luma_y < seg_min_y) continue;
                c = rs.component; y = rs.curr_y;
                has_above = !((top_mask >> c) & 1u);
                top_mask &= ~(1u << c);
                ci = c == 0 ? 0 : 1;
                w = g.bch[c];
                rowp = reinterpret_cast<const int16_t*>(g.plane[c]) + (size_t)y * w * 64;
                abovep = rowp - (size_t)w * 64;
                q = g.q[c];
                q0 = q[0];
                redge = reinterpret_cast<int16_t*>(rowbuf + (size_t)(c == 0 ? 0 : (c == 1 ? bw0 : bw0 + bw1)) * 16);
                rnz = rowbuf + nz_base + (c == 0 ? 0 : (c == 1 ? nzs0 : nzs0 + nzs1));
                icx = g.icos_x[c];
                icy = g.icos_y[c];
                x = 0; pc = 0; pa = 2; nz_left = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) left_edge[i] = 0;
                need_row = false;
                break;
            }
        }
        if (!__any_sync(FULL, alive)) break;

        const bool has_left = x > 0;
        int16_t* rcur = blk[pc];
        const int16_t* rleft = blk[pc ^ 1];
        int16_t* rabove = blk[pa];
        const int16_t* raleft = blk[pa ^ 1];
        EMicro m;
        m.st = ES_IDLE; m.addr = 0; m.bit = 0;
        int nz = 0, eobx = 0, eoby = 0, ne_h = 0, ne_v = 0;

        // ---- (1) this block and the one above -> raster copies; priors of the 7x7 coefficients; non-zero count
        if (alive) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half == 1 && !has_above) break;
                const uint4* src = reinterpret_cast<const uint4*>((half ? abovep : rowp) + (size_t)x * 64);
                int16_t* dst = half ? rabove : rcur;
#pragma unroll
                for (int v4 = 0; v4 < 8; ++v4) {
                    const uint4 u = src[v4];
                    const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        dst[c_aligned_to_raster[v4 * 8 + 2 * k]] = (int16_t)(wds[k] & 0xffff);
                        dst[c_aligned_to_raster[v4 * 8 + 2 * k + 1]] = (int16_t)(wds[k] >> 16);
                    }
                }
            }
            // compute_aavrg (model.hh:895-924) for all 49 positions, the neighbour case decided once per block
            if (has_left && has_above) {
#pragma unroll
                for (int zz = 0; zz < 49; ++zz) {
                    const int coord = c_aligned_to_raster[zz];
                    const uint32_t Lm = (uint32_t)iabs(rleft[coord]) & 0xffff, A = (uint32_t)iabs(rabove[coord]) & 0xffff;
                    const uint32_t tt = ((Lm + A) * 13u + (((uint32_t)iabs(raleft[coord]) & 0xffff) * 6u)) & 0xffff;
                    pbsr[zz] = (uint8_t)bitlen(min(tt >> 5, 1023u));
                }
            } else if (has_left || has_above) {
                const int16_t* nb = has_left ? rleft : rabove;
#pragma unroll
                for (int zz = 0; zz < 49; ++zz) {
                    const int prior = (int)(int16_t)((uint32_t)iabs(nb[c_aligned_to_raster[zz]]) & 0xffff);
                    pbsr[zz] = (uint8_t)bitlen((uint32_t)min(iabs(prior), 1023));
                }
            } else {
#pragma unroll
                for (int zz = 0; zz < 49; ++zz) pbsr[zz] = 0;
            }
#pragma unroll
            for (int zz = 0; zz < 49; ++zz) nz += rcur[c_aligned_to_raster[zz]] != 0;
#pragma unroll
            for (int k = 1; k < 8; ++k) { ne_h += rcur[k] != 0; ne_v += rcur[8 * k] != 0; }
            const int nz_above = has_above ? (int)rnz[x] : 0;
            int ctx = 0;
            if (has_above && !has_left) ctx = (nz_above + 1) / 2;
            else if (has_left && !has_above) ctx = (nz_left + 1) / 2;
            else if (has_left && has_above) ctx = (nz_above + nz_left + 2) / 4;
            e_start_count(m, nz, m_nz7(ci, s_nzbin[ctx], 0, 0), 5, 6);
            L.w = L.model[m.addr];
        }

        // ---- (2) step: 7x7 non-zero count, then the 7x7 coefficients in zig-zag order while non-zeros remain
        {
            int zz = 0, left_nz = nz;
            bool busy = alive;
            while (__any_sync(FULL, busy)) {
                if (busy) {
                    const uint32_t cur_addr = m.addr, cur_bit = m.bit;
                    const int ev = e_advance(m);
                    if (ev != EEV_NONE) {
                        bool next = true;
                        if (ev == EEV_COUNT) {
                            if (nz == 0) next = false;
                        } else {
                            const int coord = s_a2r[zz];
                            if (rcur[coord] != 0) {
                                --left_nz;
                                eobx = max(eobx, coord & 7); eoby = max(eoby, coord >> 3);
                            }
                            ++zz;
                            if (left_nz == 0 || zz == 49) next = false;
                        }
                        if (next) {
                            const int bin = s_nzbin[left_nz];
                            const int coord = s_a2r[zz];
                            if (!e_start_coef(m, rcur[coord], m_exp7(ci, bin, zz, pbsr[zz]), m_sign(ci, 0, 0), m_resn(ci, coord, bin), 0, E_NO_THR))
                                status = ST_COEF_RANGE;
                        } else {
                            busy = false;
                        }
                    }
                    e_put(L, s_rcp, cur_addr, cur_bit, busy ? m.addr : ~0u);
                }
            }
        }

        // ---- (3) Lakhani predictions of the 14 edge coefficients (model.hh:1033-1071)
        if (alive) {
#pragma unroll
            for (int k = 1; k < 8; ++k) lak[k - 1] = has_above ? e_lak(rcur, rabove, icx + k * 8, k, 8) : 0;
#pragma unroll
            for (int k = 1; k < 8; ++k) lak[6 + k] = has_left ? e_lak(rcur, rleft, icy + k * 8, 8 * k, 1) : 0;
            e_start_count(m, ne_h, m_nze(0, ci, eobx, (nz + 3) / 7, 0, 0), 2, 3);
            L.w = L.model[m.addr];
        }

        // ---- (4) step: horizontal edge (raster 1..7), then vertical edge (raster 8..56): count, then coefficients
        {
            int vert = 0, ne = 0, ln = 0;
            bool busy = alive;
            while (__any_sync(FULL, busy)) {
                if (busy) {
                    const uint32_t cur_addr = m.addr, cur_bit = m.bit;
                    const int ev = e_advance(m);
                    if (ev != EEV_NONE) {
                        bool more;
                        if (ev == EEV_COUNT) {
                            ne = vert ? ne_v : ne_h; ln = 0;
                            more = ne > 0;
                        } else {
                            if (rcur[vert ? 8 * (ln + 1) : ln + 1] != 0) --ne;
                            ++ln;
                            more = ne > 0 && ln < 7;
                        }
                        if (more) {
                            const int coord = vert ? 8 * (ln + 1) : ln + 1;
                            const int prior = lak[vert * 7 + ln];
                            const int bsr = bitlen((uint32_t)min(iabs(prior), 1023));
                            const int p16 = (int)(int16_t)prior;
                            const int sctx = p16 == 0 ? 0 : (p16 > 0 ? 1 : 2);
                            const int min_thr = g.min_thr[c][coord];
                            const int ctx_abs = iabs(prior) & 0xffff;
                            if (!e_start_coef(m, rcur[coord], m_expx(ci, ne, vert ? 7 + ln : ln, bsr), m_sign(ci, sctx, bsr), m_resn(ci, coord, ne),
                                              m_thr(ci, min(ctx_abs >> min_thr, 255), 0), min_thr))
                                status = ST_COEF_RANGE;
                        } else if (vert == 0) {
                            vert = 1;
                            e_start_count(m, ne_v, m_nze(1, ci, eoby, (nz + 3) / 7, 0, 0), 2, 3);
                        } else {
                            busy = false;
                        }
                    }
                    e_put(L, s_rcp, cur_addr, cur_bit, busy ? m.addr : ~0u);
                }
            }
        }

        // ---- (5) DC: pixels of the block without its DC, prediction from the neighbours' edge pixels (encoder.cc:293-364)
        if (alive) {
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
            int avgmed = 0, unc = 0, unc2 = 0;
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
            const int pred = (div_trunc_small(avgmed, q0) + 4) >> 3;
            const int dc = rcur[0];
            const int adv = adv_unpredict(dc, false, pred);
            if (dc != adv_unpredict((int)(int16_t)adv, true, pred)) status = ST_COEF_RANGE;
            const int lm = min(bitlen((uint32_t)iabs(unc) & 0xffff), 11), lo16 = min(bitlen((uint32_t)iabs(unc2) & 0xffff), 16);
            const int sctx = unc2 >= 0 ? (unc2 == 0 ? 3 : 2) : 1;
            e_start_coef(m, (int)(int16_t)adv, m_expdc(lm, lo16), m_sign(ci, 0, sctx), m_resdc(lm), 0, E_NO_THR);
            L.w = L.model[m.addr];
        }

        // ---- (6) step: the DC coefficient
        {
            bool busy = alive;
            while (__any_sync(FULL, busy)) {
                if (busy) {
                    const uint32_t cur_addr = m.addr, cur_bit = m.bit;
                    if (e_advance(m) == EEV_COEF) busy = false;
                    e_put(L, s_rcp, cur_addr, cur_bit, busy ? m.addr : ~0u);
                }
            }
        }

        // ---- (7) neighbour summaries (block_context.hh:44-78), next block
        if (alive) {
            const int dc = rcur[0];
            const int16_t qdc = (int16_t)((uint32_t)q0 * (uint32_t)dc);
            for (int i = 0; i < 8; ++i) {
                {
                    const int16_t cur = pix[i * 8 + 7], prev = pix[i * 8 + 6];
                    const int16_t delta = (int16_t)(cur - prev);
                    left_edge[i] = (int16_t)(cur + half_rz16(delta) + 1024 + qdc);
                }
                {
                    const int16_t cur = pix[56 + i], prev = pix[48 + i];
                    const int16_t delta = (int16_t)(cur - prev);
                    redge[(size_t)x * 8 + i] = (int16_t)(cur + half_rz16(delta) + 1024 + qdc);
                }
            }
            rnz[x] = (uint8_t)nz;
            nz_left = nz;
            // a block at or past the truncation bound is not coded unless it is the first of its row (vp8_encoder.cc:110-113,133-135)
            if (x + 1 >= w || (uint32_t)((size_t)y * w + x + 1) >= (uint32_t)g.trunc_bc[c]) need_row = true;
            else { ++x; pc ^= 1; pa ^= 1; }
        }
    }
    if (report) {
        if (status == ST_OK && L.ntok > L.tok_cap) status = ST_OUT_OVERFLOW;
        sdp->ntok = L.ntok;
        sdp->len = 0;
        sdp->status = status;
        sdp->ndecisions_lo = L.ntok;          // one token per decision
        sdp->ndecisions_hi = 0;
    }
}

}  // namespace lepb200
