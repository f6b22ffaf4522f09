// lep_decode_lockstep.cu -- sm_100a decode kernel, one thread per Lepton thread-segment, lanes in LOCK STEP.
//
// The warp-per-segment kernel (lep_decode.cu) spends 32 lanes on one serial chain; the plain thread-per-segment kernel
// (lep_decode_thread.cu) gives every lane a chain but lets the lanes drift apart inside the token grammar's nested loops,
// so the warp executes each lane's path one after the other.  Here the 32 chains of a warp advance together:
//
//   * the token grammar (parse_tokens, src/vp8/decoder/decoder.cc:167-318) is a per-lane state machine; ONE copy of the
//     bool decoder (vpx_reader, src/vp8/decoder/boolreader.hh:184-258,376-416 + Branch::record_obs_and_update,
//     src/vp8/model/branch.hh:82-100) sits in a loop that every lane with a pending decision runs at the same time, and
//     the lane's state picks the branch address and what the decoded bit means;
//   * the lanes meet at the three points of a block where the scalar predictors run, so those execute once for 32 blocks:
//     priors of the 49 inner coefficients (compute_aavrg, model.hh:895-924) -> [step: non-zero count + 7x7] ->
//     14 Lakhani edge predictions (compute_lak, model.hh:1033-1071) -> [step: edge counts + edges] ->
//     8x8 IDCT + DC estimate (adv_predict_dc_pix, model.hh:674-784) -> [step: DC] -> neighbour summaries + block store.
//
// A lane whose block needs fewer decisions than its neighbours' idles until the slowest lane of the warp is through
// the phase; segments are handed out in order of decreasing size so that the lanes of a warp run out together.
//
// Same job descriptors, model layout (one zero-filled 1.58 MB model per lane) and launch shape as the plain
// thread-per-segment kernel.  The lanes exchange nothing but votes, which is what lets tests/emu run the body one lane
// at a time on the CPU against the oracle.
#include "lep_common.cuh"
#include "lep_predict.cuh"

namespace lepb200 {

constexpr int DECL_THREADS = 32;          // one warp per CTA

struct LBool {                            // vpx_reader (boolreader.hh:184-258), per lane
    unsigned long long value;             // stream bits, left aligned
    uint32_t range;
    int valid;                            // bits of `value` that come from the stream (the rest are zero)
    const uint8_t* p;                     // next 32-bit word of the stream (streams start 16-byte aligned)
    const uint8_t* end;
};

// The reference tops its window up byte by byte whenever fewer than 8 bits are left (vpx_reader_fill); what a decision
// sees is only the top byte of the window, and bits past the end of the stream read as zero.  Here the window takes one
// aligned big-endian 32-bit word whenever fewer than 32 bits are left: same bits in the same positions, and a refill is a
// handful of instructions -- it matters because in a lock-step warp ANY lane's refill is paid by all 32.
__device__ __forceinline__ void l_refill(LBool& r) {
    uint32_t w = 0;
    const long long rem = r.end - r.p;
    if (rem > 0) {
        w = __byte_perm(__ldg(reinterpret_cast<const uint32_t*>(r.p)), 0u, 0x0123u);
        if (rem < 4) w &= 0xffffffffu << (8 * (4 - (int)rem));        // the padding behind a stream is readable but not zero
    }
    r.value |= (unsigned long long)w << (32 - r.valid);
    r.valid += 32;
    r.p += 4;
}

// VPXBoolReader::get (vpx_bool_reader.hh:45-57) = vpx_read + Branch::record_obs_and_update
__device__ __forceinline__ uint32_t l_get(LBool& r, uint16_t* model, const uint32_t* rcp, uint32_t addr) {
    const uint32_t w = model[addr];
#ifdef LEPB200_EMU_TRACE                 // tests/tools_model_reuse.py: cache footprint of a lane's model accesses
    emu_trace_model_access(model, addr);
#endif
    const uint32_t prob = branch_prob(w, rcp);
    const uint32_t split = (r.range * prob + (256 - prob)) >> 8;
    if (r.valid < 32) l_refill(r);
    const uint32_t top = (uint32_t)(r.value >> 56);               // value >= split << 56  <=>  top byte >= split
    const uint32_t bit = top >= split;
    const uint32_t range = bit ? r.range - split : split;
    if (bit) r.value -= (unsigned long long)split << 56;
    const int shift = __clz(range) - 24;
    r.range = range << shift;
    r.value <<= shift;
    r.valid -= shift;
    const bool plain = (w & 0xffu) < 254u && (w >> 8) < 254u;       // no count about to saturate, not the special state
    model[addr] = (uint16_t)(plain ? w + (bit ? 0x100u : 1u) : branch_update(w, bit));
    return bit;
}

// Per-lane position in the token grammar.  A count is read MSB first with the bits so far as context; a coefficient is
// exponent (unary, <= 11) / sign / value bits, the top value bits of an edge coefficient going through the threshold
// tables (decoder.cc:212-240, 257-300).
enum : int { LS_IDLE = 0, LS_COUNT, LS_EXP, LS_SIGN, LS_THR, LS_RES };
enum : int { LEV_NONE = 0, LEV_COUNT, LEV_COEF };

struct LMicro {
    int st;
    uint32_t addr;                        // branch of the pending decision
    // count
    uint32_t cnt_base; int cnt_shift, cnt_idx, cnt_prefix, cnt_val;
    // coefficient
    uint32_t exp_base, sign_addr, res_base, thr_ctx;
    int len, val, ri, min_thr, neg;
    uint32_t so;
};

__device__ __forceinline__ void l_start_count(LMicro& m, uint32_t base, int shift, int nbits) {
    m.st = LS_COUNT; m.cnt_base = base; m.cnt_shift = shift; m.cnt_idx = nbits - 1; m.cnt_prefix = 0; m.cnt_val = 0;
    m.addr = base + ((uint32_t)(nbits - 1) << shift);
}
__device__ __forceinline__ void l_start_coef(LMicro& m, uint32_t exp_base, uint32_t sign_addr, uint32_t res_base, uint32_t thr_ctx, int min_thr) {
    m.st = LS_EXP; m.exp_base = exp_base; m.sign_addr = sign_addr; m.res_base = res_base; m.thr_ctx = thr_ctx; m.min_thr = min_thr;
    m.len = 0; m.addr = exp_base;
}

// consumes one decoded bit; LEV_COUNT: m.cnt_val is complete, LEV_COEF: `value` is the coefficient (0 <=> exponent 0)
__device__ __forceinline__ int l_advance(LMicro& m, uint32_t bit, int& value) {
    switch (m.st) {
    case LS_COUNT:
        m.cnt_val |= (int)bit << m.cnt_idx;
        m.cnt_prefix = (m.cnt_prefix << 1) | (int)bit;
        if (--m.cnt_idx < 0) { m.st = LS_IDLE; return LEV_COUNT; }
        m.addr = m.cnt_base + ((uint32_t)m.cnt_idx << m.cnt_shift) + (uint32_t)m.cnt_prefix;
        return LEV_NONE;
    case LS_EXP:
        if (bit && ++m.len < 11) { m.addr = m.exp_base + (uint32_t)m.len; return LEV_NONE; }
        if (m.len == 0) { m.st = LS_IDLE; value = 0; return LEV_COEF; }
        m.st = LS_SIGN; m.addr = m.sign_addr;
        return LEV_NONE;
    case LS_SIGN:
        m.neg = !bit;
        m.val = 1 << (m.len - 1);
        m.ri = m.len - 2;
        if (m.ri < 0) break;
        if (m.ri >= m.min_thr) {
            m.st = LS_THR; m.so = 1;
            m.thr_ctx += (uint32_t)min(m.len - m.min_thr, 7) << 7;          // m_thr(ci, ctx, len - min_thr)
            m.addr = m.thr_ctx + 1;
        } else {
            m.st = LS_RES; m.addr = m.res_base + (uint32_t)m.ri;
        }
        return LEV_NONE;
    case LS_THR:
        m.val |= (int)bit << m.ri;
        m.so = min((m.so << 1) | bit, 127u);
        if (--m.ri < 0) break;
        if (m.ri >= m.min_thr) m.addr = m.thr_ctx + m.so;
        else { m.st = LS_RES; m.addr = m.res_base + (uint32_t)m.ri; }
        return LEV_NONE;
    case LS_RES:
        m.val |= (int)bit << m.ri;
        if (--m.ri < 0) break;
        m.addr = m.res_base + (uint32_t)m.ri;
        return LEV_NONE;
    default:
        return LEV_NONE;
    }
    m.st = LS_IDLE;
    value = m.neg ? -m.val : m.val;
    return LEV_COEF;
}

// compute_lak (model.hh:1033-1071), scalar; same arithmetic as lak_pred in lep_predict.cuh
__device__ __forceinline__ int l_lak(const int16_t* cur, const int16_t* nb, const int32_t* __restrict__ icos, int first, int step) {
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

constexpr int L_NO_THR = 127;             // min_thr of coefficients without threshold bits (7x7, DC)

__global__ void __launch_bounds__(DECL_THREADS)
lep_decode_lockstep_kernel(const ImageDesc* __restrict__ images, SegDesc* __restrict__ segs, int first, int count, const int* __restrict__ order,
                           uint16_t* __restrict__ model_pool, uint8_t* __restrict__ row_pool, size_t row_pool_stride) {
    __shared__ uint32_t s_rcp[512];
    __shared__ uint8_t s_a2r[64];         // aligned index -> raster index; lanes look up different entries (constant memory would serialise)
    __shared__ uint8_t s_nzbin[64];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_rcp[i] = i < 2 ? 0u : (uint32_t)((0x100000000ull + i - 1) / i);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { s_a2r[i] = c_aligned_to_raster[i]; s_nzbin[i] = i < 50 ? c_nonzero_to_bin[i] : 0; }
    __syncthreads();
    const int t = blockIdx.x * DECL_THREADS + threadIdx.x;
    // no lane leaves before the end: the votes below are over the full warp
    SegDesc* sdp = t < count ? &segs[order[first + t]] : nullptr;
    bool alive = sdp != nullptr && sdp->status == ST_OK;          // else rejected on the host (e.g. zero quantiser, model.hh:257-262)
    const bool report = alive;
    const ImageDesc& g = images[alive ? sdp->image : 0];
    const int seg_min_y = alive ? sdp->min_y : 0, seg_max_y = alive ? sdp->max_y : 0;
    const bool seg_last = alive ? sdp->is_last != 0 : false;
    uint8_t* rowbuf = row_pool + (size_t)(t < count ? t : 0) * row_pool_stride;
    uint16_t* model = model_pool + (size_t)(t < count ? t : 0) * M_TOTAL;       // zero-filled before the launch

    LBool br;
    br.value = 0; br.valid = 0; br.range = 255; br.p = nullptr; br.end = nullptr;
    unsigned long long ndec = 0;
    if (alive) {
        br.p = reinterpret_cast<const uint8_t*>(sdp->stream); br.end = br.p + sdp->cap;
        l_refill(br);
        // marker bit at p = 128 (boolreader.cc:26-35); no model involved
        const uint32_t split = (br.range * 128u + 128u) >> 8;
        const uint32_t bit = (uint32_t)(br.value >> 56) >= split;
        const uint32_t range = bit ? br.range - split : split;
        if (bit) br.value -= (unsigned long long)split << 56;
        const int shift = __clz(range) - 24;
        br.range = range << shift; br.value <<= shift; br.valid -= shift;
    }

    const int bw0 = g.bch[0], bw1 = g.ncmp > 1 ? g.bch[1] : 0, bw2 = g.ncmp > 2 ? g.bch[2] : 0;
    const size_t nz_base = (size_t)(bw0 + bw1 + bw2) * 16;
    const int nzs0 = (bw0 + 15) & ~15, nzs1 = (bw1 + 15) & ~15;

    // lane-local raster-order blocks: [0],[1] current / left (ping-pong), [2],[3] above / above-left (ping-pong)
    int16_t blk[4][64];
    int32_t tmp[64];
    int16_t pix[64];
    uint8_t pbsr[49];                     // 7x7: bit length of each coefficient's neighbour prior
    int32_t lak[14];                      // edge predictions: 7 horizontal, 7 vertical
    int16_t left_edge[8];                 // right-column edge prediction of the left neighbour

    // ---- row / block cursor of this lane (row iteration of lepton_codec.hh:41-100)
    int status = ST_OK;
    uint32_t top_mask = 7u, index = 0;
    int c = 0, ci = 0, y = 0, w = 0, x = 0, q0 = 1, pc = 0, pa = 2, nz_left = 0;
    bool has_above = false;
    int16_t* rowp = nullptr;
    const int16_t* abovep = nullptr;
    const uint16_t* q = g.q[0];
    int16_t* redge = nullptr;
    uint8_t* rnz = nullptr;
    const int32_t* icx = g.icos_x[0];
    const int32_t* icy = g.icos_y[0];

    bool need_row = true;
    for (;;) {
        // ---- (0) move to the next row when the previous one is finished (per lane, once per row)
        if (alive && need_row) {
            for (;;) {
                const RowSpec rs = row_spec_from_index(index++, g);
                if (rs.done || (rs.luma_y >= seg_max_y && !seg_last)) { alive = false; break; }
                if (rs.skip || rs.luma_y < seg_min_y) continue;
                c = rs.component; y = rs.curr_y;
                has_above = !((top_mask >> c) & 1u);
                top_mask &= ~(1u << c);
                ci = c == 0 ? 0 : 1;
                w = g.bch[c];
                rowp = reinterpret_cast<int16_t*>(g.plane[c]) + (size_t)y * w * 64;
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
        LMicro m;
        m.st = LS_IDLE; m.addr = 0;
        int nz = 0, eobx = 0, eoby = 0;

        // ---- (1) above block -> raster copy, clear the current block, priors of the 7x7 coefficients, count context
        if (alive) {
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
            // compute_aavrg (model.hh:895-924) for all 49 positions, the neighbour case decided once per block
            if (has_left && has_above) {
#pragma unroll
                for (int zz = 0; zz < 49; ++zz) {
                    const int coord = c_aligned_to_raster[zz];
                    const uint32_t L = (uint32_t)iabs(rleft[coord]) & 0xffff, A = (uint32_t)iabs(rabove[coord]) & 0xffff;
                    const uint32_t t = ((L + A) * 13u + (((uint32_t)iabs(raleft[coord]) & 0xffff) * 6u)) & 0xffff;
                    pbsr[zz] = (uint8_t)bitlen(min(t >> 5, 1023u));
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
            const int nz_above = has_above ? (int)rnz[x] : 0;
            int ctx = 0;
            if (has_above && !has_left) ctx = (nz_above + 1) / 2;
            else if (has_left && !has_above) ctx = (nz_left + 1) / 2;
            else if (has_left && has_above) ctx = (nz_above + nz_left + 2) / 4;
            l_start_count(m, m_nz7(ci, s_nzbin[ctx], 0, 0), 5, 6);
        }

        // ---- (2) step: 7x7 non-zero count, then the 7x7 coefficients in zig-zag order (== aligned order 0..48)
        {
            int zz = 0, left_nz = 0;
            bool busy = alive;
            while (__any_sync(FULL, busy)) {
                if (busy) {
                    const uint32_t bit = l_get(br, model, s_rcp, m.addr);
                    ++ndec;
                    int v = 0;
                    const int ev = l_advance(m, bit, v);
                    if (ev != LEV_NONE) {
                        bool next = true;
                        if (ev == LEV_COUNT) {
                            nz = m.cnt_val;
                            left_nz = nz;
                            if (nz > 49) { status = ST_STREAM_INCONSISTENT; alive = false; }
                            if (nz > 49 || nz == 0) next = false;
                        } else {
                            if (v != 0) {
                                const int coord = s_a2r[zz];
                                --left_nz;
                                eobx = max(eobx, coord & 7); eoby = max(eoby, coord >> 3);
                                rcur[coord] = (int16_t)v;
                            }
                            ++zz;
                            if (left_nz == 0 || zz == 49) next = false;
                        }
                        if (next) {
                            const int bin = s_nzbin[left_nz];
                            l_start_coef(m, m_exp7(ci, bin, zz, pbsr[zz]), m_sign(ci, 0, 0), m_resn(ci, s_a2r[zz], bin), 0, L_NO_THR);
                        } else {
                            busy = false;
                        }
                    }
                }
            }
        }

        // ---- (3) Lakhani predictions of the 14 edge coefficients: they read the 7x7 part of this block and the
        //          neighbours only, never the other edge (model.hh:1033-1071)
        if (alive) {
#pragma unroll
            for (int k = 1; k < 8; ++k) lak[k - 1] = has_above ? l_lak(rcur, rabove, icx + k * 8, k, 8) : 0;
#pragma unroll
            for (int k = 1; k < 8; ++k) lak[6 + k] = has_left ? l_lak(rcur, rleft, icy + k * 8, 8 * k, 1) : 0;
            l_start_count(m, m_nze(0, ci, eobx, (nz + 3) / 7, 0, 0), 2, 3);
        }

        // ---- (4) step: horizontal edge (raster 1..7), then vertical edge (raster 8..56): count, then coefficients
        {
            int vert = 0, ne = 0, ln = 0;
            bool busy = alive;
            while (__any_sync(FULL, busy)) {
                if (busy) {
                    const uint32_t bit = l_get(br, model, s_rcp, m.addr);
                    ++ndec;
                    int v = 0;
                    const int ev = l_advance(m, bit, v);
                    if (ev != LEV_NONE) {
                        bool more;                                  // another coefficient of this edge follows
                        if (ev == LEV_COUNT) {
                            ne = m.cnt_val; ln = 0;
                            more = ne > 0;
                        } else {
                            if (v != 0) {
                                rcur[vert ? 8 * (ln + 1) : ln + 1] = (int16_t)v;
                                --ne;
                            }
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
                            l_start_coef(m, m_expx(ci, ne, vert ? 7 + ln : ln, bsr), m_sign(ci, sctx, bsr), m_resn(ci, coord, ne),
                                         m_thr(ci, min(ctx_abs >> min_thr, 255), 0), min_thr);
                        } else if (vert == 0) {
                            vert = 1;
                            l_start_count(m, m_nze(1, ci, eoby, (nz + 3) / 7, 0, 0), 2, 3);
                        } else {
                            busy = false;
                        }
                    }
                }
            }
        }

        // ---- (5) DC: pixels of the block without its DC, prediction from the neighbours' edge pixels
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
            tmp[0] = pred;                                        // kept for step (6) (tmp is free after the IDCT)
            const int lm = min(bitlen((uint32_t)iabs(unc) & 0xffff), 11), lo16 = min(bitlen((uint32_t)iabs(unc2) & 0xffff), 16);
            const int sctx = unc2 >= 0 ? (unc2 == 0 ? 3 : 2) : 1;
            l_start_coef(m, m_expdc(lm, lo16), m_sign(ci, 0, sctx), m_resdc(lm), 0, L_NO_THR);
        }

        // ---- (6) step: the DC coefficient
        int dcv = 0;
        {
            bool busy = alive;
            while (__any_sync(FULL, busy)) {
                if (busy) {
                    const uint32_t bit = l_get(br, model, s_rcp, m.addr);
                    ++ndec;
                    int v = 0;
                    if (l_advance(m, bit, v) == LEV_COEF) { dcv = v; busy = false; }
                }
            }
        }

        // ---- (7) neighbour summaries (block_context.hh:44-78), block store in aligned order, next block
        if (alive) {
            const int dc = (int)(int16_t)adv_unpredict((int)(int16_t)dcv, true, tmp[0]);          // decoder.cc:305-309
            rcur[0] = (int16_t)dc;
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
            // a truncated image ends inside a row (component_size_in_blocks)
            if (x + 1 >= w || (uint32_t)((size_t)y * w + x + 1) >= (uint32_t)g.trunc_bc[c]) need_row = true;
            else { ++x; pc ^= 1; pa ^= 1; }
        }
    }
    if (report) {
        sdp->status = status;
        sdp->len = (uint32_t)(br.p - reinterpret_cast<const uint8_t*>(sdp->stream));        // whole words, so up to 7 bytes past the other kernels' figure
        sdp->ndecisions_lo = (uint32_t)ndec;
        sdp->ndecisions_hi = (uint32_t)(ndec >> 32);
    }
}

}  // namespace lepb200
