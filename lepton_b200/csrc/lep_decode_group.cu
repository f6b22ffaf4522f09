// lep_decode_group.cu -- sm_100a decode kernel: G lanes per Lepton thread-segment, 32/G segments per warp in LOCK STEP.
//
// Decoding is a serial chain per thread-segment (every context depends on decoded values; parse_tokens,
// src/vp8/decoder/decoder.cc:167-318), and a batch holds only a few thousand chains.  One warp per chain
// (lep_decode.cu) spends 32 lanes on a scalar chain and is bound by instruction issue; one thread per chain leaves the
// SMs with less than one warp per scheduler and scalar 8x8 transforms.  Here a warp carries S = 32/G chains:
//
//   * the G lanes of a group hold the same decoder state and execute the chain redundantly (free in SIMT), so every
//     warp instruction of the bool decoder (vpx_reader, src/vp8/decoder/boolreader.hh:184-258,376-416 +
//     Branch::record_obs_and_update, src/vp8/model/branch.hh:82-100) advances S chains;
//   * the token grammar is a per-group state machine around ONE copy of the decoder step, so groups that stand at
//     different points of the grammar still share every instruction;
//   * the groups meet at the three points of a block where the data-parallel predictors run, and those use the G lanes:
//     49 neighbour priors (compute_aavrg, model.hh:895-924) -> [steps: non-zero count + 7x7] -> 14 Lakhani edge
//     predictions (compute_lak, model.hh:1033-1071) -> [steps: edge counts + edges] -> 8x8 IDCT + DC estimate
//     (adv_predict_dc_pix, model.hh:674-784) -> [steps: DC] -> neighbour summaries + block store.
//
// Groups pull thread-segments from a queue (largest first); every segment has its own zero-filled model
// (cudaMemsetAsync before the launch), so taking the next segment never stalls the other groups of the warp.
#include "lep_common.cuh"
#include "lep_predict.cuh"

namespace lepb200 {

struct GBool {                            // vpx_reader (boolreader.hh:184-258), identical in the G lanes of a group
    unsigned long long value;             // stream bits, left aligned
    uint32_t range;
    int valid;                            // bits of `value` that come from the stream (the rest are zero)
    const uint8_t* p;                     // next 32-bit word of the stream (streams start 16-byte aligned)
    const uint8_t* end;
};

// The reference tops its window up byte by byte (vpx_reader_fill); a decision only looks at the top byte and bits past
// the end read as zero.  Here: one aligned big-endian 32-bit word whenever fewer than 32 bits are left.
__device__ __forceinline__ void g_refill(GBool& r) {
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

// VPXBoolReader::get (vpx_bool_reader.hh:45-57) = vpx_read + Branch::record_obs_and_update, given the branch word `w`
// the caller loaded from model[addr].  (Load and use are separate so that the CPU warp emulator, where lanes run one
// after the other, can put a barrier between the group's reads and its first write-back; on the device the converged
// lanes of a group read before any of them writes anyway.)
__device__ __forceinline__ uint32_t g_get(GBool& r, uint16_t* model, const uint32_t* rcp, uint32_t addr, uint32_t w) {
    const uint32_t prob = branch_prob(w, rcp);
    const uint32_t split = (r.range * prob + (256 - prob)) >> 8;
    if (r.valid < 32) g_refill(r);
    const uint32_t top = (uint32_t)(r.value >> 56);               // value >= split << 56  <=>  top byte >= split
    const uint32_t bit = top >= split;
    const uint32_t range = bit ? r.range - split : split;
    if (bit) r.value -= (unsigned long long)split << 56;
    const int shift = __clz(range) - 24;
    r.range = range << shift;
    r.value <<= shift;
    r.valid -= shift;
    const bool plain = (w & 0xffu) < 254u && (w >> 8) < 254u;       // no count about to saturate, not the special state
    model[addr] = (uint16_t)(plain ? w + (bit ? 0x100u : 1u) : branch_update(w, bit));     // all lanes of the group store the same value
    return bit;
}
#ifdef LEPB200_EMU
#define G_EMU_BARRIER() __syncwarp()
#else
#define G_EMU_BARRIER()
#endif

// Per-group position in the token grammar.  A count is read MSB first with the bits so far as context; a coefficient is
// exponent (unary, <= 11) / sign / value bits, the top value bits of an edge coefficient going through the threshold
// tables (decoder.cc:212-240, 257-300).
enum : int { GS_IDLE = 0, GS_COUNT, GS_EXP, GS_SIGN, GS_THR, GS_RES };
enum : int { GEV_NONE = 0, GEV_COUNT, GEV_COEF };

struct GMicro {
    int st;
    uint32_t addr;                        // branch of the pending decision
    uint32_t cnt_base; int cnt_shift, cnt_idx, cnt_prefix, cnt_val;
    uint32_t exp_base, sign_addr, res_base, thr_ctx;
    int len, val, ri, min_thr, neg;
    uint32_t so;
};

__device__ __forceinline__ void g_start_count(GMicro& m, uint32_t base, int shift, int nbits) {
    m.st = GS_COUNT; m.cnt_base = base; m.cnt_shift = shift; m.cnt_idx = nbits - 1; m.cnt_prefix = 0; m.cnt_val = 0;
    m.addr = base + ((uint32_t)(nbits - 1) << shift);
}
__device__ __forceinline__ void g_start_coef(GMicro& m, uint32_t exp_base, uint32_t sign_addr, uint32_t res_base, uint32_t thr_ctx, int min_thr) {
    m.st = GS_EXP; m.exp_base = exp_base; m.sign_addr = sign_addr; m.res_base = res_base; m.thr_ctx = thr_ctx; m.min_thr = min_thr;
    m.len = 0; m.addr = exp_base;
}

// consumes one decoded bit; GEV_COUNT: m.cnt_val is complete, GEV_COEF: `value` is the coefficient (0 <=> exponent 0)
__device__ __forceinline__ int g_advance(GMicro& m, uint32_t bit, int& value) {
    switch (m.st) {
    case GS_COUNT:
        m.cnt_val |= (int)bit << m.cnt_idx;
        m.cnt_prefix = (m.cnt_prefix << 1) | (int)bit;
        if (--m.cnt_idx < 0) { m.st = GS_IDLE; return GEV_COUNT; }
        m.addr = m.cnt_base + ((uint32_t)m.cnt_idx << m.cnt_shift) + (uint32_t)m.cnt_prefix;
        return GEV_NONE;
    case GS_EXP:
        if (bit && ++m.len < 11) { m.addr = m.exp_base + (uint32_t)m.len; return GEV_NONE; }
        if (m.len == 0) { m.st = GS_IDLE; value = 0; return GEV_COEF; }
        m.st = GS_SIGN; m.addr = m.sign_addr;
        return GEV_NONE;
    case GS_SIGN:
        m.neg = !bit;
        m.val = 1 << (m.len - 1);
        m.ri = m.len - 2;
        if (m.ri < 0) break;
        if (m.ri >= m.min_thr) {
            m.st = GS_THR; m.so = 1;
            m.thr_ctx += (uint32_t)min(m.len - m.min_thr, 7) << 7;          // m_thr(ci, ctx, len - min_thr)
            m.addr = m.thr_ctx + 1;
        } else {
            m.st = GS_RES; m.addr = m.res_base + (uint32_t)m.ri;
        }
        return GEV_NONE;
    case GS_THR:
        m.val |= (int)bit << m.ri;
        m.so = min((m.so << 1) | bit, 127u);
        if (--m.ri < 0) break;
        if (m.ri >= m.min_thr) m.addr = m.thr_ctx + m.so;
        else { m.st = GS_RES; m.addr = m.res_base + (uint32_t)m.ri; }
        return GEV_NONE;
    case GS_RES:
        m.val |= (int)bit << m.ri;
        if (--m.ri < 0) break;
        m.addr = m.res_base + (uint32_t)m.ri;
        return GEV_NONE;
    default:
        return GEV_NONE;
    }
    m.st = GS_IDLE;
    value = m.neg ? -m.val : m.val;
    return GEV_COEF;
}

constexpr int G_NO_THR = 127;             // min_thr of coefficients without threshold bits (7x7, DC)

// shared memory of one group: raster-order blocks and the per-block scratch of the lane-parallel phases
struct GGroupSmem {
    int16_t blk[4][64];                   // [0],[1] current / left (ping-pong); [2],[3] above / above-left (ping-pong)
    int32_t tmp[64];                      // IDCT intermediate (after the row pass)
    int16_t pix[64];                      // pixels of the block without its DC
    int32_t lak[14];                      // edge predictions: 7 horizontal, 7 vertical
    int16_t ledge[8];                     // right-column edge prediction of the left neighbour (block_context.hh:44-78)
    uint8_t pbsr[56];                     // 7x7: bit length of each coefficient's neighbour prior
};
static_assert(sizeof(GGroupSmem) == 4 * 128 + 256 + 128 + 56 + 16 + 56, "group scratch layout");

template <int G> struct GCfg {
    static constexpr int S = 32 / G;                                               // groups (thread-segments) per warp
    static constexpr int WARPS = (G >= 4) ? 4 : G;                                 // static shared memory stays under 48 KB
    static constexpr int THREADS = WARPS * 32;
};

template <int G> __device__ __forceinline__ int grp_sum(int v) {
#pragma unroll
    for (int d = 1; d < G; d <<= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}
template <int G> __device__ __forceinline__ int grp_min(int v) {
#pragma unroll
    for (int d = 1; d < G; d <<= 1) v = min(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
template <int G> __device__ __forceinline__ int grp_max(int v) {
#pragma unroll
    for (int d = 1; d < G; d <<= 1) v = max(v, __shfl_xor_sync(FULL, v, d));
    return v;
}

template <int G>
__global__ void __launch_bounds__(GCfg<G>::THREADS)
lep_decode_group_kernel(const ImageDesc* __restrict__ images, SegDesc* __restrict__ segs, int first, int count, const int* __restrict__ order,
                        int* __restrict__ work_counter, uint16_t* __restrict__ model_pool, uint8_t* __restrict__ row_pool, size_t row_pool_stride) {
    constexpr int S = GCfg<G>::S;
    constexpr int CPL = 64 / G;           // coefficients of a block per lane (aligned order)
    __shared__ uint32_t s_rcp[512];
    __shared__ uint8_t s_a2r[64];         // aligned index -> raster index (lanes look up different entries; constant memory would serialise)
    __shared__ uint8_t s_nzbin[64];
    __shared__ GGroupSmem s_grp[GCfg<G>::WARPS * S];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_rcp[i] = i < 2 ? 0u : (uint32_t)((0x100000000ull + i - 1) / i);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { s_a2r[i] = c_aligned_to_raster[i]; s_nzbin[i] = i < 50 ? c_nonzero_to_bin[i] : 0; }
    __syncthreads();
    const int lane = lane_id();
    const int sub = lane & (G - 1);                       // this lane's place in its group
    const int gbase = lane & ~(G - 1);                    // first lane of the group
    const int slot = (blockIdx.x * GCfg<G>::WARPS + (threadIdx.x >> 5)) * S + lane / G;       // row buffer of this group
    GGroupSmem& gs = s_grp[(threadIdx.x >> 5) * S + lane / G];
    uint8_t* rowbuf = row_pool + (size_t)slot * row_pool_stride;

    // ---- segment state (identical in the lanes of a group)
    bool alive = false;                   // a segment is in progress
    bool exhausted = false;               // the queue is empty: this group is done
    SegDesc* sdp = nullptr;
    const ImageDesc* gp = images;
    uint16_t* model = model_pool;
    int seg_min_y = 0, seg_max_y = 0;
    bool seg_last = false;
    GBool br;
    br.value = 0; br.valid = 0; br.range = 255; br.p = nullptr; br.end = nullptr;
    unsigned long long ndec = 0;
    int status = ST_OK;
    uint32_t top_mask = 7u, index = 0;
    int bw0 = 0, bw1 = 0, bw2 = 0, nzs0 = 0, nzs1 = 0;
    size_t nz_base = 0;
    // ---- row / block cursor (row iteration of lepton_codec.hh:41-100)
    int c = 0, ci = 0, y = 0, w = 0, x = 0, q0 = 1, pc = 0, pa = 2, nz_left = 0;
    bool has_above = false, need_row = true;
    int16_t* rowp = nullptr;
    const int16_t* abovep = nullptr;
    const uint16_t* q = nullptr;
    int16_t* redge = nullptr;
    uint8_t* rnz = nullptr;
    const int32_t* icx = nullptr;
    const int32_t* icy = nullptr;
    const uint8_t* mthr = nullptr;

    for (;;) {
        // ---- (0a) a finished segment reports; a free group takes the next segment of the queue
        const bool want_job = !alive && !exhausted;
        if (__any_sync(FULL, want_job)) {
            int job = -1;
            if (want_job && sub == 0) job = atomicAdd(work_counter, 1);
            job = __shfl_sync(FULL, job, gbase);
            if (want_job) {
                if (job >= count) {
                    exhausted = true;
                } else {
                    const int sidx = order[first + job];
                    sdp = &segs[sidx];
                    if (sdp->status == ST_OK) {               // else rejected on the host (e.g. zero quantiser, model.hh:257-262)
                        gp = &images[sdp->image];
                        model = model_pool + (size_t)job * M_TOTAL;          // zero-filled before the launch
                        seg_min_y = sdp->min_y; seg_max_y = sdp->max_y; seg_last = sdp->is_last != 0;
                        br.value = 0; br.valid = 0; br.range = 255;
                        br.p = reinterpret_cast<const uint8_t*>(sdp->stream); br.end = br.p + sdp->cap;
                        g_refill(br);
                        {   // marker bit at p = 128 (boolreader.cc:26-35); no model involved
                            const uint32_t split = (br.range * 128u + 128u) >> 8;
                            const uint32_t bit = (uint32_t)(br.value >> 56) >= split;
                            const uint32_t range = bit ? br.range - split : split;
                            if (bit) br.value -= (unsigned long long)split << 56;
                            const int shift = __clz(range) - 24;
                            br.range = range << shift; br.value <<= shift; br.valid -= shift;
                        }
                        ndec = 0; status = ST_OK; top_mask = 7u; index = 0;
                        bw0 = gp->bch[0]; bw1 = gp->ncmp > 1 ? gp->bch[1] : 0; bw2 = gp->ncmp > 2 ? gp->bch[2] : 0;
                        nz_base = (size_t)(bw0 + bw1 + bw2) * 16;
                        nzs0 = (bw0 + 15) & ~15; nzs1 = (bw1 + 15) & ~15;
                        need_row = true;
                        alive = true;
                    }
                }
            }
        }
        // ---- (0b) move to the next row when the previous one is finished (per group, once per row)
        if (alive && need_row) {
            const ImageDesc& g = *gp;
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
                mthr = g.min_thr[c];
                x = 0; pc = 0; pa = 2; nz_left = 0;
                need_row = false;
                break;
            }
            if (!alive && sub == 0) {                         // the segment is complete
                sdp->status = status;
                sdp->len = (uint32_t)(br.p - reinterpret_cast<const uint8_t*>(sdp->stream));       // whole words
                sdp->ndecisions_lo = (uint32_t)ndec;
                sdp->ndecisions_hi = (uint32_t)(ndec >> 32);
            }
        }
        if (__all_sync(FULL, exhausted)) break;
        // a group that just finished its segment sits this block round out and takes the next segment at the top

        const bool has_left = x > 0;
        int16_t* rcur = gs.blk[pc];
        const int16_t* rleft = gs.blk[pc ^ 1];
        int16_t* rabove = gs.blk[pa];
        const int16_t* raleft = gs.blk[pa ^ 1];
        GMicro m;
        m.st = GS_IDLE; m.addr = 0;
        int nz = 0, eobx = 0, eoby = 0;

        // ---- (1) above block -> raster copy, clear the current block, priors of the 7x7 coefficients, count context
        if (alive) {
            if (has_above) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(abovep + (size_t)x * 64) + sub * (CPL / 2);
#pragma unroll
                for (int k = 0; k < CPL / 2; ++k) {
                    const uint32_t u = src[k];
                    const int a = sub * CPL + 2 * k;
                    rabove[s_a2r[a]] = (int16_t)(u & 0xffff);
                    rabove[s_a2r[a + 1]] = (int16_t)(u >> 16);
                }
            }
#pragma unroll
            for (int k = 0; k < CPL / 2; ++k) reinterpret_cast<uint32_t*>(rcur)[sub * (CPL / 2) + k] = 0u;
        }
        __syncwarp();
        if (alive) {
            // compute_aavrg (model.hh:895-924) for the 49 positions, G at a time
            for (int zz = sub; zz < 49; zz += G) {
                const int coord = s_a2r[zz];
                uint32_t pr = 0;
                if (has_left && has_above) {
                    const uint32_t L = (uint32_t)iabs(rleft[coord]) & 0xffff, A = (uint32_t)iabs(rabove[coord]) & 0xffff;
                    pr = (((L + A) * 13u + (((uint32_t)iabs(raleft[coord]) & 0xffff) * 6u)) & 0xffff) >> 5;
                } else if (has_left || has_above) {
                    const int16_t nb = has_left ? rleft[coord] : rabove[coord];
                    pr = (uint32_t)iabs((int)(int16_t)((uint32_t)iabs(nb) & 0xffff));
                }
                gs.pbsr[zz] = (uint8_t)bitlen(min(pr, 1023u));
            }
            const int nz_above = has_above ? (int)rnz[x] : 0;
            int ctx = 0;
            if (has_above && !has_left) ctx = (nz_above + 1) / 2;
            else if (has_left && !has_above) ctx = (nz_left + 1) / 2;
            else if (has_left && has_above) ctx = (nz_above + nz_left + 2) / 4;
            g_start_count(m, m_nz7(ci, s_nzbin[ctx], 0, 0), 5, 6);
        }
        __syncwarp();

        // ---- (2) steps: 7x7 non-zero count, then the 7x7 coefficients in zig-zag order (== aligned order 0..48)
        {
            int zz = 0, left_nz = 0;
            bool busy = alive;
            while (__any_sync(FULL, busy)) {
                const uint32_t mw = busy ? model[m.addr] : 0u;
                G_EMU_BARRIER();
                if (busy) {
                    const uint32_t bit = g_get(br, model, s_rcp, m.addr, mw);
                    ++ndec;
                    int v = 0;
                    const int ev = g_advance(m, bit, v);
                    if (ev != GEV_NONE) {
                        bool next = true;
                        if (ev == GEV_COUNT) {
                            nz = m.cnt_val;
                            left_nz = nz;
                            if (nz > 49) { status = ST_STREAM_INCONSISTENT; alive = false; }
                            if (nz > 49 || nz == 0) next = false;
                        } else {
                            if (v != 0) {
                                const int coord = s_a2r[zz];
                                --left_nz;
                                eobx = max(eobx, coord & 7); eoby = max(eoby, coord >> 3);
                                if (sub == 0) rcur[coord] = (int16_t)v;
                            }
                            ++zz;
                            if (left_nz == 0 || zz == 49) next = false;
                        }
                        if (next) {
                            const int bin = s_nzbin[left_nz];
                            g_start_coef(m, m_exp7(ci, bin, zz, gs.pbsr[zz]), m_sign(ci, 0, 0), m_resn(ci, s_a2r[zz], bin), 0, G_NO_THR);
                        } else {
                            busy = false;
                        }
                    }
                }
            }
        }
        __syncwarp();
        // a stream that announces more than 49 coefficients ends its segment here (decoder.cc:182-184)
        if (status != ST_OK && sdp != nullptr && !alive && !need_row) {
            if (sub == 0) {
                sdp->status = status;
                sdp->len = (uint32_t)(br.p - reinterpret_cast<const uint8_t*>(sdp->stream));
                sdp->ndecisions_lo = (uint32_t)ndec;
                sdp->ndecisions_hi = (uint32_t)(ndec >> 32);
            }
            status = ST_OK; need_row = true;
        }

        // ---- (3) Lakhani predictions of the 14 edge coefficients: they read the 7x7 part of this block and the
        //          neighbours only, never the other edge (model.hh:1033-1071)
        if (alive) {
            for (int k = sub; k < 14; k += G) {
                int p = 0;
                if (k < 7) { if (has_above) p = lak_pred(rcur, rabove, icx + (k + 1) * 8, k + 1, 8); }
                else if (has_left) p = lak_pred(rcur, rleft, icy + (k - 6) * 8, 8 * (k - 6), 1);
                gs.lak[k] = p;
            }
            g_start_count(m, m_nze(0, ci, eobx, (nz + 3) / 7, 0, 0), 2, 3);
        }
        __syncwarp();

        // ---- (4) steps: horizontal edge (raster 1..7), then vertical edge (raster 8..56): count, then coefficients
        {
            int vert = 0, ne = 0, ln = 0;
            bool busy = alive;
            while (__any_sync(FULL, busy)) {
                const uint32_t mw = busy ? model[m.addr] : 0u;
                G_EMU_BARRIER();
                if (busy) {
                    const uint32_t bit = g_get(br, model, s_rcp, m.addr, mw);
                    ++ndec;
                    int v = 0;
                    const int ev = g_advance(m, bit, v);
                    if (ev != GEV_NONE) {
                        bool more;                                  // another coefficient of this edge follows
                        if (ev == GEV_COUNT) {
                            ne = m.cnt_val; ln = 0;
                            more = ne > 0;
                        } else {
                            if (v != 0) {
                                if (sub == 0) rcur[vert ? 8 * (ln + 1) : ln + 1] = (int16_t)v;
                                --ne;
                            }
                            ++ln;
                            more = ne > 0 && ln < 7;
                        }
                        if (more) {
                            const int coord = vert ? 8 * (ln + 1) : ln + 1;
                            const int prior = gs.lak[vert * 7 + ln];
                            const int bsr = bitlen((uint32_t)min(iabs(prior), 1023));
                            const int p16 = (int)(int16_t)prior;
                            const int sctx = p16 == 0 ? 0 : (p16 > 0 ? 1 : 2);
                            const int min_thr = mthr[coord];
                            const int ctx_abs = iabs(prior) & 0xffff;
                            g_start_coef(m, m_expx(ci, ne, vert ? 7 + ln : ln, bsr), m_sign(ci, sctx, bsr), m_resn(ci, coord, ne),
                                         m_thr(ci, min(ctx_abs >> min_thr, 255), 0), min_thr);
                        } else if (vert == 0) {
                            vert = 1;
                            g_start_count(m, m_nze(1, ci, eoby, (nz + 3) / 7, 0, 0), 2, 3);
                        } else {
                            busy = false;
                        }
                    }
                }
            }
        }
        __syncwarp();

        // ---- (5) DC: pixels of the block without its DC (8x8 IDCT, rows then columns, G lanes), prediction from the
        //          neighbours' edge pixels
        int32_t* tmp = gs.tmp;
        if (alive) {
            for (int r = sub; r < 8; r += G) {
                int32_t in[8], out[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) in[k] = (int32_t)rcur[r * 8 + k] * (int32_t)q[r * 8 + k];
                if (r == 0) in[0] = 0;
                idct_row(in, out);
#pragma unroll
                for (int k = 0; k < 8; ++k) tmp[r * 8 + k] = out[k];
            }
        }
        __syncwarp();
        if (alive) {
            for (int col = sub; col < 8; col += G) {
                int32_t in[8], out[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) in[k] = tmp[k * 8 + col];
                idct_col(in, out);
#pragma unroll
                for (int k = 0; k < 8; ++k) gs.pix[k * 8 + col] = (int16_t)out[k];
            }
        }
        __syncwarp();
        int pred = 0;
        {
            // adv_predict_dc_pix (model.hh:678-784), 16-bit lane arithmetic of the SSE build; 16 estimates over the G lanes
            int sl = 0, sa = 0, mnl = 32767, mxl = -32768, mna = 32767, mxa = -32768;
            if (alive) {
                for (int i = sub; i < 16; i += G) {
                    if (i < 8) {
                        if (has_left) {
                            const int16_t p0 = gs.pix[i * 8], p1 = gs.pix[i * 8 + 1];
                            const int16_t delta = (int16_t)(p0 - p1);
                            const int est = (int16_t)((int16_t)((int16_t)gs.ledge[i] - half_rz16(delta)) - (int16_t)(p0 + 1024));
                            sl += est; mnl = min(mnl, est); mxl = max(mxl, est);
                        }
                    } else if (has_above) {
                        const int j = i - 8;
                        const int16_t p0 = gs.pix[j], p1 = gs.pix[8 + j];
                        const int16_t delta = (int16_t)(p0 - p1);
                        const int est = (int16_t)((int16_t)((int16_t)redge[(size_t)x * 8 + j] - half_rz16(delta)) - (int16_t)(p0 + 1024));
                        sa += est; mna = min(mna, est); mxa = max(mxa, est);
                    }
                }
            }
            sl = grp_sum<G>(sl); sa = grp_sum<G>(sa);
            mnl = grp_min<G>(mnl); mna = grp_min<G>(mna);
            mxl = grp_max<G>(mxl); mxa = grp_max<G>(mxa);
            if (alive) {
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
                pred = (div_trunc_small(avgmed, q0) + 4) >> 3;
                const int lm = min(bitlen((uint32_t)iabs(unc) & 0xffff), 11), lo16 = min(bitlen((uint32_t)iabs(unc2) & 0xffff), 16);
                const int sctx = unc2 >= 0 ? (unc2 == 0 ? 3 : 2) : 1;
                g_start_coef(m, m_expdc(lm, lo16), m_sign(ci, 0, sctx), m_resdc(lm), 0, G_NO_THR);
            }
        }

        // ---- (6) steps: the DC coefficient
        int dcv = 0;
        {
            bool busy = alive;
            while (__any_sync(FULL, busy)) {
                const uint32_t mw = busy ? model[m.addr] : 0u;
                G_EMU_BARRIER();
                if (busy) {
                    const uint32_t bit = g_get(br, model, s_rcp, m.addr, mw);
                    ++ndec;
                    int v = 0;
                    if (g_advance(m, bit, v) == GEV_COEF) { dcv = v; busy = false; }
                }
            }
        }

        // ---- (7) neighbour summaries (block_context.hh:44-78), block store in aligned order, next block
        if (alive) {
            const int dc = (int)(int16_t)adv_unpredict((int)(int16_t)dcv, true, pred);            // decoder.cc:305-309
            if (sub == 0) rcur[0] = (int16_t)dc;
            const int16_t qdc = (int16_t)((uint32_t)q0 * (uint32_t)dc);
            for (int i = sub; i < 16; i += G) {
                if (i < 8) {   // right column -> the next block's left neighbour
                    const int16_t cur = gs.pix[i * 8 + 7], prev = gs.pix[i * 8 + 6];
                    const int16_t delta = (int16_t)(cur - prev);
                    gs.ledge[i] = (int16_t)(cur + half_rz16(delta) + 1024 + qdc);
                } else {       // bottom row -> the block below
                    const int j = i - 8;
                    const int16_t cur = gs.pix[56 + j], prev = gs.pix[48 + j];
                    const int16_t delta = (int16_t)(cur - prev);
                    redge[(size_t)x * 8 + j] = (int16_t)(cur + half_rz16(delta) + 1024 + qdc);
                }
            }
            if (sub == 0) rnz[x] = (uint8_t)nz;
            nz_left = nz;
        }
        __syncwarp();
        if (alive) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(rowp + (size_t)x * 64) + sub * (CPL / 2);
#pragma unroll
            for (int k = 0; k < CPL / 2; ++k) {
                const int a = sub * CPL + 2 * k;
                const uint32_t lo = (uint16_t)rcur[s_a2r[a]];
                const uint32_t hi = (uint16_t)rcur[s_a2r[a + 1]];
                dst[k] = lo | (hi << 16);
            }
            // a truncated image ends inside a row (component_size_in_blocks)
            if (x + 1 >= w || (uint32_t)((size_t)y * w + x + 1) >= (uint32_t)gp->trunc_bc[c]) need_row = true;
            else { ++x; pc ^= 1; pa ^= 1; }
        }
        __syncwarp();
    }
}

}  // namespace lepb200
