// lep_huffpar.cu -- baseline JPEG Huffman decode with MANY THREADS PER IMAGE.
//
// lep_huff.cu walks a scan with one warp: a chain of ~0.5 M symbols per 1080p image, 95 ms whatever the batch size, on
// the critical path of every compress call.  A baseline scan has no entry points, but Huffman streams re-synchronise: a
// decoder started at an arbitrary bit falls into step with the true symbol sequence after a while.  So the scan is cut
// into sub-sequences of `sub_bits` bits and every sub-sequence gets its own thread (Klein & Wiseman's observation as used
// by Weissenberger & Schmidt for JPEG):
//   (1) lep_huffpar_sync_kernel, iteration 0: thread i decodes from bit i * sub_bits assuming a block start and records
//       the state in which it leaves its sub-sequence (bit position, block of the MCU, zig-zag position);
//       iterations 1..: thread i starts again from the state thread i - 1 left in -- only if that state changed -- until
//       nothing changes any more.  Thread 0 starts from the true state, so after k iterations at least the first k + 1
//       exit states are the true ones; in practice a handful of iterations settle everything.  Every run also counts the
//       blocks it completed, the DC differences per component and the decision bound of its symbols;
//   (2) lep_huffpar_prefix_kernel: exclusive prefix of those counts over the sub-sequences of an image -> the block
//       index, the DC predictors and the decision count at which each sub-sequence starts;
//   (3) lep_huffpar_write_kernel: every thread decodes its sub-sequence once more from its now known state and stores
//       the coefficients, the per-MCU-row states (HuffRow) and, for the thread that completes the last block, the final
//       state of the scan -- exactly what lep_huffdecode_kernel produces.
// Anything other than a clean scan (invalid code, run past the block, trailing or missing data, restart intervals,
// single-component scans, no convergence) is left to lep_huffdecode_kernel, which runs afterwards for the jobs flagged
// here, so status codes and outputs are those of the serial walk in every case (jpgcoder.cc:2799-3302, :4893-4961).
#include "lep_huff.cu"

namespace lepb200 {

// decoder state between two symbols: [0,32) bit position | [32,38) zig-zag position (0 = the DC comes next) |
// [38,44) zig-zag position of the last non-zero of the open block seen before this point | 44 last AC symbol was non-zero |
// [48,56) block of the MCU
__device__ __forceinline__ unsigned long long hp_pack(uint32_t p, int z, int cl, int lnz, int b) {
    return (unsigned long long)p | ((unsigned long long)z << 32) | ((unsigned long long)cl << 38) | ((unsigned long long)lnz << 44) |
           ((unsigned long long)b << 48);
}

struct HpResult {
    unsigned long long exit;
    uint32_t nblk;           // blocks completed
    int dc[3];               // sum of the DC differences decoded, per component
    uint32_t tok;            // decision bound of the symbols decoded (lep_huff.cu: blk_sum + blk_last + 34 per block)
    int anomaly;
};

__device__ __forceinline__ void hp_lookup(const HuffTableDev* __restrict__ tab, uint32_t hi, int& len, int& sym) {
    len = 0; sym = 0;
    const uint32_t f = tab->fast[hi >> 23];
    if (f) { len = (int)(f >> 8); sym = (int)(f & 0xff); return; }
    const uint32_t top = hi >> 16;
    int l = 10, code = (int)(top >> 6);
    while (l <= 16 && code > tab->maxcode[l]) { ++l; code = (int)(top >> (16 - l)); }
    if (l <= 16) { len = l; sym = tab->vals[code + tab->valoff[l]]; }
}

// Decodes from `start` until the bit position reaches `end_bit` (or, when WRITE, the image is complete).
// WRITE: n0 / pdc / tok0 are the block index, DC predictors and decision count at `start`.
template <bool WRITE>
__device__ __forceinline__ void hp_span(HuffJob& jb, const HuffTableDev* __restrict__ tb, const uint8_t* __restrict__ zz,
                                        unsigned long long start, uint32_t end_bit, bool is_last,
                                        uint32_t n0, int pdc0, int pdc1, int pdc2, uint32_t tok0, HpResult& o) {
    const uint32_t* __restrict__ words = reinterpret_cast<const uint32_t*>(jb.huff);
    const uint32_t nwords = (jb.nbytes + 3) / 4, total_bits = jb.nbytes * 8u;
    const int ncmp = jb.ncmp, mcuh = jb.mcuh, mcuv = jb.mcuv;
    const int nb0 = jb.H[0] * jb.V[0], nb1 = ncmp > 1 ? nb0 + jb.H[1] * jb.V[1] : nb0, bpm = ncmp > 2 ? nb1 + jb.H[2] * jb.V[2] : nb1;
    uint32_t p = (uint32_t)start;
    int z = (int)(start >> 32) & 63, cl = (int)(start >> 38) & 63, lnz = (int)(start >> 44) & 1, b = (int)(start >> 48) & 255;
    int cmp = b < nb0 ? 0 : (b < nb1 ? 1 : 2);
    const HuffTableDev* dct;
    const HuffTableDev* act;
    uint32_t nblk = 0, tok = 0;
    int s0 = 0, s1 = 0, s2 = 0, blk_sum = 0, blk_last = 0;
    o.anomaly = 0;
    // geometry of the write pass
    const uint32_t N = (uint32_t)mcuh * (uint32_t)mcuv * (uint32_t)bpm;
    int mx = 0, my = 0;
    int16_t* blk = nullptr;
    HuffRow* rows = reinterpret_cast<HuffRow*>(jb.rows);
    auto block_ptr = [&]() -> int16_t* {
        const int b0 = cmp == 0 ? b : (cmp == 1 ? b - nb0 : b - nb1);
        const int H = jb.H[cmp], V = jb.V[cmp];
        const int sy = b0 / H, sx = b0 - sy * H;
        return reinterpret_cast<int16_t*>(jb.plane[cmp]) + ((size_t)(my * V + sy) * jb.bch[cmp] + mx * H + sx) * 64;
    };
    bool finished = false;
    if (WRITE) {
        const uint32_t mcu = n0 / (uint32_t)bpm;
        if ((int)(n0 - mcu * (uint32_t)bpm) != b) { o.anomaly = 1; return; }
        mx = (int)(mcu % (uint32_t)mcuh); my = (int)(mcu / (uint32_t)mcuh);
        blk = block_ptr();
    }
    // table of the current block: indices of the three components kept in registers (no global load per block)
    const int dct0 = jb.dc_tab[0], dct1 = jb.dc_tab[1], dct2 = jb.dc_tab[2], act0 = jb.ac_tab[0], act1 = jb.ac_tab[1], act2 = jb.ac_tab[2];
    dct = tb + (cmp == 0 ? dct0 : (cmp == 1 ? dct1 : dct2));
    act = tb + (cmp == 0 ? act0 : (cmp == 1 ? act1 : act2));
    uint32_t ck = 0xffffffffu, w0 = 0, w1 = 0;
    while (p < end_bit) {
        const uint32_t k = p >> 5, sh = p & 31;
        if (k != ck) {
            if (k == ck + 1 && ck != 0xffffffffu) w0 = w1; else w0 = be_word(words, k, nwords);
            w1 = be_word(words, k + 1, nwords);
            ck = k;
        }
        const uint32_t hi = __funnelshift_l(w1, w0, sh);       // 32 bits from p: a code (<= 16 bits) and its magnitude bits (<= 16)
        // DC and AC symbols take the same instructions (the lanes of a warp are at different places of their blocks)
        const bool isdc = z == 0;
        if (WRITE && isdc && b == 0 && mx == 0) {              // first block of an MCU row: the resumable state (HuffRow)
            HuffRow r;
            r.bitpos = p; r.mcu_y = (int16_t)my;
            r.lastdc[0] = (int16_t)(pdc0 + s0); r.lastdc[1] = (int16_t)(pdc1 + s1); r.lastdc[2] = (int16_t)(pdc2 + s2);
            r.tokens = tok0 + tok;
            rows[my] = r;
        }
        int len, sym;
        hp_lookup(isdc ? dct : act, hi, len, sym);
        const int sz = isdc ? sym : (sym & 15), run = isdc ? 0 : (sym >> 4);
        if (len == 0 || sz > 16) {                             // no such code
            if (WRITE) { o.anomaly = 1; return; }
            p += 1; continue;
        }
        int val = 0;
        if (sz) {
            const int nb = (int)((hi << len) >> (32 - sz));
            val = nb >= (1 << (sz - 1)) ? nb : nb + 1 - (1 << sz);
        }
        p += (uint32_t)(len + sz);
        bool block_done = false;
        if (isdc) {
            if (cmp == 0) s0 += val; else if (cmp == 1) s1 += val; else s2 += val;
            if (WRITE) blk[49] = (int16_t)((cmp == 0 ? pdc0 + s0 : (cmp == 1 ? pdc1 + s1 : pdc2 + s2)));
            z = 1; lnz = 1;
        } else if (sym == 0) {                                 // EOB
            if (WRITE && z > 1 && !lnz) { o.anomaly = 1; return; }           // "eob after last 0" (jpgcoder.cc:2953)
            block_done = true;
        } else if (run + z >= 64) {                            // the truncated-file fix-up path of the reference: serial kernel
            if (WRITE) { o.anomaly = 1; return; }
            block_done = true;
        } else {
            z += run;
            if (sz) { blk_sum += min(sz + 1, 11) + sz - 1; blk_last = z; }
            if (WRITE) blk[zz[z]] = (int16_t)val;
            lnz = sz != 0;
            ++z;
            if (z >= 64) block_done = true;
        }
        if (!block_done) continue;
        tok += (uint32_t)(blk_sum + (blk_last ? blk_last - cl : 0) + 34);
        blk_sum = 0; blk_last = 0; cl = 0;
        ++nblk;
        if (WRITE && p > total_bits) { o.anomaly = 1; return; }               // entropy data ends inside a block
        if (++b == bpm) {
            b = 0;
            if (WRITE && ++mx == mcuh) { mx = 0; ++my; }
        }
        cmp = b < nb0 ? 0 : (b < nb1 ? 1 : 2);
        dct = tb + (cmp == 0 ? dct0 : (cmp == 1 ? dct1 : dct2));
        act = tb + (cmp == 0 ? act0 : (cmp == 1 ? act1 : act2));
        z = 0; lnz = 1;
        if (WRITE) {
            if (n0 + nblk == N) { finished = true; break; }
            blk = block_ptr();
        }
    }
    if (WRITE) {
        if (finished) {
            // end of the scan: abitreader::unpad (bitops.hh:316-332) + padbit bookkeeping (jpgcoder.cc:3260-3271), then the
            // final row record, as at the end of lep_huffdecode_kernel
            int fb = -1;
            if ((p & 7) != 0 && p < total_bits) {
                auto bit_at = [&](uint32_t pos) -> int { return (int)((be_word(words, pos >> 5, nwords) >> (31 - (pos & 31))) & 1u); };
                int last = bit_at(p); ++p;
                fb = last;
                int offset = 1;
                while (p & 7) { last = bit_at(p); ++p; fb |= last << offset; ++offset; }
                while (offset < 7) { fb |= last << offset; ++offset; }
            }
            if (p < total_bits) { o.anomaly = 1; return; }                    // "unneeded data found after coded image data"
            HuffRow r;
            r.bitpos = p; r.mcu_y = (int16_t)my;
            r.lastdc[0] = (int16_t)(pdc0 + s0); r.lastdc[1] = (int16_t)(pdc1 + s1); r.lastdc[2] = (int16_t)(pdc2 + s2);
            r.tokens = tok0 + tok;
            rows[mcuv] = r;
            jb.padbit = fb;
            jb.end_bitpos = p;
            jb.nrows = mcuv + 1;
            jb.par_done = 1;
        } else if (is_last) {
            o.anomaly = 1;                                                    // the data ran out before the last block
        }
        return;
    }
    if (z > 0) {                                               // open block: its share of the bound so far
        tok += (uint32_t)(blk_sum + (blk_last ? blk_last - cl : 0));
        if (blk_last) cl = blk_last;
    }
    o.exit = hp_pack(p, z, cl, lnz, b);
    o.nblk = nblk; o.dc[0] = s0; o.dc[1] = s1; o.dc[2] = s2; o.tok = tok;
}

constexpr int HP_THREADS = 128;
constexpr int HP_SMEM_TABLES = 8;

struct HpArrays {
    unsigned long long* exit;    // per sub-sequence: state in which its thread left it
    uint32_t* epoch;             // iteration in which the sub-sequence has to be decoded again
    uint4* cnt;                  // (blocks, dc0, dc1, dc2) of the last run; after the prefix kernel: the values at its start
    uint32_t* tok;
    const uint32_t* sub_base;    // njobs + 1: first sub-sequence of each job
    uint32_t total;
    uint32_t sub_bits;
};

__device__ __forceinline__ int hp_find_job(const uint32_t* __restrict__ sub_base, int njobs, uint32_t g) {
    int lo = 0, hi = njobs;                                    // largest j with sub_base[j] <= g
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(sub_base + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// tables (and the zig-zag map) into shared memory when the batch uses few distinct ones
__device__ __forceinline__ const HuffTableDev* hp_stage_tables(HuffTableDev* s_tab, uint8_t* s_zz, const HuffTableDev* __restrict__ tables, int ntables) {
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_zz[i] = c_zigzag_to_aligned[i];
    const bool use_smem = ntables <= HP_SMEM_TABLES;
    if (use_smem) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_tab);
        const int nw = ntables * (int)(sizeof(HuffTableDev) / 4);
        for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    return use_smem ? s_tab : tables;
}

// (1) one synchronisation iteration.  dirty[iter + 1] counts the sub-sequences that have to run in the next one.
__global__ void __launch_bounds__(HP_THREADS)
lep_huffpar_sync_kernel(HuffJob* __restrict__ jobs, int njobs, const HuffTableDev* __restrict__ tables, int ntables, HpArrays a, int iter,
                        int last_iter, unsigned int* __restrict__ dirty) {
    __shared__ HuffTableDev s_tab[HP_SMEM_TABLES];
    __shared__ uint8_t s_zz[64];
    const HuffTableDev* tb = hp_stage_tables(s_tab, s_zz, tables, ntables);
    const uint32_t g = blockIdx.x * HP_THREADS + threadIdx.x;
    if (g >= a.total) return;
    if (iter > 0 && a.epoch[g] != (uint32_t)iter) return;
    const int j = hp_find_job(a.sub_base, njobs, g);
    HuffJob& jb = jobs[j];
    const uint32_t i = g - a.sub_base[j];
    if (iter > 0 && i == 0) return;
    const uint32_t total_bits = jb.nbytes * 8u;
    const unsigned long long start = (iter == 0 || i == 0) ? hp_pack(i * a.sub_bits, 0, 0, 1, 0) : a.exit[g - 1];
    const uint32_t end_bit = min(total_bits, (i + 1) * a.sub_bits);
    HpResult r;
    hp_span<false>(jb, tb, s_zz, start, end_bit, false, 0, 0, 0, 0, 0, r);
    a.cnt[g] = make_uint4(r.nblk, (uint32_t)r.dc[0], (uint32_t)r.dc[1], (uint32_t)r.dc[2]);
    a.tok[g] = r.tok;
    if (iter == 0 || r.exit != a.exit[g]) {
        a.exit[g] = r.exit;
        if (i + 1 < jb.nsub) {
            a.epoch[g + 1] = (uint32_t)iter + 1;
            atomicAdd(dirty + iter + 1, 1u);
            if (iter == last_iter) jb.par_redo = 1;            // no convergence within the iteration budget
        }
    }
}

// (2) exclusive prefix over the sub-sequences of each image; one thread per image
__global__ void lep_huffpar_prefix_kernel(const HuffJob* __restrict__ jobs, int njobs, HpArrays a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    const uint32_t g0 = a.sub_base[j], n = jobs[j].nsub;
    uint32_t nb = 0, d0 = 0, d1 = 0, d2 = 0, tk = 0;
    for (uint32_t g = g0; g < g0 + n; ++g) {
        const uint4 c = a.cnt[g];
        const uint32_t t = a.tok[g];
        a.cnt[g] = make_uint4(nb, d0, d1, d2);
        a.tok[g] = tk;
        nb += c.x; d0 += c.y; d1 += c.z; d2 += c.w; tk += t;
    }
}

// (3) the write pass
__global__ void __launch_bounds__(HP_THREADS)
lep_huffpar_write_kernel(HuffJob* __restrict__ jobs, int njobs, const HuffTableDev* __restrict__ tables, int ntables, HpArrays a) {
    __shared__ HuffTableDev s_tab[HP_SMEM_TABLES];
    __shared__ uint8_t s_zz[64];
    const HuffTableDev* tb = hp_stage_tables(s_tab, s_zz, tables, ntables);
    const uint32_t g = blockIdx.x * HP_THREADS + threadIdx.x;
    if (g >= a.total) return;
    const int j = hp_find_job(a.sub_base, njobs, g);
    HuffJob& jb = jobs[j];
    if (jb.par_redo) return;
    const uint32_t i = g - a.sub_base[j];
    const uint32_t total_bits = jb.nbytes * 8u;
    const unsigned long long start = i == 0 ? hp_pack(0, 0, 0, 1, 0) : a.exit[g - 1];
    const uint32_t end_bit = min(total_bits, (i + 1) * a.sub_bits);
    const uint4 c = a.cnt[g];
    const int nb0 = jb.H[0] * jb.V[0], nb1 = jb.ncmp > 1 ? nb0 + jb.H[1] * jb.V[1] : nb0, bpm = jb.ncmp > 2 ? nb1 + jb.H[2] * jb.V[2] : nb1;
    const uint32_t N = (uint32_t)jb.mcuh * (uint32_t)jb.mcuv * (uint32_t)bpm;
    const bool is_last = i + 1 == jb.nsub;
    if (c.x >= N) {                                            // the image ended in an earlier sub-sequence
        return;
    }
    HpResult r;
    hp_span<true>(jb, tb, s_zz, start, end_bit, is_last, c.x, (int)c.y, (int)c.z, (int)c.w, a.tok[g], r);
    if (r.anomaly) jb.par_redo = 1;
}

}  // namespace lepb200
