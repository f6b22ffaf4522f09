// lep_huff.cu -- baseline JPEG Huffman decode on the GPU.
//
// Moves the 16x data expansion (JPEG bytes -> 128 B/block coefficient planes) onto the device: the host only
// splits the file into header / de-stuffed entropy bytes (lep_jpeg.cc parse_jpeg) and uploads the entropy bytes;
// this kernel produces the coefficient planes in AlignedBlock order directly in HBM, plus, per MCU row, the
// resumable Huffman state the reference calls a ThreadHandoff (bit position, last DCs;
// crystallize_thread_handoff, src/lepton/jpgcoder.cc:2520-2560).
//
// A baseline scan without restart markers is one serial bit stream, so the unit of parallelism is the image: ONE WARP
// PER IMAGE, whose 32 lanes decode the codewords that would start at the next 32 bit offsets while a warp-uniform walk
// follows the true chain (see the kernel).  A batch of thousands of files keeps the chip busy: 4096 images are one wave
// of 28 warps per SM.  Besides the Huffman state the kernel accumulates, per MCU row, an upper bound of the binary
// decisions the Lepton coder will take, so the encoder's token streams can be laid out without a counting pass.
// Semantics follow decode_jpeg / decode_block_seq (jpgcoder.cc:2799-3302, :4893-4961) exactly like the host decoder in
// lep_jpeg.cc, against which it is tested.
#pragma once
#include <cstring>

#include "lep_common.cuh"

namespace lepb200 {

struct HuffTableDev {
    uint16_t fast[512];      // (len << 8) | symbol for codes of <= 9 bits, 0 = longer code
    int32_t maxcode[18];
    int32_t valoff[18];
    uint8_t vals[256];
};

// host side: the decode tables of one DHT table (counts per code length, values in code order)
static inline bool huff_build_table(const uint8_t bits[17], const uint8_t vals[256], HuffTableDev& t) {
    memset(&t, 0, sizeof(t));
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        t.valoff[len] = k - code;
        for (int i = 0; i < bits[len]; ++i, ++k, ++code) {
            if (k >= 256 || code >= (1 << len)) return false;        // over-subscribed table
            t.vals[k] = vals[k];
            if (len <= 9) {
                const int shift = 9 - len;
                for (int f = 0; f < (1 << shift); ++f) t.fast[(code << shift) | f] = (uint16_t)((len << 8) | vals[k]);
            }
        }
        t.maxcode[len] = bits[len] ? code - 1 : -1;
        if (code > (1 << len)) return false;
        code <<= 1;
    }
    t.maxcode[17] = 0x7fffffff;
    return true;
}

struct HuffRow {             // state at the start of an MCU row
    uint32_t bitpos;         // bits consumed of the de-stuffed entropy stream
    int16_t lastdc[3];
    int16_t mcu_y;
    uint32_t tokens;         // upper bound of the binary decisions the Lepton coder takes for all blocks before this row
};

constexpr int HUFF_JOB_SKIP = -1;        // HuffJob::status of a placeholder (plane slot only, nothing to decode)

struct HuffJob {
    unsigned long long huff;         // device address of the de-stuffed entropy bytes (4-byte aligned, zero padded by 8)
    unsigned long long plane[3];     // coefficient planes (pre-zeroed)
    unsigned long long rows;         // HuffRow[mcuv + 1]
    uint32_t nbytes;
    int32_t ncmp, mcuh, mcuv, rsti;
    int32_t H[3], V[3], bch[3], bcv[3], nch[3], ncv[3];
    int32_t dc_tab[3], ac_tab[3];    // indices into the table array
    // outputs
    int32_t status;                  // 0 ok; 42 UNSUPPORTED_JPEG (decode error / eob after last 0 / unneeded data); 200 not handled
    int32_t padbit;
    uint32_t end_bitpos;
    int32_t nrows;
    // many-threads-per-image path (lep_huffpar.cu): first sub-sequence and their number (0 = this kernel decodes the image)
    uint32_t sub_base, nsub;
    int32_t par_done;                // the sub-sequence pass completed the image
    int32_t par_redo;                // it met something only the serial walk below classifies (an error, trailing data, no
                                     // convergence): decode the image again here, over freshly zeroed planes
};

static __constant__ uint8_t c_zigzag_to_aligned[64] = {
    49, 50, 57, 58, 0, 51, 52, 1, 2, 59, 60, 3, 4, 5, 53, 54, 6, 7, 8, 9, 61, 62, 10, 11,
    12, 13, 14, 55, 56, 15, 16, 17, 18, 19, 20, 63, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32,
    33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48};

struct HBits {
    const uint32_t* w;       // big-endian words of the stream
    unsigned long long acc;  // next bits, MSB first, in the top `n` bits
    int n;
    uint32_t wi, nwords;
    unsigned long long bitpos;
};
__device__ __forceinline__ void hb_fill(HBits& b) {
    if (b.n <= 32) {
        uint32_t v = b.wi < b.nwords ? __ldg(b.w + b.wi) : 0u;
        b.wi++;
        v = __byte_perm(v, 0, 0x0123);
        b.acc |= (unsigned long long)v << (32 - b.n);
        b.n += 32;
    }
}
__device__ __forceinline__ uint32_t hb_peek(const HBits& b, int k) { return (uint32_t)(b.acc >> (64 - k)); }   // 1 <= k <= 32
__device__ __forceinline__ void hb_skip(HBits& b, int k) { b.acc <<= k; b.n -= k; b.bitpos += (unsigned)k; }

__device__ __forceinline__ int huff_symbol(HBits& b, const HuffTableDev* __restrict__ t) {
    hb_fill(b);
    const uint32_t top = hb_peek(b, 16);
    const uint32_t f = __ldg(&t->fast[top >> 7]);
    if (f) { hb_skip(b, (int)(f >> 8)); return (int)(f & 0xff); }
    int len = 10;
    int code = (int)(top >> 6);
    while (len <= 16 && code > __ldg(&t->maxcode[len])) { ++len; code = (int)(top >> (16 - len)); }
    if (len > 16) return -1;
    hb_skip(b, len);
    return __ldg(&t->vals[code + __ldg(&t->valoff[len])]);
}
__device__ __forceinline__ int huff_extend(HBits& b, int s) {      // DEVLI (jpgcoder.cc:117)
    if (s == 0) return 0;
    hb_fill(b);
    const int n = (int)hb_peek(b, s);
    hb_skip(b, s);
    return n >= (1 << (s - 1)) ? n : n + 1 - (1 << s);
}

constexpr int HUFF_MAX_WARPS = 8;         // warps (= images) per CTA is a launch parameter (blockDim.x / 32)
constexpr int HUFF_SMEM_TABLES = 8;      // tables staged in shared memory when the batch uses few distinct ones

// ONE WARP PER IMAGE, window-parallel Huffman decode.  A single thread walking the bit stream pays ~500 cycles per
// symbol on this machine (in-order issue, every step a dependent load).  Instead the 32 lanes decode, in parallel, the
// codeword that WOULD start at each of the next 32 bit offsets (table lookup + magnitude bits from a private 64-bit
// window); the true symbol sequence is then recovered by a warp-uniform walk that costs one shuffle per symbol:
// start at offset 0, jump by each symbol's total length, stop at the end of the block or of the 32-offset window.
// Only the first symbol of a block uses the DC table, and windows are re-based at every block start, so lane 0 alone
// ever needs it.
__device__ __forceinline__ uint32_t be_word(const uint32_t* __restrict__ w, uint32_t k, uint32_t nwords) {
    return k < nwords ? __byte_perm(__ldg(w + k), 0, 0x0123) : 0u;
}

#ifndef LEPB200_HUFF_MINBLOCKS
#define LEPB200_HUFF_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(HUFF_MAX_WARPS * 32, LEPB200_HUFF_MINBLOCKS)
lep_huffdecode_kernel(HuffJob* __restrict__ jobs, int njobs, const HuffTableDev* __restrict__ tables, int ntables) {
    __shared__ HuffTableDev s_tab[HUFF_SMEM_TABLES];
    __shared__ uint8_t s_zz[64];
    // per-component constants of each warp's image: read once per block, so they live in shared memory, not registers
    struct CmpInfo { unsigned long long plane; const HuffTableDev* dct; const HuffTableDev* act; int H, V, W, pad; };
    __shared__ CmpInfo s_cmp[HUFF_MAX_WARPS][3];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_zz[i] = c_zigzag_to_aligned[i];
    const bool use_smem = ntables <= HUFF_SMEM_TABLES;
    if (use_smem) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_tab);
        const int nw = ntables * (int)(sizeof(HuffTableDev) / 4);
        for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const HuffTableDev* tb = use_smem ? s_tab : tables;
    const int job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (job >= njobs) return;
    HuffJob& jb = jobs[job];
    if (jb.status != 0) return;
    if (jb.nsub != 0) {                              // the image went through the sub-sequence kernels first
        if (jb.par_done && !jb.par_redo) return;
        for (int q = 0; q < jb.ncmp; ++q) {          // rare: start over from zeroed planes
            uint4* pl = reinterpret_cast<uint4*>(jb.plane[q]);
            const size_t n16 = (size_t)jb.bch[q] * jb.bcv[q] * 8;
            for (size_t i = lane; i < n16; i += 32) pl[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
    }

    const uint32_t* __restrict__ words = reinterpret_cast<const uint32_t*>(jb.huff);
    const uint32_t nwords = (jb.nbytes + 3) / 4;
    const uint32_t total_bits = jb.nbytes * 8u;
    HuffRow* rows = reinterpret_cast<HuffRow*>(jb.rows);
    const int ncmp = jb.ncmp, mcuh = jb.mcuh, mcuv = jb.mcuv, rsti = jb.rsti;
    CmpInfo* const ci = s_cmp[threadIdx.x >> 5];
    if (lane < 3) {
        const int q = lane < ncmp ? lane : 0;
        CmpInfo t;
        t.plane = jb.plane[q]; t.dct = tb + jb.dc_tab[q]; t.act = tb + jb.ac_tab[q];
        t.H = lane < ncmp ? jb.H[q] : 1; t.V = lane < ncmp ? jb.V[q] : 1; t.W = lane < ncmp ? jb.bch[q] : 0; t.pad = 0;
        ci[lane] = t;
    }
    __syncwarp();
    const int H0 = ci[0].H, V0 = ci[0].V, W0 = ci[0].W;
    const int nch0 = jb.nch[0], ncv0 = jb.ncv[0];
    int16_t* const P0 = reinterpret_cast<int16_t*>(ci[0].plane);

    // warp-uniform decoder state
    uint32_t p = 0;                                  // bit position of the next symbol
    int dc0 = 0, dc1 = 0, dc2 = 0;
    int cmp = 0, sx = 0, sy = 0, mx = 0, my = 0, bx = 0, by = 0;
    int nrows = 0, status = 0, padbit = -1, final_y = 0;
    int rstw = rsti;
    int bpos = 0;
    bool last_nonzero = true;
    // Bound of the coder's decision count, per block: 12 count bits + <= 22 for the DC + for every non-zero AC its
    // exponent/sign/residual decisions + one decision per coded zero, of which there are at most (zig-zag index of the
    // last non-zero) - (number of non-zeros).  Replaces a separate counting pass over the decoded planes.
    uint32_t tokacc = 0;
    int blk_sum = 0, blk_last = 0;
    int16_t* blk = P0;
    const HuffTableDev* dct = ci[0].dct;
    const HuffTableDev* act = ci[0].act;
    int curH = H0, curV = V0;
    auto push_row = [&](int mcu_y) {
        if (lane == 0) {
            HuffRow r;
            r.bitpos = p; r.mcu_y = (int16_t)mcu_y;
            r.lastdc[0] = (int16_t)dc0; r.lastdc[1] = (int16_t)dc1; r.lastdc[2] = (int16_t)dc2;
            r.tokens = tokacc;
            rows[nrows] = r;
        }
        nrows++;
    };
    auto bit_at = [&](uint32_t pos) -> int {         // uniform single-bit read (padding bits)
        const uint32_t wv = be_word(words, pos >> 5, nwords);
        return (int)((wv >> (31 - (pos & 31))) & 1u);
    };
    push_row(0);
    bool done = false;
    while (!done) {
        // ---- every lane: the symbol that would start at bit offset p + lane
        const uint32_t off = p + (uint32_t)lane;
        const uint32_t k = off >> 5, sh = off & 31;
        const uint32_t w0 = be_word(words, k, nwords), w1 = be_word(words, k + 1, nwords);
        const uint32_t hi = __funnelshift_l(w1, w0, sh);                     // 32 bits from `off`, MSB first: a code (<= 16 bits)
                                                                             // and its magnitude bits (<= 16) always fit
        const bool is_dc = lane == 0 && bpos == 0;
        const HuffTableDev* tab = is_dc ? dct : act;
        int len = 0, sym = 0;
        {
            const uint32_t f = tab->fast[hi >> 23];
            if (f) { len = (int)(f >> 8); sym = (int)(f & 0xff); }
            else {
                const uint32_t top = hi >> 16;
                int l = 10, code = (int)(top >> 6);
                while (l <= 16 && code > tab->maxcode[l]) { ++l; code = (int)(top >> (16 - l)); }
                if (l <= 16) { len = l; sym = tab->vals[code + tab->valoff[l]]; }
            }
        }
        const int sz = is_dc ? sym : (sym & 15);
        const int run = is_dc ? 0 : (sym >> 4);
        int val = 0;
        bool bad = len == 0 || sz > 16;
        if (!bad && sz) {
            // sz magnitude bits follow the code: bits [len, len+sz) of the window
            const int nb = (int)((hi << len) >> (32 - sz));
            val = nb >= (1 << (sz - 1)) ? nb : nb + 1 - (1 << sz);
        }
        // info: [0,6) total length (0 = invalid) | [6,10) run | [10,15) size | [16,32) value
        const uint32_t info = bad ? 0u : ((uint32_t)(len + sz) | ((uint32_t)run << 6) | ((uint32_t)sz << 10) | ((uint32_t)(val & 0xffff) << 16));
        // ---- warp-uniform walk along the true symbol sequence inside this window: it only follows the chain of code
        // lengths and tracks the zig-zag position; every visited lane then stores its own coefficient (parallel stores)
        int cur = 0;
        bool block_done = false;
        int mypos = -1;                               // zig-zag position of this lane's AC symbol, if it is on the chain
        if (bpos == 0) {                              // first symbol of a block: the DC difference, decoded by lane 0
            const uint32_t inf = __shfl_sync(FULL, info, 0);
            const int tot = (int)(inf & 63);
            if (tot == 0) { status = 42; break; }
            const int v = (int)(int16_t)(inf >> 16);
            const int last = cmp == 0 ? dc0 : (cmp == 1 ? dc1 : dc2);
            const int16_t dcv = (int16_t)(v + last);
            if (cmp == 0) dc0 = dcv; else if (cmp == 1) dc1 = dcv; else dc2 = dcv;
            if (lane == 0) blk[49] = dcv;
            bpos = 1;
            last_nonzero = true;
            cur = tot;
        }
        while (cur < 32) {
            const uint32_t inf = __shfl_sync(FULL, info, cur);
            const int tot = (int)(inf & 63);
            if (tot == 0) { status = 42; break; }
            const int rz = (int)((inf >> 6) & 0x1ff);                         // run | size << 4
            if (rz == 0) {                                                    // EOB
                if (bpos > 1 && !last_nonzero) status = 42;                   // "eob after last 0" (jpgcoder.cc:2953)
                cur += tot;
                block_done = true;
                break;
            }
            const int r = rz & 15;
            if (r + bpos >= 64) { status = 200; break; }                      // truncated-file fix-up path: not handled here
            bpos += r;
            if (lane == cur) mypos = bpos;
            last_nonzero = (rz >> 4) != 0;                                    // size 0 here is ZRL, the only zero-valued AC symbol
            ++bpos;
            cur += tot;
            if (bpos >= 64) { block_done = true; break; }
        }
        if (mypos >= 0) blk[s_zz[mypos]] = (int16_t)(info >> 16);
        {
            const int z = (int)((info >> 10) & 31u);
            const bool nzsym = mypos >= 0 && z > 0;
            blk_sum += __reduce_add_sync(FULL, nzsym ? min(z + 1, 11) + z - 1 : 0);
            blk_last = max(blk_last, (int)__reduce_max_sync(FULL, nzsym ? mypos : 0));
        }
        if (status) break;
        p += (uint32_t)cur;
        if (!block_done) continue;
        if (p > total_bits) { status = 200; break; }                          // entropy data ends inside a block
        tokacc += (uint32_t)(blk_sum + blk_last + 34);
        blk_sum = 0; blk_last = 0;
        // ---- next block position (next_mcupos / next_mcuposn), warp-uniform
        int sta = 0;
        bool handoff_due = false;
        if (ncmp > 1) {
            if (++sx >= curH) { sx = 0; ++sy; }
            if (sy >= curV) {
                sy = 0;
                if (++cmp >= ncmp) {
                    cmp = 0;
                    if (++mx >= mcuh) { mx = 0; ++my; handoff_due = true; }
                    if (my >= mcuv) sta = 2;
                    else if (rsti > 0 && --rstw == 0) sta = 1;
                }
                dct = ci[cmp].dct; act = ci[cmp].act; curH = ci[cmp].H; curV = ci[cmp].V;
            }
            blk = reinterpret_cast<int16_t*>(ci[cmp].plane) + ((size_t)(my * curV + sy) * ci[cmp].W + mx * curH + sx) * 64;
        } else {
            if (++bx >= nch0) { bx = 0; ++by; }
            if (by >= ncv0) sta = 2;
            else if (rsti > 0 && --rstw == 0) sta = 1;
            const int dpos = by * W0 + bx;
            blk = P0 + (size_t)dpos * 64;
            const int per_mcu = H0 * V0;
            if (sta != 2) {
                my = (dpos / per_mcu) / mcuh;
                if ((dpos % per_mcu) == 0 && ((dpos / per_mcu) % mcuh) == 0) handoff_due = true;
            } else {
                my = mcuv;
            }
        }
        bpos = 0;
        if (p >= total_bits) sta = 2;                                         // huffr->eof
        if (sta != 0) {
            // abitreader::unpad (bitops.hh:316-332) + padbit bookkeeping (jpgcoder.cc:3260-3271)
            int fb = padbit;
            if ((p & 7) != 0 && p < total_bits) {
                int last = bit_at(p); ++p;
                fb = last;
                int offset = 1;
                while (p & 7) { last = bit_at(p); ++p; fb |= last << offset; ++offset; }
                while (offset < 7) { fb |= last << offset; ++offset; }
            }
            if (padbit != -1) { if (padbit != fb) { status = 42; break; } }
            else padbit = fb;
            if (sta == 2) { final_y = my; done = true; break; }
            dc0 = dc1 = dc2 = 0;                                              // restart interval
            rstw = rsti;
        }
        if (handoff_due) push_row(my);
    }
    if (status == 0) {
        push_row(final_y);
        if (p < total_bits) status = 42;                                      // "unneeded data found after coded image data"
    }
    if (lane == 0) {
        jb.status = status;
        jb.padbit = padbit;
        jb.end_bitpos = p;
        jb.nrows = nrows;
    }
}

}  // namespace lepb200
