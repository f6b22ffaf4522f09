// lep_huff.cu -- baseline JPEG Huffman decode on the GPU.
//
// Moves the 16x data expansion (JPEG bytes -> 128 B/block coefficient planes) onto the device: the host only
// splits the file into header / de-stuffed entropy bytes (lep_jpeg.cc parse_jpeg) and uploads the entropy bytes;
// this kernel produces the coefficient planes in AlignedBlock order directly in HBM, plus, per MCU row, the
// resumable Huffman state the reference calls a ThreadHandoff (bit position, last DCs;
// crystallize_thread_handoff, src/lepton/jpgcoder.cc:2520-2560).
//
// A baseline scan without restart markers is one serial bit stream, so the unit of parallelism is the image:
// ONE THREAD PER IMAGE (a batch of thousands of files keeps the chip busy; the kernel is latency-bound per thread,
// its duration is that of the largest file).  Semantics follow decode_jpeg / decode_block_seq
// (jpgcoder.cc:2799-3302, :4893-4961) exactly like the host decoder in lep_jpeg.cc, against which it is tested.
#include "lep_common.cuh"

namespace lepb200 {

struct HuffTableDev {
    uint16_t fast[512];      // (len << 8) | symbol for codes of <= 9 bits, 0 = longer code
    int32_t maxcode[18];
    int32_t valoff[18];
    uint8_t vals[256];
};

struct HuffRow {             // state at the start of an MCU row
    uint32_t bitpos;         // bits consumed of the de-stuffed entropy stream
    int16_t lastdc[3];
    int16_t mcu_y;
};

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

constexpr int HUFF_THREADS = 32;
constexpr int HUFF_SMEM_TABLES = 8;      // tables staged in shared memory when the batch uses few distinct ones

// One symbol per loop iteration (flat state machine): lanes of a warp decode different images, and a nested
// "for each block / while AC" loop would make every lane wait for the slowest block of the warp at each block end.
__global__ void __launch_bounds__(HUFF_THREADS)
lep_huffdecode_kernel(HuffJob* __restrict__ jobs, int njobs, const HuffTableDev* __restrict__ tables, int ntables) {
    __shared__ HuffTableDev s_tab[HUFF_SMEM_TABLES];
    __shared__ uint8_t s_zz[64];          // per-lane indices differ: constant memory would serialise the lookups
    for (int i = threadIdx.x; i < 64; i += HUFF_THREADS) s_zz[i] = c_zigzag_to_aligned[i];
    const bool use_smem = ntables <= HUFF_SMEM_TABLES;
    if (use_smem) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_tab);
        const int nw = ntables * (int)(sizeof(HuffTableDev) / 4);
        for (int i = threadIdx.x; i < nw; i += HUFF_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const HuffTableDev* tb = use_smem ? s_tab : tables;
    const int t = blockIdx.x * HUFF_THREADS + threadIdx.x;
    if (t >= njobs) return;
    HuffJob& jb = jobs[t];
    if (jb.status != 0) return;
    HBits b;
    b.w = reinterpret_cast<const uint32_t*>(jb.huff);
    b.acc = 0; b.n = 0; b.wi = 0; b.nwords = (jb.nbytes + 3) / 4; b.bitpos = 0;
    const unsigned long long total_bits = (unsigned long long)jb.nbytes * 8;
    HuffRow* rows = reinterpret_cast<HuffRow*>(jb.rows);
    // everything the per-block bookkeeping needs lives in registers; positions are tracked incrementally so that the
    // block-end path (which some lane of the warp takes on almost every iteration) has no divisions and no loads
    const int ncmp = jb.ncmp, mcuh = jb.mcuh, mcuv = jb.mcuv, rsti = jb.rsti;
    const int H0 = jb.H[0], V0 = jb.V[0], H1 = ncmp > 1 ? jb.H[1] : 1, V1 = ncmp > 1 ? jb.V[1] : 1, H2 = ncmp > 2 ? jb.H[2] : 1, V2 = ncmp > 2 ? jb.V[2] : 1;
    const int W0 = jb.bch[0], W1 = ncmp > 1 ? jb.bch[1] : 0, W2 = ncmp > 2 ? jb.bch[2] : 0;
    int16_t* const P0 = reinterpret_cast<int16_t*>(jb.plane[0]);
    int16_t* const P1 = reinterpret_cast<int16_t*>(jb.plane[1]);
    int16_t* const P2 = reinterpret_cast<int16_t*>(jb.plane[2]);
    const HuffTableDev* const D0 = tb + jb.dc_tab[0]; const HuffTableDev* const A0 = tb + jb.ac_tab[0];
    const HuffTableDev* const D1 = tb + jb.dc_tab[1]; const HuffTableDev* const A1 = tb + jb.ac_tab[1];
    const HuffTableDev* const D2 = tb + jb.dc_tab[2]; const HuffTableDev* const A2 = tb + jb.ac_tab[2];
    const int nch0 = jb.nch[0], ncv0 = jb.ncv[0];
    int dc0 = 0, dc1 = 0, dc2 = 0;                   // last DCs
    int cmp = 0, sx = 0, sy = 0, mx = 0, my = 0;     // component, block within MCU, MCU position
    int bx = 0, by = 0;                              // single-component scans: block position in the plane
    int nrows = 0, status = 0, padbit = -1;
    int rstw = rsti;
    int bpos = 0;                                    // 0: next symbol is the block's DC; 1..63: next AC position
    bool last_nonzero = true;
    int16_t* blk = P0;
    const HuffTableDev* dct = D0;
    const HuffTableDev* act = A0;
    {
        HuffRow r;
        r.bitpos = 0; r.mcu_y = 0; r.lastdc[0] = r.lastdc[1] = r.lastdc[2] = 0;
        rows[nrows++] = r;
    }
    while (true) {
        // ---- one Huffman symbol (+ its magnitude bits)
        hb_fill(b);
        const HuffTableDev* tab = bpos == 0 ? dct : act;
        const uint32_t top = hb_peek(b, 16);
        int sym;
        {
            const uint32_t f = tab->fast[top >> 7];
            if (f) { hb_skip(b, (int)(f >> 8)); sym = (int)(f & 0xff); }
            else {
                int len = 10, code = (int)(top >> 6);
                while (len <= 16 && code > tab->maxcode[len]) { ++len; code = (int)(top >> (16 - len)); }
                if (len > 16) { status = 42; break; }
                hb_skip(b, len);
                sym = tab->vals[code + tab->valoff[len]];
            }
        }
        bool block_done = false;
        if (bpos == 0) {
            if (sym > 16) { status = 42; break; }
            const int last = cmp == 0 ? dc0 : (cmp == 1 ? dc1 : dc2);
            const int16_t dcv = (int16_t)(huff_extend(b, sym) + last);
            if (cmp == 0) dc0 = dcv; else if (cmp == 1) dc1 = dcv; else dc2 = dcv;
            blk[49] = dcv;
            bpos = 1;
            last_nonzero = true;
        } else if (sym == 0) {                                    // EOB
            if (bpos > 1 && !last_nonzero) { status = 42; break; }   // "eob after last 0" (jpgcoder.cc:2953)
            block_done = true;
        } else {
            const int z = sym >> 4, sz = sym & 15;
            const int v = huff_extend(b, sz);
            if (z + bpos >= 64) { status = 200; break; }           // truncated-file fix-up path: not handled here
            bpos += z;
            blk[s_zz[bpos++]] = (int16_t)v;
            last_nonzero = v != 0;
            block_done = bpos >= 64;
        }
        if (!block_done) continue;
        if (b.bitpos > total_bits) { status = 200; break; }         // entropy data ends inside a block
        // ---- next block position (next_mcupos / next_mcuposn), incremental
        int sta = 0;
        bool handoff_due = false;
        if (ncmp > 1) {
            const int H = cmp == 0 ? H0 : (cmp == 1 ? H1 : H2), V = cmp == 0 ? V0 : (cmp == 1 ? V1 : V2);
            if (++sx >= H) { sx = 0; ++sy; }
            if (sy >= V) {                                          // component done within this MCU
                sy = 0;
                if (++cmp >= ncmp) {
                    cmp = 0;
                    if (++mx >= mcuh) { mx = 0; ++my; handoff_due = true; }
                    if (my >= mcuv) sta = 2;
                    else if (rsti > 0 && --rstw == 0) sta = 1;
                }
                dct = cmp == 0 ? D0 : (cmp == 1 ? D1 : D2);
                act = cmp == 0 ? A0 : (cmp == 1 ? A1 : A2);
            }
            const int Hn = cmp == 0 ? H0 : (cmp == 1 ? H1 : H2), Vn = cmp == 0 ? V0 : (cmp == 1 ? V1 : V2);
            const int Wn = cmp == 0 ? W0 : (cmp == 1 ? W1 : W2);
            int16_t* Pn = cmp == 0 ? P0 : (cmp == 1 ? P1 : P2);
            const int dpos = (my * Vn + sy) * Wn + mx * Hn + sx;
            blk = Pn + (size_t)dpos * 64;
        } else {
            // next_mcuposn (jpgcoder.cc:5432-5456): row-major over the nch x ncv coded blocks of the bch x bcv plane
            if (++bx >= nch0) { bx = 0; ++by; }
            if (by >= ncv0) sta = 2;
            else if (rsti > 0 && --rstw == 0) sta = 1;
            const int dpos = by * W0 + bx;
            blk = P0 + (size_t)dpos * 64;
            // handoff when the block index is a multiple of one "MCU row" worth of blocks (jpgcoder.cc:3084-3087)
            const int per_mcu = H0 * V0;
            if (sta != 2 && (dpos % per_mcu) == 0 && ((dpos / per_mcu) % mcuh) == 0) handoff_due = true;
            if (sta != 2) { my = (dpos / per_mcu) / mcuh; }
        }
        bpos = 0;
        if (b.bitpos >= total_bits) sta = 2;                          // huffr->eof
        if (sta != 0) {
            // abitreader::unpad (bitops.hh:316-332) + padbit bookkeeping (jpgcoder.cc:3260-3271)
            int fb = padbit;
            if ((b.bitpos & 7) != 0 && b.bitpos < total_bits) {
                hb_fill(b);
                int last = (int)hb_peek(b, 1); hb_skip(b, 1);
                fb = last;
                int offset = 1;
                while (b.bitpos & 7) { hb_fill(b); last = (int)hb_peek(b, 1); hb_skip(b, 1); fb |= last << offset; ++offset; }
                while (offset < 7) { fb |= last << offset; ++offset; }
            }
            if (padbit != -1) { if (padbit != fb) { status = 42; break; } }
            else padbit = fb;
            if (sta == 2) break;
            dc0 = dc1 = dc2 = 0;                                      // restart interval
            rstw = rsti;
        }
        if (handoff_due) {
            HuffRow r;
            r.bitpos = (uint32_t)b.bitpos; r.mcu_y = (int16_t)my;
            r.lastdc[0] = (int16_t)dc0; r.lastdc[1] = (int16_t)dc1; r.lastdc[2] = (int16_t)dc2;
            rows[nrows++] = r;
        }
    }
    const int final_mcu_y = ncmp > 1 ? my : (by >= ncv0 ? mcuv : my);
    if (status == 0) {
        HuffRow r;
        r.bitpos = (uint32_t)b.bitpos; r.mcu_y = (int16_t)final_mcu_y;
        r.lastdc[0] = (int16_t)dc0; r.lastdc[1] = (int16_t)dc1; r.lastdc[2] = (int16_t)dc2;
        rows[nrows++] = r;
        if (b.bitpos < total_bits) status = 42;                      // "unneeded data found after coded image data"
    }
    jb.status = status;
    jb.padbit = padbit;
    jb.end_bitpos = (uint32_t)b.bitpos;
    jb.nrows = nrows;
}

}  // namespace lepb200
