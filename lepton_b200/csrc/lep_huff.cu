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

// One WARP per group of up to 32 images WITH IDENTICAL GEOMETRY (the host sorts the batch): the walk over MCUs /
// components / blocks is then warp-uniform and costs a handful of instructions per block, and only the symbol loop
// inside a block is per-lane (lanes wait for the longest block of the group).  With per-lane geometry every branch of
// the bookkeeping is taken by some lane on almost every iteration and the warp executes all of it all the time.
__global__ void __launch_bounds__(HUFF_THREADS)
lep_huffdecode_kernel(HuffJob* __restrict__ jobs, const int2* __restrict__ groups, int ngroups,
                      const HuffTableDev* __restrict__ tables, int ntables) {
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
    if ((int)blockIdx.x >= ngroups) return;
    const int2 grp = groups[blockIdx.x];                 // (first job, count <= 32)
    const int lane = threadIdx.x;
    const bool have = lane < grp.y;
    HuffJob& jb = jobs[grp.x + (have ? lane : 0)];       // idle lanes shadow lane 0's geometry, never write
    bool live = have && jb.status == 0;

    HBits b;
    b.w = reinterpret_cast<const uint32_t*>(jb.huff);
    b.acc = 0; b.n = 0; b.wi = 0; b.nwords = (jb.nbytes + 3) / 4; b.bitpos = 0;
    const unsigned long long total_bits = (unsigned long long)jb.nbytes * 8;
    HuffRow* rows = reinterpret_cast<HuffRow*>(jb.rows);
    // geometry: identical for every lane of the group (taken from the group's first job)
    const HuffJob& g0 = jobs[grp.x];
    const int ncmp = g0.ncmp, mcuh = g0.mcuh, mcuv = g0.mcuv, rsti = g0.rsti;
    const int H0 = g0.H[0], V0 = g0.V[0], H1 = ncmp > 1 ? g0.H[1] : 1, V1 = ncmp > 1 ? g0.V[1] : 1, H2 = ncmp > 2 ? g0.H[2] : 1, V2 = ncmp > 2 ? g0.V[2] : 1;
    const int W0 = g0.bch[0], W1 = ncmp > 1 ? g0.bch[1] : 0, W2 = ncmp > 2 ? g0.bch[2] : 0;
    const int nch0 = g0.nch[0], ncv0 = g0.ncv[0];
    // per-lane: planes and tables
    int16_t* const P0 = reinterpret_cast<int16_t*>(jb.plane[0]);
    int16_t* const P1 = reinterpret_cast<int16_t*>(jb.plane[1]);
    int16_t* const P2 = reinterpret_cast<int16_t*>(jb.plane[2]);
    const HuffTableDev* const D0 = tb + jb.dc_tab[0]; const HuffTableDev* const A0 = tb + jb.ac_tab[0];
    const HuffTableDev* const D1 = tb + jb.dc_tab[1]; const HuffTableDev* const A1 = tb + jb.ac_tab[1];
    const HuffTableDev* const D2 = tb + jb.dc_tab[2]; const HuffTableDev* const A2 = tb + jb.ac_tab[2];
    int dc0 = 0, dc1 = 0, dc2 = 0;
    int nrows = 0, status = 0, padbit = -1, final_y = 0;
    int rstw = rsti;
    bool finished = false;                    // this lane reached the end of its scan (sta == 2)

    // state shared by both scan shapes: called at the end of a restart interval / of the scan
    auto unpad_and_check = [&]() {
        int fb = padbit;
        if ((b.bitpos & 7) != 0 && b.bitpos < total_bits) {
            hb_fill(b);
            int last = (int)hb_peek(b, 1); hb_skip(b, 1);
            fb = last;
            int offset = 1;
            while (b.bitpos & 7) { hb_fill(b); last = (int)hb_peek(b, 1); hb_skip(b, 1); fb |= last << offset; ++offset; }
            while (offset < 7) { fb |= last << offset; ++offset; }
        }
        if (padbit != -1) { if (padbit != fb) status = 42; }
        else padbit = fb;
    };
    auto push_row = [&](int mcu_y) {
        HuffRow r;
        r.bitpos = (uint32_t)b.bitpos; r.mcu_y = (int16_t)mcu_y;
        r.lastdc[0] = (int16_t)dc0; r.lastdc[1] = (int16_t)dc1; r.lastdc[2] = (int16_t)dc2;
        rows[nrows++] = r;
    };
    // one block for this lane: DC symbol, AC symbols until EOB / 63 (decode_block_seq, jpgcoder.cc:4893-4961)
    auto decode_block = [&](int cmp, int16_t* blk) {
        const HuffTableDev* dct = cmp == 0 ? D0 : (cmp == 1 ? D1 : D2);
        const HuffTableDev* act = cmp == 0 ? A0 : (cmp == 1 ? A1 : A2);
        int bpos = 0;
        bool last_nonzero = true;
        bool active = live && !finished && status == 0;
        while (__any_sync(FULL, active)) {
            if (active) {
                hb_fill(b);                                   // >= 33 bits: enough for one code (<= 16) + magnitude (<= 16)
                const HuffTableDev* tab = bpos == 0 ? dct : act;
                const uint32_t top = hb_peek(b, 16);
                int sym = -1;
                const uint32_t f = tab->fast[top >> 7];
                if (f) { hb_skip(b, (int)(f >> 8)); sym = (int)(f & 0xff); }
                else {
                    int len = 10, code = (int)(top >> 6);
                    while (len <= 16 && code > tab->maxcode[len]) { ++len; code = (int)(top >> (16 - len)); }
                    if (len <= 16) { hb_skip(b, len); sym = tab->vals[code + tab->valoff[len]]; }
                }
                if (sym < 0 || (bpos == 0 && sym > 16)) { status = 42; active = false; }
                else if (bpos == 0) {
                    int v = 0;
                    if (sym) { const int n = (int)hb_peek(b, sym); hb_skip(b, sym); v = n >= (1 << (sym - 1)) ? n : n + 1 - (1 << sym); }
                    const int last = cmp == 0 ? dc0 : (cmp == 1 ? dc1 : dc2);
                    const int16_t dcv = (int16_t)(v + last);
                    if (cmp == 0) dc0 = dcv; else if (cmp == 1) dc1 = dcv; else dc2 = dcv;
                    blk[49] = dcv;
                    bpos = 1;
                } else if (sym == 0) {                                        // EOB
                    if (bpos > 1 && !last_nonzero) status = 42;               // "eob after last 0" (jpgcoder.cc:2953)
                    active = false;
                } else {
                    const int z = sym >> 4, sz = sym & 15;
                    int v = 0;
                    if (sz) { const int n = (int)hb_peek(b, sz); hb_skip(b, sz); v = n >= (1 << (sz - 1)) ? n : n + 1 - (1 << sz); }
                    if (z + bpos >= 64) { status = 200; active = false; }     // truncated-file fix-up path: not handled here
                    else {
                        bpos += z;
                        blk[s_zz[bpos++]] = (int16_t)v;
                        last_nonzero = v != 0;
                        if (bpos >= 64) active = false;
                    }
                }
            }
        }
        if (live && !finished && status == 0 && b.bitpos > total_bits) status = 200;   // entropy data ends inside a block
    };
    // after a block: end of data / restart interval / end of scan handling for this lane
    auto after_block = [&](bool scan_done, bool restart_due, int mcu_y_now) {
        if (!live || finished || status) return;
        int sta = scan_done ? 2 : (restart_due ? 1 : 0);
        if (b.bitpos >= total_bits) sta = 2;                                  // huffr->eof
        if (sta == 0) return;
        unpad_and_check();
        if (status) return;
        if (sta == 2) { finished = true; final_y = mcu_y_now; return; }
        dc0 = dc1 = dc2 = 0;                                                  // restart interval
    };

    if (live) push_row(0);
    if (ncmp > 1) {
        for (int my = 0; my < mcuv; ++my) {
            for (int mx = 0; mx < mcuh; ++mx) {
                for (int cmp = 0; cmp < ncmp; ++cmp) {
                    const int H = cmp == 0 ? H0 : (cmp == 1 ? H1 : H2), V = cmp == 0 ? V0 : (cmp == 1 ? V1 : V2);
                    const int W = cmp == 0 ? W0 : (cmp == 1 ? W1 : W2);
                    int16_t* P = cmp == 0 ? P0 : (cmp == 1 ? P1 : P2);
                    for (int sy = 0; sy < V; ++sy)
                        for (int sx = 0; sx < H; ++sx) {
                            const bool last_in_mcu = cmp == ncmp - 1 && sy == V - 1 && sx == H - 1;
                            decode_block(cmp, P + ((size_t)(my * V + sy) * W + mx * H + sx) * 64);
                            if (!last_in_mcu) {
                                // the reference checks eof after every block (jpgcoder.cc:2975)
                                if (live && !finished && !status && b.bitpos >= total_bits) { unpad_and_check(); finished = true; final_y = my; }
                            }
                        }
                }
                // MCU complete (next_mcupos, recoder.cc:190-243)
                const bool scan_done = my == mcuv - 1 && mx == mcuh - 1;
                bool restart_due = false;
                if (!scan_done && rsti > 0 && --rstw == 0) { restart_due = true; rstw = rsti; }
                const int next_y = mx == mcuh - 1 ? my + 1 : my;
                after_block(scan_done, restart_due, next_y);
                if (mx == mcuh - 1 && !scan_done && live && !finished && !status) push_row(my + 1);
            }
            if (!__any_sync(FULL, live && !finished && status == 0)) break;
        }
    } else {
        // single component: row-major over the nch x ncv coded blocks (next_mcuposn, jpgcoder.cc:5432-5456)
        const int per_mcu = H0 * V0;
        for (int by = 0; by < ncv0; ++by) {
            for (int bx = 0; bx < nch0; ++bx) {
                decode_block(0, P0 + ((size_t)by * W0 + bx) * 64);
                const bool scan_done = by == ncv0 - 1 && bx == nch0 - 1;
                bool restart_due = false;
                if (!scan_done && rsti > 0 && --rstw == 0) { restart_due = true; rstw = rsti; }
                // position of the NEXT block decides the handoff (jpgcoder.cc:3084-3087)
                const int nbx = bx + 1 < nch0 ? bx + 1 : 0, nby = bx + 1 < nch0 ? by : by + 1;
                const int ndpos = nby * W0 + nbx;
                const int nmcu_y = (ndpos / per_mcu) / mcuh;
                after_block(scan_done, restart_due, scan_done ? mcuv : nmcu_y);
                if (!scan_done && (ndpos % per_mcu) == 0 && ((ndpos / per_mcu) % mcuh) == 0 && live && !finished && !status) push_row(nmcu_y);
            }
            if (!__any_sync(FULL, live && !finished && status == 0)) break;
        }
    }
    if (have) {
        if (live && status == 0) {
            if (!finished) final_y = mcuv;
            push_row(final_y);
            if (b.bitpos < total_bits) status = 42;                      // "unneeded data found after coded image data"
        }
        if (jb.status == 0) jb.status = status;
        jb.padbit = padbit;
        jb.end_bitpos = (uint32_t)b.bitpos;
        jb.nrows = nrows;
    }
}

}  // namespace lepb200
