// lep_predict.cuh -- warp-cooperative predictors shared by the encode and decode kernels:
// 8x8 integer IDCT (DC forced to zero), DC prediction from neighbour edge pixels, Lakhani edge predictor.
#pragma once
#include "lep_common.cuh"

namespace lepb200 {

// Truncating signed division n / d for |n| < 2^24, 1 <= d < 2^16 without the ~40-instruction integer divide:
// float reciprocal estimate + one-step correction (exact: the estimate is off by at most one).
__device__ __forceinline__ int div_trunc_small(int n, int d) {
    const uint32_t a = (uint32_t)(n < 0 ? -n : n);
    uint32_t q = (uint32_t)(__uint2float_rz(a) * __frcp_rn(__uint2float_rz((uint32_t)d)));
    int rem = (int)a - (int)(q * (uint32_t)d);
    if (rem < 0) { --q; rem += d; }
    if (rem >= d) { ++q; }
    return n < 0 ? -(int)q : (int)q;
}

// 8-point row pass of the reference IDCT (src/lepton/idct.cc:41-111), int32 wrap-around arithmetic.
__device__ __forceinline__ void idct_row(const int32_t in[8], int32_t out[8]) {
    const int32_t w1 = 2841, w2 = 2676, w3 = 2408, w5 = 1609, w6 = 1108, w7 = 565, r2 = 181;
    int32_t x0 = (in[0] << 11) + 128, x1 = in[4] << 11, x2 = in[6], x3 = in[2], x4 = in[1], x5 = in[7], x6 = in[5], x7 = in[3];
    int32_t x8 = w7 * (x4 + x5);
    x4 = x8 + (w1 - w7) * x4;
    x5 = x8 - (w1 + w7) * x5;
    x8 = w3 * (x6 + x7);
    x6 = x8 - (w3 - w5) * x6;
    x7 = x8 - (w3 + w5) * x7;
    x8 = x0 + x1;
    x0 -= x1;
    x1 = w6 * (x3 + x2);
    x2 = x1 - (w2 + w6) * x2;
    x3 = x1 + (w2 - w6) * x3;
    x1 = x4 + x6;
    x4 -= x6;
    x6 = x5 + x7;
    x5 -= x7;
    x7 = x8 + x3;
    x8 -= x3;
    x3 = x0 + x2;
    x0 -= x2;
    x2 = (r2 * (x4 + x5) + 128) >> 8;
    x4 = (r2 * (x4 - x5) + 128) >> 8;
    out[0] = (x7 + x1) >> 8; out[1] = (x3 + x2) >> 8; out[2] = (x0 + x4) >> 8; out[3] = (x8 + x6) >> 8;
    out[4] = (x8 - x6) >> 8; out[5] = (x0 - x4) >> 8; out[6] = (x3 - x2) >> 8; out[7] = (x7 - x1) >> 8;
}
// column pass (src/lepton/idct.cc:113-161)
__device__ __forceinline__ void idct_col(const int32_t in[8], int32_t out[8]) {
    const int32_t w1 = 2841, w2 = 2676, w3 = 2408, w5 = 1609, w6 = 1108, w7 = 565, r2 = 181;
    int32_t y0 = (in[0] << 8) + 8192, y1 = in[4] << 8, y2 = in[6], y3 = in[2], y4 = in[1], y5 = in[7], y6 = in[5], y7 = in[3];
    int32_t y8 = w7 * (y4 + y5) + 4;
    y4 = (y8 + (w1 - w7) * y4) >> 3;
    y5 = (y8 - (w1 + w7) * y5) >> 3;
    y8 = w3 * (y6 + y7) + 4;
    y6 = (y8 - (w3 - w5) * y6) >> 3;
    y7 = (y8 - (w3 + w5) * y7) >> 3;
    y8 = y0 + y1;
    y0 -= y1;
    y1 = w6 * (y3 + y2) + 4;
    y2 = (y1 - (w2 + w6) * y2) >> 3;
    y3 = (y1 + (w2 - w6) * y3) >> 3;
    y1 = y4 + y6;
    y4 -= y6;
    y6 = y5 + y7;
    y5 -= y7;
    y7 = y8 + y3;
    y8 -= y3;
    y3 = y0 + y2;
    y0 -= y2;
    y2 = (r2 * (y4 + y5) + 128) >> 8;
    y4 = (r2 * (y4 - y5) + 128) >> 8;
    out[0] = (y7 + y1) >> 11; out[1] = (y3 + y2) >> 11; out[2] = (y0 + y4) >> 11; out[3] = (y8 + y6) >> 11;
    out[4] = (y8 - y6) >> 11; out[5] = (y0 - y4) >> 11; out[6] = (y3 - y2) >> 11; out[7] = (y7 - y1) >> 11;
}

// Lane-parallel 8x8 IDCT with the DC forced to zero (adv_predict_dc_pix, model.hh:674-677).
// rast = raster-order coefficients in shared memory; result in sm.pix (int16, raster order).
__device__ __forceinline__ void warp_idct_sans_dc(const int16_t* rast, const uint16_t* __restrict__ q, int32_t* tmp, int16_t* pix, int lane) {
    if (lane < 8) {
        int32_t in[8], out[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) in[k] = (int32_t)rast[lane * 8 + k] * (int32_t)q[lane * 8 + k];
        if (lane == 0) in[0] = 0;
        idct_row(in, out);
#pragma unroll
        for (int k = 0; k < 8; ++k) tmp[lane * 8 + k] = out[k];
    }
    __syncwarp();
    if (lane < 8) {
        int32_t in[8], out[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) in[k] = tmp[k * 8 + lane];
        idct_col(in, out);
#pragma unroll
        for (int k = 0; k < 8; ++k) pix[k * 8 + lane] = (int16_t)out[k];
    }
    __syncwarp();
}

// sums / min / max over groups of 8 lanes
__device__ __forceinline__ int grp8_sum(int v) {
    v += __shfl_xor_sync(FULL, v, 1); v += __shfl_xor_sync(FULL, v, 2); v += __shfl_xor_sync(FULL, v, 4); return v;
}
__device__ __forceinline__ int grp8_min(int v) {
    v = min(v, __shfl_xor_sync(FULL, v, 1)); v = min(v, __shfl_xor_sync(FULL, v, 2)); v = min(v, __shfl_xor_sync(FULL, v, 4)); return v;
}
__device__ __forceinline__ int grp8_max(int v) {
    v = max(v, __shfl_xor_sync(FULL, v, 1)); v = max(v, __shfl_xor_sync(FULL, v, 2)); v = max(v, __shfl_xor_sync(FULL, v, 4)); return v;
}
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, s, d); if (lane >= d) s += t; }
    total = __shfl_sync(FULL, s, 31);
    return s - v;
}
__device__ __forceinline__ int16_t half_rz16(int d16) { return (int16_t)(d16 / 2); }

// DC prediction from the neighbours' edge pixels (adv_predict_dc_pix, model.hh:678-784; SSE 16-bit lanes).
// left_v: lanes 0..7 hold the left block's right-column prediction, above_h: lanes 8..15 the above block's bottom row.
struct DcPred { int pred, unc, unc2; };
__device__ __forceinline__ DcPred warp_predict_dc(const int16_t* pix, int left_v, int above_h, bool has_left, bool has_above, int q0, int lane) {
    int est = 0;
    int i = lane & 7;
    if (lane < 8) {
        if (has_left) {
            int16_t p0 = pix[i * 8], p1 = pix[i * 8 + 1];
            int16_t delta = (int16_t)(p0 - p1);
            est = (int16_t)((int16_t)((int16_t)left_v - half_rz16(delta)) - (int16_t)(p0 + 1024));
        }
    } else if (lane < 16) {
        if (has_above) {
            int16_t p0 = pix[i], p1 = pix[8 + i];
            int16_t delta = (int16_t)(p0 - p1);
            est = (int16_t)((int16_t)((int16_t)above_h - half_rz16(delta)) - (int16_t)(p0 + 1024));
        }
    }
    int s = grp8_sum(est), mn = grp8_min(est), mx = grp8_max(est);
    int sl = __shfl_sync(FULL, s, 0), sa = __shfl_sync(FULL, s, 8);
    int mnl = __shfl_sync(FULL, mn, 0), mna = __shfl_sync(FULL, mn, 8);
    int mxl = __shfl_sync(FULL, mx, 0), mxa = __shfl_sync(FULL, mx, 8);
    DcPred r; r.pred = 0; r.unc = 0; r.unc2 = 0;
    int avgmed = 0;
    if (has_left || has_above) {
        int a0, a1, mn_all, mx_all;
        if (has_left && has_above) { a0 = sl; a1 = sa; mn_all = min(mnl, mna); mx_all = max(mxl, mxa); }
        else if (has_left) { a0 = a1 = sl; mn_all = mnl; mx_all = mxl; }
        else { a0 = a1 = sa; mn_all = mna; mx_all = mxa; }
        avgmed = (a0 + a1) >> 1;
        r.unc = (mx_all - mn_all) >> 3;
        a0 -= avgmed; a1 -= avgmed;
        int far_afield = a1;
        if (iabs(a0) < iabs(a1)) far_afield = a0;
        r.unc2 = far_afield >> 3;
    }
    r.pred = (div_trunc_small(avgmed, q0) + 4) >> 3;            // |avgmed| < 2^20
    return r;
}

// adv_predict_or_unpredict_dc (model.hh:823-832)
__device__ __forceinline__ int adv_unpredict(int saved_dc, bool recover, int pred) {
    int r = saved_dc + (recover ? pred : -pred);
    if (r < -1024) r += 2049;
    if (r > 1024) r -= 2049;
    return r;
}

// NeighborSummary::set_horizontal/set_vertical (block_context.hh:44-78): lanes 0..7 -> vertical (right column),
// lanes 8..15 -> horizontal (bottom row); 16-bit wrap.
__device__ __forceinline__ int edge_pixel(const int16_t* pix, int q0, int dc, int lane) {
    int i = lane & 7;
    int16_t cur, prev;
    if (lane < 8) { cur = pix[i * 8 + 7]; prev = pix[i * 8 + 6]; }
    else { cur = pix[56 + i]; prev = pix[48 + i]; }
    int16_t delta = (int16_t)(cur - prev);
    int16_t qdc = (int16_t)((uint32_t)q0 * (uint32_t)dc);
    return (int16_t)(cur + half_rz16(delta) + 1024 + qdc);
}

// compute_lak (model.hh:1033-1071) for one edge coefficient, int32 wrap-around then C division.
//   horizontal (band = k, 1..7): x[i] = cur(k + 8i), a[i] = above(k + 8i), icos = icos_x[k*8 + i]
//   vertical   (band = 8k):      x[i] = cur(8k + i), a[i] = left(8k + i),  icos = icos_y[8k + i]
__device__ __forceinline__ int lak_pred(const int16_t* cur, const int16_t* nb, const int32_t* __restrict__ icos, int first, int step) {
    uint32_t pred = (uint32_t)(int32_t)nb[first] * (uint32_t)icos[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        int32_t t = (int32_t)cur[first + i * step] + ((i & 1) ? (int32_t)nb[first + i * step] : -(int32_t)nb[first + i * step]);
        pred -= (uint32_t)icos[i] * (uint32_t)t;
    }
    // C division by icos[0] = 8192 * q (model.hh:957,1064): trunc(trunc(p / 8192) / q) == trunc(p / (8192 q))
    const int32_t p = (int32_t)pred;
    const int32_t t = (p + ((p >> 31) & 8191)) >> 13;
    return div_trunc_small(t, icos[0] >> 13);
}

}  // namespace lepb200
