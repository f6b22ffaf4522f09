// cuda_shim.h -- just enough of the CUDA device dialect to compile the thread-per-segment decode kernels
// (lep_decode_thread.cu, lep_decode_lockstep.cu) as host C++ and run ONE lane at a time.
//
// Test infrastructure only.  These kernels give every lane its own segment and use no cross-lane data exchange: the only
// warp-level operations are votes that decide how long the lanes of a warp keep stepping together.  With a warp of one
// lane a vote is the lane's own predicate, so the per-lane arithmetic (bool decoder, token grammar, predictors, IDCT,
// block stores) runs exactly as on the device and can be checked bit for bit against the oracle without a GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __constant__ const
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)

struct emu_dim3 { unsigned x = 0, y = 0, z = 0; };
static emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }

static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __any_sync(unsigned, int p) { return p != 0; }
static inline int __all_sync(unsigned, int p) { return p != 0; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __ldg(const T* p) { return *p; }

static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
// cvt.rz.f32.u32: the largest float not above the integer
static inline float __uint2float_rz(uint32_t a) {
    float f = (float)a;
    if ((double)f > (double)a) f = std::nextafterf(f, 0.0f);
    return f;
}
static inline float __frcp_rn(float x) { return 1.0f / x; }          // IEEE round-to-nearest reciprocal

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
