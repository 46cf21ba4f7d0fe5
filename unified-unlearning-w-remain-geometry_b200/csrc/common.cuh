// common.cuh — shared device helpers for the SFR-on hot-path kernels (sm_100a).
//
// Every kernel on this path is an HBM-bound stream over a flat parameter shard:
//   * 128-bit coalesced accesses (one float4 per thread per access, a warp covers
//     512 contiguous bytes), several independent accesses in flight per thread;
//   * cache policy by stream role, measured on B200 (tools/tune/tune_stream.cu, tools/sweep.py):
//     read-modify-write kernels (K1, K2a, K3, EWC, shrink) use PLAIN ld.global / st.global — the
//     evict-first / no-allocate hints cost them 3-13 %; READ-ONLY kernels (clip norm, select
//     histograms, tie count, apply's key reads) use ld.global.cs (evict-first), which is 8-15 %
//     faster for them than default caching;
//   * small CTAs (128 threads) and LARGE grids — one tile per CTA for the update shape, SMs x 16 x 32
//     CTAs for the two-read-one-write shape: the hardware block scheduler keeps all SMs on one
//     moving window of addresses; measured 6.8-6.9 TB/s vs 5.7-6.1 TB/s for a persistent
//     SMs x resident-CTAs grid-stride loop (same sweep);
//   * reductions: warp shuffle -> shared memory -> ONE atomic per CTA.
// No fast-math: the file is compiled with -fmad=false and uses explicit
// __f*_rn / __fmaf_rn so the rounding sequence is exactly the one documented in
// DESIGN.md ("Arithmetic contract").
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sfron_b200.h"

namespace sfr {

constexpr unsigned kFullMask = 0xffffffffu;

// ---- launch geometry --------------------------------------------------------
struct DeviceGeometry {
  int sm_count = 0;
  int cc_major = 0;
  int cc_minor = 0;
  bool ok = false;
};

// Geometry of the device the calling thread has entered (DeviceScope below), cached per device ordinal.
const DeviceGeometry& device_geometry();

// Every entry point runs on the device that OWNS its buffers, not on whatever device happens to be
// current: the scope looks up the device of an anchor pointer, switches to it for the duration of the
// call (launches, function attributes, geometry) and switches back.  One process per GPU is the normal
// deployment, but a process that drives several GPUs (or has not called cudaSetDevice) gets the same results.
class DeviceScope {
 public:
  explicit DeviceScope(const void* anchor);
  ~DeviceScope();
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
  bool ok() const { return ok_; }
  int device() const { return dev_; }

 private:
  int prev_ = -1;
  int dev_ = -1;
  int outer_ = -1;
  bool switched_ = false;
  bool entered_ = false;
  bool ok_ = false;
};

// Persistent grid: enough CTAs to fill every SM `ctas_per_sm` times, but never more
// than there are tiles.
inline int persistent_grid(int64_t tiles, int ctas_per_sm) {
  const DeviceGeometry& g = device_geometry();
  int64_t cap = (int64_t)(g.sm_count > 0 ? g.sm_count : 148) * ctas_per_sm;
  if (tiles < 1) tiles = 1;
  return (int)(tiles < cap ? tiles : cap);
}

// One tile per CTA (every kernel still loops grid-stride, so the 2^31-1 cap is harmless).
inline int full_grid(int64_t tiles) {
  if (tiles < 1) tiles = 1;
  return (int)(tiles < 2147483647LL ? tiles : 2147483647LL);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- vector access ---------------------------------------------------------------
// ld_stream / st_stream: streams of read-modify-write kernels (default caching).
// ld_once: data a read-mostly kernel (the select passes) reads exactly once and never writes.  Default caching: on B200 the
// evict-first hint (__ldcs) costs these kernels 7 % (tools/tune/tune_hist1.cu, ZL0 / ZL1 and Z / Zp: 6.96 vs 6.46 TB/s).
__device__ __forceinline__ float4 ld_once(const float4* p) { return *p; }
__device__ __forceinline__ float ld_once(const float* p) { return *p; }
__device__ __forceinline__ float4 ld_stream(const float4* p) { return *p; }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { *p = v; }
__device__ __forceinline__ float ld_stream(const float* p) { return *p; }
__device__ __forceinline__ void st_stream(float* p, float v) { *p = v; }

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t b16) { return __uint_as_float(b16 << 16); }

// Four consecutive gradients starting at element 4*vec, as fp32.
template <int GT>
__device__ __forceinline__ float4 load_g4(const void* g, int64_t vec) {
  if constexpr (GT == SFR_F32) {
    return *(reinterpret_cast<const float4*>(g) + vec);
  } else {
    uint2 raw = *(reinterpret_cast<const uint2*>(g) + vec);
    float4 r;
    r.x = bf16_bits_to_f32(raw.x & 0xffffu);
    r.y = bf16_bits_to_f32(raw.x >> 16);
    r.z = bf16_bits_to_f32(raw.y & 0xffffu);
    r.w = bf16_bits_to_f32(raw.y >> 16);
    return r;
  }
}
template <int GT>
__device__ __forceinline__ float load_g1(const void* g, int64_t i) {
  if constexpr (GT == SFR_F32) {
    return *(reinterpret_cast<const float*>(g) + i);
  } else {
    return bf16_bits_to_f32(*(reinterpret_cast<const unsigned short*>(g) + i));
  }
}
template <int GT>
__device__ __forceinline__ void zero_g4(void* g, int64_t vec) {
  if constexpr (GT == SFR_F32) {
    *(reinterpret_cast<float4*>(g) + vec) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    *(reinterpret_cast<uint2*>(g) + vec) = make_uint2(0u, 0u);
  }
}
template <int GT>
__device__ __forceinline__ void zero_g1(void* g, int64_t i) {
  if constexpr (GT == SFR_F32) {
    reinterpret_cast<float*>(g)[i] = 0.f;
  } else {
    reinterpret_cast<unsigned short*>(g)[i] = 0;
  }
}

// Gradient and mask words are read once per kernel, with DEFAULT caching: an A/B at N3 of the whole library built with
// __ldcs here instead (profiles/r2_ab_load_hint.jsonl) has the clip norm 4.7 % slower and everything else within noise.
template <int GT>
__device__ __forceinline__ float4 load_g4_once(const void* g, int64_t vec) {
  if constexpr (GT == SFR_F32) {
    return *(reinterpret_cast<const float4*>(g) + vec);
  } else {
    uint2 raw = *(reinterpret_cast<const uint2*>(g) + vec);
    float4 r;
    r.x = bf16_bits_to_f32(raw.x & 0xffffu);
    r.y = bf16_bits_to_f32(raw.x >> 16);
    r.z = bf16_bits_to_f32(raw.y & 0xffffu);
    r.w = bf16_bits_to_f32(raw.y >> 16);
    return r;
  }
}
__device__ __forceinline__ uint32_t load_mask4_once(const uint8_t* mask, int64_t vec) {
  return *(reinterpret_cast<const unsigned int*>(mask) + vec);
}

// Four mask bytes (0/1) for elements 4*vec .. 4*vec+3.
__device__ __forceinline__ uint32_t load_mask4(const uint8_t* mask, int64_t vec) {
  return *(reinterpret_cast<const unsigned int*>(mask) + vec);
}
__device__ __forceinline__ float mask_byte_to_f32(uint32_t packed, int lane) {
  return (float)((packed >> (8 * lane)) & 0xffu);
}

// ---- clip coefficient ------------------------------------------------------------
// torch.nn.utils.clip_grad_norm_:  coef = clamp(max_norm / (total_norm + 1e-6), max=1.0)
// total_norm is formed from the double-precision sum of squares (at least as accurate
// as torch's fp32 norm of per-tensor norms).
__device__ __forceinline__ float clip_coef_from_sumsq(const double* sumsq, float max_norm) {
  float total_norm = (float)sqrt(*sumsq);
  float coef = __fdiv_rn(max_norm, __fadd_rn(total_norm, 1e-6f));
  return coef > 1.0f ? 1.0f : coef;  // keeps NaN, like torch.clamp(max=1.0)
}

// The same coefficient with one double sqrt per WARP instead of per thread (FP64 is the scarce pipe;
// one-tile CTAs make a per-thread sqrt run ~1.7e8 times per launch).  Must be called by full warps.
__device__ __forceinline__ float clip_coef_warp(const double* sumsq, float max_norm) {
  float c = 0.f;
  if ((threadIdx.x & 31) == 0) c = clip_coef_from_sumsq(sumsq, max_norm);
  return __shfl_sync(kFullMask, c, 0);
}

// ---- reductions -----------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// Sum over the CTA; result valid in thread 0.  `scratch` holds >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T r = T(0);
  if (warp == 0) {
    const int nwarps = (blockDim.x + 31) >> 5;
    r = lane < nwarps ? scratch[lane] : T(0);
    r = warp_sum(r);
  }
  return r;
}

// ---- order-preserving key of |x| (K2b) -----------------------------------------------
// 0 for NaN (ranks last, never selected before any number), else bits(|x|) + 1.
__device__ __forceinline__ uint32_t select_key(float x) {
  uint32_t b = __float_as_uint(x) & 0x7fffffffu;
  return b > 0x7f800000u ? 0u : b + 1u;
}

}  // namespace sfr

// ---- host-side argument checks ---------------------------------------------------------
#define SFR_REQUIRE_PTR(p) \
  do {                     \
    if ((p) == nullptr) return SFR_ERR_NULL; \
  } while (0)
#define SFR_REQUIRE_ALIGNED(p) \
  do {                         \
    if ((p) != nullptr && !sfr::aligned16(p)) return SFR_ERR_ALIGN; \
  } while (0)
#define SFR_ENTER_DEVICE(anchor)              \
  sfr::DeviceScope device_scope__(anchor);    \
  if (!device_scope__.ok()) return SFR_ERR_NO_DEVICE
#define SFR_LAUNCH_STATUS()                \
  do {                                     \
    cudaError_t e__ = cudaGetLastError();  \
    return e__ == cudaSuccess ? SFR_OK : (int)e__; \
  } while (0)
