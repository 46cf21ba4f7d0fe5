// peer_common.cuh — device-side vocabulary of the cross-GPU exchange kernels (peer.cu: loads / stores and NVLS
// multimem issued by the compute threads; peer_tma.cu: the same data moved by TMA bulk copies through shared memory).
#pragma once

#include "update_core.cuh"

namespace sfr {

constexpr int kMaxPeers = SFR_MAX_PEERS;
struct PeerPtrs {
  void* p[kMaxPeers];
};

// ---- multimem (NVLS) primitives ------------------------------------------------------------------
__device__ __forceinline__ float4 mc_ld_reduce_f32x4(const void* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
// four bf16 values (8 bytes), summed by the switch with fp32 accumulation, returned rounded to bf16
__device__ __forceinline__ uint2 mc_ld_reduce_bf16x4(const void* mc) {
  uint2 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v2.bf16x2 {%0,%1}, [%2];"
               : "=r"(r.x), "=r"(r.y)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void mc_st_b128(void* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :
               : "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mc_st_b64(void* mc, uint2 v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};"
               :
               : "l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y))
               : "memory");
}

__device__ __forceinline__ float4 widen_bf16x4(uint2 raw) {
  float4 r;
  r.x = bf16_bits_to_f32(raw.x & 0xffffu);
  r.y = bf16_bits_to_f32(raw.x >> 16);
  r.z = bf16_bits_to_f32(raw.y & 0xffffu);
  r.w = bf16_bits_to_f32(raw.y >> 16);
  return r;
}
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 packed;
  packed.x = *reinterpret_cast<uint32_t*>(&lo);
  packed.y = *reinterpret_cast<uint32_t*>(&hi);
  return packed;
}

// ---- gradient sources ------------------------------------------------------------------------------
// GS_LOCAL : fp32 reduced shard in local memory (index relative to the shard)
// GS_P2P   : `world` mapped full-vector buffers, summed in rank order 0..world-1 (deterministic)
// GS_MC    : one multimem.ld_reduce on the multicast address (the switch sums)
constexpr int GS_LOCAL = 0, GS_P2P = 1, GS_MC = 2, GS_TMA = 3;  // GS_TMA: peer_tma.cu (pointers as GS_P2P)

struct GradSrc {
  PeerPtrs ptrs;        // GS_P2P (and the ragged tail of GS_MC)
  const void* mc;       // GS_MC
  const float* local;   // GS_LOCAL
  int world;
  float divisor;        // world when averaging (DataParallel's mean over the global batch), else 1
};

// Peer gradients are read with volatile (relaxed, system-scope) loads: never through the non-coherent
// path the compiler would otherwise pick for provably read-only data, so what a peer wrote before the
// barrier is what arrives.  128-bit (fp32) / 64-bit (bf16) per thread, a warp covers 512 / 256 contiguous bytes.
template <int GT>
__device__ __forceinline__ float4 ld_peer4(const void* base, int64_t gvec) {
  if constexpr (GT == SFR_F32) {
    float4 r;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(reinterpret_cast<const float4*>(base) + gvec));
    return r;
  } else {
    uint2 raw;
    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];"
                 : "=r"(raw.x), "=r"(raw.y)
                 : "l"(reinterpret_cast<const uint2*>(base) + gvec));
    return widen_bf16x4(raw);
  }
}
template <int GT>
__device__ __forceinline__ float ld_peer1(const void* base, int64_t i) {
  if constexpr (GT == SFR_F32) {
    return *(reinterpret_cast<const volatile float*>(base) + i);
  } else {
    return bf16_bits_to_f32(*(reinterpret_cast<const volatile unsigned short*>(base) + i));
  }
}

template <int GT>
__device__ __forceinline__ float4 p2p_sum4(const GradSrc& s, int64_t gvec) {
  float4 v[kMaxPeers];
  v[0] = ld_peer4<GT>(s.ptrs.p[0], gvec);  // world >= 1
#pragma unroll
  for (int r = 1; r < kMaxPeers; ++r)
    if (r < s.world) v[r] = ld_peer4<GT>(s.ptrs.p[r], gvec);
  float4 a = v[0];
#pragma unroll
  for (int r = 1; r < kMaxPeers; ++r)
    if (r < s.world) {
      a.x = __fadd_rn(a.x, v[r].x);
      a.y = __fadd_rn(a.y, v[r].y);
      a.z = __fadd_rn(a.z, v[r].z);
      a.w = __fadd_rn(a.w, v[r].w);
    }
  return a;
}

// Four reduced (and averaged) gradients: elements 4*vec.. of the shard = 4*(lo_vec+vec).. of the vector.
template <int GT, int GS>
__device__ __forceinline__ float4 reduced_g4(const GradSrc& s, int64_t lo_vec, int64_t vec) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (GS == GS_LOCAL) {
    return *(reinterpret_cast<const float4*>(s.local) + vec);
  } else if constexpr (GS == GS_P2P) {
    a = p2p_sum4<GT>(s, lo_vec + vec);
  } else {
    if constexpr (GT == SFR_F32) a = mc_ld_reduce_f32x4(reinterpret_cast<const float4*>(s.mc) + lo_vec + vec);
    else a = widen_bf16x4(mc_ld_reduce_bf16x4(reinterpret_cast<const uint2*>(s.mc) + lo_vec + vec));
  }
  if (s.divisor != 1.0f) {
    a.x = __fdiv_rn(a.x, s.divisor);
    a.y = __fdiv_rn(a.y, s.divisor);
    a.z = __fdiv_rn(a.z, s.divisor);
    a.w = __fdiv_rn(a.w, s.divisor);
  }
  return a;
}

// One element of the ragged tail (n_total % 4 != 0, last shard only): always through the mapped pointers.
template <int GT, int GS>
__device__ __forceinline__ float reduced_g1(const GradSrc& s, int64_t lo, int64_t i) {
  if constexpr (GS == GS_LOCAL) {
    return s.local[i];
  } else {
    // constant indices only: a dynamically indexed by-value struct would be copied to local memory
    float a = ld_peer1<GT>(s.ptrs.p[0], lo + i);
#pragma unroll
    for (int r = 1; r < kMaxPeers; ++r)
      if (r < s.world) a = __fadd_rn(a, ld_peer1<GT>(s.ptrs.p[r], lo + i));
    return s.divisor != 1.0f ? __fdiv_rn(a, s.divisor) : a;
  }
}

// ---- weight sinks ------------------------------------------------------------------------------------
struct Sink {
  PeerPtrs ptrs;   // full-vector buffers of every rank
  void* mc;        // multicast address or nullptr (then P2P stores)
  int world;
  int skip;        // rank whose buffer the caller's local store already covers (-1: none)
  int on;
};

__device__ __forceinline__ void st_peer_b128(float4* dst, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" : : "l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_peer_b64(uint2* dst, uint2 v) {
  asm volatile("st.global.v2.u32 [%0], {%1,%2};" : : "l"(dst), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ void push_f32x4(const Sink& k, int64_t gvec, float4 v) {
  if (k.mc) {
    mc_st_b128(reinterpret_cast<float4*>(k.mc) + gvec, v);
  } else {
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < k.world && r != k.skip) st_peer_b128(reinterpret_cast<float4*>(k.ptrs.p[r]) + gvec, v);
  }
}
__device__ __forceinline__ void push_bf16x4(const Sink& k, int64_t gvec, uint2 v) {
  if (k.mc) {
    mc_st_b64(reinterpret_cast<uint2*>(k.mc) + gvec, v);
  } else {
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < k.world && r != k.skip) st_peer_b64(reinterpret_cast<uint2*>(k.ptrs.p[r]) + gvec, v);
  }
}
__device__ __forceinline__ void push_f32x1(const Sink& k, int64_t gi, float v) {
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)
    if (r < k.world && (r != k.skip || k.mc)) reinterpret_cast<float*>(k.ptrs.p[r])[gi] = v;
}
__device__ __forceinline__ void push_bf16x1(const Sink& k, int64_t gi, float v) {
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)
    if (r < k.world && (r != k.skip || k.mc)) reinterpret_cast<__nv_bfloat16*>(k.ptrs.p[r])[gi] = __float2bfloat16_rn(v);
}


// ---- entry points of peer_tma.cu (SFR_XP_TMA) ---------------------------------------------------------------
int launch_reduce_tma(int g_dtype, const GradSrc& src, const sfr_peer_geom* q, float* g_red, const uint8_t* mask,
                      double* sumsq, float* fisher, float fisher_div, int max_ctas, cudaStream_t s);
int launch_update_tma(int opt, int ema_mode, int gt, bool from_peers, float* p, const GradSrc& src, float* m, float* v,
                      const uint8_t* mask, float* ema, const Sink& bc32, const Sink& bc16, const sfr_peer_geom* q,
                      const UpdateConsts& c, const DevConsts* c_dev, const double* clip_sumsq, int max_ctas,
                      cudaStream_t s);

}  // namespace sfr
