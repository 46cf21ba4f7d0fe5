// peer_tma.cu — the cross-GPU exchange with the NVLink traffic carried by TMA bulk copies (SFR_XP_TMA).
//
// Why.  In peer.cu the threads that stream the shard's local state (p, m, v, ema, Fisher: HBM) also issue the peer
// loads and stores.  Measured at world = 2 on a DiT-XL/2 vector (profiles/r2_xchg_n2.jsonl): the two kinds of traffic
// do not overlap — a kernel takes (HBM time + NVLink time), not the larger of the two — because an SM's load/store
// pipeline that is backed up behind 770 GB/s of NVLink cannot run ahead on the 6.5 TB/s local streams.
// Here the remote side is taken off that pipeline:
//   pull   one elected thread per CTA issues cp.async.bulk (global -> shared) for a tile of EVERY rank's gradient into a
//          multi-stage shared-memory ring; completion is signalled on an mbarrier (complete_tx::bytes).  The compute
//          threads wait on the barrier, sum the `world` tiles out of shared memory in rank order (same arithmetic and
//          order as the P2P transport: deterministic) and stream the local state with ordinary 128-bit accesses;
//   push   the updated weights of a tile are staged in shared memory (double-buffered) and one elected thread issues a
//          cp.async.bulk (shared -> global) per peer; the CTA moves on to the next tile while the copy engine drains.
// The TMA unit has its own request queues, so NVLink latency and back-pressure no longer stall the local streams.
//
// Tiles: kTileElems consecutive elements of the shard per source (8 KB fp32 / 4 KB bf16); ring depth and CTAs per SM
// follow from the shared memory a stage needs (world tiles).  Bulk copies need 16-byte aligned addresses and sizes:
// the vector part of the shard is cut at a multiple of 16 bytes and the last few elements (< 4 fp32 / < 8 bf16) go
// through the scalar path of peer_common.cuh.
#include "peer_common.cuh"

namespace sfr {
namespace {

constexpr int kTmaThreads = 256;
constexpr int kTileElems = 2048;
constexpr int kTileVec = kTileElems / 4;
constexpr int kMaxStages = 4;
constexpr int kSmemBudget = 200 * 1024;   // of the 227 KB a CTA may opt into

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t phase) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(addr), "r"(phase)
                 : "memory");
}
// global (any mapped address, local or peer) -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global (any mapped address), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// writes of the generic proxy (st.shared) must be ordered before the async proxy (the bulk copy) reads them
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- the gradient ring ---------------------------------------------------------------------------------------
// Stage s holds `world` tiles back to back: tile of rank r at  ring + (s * world + r) * tile_bytes.
template <int GT>
struct GradRing {
  static constexpr int ES = GT == SFR_F32 ? 4 : 2;
  static constexpr uint32_t kTileBytes = kTileElems * ES;
  unsigned char* ring;
  unsigned long long* full;
  int stages, world;
  int64_t lo, nvec_elems;   // elements of the shard covered by bulk copies (a multiple of 16 bytes worth)

  __device__ __forceinline__ int64_t ntiles() const { return (nvec_elems + kTileElems - 1) / kTileElems; }
  __device__ __forceinline__ int tile_elems(int64_t tile) const {
    const int64_t left = nvec_elems - tile * kTileElems;
    return (int)(left < kTileElems ? left : kTileElems);
  }
  // thread 0 only
  __device__ __forceinline__ void issue(const GradSrc& src, int s, int64_t tile) const {
    const uint32_t bytes = (uint32_t)tile_elems(tile) * ES;
    mbar_expect_tx(full + s, bytes * (uint32_t)world);
    const int64_t off = (lo + tile * kTileElems) * ES;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)   // constant indices: the by-value struct stays in the parameter bank
      if (r < world)
        bulk_g2s(ring + ((int64_t)s * world + r) * kTileBytes, static_cast<const char*>(src.ptrs.p[r]) + off, bytes,
                 full + s);
  }
  // four reduced gradients: vector j of the tile in stage s (sum in rank order, then the divisor)
  __device__ __forceinline__ float4 reduced4(const GradSrc& src, int s, int j) const {
    const unsigned char* base = ring + (int64_t)s * world * kTileBytes;
    float4 a;
    if constexpr (GT == SFR_F32) {
      a = *(reinterpret_cast<const float4*>(base) + j);
#pragma unroll
      for (int r = 1; r < kMaxPeers; ++r)
        if (r < world) {
          const float4 x = *(reinterpret_cast<const float4*>(base + (int64_t)r * kTileBytes) + j);
          a.x = __fadd_rn(a.x, x.x);
          a.y = __fadd_rn(a.y, x.y);
          a.z = __fadd_rn(a.z, x.z);
          a.w = __fadd_rn(a.w, x.w);
        }
    } else {
      a = widen_bf16x4(*(reinterpret_cast<const uint2*>(base) + j));
#pragma unroll
      for (int r = 1; r < kMaxPeers; ++r)
        if (r < world) {
          const float4 x = widen_bf16x4(*(reinterpret_cast<const uint2*>(base + (int64_t)r * kTileBytes) + j));
          a.x = __fadd_rn(a.x, x.x);
          a.y = __fadd_rn(a.y, x.y);
          a.z = __fadd_rn(a.z, x.z);
          a.w = __fadd_rn(a.w, x.w);
        }
    }
    if (src.divisor != 1.0f) {
      a.x = __fdiv_rn(a.x, src.divisor);
      a.y = __fdiv_rn(a.y, src.divisor);
      a.z = __fdiv_rn(a.z, src.divisor);
      a.w = __fdiv_rn(a.w, src.divisor);
    }
    return a;
  }
};

template <int GT>
__host__ __device__ constexpr int64_t bulk_elems(int64_t n) {
  return GT == SFR_F32 ? (n & ~int64_t(3)) : (n & ~int64_t(7));   // 16-byte multiples
}

// ================================================================================== reduce (+K1, +norm)
template <int GT>
__global__ void __launch_bounds__(kTmaThreads, 1)
peer_reduce_tma_kernel(GradSrc src, int stages, int64_t lo, int64_t n, float* __restrict__ g_red,
                       const uint8_t* __restrict__ mask, double* __restrict__ sumsq,
                       float* __restrict__ fisher, float fisher_div) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[kMaxStages];
  __shared__ double scratch[32];
  GradRing<GT> ring{smem_raw, full, stages, src.world, lo, bulk_elems<GT>(n)};
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ntiles = ring.ntiles();
  if (threadIdx.x == 0)
    for (int s = 0; s < stages; ++s) {
      const int64_t tile = blockIdx.x + (int64_t)s * gridDim.x;
      if (tile < ntiles) ring.issue(src, s, tile);
    }
  float4* red4 = reinterpret_cast<float4*>(g_red);
  float4* f4 = reinterpret_cast<float4*>(fisher);
  double total = 0.0;
  for (int64_t k = 0;; ++k) {
    const int64_t tile = blockIdx.x + k * gridDim.x;
    if (tile >= ntiles) break;
    const int s = (int)(k % stages);
    const int nv = ring.tile_elems(tile) >> 2;
    const int64_t vec0 = tile * kTileVec;
    // the local streams of this tile are requested BEFORE waiting for the remote tiles
    float4 acc[kTileVec / kTmaThreads];
    uint32_t mk[kTileVec / kTmaThreads];
#pragma unroll
    for (int u = 0; u < kTileVec / kTmaThreads; ++u) {
      const int j = threadIdx.x + u * kTmaThreads;
      acc[u] = (fisher && j < nv) ? ld_stream(f4 + vec0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      mk[u] = (sumsq && mask && j < nv) ? load_mask4(mask, vec0 + j) : 0x01010101u;
    }
    mbar_wait(full + s, (uint32_t)((k / stages) & 1));
    float part = 0.f;
#pragma unroll
    for (int u = 0; u < kTileVec / kTmaThreads; ++u) {
      const int j = threadIdx.x + u * kTmaThreads;
      if (j >= nv) continue;
      const float4 g = ring.reduced4(src, s, j);
      if (g_red) st_stream(red4 + vec0 + j, g);
      if (fisher) {
        acc[u].x = __fadd_rn(acc[u].x, __fdiv_rn(__fmul_rn(g.x, g.x), fisher_div));
        acc[u].y = __fadd_rn(acc[u].y, __fdiv_rn(__fmul_rn(g.y, g.y), fisher_div));
        acc[u].z = __fadd_rn(acc[u].z, __fdiv_rn(__fmul_rn(g.z, g.z), fisher_div));
        acc[u].w = __fadd_rn(acc[u].w, __fdiv_rn(__fmul_rn(g.w, g.w), fisher_div));
        st_stream(f4 + vec0 + j, acc[u]);
      }
      if (sumsq) {
        const float x = __fmul_rn(g.x, mask_byte_to_f32(mk[u], 0));
        const float y = __fmul_rn(g.y, mask_byte_to_f32(mk[u], 1));
        const float z = __fmul_rn(g.z, mask_byte_to_f32(mk[u], 2));
        const float w = __fmul_rn(g.w, mask_byte_to_f32(mk[u], 3));
        part = __fmaf_rn(x, x, part);
        part = __fmaf_rn(y, y, part);
        part = __fmaf_rn(z, z, part);
        part = __fmaf_rn(w, w, part);
      }
    }
    total += (double)part;
    __syncthreads();   // every thread is done with stage s
    if (threadIdx.x == 0) {
      const int64_t next = blockIdx.x + (k + stages) * gridDim.x;
      if (next < ntiles) ring.issue(src, s, next);
    }
  }
  // the last elements that do not fill 16 bytes (< 4 fp32 / < 8 bf16): scalar path through the mapped pointers
  const int64_t tail0 = ring.nvec_elems;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    const float g = reduced_g1<GT, GS_P2P>(src, lo, i);
    if (g_red) g_red[i] = g;
    if (fisher) fisher[i] = __fadd_rn(fisher[i], __fdiv_rn(__fmul_rn(g, g), fisher_div));
    if (sumsq) {
      const float x = mask ? __fmul_rn(g, (float)mask[i]) : g;
      total += (double)x * (double)x;
    }
  }
  if (sumsq) {
    total = block_sum<double>(total, scratch);
    if (threadIdx.x == 0) atomicAdd(sumsq, total);
  }
}

// ================================================================================== K3 + exchange
// Gradient: the local reduced shard (FROM_PEERS = false: plain loads) or the ring above.  Push: staged in shared
// memory, two buffers per sink, one bulk store per peer and tile.
template <int OPT, int EMA, int GT, bool FROM_PEERS>
__global__ void __launch_bounds__(kTmaThreads, 1)
fused_update_tma_kernel(float* __restrict__ p, GradSrc src, int stages, float* __restrict__ m,
                        float* __restrict__ v, const uint8_t* __restrict__ mask, float* __restrict__ ema,
                        Sink bc32, Sink bc16, int64_t lo, int64_t n, UpdateConsts c_arg,
                        const DevConsts* __restrict__ c_dev, const double* __restrict__ clip_sumsq) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[kMaxStages];
  // layout of the dynamic shared memory: [2 x 8 KB fp32 staging][2 x 4 KB bf16 staging][gradient ring]
  float4* stg32 = reinterpret_cast<float4*>(smem_raw);
  uint2* stg16 = reinterpret_cast<uint2*>(smem_raw + 2 * kTileElems * 4);
  unsigned char* ring_base = smem_raw + 2 * kTileElems * 4 + 2 * kTileElems * 2;
  // every stream of the kernel shares ONE cut: whole 16-byte pieces of the NARROWEST stream (the bf16 push):
  // multiples of 8 elements; the < 8 elements left go through the scalar path below
  const int64_t nvec_elems = bulk_elems<SFR_BF16>(n);
  GradRing<GT> ring{ring_base, full, stages, src.world, lo, nvec_elems};

  UpdateConsts c = c_arg;
  float coef_dev = 1.0f;
  if (c_dev != nullptr) {
    apply_dev_consts<OPT>(c, c_dev);
    coef_dev = c_dev->clip_coef;
  }
  constexpr bool kHasV = OPT != SFR_OPT_SGD;
  constexpr bool kHasEma = EMA != SFR_EMA_NONE;
  const bool use_mask = (c.flags & (SFR_F_MASK | SFR_F_MASK_AFTER_CLIP)) != 0;
  const bool has_m = kHasV || c.has_momentum;
  const bool read_m = has_m && !(OPT == SFR_OPT_SGD && (c.flags & SFR_F_SGD_FIRST_STEP));
  const float coef = clip_sumsq ? (c_dev ? coef_dev : clip_coef_warp(clip_sumsq, c.max_norm)) : 1.0f;

  if (FROM_PEERS && threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ntiles = (nvec_elems + kTileElems - 1) / kTileElems;
  if (FROM_PEERS && threadIdx.x == 0)
    for (int s = 0; s < stages; ++s) {
      const int64_t tile = blockIdx.x + (int64_t)s * gridDim.x;
      if (tile < ntiles) ring.issue(src, s, tile);
    }
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  float4* e4 = reinterpret_cast<float4*>(ema);
  const float4* gl4 = reinterpret_cast<const float4*>(src.local);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int U = kTileVec / kTmaThreads;   // 2 vectors per thread and tile

  for (int64_t k = 0;; ++k) {
    const int64_t tile = blockIdx.x + k * gridDim.x;
    if (tile >= ntiles) break;
    const int s = FROM_PEERS ? (int)(k % stages) : 0;
    const int64_t left = nvec_elems - tile * kTileElems;
    const int elems = (int)(left < kTileElems ? left : kTileElems);
    const int nv = elems >> 2;
    const int64_t vec0 = tile * kTileVec;
    const int buf = (int)(k & 1);
    float4 gg[U], pp[U], mm[U], vv[U], ee[U];
    uint32_t mk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = threadIdx.x + u * kTmaThreads;
      const bool in = j < nv;
      if constexpr (!FROM_PEERS) gg[u] = in ? *(gl4 + vec0 + j) : zero4;
      mk[u] = (in && use_mask) ? load_mask4(mask, vec0 + j) : 0x01010101u;
      pp[u] = in ? ld_stream(p4 + vec0 + j) : zero4;
      mm[u] = (in && read_m) ? ld_stream(m4 + vec0 + j) : zero4;
      vv[u] = (in && kHasV) ? ld_stream(v4 + vec0 + j) : zero4;
      ee[u] = (in && kHasEma) ? ld_stream(e4 + vec0 + j) : zero4;
    }
    if constexpr (FROM_PEERS) mbar_wait(full + s, (uint32_t)((k / stages) & 1));
    // the staging buffer of this parity was handed to the copy engine two tiles ago: it must have been read out
    if (threadIdx.x == 0) bulk_wait_read<1>();
    __syncthreads();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = threadIdx.x + u * kTmaThreads;
      if (j >= nv) continue;
      if constexpr (FROM_PEERS) gg[u] = ring.reduced4(src, s, j);
      update_one<OPT, EMA>(pp[u].x, gg[u].x, mm[u].x, vv[u].x, ee[u].x, mask_byte_to_f32(mk[u], 0), coef, c);
      update_one<OPT, EMA>(pp[u].y, gg[u].y, mm[u].y, vv[u].y, ee[u].y, mask_byte_to_f32(mk[u], 1), coef, c);
      update_one<OPT, EMA>(pp[u].z, gg[u].z, mm[u].z, vv[u].z, ee[u].z, mask_byte_to_f32(mk[u], 2), coef, c);
      update_one<OPT, EMA>(pp[u].w, gg[u].w, mm[u].w, vv[u].w, ee[u].w, mask_byte_to_f32(mk[u], 3), coef, c);
      if (bc32.on) stg32[buf * kTileVec + j] = pp[u];
      if (bc16.on) stg16[buf * kTileVec + j] = pack_bf16x4(pp[u]);
      st_stream(p4 + vec0 + j, pp[u]);
      if (has_m) st_stream(m4 + vec0 + j, mm[u]);
      if constexpr (kHasV) st_stream(v4 + vec0 + j, vv[u]);
      if constexpr (kHasEma) st_stream(e4 + vec0 + j, ee[u]);
    }
    fence_proxy_async();
    __syncthreads();   // the tile is staged; the gradient stage has been consumed
    if (threadIdx.x == 0) {
      const int64_t goff = lo + tile * kTileElems;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) {
        if (bc32.on && r < bc32.world && r != bc32.skip)
          bulk_s2g(static_cast<char*>(bc32.ptrs.p[r]) + goff * 4, stg32 + buf * kTileVec, (uint32_t)elems * 4);
        if (bc16.on && r < bc16.world && r != bc16.skip)
          bulk_s2g(static_cast<char*>(bc16.ptrs.p[r]) + goff * 2, stg16 + buf * kTileVec, (uint32_t)elems * 2);
      }
      bulk_commit();
      if constexpr (FROM_PEERS) {
        const int64_t next = blockIdx.x + (k + stages) * gridDim.x;
        if (next < ntiles) ring.issue(src, s, next);
      }
    }
  }
  if (threadIdx.x == 0) bulk_wait_all();   // every push of this CTA has been written before the CTA retires

  const int64_t tail0 = nvec_elems;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    float gg = FROM_PEERS ? reduced_g1<GT, GS_P2P>(src, lo, i) : src.local[i];
    float mk = use_mask ? (float)mask[i] : 1.0f;
    float pp = p[i];
    float mm = read_m ? m[i] : 0.f;
    float vv = kHasV ? v[i] : 0.f;
    float ee = kHasEma ? ema[i] : 0.f;
    update_one<OPT, EMA>(pp, gg, mm, vv, ee, mk, coef, c);
    p[i] = pp;
    if (has_m) m[i] = mm;
    if constexpr (kHasV) v[i] = vv;
    if constexpr (kHasEma) ema[i] = ee;
    if (bc32.on) push_f32x1(bc32, lo + i, pp);
    if (bc16.on) push_bf16x1(bc16, lo + i, pp);
  }
}

// ---- launch geometry ---------------------------------------------------------------------------------------
struct TmaGeom {
  int stages, ctas_per_sm, smem;
};

TmaGeom ring_geometry(int world, int tile_bytes, int fixed_bytes) {
  const int stage = world * tile_bytes;
  // as many CTAs per SM as fit with at least 2 stages each, then the deepest ring that still fits
  int ctas = (kSmemBudget) / (fixed_bytes + 2 * stage);
  ctas = ctas < 1 ? 1 : (ctas > 2 ? 2 : ctas);   // 256 threads x 64-95 registers: two CTAs per SM at most
  int stages = (kSmemBudget / ctas - fixed_bytes) / stage;
  stages = stages < 1 ? 1 : (stages > kMaxStages ? kMaxStages : stages);
  return {stages, ctas, fixed_bytes + stages * stage};
}

template <typename K>
cudaError_t opt_in_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

template <int OPT, int EMA>
cudaError_t launch_update_tma_variant(int gt, bool from_peers, int grid, int smem, cudaStream_t s, float* p,
                                      const GradSrc& src, int stages, float* m, float* v, const uint8_t* mask,
                                      float* ema, const Sink& bc32, const Sink& bc16, int64_t lo, int64_t n,
                                      const UpdateConsts& c, const DevConsts* c_dev, const double* clip_sumsq) {
#define SFR_TMA_LAUNCH(GTV, FP)                                                                                    \
  do {                                                                                                             \
    auto kern = fused_update_tma_kernel<OPT, EMA, GTV, FP>;                                                        \
    cudaError_t e = opt_in_smem(kern, smem);                                                                       \
    if (e != cudaSuccess) return e;                                                                                \
    kern<<<grid, kTmaThreads, smem, s>>>(p, src, stages, m, v, mask, ema, bc32, bc16, lo, n, c, c_dev, clip_sumsq); \
    return cudaGetLastError();                                                                                     \
  } while (0)
  if (!from_peers) SFR_TMA_LAUNCH(SFR_F32, false);
  if (gt == SFR_F32) SFR_TMA_LAUNCH(SFR_F32, true);
  SFR_TMA_LAUNCH(SFR_BF16, true);
#undef SFR_TMA_LAUNCH
}

template <int OPT>
cudaError_t launch_update_tma_ema(int ema_mode, int gt, bool from_peers, int grid, int smem, cudaStream_t s, float* p,
                                  const GradSrc& src, int stages, float* m, float* v, const uint8_t* mask, float* ema,
                                  const Sink& bc32, const Sink& bc16, int64_t lo, int64_t n, const UpdateConsts& c,
                                  const DevConsts* c_dev, const double* clip_sumsq) {
#define SFR_TMA_EMA(E) \
  launch_update_tma_variant<OPT, E>(gt, from_peers, grid, smem, s, p, src, stages, m, v, mask, ema, bc32, bc16, lo, n, c, c_dev, clip_sumsq)
  switch (ema_mode) {
    case SFR_EMA_DDPM: return SFR_TMA_EMA(SFR_EMA_DDPM);
    case SFR_EMA_DIT: return SFR_TMA_EMA(SFR_EMA_DIT);
    case SFR_EMA_SLOWFAST: return SFR_TMA_EMA(SFR_EMA_SLOWFAST);
    default: return SFR_TMA_EMA(SFR_EMA_NONE);
  }
#undef SFR_TMA_EMA
}

}  // namespace

int launch_reduce_tma(int g_dtype, const GradSrc& src, const sfr_peer_geom* q, float* g_red, const uint8_t* mask,
                      double* sumsq, float* fisher, float fisher_div, int max_ctas, cudaStream_t s) {
  const int tile_bytes = kTileElems * (g_dtype == SFR_F32 ? 4 : 2);
  const TmaGeom geo = ring_geometry(q->world, tile_bytes, 0);
  const int64_t ntiles = (q->n_local + kTileElems - 1) / kTileElems;
  int grid = persistent_grid(ntiles, geo.ctas_per_sm);
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  cudaError_t e;
  if (g_dtype == SFR_F32) {
    e = opt_in_smem(peer_reduce_tma_kernel<SFR_F32>, geo.smem);
    if (e == cudaSuccess)
      peer_reduce_tma_kernel<SFR_F32><<<grid, kTmaThreads, geo.smem, s>>>(src, geo.stages, q->lo, q->n_local, g_red,
                                                                         mask, sumsq, fisher, fisher_div);
  } else {
    e = opt_in_smem(peer_reduce_tma_kernel<SFR_BF16>, geo.smem);
    if (e == cudaSuccess)
      peer_reduce_tma_kernel<SFR_BF16><<<grid, kTmaThreads, geo.smem, s>>>(src, geo.stages, q->lo, q->n_local, g_red,
                                                                          mask, sumsq, fisher, fisher_div);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  return e == cudaSuccess ? SFR_OK : (int)e;
}

int launch_update_tma(int opt, int ema_mode, int gt, bool from_peers, float* p, const GradSrc& src, float* m, float* v,
                      const uint8_t* mask, float* ema, const Sink& bc32, const Sink& bc16, const sfr_peer_geom* q,
                      const UpdateConsts& c, const DevConsts* c_dev, const double* clip_sumsq, int max_ctas,
                      cudaStream_t s) {
  const int fixed = 2 * kTileElems * 4 + 2 * kTileElems * 2;   // the two staging rings
  const int tile_bytes = kTileElems * (gt == SFR_F32 ? 4 : 2);
  TmaGeom geo = from_peers ? ring_geometry(q->world, tile_bytes, fixed) : TmaGeom{1, 2, fixed};
  const int64_t ntiles = (q->n_local + kTileElems - 1) / kTileElems;
  int grid = persistent_grid(ntiles, geo.ctas_per_sm);
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  cudaError_t e;
  switch (opt) {
    case SFR_OPT_SGD:
      e = launch_update_tma_ema<SFR_OPT_SGD>(ema_mode, gt, from_peers, grid, geo.smem, s, p, src, geo.stages, m, v, mask, ema, bc32, bc16, q->lo, q->n_local, c, c_dev, clip_sumsq);
      break;
    case SFR_OPT_ADAM:
      e = launch_update_tma_ema<SFR_OPT_ADAM>(ema_mode, gt, from_peers, grid, geo.smem, s, p, src, geo.stages, m, v, mask, ema, bc32, bc16, q->lo, q->n_local, c, c_dev, clip_sumsq);
      break;
    default:
      e = launch_update_tma_ema<SFR_OPT_ADAMW>(ema_mode, gt, from_peers, grid, geo.smem, s, p, src, geo.stages, m, v, mask, ema, bc32, bc16, q->lo, q->n_local, c, c_dev, clip_sumsq);
      break;
  }
  return e == cudaSuccess ? SFR_OK : (int)e;
}

}  // namespace sfr
