// peer.cu — the hot path's cross-GPU exchange steps, FUSED with the kernels they feed, over NVLink peer
// memory (one process per GPU; every rank maps every other rank's symmetric buffers, and — when the
// NVSwitch offers it — one multicast address per buffer).
//
// What it replaces.  The reference runs its loops under `torch.nn.DataParallel` (DiT/forget.py:193,
// DiT/generate_fisher.py:173, DDPM/runners/diffusion.py:110,1060): after every backward pass the per-GPU
// gradients are reduce_add-ed onto GPU 0, which then runs the whole mask / clip / step / EMA sequence
// alone and re-broadcasts the weights before the next forward.  Here the flat vector is SHARDED: rank r
// owns elements [lo, lo + n_local) and
//
//   sfr_peer_reduce         pulls its shard of every rank's gradient over NVLink (P2P loads from the
//                           `world` mapped buffers summed in rank order, or ONE multimem.ld_reduce that
//                           the switch reduces), averages it, and in the same pass feeds K1
//                           (F += g**2 / L), the clip norm (sum (g*mask)**2) and/or leaves the reduced
//                           shard in local memory                — reduce-scatter + K1 + norm, one kernel
//   sfr_peer_fused_update   K3 (mask, clip, optimizer step, EMA) on the shard, reading the gradient
//                           either from that local reduced shard or straight from the peers, and PUSHING
//                           the new weights (fp32 and/or the bf16 working copy) into every rank's
//                           full-vector buffer (P2P stores, or one multimem.st that the switch
//                           replicates)                          — reduce-scatter + K3 + all-gather, one kernel
//   sfr_peer_broadcast      the all-gather alone (a shard updated by another kernel)
//   sfr_peer_barrier        the cross-GPU ordering point, which also carries up to 8 doubles per rank and
//                           returns their sum in rank order (the clip norm's all-reduce rides on it)
//
// NVLink arithmetic (per GPU and direction, s = bytes per element, world = W): a reduce moves
// (W-1)/W * n * s out of every GPU (every other rank needs this GPU's contribution to its shard), a
// broadcast moves the same amount in.  Fused, P2P carries both in both directions ((W-1)/W * n * (s_g+s_p)
// each way); with multimem the switch does the fan-in / fan-out, so the reduce is outbound-heavy
// ((W-1)/W out, 1/W in) and the broadcast inbound-heavy: together ~n * s per direction — 1.75x less at W = 8.
//
// Ordering.  Data moves only between two sfr_peer_barrier calls on the same stream: barrier (every rank's
// backward has written g) -> reduce / fused update kernels -> barrier (every rank has finished reading g
// and its weight stores have landed).  Kernel boundaries flush this GPU's view; the barrier's
// fence.sys + st.release.sys / ld.acquire.sys pair orders it across GPUs.  Peer addresses bypass the local L2
// (B300_MICROARCH.md), and the L1 is invalidated at every kernel start, so plain loads see current data.
// The spin in the barrier is bounded by %globaltimer: on timeout it sets a status word and returns, so a
// dead peer cannot hang the GPU.
#include "peer_common.cuh"

namespace sfr {
namespace {

// ================================================================================== barrier (+ sum)
// Pad layout (u64 words of a symmetric, zero-initialised buffer; every rank has one):
//   [0] epoch of this rank (local)      [1] status (local: 0 ok, 1 timed out)
//   [8 + r]                             flag written by rank r: the epoch it has reached
//   [16 + parity*64 + r*8 + j]          j-th payload double of rank r for epochs of that parity
// A rank can be at most one epoch ahead of a peer that is still inside the barrier, so two payload
// banks (by epoch parity) are enough; flags only grow, so they are never reset.
constexpr int kPadFlag = 8, kPadPayload = 16, kPadWords = 16 + 2 * 64;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" : : "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(32, 1)
peer_barrier_kernel(PeerPtrs pads, int world, int rank, const double* __restrict__ vals,
                    double* __restrict__ sums, int nvals, unsigned long long timeout_ns) {
  const int lane = threadIdx.x;
  unsigned long long* mine = reinterpret_cast<unsigned long long*>(pads.p[rank]);
  unsigned long long e = 0;
  if (lane == 0) {
    e = mine[0] + 1;
    mine[0] = e;
  }
  e = __shfl_sync(kFullMask, e, 0);
  const int bank = (int)(e & 1ull) * 64;
  if (lane < world) {
    unsigned long long* theirs = reinterpret_cast<unsigned long long*>(pads.p[lane]);
    for (int j = 0; j < nvals; ++j)
      theirs[kPadPayload + bank + rank * 8 + j] = (unsigned long long)__double_as_longlong(vals[j]);
    __threadfence_system();  // everything this GPU wrote before the barrier (and the payload) is ordered first
    st_release_sys(theirs + kPadFlag + rank, e);
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(mine + kPadFlag + lane) < e) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        mine[1] = 1ull;
        break;
      }
    }
  }
  __syncwarp();
  if (lane < nvals) {
    double s = 0.0;
    for (int r = 0; r < world; ++r)
      s += __longlong_as_double((long long)ld_acquire_sys(mine + kPadPayload + bank + r * 8 + lane));
    sums[lane] = s;
  }
}

// ================================================================================== reduce (+K1, +norm)
// U vectors per thread, all peer loads issued before the first use (see the note on bytes in flight below).
constexpr int kRedThreads = 128;
constexpr int kRedCtasPerSm = 4;

template <int GT, int GS, int U>
__global__ void __launch_bounds__(kRedThreads, kRedCtasPerSm)
peer_reduce_kernel(GradSrc src, int64_t lo, int64_t n, float* __restrict__ g_red,
                   const uint8_t* __restrict__ mask, double* __restrict__ sumsq,
                   float* __restrict__ fisher, float fisher_div) {
  __shared__ double scratch[32];
  const int64_t nvec = n >> 2;
  const int64_t lo_vec = lo >> 2;
  const int64_t tile = (int64_t)kRedThreads * U;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  float4* red4 = reinterpret_cast<float4*>(g_red);
  float4* f4 = reinterpret_cast<float4*>(fisher);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  double total = 0.0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 g[U], acc[U];
    uint32_t mk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t vec = base + (int64_t)u * kRedThreads;
      const bool in = vec < nvec;
      g[u] = in ? reduced_g4<GT, GS>(src, lo_vec, vec) : zero4;
      acc[u] = (in && fisher) ? ld_stream(f4 + vec) : zero4;
      mk[u] = (in && sumsq && mask) ? load_mask4(mask, vec) : 0x01010101u;
    }
    float part = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t vec = base + (int64_t)u * kRedThreads;
      if (vec >= nvec) continue;
      if (g_red) st_stream(red4 + vec, g[u]);
      if (fisher) {
        // F += g**2 / L : mul, true divide, add (fisher.cu)
        acc[u].x = __fadd_rn(acc[u].x, __fdiv_rn(__fmul_rn(g[u].x, g[u].x), fisher_div));
        acc[u].y = __fadd_rn(acc[u].y, __fdiv_rn(__fmul_rn(g[u].y, g[u].y), fisher_div));
        acc[u].z = __fadd_rn(acc[u].z, __fdiv_rn(__fmul_rn(g[u].z, g[u].z), fisher_div));
        acc[u].w = __fadd_rn(acc[u].w, __fdiv_rn(__fmul_rn(g[u].w, g[u].w), fisher_div));
        st_stream(f4 + vec, acc[u]);
      }
      if (sumsq) {
        const float x = __fmul_rn(g[u].x, mask_byte_to_f32(mk[u], 0));
        const float y = __fmul_rn(g[u].y, mask_byte_to_f32(mk[u], 1));
        const float z = __fmul_rn(g[u].z, mask_byte_to_f32(mk[u], 2));
        const float w = __fmul_rn(g[u].w, mask_byte_to_f32(mk[u], 3));
        part = __fmaf_rn(x, x, part);
        part = __fmaf_rn(y, y, part);
        part = __fmaf_rn(z, z, part);
        part = __fmaf_rn(w, w, part);
      }
    }
    total += (double)part;  // fp32 partial over <= 16 elements, folded into a double per thread
  }
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    const float g = reduced_g1<GT, GS>(src, lo, i);
    if (g_red) g_red[i] = g;
    if (fisher) fisher[i] = __fadd_rn(fisher[i], __fdiv_rn(__fmul_rn(g, g), fisher_div));
    if (sumsq) {
      const float x = mask ? __fmul_rn(g, (float)mask[i]) : g;
      total += (double)x * (double)x;
    }
  }
  if (sumsq) {  // uniform across the grid
    total = block_sum<double>(total, scratch);
    if (threadIdx.x == 0) atomicAdd(sumsq, total);
  }
}

// ================================================================================== K3 + exchange
// A remote store (or a multimem.st) occupies its CTA until the fabric has taken it, and a peer load comes
// back after ~2 us: what keeps NVLink busy is BYTES IN FLIGHT PER SM, so every thread works on kXUnroll
// vectors at once — all loads first, then the arithmetic, then local stores and pushes (measured at
// world = 2, n = 675 M: one vector per thread pushed at 360 GB/s, latency-bound on CTA turnover).
constexpr int kXThreads = 128;
constexpr int kXCtasPerSm = 4;

template <int OPT, int EMA, int GT, int GS, int U>
__global__ void __launch_bounds__(kXThreads, kXCtasPerSm)
fused_update_xchg_kernel(float* __restrict__ p, GradSrc src, float* __restrict__ m,
                         float* __restrict__ v, const uint8_t* __restrict__ mask,
                         float* __restrict__ ema, Sink bc32, Sink bc16, int64_t lo, int64_t n,
                         UpdateConsts c_arg, const DevConsts* __restrict__ c_dev,
                         const double* __restrict__ clip_sumsq) {
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

  const int64_t nvec = n >> 2;
  const int64_t lo_vec = lo >> 2;
  const int64_t tile = (int64_t)kXThreads * U;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  float4* e4 = reinterpret_cast<float4*>(ema);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 gg[U], pp[U], mm[U], vv[U], ee[U];
    uint32_t mk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t vec = base + (int64_t)u * kXThreads;
      const bool in = vec < nvec;
      gg[u] = in ? reduced_g4<GT, GS>(src, lo_vec, vec) : zero4;
      mk[u] = (in && use_mask) ? load_mask4(mask, vec) : 0x01010101u;
      pp[u] = in ? ld_stream(p4 + vec) : zero4;
      mm[u] = (in && read_m) ? ld_stream(m4 + vec) : zero4;
      vv[u] = (in && kHasV) ? ld_stream(v4 + vec) : zero4;
      ee[u] = (in && kHasEma) ? ld_stream(e4 + vec) : zero4;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      update_one<OPT, EMA>(pp[u].x, gg[u].x, mm[u].x, vv[u].x, ee[u].x, mask_byte_to_f32(mk[u], 0), coef, c);
      update_one<OPT, EMA>(pp[u].y, gg[u].y, mm[u].y, vv[u].y, ee[u].y, mask_byte_to_f32(mk[u], 1), coef, c);
      update_one<OPT, EMA>(pp[u].z, gg[u].z, mm[u].z, vv[u].z, ee[u].z, mask_byte_to_f32(mk[u], 2), coef, c);
      update_one<OPT, EMA>(pp[u].w, gg[u].w, mm[u].w, vv[u].w, ee[u].w, mask_byte_to_f32(mk[u], 3), coef, c);
    }
    // pushes first: they are the long-latency side
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t vec = base + (int64_t)u * kXThreads;
      if (vec < nvec) {
        if (bc32.on) push_f32x4(bc32, lo_vec + vec, pp[u]);
        if (bc16.on) push_bf16x4(bc16, lo_vec + vec, pack_bf16x4(pp[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t vec = base + (int64_t)u * kXThreads;
      if (vec < nvec) {
        st_stream(p4 + vec, pp[u]);
        if (has_m) st_stream(m4 + vec, mm[u]);
        if constexpr (kHasV) st_stream(v4 + vec, vv[u]);
        if constexpr (kHasEma) st_stream(e4 + vec, ee[u]);
      }
    }
  }

  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    float gg = reduced_g1<GT, GS>(src, lo, i);
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

// ================================================================================== all-gather alone
template <int BYTES>  // element size of the buffer: 4 (fp32) or 2 (bf16)
__global__ void __launch_bounds__(256, 4)
peer_broadcast_kernel(const void* __restrict__ src, Sink dst, int64_t lo, int64_t n) {
  const int64_t nvec = n >> 2;
  const int64_t lo_vec = lo >> 2;
  for (int64_t vec = (int64_t)blockIdx.x * 256 + threadIdx.x; vec < nvec; vec += (int64_t)gridDim.x * 256) {
    if constexpr (BYTES == 4) push_f32x4(dst, lo_vec + vec, *(reinterpret_cast<const float4*>(src) + vec));
    else push_bf16x4(dst, lo_vec + vec, *(reinterpret_cast<const uint2*>(src) + vec));
  }
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) {
      if (r >= dst.world || (r == dst.skip && !dst.mc)) continue;
      if constexpr (BYTES == 4) reinterpret_cast<float*>(dst.ptrs.p[r])[lo + i] = reinterpret_cast<const float*>(src)[i];
      else reinterpret_cast<unsigned short*>(dst.ptrs.p[r])[lo + i] = reinterpret_cast<const unsigned short*>(src)[i];
    }
  }
}

// ---- host helpers ----------------------------------------------------------------------------------
bool geom_ok(const sfr_peer_geom* q) {
  return q && q->world >= 1 && q->world <= kMaxPeers && q->rank >= 0 && q->rank < q->world && q->lo >= 0 &&
         q->n_local >= 0 && (q->lo & 15) == 0;
}

int fill_ptrs(PeerPtrs& out, const sfr_peer_buf* b, int world) {
  for (int r = 0; r < kMaxPeers; ++r) out.p[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    if (b->ptrs[r] == nullptr) return SFR_ERR_NULL;
    if (!aligned16(b->ptrs[r])) return SFR_ERR_ALIGN;
    out.p[r] = b->ptrs[r];
  }
  if (b->multicast && !aligned16(b->multicast)) return SFR_ERR_ALIGN;
  return SFR_OK;
}

int make_src(GradSrc& s, const float* g_local, const sfr_peer_buf* g, const sfr_peer_geom* q, int transport,
             int average, int* gs) {
  s = GradSrc{};
  s.world = q->world;
  s.divisor = 1.0f;
  if (g == nullptr) {
    if (g_local == nullptr) return SFR_ERR_NULL;
    if (!aligned16(g_local)) return SFR_ERR_ALIGN;
    s.local = g_local;
    *gs = GS_LOCAL;
    return SFR_OK;
  }
  const int rc = fill_ptrs(s.ptrs, g, q->world);
  if (rc != SFR_OK) return rc;
  if (transport == SFR_XP_MULTIMEM) {
    if (g->multicast == nullptr) return SFR_ERR_NULL;
    s.mc = g->multicast;
    *gs = GS_MC;
  } else if (transport == SFR_XP_P2P) {
    *gs = GS_P2P;
  } else if (transport == SFR_XP_TMA) {
    *gs = GS_TMA;
  } else {
    return SFR_ERR_ARG;
  }
  if (average) s.divisor = (float)q->world;
  return SFR_OK;
}

int make_sink(Sink& k, const sfr_peer_buf* b, const sfr_peer_geom* q, int transport, const void* local_alias,
              int elem_bytes) {
  k = Sink{};
  if (b == nullptr) return SFR_OK;
  const int rc = fill_ptrs(k.ptrs, b, q->world);
  if (rc != SFR_OK) return rc;
  k.on = 1;
  k.world = q->world;
  k.skip = -1;
  if (transport == SFR_XP_MULTIMEM) {
    if (b->multicast == nullptr) return SFR_ERR_NULL;
    k.mc = b->multicast;
  } else if (transport != SFR_XP_P2P && transport != SFR_XP_TMA) {
    return SFR_ERR_ARG;
  }
  // the caller's own store already covers this rank's copy when its local pointer IS the shard of the buffer
  if (local_alias == static_cast<const char*>(b->ptrs[q->rank]) + q->lo * elem_bytes) k.skip = q->rank;
  return SFR_OK;
}

template <int GT, int GS>
void launch_reduce(cudaStream_t s, const GradSrc& src, const sfr_peer_geom* q, float* g_red,
                   const uint8_t* mask, double* sumsq, float* fisher, float fdiv) {
  // P2P holds `world` vectors per unrolled step in registers: unroll 4 up to four ranks, 2 beyond
  const int64_t nvec = q->n_local >> 2;
  auto grid_of = [nvec](int u) {
    return persistent_grid((nvec + (int64_t)kRedThreads * u - 1) / ((int64_t)kRedThreads * u), kRedCtasPerSm * 32);
  };
  if (GS == GS_P2P && q->world > 4)
    peer_reduce_kernel<GT, GS, 2><<<grid_of(2), kRedThreads, 0, s>>>(src, q->lo, q->n_local, g_red, mask, sumsq, fisher, fdiv);
  else
    peer_reduce_kernel<GT, GS, 4><<<grid_of(4), kRedThreads, 0, s>>>(src, q->lo, q->n_local, g_red, mask, sumsq, fisher, fdiv);
}

#define SFR_X_ARGS p, src, m, v, mask, ema, bc32, bc16, q->lo, q->n_local, c, c_dev, clip_sumsq
constexpr int kXUnrollLocal = 4;  // gradient from the local reduced shard: registers are cheap
constexpr int kXUnrollPeer = 2;   // gradient summed from up to 8 peers per vector
template <int OPT, int EMA>
void launch_xchg_src(int gt, int gs, int64_t nvec, cudaStream_t s, float* p, const GradSrc& src, float* m, float* v,
                     const uint8_t* mask, float* ema, const Sink& bc32, const Sink& bc16,
                     const sfr_peer_geom* q, const UpdateConsts& c, const DevConsts* c_dev,
                     const double* clip_sumsq) {
  auto grid_of = [nvec](int unroll) { return full_grid((nvec + (int64_t)kXThreads * unroll - 1) / ((int64_t)kXThreads * unroll)); };
  if (gs == GS_LOCAL) {
    fused_update_xchg_kernel<OPT, EMA, SFR_F32, GS_LOCAL, kXUnrollLocal><<<grid_of(kXUnrollLocal), kXThreads, 0, s>>>(SFR_X_ARGS);
  } else if (gs == GS_P2P) {
    if (gt == SFR_F32) fused_update_xchg_kernel<OPT, EMA, SFR_F32, GS_P2P, kXUnrollPeer><<<grid_of(kXUnrollPeer), kXThreads, 0, s>>>(SFR_X_ARGS);
    else fused_update_xchg_kernel<OPT, EMA, SFR_BF16, GS_P2P, kXUnrollPeer><<<grid_of(kXUnrollPeer), kXThreads, 0, s>>>(SFR_X_ARGS);
  } else {
    if (gt == SFR_F32) fused_update_xchg_kernel<OPT, EMA, SFR_F32, GS_MC, kXUnrollPeer><<<grid_of(kXUnrollPeer), kXThreads, 0, s>>>(SFR_X_ARGS);
    else fused_update_xchg_kernel<OPT, EMA, SFR_BF16, GS_MC, kXUnrollPeer><<<grid_of(kXUnrollPeer), kXThreads, 0, s>>>(SFR_X_ARGS);
  }
}
template <int OPT>
void launch_xchg_ema(int ema_mode, int gt, int gs, int64_t nvec, cudaStream_t s, float* p, const GradSrc& src,
                     float* m, float* v, const uint8_t* mask, float* ema, const Sink& bc32, const Sink& bc16,
                     const sfr_peer_geom* q, const UpdateConsts& c, const DevConsts* c_dev,
                     const double* clip_sumsq) {
#define SFR_X_CALL(E) launch_xchg_src<OPT, E>(gt, gs, nvec, s, p, src, m, v, mask, ema, bc32, bc16, q, c, c_dev, clip_sumsq)
  switch (ema_mode) {
    case SFR_EMA_DDPM: SFR_X_CALL(SFR_EMA_DDPM); break;
    case SFR_EMA_DIT: SFR_X_CALL(SFR_EMA_DIT); break;
    case SFR_EMA_SLOWFAST: SFR_X_CALL(SFR_EMA_SLOWFAST); break;
    default: SFR_X_CALL(SFR_EMA_NONE); break;
  }
#undef SFR_X_CALL
}
#undef SFR_X_ARGS

}  // namespace
}  // namespace sfr

extern "C" int64_t sfr_peer_pad_bytes(void) { return (int64_t)sfr::kPadWords * 8; }

extern "C" int sfr_peer_barrier(const sfr_peer_buf* pad, int world, int rank, const double* vals_dev,
                                double* sums_dev, int nvals, uint64_t timeout_ns, sfr_stream_t stream) {
  using namespace sfr;
  SFR_REQUIRE_PTR(pad);
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return SFR_ERR_ARG;
  if (nvals < 0 || nvals > 8) return SFR_ERR_ARG;
  if (nvals > 0) {
    SFR_REQUIRE_PTR(vals_dev);
    SFR_REQUIRE_PTR(sums_dev);
  }
  PeerPtrs pads;
  const int rc = fill_ptrs(pads, pad, world);
  if (rc != SFR_OK) return rc;
  SFR_ENTER_DEVICE(pad->ptrs[rank]);
  if (timeout_ns == 0) timeout_ns = 30ull * 1000000000ull;
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pads, world, rank, vals_dev, sums_dev,
                                                                      nvals, (unsigned long long)timeout_ns);
  SFR_LAUNCH_STATUS();
}

extern "C" int sfr_peer_reduce(const sfr_peer_buf* g, int g_dtype, const sfr_peer_geom* geom, int transport,
                               int average, float* g_red, const uint8_t* mask, double* sumsq,
                               float* fisher_acc, float fisher_divisor, int max_ctas, sfr_stream_t stream) {
  using namespace sfr;
  SFR_REQUIRE_PTR(g);
  if (!geom_ok(geom)) return SFR_ERR_ARG;
  if (g_dtype != SFR_F32 && g_dtype != SFR_BF16) return SFR_ERR_ARG;
  if (geom->n_local == 0) return SFR_OK;
  if (g_red == nullptr && sumsq == nullptr && fisher_acc == nullptr) return SFR_ERR_NULL;
  SFR_REQUIRE_ALIGNED(g_red);
  SFR_REQUIRE_ALIGNED(mask);
  SFR_REQUIRE_ALIGNED(fisher_acc);
  GradSrc src;
  int gs = GS_P2P;
  const int rc = make_src(src, nullptr, g, geom, transport, average, &gs);
  if (rc != SFR_OK) return rc;
  SFR_ENTER_DEVICE(g->ptrs[geom->rank]);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (max_ctas < 0) return SFR_ERR_ARG;
  if (gs == GS_TMA)
    return launch_reduce_tma(g_dtype, src, geom, g_red, mask, sumsq, fisher_acc, fisher_divisor, max_ctas, s);
  if (gs == GS_P2P) {
    if (g_dtype == SFR_F32) launch_reduce<SFR_F32, GS_P2P>(s, src, geom, g_red, mask, sumsq, fisher_acc, fisher_divisor);
    else launch_reduce<SFR_BF16, GS_P2P>(s, src, geom, g_red, mask, sumsq, fisher_acc, fisher_divisor);
  } else {
    if (g_dtype == SFR_F32) launch_reduce<SFR_F32, GS_MC>(s, src, geom, g_red, mask, sumsq, fisher_acc, fisher_divisor);
    else launch_reduce<SFR_BF16, GS_MC>(s, src, geom, g_red, mask, sumsq, fisher_acc, fisher_divisor);
  }
  SFR_LAUNCH_STATUS();
}

extern "C" int sfr_peer_fused_update(float* p, const float* g_red, const sfr_peer_buf* g, int g_dtype,
                                     int g_transport, int average, float* m, float* v,
                                     const uint8_t* mask, float* ema, const sfr_peer_buf* bc_f32,
                                     const sfr_peer_buf* bc_bf16, int bc_transport,
                                     const sfr_peer_geom* geom, const sfr_update_args* a,
                                     const double* clip_sumsq, long long* step_counter,
                                     void* consts_scratch, int max_ctas, sfr_stream_t stream) {
  using namespace sfr;
  SFR_REQUIRE_PTR(a);
  if (!geom_ok(geom) || max_ctas < 0) return SFR_ERR_ARG;
  if (a->opt < SFR_OPT_SGD || a->opt > SFR_OPT_ADAMW) return SFR_ERR_ARG;
  if (a->ema_mode < SFR_EMA_NONE || a->ema_mode > SFR_EMA_SLOWFAST) return SFR_ERR_ARG;
  if (g != nullptr && g_dtype != SFR_F32 && g_dtype != SFR_BF16) return SFR_ERR_ARG;
  // the gradient belongs to the peers while they read it, and the working copy travels through bc_bf16
  const uint32_t known = SFR_F_MASK | SFR_F_MASK_AFTER_CLIP | SFR_F_SGD_FIRST_STEP | SFR_F_REUSE_CONSTS;
  if (a->flags & ~known) return SFR_ERR_ARG;
  if ((a->flags & SFR_F_MASK) && (a->flags & SFR_F_MASK_AFTER_CLIP)) return SFR_ERR_ARG;
  if (a->opt != SFR_OPT_SGD && a->step < 1 && step_counter == nullptr) return SFR_ERR_ARG;
  if (geom->n_local == 0) return SFR_OK;
  const bool use_mask = (a->flags & (SFR_F_MASK | SFR_F_MASK_AFTER_CLIP)) != 0;
  const bool has_momentum = a->opt == SFR_OPT_SGD && a->momentum != 0.0;
  SFR_REQUIRE_PTR(p);
  if (a->opt != SFR_OPT_SGD || has_momentum) SFR_REQUIRE_PTR(m);
  if (a->opt != SFR_OPT_SGD) SFR_REQUIRE_PTR(v);
  if (use_mask) SFR_REQUIRE_PTR(mask);
  if (a->ema_mode != SFR_EMA_NONE) SFR_REQUIRE_PTR(ema);
  SFR_REQUIRE_ALIGNED(p);
  SFR_REQUIRE_ALIGNED(m);
  SFR_REQUIRE_ALIGNED(v);
  SFR_REQUIRE_ALIGNED(mask);
  SFR_REQUIRE_ALIGNED(ema);
  GradSrc src;
  int gs = GS_LOCAL;
  int rc = make_src(src, g_red, g, geom, g_transport, average, &gs);
  if (rc != SFR_OK) return rc;
  Sink bc32, bc16;
  rc = make_sink(bc32, bc_f32, geom, bc_transport, p, 4);
  if (rc != SFR_OK) return rc;
  rc = make_sink(bc16, bc_bf16, geom, bc_transport, nullptr, 2);
  if (rc != SFR_OK) return rc;
  SFR_ENTER_DEVICE(p);

  UpdateConsts c = make_update_consts(*a, a->step, has_momentum);
  const DevConsts* c_dev = nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((a->lr_table_dev == nullptr) != (a->lr_index_dev == nullptr)) return SFR_ERR_NULL;
  if (step_counter != nullptr || a->lr_table_dev != nullptr || (a->flags & SFR_F_REUSE_CONSTS))
    SFR_REQUIRE_PTR(consts_scratch);
  if (consts_scratch != nullptr) {
    if (!aligned16(consts_scratch)) return SFR_ERR_ALIGN;
    if (!(a->flags & SFR_F_REUSE_CONSTS))
      launch_update_consts(*a, has_momentum, step_counter, clip_sumsq, consts_scratch, s);
    c_dev = reinterpret_cast<const DevConsts*>(consts_scratch);
  }
  const int64_t nvec = geom->n_local >> 2;
  const int gt = g ? g_dtype : SFR_F32;
  const bool pushes = bc32.on || bc16.on;
  if (gs == GS_TMA || (pushes && bc_transport == SFR_XP_TMA)) {
    // one kernel moves both directions through the copy engine: no mixing with the load/store transports
    if ((gs != GS_TMA && gs != GS_LOCAL) || (pushes && bc_transport != SFR_XP_TMA)) return SFR_ERR_ARG;
    return launch_update_tma(a->opt, a->ema_mode, gt, gs == GS_TMA, p, src, m, v, mask, ema, bc32, bc16, geom, c, c_dev,
                             clip_sumsq, max_ctas, s);
  }
  switch (a->opt) {
    case SFR_OPT_SGD:
      launch_xchg_ema<SFR_OPT_SGD>(a->ema_mode, gt, gs, nvec, s, p, src, m, v, mask, ema, bc32, bc16, geom, c, c_dev, clip_sumsq);
      break;
    case SFR_OPT_ADAM:
      launch_xchg_ema<SFR_OPT_ADAM>(a->ema_mode, gt, gs, nvec, s, p, src, m, v, mask, ema, bc32, bc16, geom, c, c_dev, clip_sumsq);
      break;
    default:
      launch_xchg_ema<SFR_OPT_ADAMW>(a->ema_mode, gt, gs, nvec, s, p, src, m, v, mask, ema, bc32, bc16, geom, c, c_dev, clip_sumsq);
      break;
  }
  SFR_LAUNCH_STATUS();
}

extern "C" int sfr_peer_broadcast(const void* src_local, const sfr_peer_buf* dst, int elem_bytes,
                                  const sfr_peer_geom* geom, int transport, sfr_stream_t stream) {
  using namespace sfr;
  SFR_REQUIRE_PTR(dst);
  if (!geom_ok(geom)) return SFR_ERR_ARG;
  if (elem_bytes != 2 && elem_bytes != 4) return SFR_ERR_ARG;
  if (geom->n_local == 0) return SFR_OK;
  SFR_REQUIRE_PTR(src_local);
  SFR_REQUIRE_ALIGNED(src_local);
  Sink k;
  const int rc = make_sink(k, dst, geom, transport, src_local, elem_bytes);
  if (rc != SFR_OK) return rc;
  SFR_ENTER_DEVICE(src_local);
  const int64_t nvec = geom->n_local >> 2;
  const int grid = persistent_grid((nvec + 255) / 256, 4 * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 4) peer_broadcast_kernel<4><<<grid, 256, 0, s>>>(src_local, k, geom->lo, geom->n_local);
  else peer_broadcast_kernel<2><<<grid, 256, 0, s>>>(src_local, k, geom->lo, geom->n_local);
  SFR_LAUNCH_STATUS();
}
