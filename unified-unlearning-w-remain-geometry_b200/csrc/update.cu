// update.cu — clip norm (masked sum of squares) and K3, the fused saliency-masked
// fast/slow update of the SFR-on forget loops.
//
// Reference op sequence per forget iteration (one Python loop over named_parameters each):
//   param.grad *= mask[name].to(device)      sfron.py:201-204, runners/diffusion.py:1126-1129,
//                                            DiT/forget.py:289-292, SD gradient_ascent.py:94-99
//   clip_grad_norm_(params, max_norm)        sfron.py:205, runners/diffusion.py:1131-1136,
//                                            DiT/forget.py:293-298
//   optimizer.step()                         SGD sfron.py:167-174,206,222 ; Adam
//                                            runners/diffusion.py:1062,1138,1176 ; AdamW
//                                            DiT/forget.py:199,299,320 ; Adam SD nsfw_removal.py:81,162,173
//   EMA / slow weights                       DDPM/models/ema.py:17-24 ; DiT/forget.py:52-62 ;
//                                            sfron.py:30-37,255-257
// Here: one reduction pass (g, mask -> sum of squares) and ONE update pass that reads
// g, mask, p, m, v (, ema) once and writes p, m, v (, ema, zeroed g, bf16 p) once.
//
// Arithmetic contract (torch 2.11 CPU kernels, probed op by op):
//   add(b, alpha)   -> fma(b, alpha, a)          lerp(w<.5) -> fma(end-start, w, start)
//   addcmul(value)  -> fma(value*t1, t2, self)   addcdiv    -> self + (value*t1)/t2   (no fma)
//   mul / div / sqrt / add with a Python scalar -> the scalar rounded to fp32 first.
#include <atomic>

#include <cooperative_groups.h>

#include "update_core.cuh"

namespace sfr {
namespace {

namespace cg = cooperative_groups;

constexpr int kThreads = 256;

// ------------------------------------------------------------------ masked sum of squares
constexpr int kSumsqCtasPerSm = 8;
constexpr int kSumsqUnroll = 4;
constexpr int kSumsqGridWaves = 16;  // SMs x 8 x 16 CTAs, one double atomic each (tuned: tools/tune)

// Sum over the grid of (g*mask)^2, accumulated into *out with one double atomic per CTA.  TPB = threads per CTA.
template <int GT, bool MASK, int TPB>
__device__ __forceinline__ void masked_sumsq_body(const void* __restrict__ g, const uint8_t* __restrict__ mask,
                                                  int64_t n, double* __restrict__ out) {
  constexpr int kThreads = TPB;
  __shared__ double scratch[32];
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kThreads * kSumsqUnroll;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  double total = 0.0;

  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 gg[kSumsqUnroll];
    uint32_t mm[kSumsqUnroll];
#pragma unroll
    for (int u = 0; u < kSumsqUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kThreads;
      const bool in = v < nvec;
      gg[u] = in ? load_g4_once<GT>(g, v) : make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (MASK) mm[u] = in ? load_mask4_once(mask, v) : 0u;
    }
    // fp32 partial over the 32 elements of this tile step, folded into a double per thread:
    // the rounding error of the norm stays ~1e-8 relative for any n.
    float part = 0.f;
#pragma unroll
    for (int u = 0; u < kSumsqUnroll; ++u) {
      float x = gg[u].x, y = gg[u].y, z = gg[u].z, w = gg[u].w;
      if constexpr (MASK) {
        x = __fmul_rn(x, mask_byte_to_f32(mm[u], 0));
        y = __fmul_rn(y, mask_byte_to_f32(mm[u], 1));
        z = __fmul_rn(z, mask_byte_to_f32(mm[u], 2));
        w = __fmul_rn(w, mask_byte_to_f32(mm[u], 3));
      }
      part = __fmaf_rn(x, x, part);
      part = __fmaf_rn(y, y, part);
      part = __fmaf_rn(z, z, part);
      part = __fmaf_rn(w, w, part);
    }
    total += (double)part;
  }

  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    float x = load_g1<GT>(g, i);
    if constexpr (MASK) x = __fmul_rn(x, (float)mask[i]);
    total += (double)x * (double)x;
  }

  total = block_sum<double>(total, scratch);
  if (threadIdx.x == 0) atomicAdd(out, total);
}

template <int GT, bool MASK>
__global__ void __launch_bounds__(kThreads, kSumsqCtasPerSm)
masked_sumsq_kernel(const void* __restrict__ g, const uint8_t* __restrict__ mask, int64_t n,
                    double* __restrict__ out) {
  masked_sumsq_body<GT, MASK, kThreads>(g, mask, n, out);
}

constexpr int kUpdThreads = 128;  // small CTAs, one 128-vector tile per CTA (tuned: tools/tune/tune_stream.cu)
constexpr int kUpdCtasPerSm = 6;

template <int OPT, int EMA, int GT>
__device__ __forceinline__ void
fused_update_body(float* __restrict__ p, void* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, const uint8_t* __restrict__ mask,
                  float* __restrict__ ema, void* __restrict__ p_bf16, int64_t n,
                  const UpdateConsts& c_arg, const DevConsts* __restrict__ c_dev,
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
  // coef == 1.0f exactly when not clipping: g * 1.0f is an exact no-op
  const float coef = clip_sumsq ? (c_dev ? coef_dev : clip_coef_warp(clip_sumsq, c.max_norm)) : 1.0f;

  const int64_t nvec = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  float4* e4 = reinterpret_cast<float4*>(ema);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int64_t vec = (int64_t)blockIdx.x * kUpdThreads + threadIdx.x; vec < nvec;
       vec += (int64_t)gridDim.x * kUpdThreads) {
    // issue every load of this step before the first use: 4-6 independent 128-bit
    // requests in flight per thread
    const float4 gg = load_g4<GT>(g, vec);
    const uint32_t mk = use_mask ? load_mask4(mask, vec) : 0x01010101u;
    float4 pp = ld_stream(p4 + vec);
    float4 mm = read_m ? ld_stream(m4 + vec) : zero4;
    float4 vv = kHasV ? ld_stream(v4 + vec) : zero4;
    float4 ee = kHasEma ? ld_stream(e4 + vec) : zero4;
    update_one<OPT, EMA>(pp.x, gg.x, mm.x, vv.x, ee.x, mask_byte_to_f32(mk, 0), coef, c);
    update_one<OPT, EMA>(pp.y, gg.y, mm.y, vv.y, ee.y, mask_byte_to_f32(mk, 1), coef, c);
    update_one<OPT, EMA>(pp.z, gg.z, mm.z, vv.z, ee.z, mask_byte_to_f32(mk, 2), coef, c);
    update_one<OPT, EMA>(pp.w, gg.w, mm.w, vv.w, ee.w, mask_byte_to_f32(mk, 3), coef, c);

    st_stream(p4 + vec, pp);
    if (has_m) st_stream(m4 + vec, mm);
    if constexpr (kHasV) st_stream(v4 + vec, vv);
    if constexpr (kHasEma) st_stream(e4 + vec, ee);
    if (c.flags & SFR_F_ZERO_GRAD) zero_g4<GT>(g, vec);
    if (c.flags & SFR_F_WRITE_BF16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y);
      __nv_bfloat162 hi = __floats2bfloat162_rn(pp.z, pp.w);
      uint2 packed;
      packed.x = *reinterpret_cast<uint32_t*>(&lo);
      packed.y = *reinterpret_cast<uint32_t*>(&hi);
      *(reinterpret_cast<uint2*>(p_bf16) + vec) = packed;
    }
  }

  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    float gg = load_g1<GT>(g, i);
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
    if (c.flags & SFR_F_ZERO_GRAD) zero_g1<GT>(g, i);
    if (c.flags & SFR_F_WRITE_BF16)
      reinterpret_cast<__nv_bfloat16*>(p_bf16)[i] = __float2bfloat16_rn(pp);
  }
}

template <int OPT, int EMA, int GT>
__global__ void __launch_bounds__(kUpdThreads, kUpdCtasPerSm)
fused_update_kernel(float* __restrict__ p, void* __restrict__ g, float* __restrict__ m,
                    float* __restrict__ v, const uint8_t* __restrict__ mask,
                    float* __restrict__ ema, void* __restrict__ p_bf16, int64_t n,
                    UpdateConsts c_arg, const DevConsts* __restrict__ c_dev,
                    const double* __restrict__ clip_sumsq) {
  fused_update_body<OPT, EMA, GT>(p, g, m, v, mask, ema, p_bf16, n, c_arg, c_dev, clip_sumsq);
}

// ---- clip norm + update in ONE cooperative launch (small vectors) ----------------------------------------
// zero + masked sum of squares + step-dependent scalars + fused update are four launches of 2-10 us each when the
// vector has ~1e7 elements (ResNet-18, an 8-way shard of DiT-XL/2): launch latency, not bandwidth.  As phases of one
// persistent grid separated by two grid-wide barriers they cost one launch, and the second read of g hits the L2.
template <int OPT, int EMA, int GT>
__global__ void __launch_bounds__(kUpdThreads, kUpdCtasPerSm)
clipped_update_coop_kernel(float* __restrict__ p, void* __restrict__ g, float* __restrict__ m,
                           float* __restrict__ v, const uint8_t* __restrict__ mask,
                           float* __restrict__ ema, void* __restrict__ p_bf16, int64_t n,
                           sfr_update_args a, bool has_momentum, long long* step_counter,
                           double* __restrict__ sumsq) {
  cg::grid_group grid = cg::this_grid();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *sumsq = 0.0;
    if (step_counter) ++(*step_counter);   // optimizer state['step'] lives on the device (CUDA-graph replay)
  }
  grid.sync();
  // the norm clip_grad_norm_ sees: of the masked gradient (SFR-on order) or of the raw one (SalUn order / no mask)
  if (a.flags & SFR_F_MASK) masked_sumsq_body<GT, true, kUpdThreads>(g, mask, n, sumsq);
  else masked_sumsq_body<GT, false, kUpdThreads>(g, mask, n, sumsq);
  grid.sync();
  const long long step = step_counter ? *step_counter : (long long)a.step;
  resolve_lr(a);
  UpdateConsts c = make_update_consts(a, step, has_momentum);
  if (OPT == SFR_OPT_SGD && step_counter) {
    // first use of the momentum buffer (buf = clone(grad)) comes from the device counter
    if (step == 1) c.flags |= SFR_F_SGD_FIRST_STEP; else c.flags &= ~SFR_F_SGD_FIRST_STEP;
  }
  fused_update_body<OPT, EMA, GT>(p, g, m, v, mask, ema, p_bf16, n, c, nullptr, sumsq);
}

// ------------------------------------------------------------------------- EMA alone
template <int EMA>
__global__ void __launch_bounds__(kThreads, 4)
ema_only_kernel(const float* __restrict__ p, float* __restrict__ ema, int64_t n, UpdateConsts c) {
  const int64_t nvec = n >> 2;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  float4* e4 = reinterpret_cast<float4*>(ema);
  for (int64_t vec = (int64_t)blockIdx.x * kThreads + threadIdx.x; vec < nvec;
       vec += (int64_t)gridDim.x * kThreads) {
    float4 pp = ld_stream(p4 + vec);
    float4 ee = ld_stream(e4 + vec);
    ee.x = ema_step<EMA>(pp.x, ee.x, c);
    ee.y = ema_step<EMA>(pp.y, ee.y, c);
    ee.z = ema_step<EMA>(pp.z, ee.z, c);
    ee.w = ema_step<EMA>(pp.w, ee.w, c);
    st_stream(e4 + vec, ee);
  }
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    float pp = p[i];
    ema[i] = ema_step<EMA>(pp, ema[i], c);
  }
}

__global__ void update_consts_kernel(sfr_update_args a, bool has_momentum, long long* step_counter,
                                     const double* clip_sumsq, DevConsts* out) {
  // optimizer state['step'] lives on the device when a counter is given (graph replay)
  const long long step = step_counter ? ++(*step_counter) : (long long)a.step;
  resolve_lr(a);
  const UpdateConsts c = make_update_consts(a, step, has_momentum);
  DevConsts d;
  d.neg_lr = c.neg_lr;
  d.decay_mul = c.decay_mul;
  d.has_lr = a.lr_table_dev != nullptr;
  d.neg_step_size = c.neg_step_size;
  d.bc2_sqrt = c.bc2_sqrt;
  d.clip_coef = clip_sumsq ? clip_coef_from_sumsq(clip_sumsq, (float)a.clip_max_norm) : 1.0f;
  // first use of the momentum buffer (buf = clone(grad)): from the device counter if there is one
  d.sgd_first_step = step_counter ? (step == 1) : ((a.flags & SFR_F_SGD_FIRST_STEP) != 0);
  *out = d;
}

template <int OPT, int EMA>
void launch_update_gt(int gt, int grid, cudaStream_t s, float* p, void* g, float* m, float* v,
                      const uint8_t* mask, float* ema, void* p_bf16, int64_t n,
                      const UpdateConsts& c, const DevConsts* c_dev, const double* clip_sumsq) {
  if (gt == SFR_F32)
    fused_update_kernel<OPT, EMA, SFR_F32><<<grid, kUpdThreads, 0, s>>>(p, g, m, v, mask, ema, p_bf16, n, c, c_dev, clip_sumsq);
  else
    fused_update_kernel<OPT, EMA, SFR_BF16><<<grid, kUpdThreads, 0, s>>>(p, g, m, v, mask, ema, p_bf16, n, c, c_dev, clip_sumsq);
}

template <int OPT>
void launch_update_ema(int ema_mode, int gt, int grid, cudaStream_t s, float* p, void* g,
                       float* m, float* v, const uint8_t* mask, float* ema, void* p_bf16,
                       int64_t n, const UpdateConsts& c, const DevConsts* c_dev,
                       const double* clip_sumsq) {
  switch (ema_mode) {
    case SFR_EMA_DDPM: launch_update_gt<OPT, SFR_EMA_DDPM>(gt, grid, s, p, g, m, v, mask, ema, p_bf16, n, c, c_dev, clip_sumsq); break;
    case SFR_EMA_DIT: launch_update_gt<OPT, SFR_EMA_DIT>(gt, grid, s, p, g, m, v, mask, ema, p_bf16, n, c, c_dev, clip_sumsq); break;
    case SFR_EMA_SLOWFAST: launch_update_gt<OPT, SFR_EMA_SLOWFAST>(gt, grid, s, p, g, m, v, mask, ema, p_bf16, n, c, c_dev, clip_sumsq); break;
    default: launch_update_gt<OPT, SFR_EMA_NONE>(gt, grid, s, p, g, m, v, mask, ema, p_bf16, n, c, c_dev, clip_sumsq); break;
  }
}

// ---- cooperative launch plumbing ---------------------------------------------------------------------
template <int OPT, int EMA, int GT>
cudaError_t launch_coop(int device, cudaStream_t s, float* p, void* g, float* m, float* v, const uint8_t* mask,
                        float* ema, void* p_bf16, int64_t n, sfr_update_args a, bool has_momentum,
                        long long* step_counter, double* sumsq) {
  static std::atomic<int> cached[64];
  int cap = cached[device & 63].load(std::memory_order_acquire);
  if (cap == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, clipped_update_coop_kernel<OPT, EMA, GT>, kUpdThreads,
                                                      0) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    cap = per_sm * device_geometry().sm_count;
    cached[device & 63].store(cap, std::memory_order_release);
  }
  const int64_t tiles = ((n >> 2) + kUpdThreads - 1) / kUpdThreads;
  const int grid = (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
  void* args[] = {(void*)&p, (void*)&g, (void*)&m, (void*)&v, (void*)&mask, (void*)&ema, (void*)&p_bf16, (void*)&n,
                  (void*)&a, (void*)&has_momentum, (void*)&step_counter, (void*)&sumsq};
  return cudaLaunchCooperativeKernel((const void*)clipped_update_coop_kernel<OPT, EMA, GT>, dim3(grid),
                                     dim3(kUpdThreads), args, 0, s);
}

template <int OPT, int EMA>
cudaError_t launch_coop_gt(int gt, int device, cudaStream_t s, float* p, void* g, float* m, float* v,
                           const uint8_t* mask, float* ema, void* p_bf16, int64_t n, const sfr_update_args& a,
                           bool has_momentum, long long* step_counter, double* sumsq) {
  if (gt == SFR_F32)
    return launch_coop<OPT, EMA, SFR_F32>(device, s, p, g, m, v, mask, ema, p_bf16, n, a, has_momentum, step_counter, sumsq);
  return launch_coop<OPT, EMA, SFR_BF16>(device, s, p, g, m, v, mask, ema, p_bf16, n, a, has_momentum, step_counter, sumsq);
}

template <int OPT>
cudaError_t launch_coop_ema(int device, cudaStream_t s, float* p, void* g, float* m, float* v, const uint8_t* mask,
                            float* ema, void* p_bf16, int64_t n, const sfr_update_args& a, bool has_momentum,
                            long long* step_counter, double* sumsq) {
#define SFR_COOP(E) launch_coop_gt<OPT, E>(a.g_dtype, device, s, p, g, m, v, mask, ema, p_bf16, n, a, has_momentum, step_counter, sumsq)
  switch (a.ema_mode) {
    case SFR_EMA_DDPM: return SFR_COOP(SFR_EMA_DDPM);
    case SFR_EMA_DIT: return SFR_COOP(SFR_EMA_DIT);
    case SFR_EMA_SLOWFAST: return SFR_COOP(SFR_EMA_SLOWFAST);
    default: return SFR_COOP(SFR_EMA_NONE);
  }
#undef SFR_COOP
}

}  // namespace

void launch_update_consts(const sfr_update_args& a, bool has_momentum, long long* step_counter,
                          const double* clip_sumsq, void* consts_scratch, cudaStream_t s) {
  update_consts_kernel<<<1, 1, 0, s>>>(a, has_momentum, step_counter, clip_sumsq,
                                       reinterpret_cast<DevConsts*>(consts_scratch));
}

}  // namespace sfr

extern "C" int sfr_masked_sumsq(const void* g, int g_dtype, const uint8_t* mask, int64_t n,
                                double* out, sfr_stream_t stream) {
  using namespace sfr;
  if (n < 0) return SFR_ERR_ARG;
  if (g_dtype != SFR_F32 && g_dtype != SFR_BF16) return SFR_ERR_ARG;
  if (n == 0) return SFR_OK;
  SFR_REQUIRE_PTR(g);
  SFR_REQUIRE_PTR(out);
  SFR_REQUIRE_ALIGNED(g);
  SFR_REQUIRE_ALIGNED(mask);
  SFR_ENTER_DEVICE(g);
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kThreads * kSumsqUnroll;
  const int grid = persistent_grid((nvec + tile - 1) / tile, kSumsqCtasPerSm * kSumsqGridWaves);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (g_dtype == SFR_F32) {
    if (mask) masked_sumsq_kernel<SFR_F32, true><<<grid, kThreads, 0, s>>>(g, mask, n, out);
    else masked_sumsq_kernel<SFR_F32, false><<<grid, kThreads, 0, s>>>(g, mask, n, out);
  } else {
    if (mask) masked_sumsq_kernel<SFR_BF16, true><<<grid, kThreads, 0, s>>>(g, mask, n, out);
    else masked_sumsq_kernel<SFR_BF16, false><<<grid, kThreads, 0, s>>>(g, mask, n, out);
  }
  SFR_LAUNCH_STATUS();
}

// One cooperative launch: zero + masked sum of squares + step-dependent scalars + K3.  The norm is LEFT in *sumsq.
extern "C" int sfr_clipped_update(float* p, void* g, float* m, float* v, const uint8_t* mask,
                                  float* ema, void* p_bf16, int64_t n, const sfr_update_args* a,
                                  double* sumsq, long long* step_counter, sfr_stream_t stream) {
  using namespace sfr;
  SFR_REQUIRE_PTR(a);
  SFR_REQUIRE_PTR(sumsq);
  if (a->clip_max_norm <= 0.0) return SFR_ERR_ARG;
  // same argument contract as sfr_fused_update (checked there as well, before anything is launched)
  {
    const int rc = sfr_fused_update(p, g, m, v, mask, ema, p_bf16, -2, a, sumsq, step_counter, nullptr, stream);
    if (rc != SFR_OK) return rc;
  }
  if (n == 0) return SFR_OK;
  if (n < 0) return SFR_ERR_ARG;
  const bool has_momentum = a->opt == SFR_OPT_SGD && a->momentum != 0.0;
  SFR_ENTER_DEVICE(p);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int dev = device_scope__.device();
  cudaError_t e;
  switch (a->opt) {
    case SFR_OPT_SGD: e = launch_coop_ema<SFR_OPT_SGD>(dev, s, p, g, m, v, mask, ema, p_bf16, n, *a, has_momentum, step_counter, sumsq); break;
    case SFR_OPT_ADAM: e = launch_coop_ema<SFR_OPT_ADAM>(dev, s, p, g, m, v, mask, ema, p_bf16, n, *a, has_momentum, step_counter, sumsq); break;
    default: e = launch_coop_ema<SFR_OPT_ADAMW>(dev, s, p, g, m, v, mask, ema, p_bf16, n, *a, has_momentum, step_counter, sumsq); break;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return (int)e;
  }
  SFR_LAUNCH_STATUS();
}

extern "C" int sfr_fused_update(float* p, void* g, float* m, float* v, const uint8_t* mask,
                                float* ema, void* p_bf16, int64_t n,
                                const sfr_update_args* a, const double* clip_sumsq,
                                long long* step_counter, void* consts_scratch,
                                sfr_stream_t stream) {
  using namespace sfr;
  SFR_REQUIRE_PTR(a);
  // n == -2: validate the arguments only (sfr_clipped_update shares this contract)
  const bool validate_only = n == -2;
  if (validate_only) n = 1;
  if (n < 0) return SFR_ERR_ARG;
  if (a->opt < SFR_OPT_SGD || a->opt > SFR_OPT_ADAMW) return SFR_ERR_ARG;
  if (a->ema_mode < SFR_EMA_NONE || a->ema_mode > SFR_EMA_SLOWFAST) return SFR_ERR_ARG;
  if (a->g_dtype != SFR_F32 && a->g_dtype != SFR_BF16) return SFR_ERR_ARG;
  const uint32_t known = SFR_F_MASK | SFR_F_MASK_AFTER_CLIP | SFR_F_ZERO_GRAD |
                         SFR_F_SGD_FIRST_STEP | SFR_F_WRITE_BF16 | SFR_F_REUSE_CONSTS;
  if (a->flags & ~known) return SFR_ERR_ARG;
  if ((a->flags & SFR_F_MASK) && (a->flags & SFR_F_MASK_AFTER_CLIP)) return SFR_ERR_ARG;
  if (a->opt != SFR_OPT_SGD && a->step < 1 && step_counter == nullptr) return SFR_ERR_ARG;
  if (n == 0) return SFR_OK;
  const bool use_mask = (a->flags & (SFR_F_MASK | SFR_F_MASK_AFTER_CLIP)) != 0;
  const bool has_momentum = a->opt == SFR_OPT_SGD && a->momentum != 0.0;
  SFR_REQUIRE_PTR(p);
  SFR_REQUIRE_PTR(g);
  if (a->opt != SFR_OPT_SGD || has_momentum) SFR_REQUIRE_PTR(m);
  if (a->opt != SFR_OPT_SGD) SFR_REQUIRE_PTR(v);
  if (use_mask) SFR_REQUIRE_PTR(mask);
  if (a->ema_mode != SFR_EMA_NONE) SFR_REQUIRE_PTR(ema);
  if (a->flags & SFR_F_WRITE_BF16) SFR_REQUIRE_PTR(p_bf16);
  SFR_REQUIRE_ALIGNED(p);
  SFR_REQUIRE_ALIGNED(g);
  SFR_REQUIRE_ALIGNED(m);
  SFR_REQUIRE_ALIGNED(v);
  SFR_REQUIRE_ALIGNED(mask);
  SFR_REQUIRE_ALIGNED(ema);
  SFR_REQUIRE_ALIGNED(p_bf16);
  if (validate_only) return SFR_OK;
  SFR_ENTER_DEVICE(p);

  UpdateConsts c = make_update_consts(*a, a->step, has_momentum);
  const DevConsts* c_dev = nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((a->lr_table_dev == nullptr) != (a->lr_index_dev == nullptr)) return SFR_ERR_NULL;
  if (step_counter != nullptr || a->lr_table_dev != nullptr || (a->flags & SFR_F_REUSE_CONSTS))
    SFR_REQUIRE_PTR(consts_scratch);
  if (consts_scratch != nullptr) {
    // one-thread prep kernel: step-dependent scalars (from the device counter if given: graph
    // replay) and the clip coefficient, precomputed so the main kernel never stalls on them
    if (!aligned16(consts_scratch)) return SFR_ERR_ALIGN;
    if (!(a->flags & SFR_F_REUSE_CONSTS))
      launch_update_consts(*a, has_momentum, step_counter, clip_sumsq, consts_scratch, s);
    c_dev = reinterpret_cast<const DevConsts*>(consts_scratch);
  }

  const int64_t nvec = n >> 2;
  const int grid = full_grid((nvec + kUpdThreads - 1) / kUpdThreads);
  switch (a->opt) {
    case SFR_OPT_SGD:
      launch_update_ema<SFR_OPT_SGD>(a->ema_mode, a->g_dtype, grid, s, p, g, m, v, mask, ema, p_bf16, n, c, c_dev, clip_sumsq);
      break;
    case SFR_OPT_ADAM:
      launch_update_ema<SFR_OPT_ADAM>(a->ema_mode, a->g_dtype, grid, s, p, g, m, v, mask, ema, p_bf16, n, c, c_dev, clip_sumsq);
      break;
    default:
      launch_update_ema<SFR_OPT_ADAMW>(a->ema_mode, a->g_dtype, grid, s, p, g, m, v, mask, ema, p_bf16, n, c, c_dev, clip_sumsq);
      break;
  }
  SFR_LAUNCH_STATUS();
}

extern "C" int sfr_ema_update(const float* p, float* ema, int64_t n, int ema_mode, double ema_a,
                              sfr_stream_t stream) {
  using namespace sfr;
  if (n < 0) return SFR_ERR_ARG;
  if (ema_mode != SFR_EMA_DDPM && ema_mode != SFR_EMA_DIT) return SFR_ERR_ARG;
  if (n == 0) return SFR_OK;
  SFR_REQUIRE_PTR(p);
  SFR_REQUIRE_PTR(ema);
  SFR_REQUIRE_ALIGNED(p);
  SFR_REQUIRE_ALIGNED(ema);
  SFR_ENTER_DEVICE(p);
  UpdateConsts c{};
  fill_ema_consts(c, ema_mode, ema_a);
  const int64_t nvec = n >> 2;
  const int grid = persistent_grid((nvec + kThreads - 1) / kThreads, 32);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ema_mode == SFR_EMA_DDPM) ema_only_kernel<SFR_EMA_DDPM><<<grid, kThreads, 0, s>>>(p, ema, n, c);
  else ema_only_kernel<SFR_EMA_DIT><<<grid, kThreads, 0, s>>>(p, ema, n, c);
  SFR_LAUNCH_STATUS();
}
