// fisher.cu — K1: Fisher-diagonal accumulation over a flat parameter shard.
//
//   acc[i] <- acc[i] + (g_b[i] * g_b[i]) / divisor        b = 0 .. rows-1, in order
//
// Reference op sequence (all on the CPU there, after a per-tensor D2H copy):
//   `F[name] += param.grad.data.cpu()**2 / len(loader)`
//     Classification/unlearn/sfron.py:288-291,315-318
//     DDPM/runners/diffusion.py:1277-1281,1342-1346 (gradient clipped first, :1270-1275)
//     DiT/generate_fisher.py:236-239,276-279 ; SD/train-scripts/generate_fisher.py:73-76,123-126
//   per-sample FIM `F += tmp_i**2 / |D|`   DDPM/runners/diffusion.py:337-344   (rows > 1)
//
// HBM-bound: (4*rows [or 2*rows bf16] + 4) B read + 4 B written per element, no reuse.
// Rounding sequence = the reference's three fp32 ops: mul, true divide, add.
#include "common.cuh"

namespace sfr {
namespace {

constexpr int kThreads = 128;
constexpr int kCtasPerSm = 12;  // register cap 42; 16 resident when the variant needs <= 32
constexpr int kUnroll = 2;      // independent 128-bit loads in flight per stream per thread
constexpr int kGridWaves = 32;  // grid = SMs x 16 x 32 CTAs (tuned: tools/tune/tune_stream.cu)

// SQUARE: the Fisher term (g*coef)**2 / L.  !SQUARE: the gradient itself, for SalUn's saliency accumulation
// `gradients[name] += param.grad` after clip_grad_norm_ (DDPM/runners/diffusion.py:985-994;
// Classification/unlearn/salun.py:163-169, unclipped): two roundings, mul then add, as on the CPU.
template <bool CLIP, bool SQUARE>
__device__ __forceinline__ float fisher_term(float g, float coef, float divisor) {
  if constexpr (CLIP) g = __fmul_rn(g, coef);  // clip_grad_norm_: grad.mul_(coef)
  if constexpr (!SQUARE) return g;
  return __fdiv_rn(__fmul_rn(g, g), divisor);  // g**2 / L   (true division, as on CPU)
}

template <int GT, bool CLIP, bool SQUARE = true>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
fisher_accum_kernel(float* __restrict__ acc, const void* __restrict__ g, int64_t rows,
                    int64_t row_stride, int64_t n, float divisor,
                    const double* __restrict__ clip_sumsq, float clip_max_norm) {
  const float coef = CLIP ? clip_coef_from_sumsq(clip_sumsq, clip_max_norm) : 1.0f;
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kThreads * kUnroll;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  float4* acc4 = reinterpret_cast<float4*>(acc);
  // row_stride is a multiple of 4 elements whenever rows > 1 (checked on the host)
  const int64_t row_vec_stride = row_stride >> 2;

  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 a[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kThreads;
      a[u] = v < nvec ? ld_stream(acc4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t r = 0; r < rows; ++r) {
      float4 gg[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t v = base + (int64_t)u * kThreads;
        gg[u] = v < nvec ? load_g4<GT>(g, r * row_vec_stride + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        a[u].x = __fadd_rn(a[u].x, fisher_term<CLIP, SQUARE>(gg[u].x, coef, divisor));
        a[u].y = __fadd_rn(a[u].y, fisher_term<CLIP, SQUARE>(gg[u].y, coef, divisor));
        a[u].z = __fadd_rn(a[u].z, fisher_term<CLIP, SQUARE>(gg[u].z, coef, divisor));
        a[u].w = __fadd_rn(a[u].w, fisher_term<CLIP, SQUARE>(gg[u].w, coef, divisor));
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kThreads;
      if (v < nvec) st_stream(acc4 + v, a[u]);
    }
  }

  // ragged tail: n % 4 elements, one thread each
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    float a = acc[i];
    for (int64_t r = 0; r < rows; ++r)
      a = __fadd_rn(a, fisher_term<CLIP, SQUARE>(load_g1<GT>(g, r * row_stride + i), coef, divisor));
    acc[i] = a;
  }
}

}  // namespace
}  // namespace sfr

extern "C" int sfr_fisher_accum(float* acc, const void* g, int g_dtype, int64_t rows,
                                int64_t row_stride, int64_t n, float divisor,
                                const double* clip_sumsq, float clip_max_norm,
                                sfr_stream_t stream) {
  using namespace sfr;
  if (n < 0 || rows < 0) return SFR_ERR_ARG;
  if (n == 0 || rows == 0) return SFR_OK;
  SFR_REQUIRE_PTR(acc);
  SFR_REQUIRE_PTR(g);
  SFR_REQUIRE_ALIGNED(acc);
  SFR_REQUIRE_ALIGNED(g);
  if (g_dtype != SFR_F32 && g_dtype != SFR_BF16) return SFR_ERR_ARG;
  // every row must start 16-byte aligned: stride a multiple of 4 fp32 / 8 bf16 elements
  if (rows > 1 && (row_stride < n || (row_stride & (g_dtype == SFR_F32 ? 3 : 7)) != 0)) return SFR_ERR_ARG;
  SFR_ENTER_DEVICE(acc);

  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kThreads * kUnroll;
  const int grid = persistent_grid((nvec + tile - 1) / tile, 16 * kGridWaves);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool clip = clip_sumsq != nullptr;
#define SFR_K1(GT, CL)                                                                  \
  fisher_accum_kernel<GT, CL><<<grid, kThreads, 0, s>>>(acc, g, rows, row_stride, n,    \
                                                        divisor, clip_sumsq, clip_max_norm)
  if (g_dtype == SFR_F32) {
    if (clip) SFR_K1(SFR_F32, true); else SFR_K1(SFR_F32, false);
  } else {
    if (clip) SFR_K1(SFR_BF16, true); else SFR_K1(SFR_BF16, false);
  }
#undef SFR_K1
  SFR_LAUNCH_STATUS();
}


// SalUn saliency accumulation: acc += g [* clip coefficient]   (12 B/elem, one pass, no temporary)
extern "C" int sfr_grad_accum(float* acc, const void* g, int g_dtype, int64_t n,
                              const double* clip_sumsq, float clip_max_norm, sfr_stream_t stream) {
  using namespace sfr;
  if (n < 0) return SFR_ERR_ARG;
  if (n == 0) return SFR_OK;
  SFR_REQUIRE_PTR(acc);
  SFR_REQUIRE_PTR(g);
  SFR_REQUIRE_ALIGNED(acc);
  SFR_REQUIRE_ALIGNED(g);
  if (g_dtype != SFR_F32 && g_dtype != SFR_BF16) return SFR_ERR_ARG;
  SFR_ENTER_DEVICE(acc);
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kThreads * kUnroll;
  const int grid = persistent_grid((nvec + tile - 1) / tile, 16 * kGridWaves);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define SFR_ACC(GT, CL) \
  fisher_accum_kernel<GT, CL, false><<<grid, kThreads, 0, s>>>(acc, g, 1, n, n, 1.0f, clip_sumsq, clip_max_norm)
  if (g_dtype == SFR_F32) {
    if (clip_sumsq) SFR_ACC(SFR_F32, true); else SFR_ACC(SFR_F32, false);
  } else {
    if (clip_sumsq) SFR_ACC(SFR_BF16, true); else SFR_ACC(SFR_BF16, false);
  }
#undef SFR_ACC
  SFR_LAUNCH_STATUS();
}
