// extras.cu — the two Fisher / select CONSUMERS next to the hot path.
//
// n2  EWC / Selective-Amnesia penalty        DDPM/runners/diffusion.py:424-433 (sa_forget)
//       reference, per named tensor and per step:
//         _loss = fisher[name] * (param - params_mle[name]) ** 2 ; loss += lmbda * _loss.sum()
//       and autograd's backward of it:  grad += (lmbda * fisher) * (2 * (param - params_mle))
//     here: one pass that adds that gradient into the flat g and reduces the penalty value.
//     20 B/elem (F, p, p*, g read; g written).
//
// n3  proximal-gradient soft threshold       SD/train-scripts/proximal_gradient.py:151-183
//       threshold = k-th smallest |theta - theta0|  (reference: torch.topk over a 1.07 B-element copy on
//       a second GPU) — found with the K2b select in key mode SFR_KEY_ABSDIFF — then
//         d = p - p0 ; d > thr: d -= thr ; d < -thr: d += thr ; else d = 0 ; p = d + p0
//     12 B/elem (p, p0 read; p written).
#include "common.cuh"

namespace sfr {
namespace {

constexpr int kThreads = 128;
constexpr int kUnroll = 1;

__device__ __forceinline__ float ewc_term(float f, float p, float ps, float lambda, float& gadd) {
  const float d = __fsub_rn(p, ps);
  // backward of lmbda * sum(F * d**2):  mul-backward (lmbda * F), pow-backward (* (2 * d))
  gadd = __fmul_rn(__fmul_rn(lambda, f), __fmul_rn(2.0f, d));
  return __fmul_rn(f, __fmul_rn(d, d));  // F * d**2
}

__global__ void __launch_bounds__(kThreads, 8)
ewc_penalty_kernel(const float* __restrict__ p, const float* __restrict__ ps,
                   const float* __restrict__ fisher, float* __restrict__ g, int64_t n, float lambda,
                   double* __restrict__ penalty) {
  __shared__ double scratch[32];
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kThreads * kUnroll;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  const float4* s4 = reinterpret_cast<const float4*>(ps);
  const float4* f4 = reinterpret_cast<const float4*>(fisher);
  float4* g4 = reinterpret_cast<float4*>(g);
  double total = 0.0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 pp[kUnroll], ss[kUnroll], ff[kUnroll], gg[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kThreads;
      const bool in = v < nvec;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      pp[u] = in ? ld_stream(p4 + v) : z;
      ss[u] = in ? ld_stream(s4 + v) : z;
      ff[u] = in ? ld_stream(f4 + v) : z;
      gg[u] = in ? ld_stream(g4 + v) : z;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kThreads;
      if (v >= nvec) continue;
      float a0, a1, a2, a3, part;
      part = ewc_term(ff[u].x, pp[u].x, ss[u].x, lambda, a0);
      part = __fadd_rn(part, ewc_term(ff[u].y, pp[u].y, ss[u].y, lambda, a1));
      part = __fadd_rn(part, ewc_term(ff[u].z, pp[u].z, ss[u].z, lambda, a2));
      part = __fadd_rn(part, ewc_term(ff[u].w, pp[u].w, ss[u].w, lambda, a3));
      total += (double)part;
      gg[u].x = __fadd_rn(gg[u].x, a0);
      gg[u].y = __fadd_rn(gg[u].y, a1);
      gg[u].z = __fadd_rn(gg[u].z, a2);
      gg[u].w = __fadd_rn(gg[u].w, a3);
      st_stream(g4 + v, gg[u]);
    }
  }
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    float add;
    total += (double)ewc_term(fisher[i], p[i], ps[i], lambda, add);
    g[i] = __fadd_rn(g[i], add);
  }
  total = block_sum<double>(total, scratch);
  if (threadIdx.x == 0 && penalty != nullptr) atomicAdd(penalty, total * (double)lambda);
}

__device__ __forceinline__ float shrink(float p, float p0, float thr) {
  float d = __fsub_rn(p, p0);                    // param -= init_param
  if (d > thr) d = __fsub_rn(d, thr);            // param[larger] -= threshold
  else if (d < -thr) d = __fadd_rn(d, thr);      // param[smaller] += threshold
  else d = 0.0f;                                 // param[between] = 0  (NaN lands here too, as in the reference)
  return __fadd_rn(d, p0);                       // param += init_param
}

__global__ void __launch_bounds__(kThreads, 8)
soft_threshold_kernel(float* __restrict__ p, const float* __restrict__ p0, int64_t n,
                      const float* __restrict__ thr_dev) {
  const float thr = *thr_dev;
  const int64_t nvec = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* q4 = reinterpret_cast<const float4*>(p0);
  for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * kThreads) {
    float4 a = ld_stream(p4 + v);
    const float4 b = ld_stream(q4 + v);
    a.x = shrink(a.x, b.x, thr);
    a.y = shrink(a.y, b.y, thr);
    a.z = shrink(a.z, b.z, thr);
    a.w = shrink(a.w, b.w, thr);
    st_stream(p4 + v, a);
  }
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    p[i] = shrink(p[i], p0[i], thr);
  }
}

__global__ void select_threshold_value_kernel(const sfr_select_state* __restrict__ state,
                                              float* __restrict__ out) {
  // key = bits(|x|) + 1 for numbers, 0 for NaN; select_all / none carry no threshold
  const uint32_t key = state->thr_key;
  float v;
  if (state->select_none) v = __uint_as_float(0x7f800000u);       // nothing selected: +inf
  else if (state->select_all || key == 0u) v = 0.0f;
  else v = __uint_as_float(key - 1u);
  *out = v;
}

}  // namespace
}  // namespace sfr

extern "C" int sfr_ewc_penalty(const float* p, const float* p_star, const float* fisher, float* g,
                               int64_t n, float lambda, double* penalty, sfr_stream_t stream) {
  using namespace sfr;
  if (n < 0) return SFR_ERR_ARG;
  if (n == 0) return SFR_OK;
  SFR_REQUIRE_PTR(p);
  SFR_REQUIRE_PTR(p_star);
  SFR_REQUIRE_PTR(fisher);
  SFR_REQUIRE_PTR(g);
  SFR_REQUIRE_ALIGNED(p);
  SFR_REQUIRE_ALIGNED(p_star);
  SFR_REQUIRE_ALIGNED(fisher);
  SFR_REQUIRE_ALIGNED(g);
  SFR_ENTER_DEVICE(p);
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kThreads * kUnroll;
  const int grid = persistent_grid((nvec + tile - 1) / tile, 16 * 128);
  ewc_penalty_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p, p_star, fisher, g, n, lambda, penalty);
  SFR_LAUNCH_STATUS();
}

extern "C" int sfr_select_threshold_value(const sfr_select_state* state, float* out, sfr_stream_t stream) {
  using namespace sfr;
  SFR_REQUIRE_PTR(state);
  SFR_REQUIRE_PTR(out);
  SFR_ENTER_DEVICE(state);
  select_threshold_value_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(state, out);
  SFR_LAUNCH_STATUS();
}

extern "C" int sfr_soft_threshold(float* p, const float* p0, int64_t n, const float* threshold_dev,
                                  sfr_stream_t stream) {
  using namespace sfr;
  if (n < 0) return SFR_ERR_ARG;
  if (n == 0) return SFR_OK;
  SFR_REQUIRE_PTR(p);
  SFR_REQUIRE_PTR(p0);
  SFR_REQUIRE_PTR(threshold_dev);
  SFR_REQUIRE_ALIGNED(p);
  SFR_REQUIRE_ALIGNED(p0);
  SFR_ENTER_DEVICE(p);
  const int64_t nvec = n >> 2;
  const int grid = persistent_grid((nvec + kThreads - 1) / kThreads, 16 * 128);
  soft_threshold_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p, p0, n, threshold_dev);
  SFR_LAUNCH_STATUS();
}
