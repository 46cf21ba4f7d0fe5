// mask.cu — K2a: forget/remain Fisher-ratio saliency mask.
//
//   mask[i] = ((ff[i] + eps) / (rf[i] + eps)) >= threshold        (bool, 1 byte)
//   *zero_count += #(mask == 0)                                    ("Total sparsity")
//
// Reference (CPU, per tensor):  Classification/unlearn/sfron.py:325-334,
//   DDPM/generate_fisher_mask.py:39-45, DiT/generate_mask.py:31-39,
//   SD/train-scripts/generate_fisher_mask.py:39-45.
// The multi-threshold form does DiT/generate_mask.py:25-46's threshold loop in ONE pass
// over the two Fisher vectors instead of re-reading them per threshold.
//
// HBM-bound: 8 B read + T B written per element.  IEEE add / divide / compare: bit-exact.
#include "common.cuh"

namespace sfr {
namespace {

constexpr int kThreads = 128;
constexpr int kCtasPerSm = 8;
constexpr int kUnroll = 2;
constexpr int kGridWaves = 32;

struct Thresholds {
  float v[SFR_MAX_THRESHOLDS];
};

__device__ __forceinline__ float ratio_of(float f, float r, float eps) {
  return __fdiv_rn(__fadd_rn(f, eps), __fadd_rn(r, eps));
}

template <int T>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
ratio_mask_kernel(const float* __restrict__ ff, const float* __restrict__ rf, int64_t n,
                  Thresholds th, float eps, uint8_t* __restrict__ masks, int64_t mask_stride,
                  unsigned long long* __restrict__ zero_counts) {
  __shared__ unsigned int scratch[32];
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kThreads * kUnroll;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  const float4* ff4 = reinterpret_cast<const float4*>(ff);
  const float4* rf4 = reinterpret_cast<const float4*>(rf);
  unsigned int ones[T];
#pragma unroll
  for (int k = 0; k < T; ++k) ones[k] = 0;
  unsigned int seen = 0;

  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 f[kUnroll], r[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kThreads;
      const bool in = v < nvec;
      f[u] = in ? ld_stream(ff4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      r[u] = in ? ld_stream(rf4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kThreads;
      if (v >= nvec) continue;
      const float q0 = ratio_of(f[u].x, r[u].x, eps);
      const float q1 = ratio_of(f[u].y, r[u].y, eps);
      const float q2 = ratio_of(f[u].z, r[u].z, eps);
      const float q3 = ratio_of(f[u].w, r[u].w, eps);
      seen += 4;
#pragma unroll
      for (int k = 0; k < T; ++k) {
        const unsigned int b0 = q0 >= th.v[k], b1 = q1 >= th.v[k];
        const unsigned int b2 = q2 >= th.v[k], b3 = q3 >= th.v[k];
        ones[k] += b0 + b1 + b2 + b3;
        *(reinterpret_cast<unsigned int*>(masks + (int64_t)k * mask_stride) + v) =
            b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
      }
    }
  }

  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    const float q = ratio_of(ff[i], rf[i], eps);
    seen += 1;
#pragma unroll
    for (int k = 0; k < T; ++k) {
      const unsigned int b = q >= th.v[k];
      ones[k] += b;
      masks[(int64_t)k * mask_stride + i] = (uint8_t)b;
    }
  }

  if (zero_counts != nullptr) {
#pragma unroll
    for (int k = 0; k < T; ++k) {
      // a thread sees < 2^32 elements for any n this library accepts per launch
      unsigned int zeros = block_sum<unsigned int>(seen - ones[k], scratch);
      if (threadIdx.x == 0 && zeros) atomicAdd(zero_counts + k, (unsigned long long)zeros);
      __syncthreads();
    }
  }
}

template <int T>
int launch_ratio(const float* ff, const float* rf, int64_t n, const Thresholds& th, float eps,
                 uint8_t* masks, int64_t mask_stride, unsigned long long* zero_counts,
                 cudaStream_t s) {
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kThreads * kUnroll;
  // every CTA ends with T block reductions + T atomics (zero counts): give each at least 8 tiles
  const int64_t ntiles = (nvec + tile - 1) / tile;
  const int64_t by_work = ntiles / 8 > 1 ? ntiles / 8 : 1;
  const int cap = persistent_grid(ntiles, 16 * kGridWaves);
  const int floor_ = persistent_grid(ntiles, kCtasPerSm);
  int grid = (int)(by_work < cap ? by_work : cap);
  if (grid < floor_) grid = floor_;
  ratio_mask_kernel<T><<<grid, kThreads, 0, s>>>(ff, rf, n, th, eps, masks, mask_stride,
                                                 zero_counts);
  SFR_LAUNCH_STATUS();
}

}  // namespace
}  // namespace sfr

extern "C" int sfr_ratio_mask_multi(const float* ff, const float* rf, int64_t n,
                                    const float* thresholds_host, int n_thresholds, float eps,
                                    uint8_t* masks, int64_t mask_stride,
                                    unsigned long long* zero_counts, sfr_stream_t stream) {
  using namespace sfr;
  if (n < 0 || n_thresholds < 1 || n_thresholds > SFR_MAX_THRESHOLDS) return SFR_ERR_ARG;
  // per-thread element counters are 32-bit: one launch covers < 2^40 elements per CTA share
  if (n == 0) return SFR_OK;
  SFR_REQUIRE_PTR(ff);
  SFR_REQUIRE_PTR(rf);
  SFR_REQUIRE_PTR(masks);
  SFR_REQUIRE_PTR(thresholds_host);
  SFR_REQUIRE_ALIGNED(ff);
  SFR_REQUIRE_ALIGNED(rf);
  SFR_REQUIRE_ALIGNED(masks);
  if (n_thresholds > 1 && (mask_stride < n || (mask_stride & 15) != 0)) return SFR_ERR_ARG;
  SFR_ENTER_DEVICE(ff);
  Thresholds th;
  for (int k = 0; k < SFR_MAX_THRESHOLDS; ++k)
    th.v[k] = k < n_thresholds ? thresholds_host[k] : 0.f;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (n_thresholds) {
    case 1: return launch_ratio<1>(ff, rf, n, th, eps, masks, mask_stride, zero_counts, s);
    case 2: return launch_ratio<2>(ff, rf, n, th, eps, masks, mask_stride, zero_counts, s);
    case 3: return launch_ratio<3>(ff, rf, n, th, eps, masks, mask_stride, zero_counts, s);
    case 4: return launch_ratio<4>(ff, rf, n, th, eps, masks, mask_stride, zero_counts, s);
    case 5: return launch_ratio<5>(ff, rf, n, th, eps, masks, mask_stride, zero_counts, s);
    case 6: return launch_ratio<6>(ff, rf, n, th, eps, masks, mask_stride, zero_counts, s);
    case 7: return launch_ratio<7>(ff, rf, n, th, eps, masks, mask_stride, zero_counts, s);
    default: return launch_ratio<8>(ff, rf, n, th, eps, masks, mask_stride, zero_counts, s);
  }
}

extern "C" int sfr_ratio_mask(const float* ff, const float* rf, int64_t n, float threshold,
                              float eps, uint8_t* mask, unsigned long long* zero_count,
                              sfr_stream_t stream) {
  return sfr_ratio_mask_multi(ff, rf, n, &threshold, 1, eps, mask, n, zero_count, stream);
}
