// tune_tma.cu — does staging a once-read stream through shared memory with TMA bulk copies (cp.async.bulk +
// mbarrier) beat plain 128-bit loads on B200?  (not product code; evidence for DESIGN.md §3 "what is NOT used")
//
// Shape: the read-only reduction of the path (clip norm / histogram passes): 4 B/elem in, nothing out.
//   direct : persistent grid-stride loop, ld.global.cs.v4 (what the product kernels do)
//   tma    : one elected thread streams 16 KB tiles global -> shared with cp.async.bulk into a STAGES-deep
//            ring, completion on an mbarrier (complete_tx::bytes); all threads consume from shared memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tune/tune_tma.bin tools/tune/tune_tma.cu
// Run  : timeout 60 tools/tune/tune_tma.bin            (prints GB/s per variant and checks the sums agree)
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double block_reduce(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;
}

template <int THREADS, int UNROLL, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) sumsq_direct(const float4* __restrict__ g, int64_t nvec, double* out) {
  __shared__ double red[32];
  float acc = 0.f;
  const int64_t tile = (int64_t)THREADS * UNROLL, ntiles = (nvec + tile - 1) / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t v = base + (int64_t)u * THREADS;
      x[u] = v < nvec ? __ldcs(g + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += x[u].x * x[u].x + x[u].y * x[u].y + x[u].z * x[u].z + x[u].w * x[u].w;
  }
  const double r = block_reduce((double)acc, red);
  if (threadIdx.x == 0) atomicAdd(out, r);
}

template <int THREADS, int STAGES, int TILE_F4, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) sumsq_tma(const float4* __restrict__ g, int64_t nvec, double* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* buf = reinterpret_cast<float4*>(smem_raw);
  __shared__ __align__(8) unsigned long long full[STAGES];
  __shared__ double red[32];
  const int64_t ntiles = nvec / TILE_F4;           // the experiment uses n that is a multiple of the tile
  constexpr uint32_t kBytes = TILE_F4 * 16;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s, int64_t tile) {
    const uint32_t bar = smem_u32(&full[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(buf + (int64_t)s * TILE_F4)), "l"(g + tile * TILE_F4), "r"(kBytes), "r"(bar)
                 : "memory");
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; ++s) {
      const int64_t tile = blockIdx.x + (int64_t)s * gridDim.x;
      if (tile < ntiles) issue(s, tile);
    }
  float acc = 0.f;
  for (int64_t k = 0;; ++k) {
    const int64_t tile = blockIdx.x + k * gridDim.x;
    if (tile >= ntiles) break;
    const int s = (int)(k % STAGES);
    const uint32_t phase = (uint32_t)((k / STAGES) & 1), bar = smem_u32(&full[s]);
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    const float4* t4 = buf + (int64_t)s * TILE_F4;
#pragma unroll 4
    for (int j = threadIdx.x; j < TILE_F4; j += THREADS) {
      const float4 x = t4[j];
      acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    __syncthreads();                                   // every thread is done with stage s
    if (threadIdx.x == 0) {
      const int64_t next = blockIdx.x + (k + STAGES) * gridDim.x;
      if (next < ntiles) issue(s, next);
    }
  }
  const double r = block_reduce((double)acc, red);
  if (threadIdx.x == 0) atomicAdd(out, r);
}

template <typename F>
static float time_ms(F launch, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  std::vector<float> t;
  for (int i = 0; i < iters + 2; ++i) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (i >= 2) t.push_back(ms);
  }
  std::sort(t.begin(), t.end());
  return t[t.size() / 2];
}

__global__ void fill(float* g, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    g[i] = (float)((i * 2654435761u) & 1023) * (1.0f / 1024.0f) - 0.5f;
}

int main() {
  const int64_t n = 675129632 / 4096 * 4096, nvec = n / 4;   // DiT-XL/2-sized, multiple of the 16 KB tile
  float* g;
  double* out;
  cudaMalloc(&g, n * sizeof(float));
  cudaMalloc(&out, sizeof(double));
  fill<<<148 * 8, 256>>>(g, n);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  double ref = 0;
  auto report = [&](const char* name, float ms) {
    double got;
    cudaMemcpy(&got, out, sizeof(double), cudaMemcpyDeviceToHost);
    if (ref == 0) ref = got;
    printf("{\"variant\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f, \"sum_matches\": %s}\n", name, ms, n * 4.0 / (ms * 1e-3) / 1e9,
           (got > 0 && fabs(got - ref) <= 1e-4 * ref) ? "true" : "false");
    fflush(stdout);
  };
#define RUN(NAME, ...)                                   \
  do {                                                   \
    float ms = time_ms([&] { cudaMemsetAsync(out, 0, 8); __VA_ARGS__; }, 9); \
    cudaError_t e = cudaDeviceSynchronize();             \
    if (e != cudaSuccess) { printf("{\"variant\": \"%s\", \"error\": \"%s\"}\n", NAME, cudaGetErrorString(e)); return 1; } \
    report(NAME, ms);                                    \
  } while (0)
  RUN("direct ld.global.cs.v4, 256 thr x4, 8 CTAs/SM (product geometry)", (sumsq_direct<256, 4, 8><<<sms * 8 * 16, 256>>>(g4, nvec, out)));
  RUN("direct ld.global.cs.v4, 256 thr x4, persistent 8 CTAs/SM", (sumsq_direct<256, 4, 8><<<sms * 8, 256>>>(g4, nvec, out)));
  {
    constexpr int ST = 4, TF4 = 1024;                 // 4 stages x 16 KB = 64 KB per CTA
    auto k = sumsq_tma<256, ST, TF4, 3>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * TF4 * 16);
    RUN("TMA bulk copy, 4 x 16 KB ring, 256 thr, 3 CTAs/SM", (k<<<sms * 3, 256, ST * TF4 * 16>>>(g4, nvec, out)));
  }
  {
    constexpr int ST = 8, TF4 = 1024;                 // 8 stages x 16 KB = 128 KB per CTA
    auto k = sumsq_tma<256, ST, TF4, 1>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * TF4 * 16);
    RUN("TMA bulk copy, 8 x 16 KB ring, 256 thr, 1 CTA/SM", (k<<<sms, 256, ST * TF4 * 16>>>(g4, nvec, out)));
  }
  {
    constexpr int ST = 3, TF4 = 2048;                 // 3 stages x 32 KB = 96 KB per CTA
    auto k = sumsq_tma<512, ST, TF4, 2>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * TF4 * 16);
    RUN("TMA bulk copy, 3 x 32 KB ring, 512 thr, 2 CTAs/SM", (k<<<sms * 2, 512, ST * TF4 * 16>>>(g4, nvec, out)));
  }
  return 0;
}
