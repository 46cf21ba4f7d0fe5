// tune_hist.cu — histogram pass 0 variants (not product code): BITS-bit shared-memory histogram with COPIES
// replicated sub-histograms (lane % COPIES) to cut bank conflicts; Gaussian keys.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
__device__ __forceinline__ uint32_t key_of(float x) { uint32_t b = __float_as_uint(x) & 0x7fffffffu; return b > 0x7f800000u ? 0u : b + 1u; }
__global__ void fill(float* g, int64_t n, unsigned long long seed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; curandStatePhilox4_32_10_t st; curand_init(seed, i, 0, &st);
  for (int64_t j = i * 4; j < n; j += (int64_t)gridDim.x * blockDim.x * 4) { float4 r = curand_normal4(&st); if (j + 3 < n) { g[j] = r.x * 1e-2f; g[j+1] = r.y * 1e-2f; g[j+2] = r.z * 1e-2f; g[j+3] = r.w * 1e-2f; } }
}
template <int BITS, int COPIES, int THREADS, int UNROLL, int RUNCACHE>
__global__ void __launch_bounds__(THREADS, 1) hist(const float* __restrict__ a, int64_t n, unsigned long long* __restrict__ bins) {
  extern __shared__ unsigned int h[];
  constexpr int NB = 1 << BITS; constexpr int SHIFT = 31 - BITS;
  for (int i = threadIdx.x; i < NB * COPIES; i += THREADS) h[i] = 0;
  __syncthreads();
  unsigned int* mine = h + (threadIdx.x % COPIES) * NB;
  const int64_t nvec = n >> 2, tile = (int64_t)THREADS * UNROLL, ntiles = (nvec + tile - 1) / tile;
  const float4* a4 = (const float4*)a; uint32_t cb = 0xffffffffu, cc = 0;
  auto push = [&](uint32_t b) { if (RUNCACHE == 1) { if (b == cb) ++cc; else { if (cc) atomicAdd(mine + cb, cc); cb = b; cc = 1; } } else atomicAdd(mine + b, 1u); };
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x; float4 x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; x[u] = v < nvec ? __ldcs(a4 + v) : make_float4(0, 0, 0, 0); }
#pragma unroll
    if (RUNCACHE == 3) {   // whole thread-tile: one atomic if all 4*UNROLL bins agree (full tiles only)
      uint32_t b[UNROLL * 4]; bool same = true;
      for (int u = 0; u < UNROLL; ++u) { b[4*u] = key_of(x[u].x) >> SHIFT; b[4*u+1] = key_of(x[u].y) >> SHIFT; b[4*u+2] = key_of(x[u].z) >> SHIFT; b[4*u+3] = key_of(x[u].w) >> SHIFT; }
      for (int j = 1; j < UNROLL * 4; ++j) same &= b[j] == b[0];
      bool full = base + (int64_t)(UNROLL - 1) * THREADS < nvec;
      if (same && full) atomicAdd(mine + b[0], (unsigned)(UNROLL * 4));
      else for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; if (v < nvec) for (int q = 0; q < 4; ++q) atomicAdd(mine + b[4*u+q], 1u); }
    } else if (RUNCACHE == 2) {  // per float4
      for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; if (v < nvec) {
        uint32_t b0 = key_of(x[u].x) >> SHIFT, b1 = key_of(x[u].y) >> SHIFT, b2 = key_of(x[u].z) >> SHIFT, b3 = key_of(x[u].w) >> SHIFT;
        if (b0 == b1 && b1 == b2 && b2 == b3) atomicAdd(mine + b0, 4u); else { atomicAdd(mine + b0, 1u); atomicAdd(mine + b1, 1u); atomicAdd(mine + b2, 1u); atomicAdd(mine + b3, 1u); } } }
    } else
    for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; if (v < nvec) { push(key_of(x[u].x) >> SHIFT); push(key_of(x[u].y) >> SHIFT); push(key_of(x[u].z) >> SHIFT); push(key_of(x[u].w) >> SHIFT); } }
  }
  if (RUNCACHE == 1 && cc) atomicAdd(mine + cb, cc);
  __syncthreads();
  for (int i = threadIdx.x; i < NB; i += THREADS) { unsigned int c = 0; for (int k = 0; k < COPIES; ++k) c += h[k * NB + i]; if (c) atomicAdd(bins + i, (unsigned long long)c); }
}

// ---- round 2: load hint and register double buffer for the product form (plain shared atomics, 15 bits, 1024 threads) ----
template <int THREADS, int UNROLL, bool PLAIN, bool PREFETCH>
__global__ void __launch_bounds__(THREADS, 1) hist_p(const float* __restrict__ a, int64_t n, unsigned long long* __restrict__ bins) {
  extern __shared__ unsigned int h[];
  constexpr int NB = 1 << 15;
  for (int i = threadIdx.x; i < NB; i += THREADS) h[i] = 0;
  __syncthreads();
  const int64_t nvec = n >> 2, tile = (int64_t)THREADS * UNROLL, ntiles = (nvec + tile - 1) / tile;
  const float4* a4 = (const float4*)a;
  float4 x[UNROLL], nx[UNROLL];
  auto load = [&](float4* d, int64_t t) {
    const int64_t base = t * tile + threadIdx.x;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; d[u] = (t < ntiles && v < nvec) ? (PLAIN ? a4[v] : __ldcs(a4 + v)) : make_float4(0, 0, 0, 0); }
  };
  if (PREFETCH) load(nx, blockIdx.x);
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    if (PREFETCH) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) x[u] = nx[u];
      load(nx, t + gridDim.x);
    } else load(x, t);
    const int64_t base = t * tile + threadIdx.x;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; if (v < nvec) {
      atomicAdd(h + (key_of(x[u].x) >> 16), 1u); atomicAdd(h + (key_of(x[u].y) >> 16), 1u); atomicAdd(h + (key_of(x[u].z) >> 16), 1u); atomicAdd(h + (key_of(x[u].w) >> 16), 1u); } }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NB; i += THREADS) { unsigned int c = h[i]; if (c) atomicAdd(bins + i, (unsigned long long)c); }
}
static char* flushbuf;
template <typename F> float timeit(F f) { std::vector<float> t; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 7; ++i) { cudaMemsetAsync(flushbuf, i, 256 << 20); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (i >= 2) t.push_back(ms); }
  std::sort(t.begin(), t.end()); return t[t.size() / 2]; }
int main() {
  const int64_t n = 675129632; float* g; cudaMalloc(&g, n * 4); unsigned long long* bins; cudaMalloc(&bins, 8 << 20); cudaMalloc(&flushbuf, 256 << 20);
  fill<<<148 * 8, 256>>>(g, n, 1234); cudaDeviceSynchronize();
#define RUN(BITS, COPIES, T, U, RC, CTAS) { int smem = (1 << BITS) * COPIES * 4; cudaFuncSetAttribute(hist<BITS, COPIES, T, U, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    float ms = timeit([&] { cudaMemsetAsync(bins, 0, 8 << 20); hist<BITS, COPIES, T, U, RC><<<148 * CTAS, T, smem>>>(g, n, bins); }); \
    cudaError_t e = cudaGetLastError(); printf("bits=%d copies=%d threads=%d unroll=%d runcache=%d ctas/sm=%d  %.4f ms  %.1f GB/s %s\n", BITS, COPIES, T, U, RC, CTAS, ms, 4.0 * n / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e)); }
  for (int pass = 0; pass < 3; ++pass) {
    if (pass == 1) { cudaMemset(g, 0, n * 4); printf("--- all zeros\n"); }
    if (pass == 2) { fill<<<148 * 8, 256>>>(g, n, 99); cudaDeviceSynchronize(); cudaMemset(g, 0, (n / 10 * 9) * 4); printf("--- 90%% zeros then gaussian\n"); }
    RUN(15, 1, 1024, 4, 0, 1) RUN(15, 1, 1024, 4, 1, 1) RUN(15, 1, 1024, 4, 2, 1) RUN(15, 1, 1024, 4, 3, 1)
#define RUNP(T, U, PL, PF) { int smem = (1 << 15) * 4; cudaFuncSetAttribute(hist_p<T, U, PL, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    float ms = timeit([&] { cudaMemsetAsync(bins, 0, 8 << 20); hist_p<T, U, PL, PF><<<148, T, smem>>>(g, n, bins); }); \
    cudaError_t e = cudaGetLastError(); printf("hist_p threads=%d unroll=%d plain=%d prefetch=%d  %.4f ms  %.1f GB/s %s\n", T, U, (int)PL, (int)PF, ms, 4.0 * n / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e)); }
    RUNP(1024, 4, false, false) RUNP(1024, 4, true, false) RUNP(1024, 4, true, true) RUNP(1024, 2, true, true) RUNP(1024, 8, true, false) RUNP(512, 4, true, true) RUNP(512, 8, true, true)
  }
  return 0;
}
