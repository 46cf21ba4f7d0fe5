// tune_stream.cu — parameter sweep for the streaming kernel shapes of the hot path (not product code).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tune_stream tools/tune/tune_stream.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

enum Hint { CS = 0, NC = 1, PLAIN = 2 };

template <int H> __device__ __forceinline__ float4 ld(const float4* p) {
  if (H == CS) return __ldcs(p);
  if (H == NC) { float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r; }
  return *p;
}
template <int H> __device__ __forceinline__ void st(float4* p, float4 v) {
  if (H == CS) __stcs(p, v); else if (H == NC) __stwt(p, v); else *p = v;
}

// K1 shape: acc = acc + g*g/L   (2 reads, 1 write)
template <int THREADS, int UNROLL, int CTAS, int H>
__global__ void __launch_bounds__(THREADS, CTAS) k1(float* __restrict__ acc, const float* __restrict__ g, int64_t nvec, float L) {
  float4* a4 = (float4*)acc; const float4* g4 = (const float4*)g;
  const int64_t tile = (int64_t)THREADS * UNROLL, ntiles = (nvec + tile - 1) / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 a[UNROLL], x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; if (v < nvec) { a[u] = ld<H>(a4 + v); x[u] = ld<H>(g4 + v); } }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; if (v < nvec) {
      a[u].x += __fdiv_rn(x[u].x * x[u].x, L); a[u].y += __fdiv_rn(x[u].y * x[u].y, L);
      a[u].z += __fdiv_rn(x[u].z * x[u].z, L); a[u].w += __fdiv_rn(x[u].w * x[u].w, L); st<H>(a4 + v, a[u]); } }
  }
}

// K3 shape: 5 reads (g,p,m,v,e) 4 writes
template <int THREADS, int UNROLL, int CTAS, int H>
__global__ void __launch_bounds__(THREADS, CTAS) k3(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* __restrict__ e, int64_t nvec) {
  float4 *p4 = (float4*)p, *m4 = (float4*)m, *v4 = (float4*)v, *e4 = (float4*)e; const float4* g4 = (const float4*)g;
  const int64_t tile = (int64_t)THREADS * UNROLL, ntiles = (nvec + tile - 1) / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 G[UNROLL], P[UNROLL], M[UNROLL], V[UNROLL], E[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t i = base + (int64_t)u * THREADS; if (i < nvec) { G[u] = ld<H>(g4 + i); P[u] = ld<H>(p4 + i); M[u] = ld<H>(m4 + i); V[u] = ld<H>(v4 + i); E[u] = ld<H>(e4 + i); } }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t i = base + (int64_t)u * THREADS; if (i < nvec) {
#define UPD(c) { float gg = G[u].c; M[u].c = fmaf(gg - M[u].c, 0.1f, M[u].c); V[u].c = fmaf(0.001f * gg, gg, V[u].c * 0.999f); \
                 P[u].c = P[u].c + (-1e-4f * M[u].c) / (sqrtf(V[u].c) / 0.03f + 1e-8f); E[u].c = fmaf(P[u].c, 1e-4f, E[u].c * 0.9999f); }
      UPD(x) UPD(y) UPD(z) UPD(w)
      st<H>(p4 + i, P[u]); st<H>(m4 + i, M[u]); st<H>(v4 + i, V[u]); st<H>(e4 + i, E[u]); } }
  }
}

// read-only shape (clip norm / histogram passes): 1 stream read, register reduce
template <int THREADS, int UNROLL, int CTAS, int H>
__global__ void __launch_bounds__(THREADS, CTAS) kr(const float* __restrict__ g, int64_t nvec, unsigned* __restrict__ out) {
  const float4* g4 = (const float4*)g; unsigned acc = 0;
  const int64_t tile = (int64_t)THREADS * UNROLL, ntiles = (nvec + tile - 1) / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x; float4 x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; x[u] = v < nvec ? ld<H>(g4 + v) : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += (__float_as_uint(x[u].x) >> 16) ^ (__float_as_uint(x[u].y) >> 16) ^ (__float_as_uint(x[u].z) >> 16) ^ (__float_as_uint(x[u].w) >> 16);
  }
  if (acc == 0xdeadbeefu) out[0] = acc;
}
// apply shape: read 1 stream, write 1 byte per element
template <int THREADS, int UNROLL, int CTAS, int H>
__global__ void __launch_bounds__(THREADS, CTAS) ka(const float* __restrict__ g, unsigned* __restrict__ mask, int64_t nvec, unsigned thr) {
  const float4* g4 = (const float4*)g;
  const int64_t tile = (int64_t)THREADS * UNROLL, ntiles = (nvec + tile - 1) / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x; float4 x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; x[u] = v < nvec ? ld<H>(g4 + v) : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { int64_t v = base + (int64_t)u * THREADS; if (v < nvec) {
      unsigned m = (__float_as_uint(x[u].x) > thr) | ((__float_as_uint(x[u].y) > thr) << 8) | ((__float_as_uint(x[u].z) > thr) << 16) | ((__float_as_uint(x[u].w) > thr) << 24);
      mask[v] = m; } }
  }
}

static float* dalloc(int64_t n) { float* p; cudaMalloc(&p, n * 4); cudaMemset(p, 0, n * 4); return p; }
static char* flushbuf; 
template <typename F> float timeit(F f) {
  std::vector<float> t; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 7; ++i) { cudaMemsetAsync(flushbuf, i, 256 << 20); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (i >= 2) t.push_back(ms); }
  std::sort(t.begin(), t.end()); return t[t.size() / 2];
}

int main() {
  const int64_t n = 675129632, nvec = n / 4; const int sms = 148;
  float *p = dalloc(n), *g = dalloc(n), *m = dalloc(n), *v = dalloc(n), *e = dalloc(n);
  cudaMalloc(&flushbuf, 256 << 20);
#define RUN1(T, U, C, H, MULT) { int64_t gg = (int64_t)sms * C * MULT, nt = (nvec + (int64_t)T * U - 1) / ((int64_t)T * U); int grid = (int)(gg < nt ? gg : nt); float ms = timeit([&] { k1<T, U, C, H><<<grid, T>>>(p, g, nvec, 2000.f); }); \
    printf("k1 threads=%d unroll=%d ctas=%d hint=%d gridmult=%d  %.4f ms  %.1f GB/s\n", T, U, C, H, MULT, ms, 12.0 * n / ms / 1e6); }
#define RUN3(T, U, C, H, MULT) { int64_t gg = (int64_t)sms * C * MULT, nt = (nvec + (int64_t)T * U - 1) / ((int64_t)T * U); int grid = (int)(gg < nt ? gg : nt); float ms = timeit([&] { k3<T, U, C, H><<<grid, T>>>(p, g, m, v, e, nvec); }); \
    printf("k3 threads=%d unroll=%d ctas=%d hint=%d gridmult=%d  %.4f ms  %.1f GB/s\n", T, U, C, H, MULT, ms, 36.0 * n / ms / 1e6); }
  unsigned* outp; cudaMalloc(&outp, 64);
#define RUNR(T, U, C, H, MULT) { int64_t gg = (int64_t)sms * C * MULT, nt = (nvec + (int64_t)T * U - 1) / ((int64_t)T * U); int grid = (int)(gg < nt ? gg : nt); float ms = timeit([&] { kr<T, U, C, H><<<grid, T>>>(g, nvec, outp); }); \
    printf("kr threads=%d unroll=%d ctas=%d hint=%d gridmult=%d  %.4f ms  %.1f GB/s\n", T, U, C, H, MULT, ms, 4.0 * n / ms / 1e6); }
#define RUNA(T, U, C, H, MULT) { int64_t gg = (int64_t)sms * C * MULT, nt = (nvec + (int64_t)T * U - 1) / ((int64_t)T * U); int grid = (int)(gg < nt ? gg : nt); float ms = timeit([&] { ka<T, U, C, H><<<grid, T>>>(g, (unsigned*)m, nvec, 0x3c000000u); }); \
    printf("ka threads=%d unroll=%d ctas=%d hint=%d gridmult=%d  %.4f ms  %.1f GB/s\n", T, U, C, H, MULT, ms, 5.0 * n / ms / 1e6); }
  RUNR(256, 8, 4, CS, 1) RUNR(256, 8, 4, PLAIN, 1) RUNR(256, 4, 8, CS, 1) RUNR(256, 4, 8, CS, 8) RUNR(256, 4, 8, CS, 32) RUNR(128, 4, 16, CS, 32)
  RUNR(128, 2, 16, CS, 32) RUNR(128, 8, 8, CS, 32) RUNR(256, 4, 8, NC, 32) RUNR(256, 4, 8, PLAIN, 32) RUNR(1024, 4, 1, CS, 1) RUNR(1024, 4, 2, CS, 1) RUNR(128, 4, 16, CS, 100000) RUNR(256, 8, 4, CS, 32)
  RUNA(256, 4, 4, CS, 1) RUNA(256, 4, 4, PLAIN, 1) RUNA(256, 8, 4, CS, 2) RUNA(128, 2, 16, CS, 32) RUNA(128, 2, 16, PLAIN, 32) RUNA(128, 4, 16, CS, 32)
  RUNA(128, 4, 16, PLAIN, 32) RUNA(256, 4, 8, CS, 32) RUNA(128, 2, 16, CS, 100000) RUNA(128, 4, 8, CS, 100000) RUNA(256, 8, 4, CS, 32) RUNA(128, 8, 8, CS, 32)
  // torch-style copy for reference (read + write)
  { float ms = timeit([&] { cudaMemcpyAsync(p, g, n * 4, cudaMemcpyDeviceToDevice); }); printf("memcpy d2d %.4f ms %.1f GB/s (read+write)\n", ms, 8.0 * n / ms / 1e6); }
  return 0;
}
