// select.cu — K2b: exact global top-k saliency selection by a two-pass radix /
// histogram select over order-preserving keys, plus the threshold apply.
//
// Reference (two full CPU argsorts of an N-element vector, 16 B/elem of index temporaries):
//   all = -cat(|sum_batches g|); ranks = argsort(argsort(all)); mask = ranks < int(N*ratio)
//     DDPM/runners/diffusion.py:1009-1034, Classification/unlearn/salun.py:170-193
//   thr = -topk(-|theta - theta0|, k)[0][-1]      SD/train-scripts/proximal_gradient.py:161-165
//
// key(x) = 0 for NaN else bits(|x|)+1   (31 significant bits; monotone in |x|)
//   pass 0: histogram of key[30:16] (32768 bins) in SHARED memory, one CTA per SM
//   scan 0: largest bin B with  #(bin > B) < k  -> prefix, rank wanted inside B
//   pass 1: histogram of key[15:0] for keys whose [30:16] == prefix (65536 global bins;
//           only the few matching elements issue an atomic)
//   scan 1: threshold key, #greater, #equal, tie budget
//   apply : mask = key > thr  ||  (key == thr && index-ordered tie rank < budget)
//           With sfr_select_hist1_mask, pass 1 already wrote mask = key[30:16] > prefix for every
//           element and staged the few keys that share the prefix; apply then only walks that list
//           (and re-reads the handful of chunks holding an ordered tie): 4 + 4 + 1 B/elem in all.
// Between hist and scan a multi-GPU caller all-reduces `bins` (<= 512 KB) over NCCL;
// everything else is shard-local.  All steps are stream-ordered; the host never has to
// read anything back.
#include <atomic>

#include <cooperative_groups.h>

#include "common.cuh"

namespace sfr {
namespace {

namespace cg = cooperative_groups;

// ---- key source ------------------------------------------------------------------------
// The float whose magnitude is ranked.
template <int MODE>
__device__ __forceinline__ float key_value(float a, float b, float eps) {
  if constexpr (MODE == SFR_KEY_ABS) {
    return a;
  } else if constexpr (MODE == SFR_KEY_RATIO) {
    return __fdiv_rn(__fadd_rn(a, eps), __fadd_rn(b, eps));
  } else {  // SFR_KEY_ABSDIFF: |a - b|  (proximal_gradient.py:158-159: params -= init ; abs_())
    return __fsub_rn(a, b);
  }
}

template <int MODE>
__device__ __forceinline__ uint32_t key_from(float a, float b, float eps) {
  return select_key(key_value<MODE>(a, b, eps));
}

// Per-thread run-length cache in front of the GLOBAL-memory histogram of pass 1: consecutive items of
// one thread that fall in the same bin cost one atomic.  Dead units give long runs of exact zeros (a
// constructor-initialised DiT has 99.96 % zero gradients), which would otherwise
// serialise on one L2 counter.  Pass 0 does NOT use it: shared-memory atomics to one address run at
// full rate on B200 (all-zero input: 0.423 ms vs 0.425 ms Gaussian at n = 675 M), and the cache's
// compare-and-branch per element costs 12 % on ordinary data (tools/tune/tune_hist.cu).
template <typename Counter>
struct RunCache {
  uint32_t bin = 0xffffffffu;
  uint32_t count = 0;
  __device__ __forceinline__ void push(Counter* hist, uint32_t b) {
    if (b == bin) {
      ++count;
    } else {
      if (count) atomicAdd(hist + bin, (Counter)count);
      bin = b;
      count = 1;
    }
  }
  __device__ __forceinline__ void flush(Counter* hist) {
    if (count) atomicAdd(hist + bin, (Counter)count);
    count = 0;
    bin = 0xffffffffu;
  }
};

// ---- pass 0: 15-bit shared-memory histogram (plain shared-memory atomics) ------------------------
constexpr int kHistThreads = 1024;
constexpr int kHistUnroll = 4;

template <int MODE>
__global__ void __launch_bounds__(kHistThreads, 1)
select_hist0_kernel(const float* __restrict__ a, const float* __restrict__ b, float eps, int64_t n,
                    unsigned long long* __restrict__ bins) {
  extern __shared__ unsigned int hist[];  // SFR_SELECT_BINS0 counters, 128 KB
  for (int i = threadIdx.x; i < SFR_SELECT_BINS0; i += kHistThreads) hist[i] = 0;
  __syncthreads();

  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kHistThreads * kHistUnroll;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);

  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 x[kHistUnroll], y[kHistUnroll];
#pragma unroll
    for (int u = 0; u < kHistUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kHistThreads;
      const bool in = v < nvec;
      x[u] = in ? ld_once(a4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (MODE != SFR_KEY_ABS)
        y[u] = in ? ld_once(b4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      else
        y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kHistUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kHistThreads;
      if (v >= nvec) continue;
      atomicAdd(hist + (key_from<MODE>(x[u].x, y[u].x, eps) >> 16), 1u);
      atomicAdd(hist + (key_from<MODE>(x[u].y, y[u].y, eps) >> 16), 1u);
      atomicAdd(hist + (key_from<MODE>(x[u].z, y[u].z, eps) >> 16), 1u);
      atomicAdd(hist + (key_from<MODE>(x[u].w, y[u].w, eps) >> 16), 1u);
    }
  }
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    atomicAdd(hist + (key_from<MODE>(a[i], MODE != SFR_KEY_ABS ? b[i] : 0.f, eps) >> 16), 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SFR_SELECT_BINS0; i += kHistThreads) {
    const unsigned int c = hist[i];
    if (c) atomicAdd(bins + i, (unsigned long long)c);
  }
}

// ---- scratch layout (unsigned long long words) ------------------------------------------------
//   [0, nchunks)                  per-chunk count of threshold-equal keys
//   [nchunks, 2*nchunks)          per scan block (kScanBlock chunks): exclusive tie base (+ tie_base);
//                                 only the first ceil(nchunks / kScanBlock) words are used
//   header (kCandHeader words):   [0] overflow flag, [1] number of regions, [2] region capacity,
//                                 [3] 1 if pass 1 also wrote the provisional mask, [4] that mask's address,
//                                 [5] number of chunks listed as holding a tie, [6] 1 if that list overflowed
//                                 (the list lives in the unused tail of the second scratch region),
//                                 [7] 1 once an ordered apply has walked the list (its counters then need clearing),
//                                 [8 + r] entries staged in region r
//   candidate regions             cand_cap words: (flat index << 16) | key[15:0]
constexpr int64_t kChunkElems = 8192;       // == kChunk below (kApplyThreads * 4 * kChunkVecs)
constexpr int kMaxRegions = 2048;           // >= CTAs of the pass-1 grid
constexpr int kCandHeader = 8 + kMaxRegions;

__host__ __device__ inline int64_t scratch_nchunks(int64_t n) { return (n + kChunkElems - 1) / kChunkElems; }
__host__ __device__ inline int64_t scratch_cand_cap(int64_t n) {
  const int64_t byfrac = n / 32;            // room for 3 % of the elements sharing the 15-bit prefix
  const int64_t floor_ = (int64_t)kMaxRegions * 64;
  return byfrac > floor_ ? byfrac : floor_;
}

// ---- pass 1: filtered 16-bit histogram straight into global bins -----------------------------
// Keys whose bits [30:16] equal the chosen prefix (typically < 1 % of the elements) are also
// STAGED — (flat index, low 16 key bits) — in a region private to the CTA (one shared-memory
// counter, no global atomics).  The ordered-tie apply later counts threshold-equal keys per chunk
// from this short list instead of re-reading the whole vector; if any region overflows (degenerate
// inputs: most keys equal) a flag makes the apply fall back to the streaming count.
constexpr int kFiltThreads = 256;
constexpr int kFiltCtasPerSm = 4;     // 64 registers per thread: room for the double buffer
constexpr int kFiltUnroll = 4;

struct CandStage {
  unsigned long long* region;
  unsigned int cap;
  unsigned int* s_count;
  unsigned int* s_over;
  __device__ __forceinline__ void push(int64_t idx, uint32_t key) {
    if (*reinterpret_cast<volatile unsigned int*>(s_over)) return;
    const unsigned int pos = atomicAdd(s_count, 1u);
    if (pos < cap) region[pos] = ((unsigned long long)idx << 16) | (key & 0xffffu);
    else *reinterpret_cast<volatile unsigned int*>(s_over) = 1u;
  }
};

// The vectorised body of pass 1.  A key is  bits(|x|) + 1  (0 for NaN), so on non-negative floats key order IS float order
// and both tests of the hot loop are one FSETP on the value itself, with no key arithmetic:
//     key[30:16] >  prefix   <=>   |x| >= as_float(((prefix + 1) << 16) - 1)
//     key[30:16] == prefix   <=>   |x| >= as_float((prefix << 16) - 1)  and not the above          (prefix >= 1)
// (NaN fails every ordered compare, as its key 0 fails both tests; a bound that is itself a NaN pattern — prefix at the top
// of the range — makes its test false for every value, which is also what the keys say.  Compares are IEEE: the library is
// built without flush-to-zero, so the subnormal bounds of small prefixes order correctly.)  prefix == 0 — the threshold
// sits among zeros / subnormals, and NaN keys belong to the bin — takes the integer form (FAST = false).  Keys are built
// only for the few elements that match the prefix; one branch per float4 guards that path.
template <int MODE, bool WRITE, bool FAST>
__device__ __forceinline__ void hist1_tiles(const float* __restrict__ a, const float* __restrict__ b, float eps,
                                            int64_t nvec, uint32_t prefix, unsigned long long* __restrict__ bins,
                                            uint8_t* __restrict__ mask, RunCache<unsigned long long>& rc, CandStage& cs) {
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  const float t_hi = __uint_as_float(((prefix + 1u) << 16) - 1u);
  const float t_lo = __uint_as_float((prefix << 16) - 1u);

  // Register double buffer: the loads of this CTA's NEXT tile are in flight while the current one is tested, so the
  // bytes outstanding per SM never drop to zero between tiles (tools/tune/tune_hist1.cu, variants F / P: +7 %).
  constexpr int U = MODE == SFR_KEY_ABS ? kFiltUnroll : kFiltUnroll / 2;   // two inputs: half the tile, same bytes in flight
  const int64_t tile = (int64_t)kFiltThreads * U;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  float4 x[U], y[U], nx[U], ny[U];
  auto load_tile = [&](float4* dx, float4* dy, int64_t t) {
    const int64_t base = t * tile + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + (int64_t)u * kFiltThreads;
      const bool in = t < ntiles && v < nvec;
      dx[u] = in ? ld_once(a4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (MODE != SFR_KEY_ABS)
        dy[u] = in ? ld_once(b4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      else
        dy[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load_tile(nx, ny, blockIdx.x);
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      x[u] = nx[u];
      y[u] = ny[u];
    }
    load_tile(nx, ny, t + gridDim.x);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + (int64_t)u * kFiltThreads;
      if (v >= nvec) continue;
      if constexpr (FAST) {
        const float r0 = fabsf(key_value<MODE>(x[u].x, y[u].x, eps)), r1 = fabsf(key_value<MODE>(x[u].y, y[u].y, eps));
        const float r2 = fabsf(key_value<MODE>(x[u].z, y[u].z, eps)), r3 = fabsf(key_value<MODE>(x[u].w, y[u].w, eps));
        const bool g0 = r0 >= t_hi, g1 = r1 >= t_hi, g2 = r2 >= t_hi, g3 = r3 >= t_hi;
        const bool e0 = r0 >= t_lo && !g0, e1 = r1 >= t_lo && !g1, e2 = r2 >= t_lo && !g2, e3 = r3 >= t_lo && !g3;
        if constexpr (WRITE)
          reinterpret_cast<unsigned int*>(mask)[v] = (g0 ? 1u : 0u) | (g1 ? 0x100u : 0u) | (g2 ? 0x10000u : 0u) | (g3 ? 0x1000000u : 0u);
        if (e0 | e1 | e2 | e3) {     // a matching value is finite or +inf, never NaN: its key is bits + 1
          if (e0) { const uint32_t k = __float_as_uint(r0) + 1u; rc.push(bins, k & 0xffffu); cs.push(v * 4 + 0, k); }
          if (e1) { const uint32_t k = __float_as_uint(r1) + 1u; rc.push(bins, k & 0xffffu); cs.push(v * 4 + 1, k); }
          if (e2) { const uint32_t k = __float_as_uint(r2) + 1u; rc.push(bins, k & 0xffffu); cs.push(v * 4 + 2, k); }
          if (e3) { const uint32_t k = __float_as_uint(r3) + 1u; rc.push(bins, k & 0xffffu); cs.push(v * 4 + 3, k); }
        }
      } else {
        const uint32_t k0 = key_from<MODE>(x[u].x, y[u].x, eps);
        const uint32_t k1 = key_from<MODE>(x[u].y, y[u].y, eps);
        const uint32_t k2 = key_from<MODE>(x[u].z, y[u].z, eps);
        const uint32_t k3 = key_from<MODE>(x[u].w, y[u].w, eps);
        if ((k0 >> 16) == prefix) { rc.push(bins, k0 & 0xffffu); cs.push(v * 4 + 0, k0); }
        if ((k1 >> 16) == prefix) { rc.push(bins, k1 & 0xffffu); cs.push(v * 4 + 1, k1); }
        if ((k2 >> 16) == prefix) { rc.push(bins, k2 & 0xffffu); cs.push(v * 4 + 2, k2); }
        if ((k3 >> 16) == prefix) { rc.push(bins, k3 & 0xffffu); cs.push(v * 4 + 3, k3); }
        if constexpr (WRITE) {
          const uint32_t s0 = (k0 >> 16) > prefix, s1 = (k1 >> 16) > prefix;
          const uint32_t s2 = (k2 >> 16) > prefix, s3 = (k3 >> 16) > prefix;
          reinterpret_cast<unsigned int*>(mask)[v] = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
        }
      }
    }
  }
}

// With WRITE the pass also leaves a PROVISIONAL mask: 1 where key[30:16] > prefix, 0 elsewhere — final
// for every element except the staged ones, which sfr_select_apply then resolves from the candidate
// list alone (no third pass over the vector).
template <int MODE, bool WRITE>
__global__ void __launch_bounds__(kFiltThreads, kFiltCtasPerSm)
select_hist1_kernel(const float* __restrict__ a, const float* __restrict__ b, float eps, int64_t n,
                    const sfr_select_state* __restrict__ state,
                    unsigned long long* __restrict__ bins, unsigned long long* __restrict__ scratch,
                    uint8_t* __restrict__ mask) {
  if (state->select_none || state->select_all) return;
  __shared__ unsigned int s_count, s_over;
  if (threadIdx.x == 0) {
    s_count = 0;
    s_over = 0;
  }
  __syncthreads();
  unsigned long long* hdr = scratch + 2 * scratch_nchunks(n);
  CandStage cs;
  cs.cap = (unsigned int)(scratch_cand_cap(n) / gridDim.x);
  cs.region = hdr + kCandHeader + (int64_t)blockIdx.x * cs.cap;
  cs.s_count = &s_count;
  cs.s_over = &s_over;

  const uint32_t prefix = state->prefix;
  const int64_t nvec = n >> 2;
  RunCache<unsigned long long> rc;
  if (prefix != 0u)
    hist1_tiles<MODE, WRITE, true>(a, b, eps, nvec, prefix, bins, mask, rc, cs);
  else
    hist1_tiles<MODE, WRITE, false>(a, b, eps, nvec, prefix, bins, mask, rc, cs);
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    const uint32_t k = key_from<MODE>(a[i], MODE != SFR_KEY_ABS ? b[i] : 0.f, eps);
    if ((k >> 16) == prefix) { rc.push(bins, k & 0xffffu); cs.push(i, k); }
    if constexpr (WRITE) mask[i] = (uint8_t)((k >> 16) > prefix);
  }
  rc.flush(bins);
  __syncthreads();
  if (threadIdx.x == 0) {
    hdr[8 + blockIdx.x] = s_count < cs.cap ? s_count : cs.cap;
    if (s_over) hdr[0] = 1ull;
    if (blockIdx.x == 0) {
      hdr[1] = gridDim.x;
      hdr[2] = cs.cap;
      if constexpr (WRITE) {
        hdr[3] = 1ull;
        hdr[4] = (unsigned long long)(uintptr_t)mask;
      }
    }
  }
}

// ---- scans (multi-CTA, last CTA finishes) ------------------------------------------------------
// Finds the largest bin B with  above(B) < want <= above(B) + bins[B]  (scanning from the top).
// A single-CTA scan of 32 k / 64 k counters is a 29 / 48 us latency chain on one SM; here one CTA per
// 1024 bins forms a partial sum, and the last CTA to arrive (ticket counter, __threadfence) scans the
// <= 64 partials, then the 1024 bins of the block that holds the crossing.  Partials and the ticket
// live in the tail of the bins buffer: bins[SFR_SELECT_BINS1 .. SFR_SELECT_BINS_ALLOC).
constexpr int kScanThreads = 256;
constexpr int kScanCtaBins = 1024;                       // bins per CTA = 4 per thread
constexpr int kScanMaxCtas = SFR_SELECT_BINS1 / kScanCtaBins;   // 64
static_assert(SFR_SELECT_BINS_ALLOC >= SFR_SELECT_BINS1 + kScanMaxCtas + 1, "bins tail too small");

// Block-wide search inside `count` descending values held one per thread-slot: thread t owns
// values v[0..PER) = positions t*PER .. t*PER+PER-1 of the descending order.  Returns through shared
// memory the position (0-based from the top) of the crossing and the sum above it; -1 if want > total.
template <int PER>
__device__ void find_crossing(const unsigned long long (&v)[PER], unsigned long long want,
                              unsigned long long carry_above, int* out_pos, unsigned long long* out_above) {
  __shared__ unsigned long long warp_tot[kScanThreads / 32];
  __shared__ int s_pos;
  __shared__ unsigned long long s_above;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long mine = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) mine += v[j];
  if (threadIdx.x == 0) {
    s_pos = -1;
    s_above = 0;
  }
  unsigned long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  unsigned long long before = carry_above;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w)
    if (w < warp) before += warp_tot[w];
  const unsigned long long excl = before + (incl - mine);
  if (want > excl && want <= excl + mine) {  // exactly one thread
    unsigned long long above = excl;
    int found = -1;
    unsigned long long found_above = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (found < 0 && want <= above + v[j]) {
        found = (int)threadIdx.x * PER + j;
        found_above = above;
      }
      above += v[j];
    }
    s_pos = found;
    s_above = found_above;
  }
  __syncthreads();
  *out_pos = s_pos;
  *out_above = s_above;
  __syncthreads();
}

__global__ void __launch_bounds__(kScanThreads, 1)
select_scan_kernel(int pass, sfr_select_state* __restrict__ state,
                   unsigned long long* __restrict__ bins) {
  constexpr int kPer = kScanCtaBins / kScanThreads;   // 4
  unsigned long long* partial = bins + SFR_SELECT_BINS1;
  unsigned long long* ticket = partial + kScanMaxCtas;
  __shared__ unsigned long long red[32];
  __shared__ bool s_last;
  const int nctas = gridDim.x;                         // NBINS / kScanCtaBins
  // partial sum of this CTA's 1024 bins
  {
    const unsigned long long* mine = bins + (int64_t)blockIdx.x * kScanCtaBins;
    unsigned long long sum = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) sum += mine[threadIdx.x + j * kScanThreads];
    sum = block_sum<unsigned long long>(sum, red);
    if (threadIdx.x == 0) {
      partial[blockIdx.x] = sum;
      __threadfence();
      s_last = atomicAdd(ticket, 1ull) == (unsigned long long)(nctas - 1);
    }
    __syncthreads();
  }
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) *ticket = 0ull;                // ready for the next scan

  const bool skip = pass == 1 && (state->select_none || state->select_all);
  const unsigned long long want = pass == 0 ? state->k : state->k_in_bin;
  int bin = -1;
  unsigned long long above = 0;
  if (!skip) {
    // level 1: the CTA partials, highest block first (thread t owns descending position t)
    unsigned long long pv[1];
    const int blk_of_t = nctas - 1 - (int)threadIdx.x;
    pv[0] = ((int)threadIdx.x < nctas) ? __ldcg(partial + blk_of_t) : 0ull;
    int pos;
    unsigned long long above_blk;
    find_crossing<1>(pv, want, 0ull, &pos, &above_blk);
    if (pos >= 0) {
      // level 2: the 1024 bins of that block, highest bin first (thread t owns descending positions 4t..4t+3)
      const int blk = nctas - 1 - pos;
      const unsigned long long* b = bins + (int64_t)blk * kScanCtaBins;
      unsigned long long v[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) v[j] = __ldcg(b + (kScanCtaBins - 1 - ((int)threadIdx.x * kPer + j)));
      int p2;
      unsigned long long above2;
      find_crossing<kPer>(v, want, above_blk, &p2, &above2);
      bin = blk * kScanCtaBins + (kScanCtaBins - 1 - p2);
      above = above2;
    }
  }
  if (threadIdx.x == 0) {
    if (pass == 0) {
      const unsigned long long k = state->k;
      state->select_none = (k == 0);
      state->select_all = (k != 0 && bin < 0);  // k exceeds the element count
      state->prefix = bin < 0 ? 0u : (uint32_t)bin;
      state->k_in_bin = bin < 0 ? 0ull : k - above;
      state->count_gt = above;  // completed by scan 1
    } else if (bin >= 0) {
      state->thr_key = (state->prefix << 16) | (uint32_t)bin;
      state->count_gt += above;
      state->count_eq = __ldcg(bins + bin);
      state->tie_budget = state->k_in_bin - above;
    }
  }
}

// ---- apply -----------------------------------------------------------------------------------
// Elements are cut into fixed chunks of kChunk consecutive elements.  When more keys equal the
// threshold than the budget allows (the COMMON case at scale: 675 M Gaussian values have ~20
// elements per distinct fp32 value near the median), the lowest flat indices win, which needs the
// number of threshold-equal keys in all earlier chunks:
//   tie_count : scratch[c]            = #ties in chunk c   (from the staged candidates; streaming fallback)
//   tie scan  : two-level, see select_tie_block_sum_kernel
//   apply     : chunks without ties stream (mask = key > thr); only chunks that contain a
//               tie pay for the block-wide ordered ranking.
constexpr int kApplyThreads = 256;
constexpr int kChunkVecs = 8;                            // float4 per thread per chunk
constexpr int64_t kChunk = (int64_t)kApplyThreads * 4 * kChunkVecs;  // 8192 elements
static_assert(kChunk == kChunkElems, "scratch layout and apply kernels must agree on the chunk size");

__device__ __forceinline__ bool ties_need_order(const sfr_select_state* s) {
  return !s->select_none && !s->select_all && s->tie_budget != s->count_eq;
}

// Chunks that hold a threshold-equal key, listed by the candidate walk so that the ordered apply visits
// only those (a CTA polling 70 per-chunk counters one after the other is a 50 us chain of dependent L2 loads).
// The list occupies scratch[nchunks + nblocks, 2 * nchunks): everything the block bases do not use.
__host__ __device__ inline int64_t tie_list_offset(int64_t nchunks) { return nchunks + (nchunks + 4095) / 4096; }
__host__ __device__ inline int64_t tie_list_cap(int64_t nchunks) { return 2 * nchunks - tie_list_offset(nchunks); }

// The ordered apply can rank a chunk's ties straight from the tie-chunk list when that list is complete and short:
// "ties before chunk c" = tie_base + sum of the counters of the listed chunks below c.  That skips the two-level scan of
// all per-chunk counters (two phases and two grid-wide barriers of the fused apply).  Uniform over the grid.
constexpr unsigned long long kTieListFast = 2048;

// The candidate list is complete (no region overflowed) and pass 1 left the provisional mask in THIS
// mask buffer: only the staged candidates (and the chunks that hold an ordered tie) remain to be written.
__device__ __forceinline__ bool provisional_ok(const unsigned long long* hdr, const uint8_t* mask) {
  return hdr[0] == 0ull && hdr[3] == 1ull && hdr[4] == (unsigned long long)(uintptr_t)mask;
}
__device__ __forceinline__ bool tie_list_short(const unsigned long long* hdr, const uint8_t* mask) {
  return provisional_ok(hdr, mask) && hdr[6] == 0ull && hdr[5] <= kTieListFast;
}

// Keys of one chunk, all loads issued before the first use.  A chunk is kChunkVecs slabs of
// kApplyThreads float4; thread t owns elements [base + (slab*256 + t)*4, +4), so flat order ==
// (slab, thread, component) order.  Out-of-range elements get key 0 and valid bit 0.
template <int MODE>
__device__ __forceinline__ void load_chunk_keys(const float* __restrict__ a, const float* __restrict__ b,
                                                float eps, int64_t n, int64_t base,
                                                uint32_t (&key)[kChunkVecs][4], uint32_t& valid_bits) {
  valid_bits = 0;
  if (base + kChunk <= n) {
    float4 x[kChunkVecs], y[kChunkVecs];
#pragma unroll
    for (int sl = 0; sl < kChunkVecs; ++sl) {
      const int64_t e0 = base + ((int64_t)sl * kApplyThreads + threadIdx.x) * 4;
      x[sl] = ld_once(reinterpret_cast<const float4*>(a + e0));
      if constexpr (MODE != SFR_KEY_ABS) y[sl] = ld_once(reinterpret_cast<const float4*>(b + e0));
      else y[sl] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int sl = 0; sl < kChunkVecs; ++sl) {
      key[sl][0] = key_from<MODE>(x[sl].x, y[sl].x, eps);
      key[sl][1] = key_from<MODE>(x[sl].y, y[sl].y, eps);
      key[sl][2] = key_from<MODE>(x[sl].z, y[sl].z, eps);
      key[sl][3] = key_from<MODE>(x[sl].w, y[sl].w, eps);
    }
    valid_bits = 0xffffffffu;
  } else {  // the last, ragged chunk
#pragma unroll
    for (int sl = 0; sl < kChunkVecs; ++sl) {
      const int64_t e0 = base + ((int64_t)sl * kApplyThreads + threadIdx.x) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool ok = e0 + q < n;
        key[sl][q] = ok ? key_from<MODE>(a[e0 + q], MODE != SFR_KEY_ABS ? b[e0 + q] : 0.f, eps) : 0u;
        valid_bits |= (ok ? 1u : 0u) << (sl * 4 + q);
      }
    }
  }
}

template <int MODE>
__device__ __forceinline__ void
tie_count_body(const float* __restrict__ a, const float* __restrict__ b, float eps,
               int64_t n, const sfr_select_state* __restrict__ state,
               unsigned long long* __restrict__ scratch) {
  if (!ties_need_order(state)) return;
  const int64_t nchunks = (n + kChunk - 1) / kChunk;
  if (scratch[2 * nchunks] == 0ull) return;  // candidates did not overflow: counted from the list
  __shared__ unsigned int red[32];
  const uint32_t thr = state->thr_key;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    uint32_t key[kChunkVecs][4], valid;
    load_chunk_keys<MODE>(a, b, eps, n, c * kChunk, key, valid);
    unsigned int cnt = 0;
#pragma unroll
    for (int sl = 0; sl < kChunkVecs; ++sl)
#pragma unroll
      for (int q = 0; q < 4; ++q) cnt += ((valid >> (sl * 4 + q)) & 1u) && key[sl][q] == thr;
    cnt = block_sum<unsigned int>(cnt, red);
    if (threadIdx.x == 0) scratch[c] = cnt;
    __syncthreads();
  }
}

// Per-chunk tie counts from the staged candidates (the usual path): every region is scanned by
// one CTA; only entries whose low key bits equal the threshold's touch a (zeroed) chunk counter.
// When pass 1 left a provisional mask, the same walk also FINISHES the mask for the candidates: low bits
// above the threshold's -> 1; equal -> 1 if every tie is selected, else left 0 for the ordered kernel.
__device__ __forceinline__ void
resolve_candidates_body(int64_t n, const sfr_select_state* __restrict__ state,
                        unsigned long long* __restrict__ scratch, uint8_t* __restrict__ mask) {
  if (state->select_none || state->select_all) return;
  const int64_t nchunks = (n + kChunk - 1) / kChunk;
  unsigned long long* hdr = scratch + 2 * nchunks;
  if (hdr[0] != 0ull) return;  // overflow: the streaming kernels count and write instead
  const bool order = ties_need_order(state);
  const bool write = provisional_ok(hdr, mask);
  if (!order && !write) return;
  unsigned long long* tie_list = scratch + tie_list_offset(nchunks);
  const unsigned long long list_cap = (unsigned long long)tie_list_cap(nchunks);
  const unsigned int thr16 = state->thr_key & 0xffffu;
  const int regions = (int)hdr[1];
  const unsigned long long cap = hdr[2];
  // Work item = one slice of one region, taken by ONE WARP; a region is cut into as many slices as it takes to give
  // every warp of the grid about one item.  Each item is a chain of dependent global accesses (count -> entries ->
  // mask / counters), so what the walk costs is the number of items a warp does one after the other: with CTA-wide
  // items the 148 CTAs of a small launch each did 16 of them in turn (~30 us at 3.9e7 elements, a third of the select).
  // Every lane issues its (up to) four list loads before it touches the mask.
  constexpr int kUnroll = 4;
  if (regions <= 0) return;
  const int warps_per_cta = (int)(blockDim.x >> 5);
  const int total_warps = (int)gridDim.x * warps_per_cta;
  int split = (total_warps + regions - 1) / regions;
  split = split < 1 ? 1 : (split > 64 ? 64 : split);
  const int items = regions * split;
  const unsigned int lane = threadIdx.x & 31u;
  // consecutive items go to different CTAs (SMs), then to the next warp of each
  for (int w = (int)(threadIdx.x >> 5) * (int)gridDim.x + (int)blockIdx.x; w < items; w += total_warps) {
    const int r = w / split, part = w % split;
    const unsigned long long* region = hdr + kCandHeader + (unsigned long long)r * cap;
    const unsigned int cnt = (unsigned int)hdr[8 + r];
    const unsigned int per = (cnt + split - 1) / split;
    const unsigned int lo = part * per;
    const unsigned int hi = lo + per < cnt ? lo + per : cnt;
    for (unsigned int i0 = lo + lane; i0 < hi; i0 += 32u * kUnroll) {
      unsigned long long e[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const unsigned int i = i0 + u * 32u;
        e[u] = i < hi ? __ldcs(region + i) : ~0ull;
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (i0 + u * 32u >= hi) continue;
        const unsigned int low = (unsigned int)(e[u] & 0xffffull);
        if (low == thr16) {
          if (order) {
            const unsigned long long chunk = (e[u] >> 16) / kChunk;
            if (atomicAdd(scratch + chunk, 1ull) == 0ull) {      // first tie seen in this chunk: list it
              const unsigned long long pos = atomicAdd(hdr + 5, 1ull);
              if (pos < list_cap) tie_list[pos] = chunk;
              else hdr[6] = 1ull;
            }
          } else {
            mask[e[u] >> 16] = 1;            // write is true here
          }
        } else if (write && low > thr16) {
          mask[e[u] >> 16] = 1;
        }
      }
    }
  }
}

// Two-level exclusive scan of the per-chunk tie counts (a single-CTA scan of 82 k counters cost 96 us):
//   tie_block_sum : one CTA per kScanBlock consecutive chunks -> scratch[nchunks + b] = their tie total
//   tie_block_scan: one CTA turns the <= a few hundred block totals into exclusive bases (+ tie_base)
// The prefix INSIDE a block is formed on demand by the apply kernel, and only by the CTAs whose chunk
// contains a tie (a handful, except on degenerate inputs).
constexpr int kScanBlock = 4096;  // chunks per scan block (= 33.5 M elements)

__device__ __forceinline__ void
tie_block_sum_body(int64_t nchunks, const sfr_select_state* __restrict__ state,
                   unsigned long long* __restrict__ scratch) {
  if (!ties_need_order(state)) return;
  __shared__ unsigned long long red[32];
  const int64_t nblocks = (nchunks + kScanBlock - 1) / kScanBlock;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t lo = blk * kScanBlock;
    const int64_t hi = lo + kScanBlock < nchunks ? lo + kScanBlock : nchunks;
    unsigned long long sum = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) sum += scratch[i];
    sum = block_sum<unsigned long long>(sum, red);
    if (threadIdx.x == 0) scratch[nchunks + blk] = sum;
    __syncthreads();
  }
}

// (runs in ONE CTA of any size that is a multiple of 32 threads, up to 1024)
__device__ __forceinline__ void
tie_block_scan_body(int64_t nchunks, const sfr_select_state* __restrict__ state,
                    const unsigned long long* __restrict__ tie_base,
                    unsigned long long* __restrict__ scratch) {
  if (!ties_need_order(state)) return;
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry_s;
  const int64_t nblocks = (nchunks + kScanBlock - 1) / kScanBlock;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long carry = tie_base ? *tie_base : 0ull;
  unsigned long long* sums = scratch + nchunks;
  const int nwarps = (int)(blockDim.x >> 5);
  for (int64_t base = 0; base < nblocks; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const unsigned long long mine = i < nblocks ? sums[i] : 0ull;
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(kFullMask, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = lane < nwarps ? warp_tot[lane] : 0ull, wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(kFullMask, wi, o);
        if (lane >= o) wi += up;
      }
      warp_tot[lane] = wi - w;
      if (lane == 31) carry_s = wi;
    }
    __syncthreads();
    if (i < nblocks) sums[i] = carry + warp_tot[warp] + (incl - mine);  // exclusive base of block i
    carry += carry_s;
    __syncthreads();
  }
}

// No tie needs ordering (tie_budget == count_eq, or select all / none):
//   mask = key >= thr       — a pure stream: 4 (8) B read + 1 B written per element.
constexpr int kApplyUnroll = 4;

template <int MODE>
__device__ __forceinline__ void
apply_stream_body(const float* __restrict__ a, const float* __restrict__ b, float eps,
                  int64_t n, const sfr_select_state* __restrict__ state,
                  const unsigned long long* __restrict__ scratch, uint8_t* __restrict__ mask) {
  if (ties_need_order(state)) return;  // the ordered kernel below writes the mask instead
  const bool none = state->select_none != 0;
  const bool all = state->select_all != 0;
  // pass 1's provisional mask + the candidate walk already produced the whole mask
  if (!none && !all && provisional_ok(scratch + 2 * scratch_nchunks(n), mask)) return;
  const uint32_t thr = all ? 0u : state->thr_key;  // all: every key (NaN's 0 included) >= 0
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)kApplyThreads * kApplyUnroll;
  const int64_t ntiles = (nvec + tile - 1) / tile;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  unsigned int* m4 = reinterpret_cast<unsigned int*>(mask);
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + threadIdx.x;
    float4 x[kApplyUnroll], y[kApplyUnroll];
#pragma unroll
    for (int u = 0; u < kApplyUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kApplyThreads;
      const bool in = v < nvec;
      x[u] = in ? ld_once(a4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (MODE != SFR_KEY_ABS)
        y[u] = in ? ld_once(b4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      else
        y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kApplyUnroll; ++u) {
      const int64_t v = base + (int64_t)u * kApplyThreads;
      if (v >= nvec) continue;
      const uint32_t s0 = !none && key_from<MODE>(x[u].x, y[u].x, eps) >= thr;
      const uint32_t s1 = !none && key_from<MODE>(x[u].y, y[u].y, eps) >= thr;
      const uint32_t s2 = !none && key_from<MODE>(x[u].z, y[u].z, eps) >= thr;
      const uint32_t s3 = !none && key_from<MODE>(x[u].w, y[u].w, eps) >= thr;
      m4[v] = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
    }
  }
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < (n - tail0)) {
    const int64_t i = tail0 + threadIdx.x;
    mask[i] = (uint8_t)(!none && key_from<MODE>(a[i], MODE != SFR_KEY_ABS ? b[i] : 0.f, eps) >= thr);
  }
}

// More threshold-equal keys than the budget: lowest flat index first.
template <int MODE>
__device__ __forceinline__ void
apply_ordered_body(const float* __restrict__ a, const float* __restrict__ b, float eps,
                   int64_t n, const sfr_select_state* __restrict__ state,
                   const unsigned long long* __restrict__ tie_base,
                   const unsigned long long* __restrict__ scratch, uint8_t* __restrict__ mask) {
  if (!ties_need_order(state)) return;  // the streaming kernel above wrote the mask
  __shared__ unsigned int warp_tot[kApplyThreads / 32];
  const uint32_t thr = state->thr_key;
  const unsigned long long budget = state->tie_budget;
  const int64_t nchunks = (n + kChunk - 1) / kChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // provisional mask in place: only the chunks that hold a threshold-equal key are rewritten, taken
  // from the list the candidate walk left (or, if that overflowed, found by polling every counter)
  const unsigned long long* hdr = scratch + 2 * nchunks;
  const bool only_tie_chunks = provisional_ok(hdr, mask);
  const bool listed = only_tie_chunks && hdr[6] == 0ull;
  const bool short_list = tie_list_short(hdr, mask);          // then the block bases were never formed
  const unsigned long long* tie_list = scratch + tie_list_offset(nchunks);
  const int64_t visits = listed ? (int64_t)hdr[5] : nchunks;

  for (int64_t visit = blockIdx.x; visit < visits; visit += gridDim.x) {
    const int64_t c = listed ? (int64_t)tie_list[visit] : visit;
    const int64_t base = c * kChunk;
    const unsigned long long chunk_ties = scratch[c];        // uniform over the CTA
    if (chunk_ties == 0 && only_tie_chunks) continue;
    if (chunk_ties == 0 && base + kChunk <= n) {
      // tie-free full chunk (almost all of them): straight-line stream, mask = key > thr
      float4 x[kChunkVecs], y[kChunkVecs];
#pragma unroll
      for (int sl = 0; sl < kChunkVecs; ++sl) {
        const int64_t e0 = base + ((int64_t)sl * kApplyThreads + threadIdx.x) * 4;
        x[sl] = ld_once(reinterpret_cast<const float4*>(a + e0));
        if constexpr (MODE != SFR_KEY_ABS) y[sl] = ld_once(reinterpret_cast<const float4*>(b + e0));
        else y[sl] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int sl = 0; sl < kChunkVecs; ++sl) {
        const int64_t e0 = base + ((int64_t)sl * kApplyThreads + threadIdx.x) * 4;
        const uint32_t s0 = key_from<MODE>(x[sl].x, y[sl].x, eps) > thr;
        const uint32_t s1 = key_from<MODE>(x[sl].y, y[sl].y, eps) > thr;
        const uint32_t s2 = key_from<MODE>(x[sl].z, y[sl].z, eps) > thr;
        const uint32_t s3 = key_from<MODE>(x[sl].w, y[sl].w, eps) > thr;
        *reinterpret_cast<unsigned int*>(mask + e0) = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
      }
      continue;
    }
    // ties before this chunk = base of its scan block + counts of the earlier chunks of the block
    __shared__ unsigned long long red64[32];
    __shared__ unsigned long long run_s;
    if (short_list) {
      // straight from the list: the counters of the listed chunks below this one (+ the ties of lower ranks)
      unsigned long long part = 0;
      for (int64_t j = threadIdx.x; j < visits; j += kApplyThreads) {
        const unsigned long long other = tie_list[j];
        if ((int64_t)other < c) part += scratch[other];
      }
      part = block_sum<unsigned long long>(part, red64);
      if (threadIdx.x == 0) run_s = (tie_base ? *tie_base : 0ull) + part;
      __syncthreads();
    } else {
      const int64_t blk = c / kScanBlock;
      unsigned long long part = 0;
      for (int64_t i = blk * kScanBlock + threadIdx.x; i < c; i += kApplyThreads) part += scratch[i];
      part = block_sum<unsigned long long>(part, red64);
      if (threadIdx.x == 0) run_s = scratch[nchunks + blk] + part;
      __syncthreads();
    }
    unsigned long long run = run_s;
    uint32_t key[kChunkVecs][4], valid;
    load_chunk_keys<MODE>(a, b, eps, n, base, key, valid);
#pragma unroll
    for (int sl = 0; sl < kChunkVecs; ++sl) {
      const int64_t e0 = base + ((int64_t)sl * kApplyThreads + threadIdx.x) * 4;
      uint32_t sel[4];
      {
        unsigned int tq[4], mine = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tq[q] = ((valid >> (sl * 4 + q)) & 1u) && key[sl][q] == thr;
          mine += tq[q];
        }
        // exclusive prefix of tie counts in flat order: warp scan, then across warps
        unsigned int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int up = __shfl_up_sync(kFullMask, incl, o);
          if (lane >= o) incl += up;
        }
        __syncthreads();  // warp_tot of the previous slab fully consumed
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned int before = 0, slab_total = 0;
#pragma unroll
        for (int w = 0; w < kApplyThreads / 32; ++w) {
          const unsigned int t = warp_tot[w];
          if (w < warp) before += t;
          slab_total += t;
        }
        unsigned long long rank = run + before + (incl - mine);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          sel[q] = key[sl][q] > thr || (tq[q] && rank < budget);
          rank += tq[q];
        }
        run += slab_total;  // every thread keeps the same running count
      }
      if (((valid >> (sl * 4)) & 0xfu) == 0xfu) {
        *reinterpret_cast<unsigned int*>(mask + e0) = sel[0] | (sel[1] << 8) | (sel[2] << 16) | (sel[3] << 24);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if ((valid >> (sl * 4 + q)) & 1u) mask[e0 + q] = (uint8_t)sel[q];
      }
    }
  }
}


// ---- the whole apply stage in ONE cooperative launch -------------------------------------------------
// The stages above depend on each other grid-wide (tie counts -> block sums -> block bases -> ordered apply).
// As separate launches (2 memsets + 6 kernels, most of them a few microseconds of work) they cost 74 us of
// the 130 us select at the DDPM size (38.6 M elements) — launch latency, not bandwidth.  Here they are phases
// of one persistent grid separated by grid-wide barriers; the branches on the device-side select state are
// uniform over the grid, so every CTA reaches the same barriers.
template <int MODE>
__global__ void __launch_bounds__(kApplyThreads, MODE != SFR_KEY_ABS ? 2 : 3)
select_apply_fused_kernel(const float* __restrict__ a, const float* __restrict__ b, float eps, int64_t n,
                          const sfr_select_state* __restrict__ state,
                          const unsigned long long* __restrict__ tie_base,
                          unsigned long long* __restrict__ scratch, uint8_t* __restrict__ mask) {
  cg::grid_group grid = cg::this_grid();
  const int64_t nchunks = (n + kChunk - 1) / kChunk;
  const bool order = ties_need_order(state);
  unsigned long long* hdr = scratch + 2 * nchunks;
  if (order && hdr[7] != 0ull) {
    // an apply has already run on this pass 1's scratch (hdr[7], set below; zeroed with the header by pass 1): the
    // per-chunk counters are accumulated into and the tie-chunk list appended to, so clear them before walking again
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nchunks; i += (int64_t)gridDim.x * blockDim.x)
      scratch[i] = 0ull;
    if (blockIdx.x == 0 && threadIdx.x < 2) hdr[5 + threadIdx.x] = 0ull;
    grid.sync();
  }
  resolve_candidates_body(n, state, scratch, mask);   // list walk: finishes the provisional mask, counts ties
  tie_count_body<MODE>(a, b, eps, n, state, scratch); // (only after a candidate overflow: streaming count)
  if (order) {
    grid.sync();                                       // every CTA has read hdr[7] by now
    if (blockIdx.x == 0 && threadIdx.x == 0) hdr[7] = 1ull;
    if (!tie_list_short(hdr, mask)) {                  // long / overflowed list, or no provisional mask: scan every counter
      tie_block_sum_body(nchunks, state, scratch);
      grid.sync();
      if (blockIdx.x == 0) tie_block_scan_body(nchunks, state, tie_base, scratch);
      grid.sync();
    }
    apply_ordered_body<MODE>(a, b, eps, n, state, tie_base, scratch, mask);
  } else {
    apply_stream_body<MODE>(a, b, eps, n, state, scratch, mask);   // returns at once when the walk wrote the mask
  }
}

// co-resident CTAs of the fused kernel, per device and key mode (occupancy x SMs)
template <int MODE>
int fused_grid(int device) {
  static std::atomic<int> cached[64];
  int g = cached[device & 63].load(std::memory_order_acquire);
  if (g == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_apply_fused_kernel<MODE>, kApplyThreads, 0) !=
            cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    g = per_sm * device_geometry().sm_count;
    cached[device & 63].store(g, std::memory_order_release);
  }
  return g;
}

template <int MODE>
cudaError_t launch_apply_fused(int device, const float* a, const float* b, float eps, int64_t n,
                               const sfr_select_state* state, const unsigned long long* tie_base,
                               unsigned long long* scratch, uint8_t* mask, cudaStream_t s) {
  // a grid-wide barrier costs more the more CTAs take part, and the phases between the barriers are short: one CTA
  // per SM unless the vector is large enough to give every co-resident CTA some chunks
  const int64_t nchunks = (n + kChunk - 1) / kChunk;
  const int cap = fused_grid<MODE>(device);
  const int sms = device_geometry().sm_count;
  int64_t want = nchunks / 32;
  if (want < sms) want = sms;
  const int grid = (int)(want < cap ? want : cap);
  void* args[] = {(void*)&a, (void*)&b, (void*)&eps, (void*)&n, (void*)&state, (void*)&tie_base, (void*)&scratch,
                  (void*)&mask};
  return cudaLaunchCooperativeKernel((const void*)select_apply_fused_kernel<MODE>, dim3(grid), dim3(kApplyThreads),
                                     args, 0, s);
}

__global__ void select_init_kernel(sfr_select_state* state, unsigned long long* bins,
                                   unsigned long long k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < SFR_SELECT_BINS_ALLOC) bins[i] = 0ull;
  if (i == 0) {
    sfr_select_state z{};
    z.k = k;
    *state = z;
  }
}

}  // namespace
}  // namespace sfr

extern "C" int sfr_select_init(sfr_select_state* state, unsigned long long* bins,
                               unsigned long long k, sfr_stream_t stream) {
  using namespace sfr;
  SFR_REQUIRE_PTR(state);
  SFR_REQUIRE_PTR(bins);
  SFR_ENTER_DEVICE(state);
  select_init_kernel<<<(SFR_SELECT_BINS_ALLOC + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(state, bins, k);
  SFR_LAUNCH_STATUS();
}

static int select_hist_impl(const float* a, const float* b, int key_mode, float eps,
                            int64_t n, int pass, const sfr_select_state* state,
                            unsigned long long* bins, unsigned long long* scratch,
                            uint8_t* mask, sfr_stream_t stream) {
  using namespace sfr;
  if (n < 0 || (pass != 0 && pass != 1)) return SFR_ERR_ARG;
  if (key_mode < SFR_KEY_ABS || key_mode > SFR_KEY_ABSDIFF) return SFR_ERR_ARG;
  SFR_REQUIRE_PTR(bins);
  SFR_REQUIRE_PTR(state);
  if (n == 0) return SFR_OK;
  SFR_REQUIRE_PTR(a);
  if (key_mode != SFR_KEY_ABS) SFR_REQUIRE_PTR(b);
  SFR_REQUIRE_ALIGNED(a);
  SFR_REQUIRE_ALIGNED(b);
  SFR_ENTER_DEVICE(state);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t nvec = n >> 2;
  if (pass == 0) {
    // the opt-in to 128 KB of dynamic shared memory is per device (and this may be called from any thread)
    static std::atomic<bool> attr_done_on[64];
    const int smem = SFR_SELECT_BINS0 * (int)sizeof(unsigned int);
    std::atomic<bool>& attr_done = attr_done_on[device_scope__.device() & 63];
    if (!attr_done.load(std::memory_order_acquire)) {
      cudaFuncSetAttribute(select_hist0_kernel<SFR_KEY_ABS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(select_hist0_kernel<SFR_KEY_RATIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(select_hist0_kernel<SFR_KEY_ABSDIFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr_done.store(true, std::memory_order_release);
    }
    const int64_t tile = (int64_t)kHistThreads * kHistUnroll;
    const int grid = persistent_grid((nvec + tile - 1) / tile, 1);
    if (key_mode == SFR_KEY_ABS) select_hist0_kernel<SFR_KEY_ABS><<<grid, kHistThreads, smem, s>>>(a, b, eps, n, bins);
    else if (key_mode == SFR_KEY_RATIO) select_hist0_kernel<SFR_KEY_RATIO><<<grid, kHistThreads, smem, s>>>(a, b, eps, n, bins);
    else select_hist0_kernel<SFR_KEY_ABSDIFF><<<grid, kHistThreads, smem, s>>>(a, b, eps, n, bins);
  } else {
    const int64_t tile = (int64_t)kFiltThreads * kFiltUnroll;
    SFR_REQUIRE_PTR(scratch);
    int grid = persistent_grid((nvec + tile - 1) / tile, kFiltCtasPerSm);
    if (grid > kMaxRegions) grid = kMaxRegions;
    // zero the per-chunk tie counters and the candidate header (the regions need no clearing)
    cudaMemsetAsync(scratch, 0, (size_t)(2 * scratch_nchunks(n) + kCandHeader) * sizeof(unsigned long long), s);
#define SFR_HIST1(M)                                                                                          \
  do {                                                                                                        \
    if (mask) select_hist1_kernel<M, true><<<grid, kFiltThreads, 0, s>>>(a, b, eps, n, state, bins, scratch, mask); \
    else select_hist1_kernel<M, false><<<grid, kFiltThreads, 0, s>>>(a, b, eps, n, state, bins, scratch, nullptr);  \
  } while (0)
    if (key_mode == SFR_KEY_ABS) SFR_HIST1(SFR_KEY_ABS);
    else if (key_mode == SFR_KEY_RATIO) SFR_HIST1(SFR_KEY_RATIO);
    else SFR_HIST1(SFR_KEY_ABSDIFF);
#undef SFR_HIST1
  }
  SFR_LAUNCH_STATUS();
}

extern "C" int sfr_select_hist(const float* a, const float* b, int key_mode, float eps,
                               int64_t n, int pass, const sfr_select_state* state,
                               unsigned long long* bins, unsigned long long* scratch,
                               sfr_stream_t stream) {
  return select_hist_impl(a, b, key_mode, eps, n, pass, state, bins, scratch, nullptr, stream);
}

extern "C" int sfr_select_hist1_mask(const float* a, const float* b, int key_mode, float eps,
                                     int64_t n, const sfr_select_state* state,
                                     unsigned long long* bins, unsigned long long* scratch,
                                     uint8_t* mask, sfr_stream_t stream) {
  using namespace sfr;
  if (n > 0) {
    SFR_REQUIRE_PTR(mask);
    SFR_REQUIRE_ALIGNED(mask);
  }
  return select_hist_impl(a, b, key_mode, eps, n, 1, state, bins, scratch, mask, stream);
}

extern "C" int sfr_select_scan(int pass, sfr_select_state* state, unsigned long long* bins,
                               sfr_stream_t stream) {
  using namespace sfr;
  if (pass != 0 && pass != 1) return SFR_ERR_ARG;
  SFR_REQUIRE_PTR(state);
  SFR_REQUIRE_PTR(bins);
  SFR_ENTER_DEVICE(state);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nbins = pass == 0 ? SFR_SELECT_BINS0 : SFR_SELECT_BINS1;
  select_scan_kernel<<<nbins / kScanCtaBins, kScanThreads, 0, s>>>(pass, state, bins);
  // leave the bins clean for the next pass / the next select (the partials and the ticket stay as they are)
  cudaMemsetAsync(bins, 0, (size_t)SFR_SELECT_BINS1 * sizeof(unsigned long long), s);
  SFR_LAUNCH_STATUS();
}

extern "C" int64_t sfr_select_scratch_elems(int64_t n) {
  if (n < 0) n = 0;
  // per-chunk tie counts + their exclusive scan + candidate header + candidate regions
  return 2 * sfr::scratch_nchunks(n) + sfr::kCandHeader + sfr::scratch_cand_cap(n);
}

extern "C" int sfr_select_apply(const float* a, const float* b, int key_mode, float eps,
                                int64_t n, const sfr_select_state* state,
                                const unsigned long long* tie_base,
                                unsigned long long* scratch, uint8_t* mask,
                                sfr_stream_t stream) {
  using namespace sfr;
  if (n < 0) return SFR_ERR_ARG;
  if (key_mode < SFR_KEY_ABS || key_mode > SFR_KEY_ABSDIFF) return SFR_ERR_ARG;
  SFR_REQUIRE_PTR(state);
  if (n == 0) return SFR_OK;
  SFR_REQUIRE_PTR(a);
  if (key_mode != SFR_KEY_ABS) SFR_REQUIRE_PTR(b);
  SFR_REQUIRE_PTR(scratch);
  SFR_REQUIRE_PTR(mask);
  SFR_REQUIRE_ALIGNED(a);
  SFR_REQUIRE_ALIGNED(b);
  SFR_REQUIRE_ALIGNED(mask);
  SFR_ENTER_DEVICE(state);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  const int dev = device_scope__.device();
  if (key_mode == SFR_KEY_ABS) e = launch_apply_fused<SFR_KEY_ABS>(dev, a, b, eps, n, state, tie_base, scratch, mask, s);
  else if (key_mode == SFR_KEY_RATIO) e = launch_apply_fused<SFR_KEY_RATIO>(dev, a, b, eps, n, state, tie_base, scratch, mask, s);
  else e = launch_apply_fused<SFR_KEY_ABSDIFF>(dev, a, b, eps, n, state, tie_base, scratch, mask, s);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return (int)e;
  }
  SFR_LAUNCH_STATUS();
}
