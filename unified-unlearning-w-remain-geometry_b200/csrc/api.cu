// api.cu — library introspection, device geometry cache, and the flat-gradient gather.
#include <mutex>

#include "common.cuh"

namespace sfr {

namespace {
constexpr int kMaxDevices = 64;
thread_local int tls_scope_device = -1;  // device entered by the innermost DeviceScope of this thread

DeviceGeometry query_geometry(int dev) {
  DeviceGeometry g;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    return g;
  }
  g.sm_count = prop.multiProcessorCount;
  g.cc_major = prop.major;
  g.cc_minor = prop.minor;
  // The kernels are compiled for sm_100a only: anything else cannot run them.
  g.ok = (prop.major == 10);
  return g;
}

const DeviceGeometry& geometry_of(int dev) {
  static DeviceGeometry none;
  if (dev < 0 || dev >= kMaxDevices) return none;
  // one slot per ordinal, each initialised once (thread-safe: function-local statics in the lambdas)
  static DeviceGeometry slots[kMaxDevices];
  static std::once_flag flags[kMaxDevices];
  std::call_once(flags[dev], [dev] { slots[dev] = query_geometry(dev); });
  return slots[dev];
}
}  // namespace

const DeviceGeometry& device_geometry() {
  int dev = tls_scope_device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    dev = -1;
  }
  return geometry_of(dev);
}

DeviceScope::DeviceScope(const void* anchor) {
  if (cudaGetDevice(&prev_) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  dev_ = prev_;
  // one visible device: nothing to look up (and no driver query on the launch path at all)
  static const int device_count = [] {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
      cudaGetLastError();
      c = 0;
    }
    return c;
  }();
  if (anchor != nullptr && device_count > 1) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, anchor) == cudaSuccess) {
      if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) dev_ = at.device;
    } else {
      cudaGetLastError();
    }
  }
  if (dev_ != prev_) {
    if (cudaSetDevice(dev_) != cudaSuccess) {
      cudaGetLastError();
      dev_ = prev_;
      return;
    }
    switched_ = true;
  }
  outer_ = tls_scope_device;
  tls_scope_device = dev_;
  entered_ = true;
  ok_ = geometry_of(dev_).ok;
}

DeviceScope::~DeviceScope() {
  if (entered_) tls_scope_device = outer_;
  if (switched_) cudaSetDevice(prev_);
}

namespace {

// Gather `count` tensors into the flat vector.  One CTA handles kGatherChunk consecutive
// flat elements; the owning segment is found by binary search on the (sorted) offsets.
constexpr int kGatherThreads = 256;
constexpr int64_t kGatherChunk = 4096;

template <int ST>
__global__ void __launch_bounds__(kGatherThreads, 8)
gather_segments_kernel(float* __restrict__ flat, const void* const* __restrict__ srcs,
                       const int64_t* __restrict__ offsets, const int64_t* __restrict__ sizes,
                       int32_t count, int64_t total) {
  const int64_t nchunks = (total + kGatherChunk - 1) / kGatherChunk;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t base = c * kGatherChunk;
    for (int64_t i = base + threadIdx.x; i < base + kGatherChunk && i < total; i += kGatherThreads) {
      // last segment with offsets[seg] <= i
      int lo = 0, hi = count - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid - 1;
      }
      const int64_t j = i - offsets[lo];
      if (j < sizes[lo]) flat[i] = load_g1<ST>(srcs[lo], j);
    }
  }
}

}  // namespace
}  // namespace sfr

extern "C" int sfr_abi_version(void) { return SFR_ABI_VERSION; }

extern "C" const char* sfr_error_string(int code) {
  switch (code) {
    case SFR_OK: return "ok";
    case SFR_ERR_NULL: return "required pointer is NULL";
    case SFR_ERR_ALIGN: return "vector pointer is not 16-byte aligned";
    case SFR_ERR_ARG: return "argument out of range";
    case SFR_ERR_NO_DEVICE: return "no sm_100 CUDA device available (this library has no CPU fallback)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}

extern "C" int sfr_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  const sfr::DeviceGeometry& g = sfr::device_geometry();
  if (g.sm_count == 0) return SFR_ERR_NO_DEVICE;
  if (sm_count) *sm_count = g.sm_count;
  if (cc_major) *cc_major = g.cc_major;
  if (cc_minor) *cc_minor = g.cc_minor;
  return g.ok ? SFR_OK : SFR_ERR_NO_DEVICE;
}

extern "C" int sfr_gather_segments(float* flat, const void* const* srcs, const int64_t* offsets,
                                   const int64_t* sizes, int32_t count, int src_dtype,
                                   int64_t total, sfr_stream_t stream) {
  using namespace sfr;
  if (count < 0 || total < 0) return SFR_ERR_ARG;
  if (src_dtype != SFR_F32 && src_dtype != SFR_BF16) return SFR_ERR_ARG;
  if (count == 0 || total == 0) return SFR_OK;
  SFR_REQUIRE_PTR(flat);
  SFR_REQUIRE_PTR(srcs);
  SFR_REQUIRE_PTR(offsets);
  SFR_REQUIRE_PTR(sizes);
  SFR_ENTER_DEVICE(flat);
  const int64_t nchunks = (total + kGatherChunk - 1) / kGatherChunk;
  const int grid = persistent_grid(nchunks, 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (src_dtype == SFR_F32)
    gather_segments_kernel<SFR_F32><<<grid, kGatherThreads, 0, s>>>(flat, srcs, offsets, sizes, count, total);
  else
    gather_segments_kernel<SFR_BF16><<<grid, kGatherThreads, 0, s>>>(flat, srcs, offsets, sizes, count, total);
  SFR_LAUNCH_STATUS();
}
