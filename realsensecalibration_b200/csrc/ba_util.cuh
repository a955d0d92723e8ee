// ba_util.cuh -- error plumbing, device buffers, cub wrappers, deterministic reductions.
#pragma once
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "ba_cuda.h"

namespace ba {

// ---- thread-local last error --------------------------------------------------------
inline char* err_buf() {
  static thread_local char buf[1024] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

#define BA_CUDA_TRY(expr)                                                                       \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return ba::fail(e__ == cudaErrorMemoryAllocation ? BA_ERR_OUT_OF_MEMORY : BA_ERR_CUDA,    \
                      "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));   \
  } while (0)

#define BA_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != BA_OK) return rc__; \
  } while (0)

// ---- caching device allocator ----------------------------------------------------------
// cudaMalloc / cudaFree cost milliseconds for the buffers of a BAL-sized problem and cudaFree synchronises the
// device; ba_cuda_set_model_* allocates ~150 buffers.  Freed blocks are therefore kept in a per-device free list and
// handed out again by size (best fit).  Reuse is stream-ordered: every block remembers the stream that was current
// when it was freed; taking it from another stream first synchronises the device (rare: one stream per problem).
struct DevCache {
  struct Block { void* p; int dev; cudaStream_t st; };
  std::mutex mu;
  std::multimap<size_t, Block> free_blocks;
  std::unordered_map<void*, size_t> live;   // every block handed out -> rounded size
  size_t cached_bytes = 0;
  static DevCache& get() { static DevCache* c = new DevCache(); return *c; }  // leaked on purpose: outlives static destructors
  static cudaStream_t& current_stream() { static thread_local cudaStream_t s = nullptr; return s; }
  static size_t round_up(size_t b) {
    if (b == 0) b = 1;
    const size_t g = b < (1u << 20) ? 512 : (size_t)2 << 20;
    return (b + g - 1) / g * g;
  }
  cudaError_t malloc(void** out, size_t bytes) {
    const size_t want = round_up(bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> lk(mu);
      for (auto it = free_blocks.lower_bound(want); it != free_blocks.end() && it->first <= want + want / 4 + 512; ++it) {
        if (it->second.dev != dev) continue;
        Block b = it->second;
        const size_t sz = it->first;
        free_blocks.erase(it);
        cached_bytes -= sz;
        live[b.p] = sz;
        if (b.st != nullptr && b.st != current_stream()) cudaDeviceSynchronize();
        *out = b.p;
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and try again
      cudaGetLastError();
      trim();
      e = cudaMalloc(out, want);
    }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(mu); live[*out] = want; }
    return e;
  }
  void free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu);
    auto it = live.find(p);
    if (it == live.end()) { cudaFree(p); return; }
    int dev = 0;
    cudaGetDevice(&dev);
    free_blocks.emplace(it->second, Block{p, dev, current_stream()});
    cached_bytes += it->second;
    live.erase(it);
  }
  void mark_clean(cudaStream_t st) {  // the stream was synchronised: its freed blocks are safe for any stream
    std::lock_guard<std::mutex> lk(mu);
    for (auto& kv : free_blocks) if (kv.second.st == st) kv.second.st = nullptr;
  }
  void trim() {  // cudaFree everything cached (all devices)
    std::lock_guard<std::mutex> lk(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto& kv : free_blocks) { cudaSetDevice(kv.second.dev); cudaFree(kv.second.p); }
    cudaSetDevice(dev);
    free_blocks.clear();
    cached_bytes = 0;
  }
};

// ---- owning device buffer -----------------------------------------------------------
template <typename T>
struct DVec {
  T* p = nullptr;
  size_t n = 0;
  bool borrowed = false;   // a view into someone else's allocation (the host-built structure arena): never freed here
  DVec() = default;
  DVec(const DVec&) = delete;
  DVec& operator=(const DVec&) = delete;
  ~DVec() { release(); }
  void release() {
    if (p && !borrowed) DevCache::get().free(p);
    p = nullptr;
    n = 0;
    borrowed = false;
  }
  void borrow(T* ptr, size_t count) {
    release();
    p = ptr; n = count; borrowed = true;
  }
  int alloc(size_t count) {
    release();
    n = count;
    if (count == 0) count = 1;  // keep pointers non-null so kernels can take them
    BA_CUDA_TRY(DevCache::get().malloc((void**)&p, count * sizeof(T)));
    return BA_OK;
  }
  int alloc_zero(size_t count, cudaStream_t s) {
    BA_TRY(alloc(count));
    BA_CUDA_TRY(cudaMemsetAsync(p, 0, (count ? count : 1) * sizeof(T), s));
    return BA_OK;
  }
  int upload(const T* host, size_t count, cudaStream_t s) {
    BA_TRY(alloc(count));
    if (count) BA_CUDA_TRY(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    return BA_OK;
  }
  void swap(DVec& o) {
    T* tp = p; p = o.p; o.p = tp;
    size_t tn = n; n = o.n; o.n = tn;
    bool tb = borrowed; borrowed = o.borrowed; o.borrowed = tb;
  }
  size_t bytes() const { return n * sizeof(T); }
};

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block > 0 ? (n + block - 1) / block : 1); }

// ---- cub two-phase call helper --------------------------------------------------------
template <typename F>
int cub_call(F&& f) {
  void* tmp = nullptr;
  size_t bytes = 0;
  BA_CUDA_TRY(f(tmp, bytes));
  BA_CUDA_TRY(DevCache::get().malloc(&tmp, bytes ? bytes : 1));
  cudaError_t e = f(tmp, bytes);
  DevCache::get().free(tmp);
  BA_CUDA_TRY(e);
  return BA_OK;
}

// ---- device-side deterministic reductions ----------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;  // valid in lane 0
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}
// Sum of N values over a group of G lanes (G a power of two, the group aligned in the warp) by recursive halving: at every
// step a lane keeps one half of its values and sends the other half to its partner, so the whole reduction costs about N
// shuffles instead of N log2(G) and ends with the N sums spread over the lanes, which store them to dst[0..N).  Whole warps
// call this; the order of the additions is fixed.
template <int N, int G>
struct HalvingSum {
  static __device__ __forceinline__ void run(const double* v, int lane, int base, int cnt, double* __restrict__ dst, bool live) {
    if constexpr (G == 1) {
#pragma unroll
      for (int q = 0; q < N; ++q)
        if (live && q < cnt) dst[base + q] = v[q];
    } else {
      constexpr int LO = (N + 1) / 2, HI = N - LO, H = G / 2;
      const bool up = (lane & H) != 0;
      double w[LO > 0 ? LO : 1];
#pragma unroll
      for (int q = 0; q < LO; ++q) {
        const double hi = q < HI ? v[LO + q] : 0.0;
        const double send = up ? v[q] : hi, keep = up ? hi : v[q];
        w[q] = keep + __shfl_xor_sync(0xffffffffu, send, H);
      }
      const int c2 = up ? max(cnt - LO, 0) : min(cnt, LO);
      HalvingSum<LO, H>::run(w, lane, base + (up ? LO : 0), c2, dst, live);
    }
  }
};
template <int G>
struct HalvingSum<0, G> {
  static __device__ __forceinline__ void run(const double*, int, int, int, double* __restrict__, bool) {}
};
template <int N, int G>
__device__ __forceinline__ void group_sum_store(const double* v, int lane, double* __restrict__ dst, bool live) {
  HalvingSum<N, G>::run(v, lane, 0, N, dst, live);
}

// Sum over a CTA in a fixed tree order; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double* smem32) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.0;
  if (warp == 0) v = warp_sum(v);
  return v;
}
__device__ __forceinline__ double block_max(double v, double* smem32) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.0;
  if (warp == 0) v = warp_max(v);
  return v;
}

// Folds `n` partials (one per CTA of a previous kernel) into out[slot]; launched <<<1,1024>>>.
__global__ void k_fold_partials(const double* __restrict__ partial, int n, double* out, int slot, int is_max) {
  __shared__ double sm[32];
  double v = 0.0;
  if (is_max) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) v = fmax(v, partial[i]);
    v = block_max(v, sm);
  } else {
    // fixed strided order per thread, then fixed tree: bitwise reproducible for a given n
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partial[i];
    v = block_sum(v, sm);
  }
  if (threadIdx.x == 0) out[slot] = v;
}

// ---- small dense algebra in registers ---------------------------------------------------
template <int D>
__device__ __forceinline__ bool chol_small(double* A) {  // lower Cholesky in place, row-major
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double d = A[j * D + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[j * D + k] * A[j * D + k];
    if (!(d > 0.0) || !isfinite(d)) return false;
    const double l = sqrt(d);
    A[j * D + j] = l;
    const double inv = 1.0 / l;
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      double s = A[i * D + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[i * D + k] * A[j * D + k];
      A[i * D + j] = s * inv;
    }
  }
  return true;
}
template <int D>
__device__ __forceinline__ void fwd_small(const double* L, double* x) {  // x <- L^-1 x
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double s = x[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * D + k] * x[k];
    x[i] = s / L[i * D + i];
  }
}
template <int D>
__device__ __forceinline__ void bwd_small(const double* L, double* x) {  // x <- L^-T x
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
    double s = x[i];
#pragma unroll
    for (int k = i + 1; k < D; ++k) s -= L[k * D + i] * x[k];
    x[i] = s / L[i * D + i];
  }
}

}  // namespace ba
