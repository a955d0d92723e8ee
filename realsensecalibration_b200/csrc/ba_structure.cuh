// ba_structure.cuh -- one-time, on-device construction of the index structure of a problem.
//
// Every residual block ("obs") touches one eliminated block e (Model A: point, Model B:
// frame) and up to two kept blocks f (Model A: camera; Model B: camera and marker).  This
// file sorts the observations by e, enumerates the (e,f) incidences, the per-f gather
// lists, and the destination blocks (fa <= fb) of the reduced camera system together with
// the list of incidence pairs that contribute to each of them.  All lists are produced by
// stable radix sorts, so every later reduction runs in a fixed, reproducible order.
//
// Replaces what ceres::Problem::AddResidualBlock + Program/ParameterBlockOrdering build on
// the host (reference call sites: bundle_adjustment_manager.cpp:37,50,67,81).
#pragma once
#include <algorithm>

#include "ba_util.cuh"

namespace ba {

struct Chunks {
  int n = 0;     // number of chunks
  int nseg = 0;  // number of segments
  int ch = 0;    // max entries per chunk
  DVec<int32_t> seg;        // chunk -> segment
  DVec<int64_t> begin;      // chunk -> first entry (end = min(begin + ch, ptr[seg + 1]))
  DVec<int32_t> seg_first;  // nseg + 1: first chunk of every segment
};

struct Structure {
  int64_t nb = 0, ne = 0, nf = 0, ninc = 0, npairs = 0;
  int ndest = 0;
  int nslots = 1;  // f slots per obs
  DVec<int32_t> perm, ob_e, ob_f0, ob_f1;
  DVec<int64_t> e_ptr;
  // incidences; for nslots == 1 they alias the observation arrays (incidence id == obs id)
  DVec<int32_t> own_inc_e, own_inc_f, own_ob_inc0, ob_inc1;
  DVec<int64_t> own_einc_ptr;
  const int32_t* inc_e = nullptr;
  const int32_t* inc_f = nullptr;
  const int32_t* ob_inc0 = nullptr;  // nullptr => identity
  const int64_t* einc_ptr = nullptr;
  DVec<int64_t> incobs_ptr;  // nslots == 2 only: incidence -> (obs << 1 | slot)
  DVec<int32_t> incobs;
  DVec<int64_t> finc_ptr, fobs_ptr;
  DVec<int32_t> finc, fobs;  // fobs entries are (obs << 1 | slot)
  DVec<int32_t> dest_fa, dest_fb, diag_dest;
  DVec<uint64_t> dest_keys;  // fa * nf + fb, ascending (the RCS pattern of this rank's shard)
  // open-addressing map key -> destination index (nslots == 1: built instead of sorting every incidence pair)
  DVec<uint64_t> dh_keys;
  DVec<int32_t> dh_val;
  uint64_t dh_mask = 0;
  int dh_shift = 0;
  bool pair_lists = false;   // dpair_ptr / pairs / ch_pairs exist (generic pipeline only; built on demand for Model A)
  DVec<int64_t> dpair_ptr;
  DVec<int2> pairs;
  DVec<int64_t> dobs_ptr;  // nslots == 2 only: dest -> obs having (f0,f1) == (fa,fb)
  DVec<int32_t> dobs;
  Chunks ch_fobs, ch_finc, ch_pairs, ch_dobs;
  DVec<unsigned char> arena;   // host-built structures (build_structure_host): the one allocation every list is a view into
  const int64_t* host_f_act_ptr = nullptr;   // host build: exclusive scan of "f has an observation" (nf + 1), inside the arena
  int64_t host_n_active_e = -1, host_n_active_f = -1;
};

// ---- small setup kernels ---------------------------------------------------------------
__global__ void k_iota(int32_t* a, int64_t n, int shift) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t)(i << shift);
}
__global__ void k_gather_i32(int32_t* dst, const int32_t* src, const int32_t* idx, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src ? src[idx[i]] : -1;
}
// ptr[s] = first position whose key >= s, for s in [0, nseg]; keys sorted ascending, keys >= nseg are "invalid tail".
template <typename K>
__global__ void k_seg_ptr(const K* __restrict__ keys, int64_t n, int64_t nseg, int64_t* ptr) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n == 0) {
    if (i <= nseg) ptr[i] = 0;
    return;
  }
  if (i >= n) return;
  const int64_t k = min((int64_t)keys[i], nseg);
  const int64_t kp = (i == 0) ? -1 : min((int64_t)keys[i - 1], nseg);
  for (int64_t s = kp + 1; s <= k; ++s) ptr[s] = i;
  if (i == n - 1)
    for (int64_t s = k + 1; s <= nseg; ++s) ptr[s] = n;
}
__global__ void k_slot_keys(const int32_t* f0, const int32_t* f1, int64_t nb, int32_t* keys, int32_t* vals) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  keys[2 * i] = f0[i] >= 0 ? f0[i] : INT32_MAX;
  vals[2 * i] = (int32_t)(i << 1);
  keys[2 * i + 1] = f1[i] >= 0 ? f1[i] : INT32_MAX;
  vals[2 * i + 1] = (int32_t)(i << 1) | 1;
}
__global__ void k_ef_keys(const int32_t* e, const int32_t* f0, const int32_t* f1, int64_t nb, int64_t nf, uint64_t* keys) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  keys[2 * i] = f0[i] >= 0 ? (uint64_t)e[i] * nf + f0[i] : ~0ull;
  keys[2 * i + 1] = f1[i] >= 0 ? (uint64_t)e[i] * nf + f1[i] : ~0ull;
}
__global__ void k_split_ef(const uint64_t* keys, int64_t n, int64_t nf, int32_t* e, int32_t* f) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  e[i] = (int32_t)(keys[i] / nf);
  f[i] = (int32_t)(keys[i] % nf);
}
__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t* a, int64_t n, uint64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__global__ void k_obs_inc(const int32_t* e, const int32_t* f0, const int32_t* f1, int64_t nb, int64_t nf,
                          const uint64_t* uniq, int64_t ninc, int32_t* inc0, int32_t* inc1, int32_t* keys, int32_t* vals) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const int32_t a = f0[i] >= 0 ? (int32_t)lower_bound_u64(uniq, ninc, (uint64_t)e[i] * nf + f0[i]) : -1;
  const int32_t b = f1[i] >= 0 ? (int32_t)lower_bound_u64(uniq, ninc, (uint64_t)e[i] * nf + f1[i]) : -1;
  inc0[i] = a; inc1[i] = b;
  keys[2 * i] = a >= 0 ? a : INT32_MAX; vals[2 * i] = (int32_t)(i << 1);
  keys[2 * i + 1] = b >= 0 ? b : INT32_MAX; vals[2 * i + 1] = (int32_t)(i << 1) | 1;
}
// ordered incidence pairs (i,j) of one e-block with f_i < f_j, or f_i == f_j (both orders, and i == j)
__global__ void k_pair_count(const int64_t* einc_ptr, const int32_t* inc_f, int64_t ne, int64_t* cnt) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > ne) return;
  if (e == ne) { cnt[e] = 0; return; }
  int64_t c = 0;
  for (int64_t i = einc_ptr[e]; i < einc_ptr[e + 1]; ++i)
    for (int64_t j = einc_ptr[e]; j < einc_ptr[e + 1]; ++j) c += (inc_f[i] <= inc_f[j]) ? 1 : 0;
  cnt[e] = c;
}
__global__ void k_pair_fill(const int64_t* einc_ptr, const int32_t* inc_f, int64_t ne, int64_t nf, const int64_t* off,
                            uint64_t* keys, uint64_t* vals) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int64_t o = off[e];
  for (int64_t i = einc_ptr[e]; i < einc_ptr[e + 1]; ++i)
    for (int64_t j = einc_ptr[e]; j < einc_ptr[e + 1]; ++j)
      if (inc_f[i] <= inc_f[j]) {
        keys[o] = (uint64_t)inc_f[i] * nf + inc_f[j];
        vals[o] = (uint64_t)(uint32_t)i | ((uint64_t)(uint32_t)j << 32);
        ++o;
      }
}
__global__ void k_pair_diag_sentinels(int64_t nf, int64_t base, uint64_t* keys, uint64_t* vals) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nf) return;
  keys[base + f] = (uint64_t)f * nf + f;
  vals[base + f] = ~0ull;  // i = j = -1: contributes nothing, only guarantees the (f,f) destination exists
}
__global__ void k_dest_finish(const uint64_t* ukeys, int ndest, int64_t nf, int32_t* fa, int32_t* fb, int32_t* diag_dest) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= ndest) return;
  const int32_t a = (int32_t)(ukeys[d] / nf), b = (int32_t)(ukeys[d] % nf);
  fa[d] = a; fb[d] = b;
  if (a == b) diag_dest[a] = d;
}
// ---- destination blocks as a hash set ---------------------------------------------------
// Model A at BAL scale has ~5 incidence pairs per observation (144 M on the 30 M-observation problem) but only ~16
// destination blocks per camera: the set of distinct (fa, fb) keys is collected in an open-addressing table (almost
// every insertion is a read that finds its key), sorted once, and the table then maps key -> destination index for
// the tile builder of the fused path.  The full per-destination pair lists are only needed by the generic pipeline.
constexpr uint64_t DH_EMPTY = ~0ull;
__device__ __forceinline__ uint64_t dh_slot(uint64_t key, int shift) { return (key * 0x9E3779B97F4A7C15ull) >> shift; }
__device__ __forceinline__ bool dh_insert(uint64_t* table, uint64_t mask, int shift, uint64_t key, unsigned long long* count) {
  uint64_t h = dh_slot(key, shift);
  for (uint64_t probe = 0; probe <= mask; ++probe, h = (h + 1) & mask) {
    const uint64_t cur = __ldcg(table + h);
    if (cur == key) return true;
    if (cur == DH_EMPTY) {
      const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(table + h), (unsigned long long)DH_EMPTY, (unsigned long long)key);
      if (old == DH_EMPTY) { atomicAdd(count, 1ull); return true; }
      if (old == key) return true;
    }
  }
  return false;
}
__device__ __forceinline__ int32_t dh_find(const uint64_t* __restrict__ table, const int32_t* __restrict__ val, uint64_t mask, int shift,
                                           uint64_t key) {
  uint64_t h = dh_slot(key, shift);
  for (uint64_t probe = 0; probe <= mask; ++probe, h = (h + 1) & mask) {
    const uint64_t cur = table[h];
    if (cur == key) return val[h];
    if (cur == DH_EMPTY) return -1;
  }
  return -1;
}
__global__ void k_dh_insert_pairs(const int64_t* __restrict__ einc_ptr, const int32_t* __restrict__ inc_f, int64_t ne, int64_t nf,
                                  uint64_t* table, uint64_t mask, int shift, unsigned long long* count, int* overflow) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int64_t b = einc_ptr[e], n = einc_ptr[e + 1];
  for (int64_t i = b; i < n; ++i) {
    const int32_t fi = inc_f[i];
    for (int64_t j = b; j < n; ++j) {
      const int32_t fj = inc_f[j];
      if (fi <= fj && !dh_insert(table, mask, shift, (uint64_t)fi * nf + fj, count)) *overflow = 1;
    }
  }
}
__global__ void k_dh_insert_diag(int64_t nf, uint64_t* table, uint64_t mask, int shift, unsigned long long* count, int* overflow) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f < nf && !dh_insert(table, mask, shift, (uint64_t)f * nf + f, count)) *overflow = 1;   // the (f,f) destination always exists
}
__global__ void k_dh_map(const uint64_t* __restrict__ dest_keys, int ndest, const uint64_t* __restrict__ table, uint64_t mask, int shift,
                         int32_t* __restrict__ val) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= ndest) return;
  uint64_t h = dh_slot(dest_keys[d], shift);
  while (table[h] != dest_keys[d]) h = (h + 1) & mask;
  val[h] = d;
}

__global__ void k_unpack_pairs(const uint64_t* vals, int64_t n, int2* pairs) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  pairs[i] = make_int2((int32_t)(uint32_t)(vals[i] & 0xffffffffu), (int32_t)(uint32_t)(vals[i] >> 32));
}
__global__ void k_ff_keys(const int32_t* f0, const int32_t* f1, int64_t nb, int64_t nf, uint64_t* keys) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  keys[i] = (f0[i] >= 0 && f1[i] >= 0) ? (uint64_t)f0[i] * nf + f1[i] : ~0ull;
}
__global__ void k_dobs_ptr(const uint64_t* sorted_keys, int64_t nb, const uint64_t* dest_keys, int ndest, int64_t* ptr) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d > ndest) return;
  ptr[d] = d < ndest ? lower_bound_u64(sorted_keys, nb, dest_keys[d]) : lower_bound_u64(sorted_keys, nb, ~0ull);
}
__global__ void k_chunk_count(const int64_t* ptr, int nseg, int ch, int32_t* cnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nseg) return;
  cnt[s] = s < nseg ? (int32_t)((ptr[s + 1] - ptr[s] + ch - 1) / ch) : 0;
}
__global__ void k_chunk_fill(const int64_t* ptr, int nseg, int ch, const int32_t* first, int32_t* seg, int64_t* begin) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  int64_t b = ptr[s];
  for (int c = first[s]; c < first[s + 1]; ++c, b += ch) { seg[c] = s; begin[c] = b; }
}

inline int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}

inline int build_chunks(Chunks& C, const int64_t* ptr, int nseg, int ch, cudaStream_t st) {
  C.nseg = nseg; C.ch = ch;
  DVec<int32_t> cnt;
  BA_TRY(cnt.alloc(nseg + 1));
  BA_TRY(C.seg_first.alloc(nseg + 1));
  k_chunk_count<<<grid_for(nseg + 1, 256), 256, 0, st>>>(ptr, nseg, ch, cnt.p);
  BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt.p, C.seg_first.p, nseg + 1, st); }));
  int32_t total = 0;
  BA_CUDA_TRY(cudaMemcpyAsync(&total, C.seg_first.p + nseg, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  BA_CUDA_TRY(cudaStreamSynchronize(st));
  C.n = total;
  BA_TRY(C.seg.alloc(total));
  BA_TRY(C.begin.alloc(total));
  if (nseg > 0) k_chunk_fill<<<grid_for(nseg, 256), 256, 0, st>>>(ptr, nseg, ch, C.seg_first.p, C.seg.p, C.begin.p);
  BA_CUDA_TRY(cudaGetLastError());
  return BA_OK;
}

// sorts (key,val) int32 pairs where invalid entries carry key INT32_MAX; returns CSR over [0,nseg)
inline int sort_to_csr(const int32_t* keys_in, const int32_t* vals_in, int64_t n, int64_t nseg, DVec<int64_t>& ptr,
                       DVec<int32_t>& vals_out, cudaStream_t st) {
  DVec<int32_t> keys_sorted;
  BA_TRY(keys_sorted.alloc(n));
  BA_TRY(vals_out.alloc(n));
  BA_TRY(ptr.alloc(nseg + 1));
  if (n > 0)
    BA_TRY(cub_call([&](void* t, size_t& b) {
      return cub::DeviceRadixSort::SortPairs(t, b, keys_in, keys_sorted.p, vals_in, vals_out.p, (int)n, 0, 32, st);
    }));
  k_seg_ptr<int32_t><<<grid_for(n > nseg + 1 ? n : nseg + 1, 256), 256, 0, st>>>(keys_sorted.p, n, nseg, ptr.p);
  BA_CUDA_TRY(cudaGetLastError());
  return BA_OK;
}

// destination blocks (fa <= fb) with the list of incidence pairs contributing to each: every ordered pair is
// materialised and sorted by destination key.  Model B always; Model A only for the generic pipeline (on demand).
inline int build_pair_lists(Structure& S, cudaStream_t st) {
  const int64_t ne = S.ne, nf = S.nf;
  const int B = 256;
  DVec<uint64_t> dest_keys;
  {
    DVec<int64_t> cnt, off;
    BA_TRY(cnt.alloc(ne + 1)); BA_TRY(off.alloc(ne + 1));
    k_pair_count<<<grid_for(ne + 1, 128), 128, 0, st>>>(S.einc_ptr, S.inc_f, ne, cnt.p);
    BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt.p, off.p, (int)(ne + 1), st); }));
    int64_t np = 0;
    BA_CUDA_TRY(cudaMemcpyAsync(&np, off.p + ne, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    BA_CUDA_TRY(cudaStreamSynchronize(st));
    S.npairs = np + nf;
    if (S.npairs >= (int64_t)INT32_MAX) return fail(BA_ERR_UNSUPPORTED, "too many incidence pairs (%lld) for one GPU", (long long)S.npairs);
    DVec<uint64_t> pk, pv, pks, pvs;
    BA_TRY(pk.alloc(S.npairs)); BA_TRY(pv.alloc(S.npairs)); BA_TRY(pks.alloc(S.npairs)); BA_TRY(pvs.alloc(S.npairs));
    if (ne > 0) k_pair_fill<<<grid_for(ne, 128), 128, 0, st>>>(S.einc_ptr, S.inc_f, ne, nf, off.p, pk.p, pv.p);
    k_pair_diag_sentinels<<<grid_for(nf, B), B, 0, st>>>(nf, np, pk.p, pv.p);
    BA_TRY(cub_call([&](void* t, size_t& b) {
      return cub::DeviceRadixSort::SortPairs(t, b, pk.p, pks.p, pv.p, pvs.p, (int)S.npairs, 0, bits_for((uint64_t)nf * nf), st);
    }));
    pk.release(); pv.release();
    DVec<int64_t> run_cnt;
    DVec<int32_t> nruns;
    BA_TRY(dest_keys.alloc(S.npairs)); BA_TRY(run_cnt.alloc(S.npairs + 1)); BA_TRY(nruns.alloc(1));
    BA_TRY(cub_call([&](void* t, size_t& b) {
      return cub::DeviceRunLengthEncode::Encode(t, b, pks.p, dest_keys.p, run_cnt.p, nruns.p, (int)S.npairs, st);
    }));
    int32_t nd = 0;
    BA_CUDA_TRY(cudaMemcpyAsync(&nd, nruns.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    BA_CUDA_TRY(cudaStreamSynchronize(st));
    const bool have_dest = S.dest_keys.n != 0;   // the hashed set exists already: same keys, same (ascending) order
    if (have_dest && nd != S.ndest) return fail(BA_ERR_CUDA, "destination sets disagree (%d hashed, %d sorted)", S.ndest, nd);
    S.ndest = nd;
    BA_TRY(S.dpair_ptr.alloc(nd + 1));
    BA_CUDA_TRY(cudaMemsetAsync(run_cnt.p + nd, 0, sizeof(int64_t), st));
    BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, run_cnt.p, S.dpair_ptr.p, nd + 1, st); }));
    if (!have_dest) {
      BA_TRY(S.dest_fa.alloc(nd)); BA_TRY(S.dest_fb.alloc(nd)); BA_TRY(S.diag_dest.alloc(nf));
      k_dest_finish<<<grid_for(nd, B), B, 0, st>>>(dest_keys.p, nd, nf, S.dest_fa.p, S.dest_fb.p, S.diag_dest.p);
      BA_TRY(S.dest_keys.alloc(nd));   // exactly ndest keys
      BA_CUDA_TRY(cudaMemcpyAsync(S.dest_keys.p, dest_keys.p, sizeof(uint64_t) * nd, cudaMemcpyDeviceToDevice, st));
    }
    BA_TRY(S.pairs.alloc(S.npairs));
    k_unpack_pairs<<<grid_for(S.npairs, B), B, 0, st>>>(pvs.p, S.npairs, S.pairs.p);
    BA_CUDA_TRY(cudaStreamSynchronize(st));
  }
  S.pair_lists = true;
  return BA_OK;
}

// destination blocks only (see "destination blocks as a hash set" above); nslots == 1
inline int build_dest_hashed(Structure& S, cudaStream_t st) {
  const int64_t ne = S.ne, nf = S.nf;
  const int B = 256;
  int64_t np = 0;
  {
    DVec<int64_t> cnt, off;
    BA_TRY(cnt.alloc(ne + 1)); BA_TRY(off.alloc(ne + 1));
    k_pair_count<<<grid_for(ne + 1, 128), 128, 0, st>>>(S.einc_ptr, S.inc_f, ne, cnt.p);
    BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt.p, off.p, (int)(ne + 1), st); }));
    BA_CUDA_TRY(cudaMemcpyAsync(&np, off.p + ne, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    BA_CUDA_TRY(cudaStreamSynchronize(st));
  }
  S.npairs = np + nf;
  const int64_t bound = std::min<int64_t>(np + nf, nf < ((int64_t)1 << 20) ? nf * (nf + 1) / 2 : INT64_MAX);   // distinct keys at most
  int64_t cap = 1024;
  while (cap < 2 * std::min<int64_t>(bound, 48 * nf)) cap <<= 1;
  DVec<unsigned long long> count;
  DVec<int> overflow;
  BA_TRY(count.alloc(1)); BA_TRY(overflow.alloc(1));
  unsigned long long h_count = 0;
  for (;;) {
    int bits = 0;
    while (((int64_t)1 << bits) < cap) ++bits;
    S.dh_mask = (uint64_t)cap - 1; S.dh_shift = 64 - bits;
    BA_TRY(S.dh_keys.alloc((size_t)cap));
    BA_CUDA_TRY(cudaMemsetAsync(S.dh_keys.p, 0xff, sizeof(uint64_t) * cap, st));
    BA_CUDA_TRY(cudaMemsetAsync(count.p, 0, sizeof(unsigned long long), st));
    BA_CUDA_TRY(cudaMemsetAsync(overflow.p, 0, sizeof(int), st));
    if (ne > 0)
      k_dh_insert_pairs<<<grid_for(ne, 128), 128, 0, st>>>(S.einc_ptr, S.inc_f, ne, nf, S.dh_keys.p, S.dh_mask, S.dh_shift, count.p, overflow.p);
    k_dh_insert_diag<<<grid_for(nf, B), B, 0, st>>>(nf, S.dh_keys.p, S.dh_mask, S.dh_shift, count.p, overflow.p);
    int h_over = 0;
    BA_CUDA_TRY(cudaMemcpyAsync(&h_count, count.p, sizeof(h_count), cudaMemcpyDeviceToHost, st));
    BA_CUDA_TRY(cudaMemcpyAsync(&h_over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    BA_CUDA_TRY(cudaStreamSynchronize(st));
    if (!h_over && (int64_t)h_count * 10 <= cap * 7) break;     // load factor <= 0.7
    if (cap >= 4 * bound + 1024) return fail(BA_ERR_CUDA, "destination hash set overflowed at capacity %lld", (long long)cap);
    cap <<= 3;
  }
  if ((int64_t)h_count >= (int64_t)INT32_MAX) return fail(BA_ERR_UNSUPPORTED, "reduced camera system pattern too large");
  S.ndest = (int)h_count;
  {  // ascending keys: the empty slots (~0) sort to the end
    DVec<uint64_t> sorted;
    BA_TRY(sorted.alloc((size_t)cap));
    BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceRadixSort::SortKeys(t, b, S.dh_keys.p, sorted.p, (int)cap, 0, 64, st); }));
    BA_TRY(S.dest_keys.alloc(S.ndest));
    BA_CUDA_TRY(cudaMemcpyAsync(S.dest_keys.p, sorted.p, sizeof(uint64_t) * S.ndest, cudaMemcpyDeviceToDevice, st));
    BA_CUDA_TRY(cudaStreamSynchronize(st));
  }
  BA_TRY(S.dh_val.alloc((size_t)cap));
  k_dh_map<<<grid_for(S.ndest, B), B, 0, st>>>(S.dest_keys.p, S.ndest, S.dh_keys.p, S.dh_mask, S.dh_shift, S.dh_val.p);
  BA_TRY(S.dest_fa.alloc(S.ndest)); BA_TRY(S.dest_fb.alloc(S.ndest)); BA_TRY(S.diag_dest.alloc(nf));
  k_dest_finish<<<grid_for(S.ndest, B), B, 0, st>>>(S.dest_keys.p, S.ndest, nf, S.dest_fa.p, S.dest_fb.p, S.diag_dest.p);
  BA_CUDA_TRY(cudaGetLastError());
  S.pair_lists = false;
  return BA_OK;
}

// the pair lists and their chunk table for a structure that was built without them (Model A, generic pipeline)
inline int ensure_pair_lists(Structure& S, cudaStream_t st) {
  if (S.pair_lists) return BA_OK;
  BA_TRY(build_pair_lists(S, st));
  BA_TRY(build_chunks(S.ch_pairs, S.dpair_ptr.p, S.ndest, 256, st));
  BA_CUDA_TRY(cudaStreamSynchronize(st));
  return BA_OK;
}

// h_e / h_f0 / h_f1 are in the caller's observation order; h_f1 may be NULL (one f slot).
template <typename Validate>
inline int build_structure(Structure& S, int64_t nb, int64_t ne, int64_t nf, const int32_t* h_e, const int32_t* h_f0,
                           const int32_t* h_f1, cudaStream_t st, Validate&& validate) {
  if (nb >= (int64_t)1 << 30) return fail(BA_ERR_UNSUPPORTED, "more than 2^30 residual blocks per GPU are not supported");
  S.nb = nb; S.ne = ne; S.nf = nf; S.nslots = h_f1 ? 2 : 1;
  const int B = 256;
  DVec<int32_t> e_in, f0_in, f1_in, iota;
  BA_TRY(e_in.upload(h_e, nb, st));
  BA_TRY(f0_in.upload(h_f0, nb, st));
  if (h_f1) BA_TRY(f1_in.upload(h_f1, nb, st));
  BA_TRY(validate(e_in.p, f0_in.p));
  BA_TRY(iota.alloc(nb));
  k_iota<<<grid_for(nb, B), B, 0, st>>>(iota.p, nb, 0);
  // 1. observations sorted by e (stable)
  BA_TRY(S.ob_e.alloc(nb)); BA_TRY(S.perm.alloc(nb)); BA_TRY(S.ob_f0.alloc(nb)); BA_TRY(S.ob_f1.alloc(nb));
  if (nb > 0)
    BA_TRY(cub_call([&](void* t, size_t& b) {
      return cub::DeviceRadixSort::SortPairs(t, b, e_in.p, S.ob_e.p, iota.p, S.perm.p, (int)nb, 0, bits_for((uint64_t)ne), st);
    }));
  k_gather_i32<<<grid_for(nb, B), B, 0, st>>>(S.ob_f0.p, f0_in.p, S.perm.p, nb);
  k_gather_i32<<<grid_for(nb, B), B, 0, st>>>(S.ob_f1.p, h_f1 ? f1_in.p : nullptr, S.perm.p, nb);
  BA_TRY(S.e_ptr.alloc(ne + 1));
  k_seg_ptr<int32_t><<<grid_for(nb > ne + 1 ? nb : ne + 1, B), B, 0, st>>>(S.ob_e.p, nb, ne, S.e_ptr.p);
  BA_CUDA_TRY(cudaGetLastError());

  // 2. incidences
  if (S.nslots == 1) {
    S.ninc = nb;
    S.inc_e = S.ob_e.p; S.inc_f = S.ob_f0.p; S.ob_inc0 = nullptr; S.einc_ptr = S.e_ptr.p;
    BA_TRY(S.ob_inc1.alloc(0));
  } else {
    DVec<uint64_t> k2, k2s, uniq;
    DVec<int64_t> nuniq;
    BA_TRY(k2.alloc(2 * nb)); BA_TRY(k2s.alloc(2 * nb)); BA_TRY(uniq.alloc(2 * nb)); BA_TRY(nuniq.alloc(1));
    k_ef_keys<<<grid_for(nb, B), B, 0, st>>>(S.ob_e.p, S.ob_f0.p, S.ob_f1.p, nb, nf, k2.p);
    int64_t n_unique = 0;
    uint64_t last = 0;
    if (nb > 0) {
      BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceRadixSort::SortKeys(t, b, k2.p, k2s.p, (int)(2 * nb), 0, 64, st); }));
      BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceSelect::Unique(t, b, k2s.p, uniq.p, nuniq.p, (int)(2 * nb), st); }));
      BA_CUDA_TRY(cudaMemcpyAsync(&n_unique, nuniq.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      BA_CUDA_TRY(cudaStreamSynchronize(st));
      BA_CUDA_TRY(cudaMemcpyAsync(&last, uniq.p + (n_unique - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      BA_CUDA_TRY(cudaStreamSynchronize(st));
    }
    S.ninc = n_unique - ((n_unique > 0 && last == ~0ull) ? 1 : 0);
    BA_TRY(S.own_inc_e.alloc(S.ninc)); BA_TRY(S.own_inc_f.alloc(S.ninc));
    k_split_ef<<<grid_for(S.ninc, B), B, 0, st>>>(uniq.p, S.ninc, nf, S.own_inc_e.p, S.own_inc_f.p);
    BA_TRY(S.own_ob_inc0.alloc(nb)); BA_TRY(S.ob_inc1.alloc(nb));
    DVec<int32_t> ik, iv;
    BA_TRY(ik.alloc(2 * nb)); BA_TRY(iv.alloc(2 * nb));
    k_obs_inc<<<grid_for(nb, B), B, 0, st>>>(S.ob_e.p, S.ob_f0.p, S.ob_f1.p, nb, nf, uniq.p, S.ninc, S.own_ob_inc0.p, S.ob_inc1.p, ik.p, iv.p);
    BA_TRY(sort_to_csr(ik.p, iv.p, 2 * nb, S.ninc, S.incobs_ptr, S.incobs, st));
    BA_TRY(S.own_einc_ptr.alloc(ne + 1));
    k_seg_ptr<int32_t><<<grid_for(S.ninc > ne + 1 ? S.ninc : ne + 1, B), B, 0, st>>>(S.own_inc_e.p, S.ninc, ne, S.own_einc_ptr.p);
    S.inc_e = S.own_inc_e.p; S.inc_f = S.own_inc_f.p; S.ob_inc0 = S.own_ob_inc0.p; S.einc_ptr = S.own_einc_ptr.p;
    BA_CUDA_TRY(cudaStreamSynchronize(st));  // uniq etc. go out of scope
  }
  if (S.ninc >= (int64_t)1 << 30) return fail(BA_ERR_UNSUPPORTED, "too many incidences");

  // 3. f -> incidences, f -> observations
  {
    DVec<int32_t> inc_iota;
    BA_TRY(inc_iota.alloc(S.ninc));
    k_iota<<<grid_for(S.ninc, B), B, 0, st>>>(inc_iota.p, S.ninc, 0);
    BA_TRY(sort_to_csr(S.inc_f, inc_iota.p, S.ninc, nf, S.finc_ptr, S.finc, st));
    if (S.nslots == 1) {
      DVec<int32_t> v;
      BA_TRY(v.alloc(nb));
      k_iota<<<grid_for(nb, B), B, 0, st>>>(v.p, nb, 1);
      BA_TRY(sort_to_csr(S.ob_f0.p, v.p, nb, nf, S.fobs_ptr, S.fobs, st));
      BA_CUDA_TRY(cudaStreamSynchronize(st));
    } else {
      DVec<int32_t> k, v;
      BA_TRY(k.alloc(2 * nb)); BA_TRY(v.alloc(2 * nb));
      k_slot_keys<<<grid_for(nb, B), B, 0, st>>>(S.ob_f0.p, S.ob_f1.p, nb, k.p, v.p);
      BA_TRY(sort_to_csr(k.p, v.p, 2 * nb, nf, S.fobs_ptr, S.fobs, st));
      BA_CUDA_TRY(cudaStreamSynchronize(st));
    }
  }

  // 4. destination blocks of the reduced system (and, for the generic pipeline, their incidence-pair lists)
  if (S.nslots == 1) BA_TRY(build_dest_hashed(S, st));
  else BA_TRY(build_pair_lists(S, st));
  // 5. Model B: observations whose (f0,f1) is a destination block (off-diagonal F^T F terms)
  if (S.nslots == 2) {
    DVec<uint64_t> k, ks;
    DVec<int32_t> v;
    BA_TRY(k.alloc(nb)); BA_TRY(ks.alloc(nb)); BA_TRY(v.alloc(nb)); BA_TRY(S.dobs.alloc(nb));
    k_ff_keys<<<grid_for(nb, B), B, 0, st>>>(S.ob_f0.p, S.ob_f1.p, nb, nf, k.p);
    k_iota<<<grid_for(nb, B), B, 0, st>>>(v.p, nb, 0);
    if (nb > 0)
      BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceRadixSort::SortPairs(t, b, k.p, ks.p, v.p, S.dobs.p, (int)nb, 0, 64, st); }));
    BA_TRY(S.dobs_ptr.alloc(S.ndest + 1));
    k_dobs_ptr<<<grid_for(S.ndest + 1, B), B, 0, st>>>(ks.p, nb, S.dest_keys.p, S.ndest, S.dobs_ptr.p);
    BA_CUDA_TRY(cudaStreamSynchronize(st));
  }
  // 6. chunk tables of the gather reductions
  BA_TRY(build_chunks(S.ch_fobs, S.fobs_ptr.p, (int)nf, 256, st));
  BA_TRY(build_chunks(S.ch_finc, S.finc_ptr.p, (int)nf, 512, st));
  if (S.pair_lists) BA_TRY(build_chunks(S.ch_pairs, S.dpair_ptr.p, S.ndest, 256, st));
  if (S.nslots == 2) BA_TRY(build_chunks(S.ch_dobs, S.dobs_ptr.p, S.ndest, 128, st));
  BA_CUDA_TRY(cudaStreamSynchronize(st));
  BA_CUDA_TRY(cudaGetLastError());
  return BA_OK;
}

// ---------------------------------------------------------------------------------------
// The same structure built on the HOST, for rig-size two-slot problems (the reference's own: tens to a few thousand marker
// observations).  build_structure() above is made for 30 M observations: a dozen radix sorts and scans, each a launch or
// three, and ten stream synchronisations to read counts back -- 0.46 ms for hongo's 68 observations, as long as the whole
// solve.  Here every list is made by the same rules (stable sorts by the same keys, so the orders -- and with them the
// summation orders of every kernel -- are identical to the device build's, bit for bit), laid out in ONE host buffer,
// uploaded with ONE copy into ONE allocation (S.arena); the DVecs of the structure are views into it.
// ---------------------------------------------------------------------------------------
struct HostArena {
  std::vector<unsigned char> buf;
  template <typename T>
  size_t put(const std::vector<T>& v) {   // returns the offset
    const size_t off = (buf.size() + 255) / 256 * 256;
    buf.resize(off + std::max<size_t>(v.size(), 1) * sizeof(T), 0);
    if (!v.empty()) std::memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
    return off;
  }
};

inline void host_csr(const std::vector<int32_t>& sorted_keys, int64_t nseg, std::vector<int64_t>& ptr) {   // k_seg_ptr
  ptr.assign(nseg + 1, 0);
  size_t pos = 0;
  for (int64_t s2 = 0; s2 <= nseg; ++s2) {
    while (pos < sorted_keys.size() && (int64_t)sorted_keys[pos] < s2) ++pos;
    ptr[s2] = (int64_t)pos;
  }
}
// sort_to_csr on the host: stable by key, invalid entries (INT32_MAX) at the tail
inline void host_sort_to_csr(const std::vector<int32_t>& keys, const std::vector<int32_t>& vals, int64_t nseg, std::vector<int64_t>& ptr,
                             std::vector<int32_t>& vals_out) {
  std::vector<int32_t> order(keys.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return keys[a] < keys[b]; });
  std::vector<int32_t> ks(keys.size());
  vals_out.resize(keys.size());
  for (size_t i = 0; i < order.size(); ++i) { ks[i] = keys[order[i]]; vals_out[i] = vals[order[i]]; }
  host_csr(ks, nseg, ptr);
}
struct HostChunks { int n = 0; std::vector<int32_t> seg, seg_first; std::vector<int64_t> begin; };
inline void host_chunks(const std::vector<int64_t>& ptr, int nseg, int ch, HostChunks& C) {   // build_chunks
  C.seg_first.assign(nseg + 1, 0);
  for (int s2 = 0; s2 < nseg; ++s2) C.seg_first[s2 + 1] = C.seg_first[s2] + (int32_t)((ptr[s2 + 1] - ptr[s2] + ch - 1) / ch);
  C.n = C.seg_first[nseg];
  C.seg.resize(C.n); C.begin.resize(C.n);
  for (int s2 = 0; s2 < nseg; ++s2) {
    int64_t b = ptr[s2];
    for (int c = C.seg_first[s2]; c < C.seg_first[s2 + 1]; ++c, b += ch) { C.seg[c] = s2; C.begin[c] = b; }
  }
}

// h_f1 == NULL: one f slot per observation (Model A); the incidences then alias the observations and the pair lists of the
// generic pipeline (built on demand by the device path) come with the structure.
inline int build_structure_host(Structure& S, int64_t nb, int64_t ne, int64_t nf, const int32_t* h_e, const int32_t* h_f0,
                                const int32_t* h_f1, cudaStream_t st) {
  const bool two = h_f1 != nullptr;
  S.nb = nb; S.ne = ne; S.nf = nf; S.nslots = two ? 2 : 1;
  // 1. observations sorted by e (stable)
  std::vector<int32_t> perm(nb), ob_e(nb), ob_f0(nb), ob_f1(nb);
  for (int64_t i = 0; i < nb; ++i) perm[i] = (int32_t)i;
  std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return h_e[a] < h_e[b]; });
  for (int64_t i = 0; i < nb; ++i) { ob_e[i] = h_e[perm[i]]; ob_f0[i] = h_f0[perm[i]]; ob_f1[i] = two ? h_f1[perm[i]] : -1; }
  std::vector<int64_t> e_ptr;
  host_csr(ob_e, ne, e_ptr);
  // 2. incidences: the distinct (e, f) of the two slots, ascending
  std::vector<uint64_t> uniq;
  std::vector<int32_t> inc_e, inc_f, ob_inc0, ob_inc1, incobs;
  std::vector<int64_t> incobs_ptr, einc_ptr;
  int64_t ninc = nb;
  if (two) {
    uniq.reserve(2 * nb);
    for (int64_t i = 0; i < nb; ++i) {
      if (ob_f0[i] >= 0) uniq.push_back((uint64_t)ob_e[i] * nf + ob_f0[i]);
      if (ob_f1[i] >= 0) uniq.push_back((uint64_t)ob_e[i] * nf + ob_f1[i]);
    }
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    ninc = (int64_t)uniq.size();
    inc_e.resize(ninc); inc_f.resize(ninc); ob_inc0.resize(nb); ob_inc1.resize(nb);
    std::vector<int32_t> ik(2 * nb), iv(2 * nb);
    for (int64_t i = 0; i < ninc; ++i) { inc_e[i] = (int32_t)(uniq[i] / nf); inc_f[i] = (int32_t)(uniq[i] % nf); }
    auto find_inc = [&](uint64_t key) { return (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), key) - uniq.begin()); };
    for (int64_t i = 0; i < nb; ++i) {
      const int32_t a = ob_f0[i] >= 0 ? find_inc((uint64_t)ob_e[i] * nf + ob_f0[i]) : -1;
      const int32_t b = ob_f1[i] >= 0 ? find_inc((uint64_t)ob_e[i] * nf + ob_f1[i]) : -1;
      ob_inc0[i] = a; ob_inc1[i] = b;
      ik[2 * i] = a >= 0 ? a : INT32_MAX; iv[2 * i] = (int32_t)(i << 1);
      ik[2 * i + 1] = b >= 0 ? b : INT32_MAX; iv[2 * i + 1] = (int32_t)(i << 1) | 1;
    }
    host_sort_to_csr(ik, iv, ninc, incobs_ptr, incobs);
    host_csr(inc_e, ne, einc_ptr);
  } else {   // the incidences ARE the observations
    inc_e = ob_e; inc_f = ob_f0; einc_ptr = e_ptr;
  }
  S.ninc = ninc;
  // 3. f -> incidences, f -> observations
  std::vector<int32_t> inc_iota(ninc), finc, fobs;
  for (int64_t i = 0; i < ninc; ++i) inc_iota[i] = (int32_t)i;
  std::vector<int64_t> finc_ptr, fobs_ptr;
  host_sort_to_csr(inc_f, inc_iota, nf, finc_ptr, finc);
  if (two) {
    std::vector<int32_t> fk(2 * nb), fv(2 * nb);
    for (int64_t i = 0; i < nb; ++i) {
      fk[2 * i] = ob_f0[i] >= 0 ? ob_f0[i] : INT32_MAX; fv[2 * i] = (int32_t)(i << 1);
      fk[2 * i + 1] = ob_f1[i] >= 0 ? ob_f1[i] : INT32_MAX; fv[2 * i + 1] = (int32_t)(i << 1) | 1;
    }
    host_sort_to_csr(fk, fv, nf, fobs_ptr, fobs);
  } else {
    std::vector<int32_t> fv(nb);
    for (int64_t i = 0; i < nb; ++i) fv[i] = (int32_t)(i << 1);
    host_sort_to_csr(ob_f0, fv, nf, fobs_ptr, fobs);
  }
  // 4. destination blocks with their incidence-pair lists (build_pair_lists)
  std::vector<uint64_t> pk, pv;
  for (int64_t e = 0; e < ne; ++e)
    for (int64_t i = einc_ptr[e]; i < einc_ptr[e + 1]; ++i)
      for (int64_t j = einc_ptr[e]; j < einc_ptr[e + 1]; ++j)
        if (inc_f[i] <= inc_f[j]) {
          pk.push_back((uint64_t)inc_f[i] * nf + inc_f[j]);
          pv.push_back((uint64_t)(uint32_t)i | ((uint64_t)(uint32_t)j << 32));
        }
  for (int64_t f = 0; f < nf; ++f) { pk.push_back((uint64_t)f * nf + f); pv.push_back(~0ull); }
  const int64_t npairs = (int64_t)pk.size();
  S.npairs = npairs;
  std::vector<int64_t> order(npairs);
  for (int64_t i = 0; i < npairs; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return pk[a] < pk[b]; });
  std::vector<uint64_t> dest_keys;
  std::vector<int64_t> dpair_ptr;
  std::vector<int2> pairs(npairs);
  for (int64_t i = 0; i < npairs; ++i) {
    const uint64_t k = pk[order[i]], v = pv[order[i]];
    if (dest_keys.empty() || dest_keys.back() != k) { dest_keys.push_back(k); dpair_ptr.push_back(i); }
    pairs[i] = make_int2((int32_t)(uint32_t)(v & 0xffffffffu), (int32_t)(uint32_t)(v >> 32));
  }
  dpair_ptr.push_back(npairs);
  const int ndest = (int)dest_keys.size();
  S.ndest = ndest;
  std::vector<int32_t> dest_fa(ndest), dest_fb(ndest), diag_dest(nf, 0);
  for (int d = 0; d < ndest; ++d) {
    dest_fa[d] = (int32_t)(dest_keys[d] / nf); dest_fb[d] = (int32_t)(dest_keys[d] % nf);
    if (dest_fa[d] == dest_fb[d]) diag_dest[dest_fa[d]] = d;
  }
  // 5. observations whose (f0, f1) is a destination block
  std::vector<int32_t> dobs;
  std::vector<int64_t> dobs_ptr;
  if (two) {
    std::vector<uint64_t> ffk(nb), ffs(nb);
    dobs.resize(nb);
    for (int64_t i = 0; i < nb; ++i) { ffk[i] = (ob_f0[i] >= 0 && ob_f1[i] >= 0) ? (uint64_t)ob_f0[i] * nf + ob_f1[i] : ~0ull; dobs[i] = (int32_t)i; }
    std::stable_sort(dobs.begin(), dobs.end(), [&](int32_t a, int32_t b) { return ffk[a] < ffk[b]; });
    for (int64_t i = 0; i < nb; ++i) ffs[i] = ffk[dobs[i]];
    dobs_ptr.resize(ndest + 1);
    for (int d = 0; d <= ndest; ++d)
      dobs_ptr[d] = (int64_t)(std::lower_bound(ffs.begin(), ffs.end(), d < ndest ? dest_keys[d] : ~0ull) - ffs.begin());
  }
  // 6. chunk tables
  HostChunks c_fobs, c_finc, c_pairs, c_dobs;
  host_chunks(fobs_ptr, (int)nf, 256, c_fobs);
  host_chunks(finc_ptr, (int)nf, 512, c_finc);
  host_chunks(dpair_ptr, ndest, 256, c_pairs);
  if (two) host_chunks(dobs_ptr, ndest, 128, c_dobs);

  // which blocks take part (build_activity of ba_cuda.cu, one rank): known here for free
  std::vector<int64_t> f_act(nf + 1, 0);
  for (int64_t f = 0; f < nf; ++f) f_act[f + 1] = f_act[f] + (fobs_ptr[f + 1] > fobs_ptr[f] ? 1 : 0);
  S.host_n_active_f = f_act[nf];
  S.host_n_active_e = 0;
  for (int64_t e = 0; e < ne; ++e) S.host_n_active_e += e_ptr[e + 1] > e_ptr[e] ? 1 : 0;
  // one buffer, one allocation, one copy
  HostArena A;
  const size_t o_fact = A.put(f_act);
  const size_t o_perm = A.put(perm), o_ob_e = A.put(ob_e), o_ob_f0 = A.put(ob_f0), o_ob_f1 = A.put(ob_f1), o_e_ptr = A.put(e_ptr);
  const size_t o_inc_e = A.put(inc_e), o_inc_f = A.put(inc_f), o_inc0 = A.put(ob_inc0), o_inc1 = A.put(ob_inc1), o_einc = A.put(einc_ptr);
  const size_t o_iop = A.put(incobs_ptr), o_io = A.put(incobs), o_fip = A.put(finc_ptr), o_fi = A.put(finc), o_fop = A.put(fobs_ptr), o_fo = A.put(fobs);
  const size_t o_dfa = A.put(dest_fa), o_dfb = A.put(dest_fb), o_dd = A.put(diag_dest), o_dk = A.put(dest_keys);
  const size_t o_dpp = A.put(dpair_ptr), o_pairs = A.put(pairs), o_dop = A.put(dobs_ptr), o_do = A.put(dobs);
  struct CO { size_t seg, begin, first; };
  auto put_chunks = [&](const HostChunks& C) { return CO{A.put(C.seg), A.put(C.begin), A.put(C.seg_first)}; };
  const CO k_fobs = put_chunks(c_fobs), k_finc = put_chunks(c_finc), k_pairs = put_chunks(c_pairs), k_dobs = put_chunks(c_dobs);
  BA_TRY(S.arena.upload(A.buf.data(), A.buf.size(), st));
  unsigned char* base = S.arena.p;
  S.host_f_act_ptr = (const int64_t*)(base + o_fact);
  S.perm.borrow((int32_t*)(base + o_perm), nb); S.ob_e.borrow((int32_t*)(base + o_ob_e), nb);
  S.ob_f0.borrow((int32_t*)(base + o_ob_f0), nb); S.ob_f1.borrow((int32_t*)(base + o_ob_f1), nb);
  S.e_ptr.borrow((int64_t*)(base + o_e_ptr), ne + 1);
  if (two) {
    S.own_inc_e.borrow((int32_t*)(base + o_inc_e), ninc); S.own_inc_f.borrow((int32_t*)(base + o_inc_f), ninc);
    S.own_ob_inc0.borrow((int32_t*)(base + o_inc0), nb); S.ob_inc1.borrow((int32_t*)(base + o_inc1), nb);
    S.own_einc_ptr.borrow((int64_t*)(base + o_einc), ne + 1);
    S.inc_e = S.own_inc_e.p; S.inc_f = S.own_inc_f.p; S.ob_inc0 = S.own_ob_inc0.p; S.einc_ptr = S.own_einc_ptr.p;
    S.incobs_ptr.borrow((int64_t*)(base + o_iop), ninc + 1); S.incobs.borrow((int32_t*)(base + o_io), 2 * nb);
  } else {
    S.inc_e = S.ob_e.p; S.inc_f = S.ob_f0.p; S.ob_inc0 = nullptr; S.einc_ptr = S.e_ptr.p;
    S.ob_inc1.borrow((int32_t*)(base + o_inc1), 0);
  }
  S.finc_ptr.borrow((int64_t*)(base + o_fip), nf + 1); S.finc.borrow((int32_t*)(base + o_fi), ninc);
  S.fobs_ptr.borrow((int64_t*)(base + o_fop), nf + 1); S.fobs.borrow((int32_t*)(base + o_fo), 2 * nb);
  S.dest_fa.borrow((int32_t*)(base + o_dfa), ndest); S.dest_fb.borrow((int32_t*)(base + o_dfb), ndest);
  S.diag_dest.borrow((int32_t*)(base + o_dd), nf); S.dest_keys.borrow((uint64_t*)(base + o_dk), ndest);
  S.dpair_ptr.borrow((int64_t*)(base + o_dpp), ndest + 1); S.pairs.borrow((int2*)(base + o_pairs), npairs);
  if (two) { S.dobs_ptr.borrow((int64_t*)(base + o_dop), ndest + 1); S.dobs.borrow((int32_t*)(base + o_do), nb); }
  S.pair_lists = true;
  auto view_chunks = [&](Chunks& C, const HostChunks& H, const CO& o, int nseg, int ch) {
    C.n = H.n; C.nseg = nseg; C.ch = ch;
    C.seg.borrow((int32_t*)(base + o.seg), H.n); C.begin.borrow((int64_t*)(base + o.begin), H.n);
    C.seg_first.borrow((int32_t*)(base + o.first), nseg + 1);
  };
  view_chunks(S.ch_fobs, c_fobs, k_fobs, (int)nf, 256);
  view_chunks(S.ch_finc, c_finc, k_finc, (int)nf, 512);
  view_chunks(S.ch_pairs, c_pairs, k_pairs, ndest, 256);
  if (two) view_chunks(S.ch_dobs, c_dobs, k_dobs, ndest, 128);
  BA_CUDA_TRY(cudaStreamSynchronize(st));   // A.buf goes out of scope
  return BA_OK;
}

}  // namespace ba
