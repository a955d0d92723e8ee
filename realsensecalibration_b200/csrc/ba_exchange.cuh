// ba_exchange.cuh -- multi-GPU exchange of the partial reduced camera system when the shards' contributions are
// (nearly) DISJOINT.
//
// Points are sharded in contiguous ranges; on a problem whose points are ordered along the camera trajectory every rank
// then only touches the destination blocks and cameras of its own stretch (a boundary of one camera window excepted).
// Summing the full block-sparse system with ncclAllReduce moves 2 (N-1)/N of ALL blocks per rank and mostly adds zeros
// (measured round 1, 8 ranks, 30 M observations: 46 MB, 0.38 ms per LM iteration, 19 % of the step).  Here every rank
// sends only the blocks it has -- in its own local order, padded to the longest list -- together with its cameras' sums
// and its shard-local scalars in ONE ncclAllGather, and every rank then adds the pieces in rank order through inverse
// maps built once per problem (k_sx_unpack_*): the same fixed order everywhere, so the replicated LM steps stay bit
// identical across ranks.  Half the bytes of the all-reduce, and the camera-sum collective rides along (3 -> 2
// collectives per LM iteration).  When the shards overlap heavily (sum of the local lists > 1.5 x the union) the
// all-reduce path is kept.
#pragma once
#include "ba_util.cuh"

namespace ba {

struct SparseExchange {
  bool on = false;
  int world = 1, max_nd = 0, max_nc = 0, nc_local = 0, nvc = 0;
  size_t slot = 0;               // doubles per rank: max_nd * 36 | max_nc * nvc | 4 scalars
  DVec<double> buf;              // [world][slot]
  DVec<int32_t> inv_d;           // [world][nd]: local index of global block g on rank r, or -1
  DVec<int32_t> inv_c;           // [world][nf]: row of camera f in rank r's camera list, or -1
  DVec<int32_t> my_cams;         // cameras with observations in this shard, ascending
  DVec<int32_t> send_d;          // local destination blocks that are sent (diagonal blocks of cameras this shard never sees are not)
  int n_send = 0;
};

__global__ void k_sx_flags(const int64_t* __restrict__ fobs_ptr, int64_t nf, int32_t* __restrict__ flag) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f < nf) flag[f] = fobs_ptr[f + 1] > fobs_ptr[f] ? 1 : 0;
}
__global__ void k_sx_compact(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int64_t nf, int32_t* __restrict__ out) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f < nf && flag[f]) out[pos[f]] = (int32_t)f;
}
// inverse maps: lists[r][i] (padded with -1) -> inv[r][lists[r][i]] = i
__global__ void k_sx_invert(int world, int max_n, const int32_t* __restrict__ lists, int64_t n_global, int32_t* __restrict__ inv) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)world * max_n) return;
  const int r = (int)(t / max_n), i = (int)(t % max_n);
  const int32_t g = lists[t];
  if (g >= 0) inv[(int64_t)r * n_global + g] = i;
}
// my slot: the sent destination blocks | my cameras' sums | cost, |g_e|^2, max |g_e|, pad
// a local destination block is sent unless it is the (always present) diagonal block of a camera without local observations
__global__ void k_sx_send_flags(int ndest, const int32_t* __restrict__ fa, const int32_t* __restrict__ fb, const int32_t* __restrict__ cam_flag,
                                int32_t* __restrict__ flag) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < ndest) flag[d] = (fa[d] != fb[d] || cam_flag[fa[d]]) ? 1 : 0;
}
__global__ void k_sx_gather_l2g(int n, const int32_t* __restrict__ sel, const int32_t* __restrict__ l2g, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = l2g[sel[i]];
}
__global__ void k_sx_pack(int n_send, const int32_t* __restrict__ send_d, const double* __restrict__ Pacc, int nc_local, const int32_t* __restrict__ my_cams, int nvc,
                          const double* __restrict__ camacc, const double* __restrict__ scal, int s_cost, int s_g2e, int s_gmaxe,
                          int max_nd, int max_nc, double* __restrict__ slot) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n_p = (int64_t)max_nd * 36, n_c = (int64_t)max_nc * nvc;
  if (t < n_p) {
    const int64_t i = t / 36;
    slot[t] = i < n_send ? Pacc[(int64_t)send_d[i] * 36 + (t - 36 * i)] : 0.0;
  } else if (t < n_p + n_c) {
    const int64_t u = t - n_p;
    const int row = (int)(u / nvc), k = (int)(u % nvc);
    slot[t] = row < nc_local ? camacc[(int64_t)my_cams[row] * nvc + k] : 0.0;
  } else if (t < n_p + n_c + 4) {
    const int k = (int)(t - n_p - n_c);
    slot[t] = k == 0 ? scal[s_cost] : k == 1 ? scal[s_g2e] : k == 2 ? scal[s_gmaxe] : 0.0;
  }
}
// Sb[g] = - sum over the ranks (in rank order) of their block for g
__global__ void k_sx_unpack_blocks(int world, int64_t nd, size_t slot, const double* __restrict__ buf, const int32_t* __restrict__ inv_d,
                                   double* __restrict__ Sb) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nd * 36) return;
  const int64_t g = t / 36;
  const int q = (int)(t % 36);
  double a = 0.0;
  for (int r = 0; r < world; ++r) {
    const int32_t i = inv_d[(int64_t)r * nd + g];
    if (i >= 0) a += buf[(size_t)r * slot + (size_t)i * 36 + q];
  }
  Sb[t] = -a;
}
__global__ void k_sx_unpack_cams(int world, int64_t nf, int nvc, size_t slot, size_t cam_off, const double* __restrict__ buf,
                                 const int32_t* __restrict__ inv_c, double* __restrict__ camacc) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nf * nvc) return;
  const int64_t f = t / nvc;
  const int k = (int)(t % nvc);
  double a = 0.0;
  for (int r = 0; r < world; ++r) {
    const int32_t i = inv_c[(int64_t)r * nf + f];
    if (i >= 0) a += buf[(size_t)r * slot + cam_off + (size_t)i * nvc + k];
  }
  camacc[t] = a;
}
__global__ void k_sx_unpack_scalars(int world, size_t slot, size_t off, const double* __restrict__ buf, double* __restrict__ scal, int s_cost,
                                    int s_g2e, int s_gmaxe) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double c = 0.0, g2 = 0.0, gm = 0.0;
  for (int r = 0; r < world; ++r) {
    const double* s = buf + (size_t)r * slot + off;
    c += s[0]; g2 += s[1]; gm = fmax(gm, s[2]);
  }
  scal[s_cost] = c; scal[s_g2e] = g2; scal[s_gmaxe] = gm;
}
// Jacobi scaling applied to the stored blocks of the global pattern (the FIRST pass exchanges unscaled products)
__global__ void k_sx_scale_blocks(int64_t nd, const int32_t* __restrict__ fa, const int32_t* __restrict__ fb, const double* __restrict__ sf,
                                  double* __restrict__ Sb) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nd * 36) return;
  const int64_t d = t / 36;
  const int v = (int)(t % 36);
  Sb[t] *= sf[6 * (int64_t)fa[d] + v / 6] * sf[6 * (int64_t)fb[d] + v % 6];
}

}  // namespace ba
