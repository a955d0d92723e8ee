// ba_fused_a.cuh -- Model A (camera, point) hot path, B200-first: two point-tile passes per LM iteration in
// which the residual and the analytic Jacobian never leave the SM.
//
// The materialised-Jacobian pipeline (ba_kernels.cuh, still used for Model B) moves ~1.3 kB per observation
// and iteration through HBM; on a B200 (measured ~6.5 TB/s against ~37 TFLOP/s fp64) recomputing the 2x9
// Jacobian of an observation (~150 flops) is cheaper than re-reading its 160 bytes, and the Schur complement's
// pair products are the only heavy arithmetic.  So:
//
//   pass 1  k_fa_pass1   per tile of points (<= 448 observations, <= 256 points):
//             A1  one thread per observation: r, J_e, J_f in registers -> shared memory
//             A2  one thread per point: E^T E, E^T r, LM diagonal, 3x3 Cholesky in registers, z = L^-1 E^T r,
//                 U_i = L^-1 J_e,i^T per observation (shared memory), L and z -> HBM (72 B / point)
//             B   one thread per WORK ITEM (a fixed list of <= 16 observation pairs of one camera pair, or
//                 <= 32 observations of one camera, all inside the tile): operands from shared memory, 36 resp.
//                 33 accumulators in registers, one partial block -> HBM
//   reduce  k_reduce_items   fixed-order sum of the partial blocks per camera pair / camera (two levels)
//   pass 2  k_fa_pass2   per tile: r, J again, back-substitution, Ceres' model cost change, candidate point,
//                        candidate cost -- one pass instead of four
//
// Every list is static (built once per problem with stable radix sorts) and every sum runs in a fixed order:
// deterministic segmented reduction with shared-memory staging, no floating-point atomics.
// Replaces ReprojectionError + AutoDiff (Test1_BundleAdjustment/bundle_adjustmenter.cpp:106-148) and, inside
// Ceres, SchurEliminator::Eliminate / BackSubstitute.
#pragma once
#include "ba_kernels.cuh"
#include "ba_structure.cuh"

namespace ba {

constexpr int FA_TOBS = 384;                 // observation window of a tile
constexpr int FA_KMAX = 64;                  // max observations of one point (fused path)
constexpr int FA_CAP = FA_TOBS + FA_KMAX;    // shared-memory capacity in observations
constexpr int FA_TPTS = 256;                 // max points of a tile
constexpr int FA_CH_PAIR = 16;               // pairs per pair item
constexpr int FA_CH_CAM = 32;                // observations per camera item
constexpr int FA_CH_RED = 64;                // partial blocks per first-level reduction chunk
constexpr int FA_NVC = 33;                   // per camera: 21 packed upper F^T F | 6 F^T r | 6 sum of v_i
constexpr int FA_THREADS = 128;               // 2 CTAs of 128 threads per SM: up to 255 registers for the 6x6 accumulators
constexpr int FA_REC = 22;                    // doubles per observation record in shared memory (pass 1)
constexpr int FA_REC2 = 10;                   // pass 2

// work items of one kind (pair items or camera items) and the static reduction lists over their partial blocks
struct ItemSet {
  int64_t n_ent = 0;
  int n_items = 0, n_targets = 0;
  DVec<int32_t> ent;            // sorted entries (pair: li | lj << 16 ; camera: li)
  DVec<int64_t> item_begin;     // n_items
  DVec<int64_t> item_end;       // n_items
  DVec<int64_t> tile_item_ptr;  // n_tiles + 1
  DVec<int32_t> item_target;    // n_items: destination block / camera
  DVec<int32_t> red_items;      // item ids grouped by target (stable)
  DVec<int64_t> tgt_ptr;        // n_targets + 1 into red_items
  Chunks red_ch;                // chunks over tgt_ptr
};

struct FusedA {
  bool ready = false;
  int n_tiles = 0;
  DVec<int64_t> tile_pt_ptr;    // n_tiles + 1
  ItemSet pairs, cams;
  DVec<double> partP, partC, red1P, red1C, camacc, Lz;
};

__global__ void k_fa_tile_flags(const int64_t* __restrict__ e_ptr, int64_t ne, int32_t* __restrict__ flag, int* __restrict__ too_wide) {
  const int64_t pt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (pt >= ne) return;
  if (e_ptr[pt + 1] - e_ptr[pt] > FA_KMAX) atomicOr(too_wide, 1);
  int f = 0;
  if (pt > 0) f = (e_ptr[pt] / FA_TOBS != e_ptr[pt - 1] / FA_TOBS) || (pt / FA_TPTS != (pt - 1) / FA_TPTS);
  flag[pt] = f;
}

// ordered observation pairs of a point, keyed by (tile, destination block)
__global__ void k_fa_pair_fill(const int64_t* __restrict__ e_ptr, const int32_t* __restrict__ ob_f, int64_t ne, int64_t nf,
                               const int64_t* __restrict__ off, const int32_t* __restrict__ tile_of_pt, const int64_t* __restrict__ tile_pt_ptr,
                               const uint64_t* __restrict__ dest_keys, int ndest, uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int64_t o = off[e];
  const int tile = tile_of_pt[e];
  const int64_t ob0 = e_ptr[tile_pt_ptr[tile]];
  for (int64_t i = e_ptr[e]; i < e_ptr[e + 1]; ++i)
    for (int64_t j = e_ptr[e]; j < e_ptr[e + 1]; ++j)
      if (ob_f[i] <= ob_f[j]) {
        const int64_t d = lower_bound_u64(dest_keys, ndest, (uint64_t)ob_f[i] * nf + ob_f[j]);
        keys[o] = (uint64_t)tile * (uint64_t)ndest + (uint64_t)d;
        vals[o] = (int32_t)(i - ob0) | ((int32_t)(j - ob0) << 16);
        ++o;
      }
}
__global__ void k_fa_cam_fill(const int32_t* __restrict__ ob_e, const int32_t* __restrict__ ob_f, int64_t nb, int64_t nf,
                              const int64_t* __restrict__ e_ptr, const int32_t* __restrict__ tile_of_pt, const int64_t* __restrict__ tile_pt_ptr,
                              uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (o >= nb) return;
  const int tile = tile_of_pt[ob_e[o]];
  keys[o] = (uint64_t)tile * (uint64_t)nf + (uint64_t)ob_f[o];
  vals[o] = (int32_t)(o - e_ptr[tile_pt_ptr[tile]]);
}
__global__ void k_fa_item_meta(int n_items, const int32_t* __restrict__ seg, const int64_t* __restrict__ begin, int ch,
                               const int64_t* __restrict__ group_ptr, const uint64_t* __restrict__ group_key, uint64_t n_targets,
                               int64_t* __restrict__ item_end, int32_t* __restrict__ item_target, int32_t* __restrict__ item_tile) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const int g = seg[i];
  item_end[i] = min(begin[i] + ch, group_ptr[g + 1]);
  item_target[i] = (int32_t)(group_key[g] % n_targets);
  item_tile[i] = (int32_t)(group_key[g] / n_targets);
}

// keys (tile * n_targets + target) with their entries -> sorted entries, items of <= ch entries, reduction lists
inline int build_items(ItemSet& I, DVec<uint64_t>& keys, DVec<int32_t>& vals, int64_t n, int n_tiles, int64_t n_targets, int ch,
                       cudaStream_t st) {
  I.n_ent = n; I.n_targets = (int)n_targets;
  DVec<uint64_t> ks, gkey;
  DVec<int64_t> gcnt, gptr;
  DVec<int32_t> nruns;
  BA_TRY(ks.alloc(n)); BA_TRY(I.ent.alloc(n));
  if (n > 0)
    BA_TRY(cub_call([&](void* t, size_t& b) {
      return cub::DeviceRadixSort::SortPairs(t, b, keys.p, ks.p, vals.p, I.ent.p, (int)n, 0, bits_for((uint64_t)n_tiles * (uint64_t)n_targets), st);
    }));
  keys.release(); vals.release();
  BA_TRY(gkey.alloc(n)); BA_TRY(gcnt.alloc(n + 1)); BA_TRY(nruns.alloc(1));
  int32_t ng = 0;
  if (n > 0) {
    BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceRunLengthEncode::Encode(t, b, ks.p, gkey.p, gcnt.p, nruns.p, (int)n, st); }));
    BA_CUDA_TRY(cudaMemcpyAsync(&ng, nruns.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    BA_CUDA_TRY(cudaStreamSynchronize(st));
  }
  ks.release();
  BA_TRY(gptr.alloc(ng + 1));
  BA_CUDA_TRY(cudaMemsetAsync(gcnt.p + ng, 0, sizeof(int64_t), st));
  BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, gcnt.p, gptr.p, ng + 1, st); }));
  Chunks C;
  BA_TRY(build_chunks(C, gptr.p, ng, ch, st));
  I.n_items = C.n;
  BA_TRY(I.item_end.alloc(C.n)); BA_TRY(I.item_target.alloc(C.n));
  DVec<int32_t> item_tile, iota;
  BA_TRY(item_tile.alloc(C.n)); BA_TRY(iota.alloc(C.n));
  k_fa_item_meta<<<grid_for(C.n, 256), 256, 0, st>>>(C.n, C.seg.p, C.begin.p, ch, gptr.p, gkey.p, (uint64_t)n_targets, I.item_end.p,
                                                     I.item_target.p, item_tile.p);
  I.item_begin.swap(C.begin);
  BA_TRY(I.tile_item_ptr.alloc((size_t)n_tiles + 1));
  k_seg_ptr<int32_t><<<grid_for(C.n > n_tiles + 1 ? C.n : n_tiles + 1, 256), 256, 0, st>>>(item_tile.p, C.n, n_tiles, I.tile_item_ptr.p);
  k_iota<<<grid_for(C.n, 256), 256, 0, st>>>(iota.p, C.n, 0);
  BA_TRY(sort_to_csr(I.item_target.p, iota.p, C.n, n_targets, I.tgt_ptr, I.red_items, st));
  BA_TRY(build_chunks(I.red_ch, I.tgt_ptr.p, (int)n_targets, FA_CH_RED, st));
  BA_CUDA_TRY(cudaStreamSynchronize(st));
  BA_CUDA_TRY(cudaGetLastError());
  return BA_OK;
}

// Returns BA_ERR_UNSUPPORTED when the problem does not fit the fused path (a point with more than FA_KMAX
// observations, or nothing to do): the caller then keeps the generic pipeline.
inline int build_fused_a(FusedA& F, const Structure& S, cudaStream_t st) {
  F.ready = false;
  const int64_t ne = S.ne, nb = S.nb, nf = S.nf;
  if (ne == 0 || nb == 0 || S.nslots != 1) return BA_ERR_UNSUPPORTED;
  DVec<int32_t> flag, tile_of_pt;
  DVec<int> wide;
  BA_TRY(flag.alloc(ne)); BA_TRY(tile_of_pt.alloc(ne)); BA_TRY(wide.alloc_zero(1, st));
  k_fa_tile_flags<<<grid_for(ne, 256), 256, 0, st>>>(S.e_ptr.p, ne, flag.p, wide.p);
  BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceScan::InclusiveSum(t, b, flag.p, tile_of_pt.p, (int)ne, st); }));
  int h_wide = 0, last_tile = 0;
  BA_CUDA_TRY(cudaMemcpyAsync(&h_wide, wide.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  BA_CUDA_TRY(cudaMemcpyAsync(&last_tile, tile_of_pt.p + (ne - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  BA_CUDA_TRY(cudaStreamSynchronize(st));
  if (h_wide) return BA_ERR_UNSUPPORTED;
  F.n_tiles = last_tile + 1;
  BA_TRY(F.tile_pt_ptr.alloc((size_t)F.n_tiles + 1));
  k_seg_ptr<int32_t><<<grid_for(ne > F.n_tiles + 1 ? ne : F.n_tiles + 1, 256), 256, 0, st>>>(tile_of_pt.p, ne, F.n_tiles, F.tile_pt_ptr.p);
  // pair items
  {
    DVec<int64_t> cnt, off;
    BA_TRY(cnt.alloc(ne + 1)); BA_TRY(off.alloc(ne + 1));
    k_pair_count<<<grid_for(ne + 1, 128), 128, 0, st>>>(S.e_ptr.p, S.ob_f0.p, ne, cnt.p);
    BA_TRY(cub_call([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt.p, off.p, (int)(ne + 1), st); }));
    int64_t np = 0;
    BA_CUDA_TRY(cudaMemcpyAsync(&np, off.p + ne, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    BA_CUDA_TRY(cudaStreamSynchronize(st));
    if (np >= (int64_t)INT32_MAX) return BA_ERR_UNSUPPORTED;
    DVec<uint64_t> keys;
    DVec<int32_t> vals;
    BA_TRY(keys.alloc(np)); BA_TRY(vals.alloc(np));
    k_fa_pair_fill<<<grid_for(ne, 128), 128, 0, st>>>(S.e_ptr.p, S.ob_f0.p, ne, nf, off.p, tile_of_pt.p, F.tile_pt_ptr.p, S.dest_keys.p,
                                                      S.ndest, keys.p, vals.p);
    BA_TRY(build_items(F.pairs, keys, vals, np, F.n_tiles, S.ndest, FA_CH_PAIR, st));
  }
  // camera items
  {
    DVec<uint64_t> keys;
    DVec<int32_t> vals;
    BA_TRY(keys.alloc(nb)); BA_TRY(vals.alloc(nb));
    k_fa_cam_fill<<<grid_for(nb, 256), 256, 0, st>>>(S.ob_e.p, S.ob_f0.p, nb, nf, S.e_ptr.p, tile_of_pt.p, F.tile_pt_ptr.p, keys.p, vals.p);
    BA_TRY(build_items(F.cams, keys, vals, nb, F.n_tiles, nf, FA_CH_CAM, st));
  }
  BA_TRY(F.partP.alloc((size_t)F.pairs.n_items * 36)); BA_TRY(F.partC.alloc((size_t)F.cams.n_items * FA_NVC));
  BA_TRY(F.red1P.alloc((size_t)F.pairs.red_ch.n * 36)); BA_TRY(F.red1C.alloc((size_t)F.cams.red_ch.n * FA_NVC));
  BA_TRY(F.camacc.alloc((size_t)nf * FA_NVC)); BA_TRY(F.Lz.alloc((size_t)ne * 9));
  BA_CUDA_TRY(cudaStreamSynchronize(st));
  BA_CUDA_TRY(cudaGetLastError());
  F.ready = true;
  return BA_OK;
}

// ---------------------------------------------------------------------------------------
// residual + Jacobian of one Model A observation (the arithmetic of k_jac_a)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void fa_linearize(const double* __restrict__ T, const double* X, const double* s, double2 ob, double* r,
                                             double* je, double* jf) {
  double q[3];
  mat3_vec(T, X, q);
  const double p0 = q[0] + T[18], p1 = q[1] + T[19], p2 = q[2] + T[20];
  r[0] = T[21] * p0 / p2 + T[23] - ob.x;
  r[1] = T[22] * p1 / p2 + T[24] - ob.y;
  const double iz = 1.0 / p2;
  const double a = T[21] * iz, bb = -T[21] * p0 * iz * iz, cc = T[22] * iz, dd = -T[22] * p1 * iz * iz;
  double D[9];
  rot_deriv(T + 9, T[25] != 0.0 ? X : q, D);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    jf[k] = (a * D[k] + bb * D[6 + k]) * T[26 + k];
    jf[6 + k] = (cc * D[3 + k] + dd * D[6 + k]) * T[26 + k];
  }
  jf[3] = a * T[29]; jf[4] = 0.0;        jf[5] = bb * T[31];
  jf[9] = 0.0;       jf[10] = cc * T[30]; jf[11] = dd * T[31];
  je[0] = (a * T[0] + bb * T[6]) * s[0]; je[1] = (a * T[1] + bb * T[7]) * s[1]; je[2] = (a * T[2] + bb * T[8]) * s[2];
  je[3] = (cc * T[3] + dd * T[6]) * s[0]; je[4] = (cc * T[4] + dd * T[7]) * s[1]; je[5] = (cc * T[5] + dd * T[8]) * s[2];
}

struct FaParams {
  // structure
  const int64_t* tile_pt_ptr; const int64_t* e_ptr; const int32_t* ob_e; const int32_t* ob_f; const double2* uv;
  const int64_t* tile_pitem_ptr; const int64_t* pitem_begin; const int64_t* pitem_end; const int32_t* pent;
  const int64_t* tile_citem_ptr; const int64_t* citem_begin; const int64_t* citem_end; const int32_t* cent;
  // state
  const double* xe; const double* se; const double* tab_f; const double* radius;
  double min_diag, max_diag;
  // pass 1 outputs
  double* partP; double* partC; double* Lz; double* se_out;
  double* cost_partial; double* gmax_partial; double* g2_partial;
  // pass 2 inputs / outputs
  const double* yf; const double* tabc_f; double* xe_c;
  double* mcc_partial; double* x2_partial; double* d2_partial; double* cand_partial;
  int* status;
};

// NORMS = true: iteration 0 only, unscaled Jacobian; writes the Jacobi scaling of the points and the camera items
// (their F^T F diagonals are the camera column norms); no Schur products.
template <bool NORMS>
__global__ void __launch_bounds__(FA_THREADS, 2) k_fa_pass1(FaParams P) {
  extern __shared__ double rec[];  // [FA_CAP][FA_REC]: U (6) | Jf (12) | r (2) | w (2)
  __shared__ double red[32];
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int64_t pt0 = P.tile_pt_ptr[tile], pt1 = P.tile_pt_ptr[tile + 1];
  const int64_t ob0 = P.e_ptr[pt0], ob1 = P.e_ptr[pt1];
  const int nobs = (int)(ob1 - ob0), npts = (int)(pt1 - pt0);
  // ---- A1: one thread per observation ----
  double sq = 0.0;
  for (int l = tid; l < nobs; l += FA_THREADS) {
    const int64_t o = ob0 + l;
    const int64_t e = P.ob_e[o];
    double T[TAB];
    load_tab(P.tab_f, P.ob_f[o], T);
    const double X[3] = {P.xe[3 * e], P.xe[3 * e + 1], P.xe[3 * e + 2]};
    double s[3] = {1.0, 1.0, 1.0};
    if (!NORMS) { s[0] = P.se[3 * e]; s[1] = P.se[3 * e + 1]; s[2] = P.se[3 * e + 2]; }
    double r[2], je[6], jf[12];
    fa_linearize(T, X, s, P.uv[o], r, je, jf);
    sq += r[0] * r[0] + r[1] * r[1];
    double2* R2 = reinterpret_cast<double2*>(rec + (size_t)l * FA_REC);
#pragma unroll
    for (int k = 0; k < 3; ++k) R2[k] = make_double2(je[2 * k], je[2 * k + 1]);
#pragma unroll
    for (int k = 0; k < 6; ++k) R2[3 + k] = make_double2(jf[2 * k], jf[2 * k + 1]);
    R2[9] = make_double2(r[0], r[1]);
    R2[10] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  // ---- A2: one thread per point ----
  double gmx = 0.0, g2 = 0.0;
  const double radius = *P.radius;
  for (int lp = tid; lp < npts; lp += FA_THREADS) {
    const int64_t e = pt0 + lp;
    const int l0 = (int)(P.e_ptr[e] - ob0), l1 = (int)(P.e_ptr[e + 1] - ob0);
    double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    for (int l = l0; l < l1; ++l) {
      const double* R = rec + (size_t)l * FA_REC;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const double j0 = R[3 * rr], j1 = R[3 * rr + 1], j2 = R[3 * rr + 2], rv = R[18 + rr];
        M[0] += j0 * j0; M[1] += j0 * j1; M[2] += j0 * j2; M[4] += j1 * j1; M[5] += j1 * j2; M[8] += j2 * j2;
        g[0] += j0 * rv; g[1] += j1 * rv; g[2] += j2 * rv;
      }
    }
    if (NORMS) {
      P.se_out[3 * e] = 1.0 / (1.0 + sqrt(M[0])); P.se_out[3 * e + 1] = 1.0 / (1.0 + sqrt(M[4])); P.se_out[3 * e + 2] = 1.0 / (1.0 + sqrt(M[8]));
      continue;
    }
    if (l1 > l0) {  // |x - Plus(x, -g)| of the unscaled gradient (TrustRegionMinimizer::EvaluateGradientAndJacobian)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double xv = P.xe[3 * e + k];
        const double d = xv - (xv + (-(g[k] / P.se[3 * e + k])));
        gmx = fmax(gmx, fabs(d)); g2 += d * d;
      }
    }
    M[3] = M[1]; M[6] = M[2]; M[7] = M[5];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double d = sqrt(fmin(fmax(M[4 * k], P.min_diag), P.max_diag) / radius);
      M[4 * k] += d * d;
    }
    if (!chol_small<3>(M)) {
      atomicOr(P.status, 1);
#pragma unroll
      for (int k = 0; k < 9; ++k) M[k] = (k % 4 == 0) ? 1.0 : 0.0;
    }
    fwd_small<3>(M, g);  // z
    double* out = P.Lz + 9 * e;
    out[0] = M[0]; out[1] = M[3]; out[2] = M[4]; out[3] = M[6]; out[4] = M[7]; out[5] = M[8]; out[6] = g[0]; out[7] = g[1]; out[8] = g[2];
    for (int l = l0; l < l1; ++l) {
      double* R = rec + (size_t)l * FA_REC;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        double u[3] = {R[3 * rr], R[3 * rr + 1], R[3 * rr + 2]};
        fwd_small<3>(M, u);  // u_rr = L^-1 (J_e row rr)^T
        R[3 * rr] = u[0]; R[3 * rr + 1] = u[1]; R[3 * rr + 2] = u[2];
        R[20 + rr] = u[0] * g[0] + u[1] * g[1] + u[2] * g[2];  // w_rr = u_rr . z
      }
    }
  }
  __syncthreads();
  // ---- B: one thread per work item ----
  for (int64_t it = P.tile_citem_ptr[tile] + tid; it < P.tile_citem_ptr[tile + 1]; it += FA_THREADS) {
    double acc[FA_NVC];
#pragma unroll
    for (int k = 0; k < FA_NVC; ++k) acc[k] = 0.0;
    for (int64_t q = P.citem_begin[it]; q < P.citem_end[it]; ++q) {
      const double2* R2 = reinterpret_cast<const double2*>(rec + (size_t)P.cent[q] * FA_REC);
      double jf[12];
#pragma unroll
      for (int k = 0; k < 6; ++k) { const double2 v = R2[3 + k]; jf[2 * k] = v.x; jf[2 * k + 1] = v.y; }
      const double2 rv = R2[9], wv = R2[10];
      int c = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) acc[c++] += jf[a] * jf[b] + jf[6 + a] * jf[6 + b];
#pragma unroll
      for (int a = 0; a < 6; ++a) { acc[21 + a] += jf[a] * rv.x + jf[6 + a] * rv.y; acc[27 + a] += jf[a] * wv.x + jf[6 + a] * wv.y; }
    }
    double* out = P.partC + (size_t)it * FA_NVC;
#pragma unroll
    for (int k = 0; k < FA_NVC; ++k) out[k] = acc[k];
  }
  if (!NORMS) {
    for (int64_t it = P.tile_pitem_ptr[tile] + tid; it < P.tile_pitem_ptr[tile + 1]; it += FA_THREADS) {
      double acc[36];
#pragma unroll
      for (int k = 0; k < 36; ++k) acc[k] = 0.0;
      for (int64_t q = P.pitem_begin[it]; q < P.pitem_end[it]; ++q) {
        const int32_t ent = P.pent[q];
        const double2* Ri = reinterpret_cast<const double2*>(rec + (size_t)(ent & 0xffff) * FA_REC);
        const double2* Rj = reinterpret_cast<const double2*>(rec + (size_t)((ent >> 16) & 0xffff) * FA_REC);
        double ui[6], uj[6], fj[12];
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double2 a = Ri[k], b = Rj[k]; ui[2 * k] = a.x; ui[2 * k + 1] = a.y; uj[2 * k] = b.x; uj[2 * k + 1] = b.y; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { const double2 b = Rj[3 + k]; fj[2 * k] = b.x; fj[2 * k + 1] = b.y; }
        const double g00 = ui[0] * uj[0] + ui[1] * uj[1] + ui[2] * uj[2], g01 = ui[0] * uj[3] + ui[1] * uj[4] + ui[2] * uj[5];
        const double g10 = ui[3] * uj[0] + ui[4] * uj[1] + ui[5] * uj[2], g11 = ui[3] * uj[3] + ui[4] * uj[4] + ui[5] * uj[5];
        double t0[6], t1[6];
#pragma unroll
        for (int b = 0; b < 6; ++b) { t0[b] = g00 * fj[b] + g01 * fj[6 + b]; t1[b] = g10 * fj[b] + g11 * fj[6 + b]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double2 f0 = Ri[3 + k], f1 = Ri[6 + k];   // J_f,i rows 0 / 1, columns 2k, 2k+1
#pragma unroll
          for (int b = 0; b < 6; ++b) {
            acc[(2 * k) * 6 + b] += f0.x * t0[b] + f1.x * t1[b];
            acc[(2 * k + 1) * 6 + b] += f0.y * t0[b] + f1.y * t1[b];
          }
        }
      }
      double2* out = reinterpret_cast<double2*>(P.partP + (size_t)it * 36);
#pragma unroll
      for (int k = 0; k < 18; ++k) out[k] = make_double2(acc[2 * k], acc[2 * k + 1]);
    }
    // cost and gradient-norm partials of the tile (fixed tree)
    sq = block_sum(sq, red);
    g2 = block_sum(g2, red);
    gmx = block_max(gmx, red);
    if (tid == 0) { P.cost_partial[tile] = sq; P.g2_partial[tile] = g2; P.gmax_partial[tile] = gmx; }
  }
}

// back-substitution, model cost change, candidate point and candidate cost of one tile
__global__ void __launch_bounds__(FA_THREADS, 2) k_fa_pass2(FaParams P) {
  extern __shared__ double rec[];  // [FA_CAP][FA_REC2]: J_e (6) | r (2) | J_f yf (2) ; then [FA_TPTS][3] candidate points
  __shared__ double red[32];
  double* Xc = rec + (size_t)FA_CAP * FA_REC2;
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int64_t pt0 = P.tile_pt_ptr[tile], pt1 = P.tile_pt_ptr[tile + 1];
  const int64_t ob0 = P.e_ptr[pt0], ob1 = P.e_ptr[pt1];
  const int nobs = (int)(ob1 - ob0), npts = (int)(pt1 - pt0);
  for (int l = tid; l < nobs; l += FA_THREADS) {
    const int64_t o = ob0 + l;
    const int64_t e = P.ob_e[o];
    const int32_t c = P.ob_f[o];
    double T[TAB];
    load_tab(P.tab_f, c, T);
    const double X[3] = {P.xe[3 * e], P.xe[3 * e + 1], P.xe[3 * e + 2]};
    const double s[3] = {P.se[3 * e], P.se[3 * e + 1], P.se[3 * e + 2]};
    double r[2], je[6], jf[12];
    fa_linearize(T, X, s, P.uv[o], r, je, jf);
    double y[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) y[k] = __ldg(P.yf + 6 * (int64_t)c + k);
    double q0 = 0.0, q1 = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { q0 += jf[k] * y[k]; q1 += jf[6 + k] * y[k]; }
    double2* R2 = reinterpret_cast<double2*>(rec + (size_t)l * FA_REC2);
#pragma unroll
    for (int k = 0; k < 3; ++k) R2[k] = make_double2(je[2 * k], je[2 * k + 1]);
    R2[3] = make_double2(r[0], r[1]);
    R2[4] = make_double2(q0, q1);
  }
  __syncthreads();
  double mcc = 0.0, x2 = 0.0, d2 = 0.0;
  for (int lp = tid; lp < npts; lp += FA_THREADS) {
    const int64_t e = pt0 + lp;
    const int l0 = (int)(P.e_ptr[e] - ob0), l1 = (int)(P.e_ptr[e + 1] - ob0);
    const double* in = P.Lz + 9 * e;
    const double L[9] = {in[0], 0.0, 0.0, in[1], in[2], 0.0, in[3], in[4], in[5]};
    double t[3] = {in[6], in[7], in[8]};
    for (int l = l0; l < l1; ++l) {
      const double* R = rec + (size_t)l * FA_REC2;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        double u[3] = {R[3 * rr], R[3 * rr + 1], R[3 * rr + 2]};
        fwd_small<3>(L, u);
        const double q = R[8 + rr];
        t[0] -= u[0] * q; t[1] -= u[1] * q; t[2] -= u[2] * q;
      }
    }
    bwd_small<3>(L, t);  // y_e
    // Ceres: model_cost_change = -(J step)^T (r + J step / 2) with step = -y; the caller negates the sum
    for (int l = l0; l < l1; ++l) {
      const double* R = rec + (size_t)l * FA_REC2;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const double m = -(R[3 * rr] * t[0] + R[3 * rr + 1] * t[1] + R[3 * rr + 2] * t[2]) - R[8 + rr];
        mcc += m * (R[6 + rr] + m / 2.0);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double xv = P.xe[3 * e + k];
      const double c = xv + (-(t[k]) * P.se[3 * e + k]);
      P.xe_c[3 * e + k] = c;
      Xc[3 * lp + k] = c;
      if (l1 > l0) { x2 += xv * xv; const double d = xv - c; d2 += d * d; }
    }
  }
  __syncthreads();
  double sq = 0.0;
  for (int l = tid; l < nobs; l += FA_THREADS) {
    const int64_t o = ob0 + l;
    const int lp = (int)(P.ob_e[o] - pt0);
    const double* T = P.tabc_f + TAB * (int64_t)P.ob_f[o];
    double Rm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rm[k] = __ldg(T + k);
    const double X[3] = {Xc[3 * lp], Xc[3 * lp + 1], Xc[3 * lp + 2]};
    double q[3];
    mat3_vec(Rm, X, q);
    const double p0 = q[0] + __ldg(T + 18), p1 = q[1] + __ldg(T + 19), p2 = q[2] + __ldg(T + 20);
    const double2 ob = P.uv[o];
    const double r0 = __ldg(T + 21) * p0 / p2 + __ldg(T + 23) - ob.x;
    const double r1 = __ldg(T + 22) * p1 / p2 + __ldg(T + 24) - ob.y;
    sq += r0 * r0 + r1 * r1;
  }
  mcc = block_sum(mcc, red);
  x2 = block_sum(x2, red);
  d2 = block_sum(d2, red);
  sq = block_sum(sq, red);
  if (tid == 0) { P.mcc_partial[tile] = mcc; P.x2_partial[tile] = x2; P.d2_partial[tile] = d2; P.cand_partial[tile] = sq; }
}

// first level: one warp per chunk of <= ch partial blocks of one target, lane = value, sequential over the blocks
template <int NV>
__global__ void __launch_bounds__(128)
k_reduce_items(int nchunks, int ch, const int32_t* __restrict__ chunk_seg, const int64_t* __restrict__ chunk_begin,
               const int64_t* __restrict__ tgt_ptr, const int32_t* __restrict__ items, const double* __restrict__ part, double* __restrict__ out) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= nchunks) return;
  const int64_t begin = chunk_begin[c];
  const int64_t end = min(begin + ch, tgt_ptr[chunk_seg[c] + 1]);
  for (int v = lane; v < NV; v += 32) {
    double a = 0.0;
    for (int64_t idx = begin; idx < end; ++idx) a += part[(int64_t)items[idx] * NV + v];
    out[(int64_t)c * NV + v] = a;
  }
}
// second level: one warp per target over its chunk partials
template <int NV>
__global__ void __launch_bounds__(128)
k_reduce_final(int ntargets, const int32_t* __restrict__ seg_first, const double* __restrict__ part1, double* __restrict__ out) {
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= ntargets) return;
  const int c0 = seg_first[t], c1 = seg_first[t + 1];
  for (int v = lane; v < NV; v += 32) {
    double a = 0.0;
    for (int c = c0; c < c1; ++c) a += part1[(int64_t)c * NV + v];
    out[(int64_t)t * NV + v] = a;
  }
}

// folds up to 4 partial arrays in one launch: blockIdx.x = which; op 0 = sum, 1 = max
struct FoldJob { const double* partial[4]; int slot[4]; int is_max[4]; };
__global__ void k_fold_multi(FoldJob J, int n, double* out) {
  __shared__ double sm[32];
  const double* partial = J.partial[blockIdx.x];
  double v = 0.0;
  if (J.is_max[blockIdx.x]) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) v = fmax(v, partial[i]);
    v = block_max(v, sm);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partial[i];
    v = block_sum(v, sm);
  }
  if (threadIdx.x == 0) out[J.slot[blockIdx.x]] = v;
}

}  // namespace ba
