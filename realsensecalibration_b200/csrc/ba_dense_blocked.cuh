// ba_dense_blocked.cuh -- K3a for mid-sized reduced camera systems (160 < n, e.g. cfg3-B: 7 cameras + 99 markers,
// n = 636; cfg4 with DENSE_SCHUR forced: n = 6000): right-looking blocked Cholesky over the whole GPU in ONE
// cooperative launch, the trailing update on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64), then the two
// triangular solves by CTA 0 out of L2.
//
// Replaces Ceres' DenseSchurComplementSolver -> Eigen LLT (options.linear_solver_type = DENSE_SCHUR,
// bundle_adjustment_manager.cpp:91) where the single-CTA kernel of ba_dense.cuh no longer holds the matrix in
// shared memory.  Every sum runs in a fixed order (fixed tile -> warp map, fixed k order): bitwise reproducible.
//
//   step k (panel of NB = 32 columns), two grid barriers:
//     1  every CTA factors the 32x32 diagonal block in its own shared memory (redundant, 11 kflop; saves a barrier),
//        then solves its share of the row blocks below:  L_ik = A_ik L_kk^-T   (one warp per 32x32 block, one lane per row)
//     2  trailing update A_IJ -= L_Ik L_Jk^T on 64x64 output tiles, one CTA per tile: both panels staged in shared
//        memory (row stride 36 doubles: the DMMA fragment loads are conflict free), 8 warps x 8 m8n8 accumulators
//   solve: L t = b by row blocks (coalesced row dot products), L^T y = t by column updates (coalesced row reads)
#pragma once
#include <cooperative_groups.h>

#include "ba_util.cuh"

namespace ba {
namespace cgb = cooperative_groups;

constexpr int CB_NB = 32;       // panel width
constexpr int CB_TB = 64;       // output tile of the trailing update
constexpr int CB_LD = 36;       // shared-memory row stride of a 32-wide panel tile (doubles)
constexpr int CB_THREADS = 256;

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// A: n x n row-major, lower triangle + diagonal valid on entry; L overwrites the lower triangle.
// status |= 2 if a pivot is not positive, |= 4 if the solution is not finite (as k_chol_solve).
__global__ void __launch_bounds__(CB_THREADS)
k_chol_blocked(int n, double* __restrict__ A, const double* __restrict__ rhs, double* __restrict__ y, double* __restrict__ work,
               int* status, int* bad_flag) {
  cgb::grid_group grid = cgb::this_grid();
  __shared__ __align__(16) double Ld[CB_NB * (CB_NB + 1)];        // diagonal block / its factor
  __shared__ __align__(16) double Pa[CB_TB * CB_LD];              // panel rows of tile row I
  __shared__ __align__(16) double Pb[CB_TB * CB_LD];              // panel rows of tile row J
  __shared__ double Linv[CB_NB], colj[CB_NB];                     // 1 / diag(L_kk); the column being eliminated
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (n + CB_NB - 1) / CB_NB;
  const size_t ld = (size_t)n;

  for (int k = 0; k < nblk; ++k) {
    const int k0 = k * CB_NB;
    const int kb = min(CB_NB, n - k0);
    // ---- phase 1: diagonal block (every CTA, in shared memory), then the row blocks below ----
    if (tid == 0) s_bad = 0;
    for (int i = tid; i < CB_NB * CB_NB; i += CB_THREADS) {
      const int r = i / CB_NB, c = i % CB_NB;
      double v = (r == c) ? 1.0 : 0.0;                             // identity padding past n
      if (r < kb && c < kb && c <= r) v = A[(k0 + r) * ld + k0 + c];
      Ld[r * (CB_NB + 1) + c] = v;
    }
    __syncthreads();
    {  // 32 x 32 Cholesky by the whole CTA (a lone warp ran it in 16 us: one warp advances ~6 cycles per instruction).  Right-looking
       // on the unscaled columns (a_rc -= a_rj a_cj / d_j: one barrier per column, no scaling pass in between), the reciprocal of
       // the next pivot left in shared memory by the thread that finished it; scaled to L = L' D^-1/2 at the end.
      const int r = tid >> 3, g = tid & 7;   // thread = (row, column class): columns g, g + 8, g + 16, g + 24
      if (tid == 0) { const double d0 = Ld[0]; colj[0] = (d0 > 0.0 && isfinite(d0)) ? 1.0 / d0 : 0.0; }
      __syncthreads();
      bool ok = true;
      for (int j = 0; j < CB_NB; ++j) {
        const double inv = colj[j];   // 1 / d_j, 0 = not positive (every thread reads the same value: uniform)
        if (inv == 0.0) { ok = false; break; }
        if (r > j) {
          const double lrj = Ld[r * (CB_NB + 1) + j] * inv;
#pragma unroll
          for (int q = 0; q < CB_NB / 8; ++q) {
            const int c = g + 8 * q;
            if (c > j && c <= r) {
              const double v = Ld[r * (CB_NB + 1) + c] - lrj * Ld[c * (CB_NB + 1) + j];
              Ld[r * (CB_NB + 1) + c] = v;
              if (r == j + 1 && c == j + 1) colj[j + 1] = (v > 0.0 && isfinite(v)) ? 1.0 / v : 0.0;
            }
          }
        }
        __syncthreads();
      }
      if (!ok) {
        if (tid == 0) s_bad = 1;
      } else {
        if (tid < CB_NB) Linv[tid] = sqrt(colj[tid]);   // 1 / L_jj
        __syncthreads();
#pragma unroll
        for (int q = 0; q < CB_NB / 8; ++q) {
          const int c = g + 8 * q;
          if (c < r) Ld[r * (CB_NB + 1) + c] *= Linv[c];
          else if (c == r) Ld[r * (CB_NB + 1) + c] = sqrt(Ld[r * (CB_NB + 1) + c]);
        }
      }
    }
    __syncthreads();
    if (s_bad) {
      if (blockIdx.x == 0 && tid == 0) { atomicOr(status, 2); *bad_flag = 1; }
    } else {
      // row blocks below the diagonal: warp-per-block, lane-per-row:  x L_kk^T = a
      const int gw = blockIdx.x * (CB_THREADS / 32) + warp, nw = gridDim.x * (CB_THREADS / 32);
      for (int ib = k + 1 + gw; ib < nblk; ib += nw) {
        const int r = ib * CB_NB + lane;
        if (r < n) {
          double* row = A + r * ld + k0;
          double x[CB_NB];
#pragma unroll
          for (int c = 0; c < CB_NB; ++c) x[c] = c < kb ? row[c] : 0.0;
#pragma unroll
          for (int c = 0; c < CB_NB; ++c) {                        // column form: the updates of one column are independent
            x[c] *= Linv[c];
#pragma unroll
            for (int c2 = c + 1; c2 < CB_NB; ++c2) x[c2] -= x[c] * Ld[c2 * (CB_NB + 1) + c];
          }
#pragma unroll
          for (int c = 0; c < CB_NB; ++c)
            if (c < kb) row[c] = x[c];
        }
      }
    }
    __threadfence();
    grid.sync();
    if (*reinterpret_cast<volatile int*>(bad_flag)) break;         // uniform over the grid after the barrier
    if (blockIdx.x == 0)                                           // CTA 0 publishes L_kk (only now: the others read A_kk above)
      for (int i = tid; i < kb * kb; i += CB_THREADS) {
        const int r = i / kb, c = i % kb;
        if (c <= r) A[(k0 + r) * ld + k0 + c] = Ld[r * (CB_NB + 1) + c];
      }
    // ---- phase 2: trailing update on 64x64 tiles, DMMA ----
    const int t0 = k0 + CB_NB;                                     // first trailing row / column
    const int m = n - t0;
    if (m > 0) {
      const int mt = (m + CB_TB - 1) / CB_TB;
      const int ntiles = mt * (mt + 1) / 2;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        // t -> (I, J), I >= J, row-major over the lower triangle
        int I = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((I + 1) * (I + 2) / 2 <= t) ++I;
        while (I * (I + 1) / 2 > t) --I;
        const int J = t - I * (I + 1) / 2;
        const int ri = t0 + I * CB_TB, rj = t0 + J * CB_TB;
        __syncthreads();
        for (int i = tid; i < CB_TB * CB_NB; i += CB_THREADS) {
          const int r = i / CB_NB, c = i % CB_NB;
          Pa[r * CB_LD + c] = (ri + r < n && c < kb) ? A[(ri + r) * ld + k0 + c] : 0.0;
          Pb[r * CB_LD + c] = (rj + r < n && c < kb) ? A[(rj + r) * ld + k0 + c] : 0.0;
        }
        __syncthreads();
        double acc[8][2];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c][0] = acc[c][1] = 0.0;
        const int fr = lane >> 2, fc = lane & 3;                   // fragment row / k index of this lane
#pragma unroll
        for (int kk = 0; kk < CB_NB / 4; ++kk) {
          const double a = Pa[(8 * warp + fr) * CB_LD + 4 * kk + fc];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const double b = Pb[(8 * c + fr) * CB_LD + 4 * kk + fc];
            dmma_m8n8k4(acc[c][0], acc[c][1], a, b);
          }
        }
        const int row = ri + 8 * warp + fr;
        if (row < n) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int col = rj + 8 * c + 2 * fc;
            if (col + 1 < n && col + 1 <= row) {
              double2* p = reinterpret_cast<double2*>(A + row * ld + col);
              if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                double2 v = *p; v.x -= acc[c][0]; v.y -= acc[c][1]; *p = v;
              } else {
                A[row * ld + col] -= acc[c][0]; A[row * ld + col + 1] -= acc[c][1];
              }
            } else {
              if (col < n && col <= row) A[row * ld + col] -= acc[c][0];
              if (col + 1 < n && col + 1 <= row) A[row * ld + col + 1] -= acc[c][1];
            }
          }
        }
      }
    }
    __threadfence();
    grid.sync();
  }
  // ---- solves, CTA 0 ----
  if (blockIdx.x != 0) return;
  if (*reinterpret_cast<volatile int*>(bad_flag)) {
    for (int i = tid; i < n; i += CB_THREADS) y[i] = 0.0;
    return;
  }
  double* bv = work;                                               // n doubles in global memory (L2 resident)
  for (int i = tid; i < n; i += CB_THREADS) bv[i] = rhs[i];
  __syncthreads();
  constexpr int NW = CB_THREADS / 32;
  // forward: L t = b, row blocks of 32
  for (int k = 0; k < nblk; ++k) {
    const int k0 = k * CB_NB, kb = min(CB_NB, n - k0);
    // rows of the block minus the contribution of everything solved so far: warp per row, lanes over columns
    for (int r = warp; r < kb; r += NW) {
      const double* row = A + (size_t)(k0 + r) * ld;
      double s = 0.0;
      for (int c = lane; c < k0; c += 32) s += row[c] * bv[c];
      s = warp_sum(s);
      if (lane == 0) Ld[r] = bv[k0 + r] - s;
    }
    for (int i = tid; i < kb * kb; i += CB_THREADS) Pa[(i / kb) * CB_LD + i % kb] = A[(size_t)(k0 + i / kb) * ld + k0 + i % kb];  // L_kk
    __syncthreads();
    if (warp == 0) {                                               // 32x32 triangular solve out of shared memory, lane = row
      double mine = lane < kb ? Ld[lane] : 0.0;
      for (int j = 0; j < kb; ++j) {
        const double tj = __shfl_sync(0xffffffffu, mine, j) / Pa[j * CB_LD + j];
        if (lane == j) mine = tj;
        if (lane > j && lane < kb) mine -= Pa[lane * CB_LD + j] * tj;
      }
      if (lane < kb) bv[k0 + lane] = mine;
    }
    __syncthreads();
  }
  // backward: L^T y = t, from the last block; after a block is solved its rows update everything above (coalesced)
  for (int k = nblk - 1; k >= 0; --k) {
    const int k0 = k * CB_NB, kb = min(CB_NB, n - k0);
    for (int i = tid; i < kb * kb; i += CB_THREADS) Pa[(i / kb) * CB_LD + i % kb] = A[(size_t)(k0 + i / kb) * ld + k0 + i % kb];  // L_kk
    __syncthreads();
    if (warp == 0) {
      double mine = lane < kb ? bv[k0 + lane] : 0.0;
      for (int j = kb - 1; j >= 0; --j) {
        const double yj = __shfl_sync(0xffffffffu, mine, j) / Pa[j * CB_LD + j];
        if (lane == j) mine = yj;
        if (lane < j) mine -= Pa[j * CB_LD + lane] * yj;
      }
      if (lane < kb) { bv[k0 + lane] = mine; Ld[lane] = mine; }
    }
    __syncthreads();
    for (int c = tid; c < k0; c += CB_THREADS) {
      double s = bv[c];
      for (int j = 0; j < kb; ++j) s -= A[(size_t)(k0 + j) * ld + c] * Ld[j];
      bv[c] = s;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += CB_THREADS) {
    const double v = bv[i];
    y[i] = v;
    if (!isfinite(v)) atomicOr(status, 4);
  }
}

struct CholBlockedWork {
  DVec<double> work;
  DVec<int> bad;
  int grid = 0;
};

inline int launch_chol_blocked(CholBlockedWork& W, int n, double* S, const double* rhs, double* y, int* status, int device, cudaStream_t st) {
  if (W.grid == 0) {
    int per_sm = 0, sms = 0;
    BA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chol_blocked, CB_THREADS, 0));
    BA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (per_sm < 1) return fail(BA_ERR_CUDA, "k_chol_blocked does not fit an SM");
    W.grid = sms * std::min(per_sm, 2);
    BA_TRY(W.bad.alloc(1));
  }
  if (W.work.n < (size_t)n) BA_TRY(W.work.alloc(n));
  BA_CUDA_TRY(cudaMemsetAsync(W.bad.p, 0, sizeof(int), st));
  // no more CTAs than there is work in the first (largest) trailing update
  const int mt = (n + CB_TB - 1) / CB_TB;
  const int grid = std::max(1, std::min(W.grid, mt * (mt + 1) / 2));
  double* work = W.work.p;
  int* bad = W.bad.p;
  void* args[] = {&n, &S, &rhs, &y, &work, &status, &bad};
  BA_CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_chol_blocked, dim3(grid), dim3(CB_THREADS), args, 0, st));
  return BA_OK;
}

}  // namespace ba
