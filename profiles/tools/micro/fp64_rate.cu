// fp64_rate.cu -- what one B200 SM sustains in FP64: latency of a dependent DFMA chain, and the issue rate of K
// independent chains per warp with W warps per SM.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double x[K];
#pragma unroll
  for (int i = 0; i < K; ++i) x[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < K; ++i) x[i] = fma(x[i], a, b);
  }
  const long long t1 = clock64();
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int K>
void run(int threads, double* out, long long* cyc) {
  const int iters = 4096;
  k<K><<<1, threads>>>(out, cyc, iters, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_warp_inst = (double)h / ((double)iters * K);
  const int warps = threads / 32;
  printf("chains/warp %2d  warps/SM %2d : %.2f cycles per DFMA per warp, %.3f warp-DFMA/cycle/SM (%.1f lanes/cycle/SM)\n", K, warps, per_warp_inst,
         warps / per_warp_inst, 32.0 * warps / per_warp_inst);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8 * 8); cudaMalloc(&cyc, 64);
  for (int threads : {32, 128, 256, 512, 1024}) {
    run<1>(threads, out, cyc); run<2>(threads, out, cyc); run<4>(threads, out, cyc); run<8>(threads, out, cyc); run<16>(threads, out, cyc);
  }
  return 0;
}
