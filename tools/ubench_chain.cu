// Microbenchmark of the moment-chain consumer loops of k_moments_pipe (data already in shared memory):
// cycles per row for the sum chain, the sum-of-squares chain, and variants.  One warp per block.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_chain ubench_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ROWS = 64, PITCH = 34;
__global__ void k_chain(double* out, long long* cyc, int iters, int variant) {
  __shared__ double sz[ROWS * PITCH];
  __shared__ double sw[ROWS];
  const int lane = threadIdx.x;
  for (int i = lane; i < ROWS * PITCH; i += 32) sz[i] = 1.0 + 1e-9 * i;
  for (int i = lane; i < ROWS; i += 32) sw[i] = 0.5 + 1e-3 * i;
  __syncwarp();
  const double* zr = sz + lane;
  double acc = 0.0, acc2 = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (variant == 0) {          // sum chain: DMUL + DADD per row
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc = __dadd_rn(acc, __dmul_rn(zr[r * PITCH], sw[r]));
    } else if (variant == 1) {   // sum-of-squares chain: DMUL, DMUL, DADD
#pragma unroll
      for (int r = 0; r < ROWS; ++r) { const double z = zr[r * PITCH]; acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(z, sw[r]), z)); }
    } else if (variant == 2) {   // both chains in one warp
#pragma unroll
      for (int r = 0; r < ROWS; ++r) { const double z = zr[r * PITCH]; const double wz = __dmul_rn(z, sw[r]); acc = __dadd_rn(acc, wz); acc2 = __dadd_rn(acc2, __dmul_rn(wz, z)); }
    } else if (variant == 3) {   // pure chain: LDS + DADD
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc = __dadd_rn(acc, zr[r * PITCH]);
    } else {                     // products of all rows first (registers), then the chain
      double pr[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) { const double z = zr[r * PITCH]; pr[r] = __dmul_rn(__dmul_rn(z, sw[r]), z); }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc = __dadd_rn(acc, pr[r]);
    }
  }
  long long t1 = clock64();
  if (lane == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 32 + lane] = acc + acc2;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 32 * 8); cudaMalloc(&cyc, 1024 * 8);
  const char* names[5] = {"sum_chain(DMUL,DADD)", "sumsq_chain(DMUL,DMUL,DADD)", "both_chains_one_warp", "pure_chain(LDS,DADD)", "products_first_then_chain"};
  printf("{");
  for (int v = 0; v < 5; ++v) {
    int iters = 2000;
    k_chain<<<1, 32>>>(out, cyc, iters, v); cudaDeviceSynchronize();
    k_chain<<<1, 32>>>(out, cyc, iters, v); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("\"%s_cycles_per_row\": %.2f%s", names[v], (double)h / ((double)iters * ROWS), v < 4 ? ", " : "");
  }
  printf("}\n");
  return 0;
}
