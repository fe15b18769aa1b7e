// FP64 calibration microbenchmarks for B200 (sm_100a).
// Measures the denominators DESIGN.md quotes: DFMA and DMMA.8x8x4 issue rate,
// dependent DADD / DMMA latency, IEEE double-division rate, cuBLAS DGEMM on a
// square and on the fold-Gram shape, and a plain device copy.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_fp64 ubench_fp64.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = threadIdx.x; acc[i][1] = i; }
  double av = a + threadIdx.x * 1e-9, bv = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ddiv(double* out, int iters, double a) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 1.0 + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __ddiv_rn(a, acc[i]) + 1.5;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lat(double* out, long long* cyc, int iters, double a) {
  double x = a;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    x = __dadd_rn(x, a); x = __dadd_rn(x, a); x = __dadd_rn(x, a); x = __dadd_rn(x, a);
    x = __dadd_rn(x, a); x = __dadd_rn(x, a); x = __dadd_rn(x, a); x = __dadd_rn(x, a);
  }
  long long t1 = clock64();
  double d0 = a, d1 = a;
  for (int it = 0; it < iters; ++it) {
    dmma884(d0, d1, a, a); dmma884(d0, d1, a, a); dmma884(d0, d1, a, a); dmma884(d0, d1, a, a);
    dmma884(d0, d1, a, a); dmma884(d0, d1, a, a); dmma884(d0, d1, a, a); dmma884(d0, d1, a, a);
  }
  long long t2 = clock64();
  double f = a;
  for (int it = 0; it < iters; ++it) {
    f = fma(f, a, a); f = fma(f, a, a); f = fma(f, a, a); f = fma(f, a, a);
    f = fma(f, a, a); f = fma(f, a, a); f = fma(f, a, a); f = fma(f, a, a);
  }
  long long t3 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
  out[threadIdx.x] = x + d0 + d1 + f;
}

__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, p.multiProcessorCount, clk_khz);
  const int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 8));

  // DFMA throughput: 8 warps/SMSP-equivalents, 8 independent chains per thread
  for (int bps : {1, 2, 4}) {
    int iters = 20000;
    int blocks = sms * bps, threads = 256;
    float ms = time_ms([&] { k_dfma<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double flops = 2.0 * 16 * iters * (double)blocks * threads;
    printf(" \"dfma_tflops_bps%d\": %.2f,\n", bps, flops / ms / 1e9);
  }
  // DMMA throughput
  for (int bps : {1, 2}) {
    for (int warps : {4, 8, 16}) {
      int iters = 4000;
      int blocks = sms * bps, threads = warps * 32;
      float ms = time_ms([&] { k_dmma<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double flops = 2.0 * 256 * 16 * iters * (double)blocks * warps;
      printf(" \"dmma_tflops_bps%d_w%d_acc16\": %.2f,\n", bps, warps, flops / ms / 1e9);
    }
  }
  {
    int iters = 4000, blocks = sms, threads = 256;
    float ms = time_ms([&] { k_dmma<32><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double flops = 2.0 * 256 * 32 * iters * (double)blocks * 8;
    printf(" \"dmma_tflops_bps1_w8_acc32\": %.2f,\n", flops / ms / 1e9);
    ms = time_ms([&] { k_dmma<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    flops = 2.0 * 256 * 4 * iters * (double)blocks * 8;
    printf(" \"dmma_tflops_bps1_w8_acc4\": %.2f,\n", flops / ms / 1e9);
  }
  // division
  {
    int iters = 2000, blocks = sms * 4, threads = 256;
    float ms = time_ms([&] { k_ddiv<<<blocks, threads>>>(out, iters, 3.0); }, 5);
    double divs = 8.0 * iters * (double)blocks * threads;
    printf(" \"ddiv_gdiv_per_s\": %.2f,\n", divs / ms / 1e6);
  }
  // latencies
  {
    long long* cyc; CK(cudaMalloc(&cyc, 3 * sizeof(long long)));
    int iters = 10000;
    k_lat<<<1, 32>>>(out, cyc, iters, 1.0000001);
    CK(cudaDeviceSynchronize());
    k_lat<<<1, 32>>>(out, cyc, iters, 1.0000001);
    CK(cudaDeviceSynchronize());
    long long h[3]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf(" \"dadd_latency_cyc\": %.2f, \"dmma_latency_cyc\": %.2f, \"dfma_latency_cyc\": %.2f,\n",
           h[0] / (8.0 * iters), h[1] / (8.0 * iters), h[2] / (8.0 * iters));
    float ms = time_ms([&] { k_lat<<<1, 32>>>(out, cyc, iters, 1.0000001); }, 3);
    printf(" \"lat_kernel_ms\": %.3f, \"lat_kernel_total_cyc\": %lld,\n", ms, h[0] + h[1] + h[2]);
  }
  // copy bandwidth
  {
    size_t n = (size_t)1 << 30;  // 1 GiB each way
    double2 *a, *b; CK(cudaMalloc(&a, n)); CK(cudaMalloc(&b, n));
    CK(cudaMemset(a, 1, n));
    float ms = time_ms([&] { k_copy<<<sms * 8, 512>>>(a, b, n / sizeof(double2)); }, 10);
    printf(" \"copy_gbs\": %.1f,\n", 2.0 * n / ms / 1e6);
    ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, n, cudaMemcpyDeviceToDevice)); }, 10);
    printf(" \"memcpy_d2d_gbs\": %.1f,\n", 2.0 * n / ms / 1e6);
    CK(cudaFree(a)); CK(cudaFree(b));
  }
  // cuBLAS DGEMM
  {
    cublasHandle_t h; cublasCreate(&h);
    {
      int n = 8192;
      double *A, *B, *C; CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n));
      CK(cudaMalloc(&C, sizeof(double) * n * n));
      CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
      double one = 1, zero = 0;
      float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
      printf(" \"cublas_dgemm_8192_tflops\": %.2f,\n", 2.0 * n * n * (double)n / ms / 1e9);
      // sustained: 2 s back to back
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      int reps = (int)(2000.0f / ms) + 1;
      cudaEventRecord(e0);
      for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float tot; cudaEventElapsedTime(&tot, e0, e1);
      printf(" \"cublas_dgemm_8192_sustained_tflops\": %.2f,\n", 2.0 * n * n * (double)n * reps / tot / 1e9);
      CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
    }
    {
      // fold-Gram shape: C(510 x 500) = Z^T (510 x Nv) * X (Nv x 500); row-major X (Nv x 512 pitch)
      int K = 500, KM = 510, ld = 512; int Nv = 200000;
      double *Z, *C; CK(cudaMalloc(&Z, sizeof(double) * (size_t)Nv * ld)); CK(cudaMalloc(&C, sizeof(double) * KM * K));
      CK(cudaMemset(Z, 0, sizeof(double) * (size_t)Nv * ld));
      double one = 1, zero = 0;
      // row-major Z (Nv x ld) is column-major (ld x Nv): C_cm(KM x K) = Zcm(KM x Nv) * Zcm(K x Nv)^T
      float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, KM, K, Nv, &one, Z, ld, Z, ld, &zero, C, KM); }, 5);
      printf(" \"cublas_dgemm_gramshape_ms\": %.3f, \"cublas_dgemm_gramshape_tflops\": %.2f,\n", ms,
             2.0 * KM * K * (double)Nv / ms / 1e9);
      ms = time_ms([&] { cublasDsyrk(h, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, K, Nv, &one, Z, ld, &zero, C, K); }, 5);
      printf(" \"cublas_dsyrk_gramshape_ms\": %.3f, \"cublas_dsyrk_full_equiv_tflops\": %.2f\n", ms,
             2.0 * K * K * (double)Nv / ms / 1e9);
      CK(cudaFree(Z)); CK(cudaFree(C));
    }
    cublasDestroy(h);
  }
  printf("}\n");
  return 0;
}
