// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) throughput and dependent-chain latency on sm_100a, next to DFMA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dmma ubench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NCH independent accumulator tiles per warp
template <int NCH> __global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
  double c[NCH][2];
#pragma unroll
  for (int i = 0; i < NCH; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int NCH> __global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
  double c[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) c[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) c[i] = fma(c[i], 1.0000001, 1e-9);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += c[i];
  if (s == 123.456) out[0] = s;
}

__global__ void k_lat(double* out, long long* cyc, int n) {
  double c0 = 0.1, c1 = 0.2;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) dmma(c0, c1, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  double* d; long long* c;
  cudaMalloc(&d, 8 * 256); cudaMalloc(&c, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int ctas = 1; ctas <= 8; ctas *= 2) {
    const int grid = n_sm * ctas;
    float ms;
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_dmma<8><<<grid, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf_mma = 2.0 * 256.0 * 8 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;  // 8 warps x 8 tiles x 256 FMA
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_dfma<16><<<grid, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf_fma = 2.0 * 16.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    printf("%d CTAs/SM x 256 threads: DMMA m8n8k4 %.2f TFLOP/s   DFMA %.2f TFLOP/s\n", ctas, tf_mma, tf_fma);
  }
  {
    // one warp per SM sub-partition (4 warps per SM): what a single warp can pull
    float ms;
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_dmma<8><<<n_sm, 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("1 warp/SM: DMMA %.1f cycles per instruction (8 independent tiles)\n", ms * 1e-3 * 1.965e9 / (8.0 * iters));
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_dfma<16><<<n_sm, 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("1 warp/SM: DFMA %.1f cycles per instruction (16 independent chains)\n", ms * 1e-3 * 1.965e9 / (16.0 * iters));
  }
  k_lat<<<1, 32>>>(d, c, 4096); cudaDeviceSynchronize();
  k_lat<<<1, 32>>>(d, c, 4096); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("DMMA dependent chain: %.1f cycles per instruction\n", (double)h / 4096);
  return 0;
}
