// Dependent-chain latencies of the FP64 operations on the solver's critical path (one warp, clock64).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsqrt_fast(double a) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
  const double t = a * y0, r = fma(-t, 0.5 * y0, 0.5);
  const double q = fma(r, 1.5, 1.0), yr = y0 * r;
  return fma(yr, q, y0);
}
__device__ __forceinline__ double rcp_fast(double a) {
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
  const double e = fma(-a, y0, 1.0);
  const double q = fma(e, e, e);
  return fma(y0, q, y0);
}
template <int OP> __global__ void k(double* out, long long* cyc, double seed, int n) {
  double x = seed + threadIdx.x * 1e-3;
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    if (OP == 0) x = fma(x, 1.0000001, 1e-9);
    if (OP == 1) x = rsqrt(x) + 1.5;
    if (OP == 2) x = rsqrt_fast(x) + 1.5;
    if (OP == 3) x = 1.0 / x + 1.5;
    if (OP == 4) x = rcp_fast(x) + 1.5;
    if (OP == 5) x = log(x) + 2.5;
    if (OP == 6) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) + 1e-9;
    if (OP == 7) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + 1e-9; __syncwarp(); }
    if (OP == 8) x = sqrt(x) + 1.5;
    if (OP == 9) x = x * 1.0000001;
    if (OP == 10) { __syncthreads(); x += 1e-9; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_acc(double* out) {
  double e1 = 0, e2 = 0;
  for (int i = threadIdx.x; i < 2000000; i += blockDim.x) {
    const double a = exp2(-40.0 + 80.0 * (i / 2000000.0)) * (1.0 + 0.37 * (i % 1000) / 1000.0);
    const double r1 = rsqrt_fast(a), r0 = 1.0 / sqrt(a);
    e1 = fmax(e1, fabs(r1 - r0) / r0);
    const double c1 = rcp_fast(a), c0 = 1.0 / a;
    e2 = fmax(e2, fabs(c1 - c0) / c0);
  }
  for (int o = 16; o; o >>= 1) { e1 = fmax(e1, __shfl_xor_sync(0xffffffffu, e1, o)); e2 = fmax(e2, __shfl_xor_sync(0xffffffffu, e2, o)); }
  if (threadIdx.x == 0) { out[0] = e1; out[1] = e2; }
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 8 * 256); cudaMalloc(&c, 8);
  const char* names[] = {"dfma", "rsqrt(lib)+dadd", "rsqrt_fast+dadd", "1/x+dadd", "rcp_fast+dadd", "log+dadd", "shfl+dadd", "sts/lds+dadd", "sqrt+dadd", "dmul", "syncthreads(256)+dadd"};
  const int n = 4096;
  for (int op = 0; op <= 10; op++) {
    for (int rep = 0; rep < 2; rep++) {
      const int nt = op == 10 ? 256 : 32;
      switch (op) {
        case 0: k<0><<<1, nt>>>(d, c, 1.3, n); break; case 1: k<1><<<1, nt>>>(d, c, 1.3, n); break;
        case 2: k<2><<<1, nt>>>(d, c, 1.3, n); break; case 3: k<3><<<1, nt>>>(d, c, 1.3, n); break;
        case 4: k<4><<<1, nt>>>(d, c, 1.3, n); break; case 5: k<5><<<1, nt>>>(d, c, 1.3, n); break;
        case 6: k<6><<<1, nt>>>(d, c, 1.3, n); break; case 7: k<7><<<1, nt>>>(d, c, 1.3, n); break;
        case 8: k<8><<<1, nt>>>(d, c, 1.3, n); break; case 9: k<9><<<1, nt>>>(d, c, 1.3, n); break;
        case 10: k<10><<<1, nt>>>(d, c, 1.3, n); break;
      }
      cudaDeviceSynchronize();
    }
    long long h; double x[32]; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(x, d, 8 * 32, cudaMemcpyDeviceToHost);
    printf("%-24s %7.1f cycles per dependent op   (x=%.17g)\n", names[op], (double)h / n, x[0]);
  }
  k_acc<<<1, 32>>>(d); cudaDeviceSynchronize();
  double e[2]; cudaMemcpy(e, d, 16, cudaMemcpyDeviceToHost);
  printf("max relative error: rsqrt_fast %.3g, rcp_fast %.3g (2^-53 = 1.1e-16)\n", e[0], e[1]);
  return 0;
}
