// tvlqr.cu -- time-varying LQR pass along solved landing trajectories, batched (sm_100a, FP64).
//
// Replaces, for a whole sweep, what optimizations/landing/quadruped_SRBM_NLP.m:428-497 does per trajectory through
// CasADi functions: variational single-rigid-body dynamics A(t) [24x24], B(t) [24x12]
// (utilities_general/srbm-utilities/generateVariationalDynamics.m:9-62) and explicit-Euler backward integration of
// the Riccati differential equation  Pdot = A'P + PA - P B R^-1 B'P + Q  (generateRiccatiIntegrator.m:24,49-53),
// returning P(t_k) and the feedback gains K_k = R^-1 B_k' P_k.
//
// One CTA (192 threads = 576 / 3 entries of P per thread) per trajectory; P, A, B, A'P and PB live in shared memory;
// per step 11.5k multiply-adds (A and B are used with their fixed sparsity) and one 4.6 kB store of P (+ 2.3 kB of K):
// the kernel is bound by the sequential chain of n_steps small products (four barriers per step).
#include "kernels.cuh"

namespace srb {
namespace {

constexpr int TNT = 192, NSV = 24, NCV = 12, LDA = 25, LDB = 13;

__device__ __forceinline__ double skew_entry(const double* v, int i, int j) {  // skew(v)[i][j]
  if (i == j) return 0.0;
  const int k = 3 - i - j;
  const double s = ((j - i + 3) % 3 == 1) ? -1.0 : 1.0;
  return s * v[k];
}

__global__ void __launch_bounds__(TNT, 5) k_tvlqr(TvlqrArgs a) {
  __shared__ double P[NSV * LDA], A[NSV * LDA], W[NSV * LDA], Bm[NSV * LDB], S[NSV * LDB];
  __shared__ double xd[24], ud[12], Rt[9], Ibi[9], tau[3], fsum[3], Ibom[3], Rinv[12], sc[6];
  const long long b = blockIdx.x;
  const int tid = threadIdx.x, N = a.N;
  const double* x = a.x_star + b * a.nx;
  const landing_tvlqr& q = a.par;
  if (tid < 9) {  // inverse of the 3x3 inertia (adjugate)
    const double* I = q.Ib;
    const double det = I[0] * (I[4] * I[8] - I[5] * I[7]) - I[1] * (I[3] * I[8] - I[5] * I[6]) + I[2] * (I[3] * I[7] - I[4] * I[6]);
    const int i = tid / 3, j = tid % 3, i1 = (j + 1) % 3, i2 = (j + 2) % 3, j1 = (i + 1) % 3, j2 = (i + 2) % 3;
    Ibi[tid] = (I[i1 * 3 + j1] * I[i2 * 3 + j2] - I[i1 * 3 + j2] * I[i2 * 3 + j1]) / det;
  }
  for (int e = tid; e < NSV * NSV; e += TNT) P[(e / NSV) * LDA + e % NSV] = q.F[e];
  // A and B keep their zero pattern: only the structural entries are rewritten at every time point
  for (int e = tid; e < NSV * LDA; e += TNT) A[e] = 0.0;
  for (int e = tid; e < NSV * LDB; e += TNT) Bm[e] = 0.0;
  if (tid >= 32 && tid < 44) Rinv[tid - 32] = 1.0 / q.R[tid - 32];
  __syncthreads();
  const double hk = q.T / (double)(N - 1);
  // knot interval of the reference point: the smallest ko with t <= (ko + 1) hk (capped at N - 2).  Searched once for the
  // last time point; t only decreases afterwards, so the interval of the previous step is walked down (the search from 0 at
  // every step was 14 % of the kernel's instructions, a dependent I2F-DMUL-DSETP chain that every thread repeated).
  int ko = 0;
  {
    const double t_end = (q.n_steps - 1) * q.dt;
    while (t_end > (ko + 1) * hk && ko < N - 2) ko++;
  }
  for (int k = q.n_steps - 1; k >= 0; k--) {
    // reference point at t = k dt (quadruped_SRBM_NLP.m:478-487)
    const double t_int = k * q.dt;
    while (ko > 0 && t_int <= ko * hk) ko--;
    const double al = ((ko + 1) * hk - t_int) / hk;
    if (tid < 12) xd[tid] = al * x[12 * ko + tid] + (1.0 - al) * x[12 * (ko + 1) + tid];
    else if (tid < 24) xd[tid] = x[12 * N + 24 * ko + (tid - 12)];
    else if (tid < 36) ud[tid - 24] = x[12 * N + 24 * ko + 12 + (tid - 24)];
    __syncthreads();
    if (tid < 3) sincos(xd[3 + tid], &sc[tid], &sc[3 + tid]);  // (three lanes of warp 0, one angle each)
    __syncwarp();
    if (tid == 0) {  // R' = rpyToRotMat(rpy) (body -> world), torque, force sum, Ib * omega
      const double sr = sc[0], cr = sc[3], sp = sc[1], cp = sc[4], sy = sc[2], cy = sc[5];
      Rt[0] = cy * cp; Rt[1] = cy * sp * sr - sy * cr; Rt[2] = cy * sp * cr + sy * sr;
      Rt[3] = sy * cp; Rt[4] = sy * sp * sr + cy * cr; Rt[5] = sy * sp * cr - cy * sr;
      Rt[6] = -sp; Rt[7] = cp * sr; Rt[8] = cp * cr;
      double tw[3] = {0, 0, 0}, fs[3] = {0, 0, 0};
      for (int l = 0; l < 4; l++) {
        const double r0 = xd[12 + 3 * l] - xd[0], r1 = xd[13 + 3 * l] - xd[1], r2 = xd[14 + 3 * l] - xd[2];
        const double f0 = ud[3 * l], f1 = ud[3 * l + 1], f2 = ud[3 * l + 2];
        tw[0] += r1 * f2 - r2 * f1; tw[1] += r2 * f0 - r0 * f2; tw[2] += r0 * f1 - r1 * f0;
        fs[0] += f0; fs[1] += f1; fs[2] += f2;
      }
      for (int i = 0; i < 3; i++) {
        tau[i] = Rt[3 * i] * tw[0] + Rt[3 * i + 1] * tw[1] + Rt[3 * i + 2] * tw[2];
        fsum[i] = fs[i];
        Ibom[i] = q.Ib[3 * i] * xd[6] + q.Ib[3 * i + 1] * xd[7] + q.Ib[3 * i + 2] * xd[8];
      }
    }
    __syncthreads();
    // A, B (generateVariationalDynamics.m:32-55): 11 blocks of 3x3 on the omega rows + the constant blocks
    if (tid < 99) {
      const int blk = tid / 9, i = (tid % 9) / 3, j = tid % 3;
      double m[3];  // column j of the 3x3 matrix that Ib_inv multiplies
      if (blk == 0) {                       // skew(tau)
        for (int r = 0; r < 3; r++) m[r] = skew_entry(tau, r, j);
      } else if (blk == 2) {                // skew(Ib om) - skew(om) Ib
        const double* om = xd + 6;
        for (int r = 0; r < 3; r++) {
          double v = skew_entry(Ibom, r, j);
          for (int c = 0; c < 3; c++) v -= skew_entry(om, r, c) * q.Ib[3 * c + j];
          m[r] = v;
        }
      } else {                              // R' skew(v) (sign below)
        double v[3];
        if (blk == 1) { v[0] = fsum[0]; v[1] = fsum[1]; v[2] = fsum[2]; }
        else if (blk < 7) { const int l = blk - 3; v[0] = -ud[3 * l]; v[1] = -ud[3 * l + 1]; v[2] = -ud[3 * l + 2]; }
        else { const int l = blk - 7; v[0] = xd[12 + 3 * l] - xd[0]; v[1] = xd[13 + 3 * l] - xd[1]; v[2] = xd[14 + 3 * l] - xd[2]; }
        for (int r = 0; r < 3; r++) {
          double acc = 0.0;
          for (int c = 0; c < 3; c++) acc += Rt[3 * r + c] * skew_entry(v, c, j);
          m[r] = acc;
        }
      }
      const double val = Ibi[3 * i] * m[0] + Ibi[3 * i + 1] * m[1] + Ibi[3 * i + 2] * m[2];
      if (blk == 0) A[(6 + i) * LDA + 3 + j] = val;
      else if (blk == 1) A[(6 + i) * LDA + j] = val;
      else if (blk == 2) A[(6 + i) * LDA + 6 + j] = val;
      else if (blk < 7) A[(6 + i) * LDA + 12 + 3 * (blk - 3) + j] = val;
      else Bm[(6 + i) * LDB + 3 * (blk - 7) + j] = val;
    } else if (tid < 108) {
      const int i = (tid - 99) / 3, j = (tid - 99) % 3;
      A[(3 + i) * LDA + 3 + j] = -skew_entry(xd + 6, i, j);
      if (i == j) {
        A[i * LDA + 9 + i] = 1.0;
        A[(3 + i) * LDA + 6 + i] = 1.0;
        for (int l = 0; l < 4; l++) Bm[(9 + i) * LDB + 3 * l + i] = 1.0 / q.mass;
      }
    } else if (tid < 120) {
      const int i = tid - 108;
      A[(12 + i) * LDA + 12 + i] = -0.00001;
    }
    __syncthreads();
    // W = A'P and S = P B with the fixed sparsity of A and B (generateVariationalDynamics.m:32-55): only the omega rows
    // 6..8 of A are dense (columns 0-8, 12-23); the rest is I blocks, -skew(omega) and the -1e-5 diagonal; B has the
    // omega rows and I/m on the v rows.  6 + 4 multiply-adds per entry instead of 24 + 24.
    for (int e = tid; e < NSV * NSV; e += TNT) {
      const int i = e / NSV, j = e % NSV;
      double acc = A[6 * LDA + i] * P[6 * LDA + j] + A[7 * LDA + i] * P[7 * LDA + j] + A[8 * LDA + i] * P[8 * LDA + j];
      if (i >= 3 && i < 6) acc += A[3 * LDA + i] * P[3 * LDA + j] + A[4 * LDA + i] * P[4 * LDA + j] + A[5 * LDA + i] * P[5 * LDA + j];
      else if (i >= 6 && i < 9) acc += P[(i - 3) * LDA + j];
      else if (i >= 9 && i < 12) acc += P[(i - 9) * LDA + j];
      else if (i >= 12) acc += -0.00001 * P[i * LDA + j];
      W[i * LDA + j] = acc;
    }
    for (int e = tid; e < NSV * NCV; e += TNT) {
      const int i = e / NCV, c = e % NCV;
      S[i * LDB + c] = P[i * LDA + 6] * Bm[6 * LDB + c] + P[i * LDA + 7] * Bm[7 * LDB + c] + P[i * LDA + 8] * Bm[8 * LDB + c] +
                       P[i * LDA + 9 + c % 3] * Bm[(9 + c % 3) * LDB + c];
    }
    __syncthreads();
    // outputs of this time point, then the Euler step towards t - dt
    double* Pout = a.P_out ? a.P_out + (b * q.n_steps + k) * (NSV * NSV) : nullptr;
    double* Kout = a.K_out ? a.K_out + (b * q.n_steps + k) * (NCV * NSV) : nullptr;
    double pn[3];
    int cnt = 0;
    for (int e = tid; e < NSV * NSV; e += TNT, cnt++) {
      const int i = e / NSV, j = e % NSV;
      const double pij = P[i * LDA + j];
      if (Pout) Pout[e] = pij;
      double acc = W[i * LDA + j] + W[j * LDA + i] + q.Q[e];
#pragma unroll
      for (int c = 0; c < NCV; c++) acc -= S[i * LDB + c] * Rinv[c] * S[j * LDB + c];
      pn[cnt] = pij + q.dt * acc;
    }
    if (Kout)
      for (int e = tid; e < NCV * NSV; e += TNT) Kout[e] = S[(e % NSV) * LDB + e / NSV] * Rinv[e / NSV];
    __syncthreads();
    cnt = 0;
    for (int e = tid; e < NSV * NSV; e += TNT, cnt++) P[(e / NSV) * LDA + e % NSV] = pn[cnt];
    __syncthreads();
  }
}

}  // namespace

int launch_tvlqr(const TvlqrArgs& a, long long B, cudaStream_t st) {
  k_tvlqr<<<(unsigned)B, TNT, 0, st>>>(a);
  return 1;
}

}  // namespace srb
