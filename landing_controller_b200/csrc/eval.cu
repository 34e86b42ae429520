// eval.cu -- batched evaluation kernels for the landing NLP functions (sm_100a, FP64).
//
// Replaces the reference's straight-line generated C
//   nlp_f :10995, nlp_g :11161, nlp_grad :22015, nlp_grad_f :52602, nlp_hess_l :53527,
//   nlp_jac_g :94014   (optimizations/landing/codegen_casadi/landingCtrller_IPOPT.c)
// for B scenarios at once.  Mapping: blockIdx.y = knot, threadIdx = scenario, so one warp is one
// knot of 32 consecutive scenarios; with the SoA layout every load/store of the warp is a single
// coalesced 256-byte transaction.  blockIdx.y == N-1 is the "boundary" slice (initial/terminal
// rows, objective, terminal Hessian diagonal).  HBM-bound: 8*(n_x+n_p+m+nnz) bytes per scenario.
#include "kernels.cuh"

namespace srb {

namespace {

constexpr int TPB = 128;

__device__ __forceinline__ void load_knot(const EvalArgs& a, int k, long long b, Knot& kn, bool last) {
  const int N = a.pl.N;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    kn.X[i] = a.x.get(12 * k + i, b);
    kn.Xn[i] = a.x.get(12 * (k + 1) + i, b);
    kn.c[i] = a.x.get(12 * N + 24 * k + i, b);
    kn.f[i] = a.x.get(12 * N + 24 * k + 12 + i, b);
    kn.cn[i] = last ? 0.0 : a.x.get(12 * N + 24 * (k + 1) + i, b);
  }
  kn.h = a.p.get(a.pl.off.dt + k, b);
  kn.mu = a.p.get(a.pl.off.mu, b);
  kn.mass = a.p.get(a.pl.off.mass, b);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    kn.Ib[i] = a.p.get(a.pl.off.Ib + i, b);
    kn.Ibinv[i] = a.p.get(a.pl.off.Ibinv + i, b);
  }
}

struct LamRow {  // multiplier of knot-local row
  CView lam;
  long long b;
  int base;
  __device__ __forceinline__ double operator()(int row) const { return lam.get(base + row, b); }
};

struct ScatterSink {  // g / jac / hess values straight into the CCS value arrays
  const EvalArgs& a;
  long long b;
  int rowbase;
  const int *jm, *hm;
  bool bad;
  __device__ __forceinline__ void g(int row, double v) {
    a.g.at(rowbase + row, b) = v;
    bad |= !isfinite(v);
  }
  __device__ __forceinline__ void j(int e, int, int, double v) {
    a.jac.at(__ldg(jm + e), b) = v;
    bad |= !isfinite(v);
  }
  __device__ __forceinline__ void h(int e, int, int, double v) {
    a.hess.at(__ldg(hm + e), b) = v;
    bad |= !isfinite(v);
  }
};

// J^T lam of one knot, dealt to NGP threads by VARIABLE: part q accumulates (and later stores) the own-knot variables
// it owns -- the six variables (c_l, f_l) of leg q and the states X_i with i % 4 == q -- so no two threads ever touch
// the same output, and the compile-time filter leaves each thread a quarter of the Jacobian entries and 9 accumulators
// (one thread per knot needs 255 registers + spills and runs at 8 warps per SM).  Measured on B200 (16k scenarios, N = 30,
// SoA): NGP = 1 0.56 ms, NGP = 4 0.84 ms -- the parts repeat the rotation / wrench sub-expressions and re-read x, which
// costs more than the registers it frees -- so one part is the default.  Dealing the entries by ROWS to the four warps of
// a 32-scenario CTA instead (the split of k_eval's Jacobian kernel; c_q / f_q sums complete in part q, the X_k partial
// sums through shared memory) was measured too: 0.60 ms at 2 CTAs per SM (255 registers), 1.00 ms at 3, 0.83 ms at 4.
#ifndef EVAL_NP_GRAD
#define EVAL_NP_GRAD 1
#endif
constexpr int NGP = EVAL_NP_GRAD;

__device__ __forceinline__ constexpr int grad_owner(int var) {  // var: 0-11 X | 12-23 c | 24-35 f ; X+, c+: nobody
  return var < 12 ? (var & 3) : (var < 36 ? ((var - 12) % 12) / 3 : -1);
}
template <int PART> struct GradSink {  // acc[var] += lam[row] * dg_row/dvar for the variables of this part
  LamRow lam;
  double acc[36];
  __device__ __forceinline__ void g(int, double) {}
  __device__ __forceinline__ void j(int, int row, int var, double v) { if (var < 36 && grad_owner(var) % NGP == PART) acc[var] += lam(row) * v; }
  __device__ __forceinline__ void h(int, int, int, double) {}
};

// ---- the knot's entries are dealt to NPART threads (blockIdx.z): every thread runs the same knot template behind a
// compile-time filter, so the compiler drops whatever feeds only other parts' entries.  One thread per knot needs
// 255 registers + 1-2 kB of spills and leaves 8 warps per SM; a quarter of the entries per thread fits the register
// file and puts four times as many store streams in flight (the kernels are HBM-write bound).
template <int PART, int NP, bool LAST> struct PartSink {  // NP = 1: everything; 2: legs {0,2} | {1,3}; 4: one leg each
  ScatterSink& s;
  __device__ __forceinline__ void g(int row, double v) { if (g_owner<LAST>(row) % NP == PART) s.g(row, v); }
  __device__ __forceinline__ void j(int e, int row, int var, double v) { if (j_owner<LAST>(row, var) % NP == PART) s.j(e, row, var, v); }
  __device__ __forceinline__ void h(int e, int va, int vb, double v) { if (h_owner(e, va, vb) % NP == PART) s.h(e, va, vb, v); }
};
template <int PART, int NP, bool WG, bool WJ, bool WH>
__device__ __forceinline__ void eval_part(const Knot& kn, ScatterSink& s, const LamRow& lam, bool last) {
  if (last) {
    PartSink<PART, NP, true> ps{s};
    knot_eval<true, WG, WJ, WH>(kn, ps, lam);
  } else {
    PartSink<PART, NP, false> ps{s};
    knot_eval<false, WG, WJ, WH>(kn, ps, lam);
  }
}

// Measured on B200 (tools/bench_eval.py, 16k scenarios, SoA): the Jacobian kernels gain from the split (59 % -> 75 % of
// the measured HBM copy bandwidth with four parts at 168 registers); nlp_g and nlp_hess_l lose (their loads of x and of
// the multipliers are repeated per part) and keep one thread per knot (88 % / 79 %).
#ifndef EVAL_NP_J
#define EVAL_NP_J 4
#endif
#ifndef EVAL_NP_H
#define EVAL_NP_H 1
#endif
#ifndef EVAL_NP_G
#define EVAL_NP_G 1
#endif
template <bool WG, bool WJ, bool WH, int NP>
__global__ void __launch_bounds__(TPB, NP > 1 ? 3 : 1) k_eval(EvalArgs a) {
  const long long b = (long long)blockIdx.x * TPB + threadIdx.x;
  const int k = blockIdx.y + a.k0, N = a.pl.N;
  if (b >= a.B) return;
  bool bad = false;
  if (k == N - 1) {
    if (blockIdx.z != 0) return;
    // ---- boundary slice: rows 0-35 (generate_landingCtrller_IPOPT.m:90-97), objective (:83-85)
    const int xo = 12 * (N - 1);
    if (WG) {
      for (int i = 0; i < 12; i++) a.g.at(i, b) = a.x.get(i, b);
      for (int i = 0; i < 6; i++) {
        const double q = a.x.get(xo + i, b), qd = a.x.get(xo + 6 + i, b);
        a.g.at(12 + i, b) = q;
        a.g.at(18 + i, b) = q;
        a.g.at(24 + i, b) = qd;
        a.g.at(30 + i, b) = qd;
      }
    }
    if (WJ)
      for (int i = 0; i < 36; i++) a.jac.at(__ldg(a.pl.jbnd + i), b) = 1.0;
    if (WH) {
      const double lf = a.lam_f.get(0, b);
      for (int i = 0; i < 12; i++) {
        const double v = 2.0 * a.p.get(a.pl.off.QN + i, b) * lf;
        a.hess.at(__ldg(a.pl.hterm + i), b) = v;
        bad |= !isfinite(v);
      }
    }
    if (a.f.p || a.grad_f.p) {
      double fv = 0.0;
      for (int i = 0; i < 12; i++) {
        const double d = a.x.get(xo + i, b) - a.p.get(xo + i, b);
        const double qn = a.p.get(a.pl.off.QN + i, b);
        fv += qn * d * d;
        if (a.grad_f.p) a.grad_f.at(xo + i, b) = 2.0 * qn * d;
      }
      if (a.f.p) a.f.at(0, b) = fv;
      bad |= !isfinite(fv);
    }
  } else {
    Knot kn;
    const bool last = (k == N - 2);
    load_knot(a, k, b, kn, last);
    ScatterSink s{a, b, 36 + 104 * k, a.pl.jmap + k * NJ_INT, a.pl.hmap + k * NH_INT, false};
    LamRow lam{a.lam_g, b, 36 + 104 * k};
    if (NP == 1) {
      eval_part<0, 1, WG, WJ, WH>(kn, s, lam, last);
    } else if (NP == 2) {
      if (blockIdx.z == 0) eval_part<0, 2, WG, WJ, WH>(kn, s, lam, last);  // (block-uniform)
      else eval_part<1, 2, WG, WJ, WH>(kn, s, lam, last);
    } else {
      switch (blockIdx.z) {
        case 0: eval_part<0, NP, WG, WJ, WH>(kn, s, lam, last); break;
        case 1: eval_part<1 % NP, NP, WG, WJ, WH>(kn, s, lam, last); break;
        case 2: eval_part<2 % NP, NP, WG, WJ, WH>(kn, s, lam, last); break;
        default: eval_part<3 % NP, NP, WG, WJ, WH>(kn, s, lam, last); break;
      }
    }
    bad = s.bad;
  }
  if (bad && a.status) a.status[b] = -1;
}

template <int PART>
__device__ __forceinline__ void grad_x_part(const EvalArgs& a, const Knot& kn, const LamRow& lam, int k, long long b, bool last) {
  const int N = a.pl.N;
  GradSink<PART> s;
  s.lam = lam;
#pragma unroll
  for (int i = 0; i < 36; i++) s.acc[i] = 0.0;
  if (last)
    knot_eval<true, false, true, false>(kn, s, lam);
  else
    knot_eval<false, false, true, false>(kn, s, lam);
  // what the previous knot (or the initial-state rows) contributes to X_k, c_k
  if (k == 0) {
#pragma unroll
    for (int i = 0; i < 12; i++)
      if (grad_owner(i) % NGP == PART) s.acc[i] += a.lam_g.get(i, b);
  } else {
    const int pbase = 36 + 104 * (k - 1);  // (knot k-1 is never the last knot: interior row layout)
#pragma unroll
    for (int i = 0; i < 12; i++)
      if (grad_owner(i) % NGP == PART) s.acc[i] += a.lam_g.get(pbase + (i < 6 ? i : (i < 9 ? i + 3 : i - 3)), b);
#pragma unroll
    for (int l = 0; l < 4; l++) {
      if (l % NGP != PART) continue;  // this part's legs
      const double fzp = a.x.get(12 * N + 24 * (k - 1) + 12 + 3 * l + 2, b);
      const int L = pbase + Rows<false>::leg(l);
#pragma unroll
      for (int i = 0; i < 3; i++) s.acc[12 + 3 * l + i] += fzp * (a.lam_g.get(L + 2 + i, b) + a.lam_g.get(L + 5 + i, b));
    }
  }
#pragma unroll
  for (int vv = 0; vv < 36; vv++)
    if (grad_owner(vv) % NGP == PART) a.grad_x.at(global_var(N, k, vv), b) = s.acc[vv];
}

// grad_gamma_x = lam_f grad f + J^T lam_g ; grad_gamma_p (nlp_grad :22015).
// grad_x is written by OWNER threads, no atomics and no pre-zeroing: thread (scenario, knot k) owns X_k, c_k, f_k and
// adds to its own knot's J^T lam what knot k-1 contributes through its X+ / c+ columns -- identity entries of the
// dynamics rows and f_z,l of knot k-1 in the no-slip rows, both in closed form (no second template evaluation).  The
// boundary thread owns X_{N-1}.  (The former version did 48-60 RED.ADD.F64 per thread on a pre-zeroed array: 29 % of the
// HBM bandwidth.)  grad_p is pre-zeroed; only the eight scalar parameters shared by all knots are accumulated atomically.
__device__ __forceinline__ void grad_boundary(const EvalArgs& a, long long b) {  // X_{N-1}: objective + terminal rows
  const int N = a.pl.N;
  const ParamOff& o = a.pl.off;
  const int xo = 12 * (N - 1), pb = 36 + 104 * (N - 2);  // rows of the last knot (80-row layout: dynamics first)
  const double lf = a.lam_f.get(0, b);
  for (int i = 0; i < 12; i++) {
    const double d = a.x.get(xo + i, b) - a.p.get(xo + i, b);
    const double qn = a.p.get(o.QN + i, b);
    if (a.grad_x.p) {
      const int r1 = i < 6 ? 12 + i : 24 + (i - 6);
      // X+ column of the last knot's dynamics rows: pos, rpy -> rows 0-5; omega -> rows 9-11; v -> rows 6-8
      const int dr = i < 6 ? i : (i < 9 ? i + 3 : i - 3);
      a.grad_x.at(xo + i, b) = lf * 2.0 * qn * d + a.lam_g.get(r1, b) + a.lam_g.get(r1 + 6, b) + a.lam_g.get(pb + dr, b);
    }
    if (a.grad_p.p) {
      a.grad_p.at(xo + i, b) = -lf * 2.0 * qn * d;
      a.grad_p.at(o.QN + i, b) = lf * d * d;
    }
  }
}

// parameter sensitivities of one knot's rows (grad_gamma_p; pre-zeroed, shared scalars accumulated atomically)
__device__ __forceinline__ void grad_p_knot(const EvalArgs& a, const Knot& kn, const LamRow& lam, int k, long long b, bool last) {
  const ParamOff& o = a.pl.off;
  if (a.grad_p.p) {
    // parameter sensitivities of the knot rows: dt_k, mu, mass, Ib, Ib_inv
    double sf, cf, st, ct, sp, cp;
    sincos(kn.X[3], &sf, &cf);
    sincos(kn.X[4], &st, &ct);
    sincos(kn.X[5], &sp, &cp);
    const double R[9] = {cp * ct, -cf * sp + sf * cp * st, sf * sp + cf * cp * st,
                         sp * ct, cf * cp + sf * sp * st,  -sf * cp + cf * sp * st,
                         -st,     sf * ct,                 cf * ct};
    double F[3] = {0, 0, 0}, tau[3] = {0, 0, 0};
#pragma unroll
    for (int l = 0; l < 4; l++) {
      double arm[3], t[3];
#pragma unroll
      for (int i = 0; i < 3; i++) arm[i] = kn.c[3 * l + i] - kn.X[i];
      cross3(arm, kn.f + 3 * l, t);
#pragma unroll
      for (int i = 0; i < 3; i++) { tau[i] += t[i]; F[i] += kn.f[3 * l + i]; }
    }
    double tb[3];
    mtv(R, tau, tb);
    const double* om = kn.X + 6;
    const double* Ib = kn.Ib;
    const double w[3] = {om[1] * (Ib[2] * om[2]) - om[2] * (Ib[1] * om[1]),
                         om[2] * (Ib[0] * om[0]) - om[0] * (Ib[2] * om[2]),
                         om[0] * (Ib[1] * om[1]) - om[1] * (Ib[0] * om[0])};
    const double ea = sf * om[1] + cf * om[2], eb = cf * om[1] - sf * om[2];
    const double e[3] = {om[0] + st / ct * ea, eb, ea / ct};
    double sdt = 0.0, smass = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double acc = F[i] / kn.mass + (i == 2 ? GRAV_Z : 0.0);
      sdt += -lam(i) * kn.X[9 + i] - lam(3 + i) * e[i] - lam(6 + i) * acc -
             lam(9 + i) * kn.Ibinv[i] * (tb[i] - w[i]);
      smass += lam(6 + i) * kn.h * F[i] / (kn.mass * kn.mass);
      atomicAdd(&a.grad_p.at(o.Ibinv + i, b), -lam(9 + i) * kn.h * (tb[i] - w[i]));
    }
    a.grad_p.at(o.dt + k, b) = sdt;
    atomicAdd(&a.grad_p.at(o.mass, b), smass);
    const double c0 = lam(9) * kn.h * kn.Ibinv[0], c1 = lam(10) * kn.h * kn.Ibinv[1],
                 c2 = lam(11) * kn.h * kn.Ibinv[2];
    atomicAdd(&a.grad_p.at(o.Ib + 0, b), c1 * (om[2] * om[0]) - c2 * (om[1] * om[0]));
    atomicAdd(&a.grad_p.at(o.Ib + 1, b), -c0 * (om[2] * om[1]) + c2 * (om[0] * om[1]));
    atomicAdd(&a.grad_p.at(o.Ib + 2, b), c0 * (om[1] * om[2]) - c1 * (om[0] * om[2]));
    double smu = 0.0;
    const int fr = last ? Rows<true>::fric : Rows<false>::fric;
#pragma unroll
    for (int l = 0; l < 4; l++)
      smu += -FRIC * kn.f[3 * l + 2] * (lam(fr + l) + lam(fr + 4 + l) + lam(fr + 8 + l) + lam(fr + 12 + l));
    atomicAdd(&a.grad_p.at(o.mu, b), smu);
  }
}

__global__ void __launch_bounds__(TPB, NGP > 1 ? 3 : 1) k_grad(EvalArgs a) {
  const long long b = (long long)blockIdx.x * TPB + threadIdx.x;
  const int k = blockIdx.y, N = a.pl.N;
  if (b >= a.B) return;
  if (k == N - 1) {
    if (blockIdx.z == 0) grad_boundary(a, b);
    return;
  }
  Knot kn;
  const bool last = (k == N - 2);
  load_knot(a, k, b, kn, last);
  LamRow lam{a.lam_g, b, 36 + 104 * k};
  if (a.grad_x.p) {
    switch (blockIdx.z) {  // (block-uniform)
      case 0: grad_x_part<0>(a, kn, lam, k, b, last); break;
      case 1: grad_x_part<1 % NGP>(a, kn, lam, k, b, last); break;
      case 2: grad_x_part<2 % NGP>(a, kn, lam, k, b, last); break;
      default: grad_x_part<3 % NGP>(a, kn, lam, k, b, last); break;
    }
  }
  if (blockIdx.z != 0) return;
  grad_p_knot(a, kn, lam, k, b, last);
}

// lbg(p), ubg(p) -- optistack_internal.cpp:742-870 applied to generate_landingCtrller_IPOPT.m:90-169
__global__ void __launch_bounds__(TPB) k_bounds(DevicePlan pl, long long B, CView p, View lb, View ub) {
  const long long b = (long long)blockIdx.x * TPB + threadIdx.x;
  const int k = blockIdx.y, N = pl.N;
  if (b >= B) return;
  const double INF = HUGE_VAL;
  const ParamOff& o = pl.off;
  if (k == N - 1) {
    for (int i = 0; i < 6; i++) {
      lb.at(i, b) = ub.at(i, b) = p.get(o.qinit + i, b);
      lb.at(6 + i, b) = ub.at(6 + i, b) = p.get(o.qdinit + i, b);
      lb.at(12 + i, b) = p.get(o.qtmin + i, b); ub.at(12 + i, b) = INF;
      lb.at(18 + i, b) = -INF; ub.at(18 + i, b) = p.get(o.qtmax + i, b);
      lb.at(24 + i, b) = p.get(o.qdtmin + i, b); ub.at(24 + i, b) = INF;
      lb.at(30 + i, b) = -INF; ub.at(30 + i, b) = p.get(o.qdtmax + i, b);
    }
    return;
  }
  const bool last = (k == N - 2);
  const int base = 36 + 104 * k;
  auto set = [&](int row, double l, double u) { lb.at(base + row, b) = l; ub.at(base + row, b) = u; };
  for (int i = 0; i < 12; i++) set(i, 0.0, 0.0);
  const double fmax = p.get(o.fmax, b), lmax = p.get(o.lleg, b);
  for (int l = 0; l < 4; l++) {
    set(12 + l, 0.0, fmax);
    const int L = 16 + (last ? 6 : 12) * l, K = L + (last ? 2 : 8);
    set(L, 0.0, INF);
    set(L + 1, -INF, 0.001);
    if (!last)
      for (int i = 0; i < 3; i++) { set(L + 2 + i, -INF, 0.01); set(L + 5 + i, -0.01, INF); }
    set(K, -0.15, 0.15);
    set(K + 1, -0.15, 0.15);
    set(K + 2, -0.30, 0.0);
    set(K + 3, -INF, lmax * lmax);
  }
  const int fr = last ? 40 : 64, stb = last ? 56 : 80;
  for (int i = 0; i < 16; i++) set(fr + i, -INF, 0.0);
  for (int i = 0; i < 6; i++) {
    set(stb + i, -INF, p.get(o.qmax + i, b));
    set(stb + 6 + i, p.get(o.qmin + i, b), INF);
    set(stb + 12 + i, -INF, p.get(o.qdmax + i, b));
    set(stb + 18 + i, p.get(o.qdmin + i, b), INF);
  }
}

struct ProblemDev {
  landing_problem pb;
};

// p and x0 from drop conditions (generate_landingCtrller_IPOPT.m:199-208,336)
__global__ void __launch_bounds__(TPB) k_build(DevicePlan pl, long long B, ProblemDev P, const double* __restrict__ dtv,
                                               const double* __restrict__ drops, View p, View x0) {
  const long long b = (long long)blockIdx.x * TPB + threadIdx.x;
  const int k = blockIdx.y, N = pl.N;  // k in [0, N]: knot columns, k == N -> scalar parameters
  if (b >= B) return;
  const landing_problem& pb = P.pb;
  const ParamOff& o = pl.off;
  const double* d = drops + 12 * b;
  if (k == N) {
    if (!p.p) return;
    for (int i = 0; i < 6; i++) {
      p.at(o.qmin + i, b) = pb.q_min[i]; p.at(o.qmax + i, b) = pb.q_max[i];
      p.at(o.qdmin + i, b) = pb.qd_min[i]; p.at(o.qdmax + i, b) = pb.qd_max[i];
      p.at(o.qinit + i, b) = d[i]; p.at(o.qdinit + i, b) = d[6 + i];
      p.at(o.qtmin + i, b) = pb.q_term_min[i]; p.at(o.qtmax + i, b) = pb.q_term_max[i];
      p.at(o.qdtmin + i, b) = pb.qd_term_min[i]; p.at(o.qdtmax + i, b) = pb.qd_term_max[i];
    }
    for (int i = 0; i < 12; i++) p.at(o.QN + i, b) = pb.QN[i];
    p.at(o.mu, b) = pb.mu;
    p.at(o.lleg, b) = pb.l_leg_max;
    p.at(o.fmax, b) = pb.f_max;
    p.at(o.mass, b) = pb.mass;
    for (int i = 0; i < 3; i++) { p.at(o.Ib + i, b) = pb.Ib[i]; p.at(o.Ibinv + i, b) = pb.Ib_inv[i]; }
    return;
  }
  const double t = (double)k / (double)(N - 1);
  double xr[12];
  for (int i = 0; i < 6; i++) {
    // unfused multiply-add so that Xref is bit-identical to the host-side construction
    xr[i] = (k == N - 1) ? pb.q_term_ref[i] : __dadd_rn(d[i], __dmul_rn(pb.q_term_ref[i] - d[i], t));
    xr[6 + i] = (k == N - 1) ? pb.qd_term_ref[i]
                             : __dadd_rn(d[6 + i], __dmul_rn(pb.qd_term_ref[i] - d[6 + i], t));
  }
  for (int i = 0; i < 12; i++) {
    if (p.p) p.at(12 * k + i, b) = xr[i];
    if (x0.p) x0.at(12 * k + i, b) = xr[i];
  }
  if (k < N - 1) {
    if (p.p) p.at(o.dt + k, b) = dtv[k];
    if (x0.p)
      for (int l = 0; l < 4; l++)
        for (int i = 0; i < 3; i++) {
          x0.at(12 * N + 24 * k + 3 * l + i, b) = xr[i] + pb.c_ref[3 * l + i];
          x0.at(12 * N + 24 * k + 12 + 3 * l + i, b) = 0.0;
        }
  }
}

// dst[c][r] = src[r][c]: 32 x 32 tiles through shared memory (33-column padding), 256 threads; both the reads and
// the writes of a warp are 256 contiguous bytes
__global__ void __launch_bounds__(256) k_transpose(const double* __restrict__ src, double* __restrict__ dst,
                                                   long long rows, long long cols) {
  __shared__ double tile[32][33];
  const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = ty; j < 32; j += 8)
    if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = src[(r0 + j) * cols + c0 + tx];
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8)
    if (c0 + j < cols && r0 + tx < rows) dst[(c0 + j) * rows + r0 + tx] = tile[tx][j];
}

}  // namespace

int launch_transpose(const double* src, double* dst, long long rows, long long cols, cudaStream_t st) {
  const dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  k_transpose<<<grid, 256, 0, st>>>(src, dst, rows, cols);
  return 1;
}

int launch_eval(const EvalArgs& a, cudaStream_t st) {
  int launches = 0;
  const dim3 grid((unsigned)((a.B + TPB - 1) / TPB), (unsigned)a.pl.N);
  const bool wg = a.g.p != nullptr, wj = a.jac.p != nullptr, wh = a.hess.p != nullptr;
  const bool wf = a.f.p != nullptr || a.grad_f.p != nullptr;
  if (a.grad_f.p) cudaMemsetAsync(a.grad_f.p, 0, sizeof(double) * a.pl.nx * a.B, st);
  if (a.status) cudaMemsetAsync(a.status, 0, sizeof(int) * a.B, st);
  if (wg || wj || wh || wf) {
    // one fused launch per requested output combination (nlp_jac_g returns g and jac together)
    EvalArgs c = a;
    dim3 gr = grid;
    constexpr int NPJ = EVAL_NP_J, NPH = EVAL_NP_H, NPG = EVAL_NP_G;
    if (wj && !wh) gr.z = NPJ;
    if (wh && !wj && !wg) gr.z = NPH;
    if (wg && !wj && !wh) gr.z = NPG;
    switch ((wg ? 1 : 0) | (wj ? 2 : 0) | (wh ? 4 : 0)) {
      case 0:  // f / grad_f only: just the boundary slice
        c.k0 = a.pl.N - 1;
        gr.y = 1;
        gr.z = 1;
        k_eval<false, false, false, 1><<<gr, TPB, 0, st>>>(c);
        break;
      case 1: k_eval<true, false, false, NPG><<<gr, TPB, 0, st>>>(c); break;
      case 2: k_eval<false, true, false, NPJ><<<gr, TPB, 0, st>>>(c); break;
      case 3: k_eval<true, true, false, NPJ><<<gr, TPB, 0, st>>>(c); break;
      case 4: k_eval<false, false, true, NPH><<<gr, TPB, 0, st>>>(c); break;
      case 5: k_eval<true, false, true, 1><<<gr, TPB, 0, st>>>(c); break;
      case 6: k_eval<false, true, true, 1><<<gr, TPB, 0, st>>>(c); break;
      default: k_eval<true, true, true, 1><<<gr, TPB, 0, st>>>(c); break;
    }
    launches++;
  }
  if (a.grad_x.p || a.grad_p.p) {
    if (a.grad_p.p) cudaMemsetAsync(a.grad_p.p, 0, sizeof(double) * a.pl.np * a.B, st);  // (grad_x: owner writes)
    dim3 gg = grid;
    gg.z = a.grad_x.p ? NGP : 1;
    k_grad<<<gg, TPB, 0, st>>>(a);
    launches++;
  }
  return launches;
}

int launch_bounds(const DevicePlan& pl, long long B, CView p, View lbg, View ubg, cudaStream_t st) {
  const dim3 grid((unsigned)((B + TPB - 1) / TPB), (unsigned)pl.N);
  k_bounds<<<grid, TPB, 0, st>>>(pl, B, p, lbg, ubg);
  return 1;
}

int launch_build(const DevicePlan& pl, long long B, const landing_problem& pb, const double* dtv, const double* drops,
                 View p, View x0, cudaStream_t st) {
  const dim3 grid((unsigned)((B + TPB - 1) / TPB), (unsigned)pl.N + 1);
  ProblemDev P{pb};
  k_build<<<grid, TPB, 0, st>>>(pl, B, P, dtv, drops, p, x0);
  return 1;
}

}  // namespace srb
