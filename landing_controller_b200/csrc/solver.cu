// solver.cu -- batched primal-dual interior-point solver for the SRB landing NLP (sm_100a, FP64).
//
// Replaces, for a whole batch of drop conditions, what the reference does one scenario at a time
// through CasADi's Nlpsol + IPOPT (nlpsol.cpp:555-635 -> IpoptInterface::solve; option set
// generate_landingCtrller_IPOPT.m:232-263; callers main_scripts/landing_optimization.m:305-311,
// generate_data/generate_training_data_automated.m:130-136).
//
// B200 design
//  * ONE persistent kernel: one CTA per SM, 7 warps per CTA, ONE WARP = ONE SCENARIO for the whole
//    solve (all interior-point iterations, line searches and inertia corrections run on the device;
//    no host round trip, no lock-step between scenarios).  Warps pull scenario ids from an atomic
//    work queue, so a slow scenario never stalls the others.
//  * Evaluation: one LANE per knot (srb_knot.cuh) -> g rows and the sparse J / H entry lists.
//  * Linear algebra: slacks and bound multipliers are eliminated; the remaining equality-constrained
//    QP (linearised Euler dynamics) is solved by a Riccati recursion with state (X_k, c_k) [24] and
//    control (f_k, c_{k+1}) [24].  Each stage is built in shared memory straight from the entry
//    lists by a host-made "condensing schedule" (no atomics, deterministic), factored by a
//    warp-cooperative Cholesky (warp shuffles broadcast the pivots), and only the factors needed by
//    the forward sweep are written to the per-warp scratch in HBM/L2.
//  * Reductions (errors, step lengths, merit function) are warp-shuffle butterflies: every lane
//    ends up with the identical value, so all control flow is warp-uniform and deterministic.
//
// The algorithm mirrors oracle/ip_ref.c step by step (that file is the CPU restatement used only
// by the tests and the CPU baseline).  There is no CPU path in this library.
#include <algorithm>
#include <cstdio>
#include <map>
#include <vector>

#include "solver.cuh"
#include "srb_knot.cuh"

namespace srb {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS = 7;        // warps (= concurrent scenarios) per CTA; 1 CTA per SM
constexpr int NS = 24;          // stage state / control size
constexpr int NW = 48;          // stage variables X(12) c(12) f(12) c+(12)
constexpr int LDM = 49, LDP = 25, LDG = 36, LDGF = 37;  // leading dimensions (padding against bank conflicts)
constexpr int MAXFILTER = 64;
constexpr int RK = 104;         // rows per knot (interior numbering)

// ---- shared memory carve-up (doubles) per warp
constexpr int SM_M = 0;                       // 48 x 49
constexpr int SM_P = SM_M + NW * LDM;         // 24 x 25   P_{k+1}
constexpr int SM_G = SM_P + NS * LDP;         // 12 x 37
constexpr int SM_T = SM_G + 12 * LDG;         // 12 x 37   Pxx*G
constexpr int SM_V = SM_T + 12 * LDG;         // vectors
constexpr int V_Q = 0, V_QH = 48, V_R = 96, V_T = 108, V_YV = 120, V_PN = 144, V_XI = 168, V_U = 192,
              V_NEXT = V_XI /* dual_inf_x only; never live together with xi */, V_END = 216;
constexpr int SM_WARP = SM_V + V_END;         // 4032 doubles = 32256 B per warp
static_assert((WARPS * SM_WARP + 4 * (36 + 104)) * 8 <= 232448, "shared memory budget (227 KB per CTA)");
// per-CTA bound tables
constexpr int SM_TAB = WARPS * SM_WARP;       // lb[140] ub[140] lbo[140] ubo[140]
constexpr int NROWTAB = 36 + RK;
constexpr int SM_TOTAL = SM_TAB + 4 * NROWTAB;

struct KParams {
  int N, K, nx, MR;
  long long B;
  landing_problem pb;
  landing_options opt;
  const double* drops;
  const double* x0;
  double *x_star, *f_star, *lam_g, *viol;
  int *status, *iters;
  int* counter;
  double* scratch;
  long long slot;  // doubles per warp slot
  SolverTables tab;
};

struct Ws {  // pointers into one warp's scratch slot
  double *x, *xt, *dx;
  double *S, *Y, *ZL, *ZU, *G, *GT, *DS, *YN, *DZL, *DZU, *SIG, *YH;
  double *JL, *HL;
  double *Lf, *Yf, *Gf, *rf, *yvf, *PX, *PV, *L0;
  double *FT, *FP;
};

__host__ __device__ inline long long slot_doubles(int N) {
  const long long K = N - 1, nx = 36LL * N - 24, MR = 36 + RK * K;
  return 3 * nx + 12 * MR + K * (NJ_INT + NH_INT) + K * (576 + 576 + 432 + 12 + 24) + (K + 1) * (288 + 24) +
         144 + 2 * MAXFILTER + 64;
}

__device__ inline Ws carve(double* base, int N) {
  const long long K = N - 1, nx = 36LL * N - 24, MR = 36 + RK * K;
  Ws w;
  double* p = base;
  w.x = p; p += nx; w.xt = p; p += nx; w.dx = p; p += nx;
  w.S = p; p += MR; w.Y = p; p += MR; w.ZL = p; p += MR; w.ZU = p; p += MR; w.G = p; p += MR;
  w.GT = p; p += MR; w.DS = p; p += MR; w.YN = p; p += MR; w.DZL = p; p += MR; w.DZU = p; p += MR;
  w.SIG = p; p += MR; w.YH = p; p += MR;
  w.JL = p; p += K * NJ_INT; w.HL = p; p += K * NH_INT;
  w.Lf = p; p += K * 576; w.Yf = p; p += K * 576; w.Gf = p; p += K * 432; w.rf = p; p += K * 12;
  w.yvf = p; p += K * 24; w.PX = p; p += (K + 1) * 288; w.PV = p; p += (K + 1) * 24;
  w.L0 = p; p += 144; w.FT = p; p += MAXFILTER; w.FP = p; p += MAXFILTER;
  return w;
}

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ int wsumi(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// last-knot template row -> interior row numbering
template <bool LAST> __device__ __forceinline__ constexpr int rowmap(int r) {
  if (!LAST) return r;
  if (r < 16) return r;
  if (r < 40) return 16 + 12 * ((r - 16) / 6) + (((r - 16) % 6) < 2 ? ((r - 16) % 6) : ((r - 16) % 6) + 6);
  return r + 24;
}
__device__ __forceinline__ bool is_noslip(int rho) { return rho >= 16 && rho < 64 && ((rho - 16) % 12) >= 2 && ((rho - 16) % 12) < 8; }
// dynamics row (0..11: pos,rpy,v,om) <-> state index (pos,rpy,om,v)
__device__ __forceinline__ int dyn_state(int rho) { return rho < 6 ? rho : (rho < 9 ? rho + 3 : rho - 3); }

// row kinds
enum { ROW_EQ = 0, ROW_INEQ = 1, ROW_FREE = 2 };
__device__ __forceinline__ int row_kind(int idx, int K) {
  if (idx < 12) return ROW_EQ;
  if (idx < 36) return ROW_INEQ;
  const int k = (idx - 36) / RK, rho = (idx - 36) - k * RK;
  if (rho < 12) return ROW_EQ;
  if (k == K - 1 && is_noslip(rho)) return ROW_FREE;
  return ROW_INEQ;
}
__device__ __forceinline__ int row_tab(int idx) { return idx < 36 ? idx : 36 + (idx - 36) % RK; }

// ---- sinks for the lane-per-knot evaluation
template <bool LAST> struct ListSinkA {
  double *gp, *jl, *hl;  // g rows of this knot (interior numbering), entry lists of this knot
  const int *jmap, *hmap;
  __device__ __forceinline__ void g(int r, double v) { gp[rowmap<LAST>(r)] = v; }
  __device__ __forceinline__ void j(int e, int, int, double v) { jl[LAST ? __ldg(jmap + e) : e] = v; }
  __device__ __forceinline__ void h(int e, int, int, double v) { hl[LAST ? __ldg(hmap + e) : e] = v; }
};
template <bool LAST> struct GSink {
  double* gp;
  __device__ __forceinline__ void g(int r, double v) { gp[rowmap<LAST>(r)] = v; }
  __device__ __forceinline__ void j(int, int, int, double) {}
  __device__ __forceinline__ void h(int, int, int, double) {}
};
template <bool LAST> struct LamY {
  const double* y;  // multipliers of this knot's rows (interior numbering)
  __device__ __forceinline__ double operator()(int r) const { return y[rowmap<LAST>(r)]; }
};

__device__ __forceinline__ void load_knot(const KParams& P, const double* x, int k, Knot& kn) {
  const int N = P.N;
  const bool last = (k == N - 2);
#pragma unroll
  for (int i = 0; i < 12; i++) {
    kn.X[i] = x[12 * k + i];
    kn.Xn[i] = x[12 * (k + 1) + i];
    kn.c[i] = x[12 * N + 24 * k + i];
    kn.f[i] = x[12 * N + 24 * k + 12 + i];
    kn.cn[i] = last ? 0.0 : x[12 * N + 24 * (k + 1) + i];
  }
  kn.h = P.pb.T / (double)(N - 1);
  kn.mu = P.pb.mu;
  kn.mass = P.pb.mass;
#pragma unroll
  for (int i = 0; i < 3; i++) { kn.Ib[i] = P.pb.Ib[i]; kn.Ibinv[i] = P.pb.Ib_inv[i]; }
}

// g(x) (and, when LISTS, the J/H entry lists with multipliers Y) for all knots; returns f(x)
template <bool LISTS>
__device__ __noinline__ double eval_all(const KParams& P, const Ws& w, const double* x, double* gout, int lane) {
  const int N = P.N, K = P.K;
  for (int k = lane; k < K; k += 32) {
    Knot kn;
    load_knot(P, x, k, kn);
    double* gk = gout + 36 + RK * k;
    if (k == K - 1) {
      if (LISTS) {
        ListSinkA<true> s;
        s.gp = gk; s.jl = w.JL + (long long)k * NJ_INT; s.hl = w.HL + (long long)k * NH_INT;
        s.jmap = P.tab.jl_last; s.hmap = P.tab.hl_last;
        LamY<true> lam{w.Y + 36 + RK * k};
        knot_eval<true, true, true, true>(kn, s, lam);
      } else {
        GSink<true> s{gk};
        NoLam nl;
        knot_eval<true, true, false, false>(kn, s, nl);
      }
    } else {
      if (LISTS) {
        ListSinkA<false> s;
        s.gp = gk; s.jl = w.JL + (long long)k * NJ_INT; s.hl = w.HL + (long long)k * NH_INT;
        s.jmap = nullptr; s.hmap = nullptr;
        LamY<false> lam{w.Y + 36 + RK * k};
        knot_eval<false, true, true, true>(kn, s, lam);
      } else {
        GSink<false> s{gk};
        NoLam nl;
        knot_eval<false, true, false, false>(kn, s, nl);
      }
    }
  }
  // boundary rows 0..35 and objective (generate_landingCtrller_IPOPT.m:83-97)
  double fl = 0.0;
  const int xo = 12 * (N - 1);
  if (lane < 12) {
    gout[lane] = x[lane];
    const double q = x[xo + lane];
    const int r1 = lane < 6 ? 12 + lane : 24 + (lane - 6);
    gout[r1] = q;
    gout[r1 + 6] = q;
    const double ref = lane < 6 ? P.pb.q_term_ref[lane] : P.pb.qd_term_ref[lane - 6];
    fl = P.pb.QN[lane] * (q - ref) * (q - ref);
  }
  __syncwarp();
  return wsum(fl);
}

// in-place lower Cholesky of the n x n block at A (leading dimension ld), n <= 32
__device__ __forceinline__ bool chol_warp(double* A, int n, int ld, int lane) {
  for (int j = 0; j < n; j++) {
    double v = 0.0;
    if (lane >= j && lane < n) {
      v = A[lane * ld + j];
      for (int l = 0; l < j; l++) v -= A[lane * ld + l] * A[j * ld + l];
    }
    const double d = __shfl_sync(FULL, v, j);
    if (!(d > 1e-14)) return false;
    const double sd = sqrt(d);
    if (lane == j) A[j * ld + j] = sd;
    else if (lane > j && lane < n) A[lane * ld + j] = v / sd;
    __syncwarp();
  }
  return true;
}

struct Scal {  // warp-uniform scalars of one scenario
  double mu, f;
};

// ---------------------------------------------------------------- backward sweep
// Condenses every stage from the entry lists, factors it and propagates P, p.  false -> not PD.
__device__ __noinline__ bool backward_sweep(const KParams& P, const Ws& w, double* sm, const double* tab,
                                            double dwreg, int lane) {
  const int N = P.N, K = P.K;
  double* M = sm + SM_M;
  double* Pn = sm + SM_P;
  double* Gs = sm + SM_G;
  double* Ts = sm + SM_T;
  double* V = sm + SM_V;
  const SolverTables& tb = P.tab;
  // terminal block P_K, p_K
  for (int i = lane; i < NS * LDP; i += 32) Pn[i] = 0.0;
  if (lane < NS) V[V_PN + lane] = 0.0;
  __syncwarp();
  if (lane < 12) {
    const int r1 = lane < 6 ? 12 + lane : 24 + (lane - 6), r2 = r1 + 6;
    const double q = w.x[12 * (N - 1) + lane];
    const double ref = lane < 6 ? P.pb.q_term_ref[lane] : P.pb.qd_term_ref[lane - 6];
    Pn[lane * LDP + lane] = 2.0 * P.pb.QN[lane] + w.SIG[r1] + w.SIG[r2] + dwreg;
    V[V_PN + lane] = 2.0 * P.pb.QN[lane] * (q - ref) + w.YH[r1] + w.YH[r2];
  }
  __syncwarp();
  for (int i = lane; i < 288; i += 32) w.PX[(long long)K * 288 + i] = Pn[(i / 24) * LDP + (i % 24)];
  if (lane < NS) w.PV[K * 24 + lane] = V[V_PN + lane];

  for (int k = K - 1; k >= 0; k--) {
    const double* Jk = w.JL + (long long)k * NJ_INT;
    const double* Hk = w.HL + (long long)k * NH_INT;
    const int rb = 36 + RK * k;
    // 1. clear
    for (int i = lane; i < NW * LDM; i += 32) M[i] = 0.0;
    for (int i = lane; i < 12 * LDG; i += 32) Gs[i] = 0.0;
    __syncwarp();
    // 2. Hessian entries, dynamics Jacobian, defects
    for (int e = lane; e < NH_INT; e += 32) {
      const int t = __ldg(tb.h_t + e), i = t / NW, j = t - i * NW;
      const double v = Hk[e];
      M[i * LDM + j] += v;
      if (i != j) M[j * LDM + i] += v;
    }
    for (int n = lane; n < tb.g_n; n += 32) {
      const int t = __ldg(tb.g_t + n);
      Gs[(t / 36) * LDG + (t % 36)] = -Jk[__ldg(tb.g_e + n)];
    }
    if (lane < 12) V[V_R + dyn_state(lane)] = -w.G[rb + lane];
    __syncwarp();
    // 3. sigma-weighted outer products of the inequality rows, stage gradient
    for (int t = lane; t < tb.t_n; t += 32) {
      double acc = 0.0;
      const int p0 = __ldg(tb.t_ptr + t), p1 = __ldg(tb.t_ptr + t + 1);
      for (int p = p0; p < p1; p++) {
        const int term = __ldg(tb.t_terms + p);
        acc += w.SIG[rb + (term >> 20)] * Jk[(term >> 10) & 1023] * Jk[term & 1023];
      }
      const int ij = __ldg(tb.t_ij + t), i = ij / NW, j = ij - i * NW;
      M[i * LDM + j] += acc;
      if (i != j) M[j * LDM + i] += acc;
    }
    for (int v = lane; v < NW; v += 32) {
      double acc = 0.0;
      const int p0 = __ldg(tb.q_ptr + v), p1 = __ldg(tb.q_ptr + v + 1);
      for (int p = p0; p < p1; p++) {
        const int term = __ldg(tb.q_terms + p);
        acc += w.YH[rb + (term >> 10)] * Jk[term & 1023];
      }
      V[V_Q + v] = acc;
    }
    __syncwarp();
    if (lane < 12) {
      for (int r = 0; r < 3; r++) M[(lane + 12 * r) * LDM + lane + 12 * r] += dwreg;
      if (k == K - 1) M[(36 + lane) * LDM + 36 + lane] += 1.0;  // dummy c+ of the last stage
    }
    // 4. add [G' Pxx G, G' Pxc; Pcx G, Pcc] and the vector terms
    for (int idx = lane; idx < 12 * 36; idx += 32) {
      const int i = idx / 36, j = idx - i * 36;
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < 12; l++) s += Pn[i * LDP + l] * Gs[l * LDG + j];
      Ts[i * LDG + j] = s;
    }
    if (lane < 12) {
      double s = V[V_PN + lane];
      for (int l = 0; l < 12; l++) s += Pn[lane * LDP + l] * V[V_R + l];
      V[V_T + lane] = s;
    }
    __syncwarp();
    for (int idx = lane; idx < 36 * 36; idx += 32) {
      const int i = idx / 36, j = idx - i * 36;
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < 12; l++) s += Gs[l * LDG + i] * Ts[l * LDG + j];
      M[i * LDM + j] += s;
    }
    for (int idx = lane; idx < 36 * 12; idx += 32) {
      const int i = idx / 12, j = idx - i * 12;
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < 12; l++) s += Gs[l * LDG + i] * Pn[l * LDP + 12 + j];
      M[i * LDM + 36 + j] += s;
      M[(36 + j) * LDM + i] += s;
    }
    for (int idx = lane; idx < 144; idx += 32) {
      const int i = idx / 12, j = idx - i * 12;
      M[(36 + i) * LDM + 36 + j] += Pn[(12 + i) * LDP + 12 + j];
    }
    for (int j = lane; j < NW; j += 32) {
      double s = V[V_Q + j];
      if (j < 36) {
        for (int l = 0; l < 12; l++) s += Gs[l * LDG + j] * V[V_T + l];
      } else {
        s += V[V_PN + 12 + (j - 36)];
        for (int l = 0; l < 12; l++) s += Pn[(12 + j - 36) * LDP + l] * V[V_R + l];
      }
      V[V_QH + j] = s;
    }
    __syncwarp();
    // 5. Cholesky of the control block (rows/cols 24..47), in place
    if (!chol_warp(M + 24 * LDM + 24, NS, LDM, lane)) return false;
    // 6. Y = L^-1 M_ux (lane = column), yv = L^-1 qh_u (lane 24)
    if (lane < 25) {
      for (int i = 0; i < NS; i++) {
        double v = lane < NS ? M[(24 + i) * LDM + lane] : V[V_QH + 24 + i];
        for (int l = 0; l < i; l++)
          v -= M[(24 + i) * LDM + 24 + l] * (lane < NS ? M[(24 + l) * LDM + lane] : V[V_YV + l]);
        v /= M[(24 + i) * LDM + 24 + i];
        if (lane < NS) M[(24 + i) * LDM + lane] = v; else V[V_YV + i] = v;
      }
    }
    __syncwarp();
    // 7. P_k = M_xx - Y'Y, p_k = qh_x - Y' yv
    for (int idx = lane; idx < NS * NS; idx += 32) {
      const int i = idx / NS, j = idx - i * NS, a = i < j ? i : j, b = i < j ? j : i;
      double v = M[a * LDM + b];
#pragma unroll 8
      for (int l = 0; l < NS; l++) v -= M[(24 + l) * LDM + a] * M[(24 + l) * LDM + b];
      Pn[i * LDP + j] = v;
    }
    if (lane < NS) {
      double v = V[V_QH + lane];
      for (int l = 0; l < NS; l++) v -= M[(24 + l) * LDM + lane] * V[V_YV + l];
      V[V_PN + lane] = v;
    }
    __syncwarp();
    // 8. keep what the forward sweep needs
    double* Lf = w.Lf + (long long)k * 576;
    double* Yf = w.Yf + (long long)k * 576;
    double* Gf = w.Gf + (long long)k * 432;
    for (int idx = lane; idx < 576; idx += 32) {
      const int i = idx / NS, j = idx - i * NS;
      Lf[idx] = M[(24 + i) * LDM + 24 + j];
      Yf[idx] = M[(24 + i) * LDM + j];
    }
    for (int idx = lane; idx < 432; idx += 32) Gf[idx] = Gs[(idx / 36) * LDG + (idx % 36)];
    for (int i = lane; i < 288; i += 32) w.PX[(long long)k * 288 + i] = Pn[(i / 24) * LDP + (i % 24)];
    if (lane < 12) w.rf[k * 12 + lane] = V[V_R + lane];
    if (lane < NS) {
      w.yvf[k * 24 + lane] = V[V_YV + lane];
      w.PV[k * 24 + lane] = V[V_PN + lane];
    }
    __syncwarp();
  }
  // free initial foot positions: Cholesky of P_0's (c,c) block
  for (int idx = lane; idx < 144; idx += 32) M[(idx / 12) * LDM + (idx % 12)] = Pn[(12 + idx / 12) * LDP + 12 + (idx % 12)];
  __syncwarp();
  if (!chol_warp(M, 12, LDM, lane)) return false;
  for (int idx = lane; idx < 144; idx += 32) w.L0[idx] = M[(idx / 12) * LDM + (idx % 12)];
  // P_0 row block for the forward start (Pn still holds P_0, V_PN p_0)
  __syncwarp();
  return true;
}

struct StepInfo {
  double a_pr, a_du, dphi_bar, phi_bar, theta;  // barrier parts of dphi / phi, and theta at the current point
};

// per-row step recovery for one inequality row: ds, new multiplier, dz, step limits, merit pieces
__device__ __forceinline__ void row_step(const Ws& w, int idx, double lb, double ub, double jdx, double mu, double tau,
                                         StepInfo& si) {
  const double s = w.S[idx], rd = w.G[idx] - s;
  const double ds = jdx + rd;
  double yn = w.SIG[idx] * ds, dzl = 0.0, dzu = 0.0;
  si.theta += fabs(rd);
  if (isfinite(lb)) {
    const double d = s - lb, z = w.ZL[idx];
    yn -= mu / d;
    dzl = mu / d - z - z / d * ds;
    if (ds < 0) si.a_pr = fmin(si.a_pr, -tau * d / ds);
    if (dzl < 0) si.a_du = fmin(si.a_du, -tau * z / dzl);
    si.dphi_bar -= mu * ds / d;
    si.phi_bar -= mu * log(d);
  }
  if (isfinite(ub)) {
    const double d = ub - s, z = w.ZU[idx];
    yn += mu / d;
    dzu = mu / d - z + z / d * ds;
    if (ds > 0) si.a_pr = fmin(si.a_pr, tau * d / ds);
    if (dzu < 0) si.a_du = fmin(si.a_du, -tau * z / dzu);
    si.dphi_bar += mu * ds / d;
    si.phi_bar -= mu * log(d);
  }
  w.DS[idx] = ds;
  w.YN[idx] = yn;
  w.DZL[idx] = dzl;
  w.DZU[idx] = dzu;
}

// ---------------------------------------------------------------- forward sweep
__device__ __noinline__ void forward_sweep(const KParams& P, const Ws& w, double* sm, const double* tab,
                                           const double* drop, double mu, double tau, StepInfo& si, int lane) {
  const int N = P.N, K = P.K;
  double* Ls = sm + SM_M;            // 24 x 25
  double* Ys = Ls + NS * LDP;        // 24 x 25
  double* Gs = Ys + NS * LDP;        // 12 x 37
  double* PXn = Gs + 12 * LDGF;       // 12 x 25
  double* Pn = sm + SM_P;            // holds P_0 on entry
  double* V = sm + SM_V;
  double* xi = V + V_XI;
  double* u = V + V_U;
  const SolverTables& tb = P.tab;
  si.a_pr = 1.0; si.a_du = 1.0; si.dphi_bar = 0.0; si.phi_bar = 0.0; si.theta = 0.0;
  // initial state step and free initial feet
  if (lane < 12) {
    const double c0 = w.G[lane] - drop[lane];
    xi[lane] = -c0;
    si.theta += fabs(c0);
  }
  __syncwarp();
  {
    // b = -(p_c + P_cx dX0); solve L0 L0' dc0 = b   (12 x 12; lane 0..11 own rows, shuffles broadcast)
    double b = 0.0;
    if (lane < 12) {
      b = -V[V_PN + 12 + lane];
      for (int l = 0; l < 12; l++) b -= Pn[(12 + lane) * LDP + l] * xi[l];
    }
    for (int i = 0; i < 12; i++) {  // forward
      const double bi = __shfl_sync(FULL, b, i) / w.L0[i * 12 + i];
      if (lane == i) b = bi;
      else if (lane > i && lane < 12) b -= w.L0[lane * 12 + i] * bi;
    }
    for (int i = 11; i >= 0; i--) {  // backward
      const double bi = __shfl_sync(FULL, b, i) / w.L0[i * 12 + i];
      if (lane == i) b = bi;
      else if (lane < i) b -= w.L0[i * 12 + lane] * bi;
    }
    if (lane < 12) xi[12 + lane] = b;
  }
  __syncwarp();
  if (lane < 12) {  // multipliers of the initial-state rows: -dV0/dX
    double v = V[V_PN + lane];
    for (int l = 0; l < NS; l++) v += Pn[lane * LDP + l] * xi[l];
    w.YN[lane] = -v;
    w.DS[lane] = 0.0;
  }
  __syncwarp();
  for (int k = 0; k < K; k++) {
    const bool last = (k == K - 1);
    const double* Lf = w.Lf + (long long)k * 576;
    const double* Yf = w.Yf + (long long)k * 576;
    const double* Gf = w.Gf + (long long)k * 432;
    for (int idx = lane; idx < 576; idx += 32) {
      Ls[(idx / NS) * LDP + (idx % NS)] = Lf[idx];
      Ys[(idx / NS) * LDP + (idx % NS)] = Yf[idx];
    }
    for (int idx = lane; idx < 432; idx += 32) Gs[(idx / 36) * LDGF + (idx % 36)] = Gf[idx];
    for (int idx = lane; idx < 288; idx += 32) PXn[(idx / NS) * LDP + (idx % NS)] = w.PX[(long long)(k + 1) * 288 + idx];
    __syncwarp();
    // u = -L^-T (Y xi + yv)
    double my = 0.0;
    if (lane < NS) {
      double v = w.yvf[k * 24 + lane];
#pragma unroll 8
      for (int l = 0; l < NS; l++) v += Ys[lane * LDP + l] * xi[l];
      my = -v;
    }
    for (int i = NS - 1; i >= 0; i--) {
      const double ui = __shfl_sync(FULL, my, i) / Ls[i * LDP + i];
      if (lane == i) my = ui;
      else if (lane < i) my -= Ls[i * LDP + lane] * ui;
    }
    if (lane < NS) u[lane] = my;
    __syncwarp();
    // next state
    double xn = 0.0;
    if (lane < 12) {
      double v = w.rf[k * 12 + lane];
#pragma unroll 4
      for (int l = 0; l < NS; l++) v += Gs[lane * LDGF + l] * xi[l];
#pragma unroll 4
      for (int l = 0; l < 12; l++) v += Gs[lane * LDGF + 24 + l] * u[l];
      xn = v;
    } else if (lane < NS) {
      xn = last ? 0.0 : u[lane];
    }
    // step of this knot's variables
    if (lane < 12) {
      w.dx[12 * k + lane] = xi[lane];
      w.dx[12 * N + 24 * k + lane] = xi[12 + lane];
      w.dx[12 * N + 24 * k + 12 + lane] = u[lane];
    }
    // inequality rows of knot k: ds = J_row . dw + (g - s), multipliers, step limits
    const double* Jk = w.JL + (long long)k * NJ_INT;
    const int rb = 36 + RK * k;
    for (int rho = 12 + lane; rho < RK; rho += 32) {
      if (last && is_noslip(rho)) continue;
      double jdx = 0.0;
      const int p0 = __ldg(tb.r_ptr + rho), p1 = __ldg(tb.r_ptr + rho + 1);
      for (int p = p0; p < p1; p++) {
        const int term = __ldg(tb.r_terms + p), sv = term >> 10;
        jdx += Jk[term & 1023] * (sv < NS ? xi[sv] : u[sv - NS]);
      }
      row_step(w, rb + rho, tab[36 + rho], tab[NROWTAB + 36 + rho], jdx, mu, tau, si);
    }
    __syncwarp();
    if (lane < NS) xi[lane] = xn;
    __syncwarp();
    // costate = multiplier of the dynamics rows of knot k: -(P_{k+1} xi_{k+1} + p_{k+1})
    if (lane < 12) {
      double v = w.PV[(k + 1) * 24 + lane];
#pragma unroll 8
      for (int l = 0; l < NS; l++) v += PXn[lane * LDP + l] * xi[l];
      const int rho = lane < 6 ? lane : (lane < 9 ? lane + 3 : lane - 3);
      w.YN[rb + rho] = -v;
      w.DS[rb + rho] = 0.0;
      si.theta += fabs(w.G[rb + lane]);
    }
    __syncwarp();
  }
  if (lane < 12) w.dx[12 * (N - 1) + lane] = xi[lane];
  // terminal inequality rows 12..35: g = X_{N-1}
  if (lane < 24) {
    const int i = lane < 12 ? lane : lane - 12;           // state index
    const int r1 = (i < 6 ? 12 + i : 24 + (i - 6)) + (lane < 12 ? 0 : 6);
    row_step(w, r1, tab[r1], tab[NROWTAB + r1], xi[i], mu, tau, si);
  }
  __syncwarp();
  si.a_pr = wmin(si.a_pr);
  si.a_du = wmin(si.a_du);
  si.dphi_bar = wsum(si.dphi_bar);
  si.phi_bar = wsum(si.phi_bar);
  si.theta = wsum(si.theta);
}

// merit function pieces at the trial point (x + a dx, s + a ds) with g(trial) in GT
__device__ __forceinline__ void merit_trial(const KParams& P, const Ws& w, const double* tab, const double* drop,
                                            double alpha, double mu, double& phi_bar, double& theta, int lane) {
  const int K = P.K, MR = P.MR;
  double ph = 0.0, th = 0.0;
  for (int idx = lane; idx < MR; idx += 32) {
    const int kind = row_kind(idx, K);
    if (kind == ROW_FREE) continue;
    if (kind == ROW_EQ) {
      th += fabs(w.GT[idx] - (idx < 12 ? drop[idx] : 0.0));
      continue;
    }
    const int t = row_tab(idx);
    const double lb = tab[t], ub = tab[NROWTAB + t];
    const double s = w.S[idx] + alpha * w.DS[idx];
    th += fabs(w.GT[idx] - s);
    if (isfinite(lb)) ph -= mu * log(s - lb);
    if (isfinite(ub)) ph -= mu * log(ub - s);
  }
  phi_bar = wsum(ph);
  theta = wsum(th);
}

struct Errs {
  double dual, prim, c0, cmu, ysum, zsum, viol;
  int nzb;
};

// sigma per row and the pieces of the optimality error (oracle/ip_ref.c: assemble)
__device__ __forceinline__ void row_errors(const KParams& P, const Ws& w, const double* tab, const double* drop,
                                           double mu, Errs& e, int lane) {
  const int K = P.K, MR = P.MR;
  double dual = 0, prim = 0, c0 = 0, cmu = 0, ys = 0, zs = 0, viol = 0;
  int nb = 0;
  for (int idx = lane; idx < MR; idx += 32) {
    const int kind = row_kind(idx, K);
    if (kind == ROW_FREE) continue;
    const double g = w.G[idx], y = w.Y[idx];
    ys += fabs(y);
    if (kind == ROW_EQ) {
      const double c = fabs(g - (idx < 12 ? drop[idx] : 0.0));
      prim = fmax(prim, c);
      viol = fmax(viol, c);
      w.SIG[idx] = 0.0;
      continue;
    }
    const int t = row_tab(idx);
    const double lb = tab[t], ub = tab[NROWTAB + t];
    viol = fmax(viol, fmax(tab[2 * NROWTAB + t] - g, g - tab[3 * NROWTAB + t]));
    const double s = w.S[idx];
    double sg = 0, rs = -y;
    if (isfinite(lb)) {
      const double d = s - lb, z = w.ZL[idx];
      sg += z / d;
      rs -= z;
      c0 = fmax(c0, fabs(z * d));
      cmu = fmax(cmu, fabs(z * d - mu));
      zs += z;
      nb++;
    }
    if (isfinite(ub)) {
      const double d = ub - s, z = w.ZU[idx];
      sg += z / d;
      rs += z;
      c0 = fmax(c0, fabs(z * d));
      cmu = fmax(cmu, fabs(z * d - mu));
      zs += z;
      nb++;
    }
    prim = fmax(prim, fabs(g - s));
    dual = fmax(dual, fabs(rs));
    w.SIG[idx] = sg;
  }
  e.dual = wmax(dual); e.prim = wmax(prim); e.c0 = wmax(c0); e.cmu = wmax(cmu);
  e.ysum = wsum(ys); e.zsum = wsum(zs); e.viol = wmax(viol); e.nzb = wsumi(nb);
}

__device__ __forceinline__ double compl_at(const KParams& P, const Ws& w, const double* tab, double mu, int lane) {
  const int K = P.K, MR = P.MR;
  double cmu = 0;
  for (int idx = lane; idx < MR; idx += 32) {
    if (row_kind(idx, K) != ROW_INEQ) continue;
    const int t = row_tab(idx);
    const double lb = tab[t], ub = tab[NROWTAB + t], s = w.S[idx];
    if (isfinite(lb)) cmu = fmax(cmu, fabs(w.ZL[idx] * (s - lb) - mu));
    if (isfinite(ub)) cmu = fmax(cmu, fabs(w.ZU[idx] * (ub - s) - mu));
  }
  return wmax(cmu);
}

// yhat = sigma (g - s) - mu/(s-lb) + mu/(ub-s)
__device__ __forceinline__ void row_yhat(const KParams& P, const Ws& w, const double* tab, double mu, int lane) {
  const int K = P.K, MR = P.MR;
  for (int idx = lane; idx < MR; idx += 32) {
    if (row_kind(idx, K) != ROW_INEQ) { w.YH[idx] = 0.0; continue; }
    const int t = row_tab(idx);
    const double lb = tab[t], ub = tab[NROWTAB + t], s = w.S[idx];
    double yh = w.SIG[idx] * (w.G[idx] - s);
    if (isfinite(lb)) yh -= mu / (s - lb);
    if (isfinite(ub)) yh += mu / (ub - s);
    w.YH[idx] = yh;
  }
}

// max |grad f + J' y| (gradient of the Lagrangian w.r.t. x)
__device__ __noinline__ double dual_inf_x(const KParams& P, const Ws& w, double* sm, int lane) {
  const int N = P.N, K = P.K;
  double* nxt = sm + SM_V + V_NEXT;  // contributions of knot k-1 to (X_k, c_k)
  const SolverTables& tb = P.tab;
  if (lane < NS) nxt[lane] = lane < 12 ? w.Y[lane] : 0.0;  // initial-state rows act on X_0
  __syncwarp();
  double dmax = 0.0;
  for (int k = 0; k < K; k++) {
    const double* Jk = w.JL + (long long)k * NJ_INT;
    const double* yk = w.Y + 36 + RK * k;
    double acc[2] = {0.0, 0.0};
    for (int r = 0; r < 2; r++) {
      const int v = lane + 32 * r;
      if (v < 60) {
        const int p0 = __ldg(tb.c_ptr + v), p1 = __ldg(tb.c_ptr + v + 1);
        double a = 0.0;
        for (int p = p0; p < p1; p++) {
          const int term = __ldg(tb.c_terms + p);
          a += Jk[term & 1023] * yk[term >> 10];
        }
        acc[r] = a;
      }
    }
    // own variables 0..35 (+ what knot k-1 left for X_k, c_k); vars 36..59 go to the next knot
    if (lane < NS) acc[0] += nxt[lane];
    dmax = fmax(dmax, fabs(acc[0]));           // vars 0..31
    if (lane < 4) dmax = fmax(dmax, fabs(acc[1]));  // vars 32..35
    __syncwarp();
    if (lane >= 4 && lane < 28) nxt[lane - 4] = acc[1];  // vars 36..59 -> (X_{k+1}, c_{k+1})
    __syncwarp();
  }
  if (lane < 12) {  // terminal state: objective gradient and terminal rows
    const int r1 = lane < 6 ? 12 + lane : 24 + (lane - 6);
    const double q = w.x[12 * (N - 1) + lane];
    const double ref = lane < 6 ? P.pb.q_term_ref[lane] : P.pb.qd_term_ref[lane - 6];
    dmax = fmax(dmax, fabs(nxt[lane] + 2.0 * P.pb.QN[lane] * (q - ref) + w.Y[r1] + w.Y[r1 + 6]));
  }
  return wmax(dmax);
}

// slacks pushed inside their bounds at the current g; mu-based bound multipliers (ip_ref.c: init_slacks)
__device__ __forceinline__ void init_slacks(const KParams& P, const Ws& w, const double* tab, double mu, int lane) {
  const int K = P.K, MR = P.MR;
  const double bp = P.opt.bound_push, bf = P.opt.bound_frac;
  for (int idx = lane; idx < MR; idx += 32) {
    w.Y[idx] = 0.0; w.ZL[idx] = 0.0; w.ZU[idx] = 0.0; w.S[idx] = 0.0;
    if (row_kind(idx, K) != ROW_INEQ) continue;
    const int t = row_tab(idx);
    const double l = tab[t], u = tab[NROWTAB + t];
    double sv = w.G[idx];
    if (isfinite(l) && isfinite(u)) {
      const double pL = fmin(bp * fmax(1.0, fabs(l)), bf * (u - l));
      const double pU = fmin(bp * fmax(1.0, fabs(u)), bf * (u - l));
      sv = fmin(fmax(sv, l + pL), u - pU);
    } else if (isfinite(l)) {
      sv = fmax(sv, l + bp * fmax(1.0, fabs(l)));
    } else {
      sv = fmin(sv, u - bp * fmax(1.0, fabs(u)));
    }
    w.S[idx] = sv;
    double zl = 0.0, zu = 0.0;
    if (isfinite(l)) zl = mu / (sv - l);
    if (isfinite(u)) zu = mu / (u - sv);
    w.ZL[idx] = zl;
    w.ZU[idx] = zu;
    w.Y[idx] = zu - zl;
  }
}

// bound tables of one CTA: [lb | ub | lb_orig | ub_orig] x (36 boundary rows + 104 interior knot rows)
__device__ void build_tables(const KParams& P, double* tab) {
  const landing_problem& pb = P.pb;
  const double INF = HUGE_VAL;
  for (int t = threadIdx.x; t < NROWTAB; t += blockDim.x) {
    double lb = 0.0, ub = 0.0;
    if (t < 12) { lb = ub = 0.0; }
    else if (t < 18) { lb = pb.q_term_min[t - 12]; ub = INF; }
    else if (t < 24) { lb = -INF; ub = pb.q_term_max[t - 18]; }
    else if (t < 30) { lb = pb.qd_term_min[t - 24]; ub = INF; }
    else if (t < 36) { lb = -INF; ub = pb.qd_term_max[t - 30]; }
    else {
      const int rho = t - 36;
      if (rho < 12) { lb = ub = 0.0; }
      else if (rho < 16) { lb = 0.0; ub = pb.f_max; }
      else if (rho < 64) {
        const int j = (rho - 16) % 12;
        if (j == 0) { lb = 0.0; ub = INF; }
        else if (j == 1) { lb = -INF; ub = 0.001; }
        else if (j < 5) { lb = -INF; ub = 0.01; }
        else if (j < 8) { lb = -0.01; ub = INF; }
        else if (j < 10) { lb = -0.15; ub = 0.15; }
        else if (j == 10) { lb = -0.30; ub = 0.0; }
        else { lb = -INF; ub = pb.l_leg_max * pb.l_leg_max; }
      }
      else if (rho < 80) { lb = -INF; ub = 0.0; }
      else if (rho < 86) { lb = -INF; ub = pb.q_max[rho - 80]; }
      else if (rho < 92) { lb = pb.q_min[rho - 86]; ub = INF; }
      else if (rho < 98) { lb = -INF; ub = pb.qd_max[rho - 92]; }
      else { lb = pb.qd_min[rho - 98]; ub = INF; }
    }
    tab[2 * NROWTAB + t] = lb;
    tab[3 * NROWTAB + t] = ub;
    if (lb != ub) {  // bound_relax_factor on inequality rows
      if (isfinite(lb)) lb -= P.opt.bound_relax_factor * fmax(1.0, fabs(lb));
      if (isfinite(ub)) ub += P.opt.bound_relax_factor * fmax(1.0, fabs(ub));
    }
    tab[t] = lb;
    tab[NROWTAB + t] = ub;
  }
}

// ---------------------------------------------------------------- one scenario
__device__ void solve_one(const KParams& P, const Ws& w, double* sm, const double* tab, long long b, int lane) {
  const int N = P.N, K = P.K, nx = P.nx, MR = P.MR;
  const landing_options& opt = P.opt;
  const double kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99;
  const double gamma_theta = 1e-5, gamma_phi = 1e-5, eta_phi = 1e-8, s_theta = 1.1, s_phi = 2.3, delta_sw = 1.0;
  const double kappa_sigma = 1e10, s_max = 100.0;
  const double* drop = P.drops + 12 * b;

  // initial guess: user x0 or the reference's [Xref(:); Uref(:)] (generate_landingCtrller_IPOPT.m:199-208,336)
  if (P.x0) {
    for (int i = lane; i < nx; i += 32) w.x[i] = P.x0[b * nx + i];
  } else {
    for (int k = lane; k < N; k += 32) {
      const double t = (double)k / (double)(N - 1);
      double xr[12];
      for (int i = 0; i < 6; i++) {
        xr[i] = (k == N - 1) ? P.pb.q_term_ref[i] : __dadd_rn(drop[i], __dmul_rn(P.pb.q_term_ref[i] - drop[i], t));
        xr[6 + i] = (k == N - 1) ? P.pb.qd_term_ref[i]
                                 : __dadd_rn(drop[6 + i], __dmul_rn(P.pb.qd_term_ref[i] - drop[6 + i], t));
      }
      for (int i = 0; i < 12; i++) w.x[12 * k + i] = xr[i];
      if (k < N - 1)
        for (int l = 0; l < 4; l++)
          for (int i = 0; i < 3; i++) {
            w.x[12 * N + 24 * k + 3 * l + i] = xr[i] + P.pb.c_ref[3 * l + i];
            w.x[12 * N + 24 * k + 12 + 3 * l + i] = 0.0;
          }
    }
  }
  // rows / list slots the last knot never writes (no no-slip rows there)
  for (int i = lane; i < NJ_INT; i += 32) w.JL[(long long)(K - 1) * NJ_INT + i] = 0.0;
  for (int i = lane; i < NH_INT; i += 32) w.HL[(long long)(K - 1) * NH_INT + i] = 0.0;
  for (int i = lane; i < MR; i += 32) {
    w.G[i] = 0.0; w.GT[i] = 0.0; w.DS[i] = 0.0; w.YN[i] = 0.0; w.DZL[i] = 0.0; w.DZU[i] = 0.0;
    w.SIG[i] = 0.0; w.YH[i] = 0.0; w.Y[i] = 0.0;
  }
  __syncwarp();
  double f = eval_all<false>(P, w, w.x, w.G, lane);
  __syncwarp();
  double mu = opt.mu_init;
  init_slacks(P, w, tab, mu, lane);
  __syncwarp();

  int nfilt = 0, restarts = 0, status = LANDING_ST_MAX_ITER, it = 0;
  double theta0 = -1.0, dw_last = 0.0, viol = 0.0;
  for (it = 0; it <= opt.max_iter; it++) {
    f = eval_all<true>(P, w, w.x, w.G, lane);
    __syncwarp();
    Errs er;
    row_errors(P, w, tab, drop, mu, er, lane);
    __syncwarp();
    const double dual = fmax(er.dual, dual_inf_x(P, w, sm, lane));
    const double s_d = fmax(s_max, (er.ysum + er.zsum) / (double)(P.opt.reserved[0] + er.nzb)) / s_max;
    const double s_c = fmax(s_max, er.zsum / (double)(er.nzb > 0 ? er.nzb : 1)) / s_max;
    const double E0 = fmax(fmax(dual / s_d, er.prim), er.c0 / s_c);
    viol = er.viol;
    if (!isfinite(E0) || !isfinite(f)) { status = LANDING_ST_NAN; break; }
    if (E0 <= opt.tol && dual <= opt.dual_inf_tol && viol <= opt.constr_viol_tol && er.c0 <= opt.compl_inf_tol) {
      status = LANDING_ST_CONVERGED;
      break;
    }
    if (it == opt.max_iter) { status = LANDING_ST_MAX_ITER; break; }
    // monotone barrier update
    {
      double cmu = er.cmu;
      bool changed = false;
      for (;;) {
        const double Emu = fmax(fmax(dual / s_d, er.prim), cmu / s_c);
        if (!(Emu <= kappa_eps * mu) || mu <= opt.tol / 10.0 * 1.0000001) break;
        mu = fmax(opt.tol / 10.0, fmin(kappa_mu * mu, pow(mu, theta_mu)));
        changed = true;
        cmu = compl_at(P, w, tab, mu, lane);
      }
      if (changed) nfilt = 0;
    }
    row_yhat(P, w, tab, mu, lane);
    __syncwarp();
    const double tau = fmax(tau_min, 1.0 - mu);
    // factorise with inertia correction (IPOPT's delta_w schedule)
    double dwreg = 0.0;
    bool ok = false;
    int tries = 0;
    // while the previous iteration needed regularisation start from a third of it (ip_ref.c)
    if (dw_last > 0.0) { dwreg = dw_last / 3.0; if (dwreg < 1e-7) dwreg = 0.0; }
    for (;;) {
      if (backward_sweep(P, w, sm, tab, dwreg, lane)) { ok = true; break; }
      __syncwarp();
      if (dwreg == 0.0) dwreg = (dw_last == 0.0) ? 1e-4 : fmax(1e-20, dw_last / 3.0);
      else dwreg *= (dw_last == 0.0 && tries < 8) ? 100.0 : 8.0;
      tries++;
      if (dwreg > 1e40) break;
    }
    if (!ok) { status = LANDING_ST_FACTOR_FAIL; break; }
    dw_last = dwreg;
    StepInfo si;
    forward_sweep(P, w, sm, tab, drop, mu, tau, si, lane);
    __syncwarp();
    // filter line search
    const double theta = si.theta, phi = f + si.phi_bar;
    if (theta0 < 0) theta0 = theta;
    const double theta_max = 1e4 * fmax(1.0, theta0), theta_min = 1e-4 * fmax(1.0, theta0);
    double dphi = si.dphi_bar;
    {
      double d = 0.0;
      if (lane < 12) {
        const double q = w.x[12 * (N - 1) + lane];
        const double ref = lane < 6 ? P.pb.q_term_ref[lane] : P.pb.qd_term_ref[lane - 6];
        d = 2.0 * P.pb.QN[lane] * (q - ref) * w.dx[12 * (N - 1) + lane];
      }
      dphi += wsum(d);
    }
    double alpha = si.a_pr, ft = f;
    bool accepted = false, ftype = false;
    int ls = 0;
    while (alpha > 1e-12 * si.a_pr && ls < 40) {
      for (int i = lane; i < nx; i += 32) w.xt[i] = w.x[i] + alpha * w.dx[i];
      __syncwarp();
      ft = eval_all<false>(P, w, w.xt, w.GT, lane);
      __syncwarp();
      double phb, tht;
      merit_trial(P, w, tab, drop, alpha, mu, phb, tht, lane);
      const double pht = ft + phb;
      bool filt_ok = true;
      for (int i = lane; i < nfilt; i += 32)
        if (tht >= w.FT[i] && pht >= w.FP[i]) filt_ok = false;
      filt_ok = __all_sync(FULL, filt_ok);
      if (isfinite(pht) && isfinite(tht) && tht <= theta_max && filt_ok) {
        const bool sw = (theta <= theta_min) && (dphi < 0) && (alpha * pow(-dphi, s_phi) > delta_sw * pow(theta, s_theta));
        if (sw) {
          if (pht <= phi + eta_phi * alpha * dphi) { accepted = true; ftype = true; }
        } else if (tht <= (1.0 - gamma_theta) * theta || pht <= phi - gamma_phi * theta) {
          accepted = true;
        }
      }
      if (accepted) break;
      alpha *= 0.5;
      ls++;
    }
    if (!accepted) {
      if (restarts < 20) {  // re-centre: slacks back inside their bounds, multipliers reset, mu = mu_init
        restarts++;
        mu = opt.mu_init;
        init_slacks(P, w, tab, mu, lane);
        nfilt = 0;
        theta0 = -1.0;
        __syncwarp();
        continue;
      }
      status = LANDING_ST_LINESEARCH_FAIL;
      break;
    }
    if (!ftype) {
      if (nfilt == MAXFILTER) {
        double a = 0, c = 0;
        for (int base = 0; base < MAXFILTER; base += 32) {
          const int i = base + lane;
          if (i + 1 < MAXFILTER) { a = w.FT[i + 1]; c = w.FP[i + 1]; }
          __syncwarp();
          if (i + 1 < MAXFILTER) { w.FT[i] = a; w.FP[i] = c; }
          __syncwarp();
        }
        nfilt--;
      }
      if (lane == 0) { w.FT[nfilt] = (1.0 - gamma_theta) * theta; w.FP[nfilt] = phi - gamma_phi * theta; }
      nfilt++;
    }
    // accept the trial point
    for (int i = lane; i < nx; i += 32) w.x[i] = w.xt[i];
    f = ft;
    for (int idx = lane; idx < MR; idx += 32) {
      const int kind = row_kind(idx, K);
      if (kind == ROW_FREE) continue;
      w.G[idx] = w.GT[idx];
      const double y = w.Y[idx];
      w.Y[idx] = y + alpha * (w.YN[idx] - y);
      if (kind == ROW_EQ) continue;
      const int t = row_tab(idx);
      const double lb = tab[t], ub = tab[NROWTAB + t];
      const double s = w.S[idx] + alpha * w.DS[idx];
      w.S[idx] = s;
      if (isfinite(lb)) {
        const double d = s - lb;
        double z = w.ZL[idx] + si.a_du * w.DZL[idx];
        z = fmax(fmin(z, kappa_sigma * mu / d), mu / (kappa_sigma * d));
        w.ZL[idx] = z;
      }
      if (isfinite(ub)) {
        const double d = ub - s;
        double z = w.ZU[idx] + si.a_du * w.DZU[idx];
        z = fmax(fmin(z, kappa_sigma * mu / d), mu / (kappa_sigma * d));
        w.ZU[idx] = z;
      }
    }
    __syncwarp();
  }
  // results (AoS, CasADi order)
  for (int i = lane; i < nx; i += 32) P.x_star[b * nx + i] = w.x[i];
  if (P.lam_g) {
    const long long m = 104LL * N - 92;
    for (int idx = lane; idx < MR; idx += 32) {
      if (idx < 36 + RK * (K - 1)) { P.lam_g[b * m + idx] = w.Y[idx]; continue; }
      // last knot: interior numbering -> the 80-row layout of the generated functions
      const int rho = idx - 36 - RK * (K - 1);
      if (is_noslip(rho)) continue;
      int r = rho;
      if (rho >= 64) r = rho - 24;
      else if (rho >= 16) { const int l = (rho - 16) / 12, j = (rho - 16) % 12; r = 16 + 6 * l + (j < 2 ? j : j - 6); }
      P.lam_g[b * m + 36 + RK * (K - 1) + r] = w.Y[idx];
    }
  }
  if (lane == 0) {
    P.f_star[b] = f;
    P.status[b] = status;
    P.iters[b] = it;
    if (P.viol) P.viol[b] = viol;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(WARPS * 32, 1) k_solve(KParams P) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* tab = smem + SM_TAB;
  build_tables(P, tab);
  __syncthreads();
  double* sm = smem + warp * SM_WARP;
  const Ws w = carve(P.scratch + (long long)(blockIdx.x * WARPS + warp) * P.slot, P.N);
  for (;;) {
    long long b = 0;
    if (lane == 0) b = atomicAdd(P.counter, 1);
    b = __shfl_sync(FULL, b, 0);
    if (b >= P.B) break;
    solve_one(P, w, sm, tab, b, lane);
  }
}

// ---------------------------------------------------------------- host: tables
struct HostTables {
  std::vector<int> all;
  size_t o_jl, o_hl, o_ge, o_gt, o_ht, o_tptr, o_tij, o_tterms, o_qptr, o_qterms, o_rptr, o_rterms, o_cptr, o_cterms;
  int g_n, t_n;
};

int sidx_host(int v) { return v < 36 ? v : (v >= 48 ? v - 12 : -1); }
int dyn_state_host(int rho) { return rho < 6 ? rho : (rho < 9 ? rho + 3 : rho - 3); }
int rowmap_last_host(int r) {
  if (r < 16) return r;
  if (r < 40) { const int l = (r - 16) / 6, j = (r - 16) % 6; return 16 + 12 * l + (j < 2 ? j : j + 6); }
  return r + 24;
}

HostTables build_tables_host() {
  Knot z{};
  z.h = 0.03; z.mu = 1; z.mass = 1;
  for (int i = 0; i < 3; i++) { z.Ib[i] = 1; z.Ibinv[i] = 1; }
  PatternSink pi, pl;
  NoLam nl;
  knot_eval<false, false, true, true>(z, pi, nl);
  knot_eval<true, false, true, true>(z, pl, nl);
  HostTables T;
  auto push = [&](const std::vector<int>& v) { size_t o = T.all.size(); T.all.insert(T.all.end(), v.begin(), v.end()); return o; };
  // last-knot emission index -> interior emission index
  std::vector<int> jl(NJ_LAST), hl(NH_LAST);
  for (int e = 0; e < NJ_LAST; e++) {
    const int row = rowmap_last_host(pl.jac[e].first), var = pl.jac[e].second;
    int f = -1;
    for (int q = 0; q < NJ_INT; q++) if (pi.jac[q].first == row && pi.jac[q].second == var) f = q;
    jl[e] = f;
  }
  for (int e = 0; e < NH_LAST; e++) {
    int f = -1;
    for (int q = 0; q < NH_INT; q++) if (pi.hes[q] == pl.hes[e]) f = q;
    hl[e] = f;
  }
  T.o_jl = push(jl);
  T.o_hl = push(hl);
  // dynamics entries -> G
  std::vector<int> ge, gt;
  for (int e = 0; e < NJ_INT; e++)
    if (pi.jac[e].first < 12 && pi.jac[e].second < 36) {
      ge.push_back(e);
      gt.push_back(dyn_state_host(pi.jac[e].first) * 36 + pi.jac[e].second);
    }
  T.g_n = (int)ge.size();
  T.o_ge = push(ge);
  T.o_gt = push(gt);
  // Hessian targets
  std::vector<int> ht(NH_INT);
  for (int e = 0; e < NH_INT; e++) ht[e] = sidx_host(pi.hes[e].first) * NW + sidx_host(pi.hes[e].second);
  T.o_ht = push(ht);
  // condensing targets (i <= j) <- terms (rho, ea, eb)
  std::map<int, std::vector<int>> tg;
  std::vector<std::vector<int>> rows(RK);
  for (int e = 0; e < NJ_INT; e++) rows[pi.jac[e].first].push_back(e);
  for (int rho = 12; rho < RK; rho++)
    for (size_t a = 0; a < rows[rho].size(); a++)
      for (size_t b2 = a; b2 < rows[rho].size(); b2++) {
        const int ea = rows[rho][a], eb = rows[rho][b2];
        int i = sidx_host(pi.jac[ea].second), j = sidx_host(pi.jac[eb].second);
        if (i > j) std::swap(i, j);
        tg[i * NW + j].push_back((rho << 20) | (ea << 10) | eb);
        if (i == j && ea != eb) tg[i * NW + j].push_back((rho << 20) | (eb << 10) | ea);
      }
  // deal targets to lanes longest-first for balance
  std::vector<std::pair<int, int>> order;
  for (auto& kv : tg) order.push_back({-(int)kv.second.size(), kv.first});
  std::sort(order.begin(), order.end());
  std::vector<int> tptr{0}, tij, tterms;
  for (auto& o : order) {
    tij.push_back(o.second);
    for (int t : tg[o.second]) tterms.push_back(t);
    tptr.push_back((int)tterms.size());
  }
  T.t_n = (int)tij.size();
  T.o_tptr = push(tptr);
  T.o_tij = push(tij);
  T.o_tterms = push(tterms);
  // stage gradient: per stage variable
  std::vector<int> qptr{0}, qterms;
  for (int v = 0; v < NW; v++) {
    for (int e = 0; e < NJ_INT; e++)
      if (pi.jac[e].first >= 12 && sidx_host(pi.jac[e].second) == v) qterms.push_back((pi.jac[e].first << 10) | e);
    qptr.push_back((int)qterms.size());
  }
  T.o_qptr = push(qptr);
  T.o_qterms = push(qterms);
  // row schedule: per row, (stage var, e)
  std::vector<int> rptr(RK + 1, 0), rterms;
  for (int rho = 0; rho < RK; rho++) {
    if (rho >= 12)
      for (int e : rows[rho]) rterms.push_back((sidx_host(pi.jac[e].second) << 10) | e);
    rptr[rho + 1] = (int)rterms.size();
  }
  T.o_rptr = push(rptr);
  T.o_rterms = push(rterms);
  // column schedule: per local variable (60), (rho, e)
  std::vector<int> cptr{0}, cterms;
  for (int v = 0; v < 60; v++) {
    for (int e = 0; e < NJ_INT; e++)
      if (pi.jac[e].second == v) cterms.push_back((pi.jac[e].first << 10) | e);
    cptr.push_back((int)cterms.size());
  }
  T.o_cptr = push(cptr);
  T.o_cterms = push(cterms);
  return T;
}

}  // namespace

void solver_free(SolverWorkspace& ws) {
  if (ws.scratch) cudaFree(ws.scratch);
  if (ws.io) cudaFree(ws.io);
  if (ws.counter) cudaFree(ws.counter);
  if (ws.tab.dev) cudaFree(ws.tab.dev);
  ws = SolverWorkspace{};
}

#define CUS(call)                                                            \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      *err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
      return LANDING_ERR_CUDA;                                               \
    }                                                                        \
  } while (0)

int solver_run(SolverWorkspace& ws, const DevicePlan& pl, long long B, int memspace,
               const landing_problem& pb, const landing_options& opt, const landing_solve_io& io,
               cudaStream_t st, int* launches, std::string* err) {
  const int N = pl.N;
  if (N - 1 > 1023 / 1) { *err = "landing_solve_batch: N too large"; return LANDING_ERR_ARG; }
  if (!io.x_star || !io.f_star || !io.status || !io.iters) {
    *err = "landing_solve_batch: x_star, f_star, status and iters are required";
    return LANDING_ERR_ARG;
  }
  if (!ws.tab.dev) {
    HostTables T = build_tables_host();
    CUS(cudaMalloc(&ws.tab.dev, sizeof(int) * T.all.size()));
    CUS(cudaMemcpy(ws.tab.dev, T.all.data(), sizeof(int) * T.all.size(), cudaMemcpyHostToDevice));
    const int* d = ws.tab.dev;
    ws.tab.jl_last = d + T.o_jl; ws.tab.hl_last = d + T.o_hl;
    ws.tab.g_e = d + T.o_ge; ws.tab.g_t = d + T.o_gt; ws.tab.g_n = T.g_n;
    ws.tab.h_t = d + T.o_ht;
    ws.tab.t_ptr = d + T.o_tptr; ws.tab.t_ij = d + T.o_tij; ws.tab.t_terms = d + T.o_tterms; ws.tab.t_n = T.t_n;
    ws.tab.q_ptr = d + T.o_qptr; ws.tab.q_terms = d + T.o_qterms;
    ws.tab.r_ptr = d + T.o_rptr; ws.tab.r_terms = d + T.o_rterms;
    ws.tab.c_ptr = d + T.o_cptr; ws.tab.c_terms = d + T.o_cterms;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&ws.n_sm, cudaDevAttrMultiProcessorCount, dev);
    CUS(cudaMalloc(&ws.counter, sizeof(int)));
    CUS(cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM_TOTAL * sizeof(double))));
  }
  const long long slot = slot_doubles(N);
  const long long nslots = (long long)ws.n_sm * WARPS;
  const size_t need = sizeof(double) * slot * nslots;
  if (need > ws.scratch_bytes) {
    if (ws.scratch) cudaFree(ws.scratch);
    ws.scratch = nullptr;
    ws.scratch_bytes = 0;
    CUS(cudaMalloc(&ws.scratch, need));
    ws.scratch_bytes = need;
  }
  const long long nx = pl.nx, m = pl.m;
  KParams P{};
  P.N = N; P.K = N - 1; P.nx = pl.nx; P.MR = 36 + RK * (N - 1);
  P.B = B; P.pb = pb; P.opt = opt;
  P.opt.reserved[0] = pl.m;  // m, for the dual scaling s_d
  P.counter = ws.counter; P.scratch = ws.scratch; P.slot = slot; P.tab = ws.tab;
  if (memspace == LANDING_HOST) {
    size_t bytes = sizeof(double) * B * (12 + nx + 2 + (io.x0 ? nx : 0) + (io.lam_g ? m : 0)) + sizeof(int) * 2 * B + 64;
    if (bytes > ws.io_bytes) {
      if (ws.io) cudaFree(ws.io);
      ws.io = nullptr;
      ws.io_bytes = 0;
      CUS(cudaMalloc(&ws.io, bytes));
      ws.io_bytes = bytes;
    }
    double* p = (double*)ws.io;
    double* d_drops = p; p += 12 * B;
    P.x_star = p; p += nx * B;
    P.f_star = p; p += B;
    P.viol = p; p += B;
    double* d_x0 = nullptr;
    if (io.x0) { d_x0 = p; p += nx * B; }
    if (io.lam_g) { P.lam_g = p; p += m * B; }
    P.status = (int*)p;
    P.iters = P.status + B;
    CUS(cudaMemcpyAsync(d_drops, io.drops, sizeof(double) * 12 * B, cudaMemcpyHostToDevice, st));
    if (io.x0) CUS(cudaMemcpyAsync(d_x0, io.x0, sizeof(double) * nx * B, cudaMemcpyHostToDevice, st));
    P.drops = d_drops;
    P.x0 = d_x0;
  } else {
    P.drops = io.drops; P.x0 = io.x0; P.x_star = io.x_star; P.f_star = io.f_star; P.lam_g = io.lam_g;
    P.viol = io.viol; P.status = io.status; P.iters = io.iters;
  }
  CUS(cudaMemsetAsync(ws.counter, 0, sizeof(int), st));
  if (P.lam_g) CUS(cudaMemsetAsync(P.lam_g, 0, sizeof(double) * m * B, st));
  const int grid = (int)std::min<long long>(ws.n_sm, (B + WARPS - 1) / WARPS);
  k_solve<<<grid, WARPS * 32, SM_TOTAL * sizeof(double), st>>>(P);
  *launches += 1;
  CUS(cudaGetLastError());
  if (memspace == LANDING_HOST) {
    CUS(cudaMemcpyAsync(io.x_star, P.x_star, sizeof(double) * nx * B, cudaMemcpyDeviceToHost, st));
    CUS(cudaMemcpyAsync(io.f_star, P.f_star, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    if (io.viol) CUS(cudaMemcpyAsync(io.viol, P.viol, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    if (io.lam_g) CUS(cudaMemcpyAsync(io.lam_g, P.lam_g, sizeof(double) * m * B, cudaMemcpyDeviceToHost, st));
    CUS(cudaMemcpyAsync(io.status, P.status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    CUS(cudaMemcpyAsync(io.iters, P.iters, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    CUS(cudaStreamSynchronize(st));
  }
  return LANDING_OK;
}

}  // namespace srb
