// solver_dev.cuh -- device side of the batched interior-point solver (included by solver.cu only).
//
// ONE CTA (SRB_NT threads, 256 in the shipped build) = ONE SCENARIO for the whole solve; SRB_CTAS CTAs (2) are resident
// per SM and pull scenario ids from an atomic queue (ordered longest-expected-first by k_order).  Phases of one
// interior-point iteration:
//   (parallel over knots / rows, all threads)
//     eval        J lists on the first half of the CTA, H lists on the second half, one knot per thread
//     row passes  sigma, yhat, optimality-error pieces, grad of the Lagrangian, step recovery,
//                 fraction-to-the-boundary, merit function  -> block reductions
//     condensing  everything of a stage that does not depend on P_{k+1} (sweeps.cuh: condense_all)
//   (sequential over stages, all threads cooperate on one stage)
//     backward    condensed stage data -> 49x48 stage matrix as FP64 tensor-core tiles in registers (lower, elimination
//                 order) + W'PW terms -> blocked partial Cholesky (24 pivots, panels of 8) -> Yt, P_k, p_k
//     forward     u = -L^-T (Y xi + yv), xi+ = G [xi;u] + r, costates
// All reductions end with the identical value in every thread, so control flow is block-uniform.
#pragma once

namespace srb {
namespace {

constexpr unsigned FULL = 0xffffffffu;
#define TID ((int)threadIdx.x)
// Threads per CTA (= per scenario) and CTAs per SM are build-time choices (make EXTRA="-DSRB_NT=.. -DSRB_CTAS=.."):
// the solve of ONE scenario is a chain of short dependent phases (24 pivots per stage, K stages per sweep), so what fills
// an SM is the number of INDEPENDENT scenarios resident on it, not the number of threads of one scenario.  Shared memory
// per CTA (below) and the register file bound the choice: 256 x 2, 128 x 4 or 64 x 5 (threads x CTAs per SM).
#ifndef SRB_NT
#define SRB_NT 256
#endif
constexpr int NT = SRB_NT;       // threads per CTA (= per scenario)
constexpr int NWARP = NT / 32;
static_assert(NT % 32 == 0 && NT >= 64 && NT <= 384, "threads per scenario");
#ifndef SRB_CTAS
#define SRB_CTAS (SRB_NT == 256 ? 2 : (SRB_NT == 128 ? 4 : 5))
#endif
constexpr int CTAS_PER_SM = SRB_CTAS;
constexpr int NS = 24;           // stage state / control size
constexpr int NW = 48;           // stage variables X(12) c(12) f(12) c+(12)
constexpr int LDC = 52, LDMS = 28, LDP = 28, LDG = 37;  // leading dimensions: 4 or 12 (mod 16) doubles keeps the DMMA fragment loads conflict free
constexpr int MAXFILTER = 64;
constexpr int RK = 104;          // rows per knot (interior numbering)
constexpr int NROWTAB = 36 + RK;
constexpr int NJ_PAD = 388, NH_PAD = 192;  // strides of the per-knot entry lists in the scratch (16-byte chunks; padding = 0)

// condensed stage data: 278 target sums | 48 stage-gradient entries | 137 G values | 12 dynamics defects
constexpr int CT_Q = 280, CT_G = 328, CT_R = 468, CT_STRIDE = 480;

// ---- shared memory carve-up (doubles) per CTA.  The condensed stage matrix itself never sits in shared memory: the
// accumulator tiles of the factorisation are initialised straight from the condensed sums through a host-made inverse
// table (SolverTables::tinit), and the panel buffer shares its region with T (dead once the tiles are formed).
constexpr int SM_P = 0;                        // 24 x 28   P_{k+1} (full, symmetric)
constexpr int SM_W = SM_P + NS * LDP;          // 24 x 52   W = [G^ ; E]: next stage state = W (stage variables)
constexpr int SM_T = SM_W + NS * LDC;          // 24 x 52   T = P W
constexpr int SM_MS = SM_T;                    // 50 x 28   panel buffer: L | Yt | yv of the stage being factored (aliases T)
constexpr int SM_SWEEP_END = SM_MS + 50 * LDMS;  // (49 rows + a zero row)
constexpr int SM_SWEEP = SM_SWEEP_END - SM_P;  // doubles of the sweep regions (ring buffers of the row passes alias them)
// stage list buffers (double-buffered, filled by cp.async): J | H | sigma | yhat | dynamics defects
constexpr int LB_J = 0, LB_H = 388, LB_SIG = 580, LB_YH = 684, LB_GD = 788, LB_SIZE = 800;
constexpr int SM_LB0 = SM_SWEEP_END;
constexpr int LB_REGION = 2 * CT_STRIDE;       // two condensed-stage buffers (backward) / one list buffer (pre-pass)
constexpr int SM_V = SM_LB0 + LB_REGION;       // vectors
constexpr int V_Q = 0, V_Z = 48, V_R = 96, V_T = 108, V_YV = 132, V_PN = 156, V_XI = 180, V_U = 204, V_END = 228;  // V_Z: 24 zeros
constexpr int SM_RED = SM_V + V_END;           // block-reduction scratch (up to 12 warps) x 8
constexpr int SM_TAB = SM_RED + 16 * 8;        // lb[140] ub[140] lbo[140] ubo[140]
constexpr int TBL_INTS = 2688;                 // index tables of the sweeps (SolverTables::sm_src)
// the index tables live in shared memory (one copy per CTA) when the CTA is large enough to afford it, else they are read
// from global memory through the L1 (one copy per SM in effect)
#ifndef SRB_TBL_SMEM
#define SRB_TBL_SMEM (SRB_NT >= 128)
#endif
constexpr bool TBL_SMEM = SRB_TBL_SMEM;
constexpr int SM_TBL = SM_TAB + 4 * NROWTAB;
constexpr int SM_TOTAL = SM_TBL + (TBL_SMEM ? TBL_INTS / 2 : 0);
// row-kind tables (bytes, GLOBAL memory, built once per launch by k_kinds): kind of knot-local row rho for every knot
// class | class of every knot | boundary rows
constexpr int KT_CLASSES = 32, KT_CLS = KT_CLASSES * RK, KT_MAXK = 1024, KT_BND = KT_CLS + KT_MAXK, KT_BYTES = KT_BND + 40;
static_assert(SM_W % 2 == 0 && SM_T % 2 == 0 && SM_LB0 % 2 == 0 && SM_V % 2 == 0, "16-byte alignment of the regions");
static_assert(CTAS_PER_SM * (SM_TOTAL * 8 + 2048 + 1024) <= 232448, "shared memory budget (227 KB per SM)");

struct KParams {
  int N, K, nx, MR;
  long long B;
  landing_problem pb;
  landing_options opt;
  const double* drops;
  const double* x0;
  const double* dtv;  // knot spacings dt[0..K-1] (device)
  const unsigned char* csm;  // fixed-schedule formulation: contact bit mask of every knot (device), else nullptr
  const double* zeros;       // 16 zeros in global memory (c+ of the last knot)
  double *x_star, *f_star, *lam_g, *viol;
  int *status, *iters;
  int run_qx;  // any QX != 0
  int* counter;
  double* scratch;
  long long slot;  // doubles per CTA slot
  SolverTables tab;
  unsigned long long* prof;  // optional per-phase cycle counters (LANDING_PROF=1), else nullptr
  const int* order;          // queue position -> scenario id (nullptr: input order)
  unsigned char* kt;         // row-kind tables (KT_BYTES, global memory; filled by k_kinds before k_solve)
};

// phase ids of the optional cycle profile
enum { PH_EVAL = 0, PH_ERR, PH_DUAL, PH_MU, PH_BACK, PH_FWD, PH_ROWS, PH_LS, PH_ACCEPT, PH_NBACK, PH_NITER,
       PH_B_WAIT, PH_B_P1, PH_B_P2, PH_B_P3, PH_B_P4, PH_B_CHOL, PH_B_P6, PH_B_STAGES, PH_C_DIAG, PH_C_TRAIL,
       PH_F_WAIT, PH_F_RHS, PH_F_SOLVE, PH_F_NEXT, PH_CAL,
       PH_X0, PH_X1, PH_X2, PH_X3, PH_X4, PH_X5, PH_X6, PH_X7, PH_COUNT };
struct Prof {
  unsigned long long* c;
  long long t;
  int who = 0;  // the thread that measures
#ifndef SRB_PROF  // build with make EXTRA=-DSRB_PROF for the per-phase cycle profile (costs ~4 %)
  __device__ __forceinline__ void start() {}
  __device__ __forceinline__ void lap(int) {}
  __device__ __forceinline__ void count(int) {}
#else
  __device__ __forceinline__ void start() { if (c && TID == who) t = clock64(); }
  __device__ __forceinline__ void lap(int ph) {
    if (c && TID == who) { const long long n = clock64(); atomicAdd(c + ph, (unsigned long long)(n - t)); t = n; }
  }
  __device__ __forceinline__ void count(int ph) { if (c && TID == 0) atomicAdd(c + ph, 1ull); }
#endif
};

struct Ws {  // pointers into one CTA's scratch slot
  double *x, *xt, *dx;
  double *S, *Y, *ZL, *ZU, *G, *GT, *DS, *YN, *DZL, *DZU, *SIG, *YH;
  double *JL, *HL;
  double *CT;  // condensed stage data [K][CT_STRIDE], see sweeps.cuh:condense_all
  double *FY, *rf, *yvf, *PX, *PV, *L0;  // FY: per stage [48][24] = L (rows 0-23) | Yt (rows 24-47)
  double *FT, *FP;
};

__host__ __device__ inline long long slot_doubles(int N) {
  const long long K = N - 1, nx = 36LL * N - 24, MR = 36 + RK * K;
  const long long n = 3 * nx + 12 * MR + K * (NJ_PAD + NH_PAD + CT_STRIDE) + K * (1152 + 12 + 24) + (K + 1) * (288 + 24) + 144 +
                      2 * MAXFILTER;
  return (n + 31) / 32 * 32;
}

__device__ inline Ws carve(double* base, int N) {
  const long long K = N - 1, nx = 36LL * N - 24, MR = 36 + RK * K;
  Ws w;
  double* p = base;  // every array starts 16-byte aligned (all sizes are even)
  w.x = p; p += nx; w.xt = p; p += nx; w.dx = p; p += nx;
  w.S = p; p += MR; w.Y = p; p += MR; w.ZL = p; p += MR; w.ZU = p; p += MR; w.G = p; p += MR;
  w.GT = p; p += MR; w.DS = p; p += MR; w.YN = p; p += MR; w.DZL = p; p += MR; w.DZU = p; p += MR;
  w.SIG = p; p += MR; w.YH = p; p += MR;
  w.JL = p; p += K * NJ_PAD; w.HL = p; p += K * NH_PAD;
  w.CT = p; p += K * CT_STRIDE;
  w.FY = p; p += K * 1152; w.rf = p; p += K * 12;
  w.yvf = p; p += K * 24; w.PX = p; p += (K + 1) * 288; w.PV = p; p += (K + 1) * 24;
  w.L0 = p; p += 144; w.FT = p; p += MAXFILTER; w.FP = p; p += MAXFILTER;
  return w;
}

// ---------------------------------------------------------------- block reductions
enum { R_SUM = 0, R_MAX = 1, R_MIN = 2 };
__device__ __forceinline__ double rcomb(int op, double a, double b) {
  return op == R_SUM ? a + b : (op == R_MAX ? fmax(a, b) : fmin(a, b));
}
// Reduces one quantity over the CTA (shuffles, then the eight warp results serially); every thread returns with
// the identical result.
template <int OP>
__device__ __forceinline__ double block_reduce1(double* red, double v) {
  const int lane = TID & 31, warp = TID >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v = rcomb(OP, v, __shfl_xor_sync(FULL, v, o));
  if (lane == 0) red[warp * 8] = v;
  __syncthreads();
  double a = red[0];
#pragma unroll
  for (int w2 = 1; w2 < NWARP; w2++) a = rcomb(OP, a, red[w2 * 8]);
  __syncthreads();
  return a;
}
// Reduces sizeof...(OPS) <= 8 quantities over the CTA: the values are transposed through the (idle) sweep regions of
// shared memory and warp q (mod NWARP) reduces quantity q, so the code is one short loop instead of NQ unrolled shuffle
// trees (1.2k instructions for eight quantities; the kernel is instruction-fetch bound).
// Every thread returns with the identical results in v[].  `red` = smem + SM_RED.
template <int... OPS>
__device__ __forceinline__ void block_reduce(double* red, double (&v)[sizeof...(OPS)]) {
  constexpr int NQ = sizeof...(OPS);
  static_assert(NQ <= 8 && NQ * NT <= SM_SWEEP, "the transpose buffer aliases the sweep regions");
  double* buf = red - SM_RED + SM_P;
  const int tid = TID, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int q = 0; q < NQ; q++) buf[q * NT + tid] = v[q];
  __syncthreads();
  constexpr int ops[NQ] = {OPS...};
#pragma unroll 1
  for (int q = warp; q < NQ; q += NWARP) {
    int op = ops[0];
#pragma unroll
    for (int j = 1; j < NQ; j++) op = (q == j) ? ops[j] : op;
    const double* b = buf + q * NT + lane;
    double a = b[0];
#pragma unroll
    for (int j = 1; j < NT / 32; j++) a = rcomb(op, a, b[32 * j]);
#pragma unroll 1
    for (int o = 16; o; o >>= 1) a = rcomb(op, a, __shfl_xor_sync(FULL, a, o));
    if (lane == 0) red[q] = a;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; q++) v[q] = red[q];
  __syncthreads();
}
// one copy of each in the kernel (code size: the kernel is instruction-fetch bound)
__device__ __noinline__ double bsum(double* red, double x) {
  return block_reduce1<R_SUM>(red, x);
}
__device__ __noinline__ double bmax(double* red, double x) {
  return block_reduce1<R_MAX>(red, x);
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

// row kinds: hard equality (initial state, dynamics: the constraints of the Riccati recursion) | inequality with a slack
// | free (not a constraint of this formulation / knot) | dual-regularised equality (fixed-schedule rows, ip_ref.h)
enum { ROW_EQ = 0, ROW_INEQ = 1, ROW_FREE = 2, ROW_EQS = 3 };
// kind of knot-local row rho of a knot of class cls = contact bits (15 in formulation 0) + 16 (last knot)
__device__ __forceinline__ int row_kind_of(bool sched, int cls, int rho) {
  const bool last = cls >= 16;
  const unsigned cs = cls & 15;
  if (rho < 12) return ROW_EQ;
  if (!sched) return (last && is_noslip(rho)) ? ROW_FREE : ROW_INEQ;
  if (rho < 16) return (cs >> (rho - 12) & 1) ? ROW_INEQ : ROW_EQS;  // f_z in [0, cs f_max]: f_z = 0 in flight
  if (rho < 64) {
    const int l = (rho - 16) / 12, j = (rho - 16) - 12 * l;
    const bool on = cs >> l & 1;
    if (j == 0) return on ? ROW_EQS : ROW_FREE;                       // cs c_z = 0
    if (j == 1) return ROW_FREE;
    if (j < 5) return (on && !last) ? ROW_EQS : ROW_FREE;             // cs (c+ - c) = 0
    if (j < 8) return ROW_FREE;
    return ROW_INEQ;                                                  // kinematic box, leg length
  }
  return ROW_INEQ;
}
// The row passes look the kind of a row up in a byte table (one division by 104 and two byte loads, L1 hits) instead of
// re-deriving it: the derivation was ~100 instructions inlined at a dozen sites, and the kernel is instruction-fetch
// bound.  The table depends only on the problem, so one copy in global memory serves all CTAs.
__global__ void k_kinds(const KParams P) {
  unsigned char* kt = P.kt;
  const bool sched = P.pb.formulation == 1;
  for (int i = TID; i < KT_CLASSES * RK; i += blockDim.x) kt[i] = (unsigned char)row_kind_of(sched, i / RK, i % RK);
  for (int k = TID; k < P.K; k += blockDim.x) kt[KT_CLS + k] = (unsigned char)((sched ? __ldg(P.csm + k) & 15 : 15) + (k == P.K - 1 ? 16 : 0));
  if (TID < 36) kt[KT_BND + TID] = (unsigned char)(TID < 12 ? ROW_EQ : (sched ? ROW_FREE : ROW_INEQ));
}
__device__ __forceinline__ int row_kind(const unsigned char* kt, int idx) {
  if (idx < 36) return kt[KT_BND + idx];
  const int k = (idx - 36) / RK, rho = (idx - 36) - k * RK;
  return kt[kt[KT_CLS + k] * RK + rho];
}
__device__ __forceinline__ int row_tab(int idx) { return idx < 36 ? idx : 36 + (idx - 36) % RK; }

template <bool LAST> struct LamY {
  const double* y;  // multipliers of this knot's rows (interior numbering)
  __device__ __forceinline__ double operator()(int r) const { return y[rowmap<LAST>(r)]; }
};

__device__ __forceinline__ void load_knot(const KParams& P, const double* x, const double* zeros, int k, KnotRef& kn) {
  const int N = P.N;
  kn.X = x + 12 * k;
  kn.Xn = x + 12 * (k + 1);
  kn.c = x + 12 * N + 24 * k;
  kn.f = kn.c + 12;
  kn.cn = (k == N - 2) ? zeros : kn.c + 24;  // (the last knot has no c+: its no-slip rows are unused)
  kn.h = __ldg(P.dtv + k);
  if (P.csm) {
    const unsigned cs = __ldg(P.csm + k);
#pragma unroll
    for (int l = 0; l < 4; l++) kn.csv[l] = (double)(cs >> l & 1);
  }
  kn.mu = P.pb.mu;
  kn.mass = P.pb.mass;
#pragma unroll
  for (int i = 0; i < 3; i++) { kn.Ib[i] = P.pb.Ib[i]; kn.Ibinv[i] = P.pb.Ib_inv[i]; }
}

// The LAST knot is evaluated with the interior template too (c+ := 0): its no-slip rows and their entries are
// never used -- those rows are ROW_FREE (sigma = y = 0 for the whole solve, skipped by every row pass) -- and a
// second instantiation would run serially in the same warp (divergence) and double the code the warp streams.
// (Splitting one knot over several threads behind a compile-time filter, as the batched evaluation kernels do, was
// measured here at +1 % for 76 kB more code and is not kept.)
struct JSinkP {  // g rows + Jacobian entry list
  double *gp, *jl;
  __device__ __forceinline__ void g(int r, double v) { gp[r] = v; }
  __device__ __forceinline__ void j(int e, int, int, double v) { jl[e] = v; }
  __device__ __forceinline__ void h(int, int, int, double) {}
};
struct HSinkP {  // Hessian entry list
  double* hl;
  __device__ __forceinline__ void g(int, double) {}
  __device__ __forceinline__ void j(int, int, int, double) {}
  __device__ __forceinline__ void h(int e, int, int, double v) { hl[e] = v; }
};
struct GSinkP {
  double* gp;
  __device__ __forceinline__ void g(int r, double v) { gp[r] = v; }
  __device__ __forceinline__ void j(int, int, int, double) {}
  __device__ __forceinline__ void h(int, int, int, double) {}
};
__device__ __noinline__ void eval_knot_j(const KParams& P, const Ws& w, const double* x, double* gout, int k0, int kstep) {
  for (int k = k0; k < P.K; k += kstep) {
    KnotRef kn;
    load_knot(P, x, P.zeros, k, kn);
    NoLam nl;
    JSinkP s{gout + 36 + RK * k, w.JL + (long long)k * NJ_PAD};
    if (P.pb.formulation == 1) knot_eval<false, true, true, false, JSinkP, NoLam, true, KnotRef>(kn, s, nl);
    else knot_eval<false, true, true, false, JSinkP, NoLam, false, KnotRef>(kn, s, nl);
  }
}
__device__ __noinline__ void eval_knot_h(const KParams& P, const Ws& w, const double* x, int k0, int kstep) {
  for (int k = k0; k < P.K; k += kstep) {
    KnotRef kn;
    load_knot(P, x, P.zeros, k, kn);
    HSinkP s{w.HL + (long long)k * NH_PAD};
    LamY<false> lam{w.Y + 36 + RK * k};
    if (P.pb.formulation == 1) knot_eval<false, false, false, true, HSinkP, LamY<false>, true, KnotRef>(kn, s, lam);
    else knot_eval<false, false, false, true, HSinkP, LamY<false>, false, KnotRef>(kn, s, lam);
  }
}
__device__ __noinline__ void eval_knot_g(const KParams& P, const double* x, double* gout, int k0, int kstep) {
  for (int k = k0; k < P.K; k += kstep) {
    KnotRef kn;
    load_knot(P, x, P.zeros, k, kn);
    NoLam nl;
    GSinkP s{gout + 36 + RK * k};
    if (P.pb.formulation == 1) knot_eval<false, true, false, false, GSinkP, NoLam, true, KnotRef>(kn, s, nl);
    else knot_eval<false, true, false, false, GSinkP, NoLam, false, KnotRef>(kn, s, nl);
  }
}

// running costs: GRF cost of the "CCC" variant (landing_problem.Qf) and the running state cost of the fixed-schedule
// formulation (landing_problem.QX, quadruped_SRBM_NLP.m:85-92); all zero for the landingCtrller_IPOPT problem
__device__ __forceinline__ bool has_run_cost(const KParams& P) {
  return P.pb.Qf[0] != 0.0 || P.pb.Qf[1] != 0.0 || P.pb.Qf[2] != 0.0 || P.run_qx;
}
// Xref of knot k < N-1, component i (linspace of the drop condition towards the terminal reference; the roundings of
// oracle/srb_ref.c: srb_build_p_x0)
__device__ __forceinline__ double xref_at(const KParams& P, const double* drop, int k, int i) {
  const double ref = i < 6 ? P.pb.q_term_ref[i] : P.pb.qd_term_ref[i - 6];
  return __dadd_rn(drop[i], __dmul_rn(ref - drop[i], (double)k / (double)(P.N - 1)));
}

// this thread's share of sum_k dt_k (sum_j Qf[j%3] f_kj^2 + sum_i QX_i (X_ki - Xref_ki)^2) (dx == nullptr) or of its
// directional derivative along dx
__device__ __noinline__ double run_cost_part(const KParams& P, const double* drop, const double* x, const double* dx) {
  const int N = P.N;
  double acc = 0.0;
  for (int item = TID; item < P.K * 12; item += NT) {
    const int k = item / 12, j = item - 12 * k, iv = 12 * N + 24 * k + 12 + j;
    const double h = __ldg(P.dtv + k);
    const double q = P.pb.Qf[j % 3] * h * x[iv];
    acc += dx ? 2.0 * q * dx[iv] : q * x[iv];
    if (P.run_qx) {
      const int ix = 12 * k + j;
      const double e = x[ix] - xref_at(P, drop, k, j), qe = P.pb.QX[j] * h * e;
      acc += dx ? 2.0 * qe * dx[ix] : qe * e;
    }
  }
  return acc;
}

// g(x) (and, when LISTS, the J/H entry lists with multipliers Y) for all knots; returns f(x)
template <bool LISTS>
__device__ double eval_all(const KParams& P, const Ws& w, const double* drop, const double* x, double* gout, double* red) {
  const int N = P.N, tid = TID;
  if (LISTS) {
    // Jacobian lists on the warps of the first half of the CTA, Hessian lists on the second half; thread = knot
    constexpr int HALF = NT / 2;
    if (tid < HALF) eval_knot_j(P, w, x, gout, tid, HALF);
    else eval_knot_h(P, w, x, tid - HALF, HALF);
  } else {
    eval_knot_g(P, x, gout, tid, NT);
  }
  // boundary rows 0..35 and objective (generate_landingCtrller_IPOPT.m:83-97)
  double fl = 0.0;
  const int xo = 12 * (N - 1);
  if (tid >= NT - 32 && tid < NT - 20) {
    const int i = tid - (NT - 32);
    gout[i] = x[i];
    const double q = x[xo + i];
    const int r1 = i < 6 ? 12 + i : 24 + (i - 6);
    gout[r1] = q;
    gout[r1 + 6] = q;
    const double ref = i < 6 ? P.pb.q_term_ref[i] : P.pb.qd_term_ref[i - 6];
    fl = P.pb.QN[i] * (q - ref) * (q - ref);
  }
  if (has_run_cost(P)) fl += run_cost_part(P, drop, x, nullptr);  // (block-uniform)
  return bsum(red, fl);  // (syncs: gout / lists are visible to the whole CTA afterwards)
}

struct StepInfo {
  double a_pr, a_du, dphi_bar, phi_bar, theta;  // barrier part of dphi; -sum of logs and theta at the current point (if asked)
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
    si.phi_bar -= log(d);
  }
  if (isfinite(ub)) {
    const double d = ub - s, z = w.ZU[idx];
    yn += mu / d;
    dzu = mu / d - z + z / d * ds;
    if (ds > 0) si.a_pr = fmin(si.a_pr, tau * d / ds);
    if (dzu < 0) si.a_du = fmin(si.a_du, -tau * z / dzu);
    si.dphi_bar += mu * ds / d;
    si.phi_bar -= log(d);
  }
  w.DS[idx] = ds;
  w.YN[idx] = yn;
  w.DZL[idx] = dzl;
  w.DZU[idx] = dzu;
}

// index tables of the sweeps: shared memory (one copy per CTA) or global memory through the L1 (TBL_SMEM)
__device__ __forceinline__ const int* sweep_tables(const KParams& P, const double* smem) {
  if constexpr (TBL_SMEM) return reinterpret_cast<const int*>(smem + SM_TBL);
  else return P.tab.sm_src;
}
__device__ __forceinline__ int tld(const int* p) {
  if constexpr (TBL_SMEM) return *p;
  else return __ldg(p);
}

#include "sweeps.cuh"

// stage variable (X 0-11, c 12-23, f 24-35, c+ 36-47) of knot k -> index into x / dx
__device__ __forceinline__ int stage_var(int N, int k, int sv) {
  if (sv < 12) return 12 * k + sv;
  if (sv < 36) return 12 * N + 24 * k + (sv - 12);
  return 12 * N + 24 * (k + 1) + (sv - 36);
}

// row buffer of one knot for row_steps (doubles): J list | s | g | sigma | zL | zU | dx of the 48 stage variables
constexpr int RB_J = 0, RB_S = NJ_PAD, RB_G = RB_S + RK, RB_SIG = RB_G + RK, RB_ZL = RB_SIG + RK, RB_ZU = RB_ZL + RK,
              RB_DX = RB_ZU + RK, RB_SIZE = RB_DX + NW + 4;
static_assert(RB_SIZE % 2 == 0 && RB_SIZE <= LB_REGION && 3 * RB_SIZE <= SM_SWEEP, "row buffers alias the sweep regions");

// knots per round of row_steps / threads per knot
constexpr int RG = NT >= 256 ? 2 : 1, TG = NT / RG, RS_DIST = 4 / RG - 1;  // RS_DIST: rounds prefetched ahead (ring of 4)

__device__ __forceinline__ void prefetch_rows(const Ws& w, int N, int K, int k, double* rb) {
  // called by the TG threads of one group (t = TID % TG) for knot k
  const int t = TID % TG, r0 = 36 + RK * k;
  const double* Jk = w.JL + (long long)k * NJ_PAD;
  constexpr int H = RK / 2, C_J = NJ_PAD / 2, C_S = C_J + H, C_G = C_S + H, C_SIG = C_G + H, C_ZL = C_SIG + H, C_ZU = C_ZL + H,
                C_X = C_ZU + 6, C_CF = C_X + 12, C_END = C_CF + 6;
  for (int i = t; i < C_END; i += TG) {
    if (i < C_J) cp_async16(rb + RB_J + 2 * i, Jk + 2 * i);
    else if (i < C_S) cp_async16(rb + RB_S + 2 * (i - C_J), w.S + r0 + 2 * (i - C_J));
    else if (i < C_G) cp_async16(rb + RB_G + 2 * (i - C_S), w.G + r0 + 2 * (i - C_S));
    else if (i < C_SIG) cp_async16(rb + RB_SIG + 2 * (i - C_G), w.SIG + r0 + 2 * (i - C_G));
    else if (i < C_ZL) cp_async16(rb + RB_ZL + 2 * (i - C_SIG), w.ZL + r0 + 2 * (i - C_SIG));
    else if (i < C_ZU) cp_async16(rb + RB_ZU + 2 * (i - C_ZL), w.ZU + r0 + 2 * (i - C_ZL));
    else if (i < C_X) cp_async16(rb + RB_DX + 2 * (i - C_ZU), w.dx + 12 * k + 2 * (i - C_ZU));                      // dX_k
    else if (i < C_CF) cp_async16(rb + RB_DX + 12 + 2 * (i - C_X), w.dx + 12 * N + 24 * k + 2 * (i - C_X));          // dc_k, df_k
    else if (k + 1 < K) cp_async16(rb + RB_DX + 36 + 2 * (i - C_CF), w.dx + 12 * N + 24 * (k + 1) + 2 * (i - C_CF)); // dc_{k+1}
    else { rb[RB_DX + 36 + 2 * (i - C_CF)] = 0.0; rb[RB_DX + 37 + 2 * (i - C_CF)] = 0.0; }
  }
}

// per-row step recovery from shared-memory copies of the row data.  MERIT: also the pieces of the merit function at the
// CURRENT point (sum of logs, theta); they are normally carried over from the accepted trial point of the previous
// iteration (same numbers), so the two logarithms per row are only evaluated after a (re)start.
__device__ __forceinline__ void row_step_sm(const bool MERIT, const Ws& w, int idx, const double* rb, int rho, double lb, double ub,
                                            double jdx, double mu, double tau, StepInfo& si) {
  const double s = rb[RB_S + rho], rd = rb[RB_G + rho] - s;
  const double ds = jdx + rd;
  double yn = rb[RB_SIG + rho] * ds, dzl = 0.0, dzu = 0.0;
  if (MERIT) si.theta += fabs(rd);
  if (isfinite(lb)) {
    const double d = s - lb, z = rb[RB_ZL + rho], id = 1.0 / d, mid = mu * id;
    yn -= mid;
    dzl = mid - z - z * id * ds;
    if (ds < 0) si.a_pr = fmin(si.a_pr, -tau * d / ds);
    if (dzl < 0) si.a_du = fmin(si.a_du, -tau * z / dzl);
    si.dphi_bar -= mid * ds;
    if (MERIT) si.phi_bar -= log(d);
  }
  if (isfinite(ub)) {
    const double d = ub - s, z = rb[RB_ZU + rho], id = 1.0 / d, mid = mu * id;
    yn += mid;
    dzu = mid - z + z * id * ds;
    if (ds > 0) si.a_pr = fmin(si.a_pr, tau * d / ds);
    if (dzu < 0) si.a_du = fmin(si.a_du, -tau * z / dzu);
    si.dphi_bar += mid * ds;
    if (MERIT) si.phi_bar -= log(d);
  }
  w.DS[idx] = ds;
  w.YN[idx] = yn;
  w.DZL[idx] = dzl;
  w.DZU[idx] = dzu;
}

// ds = J_row . dx + (g - s), new multipliers, dz, fraction-to-the-boundary limits, merit pieces.
// RG knots per round (TG threads each, one or two rows per thread); the knots' J lists, row data and steps arrive in
// shared memory through a 4-buffer cp.async ring, so no thread waits on a chain of dependent L2 round trips.
__device__ __noinline__ void row_steps(const bool MERIT, const KParams& P, const Ws& w, double* smem, const double* tab, const double* drop,
                                       double* red, double mu, double tau, StepInfo& si) {
  const unsigned char* kt = P.kt;
  const int N = P.N, K = P.K, tid = TID, grp = tid / TG, t = tid % TG;
  const int* t_rptr = sweep_tables(P, smem) + P.tab.o_rptr;
  const int* t_rterms = sweep_tables(P, smem) + P.tab.o_rterms;
  si.a_pr = 1.0; si.a_du = 1.0; si.dphi_bar = 0.0; si.phi_bar = 0.0; si.theta = 0.0;
  auto ring = [&](int k) {
    const int j = k & 3;
    return smem + (j == 0 ? SM_LB0 : SM_P + (j - 1) * RB_SIZE);
  };
  const int rounds = (K + RG - 1) / RG;
#pragma unroll
  for (int r = 0; r < RS_DIST; r++) {
    const int k = RG * r + grp;
    if (k < K) prefetch_rows(w, N, K, k, ring(k));
    cp_async_commit();
  }
  // boundary rows meanwhile (initial-state rows, terminal inequality rows)
  if (tid < 12) si.theta += fabs(w.G[tid] - drop[tid]);
  else if (tid < 36 && P.pb.formulation != 1) {  // (the fixed-schedule formulation has no terminal rows)
    const int j = tid - 12, i = (j < 12 ? j % 6 : 6 + (j - 12) % 6);
    row_step(w, tid, tab[tid], tab[NROWTAB + tid], w.dx[12 * (N - 1) + i], mu, tau, si);
  }
  Prof pf{P.prof, 0, 20};  // (thread 20: an inequality row)
  pf.start();
  for (int r = 0; r < rounds; r++) {
    const int k = RG * r + grp, kn = k + RG * RS_DIST;
    if (kn < K) prefetch_rows(w, N, K, kn, ring(kn));
    cp_async_commit();
    pf.lap(PH_X0);
    cp_async_wait_group<RS_DIST>();
    __syncthreads();
    pf.lap(PH_X1);
    if (k < K) {
      const double* rb = ring(k);
      const unsigned char* kinds = kt + kt[KT_CLS + k] * RK;
      for (int rho = t; rho < RK; rho += TG) {
        const int idx = 36 + RK * k + rho;
        const int kind = kinds[rho];
        if (kind == ROW_EQ) {
          si.theta += fabs(rb[RB_G + rho]);
        } else if (kind != ROW_FREE) {
          double jdx = 0.0, jd1 = 0.0;
          const int p0 = tld(t_rptr + rho), p1 = tld(t_rptr + rho + 1);
          for (int p = p0; p < p1; p += 4) {  // (lists padded to fours with null terms; dc+ is zero at the last knot)
            const int u0 = tld(t_rterms + p), u1 = tld(t_rterms + p + 1), u2 = tld(t_rterms + p + 2), u3 = tld(t_rterms + p + 3);
            jdx += rb[RB_J + (u0 & 1023)] * rb[RB_DX + (u0 >> 10)];
            jd1 += rb[RB_J + (u1 & 1023)] * rb[RB_DX + (u1 >> 10)];
            jdx += rb[RB_J + (u2 & 1023)] * rb[RB_DX + (u2 >> 10)];
            jd1 += rb[RB_J + (u3 & 1023)] * rb[RB_DX + (u3 >> 10)];
          }
          jdx += jd1;
          pf.lap(PH_X2);
          if (kind == ROW_EQS) {  // y+ = y + sigma (J dx + c); no slack, no step-length limit
            const double c = rb[RB_G + rho];
            si.theta += fabs(c);
            w.YN[idx] = w.Y[idx] + rb[RB_SIG + rho] * (jdx + c);
            w.DS[idx] = 0.0; w.DZL[idx] = 0.0; w.DZU[idx] = 0.0;
          } else {
            row_step_sm(MERIT, w, idx, rb, rho, tab[36 + rho], tab[NROWTAB + 36 + rho], jdx, mu, tau, si);
          }
          pf.lap(PH_X3);
        }
      }
    }
    __syncthreads();
    pf.lap(PH_X4);
  }
  cp_async_wait_group<0>();
  double v[5] = {si.a_pr, si.a_du, si.dphi_bar, si.phi_bar, si.theta};
  block_reduce<R_MIN, R_MIN, R_SUM, R_SUM, R_SUM>(red, v);
  si.a_pr = v[0]; si.a_du = v[1]; si.dphi_bar = v[2]; si.phi_bar = v[3]; si.theta = v[4];
}

// merit function pieces at the trial point (x + a dx, s + a ds) with g(trial) in GT: phi_bar = -sum of logs (times mu
// gives the barrier part), theta
// Row passes over the 104 K + 36 rows of the iterate: one row per thread and round.  The rows of RU consecutive rounds
// are LOADED FIRST and then processed (the arrays live in the L2-resident scratch: one exposed round trip per RU rounds
// instead of one per round; with 8 warps per scenario nothing else hides that latency).
#ifndef SRB_RU
#define SRB_RU 4
#endif
constexpr int RU = SRB_RU;

__device__ __noinline__ void merit_trial(const KParams& P, const Ws& w, const double* tab, const double* drop,
                                            double* red, double alpha, double mu, double& phi_bar, double& theta) {
  const unsigned char* kt = P.kt;
  const int K = P.K, MR = P.MR;
  const double* __restrict__ GT = w.GT;
  const double* __restrict__ S = w.S;
  const double* __restrict__ DS = w.DS;
  double ph = 0.0, th = 0.0;
  for (int base = TID; base < MR; base += RU * NT) {
    double gt[RU], sv[RU], ds[RU];
#pragma unroll
    for (int u = 0; u < RU; u++) {
      const int idx = base + u * NT;
      const bool in = idx < MR;
      gt[u] = in ? GT[idx] : 0.0; sv[u] = in ? S[idx] : 0.0; ds[u] = in ? DS[idx] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < RU; u++) {
      const int idx = base + u * NT;
      if (idx >= MR) break;
      const int kind = row_kind(kt, idx);
      if (kind == ROW_FREE) continue;
      if (kind == ROW_EQ || kind == ROW_EQS) {
        th += fabs(gt[u] - (idx < 12 ? drop[idx] : 0.0));
        continue;
      }
      const int t = row_tab(idx);
      const double lb = tab[t], ub = tab[NROWTAB + t];
      const double s = sv[u] + alpha * ds[u];
      th += fabs(gt[u] - s);
      if (isfinite(lb)) ph -= log(s - lb);
      if (isfinite(ub)) ph -= log(ub - s);
    }
  }
  double v[2] = {ph, th};
  block_reduce<R_SUM, R_SUM>(red, v);
  phi_bar = v[0];
  theta = v[1];
}

struct Errs {
  double dual, prim, c0, cmu, ysum, zsum, viol;
  int nzb;
};

// sigma per row and the pieces of the optimality error (oracle/ip_ref.c: assemble)
__device__ __noinline__ void row_errors(const KParams& P, const Ws& w, const double* tab, const double* drop,
                                           double* red, double mu, Errs& e) {
  const unsigned char* kt = P.kt;
  const int K = P.K, MR = P.MR;
  const double* __restrict__ G = w.G;
  const double* __restrict__ Y = w.Y;
  const double* __restrict__ S = w.S;
  const double* __restrict__ ZL = w.ZL;
  const double* __restrict__ ZU = w.ZU;
  double* __restrict__ SIG = w.SIG;
  double dual = 0, prim = 0, c0 = 0, cmu = 0, ys = 0, zs = 0, viol = 0, nb = 0;
  for (int base = TID; base < MR; base += RU * NT) {
    double gv[RU], yv[RU], sv[RU], zl[RU], zu[RU];
#pragma unroll
    for (int u = 0; u < RU; u++) {
      const int idx = base + u * NT;
      const bool in = idx < MR;
      gv[u] = in ? G[idx] : 0.0; yv[u] = in ? Y[idx] : 0.0; sv[u] = in ? S[idx] : 0.0;
      zl[u] = in ? ZL[idx] : 0.0; zu[u] = in ? ZU[idx] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < RU; u++) {
      const int idx = base + u * NT;
      if (idx >= MR) break;
      const int kind = row_kind(kt, idx);
      if (kind == ROW_FREE) continue;
      const double g = gv[u], y = yv[u];
      ys += fabs(y);
      if (kind == ROW_EQ || kind == ROW_EQS) {
        const double c = fabs(g - (idx < 12 ? drop[idx] : 0.0));
        prim = fmax(prim, c);
        viol = fmax(viol, c);
        SIG[idx] = kind == ROW_EQS ? 1.0 / P.pb.delta_c : 0.0;  // dual-regularised equality: sigma = 1 / delta_c
        continue;
      }
      const int t = row_tab(idx);
      const double lb = tab[t], ub = tab[NROWTAB + t];
      viol = fmax(viol, fmax(tab[2 * NROWTAB + t] - g, g - tab[3 * NROWTAB + t]));
      const double s = sv[u];
      double sg = 0, rs = -y;
      if (isfinite(lb)) {
        const double d = s - lb, z = zl[u];
        sg += z / d;
        rs -= z;
        c0 = fmax(c0, fabs(z * d));
        cmu = fmax(cmu, fabs(z * d - mu));
        zs += z;
        nb += 1.0;
      }
      if (isfinite(ub)) {
        const double d = ub - s, z = zu[u];
        sg += z / d;
        rs += z;
        c0 = fmax(c0, fabs(z * d));
        cmu = fmax(cmu, fabs(z * d - mu));
        zs += z;
        nb += 1.0;
      }
      prim = fmax(prim, fabs(g - s));
      dual = fmax(dual, fabs(rs));
      SIG[idx] = sg;
    }
  }
  double v[8] = {dual, prim, c0, cmu, ys, zs, viol, nb};
  block_reduce<R_MAX, R_MAX, R_MAX, R_MAX, R_SUM, R_SUM, R_MAX, R_SUM>(red, v);
  e.dual = v[0]; e.prim = v[1]; e.c0 = v[2]; e.cmu = v[3]; e.ysum = v[4]; e.zsum = v[5]; e.viol = v[6];
  e.nzb = (int)(v[7] + 0.5);
}

__device__ __noinline__ double compl_at(const KParams& P, const Ws& w, const double* tab, double* red, double mu) {
  const unsigned char* kt = P.kt;
  const int K = P.K, MR = P.MR;
  double cmu = 0;
  for (int idx = TID; idx < MR; idx += NT) {
    if (row_kind(kt, idx) != ROW_INEQ) continue;
    const int t = row_tab(idx);
    const double lb = tab[t], ub = tab[NROWTAB + t], s = w.S[idx];
    if (isfinite(lb)) cmu = fmax(cmu, fabs(w.ZL[idx] * (s - lb) - mu));
    if (isfinite(ub)) cmu = fmax(cmu, fabs(w.ZU[idx] * (ub - s) - mu));
  }
  return bmax(red, cmu);
}

// yhat = sigma (g - s) - mu/(s-lb) + mu/(ub-s)
__device__ __noinline__ void row_yhat(const KParams& P, const Ws& w, const double* tab, double mu) {
  const unsigned char* kt = P.kt;
  const int K = P.K, MR = P.MR;
  const double* __restrict__ G = w.G;
  const double* __restrict__ S = w.S;
  const double* __restrict__ SIG = w.SIG;
  double* __restrict__ YH = w.YH;
  for (int base = TID; base < MR; base += RU * NT) {
    double gv[RU], sv[RU], sg[RU];
#pragma unroll
    for (int u = 0; u < RU; u++) {
      const int idx = base + u * NT;
      const bool in = idx < MR;
      gv[u] = in ? G[idx] : 0.0; sv[u] = in ? S[idx] : 0.0; sg[u] = in ? SIG[idx] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < RU; u++) {
      const int idx = base + u * NT;
      if (idx >= MR) break;
      const int kind = row_kind(kt, idx);
      if (kind == ROW_EQS) { YH[idx] = w.Y[idx] + sg[u] * gv[u]; continue; }  // yhat = y + sigma c
      if (kind != ROW_INEQ) { YH[idx] = 0.0; continue; }
      const int t = row_tab(idx);
      const double lb = tab[t], ub = tab[NROWTAB + t], s = sv[u];
      double yh = sg[u] * (gv[u] - s);
      if (isfinite(lb)) yh -= mu / (s - lb);
      if (isfinite(ub)) yh += mu / (ub - s);
      YH[idx] = yh;
    }
  }
}

// max |grad f + J' y| (gradient of the Lagrangian w.r.t. x): one (knot, variable) item per thread
__device__ __noinline__ double dual_inf_x(const KParams& P, const Ws& w, const double* drop, const double* smem, double* red) {
  const int N = P.N, K = P.K, tid = TID;
  const int* t_cptr = sweep_tables(P, smem) + P.tab.o_cptr;
  const int* t_cterms = sweep_tables(P, smem) + P.tab.o_cterms;
  double dmax = 0.0;
  for (int item = tid; item < K * 36; item += NT) {
    const int k = item / 36, v = item - k * 36;
    const double* Jk = w.JL + (long long)k * NJ_PAD;
    const double* yk = w.Y + 36 + RK * k;
    double a = 0.0, a1 = 0.0;
    {
      const int p0 = tld(t_cptr + v), p1 = tld(t_cptr + v + 1);
      for (int p = p0; p < p1; p += 4) {  // (lists padded to fours with null terms)
        const int u0 = tld(t_cterms + p), u1 = tld(t_cterms + p + 1), u2 = tld(t_cterms + p + 2), u3 = tld(t_cterms + p + 3);
        a += Jk[u0 & 1023] * yk[u0 >> 10];
        a1 += Jk[u1 & 1023] * yk[u1 >> 10];
        a += Jk[u2 & 1023] * yk[u2 >> 10];
        a1 += Jk[u3 & 1023] * yk[u3 >> 10];
      }
    }
    if (v >= 24 && has_run_cost(P)) a += 2.0 * P.pb.Qf[(v - 24) % 3] * w.x[12 * N + 24 * k + 12 + (v - 24)] * __ldg(P.dtv + k);
    if (v < 12 && P.run_qx) a += 2.0 * P.pb.QX[v] * (w.x[12 * k + v] - xref_at(P, drop, k, v)) * __ldg(P.dtv + k);
    if (v < NS) {
      if (k > 0) {  // what knot k-1 contributes to (X_k, c_k) through its X+ / c+ columns
        const double* Jp = Jk - NJ_PAD;
        const double* yp = yk - RK;
        const int p0 = tld(t_cptr + 36 + v), p1 = tld(t_cptr + 36 + v + 1);
        for (int p = p0; p < p1; p += 4) {
          const int u0 = tld(t_cterms + p), u1 = tld(t_cterms + p + 1), u2 = tld(t_cterms + p + 2), u3 = tld(t_cterms + p + 3);
          a += Jp[u0 & 1023] * yp[u0 >> 10];
          a1 += Jp[u1 & 1023] * yp[u1 >> 10];
          a += Jp[u2 & 1023] * yp[u2 >> 10];
          a1 += Jp[u3 & 1023] * yp[u3 >> 10];
        }
      } else if (v < 12) {
        a += w.Y[v];  // initial-state rows act on X_0
      }
    }
    a += a1;
    dmax = fmax(dmax, fabs(a));
  }
  if (tid < 12) {  // terminal state: last knot's X+ columns, objective gradient and terminal rows
    const double* Jk = w.JL + (long long)(K - 1) * NJ_PAD;
    const double* yk = w.Y + 36 + RK * (K - 1);
    double a = 0.0;
    const int p0 = tld(t_cptr + 36 + tid), p1 = tld(t_cptr + 36 + tid + 1);
    for (int p = p0; p < p1; p++) {
      const int term = tld(t_cterms + p);  // (null terms add 0 * y_0)
      a += Jk[term & 1023] * yk[term >> 10];
    }
    const int r1 = tid < 6 ? 12 + tid : 24 + (tid - 6);
    const double q = w.x[12 * (N - 1) + tid];
    const double ref = tid < 6 ? P.pb.q_term_ref[tid] : P.pb.qd_term_ref[tid - 6];
    dmax = fmax(dmax, fabs(a + 2.0 * P.pb.QN[tid] * (q - ref) + w.Y[r1] + w.Y[r1 + 6]));
  }
  return bmax(red, dmax);
}

// slacks pushed inside their bounds at the current g; mu-based bound multipliers (ip_ref.c: init_slacks)
__device__ __noinline__ void init_slacks(const KParams& P, const Ws& w, const double* tab, double mu) {
  const unsigned char* kt = P.kt;
  const int K = P.K, MR = P.MR;
  const double bp = P.opt.bound_push, bf = P.opt.bound_frac;
  for (int idx = TID; idx < MR; idx += NT) {
    w.Y[idx] = 0.0; w.ZL[idx] = 0.0; w.ZU[idx] = 0.0; w.S[idx] = 0.0;
    if (row_kind(kt, idx) != ROW_INEQ) continue;
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
  __syncthreads();
}

// bound tables of one CTA: [lb | ub | lb_orig | ub_orig] x (36 boundary rows + 104 interior knot rows)
__device__ void build_tables(const KParams& P, double* tab) {
  const landing_problem& pb = P.pb;
  const double INF = HUGE_VAL;
  for (int t = TID; t < NROWTAB; t += blockDim.x) {
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
        else if (j < 10) { lb = -pb.kin_box[j - 8]; ub = pb.kin_box[j - 8]; }
        else if (j == 10) { lb = -pb.kin_box[2]; ub = 0.0; }
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
__device__ void solve_one(const KParams& P, Ws& w, double* smem, long long b) {
  // w (shared memory): x / xt and G / GT are swapped instead of copied when a trial point is accepted
  const int N = P.N, K = P.K, nx = P.nx, MR = P.MR, tid = TID;
  const landing_options& opt = P.opt;
  const double kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99;
  const double gamma_theta = 1e-5, gamma_phi = 1e-5, eta_phi = 1e-8, s_theta = 1.1, s_phi = 2.3, delta_sw = 1.0;
  const double kappa_sigma = 1e10, s_max = 100.0;
  const double* drop = P.drops + 12 * b;
  const double* tab = smem + SM_TAB;
  const unsigned char* kt = P.kt;
  double* red = smem + SM_RED;

  // initial guess: user x0 or the reference's [Xref(:); Uref(:)] (generate_landingCtrller_IPOPT.m:199-208,336)
  if (P.x0) {
    for (int i = tid; i < nx; i += NT) w.x[i] = P.x0[b * nx + i];
  } else {
    for (int k = tid; k < N; k += NT) {
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
  // list slots never written: the padding of every knot and the no-slip entries of the last knot
  for (int i = tid; i < K * NJ_PAD; i += NT) w.JL[i] = 0.0;
  for (int i = tid; i < K * NH_PAD; i += NT) w.HL[i] = 0.0;
  for (int i = tid; i < MR; i += NT) {
    w.G[i] = 0.0; w.GT[i] = 0.0; w.DS[i] = 0.0; w.YN[i] = 0.0; w.DZL[i] = 0.0; w.DZU[i] = 0.0;
    w.SIG[i] = 0.0; w.YH[i] = 0.0; w.Y[i] = 0.0;
  }
  __syncthreads();
  double f = eval_all<false>(P, w, drop, w.x, w.G, red);
  double mu = opt.mu_init;
  init_slacks(P, w, tab, mu);

  int nfilt = 0, restarts = 0, status = LANDING_ST_MAX_ITER, it = 0, tiny = 0;
  // merit pieces of the current point, carried over from the accepted trial point (-sum of logs, theta)
  double slog_cur = 0.0, theta_cur = 0.0;
  bool have_cur = false;
  double theta0 = -1.0, dw_last = 0.0, viol = 0.0;
  Prof pf{P.prof, 0};
  for (it = 0; it <= opt.max_iter; it++) {
    pf.start();
    f = eval_all<true>(P, w, drop, w.x, w.G, red);
    pf.lap(PH_EVAL);
    Errs er;
    row_errors(P, w, tab, drop, red, mu, er);
    pf.lap(PH_ERR);
    const double dual = fmax(er.dual, dual_inf_x(P, w, drop, smem, red));
    pf.lap(PH_DUAL);
    pf.count(PH_NITER);
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
        cmu = compl_at(P, w, tab, red, mu);
      }
      if (changed) nfilt = 0;
    }
    row_yhat(P, w, tab, mu);
    __syncthreads();
    condense_all(P, w, smem, drop);
    pf.lap(PH_MU);
    const double tau = fmax(tau_min, 1.0 - mu);
    // factorise with inertia correction (IPOPT's delta_w schedule)
    double dwreg = 0.0;
    bool ok = false;
    int tries = 0;
    // while the previous iteration needed regularisation start from a third of it (ip_ref.c)
    if (dw_last > 0.0) { dwreg = dw_last / 3.0; if (dwreg < 1e-7) dwreg = 0.0; }
    for (;;) {
      pf.count(PH_NBACK);
      if (backward_sweep(P, w, smem, dwreg)) { ok = true; break; }
      __syncthreads();
      if (dwreg == 0.0) dwreg = (dw_last == 0.0) ? 1e-4 : fmax(1e-20, dw_last / 3.0);
      else dwreg *= (dw_last == 0.0 && tries < 8) ? 100.0 : 8.0;
      tries++;
      if (dwreg > 1e40) break;
    }
    if (!ok) { status = LANDING_ST_FACTOR_FAIL; break; }
    dw_last = dwreg;
    pf.lap(PH_BACK);
    forward_sweep(P, w, smem, drop);
    costates(P, w);
    pf.lap(PH_FWD);
    StepInfo si;
    if (have_cur) {
      row_steps(false, P, w, smem, tab, drop, red, mu, tau, si);
      si.phi_bar = slog_cur;
      si.theta = theta_cur;
    } else {
      row_steps(true, P, w, smem, tab, drop, red, mu, tau, si);
    }
    pf.lap(PH_ROWS);
    // filter line search
    const double theta = si.theta, phi = f + mu * si.phi_bar;
    if (theta0 < 0) theta0 = theta;
    const double theta_max = 1e4 * fmax(1.0, theta0), theta_min = 1e-4 * fmax(1.0, theta0);
    double dphi = si.dphi_bar;
    {
      double d = 0.0;
      if (tid < 12) {
        const double q = w.x[12 * (N - 1) + tid];
        const double ref = tid < 6 ? P.pb.q_term_ref[tid] : P.pb.qd_term_ref[tid - 6];
        d = 2.0 * P.pb.QN[tid] * (q - ref) * w.dx[12 * (N - 1) + tid];
      }
      if (has_run_cost(P)) d += run_cost_part(P, drop, w.x, w.dx);
      dphi += bsum(red, d);
    }
    double alpha = si.a_pr, ft = f, phb = 0.0, tht = 0.0;
    bool accepted = false, ftype = false;
    int ls = 0;
    while (alpha > 1e-12 * si.a_pr && ls < 40) {
      for (int i = tid; i < nx; i += NT) w.xt[i] = w.x[i] + alpha * w.dx[i];
      __syncthreads();
      ft = eval_all<false>(P, w, drop, w.xt, w.GT, red);
      merit_trial(P, w, tab, drop, red, alpha, mu, phb, tht);
      const double pht = ft + mu * phb;
      int filt_ok = 1;
      for (int i = tid; i < nfilt; i += NT)
        if (tht >= w.FT[i] && pht >= w.FP[i]) filt_ok = 0;
      filt_ok = __syncthreads_and(filt_ok);
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
    pf.lap(PH_LS);
    // watchdog against jamming at the fraction-to-the-boundary rule: a run of tiny accepted steps is treated like a
    // failed line search (ip_ref.c)
    if (accepted) tiny = (alpha < opt.jam_alpha) ? tiny + 1 : 0;
    if (accepted && opt.jam_iters > 0 && tiny >= opt.jam_iters && restarts < opt.max_restarts) accepted = false;
    if (!accepted) {
      tiny = 0;
      if (restarts < opt.max_restarts) {  // re-centre: slacks back inside their bounds, multipliers reset, mu = restart_mu
        restarts++;
        mu = opt.restart_mu > 0.0 ? opt.restart_mu : opt.mu_init;
        init_slacks(P, w, tab, mu);
        nfilt = 0;
        theta0 = -1.0;
        have_cur = false;
        continue;
      }
      status = LANDING_ST_LINESEARCH_FAIL;
      break;
    }
    if (!ftype) {
      if (nfilt == MAXFILTER) {  // drop the oldest entry
        double a = 0, c = 0;
        if (tid + 1 < MAXFILTER) { a = w.FT[tid + 1]; c = w.FP[tid + 1]; }
        __syncthreads();
        if (tid + 1 < MAXFILTER) { w.FT[tid] = a; w.FP[tid] = c; }
        nfilt--;
      }
      __syncthreads();
      if (tid == 0) { w.FT[nfilt] = (1.0 - gamma_theta) * theta; w.FP[nfilt] = phi - gamma_phi * theta; }
      nfilt++;
    }
    // accept the trial point (x <-> xt, G <-> GT by pointer)
    __syncthreads();
    if (tid == 0) { double* t = w.x; w.x = w.xt; w.xt = t; t = w.G; w.G = w.GT; w.GT = t; }
    __syncthreads();
    f = ft;
    slog_cur = phb; theta_cur = tht; have_cur = true;
    {
      double* __restrict__ Yp = w.Y;
      double* __restrict__ Sp = w.S;
      double* __restrict__ ZLp = w.ZL;
      double* __restrict__ ZUp = w.ZU;
      const double* __restrict__ YNp = w.YN;
      const double* __restrict__ DSp = w.DS;
      const double* __restrict__ DZLp = w.DZL;
      const double* __restrict__ DZUp = w.DZU;
      for (int base = tid; base < MR; base += RU * NT) {
        double yv[RU], yn[RU], sv[RU], ds[RU], zl[RU], zu[RU], dzl[RU], dzu[RU];
#pragma unroll
        for (int u = 0; u < RU; u++) {
          const int idx = base + u * NT;
          const bool in = idx < MR;
          yv[u] = in ? Yp[idx] : 0.0; yn[u] = in ? YNp[idx] : 0.0; sv[u] = in ? Sp[idx] : 0.0; ds[u] = in ? DSp[idx] : 0.0;
          zl[u] = in ? ZLp[idx] : 0.0; zu[u] = in ? ZUp[idx] : 0.0; dzl[u] = in ? DZLp[idx] : 0.0; dzu[u] = in ? DZUp[idx] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < RU; u++) {
          const int idx = base + u * NT;
          if (idx >= MR) break;
          const int kind = row_kind(kt, idx);
          if (kind == ROW_FREE) continue;
          Yp[idx] = yv[u] + alpha * (yn[u] - yv[u]);
          if (kind != ROW_INEQ) continue;
          const int t = row_tab(idx);
          const double lb = tab[t], ub = tab[NROWTAB + t];
          const double s = sv[u] + alpha * ds[u];
          Sp[idx] = s;
          if (isfinite(lb)) {
            const double d = s - lb;
            double z = zl[u] + si.a_du * dzl[u];
            z = fmax(fmin(z, kappa_sigma * mu / d), mu / (kappa_sigma * d));
            ZLp[idx] = z;
          }
          if (isfinite(ub)) {
            const double d = ub - s;
            double z = zu[u] + si.a_du * dzu[u];
            z = fmax(fmin(z, kappa_sigma * mu / d), mu / (kappa_sigma * d));
            ZUp[idx] = z;
          }
        }
      }
    }
    __syncthreads();
    pf.lap(PH_ACCEPT);
  }
  // results (AoS, CasADi order)
  for (int i = tid; i < nx; i += NT) P.x_star[b * nx + i] = w.x[i];
  if (P.lam_g) {
    const long long m = 104LL * N - 92;
    for (int idx = tid; idx < MR; idx += NT) {
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
  if (tid == 0) {
    P.f_star[b] = f;
    P.status[b] = status;
    P.iters[b] = it;
    if (P.viol) P.viol[b] = viol;
  }
  __syncthreads();
}

// The kernel parameters and the scratch pointers are read by every phase through references handed to the
// (deliberately not inlined) phase functions.  As a by-value kernel parameter / local struct they would live in LOCAL
// memory (a 3.4 kB stack frame) and be reloaded after every store -- ncu: 2.5 x more local than global loads, long-scoreboard
// stalls on lines that touch nothing but P.* -- so each CTA keeps one copy of both in SHARED memory.
__global__ void __launch_bounds__(NT, CTAS_PER_SM) k_solve(const __grid_constant__ KParams Pg) {
  extern __shared__ double smem[];
  __shared__ long long s_next;
  __shared__ KParams sP;
  __shared__ Ws sW;
  {
    static_assert(sizeof(KParams) % 4 == 0, "copied as 32-bit words");
    const int* src = reinterpret_cast<const int*>(&Pg);
    int* dst = reinterpret_cast<int*>(&sP);
    for (int i = TID; i < (int)(sizeof(KParams) / 4); i += NT) dst[i] = src[i];
  }
  __syncthreads();
  const KParams& P = sP;
  build_tables(P, smem + SM_TAB);
  if constexpr (TBL_SMEM) {
    int* tbl = reinterpret_cast<int*>(smem + SM_TBL);
    for (int i = TID; i < P.tab.sm_count; i += NT) tbl[i] = __ldg(P.tab.sm_src + i);
  }
  if (TID == 0) sW = carve(P.scratch + (long long)blockIdx.x * P.slot, P.N);
  __syncthreads();
  for (;;) {
    if (TID == 0) {
      const long long q = atomicAdd(P.counter, 1);
      s_next = (P.order && q < P.B) ? P.order[q] : q;  // queue position -> scenario id
    }
    __syncthreads();
    const long long b = s_next;
    __syncthreads();
    if (b >= P.B) break;
    solve_one(P, sW, smem, b);
  }
}

}  // namespace
}  // namespace srb
