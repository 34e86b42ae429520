/*
 * oracle/ip_ref.c -- TEST INFRASTRUCTURE ONLY; see ip_ref.h for what this restates and why.
 *
 * Primal-dual filter line-search interior point for
 *     min f(x)  s.t.  g_E(x) = b_E,   lb <= g_I(x) <= ub
 * with slacks s on every inequality row (as IPOPT does for CasADi's lbg/ubg form), options named
 * after the reference's IPOPT settings (generate_landingCtrller_IPOPT.m:232-263).
 *
 * Linear algebra: slacks and bound multipliers are eliminated, leaving the equality-constrained QP
 *   min 1/2 dx'(W + J_I' S J_I + dw I)dx + (grad f + J_I' yhat)'dx   s.t.  J_E dx + c_E = 0
 * whose constraints are the linearised Euler dynamics.  It is solved by a Riccati recursion over
 * stages with state (X_k, c_k) [24] and control (f_k, c_{k+1}) [24]; foot positions are promoted
 * to states because the no-slip rows f_z,k (c_{k+1} - c_k) couple neighbouring stages.  A failed
 * Cholesky of a control block means wrong inertia -> dw is increased (IPOPT's inertia correction).
 */
#include "ip_ref.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NS 24 /* stage state / control dimension */
#define NW 48 /* stage variables: X(12) c(12) f(12) c+(12) */
#define MAXFILTER 64

void ip_options_default(ip_options *o) {
  o->max_iter = 3000;
  o->tol = 1e-4;
  o->constr_viol_tol = 1e-3;
  o->dual_inf_tol = 1.0;
  o->compl_inf_tol = 1e-4;
  o->mu_init = 0.1;
  o->bound_push = 0.5;
  o->bound_frac = 0.5;
  o->bound_relax_factor = 1e-6;
  o->max_soc = 4;
  o->verbose = 0;
  o->jam_alpha = 0.02;
  o->jam_iters = 5;
  o->max_restarts = 8;
  o->run_Qf[0] = o->run_Qf[1] = o->run_Qf[2] = 0.0;
  o->kin_box[0] = 0.15; o->kin_box[1] = 0.15; o->kin_box[2] = 0.30;
  o->formulation = 0;
  o->cs = NULL;
  for (int i = 0; i < 12; i++) o->QX[i] = 0.0;
  o->delta_c = 1e-7;
  o->acceptable_tol = 1e-4;
  o->acceptable_iter = 0; /* (off by default: see DESIGN.md 3) */
  o->restart_mu = 0.0; /* = mu_init */
}

typedef struct {
  const srb_plan *pl;
  const ip_options *opt; /* (variant data: run_Qf, kin_box) */
  int N, nx, m;
  srb_jpat jpi[SRB_NJ_INT], jpl[SRB_NJ_LAST];
  srb_hpat hpi[SRB_NH_INT], hpl[SRB_NH_LAST];
  /* iterate */
  double *x, *s, *y, *zL, *zU, *lb, *ub, *lbo, *ubo, *g;
  double *xt, *st, *gt;
  double *dx, *ds, *yn, *dzL, *dzU, *sig, *yhat, *gradL;
  /* per-stage data */
  double *Jl, *Hl;              /* [(N-1) x 385], [(N-1) x 189] */
  double *M, *G, *q, *r;        /* [(N-1) x 48x48], [(N-1) x 12x36], [(N-1) x 48], [(N-1) x 12] */
  double *L, *Y, *P, *pv, *yv;  /* factors: [(N-1) x 24x24] x2, P [N x 24x24], pv [N x 24], yv */
  double L0[144];               /* Cholesky of P_0's (c,c) block */
  double dw[NW];
  double filt_th[MAXFILTER], filt_ph[MAXFILTER];
  int nfilt;
} ipws;

/* row classes: hard equality (initial state, dynamics: the constraints of the Riccati recursion) | inequality with a
 * slack | free (not a constraint of this formulation / scenario) | dual-regularised equality (fixed-schedule rows) */
enum { K_EQ = 0, K_INEQ = 1, K_FREE = 2, K_EQS = 3 };
static int row_kind_f(int N, int formulation, const int *cs, int i) {
  const int sched = formulation == 1;
  if (i < 12) return K_EQ;
  if (i < 36) return sched ? K_FREE : K_INEQ;
  const int k = (i - 36) / 104, rho = (i - 36) % 104;
  if (rho < 12) return K_EQ;
  if (!sched) return K_INEQ;
  const int last = (k == N - 2), stride = last ? 6 : 12;
  const int *c4 = cs + 4 * k;
  if (rho < 16) return c4[rho - 12] ? K_INEQ : K_EQS;          /* f_z in [0, cs f_max]: f_z = 0 in flight */
  if (rho < 16 + 4 * stride) {
    const int l = (rho - 16) / stride, j = (rho - 16) % stride;
    if (j == 0) return c4[l] ? K_EQS : K_FREE;                 /* cs c_z = 0 */
    if (j == 1) return K_FREE;                                 /* (f_z c_z row of the contact-implicit form) */
    if (!last) {
      if (j < 5) return c4[l] ? K_EQS : K_FREE;                /* cs (c+ - c) = 0 */
      if (j < 8) return K_FREE;
    }
    return K_INEQ;                                             /* kinematic box, leg length */
  }
  return K_INEQ;
}
#define ROWK(w, i) row_kind_f((w)->N, (w)->opt->formulation, (w)->opt->cs, (i))
static int is_eq_row(int N, int i) { /* hard equality rows (same in both formulations) */
  (void)N;
  if (i < 12) return 1;
  if (i < 36) return 0;
  return ((i - 36) % 104) < 12;
}
/* dynamics row (knot-local 0..11) -> state index of X_{k+1}: rows are pos,rpy,v,om; X is pos,rpy,om,v */
static int dynrow_state(int rho) { return rho < 6 ? rho : (rho < 9 ? rho + 3 : rho - 3); }
/* knot-local variable -> stage index (X 0-11, c 12-23, f 24-35, c+ 36-47); X+ is not a stage variable */
static int sidx(int v) { return v < 36 ? v : (v >= 48 ? v - 12 : -1); }
static int gvar(int N, int k, int v) {
  if (v < 12) return 12 * k + v;
  if (v < 36) return 12 * N + 24 * k + (v - 12);
  if (v < 48) return 12 * (k + 1) + (v - 36);
  return 12 * N + 24 * (k + 1) + (v - 48);
}

static ipws *ws_create(const srb_plan *pl) {
  ipws *w = (ipws *)calloc(1, sizeof(ipws));
  int N = pl->N, nx = pl->nx, m = pl->m, K = N - 1;
  w->pl = pl; w->N = N; w->nx = nx; w->m = m;
  srb_knot_pattern(0, w->jpi, w->hpi);
  srb_knot_pattern(1, w->jpl, w->hpl);
#define AL(n) ((double *)calloc((size_t)(n), sizeof(double)))
  w->x = AL(nx); w->xt = AL(nx); w->dx = AL(nx); w->gradL = AL(nx);
  w->s = AL(m); w->y = AL(m); w->zL = AL(m); w->zU = AL(m); w->lb = AL(m); w->ub = AL(m);
  w->lbo = AL(m); w->ubo = AL(m); w->g = AL(m); w->st = AL(m); w->gt = AL(m);
  w->ds = AL(m); w->yn = AL(m); w->dzL = AL(m); w->dzU = AL(m); w->sig = AL(m); w->yhat = AL(m);
  w->Jl = AL(K * SRB_NJ_INT); w->Hl = AL(K * SRB_NH_INT);
  w->M = AL(K * NW * NW); w->G = AL(K * 12 * 36); w->q = AL(K * NW); w->r = AL(K * 12);
  w->L = AL(K * NS * NS); w->Y = AL(K * NS * NS); w->P = AL(N * NS * NS); w->pv = AL(N * NS);
  w->yv = AL(K * NS);
#undef AL
  return w;
}
static void ws_free(ipws *w) {
  double **ptrs[] = {&w->x, &w->xt, &w->dx, &w->gradL, &w->s, &w->y, &w->zL, &w->zU, &w->lb, &w->ub,
                     &w->lbo, &w->ubo, &w->g, &w->st, &w->gt, &w->ds, &w->yn, &w->dzL, &w->dzU,
                     &w->sig, &w->yhat, &w->Jl, &w->Hl, &w->M, &w->G, &w->q, &w->r, &w->L, &w->Y,
                     &w->P, &w->pv, &w->yv};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(*ptrs[i]);
  free(w);
}

/* ---------------------------------------------------------------- evaluation */
/* running GRF cost sum_k sum_j Qf[j%3] f_kj^2 dt_k (generate_quadruped_SRBM_CCC.m:80-88 with Uref forces 0) */
static int has_QX(const ipws *w) {
  for (int i = 0; i < 12; i++) if (w->opt->QX[i] != 0.0) return 1;
  return 0;
}
static double run_cost(const ipws *w, const double *p, const double *x) {
  const int N = w->N;
  const double *Qf = w->opt->run_Qf, *QX = w->opt->QX;
  const int hf = (Qf[0] != 0.0 || Qf[1] != 0.0 || Qf[2] != 0.0), hx = has_QX(w);
  if (!hf && !hx) return 0.0;
  double c = 0.0;
  for (int k = 0; k < N - 1; k++) {
    const double h = p[w->pl->o_dt + k];
    if (hf)
      for (int j = 0; j < 12; j++) {
        const double fj = x[12 * N + 24 * k + 12 + j];
        c += Qf[j % 3] * fj * fj * h;
      }
    if (hx) /* running state cost (quadruped_SRBM_NLP.m:86-90), Xref = p[0 .. 12N) */
      for (int i = 0; i < 12; i++) {
        const double e = x[12 * k + i] - p[12 * k + i];
        c += QX[i] * e * e * h;
      }
  }
  return c;
}

/* knot lists of the fixed-schedule formulation: the contact-implicit template with the rows c_z, f_z (c+ - c) replaced
 * by cs c_z and cs (c+ - c) (linear: no Hessian contribution) -- values, Jacobian entries and multipliers patched */
static void knot_lists_sched(const ipws *w, const double *x, const double *p, int k, const double *lam_local,
                             double *gl, double *Jl, double *Hl) {
  const int N = w->N, last = (k == N - 2), nj = last ? SRB_NJ_LAST : SRB_NJ_INT, stride = last ? 6 : 12;
  const srb_jpat *jp = last ? w->jpl : w->jpi;
  const int *c4 = w->opt->cs + 4 * k;
  double lam[104];
  const int nrow = last ? 80 : 104;
  for (int r = 0; r < nrow; r++) lam[r] = lam_local[r];
  for (int l = 0; l < 4; l++) { /* rows that are linear (or absent) here must not reach the Hessian */
    const int L = 16 + stride * l;
    lam[L + 1] = 0.0;
    if (!last) for (int j = 2; j < 8; j++) lam[L + j] = 0.0;
  }
  srb_knot_lists(w->pl, x, p, k, lam, gl, Jl, Hl);
  const double *U = x + 12 * N + 24 * k;
  for (int l = 0; l < 4; l++) {
    const int L = 16 + stride * l;
    const double cs = (double)c4[l];
    gl[L] = cs * U[3 * l + 2];
    gl[L + 1] = 0.0;
    if (!last)
      for (int a = 0; a < 3; a++) {
        gl[L + 2 + a] = cs * (U[24 + 3 * l + a] - U[3 * l + a]);
        gl[L + 5 + a] = 0.0;
      }
  }
  for (int e = 0; e < nj; e++) {
    const int r = jp[e].r, v = jp[e].v;
    if (r < 16 || r >= 16 + 4 * stride) continue;
    const int l = (r - 16) / stride, j = (r - 16) % stride;
    const double cs = (double)c4[l];
    if (j == 0) Jl[e] = cs;                       /* d(cs c_z)/d c_z */
    else if (j == 1) Jl[e] = 0.0;
    else if (!last && j < 5) {
      if (v >= 12 && v < 24) Jl[e] = -cs;         /* c_k */
      else if (v >= 48) Jl[e] = cs;               /* c_{k+1} */
      else Jl[e] = 0.0;                           /* f_z */
    } else if (!last && j < 8) Jl[e] = 0.0;
  }
}
static double eval_g(ipws *w, const double *p, const double *x, double *g) {
  double f;
  srb_f(w->pl, x, p, &f);
  srb_g(w->pl, x, p, g);
  if (w->opt->formulation == 1) { /* rows of the fixed-schedule formulation (see knot_lists_sched) */
    const int N = w->N;
    for (int k = 0; k < N - 1; k++) {
      const int last = (k == N - 2), stride = last ? 6 : 12;
      const double *U = x + 12 * N + 24 * k;
      double *gl = g + 36 + 104 * k;
      for (int l = 0; l < 4; l++) {
        const int L = 16 + stride * l;
        const double cs = (double)w->opt->cs[4 * k + l];
        gl[L] = cs * U[3 * l + 2];
        gl[L + 1] = 0.0;
        if (!last)
          for (int a = 0; a < 3; a++) {
            gl[L + 2 + a] = cs * (U[24 + 3 * l + a] - U[3 * l + a]);
            gl[L + 5 + a] = 0.0;
          }
      }
    }
  }
  return f + run_cost(w, p, x);
}

/* barrier objective and constraint violation (1-norm) at (x, s) with g = g(x) */
static void merit(const ipws *w, double f, const double *g, const double *s, double mu, double *phi,
                  double *theta) {
  double ph = f, th = 0;
  for (int i = 0; i < w->m; i++) {
    const int kind = ROWK(w, i);
    if (kind == K_FREE) continue;
    if (kind == K_EQ) {
      th += fabs(g[i] - w->lb[i]);
    } else if (kind == K_EQS) {
      th += fabs(g[i]);
    } else {
      th += fabs(g[i] - s[i]);
      if (isfinite(w->lb[i])) ph -= mu * log(s[i] - w->lb[i]);
      if (isfinite(w->ub[i])) ph -= mu * log(w->ub[i] - s[i]);
    }
  }
  *phi = ph;
  *theta = th;
}

/* ---------------------------------------------------------------- stage assembly */
/* Builds, for the current iterate, sigma/yhat, the stage blocks M,G,q,r and the terminal block.
 * Returns the pieces of the optimality error through err[3] = {dual_inf, primal_inf, compl(mu=0)}
 * and compl at mu in *compl_mu. */
static void assemble(ipws *w, const double *p, double mu, double *err, double *compl_mu, double *ysum,
                     double *zsum, int *nzb) {
  const int N = w->N, m = w->m, K = N - 1;
  const srb_plan *pl = w->pl;
  double dual = 0, prim = 0, c0 = 0, cmu = 0, ys = 0, zs = 0;
  int nb = 0;
  /* barrier terms per inequality row */
  for (int i = 0; i < m; i++) {
    const int kind = ROWK(w, i);
    if (kind == K_FREE) { w->sig[i] = 0; w->yhat[i] = 0; continue; }
    ys += fabs(w->y[i]);
    if (kind == K_EQ) {
      prim = fmax(prim, fabs(w->g[i] - w->lb[i]));
      w->sig[i] = 0;
      w->yhat[i] = 0;
      continue;
    }
    if (kind == K_EQS) { /* dual-regularised equality: sigma = 1/delta_c, yhat = y + sigma c */
      prim = fmax(prim, fabs(w->g[i]));
      w->sig[i] = 1.0 / w->opt->delta_c;
      w->yhat[i] = w->y[i] + w->sig[i] * w->g[i];
      continue;
    }
    double sg = 0, yh = 0, rs = -w->y[i];
    if (isfinite(w->lb[i])) {
      double d = w->s[i] - w->lb[i];
      sg += w->zL[i] / d;
      yh -= mu / d;
      rs -= w->zL[i];
      c0 = fmax(c0, fabs(w->zL[i] * d));
      cmu = fmax(cmu, fabs(w->zL[i] * d - mu));
      zs += w->zL[i];
      nb++;
    }
    if (isfinite(w->ub[i])) {
      double d = w->ub[i] - w->s[i];
      sg += w->zU[i] / d;
      yh += mu / d;
      rs += w->zU[i];
      c0 = fmax(c0, fabs(w->zU[i] * d));
      cmu = fmax(cmu, fabs(w->zU[i] * d - mu));
      zs += w->zU[i];
      nb++;
    }
    double rd = w->g[i] - w->s[i];
    prim = fmax(prim, fabs(rd));
    dual = fmax(dual, fabs(rs));
    w->sig[i] = sg;
    w->yhat[i] = sg * rd + yh;
  }
  /* gradient of the Lagrangian: grad f + J' y */
  memset(w->gradL, 0, sizeof(double) * w->nx);
  {
    double f;
    double *gf = w->dx; /* scratch */
    srb_grad_f(pl, w->x, p, &f, gf);
    for (int i = 0; i < 12; i++) w->gradL[12 * (N - 1) + i] = gf[12 * (N - 1) + i];
    for (int i = 0; i < 12; i++) w->gradL[i] += w->y[i];
    for (int i = 0; i < 6; i++) {
      w->gradL[12 * (N - 1) + i] += w->y[12 + i] + w->y[18 + i];
      w->gradL[12 * (N - 1) + 6 + i] += w->y[24 + i] + w->y[30 + i];
    }
  }
  for (int k = 0; k < K; k++) {
    const int last = (k == N - 2);
    const int nj = last ? SRB_NJ_LAST : SRB_NJ_INT, nh = last ? SRB_NH_LAST : SRB_NH_INT;
    const srb_jpat *jp = last ? w->jpl : w->jpi;
    const srb_hpat *hp = last ? w->hpl : w->hpi;
    double *Jl = w->Jl + k * SRB_NJ_INT, *Hl = w->Hl + k * SRB_NH_INT;
    double gl[104];
    const int rb = 36 + 104 * k;
    if (w->opt->formulation == 1) knot_lists_sched(w, w->x, p, k, w->y + rb, gl, Jl, Hl);
    else srb_knot_lists(pl, w->x, p, k, w->y + rb, gl, Jl, Hl);
    double *M = w->M + (size_t)k * NW * NW, *G = w->G + k * 12 * 36, *q = w->q + k * NW, *r = w->r + k * 12;
    memset(M, 0, sizeof(double) * NW * NW);
    memset(G, 0, sizeof(double) * 12 * 36);
    memset(q, 0, sizeof(double) * NW);
    for (int i = 0; i < 12; i++) r[dynrow_state(i)] = -gl[i];
    /* Jacobian entries: dynamics rows -> G ; inequality rows -> sigma-weighted outer products */
    int e = 0;
    while (e < nj) {
      int rho = jp[e].r, e1 = e;
      while (e1 < nj && jp[e1].r == rho) e1++;
      for (int a = e; a < e1; a++) w->gradL[gvar(N, k, jp[a].v)] += Jl[a] * w->y[rb + rho];
      if (rho < 12) {
        for (int a = e; a < e1; a++)
          if (jp[a].v < 36) G[dynrow_state(rho) * 36 + jp[a].v] = -Jl[a];
      } else {
        const double sg = w->sig[rb + rho], yh = w->yhat[rb + rho];
        for (int a = e; a < e1; a++) {
          const int ia = sidx(jp[a].v);
          q[ia] += yh * Jl[a];
          for (int b = e; b < e1; b++) M[ia * NW + sidx(jp[b].v)] += sg * Jl[a] * Jl[b];
        }
      }
      e = e1;
    }
    for (int a = 0; a < nh; a++) {
      const int ia = sidx(hp[a].a), ib = sidx(hp[a].b);
      M[ia * NW + ib] += Hl[a];
      if (ia != ib) M[ib * NW + ia] += Hl[a];
    }
    if (last)
      for (int i = 36; i < 48; i++) M[i * NW + i] = 1.0; /* dummy c+ of the last stage */
    {                                                      /* running GRF cost: gradient and Hessian diagonal */
      const double *Qf = w->opt->run_Qf, h = p[pl->o_dt + k];
      if (Qf[0] != 0.0 || Qf[1] != 0.0 || Qf[2] != 0.0)
        for (int j = 0; j < 12; j++) {
          const double gj = 2.0 * Qf[j % 3] * w->x[12 * N + 24 * k + 12 + j] * h;
          q[24 + j] += gj;
          M[(24 + j) * NW + 24 + j] += 2.0 * Qf[j % 3] * h;
          w->gradL[gvar(N, k, 24 + j)] += gj;
        }
      if (has_QX(w)) /* running state cost */
        for (int i = 0; i < 12; i++) {
          const double gi = 2.0 * w->opt->QX[i] * (w->x[12 * k + i] - p[12 * k + i]) * h;
          q[i] += gi;
          M[i * NW + i] += 2.0 * w->opt->QX[i] * h;
          w->gradL[12 * k + i] += gi;
        }
    }
  }
  for (int i = 0; i < w->nx; i++) dual = fmax(dual, fabs(w->gradL[i]));
  err[0] = dual; err[1] = prim; err[2] = c0;
  *compl_mu = cmu; *ysum = ys; *zsum = zs; *nzb = nb;
}

/* ---------------------------------------------------------------- Riccati */
/* terminal block P_{N-1}, p_{N-1} from the terminal cost and the terminal inequality rows 12..35 */
static void terminal_block(ipws *w, const double *p, double dw_reg, int with_P) {
  const int N = w->N;
  double *P = w->P + (size_t)(N - 1) * NS * NS, *pv = w->pv + (N - 1) * NS;
  if (with_P) memset(P, 0, sizeof(double) * NS * NS);
  memset(pv, 0, sizeof(double) * NS);
  for (int i = 0; i < 12; i++) {
    const int r1 = i < 6 ? 12 + i : 24 + (i - 6), r2 = r1 + 6;
    if (with_P) P[i * NS + i] = 2.0 * p[w->pl->o_QN + i] + w->sig[r1] + w->sig[r2] + dw_reg;
    const double d = w->x[12 * (N - 1) + i] - p[12 * (N - 1) + i];
    pv[i] = 2.0 * p[w->pl->o_QN + i] * d + w->yhat[r1] + w->yhat[r2];
  }
}

/* in-place Cholesky (lower) of the n x n matrix A (row-major, leading dimension ld). 0 on success */
static int chol(double *A, int n, int ld) {
  for (int j = 0; j < n; j++) {
    double d = A[j * ld + j];
    for (int k = 0; k < j; k++) d -= A[j * ld + k] * A[j * ld + k];
    if (!(d > 1e-14)) return -1;
    d = sqrt(d);
    A[j * ld + j] = d;
    for (int i = j + 1; i < n; i++) {
      double v = A[i * ld + j];
      for (int k = 0; k < j; k++) v -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = v / d;
    }
  }
  return 0;
}

/* backward factorisation; returns 0, or -1 if some control block is not positive definite */
static int riccati_factor(ipws *w, const double *p, double dw_reg) {
  const int N = w->N, K = N - 1;
  terminal_block(w, p, dw_reg, 1);
  double Mh[NW * NW], T[12 * 36];
  for (int k = K - 1; k >= 0; k--) {
    const double *M = w->M + (size_t)k * NW * NW, *G = w->G + k * 12 * 36;
    const double *Pn = w->P + (size_t)(k + 1) * NS * NS;
    double *Lk = w->L + (size_t)k * NS * NS, *Yk = w->Y + (size_t)k * NS * NS, *Pk = w->P + (size_t)k * NS * NS;
    memcpy(Mh, M, sizeof Mh);
    for (int i = 0; i < 36; i++) Mh[i * NW + i] += dw_reg;
    /* T = Pxx G (12x36) ; Mh[0:36,0:36] += G' T ; Mh[0:36,36:48] += G' Pxc ; Mh[36:48,36:48] += Pcc */
    for (int i = 0; i < 12; i++)
      for (int j = 0; j < 36; j++) {
        double sacc = 0;
        for (int l = 0; l < 12; l++) sacc += Pn[i * NS + l] * G[l * 36 + j];
        T[i * 36 + j] = sacc;
      }
    for (int i = 0; i < 36; i++)
      for (int j = 0; j < 36; j++) {
        double sacc = 0;
        for (int l = 0; l < 12; l++) sacc += G[l * 36 + i] * T[l * 36 + j];
        Mh[i * NW + j] += sacc;
      }
    for (int i = 0; i < 36; i++)
      for (int j = 0; j < 12; j++) {
        double sacc = 0;
        for (int l = 0; l < 12; l++) sacc += G[l * 36 + i] * Pn[l * NS + 12 + j];
        Mh[i * NW + 36 + j] += sacc;
        Mh[(36 + j) * NW + i] += sacc;
      }
    for (int i = 0; i < 12; i++)
      for (int j = 0; j < 12; j++) Mh[(36 + i) * NW + 36 + j] += Pn[(12 + i) * NS + 12 + j];
    /* control block (rows/cols 24..47) */
    for (int i = 0; i < NS; i++)
      for (int j = 0; j < NS; j++) Lk[i * NS + j] = Mh[(24 + i) * NW + 24 + j];
    if (chol(Lk, NS, NS)) return -1;
    /* Y = L^-1 M_uxi  (24x24) */
    for (int j = 0; j < NS; j++)
      for (int i = 0; i < NS; i++) {
        double v = Mh[(24 + i) * NW + j];
        for (int l = 0; l < i; l++) v -= Lk[i * NS + l] * Yk[l * NS + j];
        Yk[i * NS + j] = v / Lk[i * NS + i];
      }
    /* P_k = M_xixi - Y'Y */
    for (int i = 0; i < NS; i++)
      for (int j = i; j < NS; j++) {
        double v = 0.5 * (Mh[i * NW + j] + Mh[j * NW + i]);
        for (int l = 0; l < NS; l++) v -= Yk[l * NS + i] * Yk[l * NS + j];
        Pk[i * NS + j] = v;
        Pk[j * NS + i] = v;
      }
  }
  /* free initial foot positions: (c,c) block of P_0 must be positive definite */
  for (int i = 0; i < 12; i++)
    for (int j = 0; j < 12; j++) w->L0[i * 12 + j] = w->P[(12 + i) * NS + 12 + j];
  if (chol(w->L0, 12, 12)) return -1;
  return 0;
}

/* solve with the current factorisation. rhs: stage gradients q, dynamics defects r, terminal pv
 * (set by terminal_block), initial-state step dX0.  Outputs dx (full vector) and the equality
 * multipliers of the QP (initial rows and dynamics rows) written into yn. */
static void riccati_solve(ipws *w, const double *dX0) {
  const int N = w->N, K = N - 1;
  double qh[NW], t[12];
  for (int k = K - 1; k >= 0; k--) {
    const double *G = w->G + k * 12 * 36, *q = w->q + k * NW, *r = w->r + k * 12;
    const double *Pn = w->P + (size_t)(k + 1) * NS * NS, *pn = w->pv + (k + 1) * NS;
    const double *Lk = w->L + (size_t)k * NS * NS, *Yk = w->Y + (size_t)k * NS * NS;
    double *pk = w->pv + k * NS, *yv = w->yv + k * NS;
    for (int i = 0; i < 12; i++) {
      double sacc = pn[i];
      for (int l = 0; l < 12; l++) sacc += Pn[i * NS + l] * r[l];
      t[i] = sacc;
    }
    for (int j = 0; j < 36; j++) {
      double sacc = q[j];
      for (int l = 0; l < 12; l++) sacc += G[l * 36 + j] * t[l];
      qh[j] = sacc;
    }
    for (int j = 0; j < 12; j++) {
      double sacc = q[36 + j] + pn[12 + j];
      for (int l = 0; l < 12; l++) sacc += Pn[(12 + j) * NS + l] * r[l];
      qh[36 + j] = sacc;
    }
    for (int i = 0; i < NS; i++) {
      double v = qh[24 + i];
      for (int l = 0; l < i; l++) v -= Lk[i * NS + l] * yv[l];
      yv[i] = v / Lk[i * NS + i];
    }
    for (int i = 0; i < NS; i++) {
      double v = qh[i];
      for (int l = 0; l < NS; l++) v -= Yk[l * NS + i] * yv[l];
      pk[i] = v;
    }
  }
  /* forward */
  double xi[NS], u[NS], rhs[NS];
  for (int i = 0; i < 12; i++) xi[i] = dX0[i];
  {
    const double *P0 = w->P, *p0 = w->pv;
    double b[12];
    for (int i = 0; i < 12; i++) {
      double v = -p0[12 + i];
      for (int l = 0; l < 12; l++) v -= P0[(12 + i) * NS + l] * dX0[l];
      b[i] = v;
    }
    for (int i = 0; i < 12; i++) {
      double v = b[i];
      for (int l = 0; l < i; l++) v -= w->L0[i * 12 + l] * b[l];
      b[i] = v / w->L0[i * 12 + i];
    }
    for (int i = 11; i >= 0; i--) {
      double v = b[i];
      for (int l = i + 1; l < 12; l++) v -= w->L0[l * 12 + i] * b[l];
      b[i] = v / w->L0[i * 12 + i];
    }
    for (int i = 0; i < 12; i++) xi[12 + i] = b[i];
    /* multipliers of the initial-state rows: y0 = -dV0/dX */
    for (int i = 0; i < 12; i++) {
      double v = p0[i];
      for (int l = 0; l < NS; l++) v += P0[i * NS + l] * xi[l];
      w->yn[i] = -v;
    }
  }
  for (int k = 0; k < K; k++) {
    const int last = (k == N - 2);
    const double *G = w->G + k * 12 * 36, *r = w->r + k * 12;
    const double *Lk = w->L + (size_t)k * NS * NS, *Yk = w->Y + (size_t)k * NS * NS, *yv = w->yv + k * NS;
    for (int i = 0; i < NS; i++) {
      double v = yv[i];
      for (int l = 0; l < NS; l++) v += Yk[i * NS + l] * xi[l];
      rhs[i] = -v;
    }
    for (int i = NS - 1; i >= 0; i--) {
      double v = rhs[i];
      for (int l = i + 1; l < NS; l++) v -= Lk[l * NS + i] * u[l];
      u[i] = v / Lk[i * NS + i];
    }
    for (int i = 0; i < 12; i++) {
      w->dx[12 * k + i] = xi[i];
      w->dx[12 * N + 24 * k + i] = xi[12 + i];
      w->dx[12 * N + 24 * k + 12 + i] = u[i];
    }
    double xn[NS];
    for (int i = 0; i < 12; i++) {
      double v = r[i];
      for (int l = 0; l < 12; l++) v += G[i * 36 + l] * xi[l] + G[i * 36 + 12 + l] * xi[12 + l] + G[i * 36 + 24 + l] * u[l];
      xn[i] = v;
      xn[12 + i] = last ? 0.0 : u[12 + i];
    }
    /* costate = multiplier of dynamics row k: -dV_{k+1}/dX */
    const double *Pn = w->P + (size_t)(k + 1) * NS * NS, *pn = w->pv + (k + 1) * NS;
    for (int i = 0; i < 12; i++) {
      double v = pn[i];
      for (int l = 0; l < NS; l++) v += Pn[i * NS + l] * xn[l];
      /* state index i -> dynamics row */
      const int rho = i < 6 ? i : (i < 9 ? i + 3 : i - 3);
      w->yn[36 + 104 * k + rho] = -v;
    }
    memcpy(xi, xn, sizeof xi);
  }
  for (int i = 0; i < 12; i++) w->dx[12 * (N - 1) + i] = xi[i];
}

/* ds, new inequality multipliers and bound-multiplier steps from dx */
static void recover_steps(ipws *w, double mu) {
  const int N = w->N, K = N - 1;
  /* terminal inequality rows */
  for (int i = 0; i < 12; i++) {
    const int r1 = i < 6 ? 12 + i : 24 + (i - 6), r2 = r1 + 6;
    w->ds[r1] = w->dx[12 * (N - 1) + i] + (w->g[r1] - w->s[r1]);
    w->ds[r2] = w->dx[12 * (N - 1) + i] + (w->g[r2] - w->s[r2]);
  }
  for (int k = 0; k < K; k++) {
    const int last = (k == N - 2), nj = last ? SRB_NJ_LAST : SRB_NJ_INT, rb = 36 + 104 * k;
    const srb_jpat *jp = last ? w->jpl : w->jpi;
    const double *Jl = w->Jl + k * SRB_NJ_INT;
    const int nrow = last ? 80 : 104;
    for (int rho = 12; rho < nrow; rho++) w->ds[rb + rho] = w->g[rb + rho] - w->s[rb + rho];
    for (int e = 0; e < nj; e++)
      if (jp[e].r >= 12) w->ds[rb + jp[e].r] += Jl[e] * w->dx[gvar(N, k, jp[e].v)];
  }
  for (int i = 0; i < w->m; i++) {
    if (is_eq_row(N, i)) { w->ds[i] = 0; w->dzL[i] = w->dzU[i] = 0; continue; }
    const int kind = ROWK(w, i);
    if (kind != K_INEQ) { /* regularised equality: y+ = y + sigma (J dx + c), no slack; free rows: nothing */
      w->yn[i] = kind == K_EQS ? w->y[i] + w->sig[i] * w->ds[i] : 0.0;
      w->ds[i] = 0; w->dzL[i] = w->dzU[i] = 0;
      continue;
    }
    double yn = w->sig[i] * w->ds[i];
    w->dzL[i] = w->dzU[i] = 0;
    if (isfinite(w->lb[i])) {
      const double d = w->s[i] - w->lb[i];
      yn -= mu / d;
      w->dzL[i] = mu / d - w->zL[i] - w->zL[i] / d * w->ds[i];
    }
    if (isfinite(w->ub[i])) {
      const double d = w->ub[i] - w->s[i];
      yn += mu / d;
      w->dzU[i] = mu / d - w->zU[i] + w->zU[i] / d * w->ds[i];
    }
    w->yn[i] = yn;
  }
}

/* ---------------------------------------------------------------- driver */
static int filter_ok(const ipws *w, double th, double ph) {
  for (int i = 0; i < w->nfilt; i++)
    if (th >= w->filt_th[i] && ph >= w->filt_ph[i]) return 0;
  return 1;
}
static void filter_add(ipws *w, double th, double ph) {
  if (w->nfilt == MAXFILTER) { /* drop the oldest entry */
    memmove(w->filt_th, w->filt_th + 1, sizeof(double) * (MAXFILTER - 1));
    memmove(w->filt_ph, w->filt_ph + 1, sizeof(double) * (MAXFILTER - 1));
    w->nfilt--;
  }
  w->filt_th[w->nfilt] = th;
  w->filt_ph[w->nfilt] = ph;
  w->nfilt++;
}

/* slack / multiplier (re)initialisation at the current x, g: push inside the bounds
 * (bound_push / bound_frac), mu-based bound multipliers, y_I = zU - zL, y_E = 0 */
static void init_slacks(ipws *w, const ip_options *opt, double mu) {
  const int N = w->N, m = w->m;
  for (int i = 0; i < m; i++) {
    w->y[i] = 0; w->zL[i] = 0; w->zU[i] = 0; w->s[i] = 0;
    if (is_eq_row(N, i) || ROWK(w, i) != K_INEQ) continue;
    const double l = w->lb[i], u = w->ub[i];
    double sv = w->g[i];
    if (isfinite(l) && isfinite(u)) {
      const double pL = fmin(opt->bound_push * fmax(1.0, fabs(l)), opt->bound_frac * (u - l));
      const double pU = fmin(opt->bound_push * fmax(1.0, fabs(u)), opt->bound_frac * (u - l));
      sv = fmin(fmax(sv, l + pL), u - pU);
    } else if (isfinite(l)) {
      sv = fmax(sv, l + opt->bound_push * fmax(1.0, fabs(l)));
    } else {
      sv = fmin(sv, u - opt->bound_push * fmax(1.0, fabs(u)));
    }
    w->s[i] = sv;
    /* mu-based multiplier initialisation (IPOPT bound_mult_init_method = mu-based) */
    if (isfinite(l)) w->zL[i] = mu / (sv - l);
    if (isfinite(u)) w->zU[i] = mu / (u - sv);
    w->y[i] = w->zU[i] - w->zL[i];
  }
}

static int ip_solve_ws(ipws *w, const double *p, const double *x0, const ip_options *opt,
                       double *x_out, double *lam_g_out, ip_result *res) {
  const int N = w->N, nx = w->nx, m = w->m;
  const double kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99;
  const double gamma_theta = 1e-5, gamma_phi = 1e-5, eta_phi = 1e-8, s_theta = 1.1, s_phi = 2.3, delta_sw = 1.0;
  const double kappa_sigma = 1e10, s_max = 100.0;
  memset(res, 0, sizeof *res);
  memcpy(w->x, x0, sizeof(double) * nx);
  srb_bounds(w->pl, p, w->lbo, w->ubo);
  w->opt = opt;
  for (int k = 0; k < N - 1; k++) /* kinematic box of the variant */
    for (int l = 0; l < 4; l++) {
      const int last = (k == N - 2), kin = 36 + 104 * k + 16 + (last ? 6 : 12) * l + (last ? 2 : 8);
      w->lbo[kin] = -opt->kin_box[0]; w->ubo[kin] = opt->kin_box[0];
      w->lbo[kin + 1] = -opt->kin_box[1]; w->ubo[kin + 1] = opt->kin_box[1];
      w->lbo[kin + 2] = -opt->kin_box[2]; w->ubo[kin + 2] = 0.0;
    }
  if (opt->formulation == 1)
    for (int i = 0; i < m; i++) {
      const int kind = ROWK(w, i);
      if (kind == K_FREE) { w->lbo[i] = -HUGE_VAL; w->ubo[i] = HUGE_VAL; }
      else if (kind == K_EQS) { w->lbo[i] = 0.0; w->ubo[i] = 0.0; }
    }
  /* relaxed bounds (bound_relax_factor) on inequality rows */
  for (int i = 0; i < m; i++) {
    w->lb[i] = w->lbo[i];
    w->ub[i] = w->ubo[i];
    if (opt->formulation == 1 && ROWK(w, i) == K_EQS) { w->lb[i] = -HUGE_VAL; w->ub[i] = HUGE_VAL; continue; }
    if (!is_eq_row(N, i)) {
      if (isfinite(w->lb[i])) w->lb[i] -= opt->bound_relax_factor * fmax(1.0, fabs(w->lb[i]));
      if (isfinite(w->ub[i])) w->ub[i] += opt->bound_relax_factor * fmax(1.0, fabs(w->ub[i]));
    }
  }
  double f = eval_g(w, p, w->x, w->g);
  double mu = opt->mu_init;
  init_slacks(w, opt, mu);
  w->nfilt = 0;
  double theta0 = -1, dw_last = 0;
  int status = 1, it = 0, restarts = 0, tiny = 0, n_acc = 0;
  for (it = 0; it <= opt->max_iter; it++) {
    double err[3], cmu, ysum, zsum;
    int nzb;
    assemble(w, p, mu, err, &cmu, &ysum, &zsum, &nzb);
    const double s_d = fmax(s_max, (ysum + zsum) / (double)(m + nzb)) / s_max;
    const double s_c = fmax(s_max, zsum / (double)(nzb > 0 ? nzb : 1)) / s_max;
    const double E0 = fmax(fmax(err[0] / s_d, err[1]), err[2] / s_c);
    res->dual_inf = err[0]; res->compl_inf = err[2]; res->mu = mu; res->restarts = restarts;
    /* constraint violation w.r.t. the original bounds */
    double viol = 0;
    for (int i = 0; i < m; i++) viol = fmax(viol, fmax(w->lbo[i] - w->g[i], w->g[i] - w->ubo[i]));
    res->viol = viol; res->f = f;
    if (!isfinite(E0) || !isfinite(f)) { status = 3; break; }
    if (opt->verbose)
      printf("it %4d f %.6e inf_pr %.2e inf_du %.2e compl %.2e mu %.1e E0 %.2e nf %d\n", it, f, err[1], err[0],
             err[2], mu, E0, w->nfilt);
    if (E0 <= opt->tol && err[0] <= opt->dual_inf_tol && viol <= opt->constr_viol_tol &&
        err[2] <= opt->compl_inf_tol) {
      status = 0;
      break;
    }
    if (opt->acceptable_iter > 0) { /* IPOPT: "solved to acceptable level" */
      if (E0 <= opt->acceptable_tol && err[0] <= 1e10 && viol <= 1e-2 && err[2] <= 1e-2) n_acc++;
      else n_acc = 0;
      if (n_acc >= opt->acceptable_iter) { status = 5; break; }
    }
    if (it == opt->max_iter) { status = 1; break; }
    /* monotone barrier update (Fiacco-McCormick) */
    {
      int changed = 0;
      for (;;) {
        const double Emu = fmax(fmax(err[0] / s_d, err[1]), cmu / s_c);
        if (!(Emu <= kappa_eps * mu) || mu <= opt->tol / 10.0 * 1.0000001) break;
        mu = fmax(opt->tol / 10.0, fmin(kappa_mu * mu, pow(mu, theta_mu)));
        changed = 1;
        /* complementarity error at the new mu */
        cmu = 0;
        for (int i = 0; i < m; i++) {
          if (is_eq_row(N, i)) continue;
          if (isfinite(w->lb[i])) cmu = fmax(cmu, fabs(w->zL[i] * (w->s[i] - w->lb[i]) - mu));
          if (isfinite(w->ub[i])) cmu = fmax(cmu, fabs(w->zU[i] * (w->ub[i] - w->s[i]) - mu));
        }
      }
      if (changed) {
        w->nfilt = 0;
        assemble(w, p, mu, err, &cmu, &ysum, &zsum, &nzb); /* yhat depends on mu */
      }
    }
    const double tau = fmax(tau_min, 1.0 - mu);
    /* factorise with inertia correction */
    double dw_reg = 0;
    int ok = 0, tries = 0;
    /* deviation from IPOPT (which always tries dw = 0 first): while the previous iteration needed
     * regularisation, start from a third of it -- saves one failed factorisation per iteration in
     * the non-convex phase; falls back to 0 once it has decayed below 1e-7 */
    if (dw_last > 0) { dw_reg = dw_last / 3.0; if (dw_reg < 1e-7) dw_reg = 0; }
    for (;;) {
      res->n_factor++;
      if (riccati_factor(w, p, dw_reg) == 0) { ok = 1; break; }
      if (dw_reg == 0) dw_reg = (dw_last == 0) ? 1e-4 : fmax(1e-20, dw_last / 3.0);
      else dw_reg *= (dw_last == 0 && tries < 8) ? 100.0 : 8.0;
      tries++;
      if (dw_reg > 1e40) break;
    }
    if (!ok) { status = 4; break; }
    dw_last = dw_reg;
    terminal_block(w, p, dw_reg, 0);
    double dX0[12];
    for (int i = 0; i < 12; i++) dX0[i] = -(w->g[i] - w->lb[i]);
    riccati_solve(w, dX0);
    recover_steps(w, mu);
    /* fraction to the boundary */
    double a_pr = 1.0, a_du = 1.0;
    for (int i = 0; i < m; i++) {
      if (is_eq_row(N, i)) continue;
      if (isfinite(w->lb[i])) {
        if (w->ds[i] < 0) a_pr = fmin(a_pr, -tau * (w->s[i] - w->lb[i]) / w->ds[i]);
        if (w->dzL[i] < 0) a_du = fmin(a_du, -tau * w->zL[i] / w->dzL[i]);
      }
      if (isfinite(w->ub[i])) {
        if (w->ds[i] > 0) a_pr = fmin(a_pr, tau * (w->ub[i] - w->s[i]) / w->ds[i]);
        if (w->dzU[i] < 0) a_du = fmin(a_du, -tau * w->zU[i] / w->dzU[i]);
      }
    }
    /* filter line search */
    double phi, theta;
    merit(w, f, w->g, w->s, mu, &phi, &theta);
    if (theta0 < 0) theta0 = theta;
    const double theta_max = 1e4 * fmax(1.0, theta0), theta_min = 1e-4 * fmax(1.0, theta0);
    double dphi = 0;
    for (int i = 0; i < 12; i++)
      dphi += 2.0 * p[w->pl->o_QN + i] * (w->x[12 * (N - 1) + i] - p[12 * (N - 1) + i]) * w->dx[12 * (N - 1) + i];
    if (opt->run_Qf[0] != 0.0 || opt->run_Qf[1] != 0.0 || opt->run_Qf[2] != 0.0)
      for (int k = 0; k < N - 1; k++)
        for (int j = 0; j < 12; j++) {
          const int iv = 12 * N + 24 * k + 12 + j;
          dphi += 2.0 * opt->run_Qf[j % 3] * w->x[iv] * p[w->pl->o_dt + k] * w->dx[iv];
        }
    if (has_QX(w))
      for (int k = 0; k < N - 1; k++)
        for (int i = 0; i < 12; i++)
          dphi += 2.0 * opt->QX[i] * (w->x[12 * k + i] - p[12 * k + i]) * p[w->pl->o_dt + k] * w->dx[12 * k + i];
    for (int i = 0; i < m; i++) {
      if (is_eq_row(N, i)) continue;
      if (isfinite(w->lb[i])) dphi -= mu * w->ds[i] / (w->s[i] - w->lb[i]);
      if (isfinite(w->ub[i])) dphi += mu * w->ds[i] / (w->ub[i] - w->s[i]);
    }
    double alpha = a_pr, ft = f;
    int accepted = 0, ftype = 0, ls = 0;
    const double alpha_min_fac = 1e-12;
    while (alpha > alpha_min_fac * a_pr && ls < 40) {
      for (int i = 0; i < nx; i++) w->xt[i] = w->x[i] + alpha * w->dx[i];
      for (int i = 0; i < m; i++) w->st[i] = w->s[i] + alpha * w->ds[i];
      ft = eval_g(w, p, w->xt, w->gt);
      double pht, tht;
      merit(w, ft, w->gt, w->st, mu, &pht, &tht);
      if (isfinite(pht) && isfinite(tht) && tht <= theta_max && filter_ok(w, tht, pht)) {
        const int sw = (theta <= theta_min) && (dphi < 0) &&
                       (alpha * pow(-dphi, s_phi) > delta_sw * pow(theta, s_theta));
        if (sw) {
          if (pht <= phi + eta_phi * alpha * dphi) { accepted = 1; ftype = 1; }
        } else if (tht <= (1.0 - gamma_theta) * theta || pht <= phi - gamma_phi * theta) {
          accepted = 1;
        }
      }
      if (accepted) break;
      if (opt->verbose > 2)
        printf("        ls %d alpha %.3e theta %.6e -> %.6e phi %.8e -> %.8e dphi %.3e filt_ok %d\n", ls, alpha, theta,
               tht, phi, pht, dphi, filter_ok(w, tht, pht));
      alpha *= 0.5;
      ls++;
    }
    /* watchdog against jamming at the fraction-to-the-boundary rule: a run of tiny accepted steps is treated like
     * a failed line search (IPOPT would leave such a phase through its restoration phase) */
    if (accepted) tiny = (alpha < opt->jam_alpha) ? tiny + 1 : 0;
    if (accepted && opt->jam_iters > 0 && tiny >= opt->jam_iters && restarts < opt->max_restarts) { accepted = 0; }
    if (!accepted) {
      tiny = 0;
      /* no restoration phase: re-centre instead -- slacks pushed back inside their bounds at the
       * current x, multipliers reset, barrier parameter back to restart_mu, filter cleared */
      if (restarts < opt->max_restarts) {
        restarts++;
        mu = opt->restart_mu > 0.0 ? opt->restart_mu : opt->mu_init;
        init_slacks(w, opt, mu);
        w->nfilt = 0;
        theta0 = -1;
        if (opt->verbose) printf("   -- line search failed: restart %d (theta %.6e, it %d)\n", restarts, theta, it);
        continue;
      }
      status = 2;
      break;
    }
    if (!ftype) filter_add(w, (1.0 - gamma_theta) * theta, phi - gamma_phi * theta);
    if (opt->verbose > 1) printf("      alpha_pr %.3e alpha_du %.3e ls %d dw %.1e ftype %d\n", alpha, a_du, ls, dw_reg, ftype);
    /* accept */
    memcpy(w->x, w->xt, sizeof(double) * nx);
    memcpy(w->g, w->gt, sizeof(double) * m);
    f = ft;
    for (int i = 0; i < m; i++) {
      if (is_eq_row(N, i)) {
        w->y[i] += alpha * (w->yn[i] - w->y[i]);
        continue;
      }
      w->s[i] = w->st[i];
      w->y[i] += alpha * (w->yn[i] - w->y[i]);
      if (isfinite(w->lb[i])) {
        const double d = w->s[i] - w->lb[i];
        double z = w->zL[i] + a_du * w->dzL[i];
        z = fmax(fmin(z, kappa_sigma * mu / d), mu / (kappa_sigma * d));
        w->zL[i] = z;
      }
      if (isfinite(w->ub[i])) {
        const double d = w->ub[i] - w->s[i];
        double z = w->zU[i] + a_du * w->dzU[i];
        z = fmax(fmin(z, kappa_sigma * mu / d), mu / (kappa_sigma * d));
        w->zU[i] = z;
      }
    }
  }
  res->status = status;
  res->iters = it;
  if (x_out) memcpy(x_out, w->x, sizeof(double) * nx);
  if (lam_g_out) memcpy(lam_g_out, w->y, sizeof(double) * m);
  return status;
}

int ip_solve(const srb_plan *pl, const double *p, const double *x0, const ip_options *opt,
             double *x_out, double *lam_g_out, ip_result *res) {
  ipws *w = ws_create(pl);
  int rc = ip_solve_ws(w, p, x0, opt, x_out, lam_g_out, res);
  ws_free(w);
  return rc;
}

int ip_solve_batch(int N, int B, const double *drops, const srb_problem *pb, const ip_options *opt,
                   double *x_out, ip_result *res, int nthreads) {
  srb_plan *pl = srb_plan_create(N);
  if (!pl) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel
  {
    ipws *w = ws_create(pl);
    double *p = (double *)malloc(sizeof(double) * pl->np);
    double *x0 = (double *)malloc(sizeof(double) * pl->nx);
#pragma omp for schedule(dynamic)
    for (int b = 0; b < B; b++) {
      srb_build_p_x0(pl, pb, drops + 12 * b, drops + 12 * b + 6, p, x0);
      ip_solve_ws(w, p, x0, opt, x_out ? x_out + (size_t)b * pl->nx : NULL, NULL, &res[b]);
    }
    free(p);
    free(x0);
    ws_free(w);
  }
  srb_plan_free(pl);
  return 0;
}
