/*
 * oracle/srb_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle); see srb_ref.h.
 *
 * Generic-N restatement of the reference's CasADi-generated NLP functions.
 * Math spec (reference file:line):
 *   model / rows      generate_landingCtrller_IPOPT.m:83-169
 *   rotation ZYX      utilities_general/dynamics-utilities/rpyToRotMat.m:2,
 *                     spatial_v2/3D/rx.m:11-13, ry.m:11-13, rz.m:11-13
 *   Euler-rate map    utilities_general/dynamics-utilities/Binv.m:13-17
 *   hip offsets       utilities_general/dynamics-utilities/get_robot_params.m:90-91
 *   gravity           utilities_general/dynamics-utilities/get_robot_model.m:140
 *   row canonical form optistack_internal.cpp:742-870
 * Derivatives are derived by hand (product rule on Rz*Ry*Rx and Binv) -- the
 * reference obtains them by CasADi AD -- and are pinned numerically against the
 * compiled reference C at N=21.
 */
#include "srb_ref.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double HIP[4][3] = {
    {0.19, -0.1, 0.0}, {0.19, 0.1, 0.0}, {-0.19, -0.1, 0.0}, {-0.19, 0.1, 0.0}};
static const double GRAV[3] = {0.0, 0.0, -9.81};
#define FRIC 0.71

/* knot-local variable numbering: 0-11 X, 12-23 c, 24-35 f, 36-47 X+, 48-59 c+ */
typedef srb_jpat jpat_t;
typedef srb_hpat hpat_t;

static void mat3mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void matvec(const double A[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
static void mattvec(const double A[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
static void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

static const int PAIR[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}}; /* (i,j)->00,01,02,11,12,22 */

typedef struct {
  double R[9], dR[3][9], ddR[6][9];
  double B[9], dB[3][9], ddB[6][9]; /* dB[0]=0, ddB[0..2]=0 (no roll dependence) */
} rot_t;

static void rot_setup(const double th[3], rot_t *q, int order) {
  double sf = sin(th[0]), cf = cos(th[0]);
  double st = sin(th[1]), ct = cos(th[1]);
  double sp = sin(th[2]), cp = cos(th[2]);
  double Rx[3][9] = {{1, 0, 0, 0, cf, -sf, 0, sf, cf},
                     {0, 0, 0, 0, -sf, -cf, 0, cf, -sf},
                     {0, 0, 0, 0, -cf, sf, 0, -sf, -cf}};
  double Ry[3][9] = {{ct, 0, st, 0, 1, 0, -st, 0, ct},
                     {-st, 0, ct, 0, 0, 0, -ct, 0, -st},
                     {-ct, 0, -st, 0, 0, 0, st, 0, -ct}};
  double Rz[3][9] = {{cp, -sp, 0, sp, cp, 0, 0, 0, 1},
                     {-sp, -cp, 0, cp, -sp, 0, 0, 0, 0},
                     {-cp, sp, 0, -sp, -cp, 0, 0, 0, 0}};
  double T[9];
#define RPROD(dz, dy, dx, out) \
  do { mat3mul(Rz[dz], Ry[dy], T); mat3mul(T, Rx[dx], out); } while (0)
  RPROD(0, 0, 0, q->R);
  if (order >= 1) {
    RPROD(0, 0, 1, q->dR[0]);
    RPROD(0, 1, 0, q->dR[1]);
    RPROD(1, 0, 0, q->dR[2]);
  }
  if (order >= 2) {
    RPROD(0, 0, 2, q->ddR[0]);
    RPROD(0, 1, 1, q->ddR[1]);
    RPROD(1, 0, 1, q->ddR[2]);
    RPROD(0, 2, 0, q->ddR[3]);
    RPROD(1, 1, 0, q->ddR[4]);
    RPROD(2, 0, 0, q->ddR[5]);
  }
#undef RPROD
  /* Binv.m:13-17 */
  double ic = 1.0 / ct, tt = st / ct;
  double B[9] = {cp * ic, sp * ic, 0, -sp, cp, 0, cp * tt, sp * tt, 1};
  memcpy(q->B, B, sizeof B);
  if (order >= 1) {
    double ic2 = ic * ic;
    double dBt[9] = {cp * st * ic2, sp * st * ic2, 0, 0, 0, 0, cp * ic2, sp * ic2, 0};
    double dBp[9] = {-sp * ic, cp * ic, 0, -cp, -sp, 0, -sp * tt, cp * tt, 0};
    memset(q->dB[0], 0, sizeof dBt);
    memcpy(q->dB[1], dBt, sizeof dBt);
    memcpy(q->dB[2], dBp, sizeof dBp);
    if (order >= 2) {
      double ic3 = ic2 * ic, a = (1.0 + st * st) * ic3, b = 2.0 * st * ic3;
      double tt_[9] = {cp * a, sp * a, 0, 0, 0, 0, cp * b, sp * b, 0};
      double tp_[9] = {-sp * st * ic2, cp * st * ic2, 0, 0, 0, 0, -sp * ic2, cp * ic2, 0};
      double pp_[9] = {-cp * ic, -sp * ic, 0, sp, -cp, 0, -cp * tt, -sp * tt, 0};
      memset(q->ddB[0], 0, sizeof tt_);
      memset(q->ddB[1], 0, sizeof tt_);
      memset(q->ddB[2], 0, sizeof tt_);
      memcpy(q->ddB[3], tt_, sizeof tt_);
      memcpy(q->ddB[4], tp_, sizeof tp_);
      memcpy(q->ddB[5], pp_, sizeof pp_);
    }
  }
}

/* knot-local row bases */
#define LEGB(l) (16 + (last ? 6 : 12) * (l))
#define KINB(l) (LEGB(l) + (last ? 2 : 8))
#define FRB (last ? 40 : 64)
#define STB (last ? 56 : 80)

/*
 * One knot: rows of generate_landingCtrller_IPOPT.m:106-169.
 * X (12), c (12), f (12) of knot k; Xn, cn of knot k+1 (cn unused when last).
 * prm = {h, mu, mass, Ib[3], Ibinv[3]}.
 * g  : 104 (80 when last) row values, or NULL.
 * Jv : Jacobian values in emission order (385 / 313), or NULL; jp: pattern capture.
 * lam/Hv : local multipliers -> Hessian values in emission order (189 / 177); hp pattern.
 */
static void knot_eval(const double *X, const double *c, const double *f, const double *Xn,
                      const double *cn, const double *prm, int last, double *g, double *Jv,
                      jpat_t *jp, const double *lam, double *Hv, hpat_t *hp) {
  const double h = prm[0], mu = prm[1], mass = prm[2];
  const double *Ib = prm + 3, *Ibinv = prm + 6;
  const double *r = X, *th = X + 3, *om = X + 6, *v = X + 9;
  const int wantJ = (Jv != NULL) || (jp != NULL);
  const int wantH = (Hv != NULL) || (hp != NULL);
  int nj = 0, nh = 0;
#define JSET(row, var, val) \
  do { if (jp) { jp[nj].r = (short)(row); jp[nj].v = (short)(var); } else Jv[nj] = (val); nj++; } while (0)
#define HSET(va, vb, val) \
  do { if (hp) { hp[nh].a = (short)(va); hp[nh].b = (short)(vb); } else Hv[nh] = (val); nh++; } while (0)

  rot_t q;
  rot_setup(th, &q, wantH ? 2 : (wantJ ? 1 : 0));

  /* ---- shared quantities ---- */
  double F[3] = {0, 0, 0}, tau[3] = {0, 0, 0}, arm[4][3];
  for (int l = 0; l < 4; l++) {
    double t[3];
    for (int a = 0; a < 3; a++) arm[l][a] = c[3 * l + a] - r[a];
    cross3(arm[l], f + 3 * l, t);
    for (int a = 0; a < 3; a++) { tau[a] += t[a]; F[a] += f[3 * l + a]; }
  }
  double u[3], e[3], Rt_tau[3], w[3];
  matvec(q.R, om, u);    /* world angular velocity */
  matvec(q.B, u, e);     /* Euler rates */
  mattvec(q.R, tau, Rt_tau);
  w[0] = om[1] * (Ib[2] * om[2]) - om[2] * (Ib[1] * om[1]);
  w[1] = om[2] * (Ib[0] * om[0]) - om[0] * (Ib[2] * om[2]);
  w[2] = om[0] * (Ib[1] * om[1]) - om[1] * (Ib[0] * om[0]);
  double kap[3] = {-h * Ibinv[0], -h * Ibinv[1], -h * Ibinv[2]};

  double hw[4][3], prel[4][3];
  for (int l = 0; l < 4; l++) {
    matvec(q.R, HIP[l], hw[l]);
    for (int a = 0; a < 3; a++) prel[l][a] = arm[l][a] - hw[l][a];
  }

  /* ---- values ---- */
  if (g) {
    for (int a = 0; a < 3; a++) {
      g[a] = Xn[a] - r[a] - v[a] * h;
      g[3 + a] = Xn[3 + a] - th[a] - e[a] * h;
      g[6 + a] = Xn[9 + a] - v[a] - (F[a] / mass + GRAV[a]) * h;
      g[9 + a] = Xn[6 + a] - om[a] - Ibinv[a] * (Rt_tau[a] - w[a]) * h;
    }
    for (int l = 0; l < 4; l++) {
      double fz = f[3 * l + 2], cz = c[3 * l + 2];
      g[12 + l] = fz;
      int L = LEGB(l), K = KINB(l);
      g[L] = cz;
      g[L + 1] = fz * cz;
      if (!last)
        for (int a = 0; a < 3; a++) {
          double ns = fz * (cn[3 * l + a] - c[3 * l + a]);
          g[L + 2 + a] = ns;
          g[L + 5 + a] = ns;
        }
      g[K] = prel[l][0];
      g[K + 1] = prel[l][1];
      g[K + 2] = prel[l][2] + 0.05;
      g[K + 3] = dot3(prel[l], prel[l]);
      g[FRB + l] = f[3 * l] - FRIC * mu * fz;
      g[FRB + 4 + l] = -FRIC * mu * fz - f[3 * l];
      g[FRB + 8 + l] = f[3 * l + 1] - FRIC * mu * fz;
      g[FRB + 12 + l] = -FRIC * mu * fz - f[3 * l + 1];
    }
    for (int i = 0; i < 6; i++) {
      g[STB + i] = X[i];
      g[STB + 6 + i] = X[i];
      g[STB + 12 + i] = X[6 + i];
      g[STB + 18 + i] = X[6 + i];
    }
  }
  if (!wantJ && !wantH) return;

  /* ---- first derivatives of e = Binv * R * om ---- */
  double E[9], dE[3][9], du[3][3], de[3][3];
  mat3mul(q.B, q.R, E);
  for (int i = 0; i < 3; i++) {
    double T1[9], T2[9];
    mat3mul(q.dB[i], q.R, T1);
    mat3mul(q.B, q.dR[i], T2);
    for (int n = 0; n < 9; n++) dE[i][n] = T1[n] + T2[n];
    matvec(q.dR[i], om, du[i]);
    matvec(dE[i], om, de[i]);
  }
  double dh[4][3][3]; /* d(R hip)/dtheta_i */
  for (int l = 0; l < 4; l++)
    for (int i = 0; i < 3; i++) matvec(q.dR[i], HIP[l], dh[l][i]);

  if (wantJ) {
    const double E3[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    /* rows 0-2 */
    for (int a = 0; a < 3; a++) {
      JSET(a, 36 + a, 1.0);
      JSET(a, a, -1.0);
      JSET(a, 9 + a, -h);
    }
    /* rows 3-5 */
    for (int a = 0; a < 3; a++) {
      JSET(3 + a, 36 + 3 + a, 1.0);
      for (int i = 0; i < 3; i++) JSET(3 + a, 3 + i, -(a == i ? 1.0 : 0.0) - h * de[i][a]);
      for (int j = 0; j < 3; j++) JSET(3 + a, 6 + j, -h * E[3 * a + j]);
    }
    /* rows 6-8 */
    for (int a = 0; a < 3; a++) {
      JSET(6 + a, 36 + 9 + a, 1.0);
      JSET(6 + a, 9 + a, -1.0);
      for (int l = 0; l < 4; l++) JSET(6 + a, 24 + 3 * l + a, -h / mass);
    }
    /* rows 9-11 */
    double dw[3][3] = {{0, Ib[2] * om[2] - om[2] * Ib[1], om[1] * Ib[2] - Ib[1] * om[1]},
                       {Ib[0] * om[2] - om[2] * Ib[2], 0, om[0] * Ib[0] - Ib[2] * om[0]},
                       {Ib[1] * om[1] - om[1] * Ib[0], om[0] * Ib[1] - Ib[0] * om[0], 0}};
    for (int m = 0; m < 3; m++) {
      JSET(9 + m, 36 + 6 + m, 1.0);
      for (int b = 0; b < 3; b++) JSET(9 + m, 6 + b, -(m == b ? 1.0 : 0.0) - kap[m] * dw[m][b]);
      for (int b = 0; b < 3; b++) { /* d/dr_b : R^T (F x e_b) */
        double t[3], o[3];
        cross3(F, E3[b], t);
        mattvec(q.R, t, o);
        JSET(9 + m, b, kap[m] * o[m]);
      }
      for (int i = 0; i < 3; i++) {
        if (m == 0 && i == 0) continue; /* first column of R has no roll dependence */
        double o[3];
        mattvec(q.dR[i], tau, o);
        JSET(9 + m, 3 + i, kap[m] * o[m]);
      }
      for (int l = 0; l < 4; l++)
        for (int b = 0; b < 3; b++) {
          double t[3], o[3];
          cross3(E3[b], f + 3 * l, t);
          mattvec(q.R, t, o);
          JSET(9 + m, 12 + 3 * l + b, kap[m] * o[m]);
          cross3(arm[l], E3[b], t);
          mattvec(q.R, t, o);
          JSET(9 + m, 24 + 3 * l + b, kap[m] * o[m]);
        }
    }
    for (int l = 0; l < 4; l++) {
      double fz = f[3 * l + 2], cz = c[3 * l + 2];
      int L = LEGB(l), K = KINB(l);
      JSET(12 + l, 24 + 3 * l + 2, 1.0);
      JSET(L, 12 + 3 * l + 2, 1.0);
      JSET(L + 1, 12 + 3 * l + 2, fz);
      JSET(L + 1, 24 + 3 * l + 2, cz);
      if (!last)
        for (int rep = 0; rep < 2; rep++)
          for (int a = 0; a < 3; a++) {
            int row = L + 2 + 3 * rep + a;
            JSET(row, 12 + 3 * l + a, -fz);
            JSET(row, 24 + 3 * l + 2, cn[3 * l + a] - c[3 * l + a]);
            JSET(row, 48 + 3 * l + a, fz);
          }
      for (int a = 0; a < 3; a++) {
        JSET(K + a, 12 + 3 * l + a, 1.0);
        JSET(K + a, a, -1.0);
        for (int i = 0; i < 3; i++) {
          if (a == 2 && i == 2) continue; /* hip_z = 0: p_z independent of yaw */
          JSET(K + a, 3 + i, -dh[l][i][a]);
        }
      }
      for (int a = 0; a < 3; a++) {
        JSET(K + 3, 12 + 3 * l + a, 2.0 * prel[l][a]);
        JSET(K + 3, a, -2.0 * prel[l][a]);
      }
      for (int i = 0; i < 3; i++) JSET(K + 3, 3 + i, -2.0 * dot3(prel[l], dh[l][i]));
      JSET(FRB + l, 24 + 3 * l, 1.0);
      JSET(FRB + l, 24 + 3 * l + 2, -FRIC * mu);
      JSET(FRB + 4 + l, 24 + 3 * l, -1.0);
      JSET(FRB + 4 + l, 24 + 3 * l + 2, -FRIC * mu);
      JSET(FRB + 8 + l, 24 + 3 * l + 1, 1.0);
      JSET(FRB + 8 + l, 24 + 3 * l + 2, -FRIC * mu);
      JSET(FRB + 12 + l, 24 + 3 * l + 1, -1.0);
      JSET(FRB + 12 + l, 24 + 3 * l + 2, -FRIC * mu);
    }
    for (int i = 0; i < 6; i++) {
      JSET(STB + i, i, 1.0);
      JSET(STB + 6 + i, i, 1.0);
      JSET(STB + 12 + i, 6 + i, 1.0);
      JSET(STB + 18 + i, 6 + i, 1.0);
    }
  }

  if (wantH) {
    double zero104[104];
    if (!lam) { memset(zero104, 0, sizeof zero104); lam = zero104; }
    /* multiplier-weighted body-frame vector for the omega rows and its world images */
    double lk[3] = {lam[9] * kap[0], lam[10] * kap[1], lam[11] * kap[2]};
    double y[3], yi[3][3], yy[6][3];
    matvec(q.R, lk, y);
    for (int i = 0; i < 3; i++) matvec(q.dR[i], lk, yi[i]);
    for (int n = 0; n < 6; n++) matvec(q.ddR[n], lk, yy[n]);
    double lpp[4], lns[4][3], lpa[4][3], slpp = 0;
    for (int l = 0; l < 4; l++) {
      int L = LEGB(l), K = KINB(l);
      lpp[l] = lam[K + 3];
      slpp += lpp[l];
      for (int a = 0; a < 3; a++) {
        lpa[l][a] = lam[K + a];
        lns[l][a] = last ? 0.0 : lam[L + 2 + a] + lam[L + 5 + a];
      }
    }
    /* (X,X): r_a - r_a */
    for (int a = 0; a < 3; a++) HSET(a, a, 2.0 * slpp);
    /* r_a - theta_i */
    for (int i = 0; i < 3; i++) {
      double t[3];
      cross3(yi[i], F, t);
      for (int a = 0; a < 3; a++) {
        double s = t[a];
        for (int l = 0; l < 4; l++) s += 2.0 * lpp[l] * dh[l][i][a];
        HSET(a, 3 + i, s);
      }
    }
    /* theta_i - theta_j */
    for (int i = 0; i < 3; i++)
      for (int j = i; j < 3; j++) {
        int n = PAIR[i][j];
        /* d2 e / dth_i dth_j = ddB u + dB_i du_j + dB_j du_i + B ddu */
        double ddu[3], t1[3], t2[3], t3[3], t4[3];
        matvec(q.ddR[n], om, ddu);
        matvec(q.ddB[n], u, t1);
        matvec(q.dB[i], du[j], t2);
        matvec(q.dB[j], du[i], t3);
        matvec(q.B, ddu, t4);
        double s = 0;
        for (int a = 0; a < 3; a++) s += lam[3 + a] * (-h) * (t1[a] + t2[a] + t3[a] + t4[a]);
        s += dot3(yy[n], tau);
        for (int l = 0; l < 4; l++) {
          double ddh[3];
          matvec(q.ddR[n], HIP[l], ddh);
          s -= lpa[l][0] * ddh[0] + lpa[l][1] * ddh[1] + lpa[l][2] * ddh[2];
          s += lpp[l] * (2.0 * dot3(dh[l][i], dh[l][j]) - 2.0 * dot3(prel[l], ddh));
        }
        HSET(3 + i, 3 + j, s);
      }
    /* theta_i - omega_j */
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        if (i == 0 && j == 0) continue;
        double s = 0;
        for (int a = 0; a < 3; a++) s += lam[3 + a] * (-h) * dE[i][3 * a + j];
        HSET(3 + i, 6 + j, s);
      }
    /* omega - omega : rows 9-11, w = om x (Ib o om) */
    HSET(6, 7, -lk[2] * (Ib[1] - Ib[0]));
    HSET(6, 8, -lk[1] * (Ib[0] - Ib[2]));
    HSET(7, 8, -lk[0] * (Ib[2] - Ib[1]));
    /* (X,U) and (U,U) per leg */
    for (int l = 0; l < 4; l++) {
      int vc = 12 + 3 * l, vf = 24 + 3 * l;
      /* r_a - c_a */
      for (int a = 0; a < 3; a++) HSET(a, vc + a, -2.0 * lpp[l]);
      /* r_a - f_b, a != b : -eps_nab y_n */
      HSET(0, vf + 1, -y[2]);
      HSET(0, vf + 2, y[1]);
      HSET(1, vf + 0, y[2]);
      HSET(1, vf + 2, -y[0]);
      HSET(2, vf + 0, -y[1]);
      HSET(2, vf + 1, y[0]);
      /* theta_i - c_b, theta_i - f_b */
      for (int i = 0; i < 3; i++) {
        double t[3];
        cross3(f + 3 * l, yi[i], t);
        for (int b = 0; b < 3; b++) HSET(3 + i, vc + b, t[b] - 2.0 * lpp[l] * dh[l][i][b]);
        cross3(yi[i], arm[l], t);
        for (int b = 0; b < 3; b++) HSET(3 + i, vf + b, t[b]);
      }
      /* c_a - c_a */
      for (int a = 0; a < 3; a++) HSET(vc + a, vc + a, 2.0 * lpp[l]);
      /* c_a - f_b : eps_nab y_n, no-slip (-lns) on f_z, complementarity on cz-fz */
      HSET(vc + 0, vf + 1, y[2]);
      HSET(vc + 0, vf + 2, -y[1] - lns[l][0]);
      HSET(vc + 1, vf + 0, -y[2]);
      HSET(vc + 1, vf + 2, y[0] - lns[l][1]);
      HSET(vc + 2, vf + 0, y[1]);
      HSET(vc + 2, vf + 1, -y[0]);
      HSET(vc + 2, vf + 2, lam[LEGB(l) + 1] - lns[l][2]);
      /* f_z - c+_a */
      if (!last)
        for (int a = 0; a < 3; a++) HSET(vf + 2, 48 + 3 * l + a, lns[l][a]);
    }
  }
#undef JSET
#undef HSET
}

/* ------------------------------------------------------------------ plan */
static int gvar(int N, int k, int v) {
  if (v < 12) return 12 * k + v;
  if (v < 36) return 12 * N + 24 * k + (v - 12);
  if (v < 48) return 12 * (k + 1) + (v - 36);
  return 12 * N + 24 * (k + 1) + (v - 48);
}

typedef struct { int row, col, id; } trip_t;
static int trip_cmp(const void *a, const void *b) {
  const trip_t *x = (const trip_t *)a, *y = (const trip_t *)b;
  if (x->col != y->col) return x->col < y->col ? -1 : 1;
  if (x->row != y->row) return x->row < y->row ? -1 : 1;
  return 0;
}

static long long *build_ccs(trip_t *t, int nnz, int nrow, int ncol, int *id2nz) {
  qsort(t, nnz, sizeof(trip_t), trip_cmp);
  long long *sp = (long long *)calloc(2 + ncol + 1 + nnz, sizeof(long long));
  sp[0] = nrow;
  sp[1] = ncol;
  long long *colind = sp + 2, *row = sp + 2 + ncol + 1;
  for (int n = 0; n < nnz; n++) {
    colind[t[n].col + 1]++;
    row[n] = t[n].row;
    id2nz[t[n].id] = n;
  }
  for (int cidx = 0; cidx < ncol; cidx++) colind[cidx + 1] += colind[cidx];
  return sp;
}

srb_plan *srb_plan_create(int N) {
  if (N < 3) return NULL;
  srb_plan *pl = (srb_plan *)calloc(1, sizeof(srb_plan));
  pl->N = N;
  pl->nx = 36 * N - 24;
  pl->np = 13 * N + 81;
  pl->m = 104 * N - 92;
  pl->nnzJ = 385 * N - 421;
  pl->nnzH = 189 * (N - 1);
  int o = 12 * N;
  pl->o_dt = o; o += N - 1;
  pl->o_qmin = o; o += 6;
  pl->o_qmax = o; o += 6;
  pl->o_qdmin = o; o += 6;
  pl->o_qdmax = o; o += 6;
  pl->o_qinit = o; o += 6;
  pl->o_qdinit = o; o += 6;
  pl->o_qtmin = o; o += 6;
  pl->o_qtmax = o; o += 6;
  pl->o_qdtmin = o; o += 6;
  pl->o_qdtmax = o; o += 6;
  pl->o_QN = o; o += 12;
  pl->o_mu = o++;
  pl->o_lleg = o++;
  pl->o_fmax = o++;
  pl->o_mass = o++;
  pl->o_Ib = o; o += 3;
  pl->o_Ibinv = o; o += 3;

  /* knot templates by a pattern-capture pass */
  jpat_t jpi[SRB_NJ_INT], jpl[SRB_NJ_LAST];
  hpat_t hpi[SRB_NH_INT], hpl[SRB_NH_LAST];
  double z[60] = {0}, prm[9] = {0.03, 1, 1, 1, 1, 1, 1, 1, 1};
  knot_eval(z, z + 12, z + 24, z + 36, z + 48, prm, 0, NULL, NULL, jpi, NULL, NULL, hpi);
  knot_eval(z, z + 12, z + 24, z + 36, z + 48, prm, 1, NULL, NULL, jpl, NULL, NULL, hpl);

  /* Jacobian */
  trip_t *tj = (trip_t *)malloc(sizeof(trip_t) * pl->nnzJ);
  int *id2nz = (int *)malloc(sizeof(int) * (36 + (N - 1) * SRB_NJ_INT));
  int n = 0;
  for (int i = 0; i < 12; i++) { tj[n].row = i; tj[n].col = i; tj[n].id = i; n++; }
  for (int i = 0; i < 6; i++) {
    tj[n].row = 12 + i; tj[n].col = 12 * (N - 1) + i; tj[n].id = 12 + i; n++;
    tj[n].row = 18 + i; tj[n].col = 12 * (N - 1) + i; tj[n].id = 18 + i; n++;
    tj[n].row = 24 + i; tj[n].col = 12 * (N - 1) + 6 + i; tj[n].id = 24 + i; n++;
    tj[n].row = 30 + i; tj[n].col = 12 * (N - 1) + 6 + i; tj[n].id = 30 + i; n++;
  }
  for (int k = 0; k < N - 1; k++) {
    int last = (k == N - 2), cnt = last ? SRB_NJ_LAST : SRB_NJ_INT;
    const jpat_t *jp = last ? jpl : jpi;
    for (int e = 0; e < cnt; e++) {
      tj[n].row = 36 + 104 * k + jp[e].r;
      tj[n].col = gvar(N, k, jp[e].v);
      tj[n].id = 36 + k * SRB_NJ_INT + e;
      n++;
    }
  }
  pl->spJ = build_ccs(tj, pl->nnzJ, pl->m, pl->nx, id2nz);
  pl->jmap = (int *)malloc(sizeof(int) * (N - 1) * SRB_NJ_INT);
  for (int i = 0; i < 36; i++) pl->jbnd[i] = id2nz[i];
  for (int k = 0; k < N - 1; k++) {
    int cnt = (k == N - 2) ? SRB_NJ_LAST : SRB_NJ_INT;
    for (int e = 0; e < SRB_NJ_INT; e++)
      pl->jmap[k * SRB_NJ_INT + e] = e < cnt ? id2nz[36 + k * SRB_NJ_INT + e] : -1;
  }
  free(tj);
  free(id2nz);

  /* Hessian (upper triangle) */
  trip_t *thh = (trip_t *)malloc(sizeof(trip_t) * pl->nnzH);
  id2nz = (int *)malloc(sizeof(int) * (12 + (N - 1) * SRB_NH_INT));
  n = 0;
  for (int i = 0; i < 12; i++) {
    thh[n].row = thh[n].col = 12 * (N - 1) + i; thh[n].id = i; n++;
  }
  for (int k = 0; k < N - 1; k++) {
    int last = (k == N - 2), cnt = last ? SRB_NH_LAST : SRB_NH_INT;
    const hpat_t *hp = last ? hpl : hpi;
    for (int e = 0; e < cnt; e++) {
      thh[n].row = gvar(N, k, hp[e].a);
      thh[n].col = gvar(N, k, hp[e].b);
      thh[n].id = 12 + k * SRB_NH_INT + e;
      n++;
    }
  }
  pl->spH = build_ccs(thh, pl->nnzH, pl->nx, pl->nx, id2nz);
  pl->hmap = (int *)malloc(sizeof(int) * (N - 1) * SRB_NH_INT);
  for (int i = 0; i < 12; i++) pl->hterm[i] = id2nz[i];
  for (int k = 0; k < N - 1; k++) {
    int cnt = (k == N - 2) ? SRB_NH_LAST : SRB_NH_INT;
    for (int e = 0; e < SRB_NH_INT; e++)
      pl->hmap[k * SRB_NH_INT + e] = e < cnt ? id2nz[12 + k * SRB_NH_INT + e] : -1;
  }
  free(thh);
  free(id2nz);
  return pl;
}

void srb_plan_free(srb_plan *pl) {
  if (!pl) return;
  free(pl->spJ);
  free(pl->spH);
  free(pl->jmap);
  free(pl->hmap);
  free(pl);
}

/* ------------------------------------------------------------------ functions */
static void knot_prm(const srb_plan *pl, const double *p, int k, double prm[9]) {
  prm[0] = p[pl->o_dt + k];
  prm[1] = p[pl->o_mu];
  prm[2] = p[pl->o_mass];
  for (int i = 0; i < 3; i++) { prm[3 + i] = p[pl->o_Ib + i]; prm[6 + i] = p[pl->o_Ibinv + i]; }
}

static int all_finite(const double *a, int n) {
  for (int i = 0; i < n; i++)
    if (!isfinite(a[i])) return 0;
  return 1;
}

static double objective(const srb_plan *pl, const double *x, const double *p, double *grad12) {
  int N = pl->N;
  double fv = 0;
  for (int i = 0; i < 12; i++) {
    double d = x[12 * (N - 1) + i] - p[12 * (N - 1) + i];
    fv += p[pl->o_QN + i] * d * d;
    if (grad12) grad12[i] = 2.0 * p[pl->o_QN + i] * d;
  }
  return fv;
}

int srb_f(const srb_plan *pl, const double *x, const double *p, double *f) {
  *f = objective(pl, x, p, NULL);
  return isfinite(*f) ? 0 : -1;
}

int srb_grad_f(const srb_plan *pl, const double *x, const double *p, double *f, double *grad) {
  double g12[12];
  double fv = objective(pl, x, p, g12);
  if (f) *f = fv;
  if (grad) {
    memset(grad, 0, sizeof(double) * pl->nx);
    memcpy(grad + 12 * (pl->N - 1), g12, sizeof g12);
  }
  return (isfinite(fv) && all_finite(g12, 12)) ? 0 : -1;
}

static void boundary_rows(const srb_plan *pl, const double *x, double *g) {
  int N = pl->N;
  for (int i = 0; i < 12; i++) g[i] = x[i];
  for (int i = 0; i < 6; i++) {
    g[12 + i] = x[12 * (N - 1) + i];
    g[18 + i] = x[12 * (N - 1) + i];
    g[24 + i] = x[12 * (N - 1) + 6 + i];
    g[30 + i] = x[12 * (N - 1) + 6 + i];
  }
}

int srb_g(const srb_plan *pl, const double *x, const double *p, double *g) {
  return srb_jac_g(pl, x, p, g, NULL);
}

int srb_jac_g(const srb_plan *pl, const double *x, const double *p, double *g, double *jac) {
  int N = pl->N;
  double gl[104], Jl[SRB_NJ_INT];
  if (g) boundary_rows(pl, x, g);
  if (jac)
    for (int i = 0; i < 36; i++) jac[pl->jbnd[i]] = 1.0;
  for (int k = 0; k < N - 1; k++) {
    int last = (k == N - 2);
    double prm[9];
    knot_prm(pl, p, k, prm);
    const double *U = x + 12 * N + 24 * k;
    knot_eval(x + 12 * k, U, U + 12, x + 12 * (k + 1), last ? NULL : U + 24, prm, last,
              g ? gl : NULL, jac ? Jl : NULL, NULL, NULL, NULL, NULL);
    if (g) memcpy(g + 36 + 104 * k, gl, sizeof(double) * (last ? 80 : 104));
    if (jac) {
      const int *map = pl->jmap + k * SRB_NJ_INT;
      int cnt = last ? SRB_NJ_LAST : SRB_NJ_INT;
      for (int e = 0; e < cnt; e++) jac[map[e]] = Jl[e];
    }
  }
  int ok = 1;
  if (g) ok &= all_finite(g, pl->m);
  if (jac) ok &= all_finite(jac, pl->nnzJ);
  return ok ? 0 : -1;
}

int srb_hess_l(const srb_plan *pl, const double *x, const double *p, double lam_f,
               const double *lam_g, double *hess) {
  int N = pl->N;
  double Hl[SRB_NH_INT];
  for (int i = 0; i < 12; i++) hess[pl->hterm[i]] = 2.0 * p[pl->o_QN + i] * lam_f;
  for (int k = 0; k < N - 1; k++) {
    int last = (k == N - 2);
    double prm[9];
    knot_prm(pl, p, k, prm);
    const double *U = x + 12 * N + 24 * k;
    knot_eval(x + 12 * k, U, U + 12, x + 12 * (k + 1), last ? NULL : U + 24, prm, last, NULL, NULL,
              NULL, lam_g ? lam_g + 36 + 104 * k : NULL, Hl, NULL);
    const int *map = pl->hmap + k * SRB_NH_INT;
    int cnt = last ? SRB_NH_LAST : SRB_NH_INT;
    for (int e = 0; e < cnt; e++) hess[map[e]] = Hl[e];
  }
  return all_finite(hess, pl->nnzH) ? 0 : -1;
}

int srb_grad(const srb_plan *pl, const double *x, const double *p, double lam_f,
             const double *lam_g, double *f, double *g, double *ggx, double *ggp) {
  int N = pl->N, ok = 1;
  double g12[12];
  double fv = objective(pl, x, p, g12);
  if (f) *f = fv;
  ok &= isfinite(fv);
  double *gtmp = (double *)malloc(sizeof(double) * pl->m);
  double *jac = (double *)malloc(sizeof(double) * pl->nnzJ);
  ok &= (srb_jac_g(pl, x, p, gtmp, jac) == 0);
  if (g) memcpy(g, gtmp, sizeof(double) * pl->m);
  if (ggx) {
    const long long *colind = pl->spJ + 2, *row = pl->spJ + 2 + pl->nx + 1;
    for (int cidx = 0; cidx < pl->nx; cidx++) {
      double s = 0;
      for (long long n = colind[cidx]; n < colind[cidx + 1]; n++)
        s += jac[n] * (lam_g ? lam_g[row[n]] : 0.0);
      ggx[cidx] = s;
    }
    for (int i = 0; i < 12; i++) ggx[12 * (N - 1) + i] += lam_f * g12[i];
    ok &= all_finite(ggx, pl->nx);
  }
  if (ggp) {
    memset(ggp, 0, sizeof(double) * pl->np);
    for (int i = 0; i < 12; i++) {
      double d = x[12 * (N - 1) + i] - p[12 * (N - 1) + i];
      ggp[12 * (N - 1) + i] = -lam_f * 2.0 * p[pl->o_QN + i] * d;
      ggp[pl->o_QN + i] = lam_f * d * d;
    }
    if (lam_g)
      for (int k = 0; k < N - 1; k++) {
        int last = (k == N - 2);
        const double *lam = lam_g + 36 + 104 * k;
        const double *X = x + 12 * k, *U = x + 12 * N + 24 * k;
        const double *c = U, *fo = U + 12, *om = X + 6, *v = X + 9;
        double h = p[pl->o_dt + k], mu = p[pl->o_mu], mass = p[pl->o_mass];
        const double *Ib = p + pl->o_Ib, *Ibinv = p + pl->o_Ibinv;
        (void)mu;
        rot_t q;
        rot_setup(X + 3, &q, 0);
        double F[3] = {0, 0, 0}, tau[3] = {0, 0, 0};
        for (int l = 0; l < 4; l++) {
          double arm[3], t[3];
          for (int a = 0; a < 3; a++) arm[a] = c[3 * l + a] - X[a];
          cross3(arm, fo + 3 * l, t);
          for (int a = 0; a < 3; a++) { tau[a] += t[a]; F[a] += fo[3 * l + a]; }
        }
        double u[3], e[3], Rt_tau[3], w[3];
        matvec(q.R, om, u);
        matvec(q.B, u, e);
        mattvec(q.R, tau, Rt_tau);
        w[0] = om[1] * (Ib[2] * om[2]) - om[2] * (Ib[1] * om[1]);
        w[1] = om[2] * (Ib[0] * om[0]) - om[0] * (Ib[2] * om[2]);
        w[2] = om[0] * (Ib[1] * om[1]) - om[1] * (Ib[0] * om[0]);
        double s = 0;
        for (int a = 0; a < 3; a++) {
          s += lam[a] * (-v[a]) + lam[3 + a] * (-e[a]) + lam[6 + a] * (-(F[a] / mass + GRAV[a])) +
               lam[9 + a] * (-Ibinv[a] * (Rt_tau[a] - w[a]));
          ggp[pl->o_mass] += lam[6 + a] * h * F[a] / (mass * mass);
          ggp[pl->o_Ibinv + a] += lam[9 + a] * (-h) * (Rt_tau[a] - w[a]);
        }
        ggp[pl->o_dt + k] += s;
        /* d w / d Ib */
        double c0 = lam[9] * h * Ibinv[0], c1 = lam[10] * h * Ibinv[1], c2 = lam[11] * h * Ibinv[2];
        ggp[pl->o_Ib + 0] += c1 * (om[2] * om[0]) - c2 * (om[1] * om[0]);
        ggp[pl->o_Ib + 1] += -c0 * (om[2] * om[1]) + c2 * (om[0] * om[1]);
        ggp[pl->o_Ib + 2] += c0 * (om[1] * om[2]) - c1 * (om[0] * om[2]);
        for (int l = 0; l < 4; l++) {
          double fz = fo[3 * l + 2];
          ggp[pl->o_mu] += -FRIC * fz *
                           (lam[FRB + l] + lam[FRB + 4 + l] + lam[FRB + 8 + l] + lam[FRB + 12 + l]);
        }
      }
    ok &= all_finite(ggp, pl->np);
  }
  free(gtmp);
  free(jac);
  return ok ? 0 : -1;
}

void srb_knot_pattern(int last, srb_jpat *jp, srb_hpat *hp) {
  double z[60] = {0}, prm[9] = {0.03, 1, 1, 1, 1, 1, 1, 1, 1};
  knot_eval(z, z + 12, z + 24, z + 36, z + 48, prm, last, NULL, NULL, jp, NULL, NULL, hp);
}

void srb_knot_lists(const srb_plan *pl, const double *x, const double *p, int k,
                    const double *lam_local, double *gl, double *Jl, double *Hl) {
  int N = pl->N, last = (k == N - 2);
  double prm[9];
  knot_prm(pl, p, k, prm);
  const double *U = x + 12 * N + 24 * k;
  knot_eval(x + 12 * k, U, U + 12, x + 12 * (k + 1), last ? NULL : U + 24, prm, last, gl, Jl, NULL,
            lam_local, Hl, NULL);
}

/* ------------------------------------------------------------------ bounds */
void srb_bounds(const srb_plan *pl, const double *p, double *lbg, double *ubg) {
  int N = pl->N;
  const double INF = HUGE_VAL;
  for (int i = 0; i < 6; i++) {
    lbg[i] = ubg[i] = p[pl->o_qinit + i];
    lbg[6 + i] = ubg[6 + i] = p[pl->o_qdinit + i];
    lbg[12 + i] = p[pl->o_qtmin + i]; ubg[12 + i] = INF;
    lbg[18 + i] = -INF; ubg[18 + i] = p[pl->o_qtmax + i];
    lbg[24 + i] = p[pl->o_qdtmin + i]; ubg[24 + i] = INF;
    lbg[30 + i] = -INF; ubg[30 + i] = p[pl->o_qdtmax + i];
  }
  double lmax = p[pl->o_lleg];
  for (int k = 0; k < N - 1; k++) {
    int last = (k == N - 2);
    double *lb = lbg + 36 + 104 * k, *ub = ubg + 36 + 104 * k;
    for (int i = 0; i < 12; i++) lb[i] = ub[i] = 0.0;
    for (int l = 0; l < 4; l++) {
      lb[12 + l] = 0.0; ub[12 + l] = p[pl->o_fmax];
      int L = LEGB(l), K = KINB(l);
      lb[L] = 0.0; ub[L] = INF;
      lb[L + 1] = -INF; ub[L + 1] = 0.001;
      if (!last)
        for (int a = 0; a < 3; a++) {
          lb[L + 2 + a] = -INF; ub[L + 2 + a] = 0.01;
          lb[L + 5 + a] = -0.01; ub[L + 5 + a] = INF;
        }
      lb[K] = -0.15; ub[K] = 0.15;
      lb[K + 1] = -0.15; ub[K + 1] = 0.15;
      lb[K + 2] = -0.30; ub[K + 2] = 0.0;
      lb[K + 3] = -INF; ub[K + 3] = lmax * lmax;
    }
    for (int i = 0; i < 16; i++) { lb[FRB + i] = -INF; ub[FRB + i] = 0.0; }
    for (int i = 0; i < 6; i++) {
      lb[STB + i] = -INF; ub[STB + i] = p[pl->o_qmax + i];
      lb[STB + 6 + i] = p[pl->o_qmin + i]; ub[STB + 6 + i] = INF;
      lb[STB + 12 + i] = -INF; ub[STB + 12 + i] = p[pl->o_qdmax + i];
      lb[STB + 18 + i] = p[pl->o_qdmin + i]; ub[STB + 18 + i] = INF;
    }
  }
}

/* ------------------------------------------------------------------ problem data */
void srb_problem_default(srb_problem *pb) {
  /* generate_landingCtrller_IPOPT.m:173-196 */
  static const double qmin[6] = {-10, -10, 0.1, -10, -10, -10}, qmax[6] = {10, 10, 1.0, 10, 10, 10};
  static const double qdmin[6] = {-10, -10, -10, -40, -40, -40}, qdmax[6] = {10, 10, 10, 40, 40, 40};
  static const double qtmin[6] = {-10, -10, 0.2, -0.1, -0.1, -10}, qtmax[6] = {10, 10, 5, 0.1, 0.1, 10};
  static const double qtref[6] = {0, 0, 0.275, 0, 0, 0};
  static const double QN[12] = {0, 0, 100, 100, 100, 0, 10, 10, 10, 10, 10, 10};
  static const double sgn[12] = {1, -1, 1, 1, 1, 1, -1, -1, 1, -1, 1, 1};
  static const double cr[3] = {0.2, 0.1, -0.2};
  pb->T = 0.6;
  pb->dt = NULL;
  for (int i = 0; i < 6; i++) {
    pb->q_min[i] = qmin[i]; pb->q_max[i] = qmax[i];
    pb->qd_min[i] = qdmin[i]; pb->qd_max[i] = qdmax[i];
    pb->q_term_min[i] = qtmin[i]; pb->q_term_max[i] = qtmax[i];
    pb->qd_term_min[i] = qdmin[i]; pb->qd_term_max[i] = qdmax[i];
    pb->q_term_ref[i] = qtref[i]; pb->qd_term_ref[i] = 0.0;
  }
  for (int i = 0; i < 12; i++) { pb->QN[i] = QN[i]; pb->c_ref[i] = sgn[i] * cr[i % 3]; }
  pb->mu = 1.0;
  pb->l_leg_max = 0.35;
  pb->f_max = 200.0;
  /* CRBA at q_home (get_mass_matrix.m:19-54); restated in oracle/crba_constants.py */
  pb->mass = 8.251999999999999;
  pb->Ib[0] = 0.05757729852959269; pb->Ib[1] = 0.23400899479539086; pb->Ib[2] = 0.2796738482657981;
  pb->Ib_inv[0] = 17.37746888893693; pb->Ib_inv[1] = 4.27334000932043; pb->Ib_inv[2] = 3.577551923825657;
}

void srb_build_p_x0(const srb_plan *pl, const srb_problem *pb, const double *q_init,
                    const double *qd_init, double *p, double *x0) {
  int N = pl->N;
  /* linspace(a,b,N): MATLAB computes a + (b-a)*i/(N-1) with exact endpoints */
  for (int k = 0; k < N; k++) {
    double t = (double)k / (double)(N - 1);
    for (int i = 0; i < 6; i++) {
      p[12 * k + i] = (k == N - 1) ? pb->q_term_ref[i] : q_init[i] + (pb->q_term_ref[i] - q_init[i]) * t;
      p[12 * k + 6 + i] =
          (k == N - 1) ? pb->qd_term_ref[i] : qd_init[i] + (pb->qd_term_ref[i] - qd_init[i]) * t;
    }
  }
  for (int k = 0; k < N - 1; k++) p[pl->o_dt + k] = pb->dt ? pb->dt[k] : pb->T / (double)(N - 1);
  for (int i = 0; i < 6; i++) {
    p[pl->o_qmin + i] = pb->q_min[i]; p[pl->o_qmax + i] = pb->q_max[i];
    p[pl->o_qdmin + i] = pb->qd_min[i]; p[pl->o_qdmax + i] = pb->qd_max[i];
    p[pl->o_qinit + i] = q_init[i]; p[pl->o_qdinit + i] = qd_init[i];
    p[pl->o_qtmin + i] = pb->q_term_min[i]; p[pl->o_qtmax + i] = pb->q_term_max[i];
    p[pl->o_qdtmin + i] = pb->qd_term_min[i]; p[pl->o_qdtmax + i] = pb->qd_term_max[i];
  }
  for (int i = 0; i < 12; i++) p[pl->o_QN + i] = pb->QN[i];
  p[pl->o_mu] = pb->mu;
  p[pl->o_lleg] = pb->l_leg_max;
  p[pl->o_fmax] = pb->f_max;
  p[pl->o_mass] = pb->mass;
  for (int i = 0; i < 3; i++) { p[pl->o_Ib + i] = pb->Ib[i]; p[pl->o_Ibinv + i] = pb->Ib_inv[i]; }
  if (x0) {
    memcpy(x0, p, sizeof(double) * 12 * N);
    for (int k = 0; k < N - 1; k++) {
      double *U = x0 + 12 * N + 24 * k;
      for (int l = 0; l < 4; l++)
        for (int a = 0; a < 3; a++) {
          U[3 * l + a] = p[12 * k + a] + pb->c_ref[3 * l + a];
          U[12 + 3 * l + a] = 0.0;
        }
    }
  }
}
