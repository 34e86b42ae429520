// srb_knot.cuh -- one knot of the SRB contact-implicit landing NLP, FP64, sm_100a.
//
// Computes, for knot k, the 104 (80 for the last knot) constraint rows of
//   /root/reference/optimizations/landing/generate_solver/generate_landingCtrller_IPOPT.m:106-169
// and their first / multiplier-weighted second derivatives, i.e. the per-knot slice of the
// reference's generated functions nlp_g (landingCtrller_IPOPT.c:11161), nlp_jac_g (:94014)
// and nlp_hess_l (:53527).
//
// B200 formulation (not the reference's AD trace): R = Rz*Ry*Rx (rpyToRotMat.m:2) is built once
// together with dR/dpitch and d2R/dpitch2; roll derivatives are right-multiplications by the x
// generator (column shuffles), yaw derivatives left-multiplications by the z generator (row
// shuffles), so every derivative vector is a 3x3 mat-vec plus a cross product.  The Euler-rate row
// uses the closed form Binv(th)*R(th) = [1 sf*tt cf*tt; 0 cf -sf; 0 sf/ct cf/ct]  (Binv.m:13-17),
// which does not depend on yaw.  Entries that the reference's CasADi pattern keeps but that are
// analytically zero (e.g. d(rpy row)/dyaw) are emitted as exact zeros.
//
// Output goes through a Sink so the same code serves the batched ABI kernels (scatter into the
// CCS value arrays) and the interior-point kernel (stage condensing).  After full unrolling the
// emission index `e` of every entry is a compile-time constant.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SRB_HD __host__ __device__ __forceinline__
#else
#define SRB_HD inline
#endif

namespace srb {

constexpr int NJ_INT = 385, NJ_LAST = 313;  // Jacobian nz per knot
constexpr int NH_INT = 189, NH_LAST = 177;  // upper-tri Hessian nz per knot
constexpr double FRIC = 0.71;               // generate_landingCtrller_IPOPT.m:160-163
constexpr double GRAV_Z = -9.81;            // get_robot_model.m:140

// knot-local variable ids: 0-11 X_k, 12-23 c_k, 24-35 f_k, 36-47 X_{k+1}, 48-59 c_{k+1}
struct Knot {
  double X[12], c[12], f[12], Xn[12], cn[12];
  double h, mu, mass, Ib[3], Ibinv[3];
  double csv[4];  // contact schedule of this knot (0 / 1 per leg); read by the SCHED instantiation only
};

// hip offsets FR, FL, BR, BL: get_robot_params.m:90-91 (hip_z = 0)
SRB_HD constexpr double hip_x(int l) { return l < 2 ? 0.19 : -0.19; }
SRB_HD constexpr double hip_y(int l) { return (l & 1) ? 0.1 : -0.1; }

template <bool LAST> struct Rows {
  static constexpr int leg(int l) { return 16 + (LAST ? 6 : 12) * l; }
  static constexpr int kin(int l) { return leg(l) + (LAST ? 2 : 8); }
  static constexpr int fric = LAST ? 40 : 64;
  static constexpr int state = LAST ? 56 : 80;
  static constexpr int count = LAST ? 80 : 104;
};

struct NoLam {
  SRB_HD double operator()(int) const { return 0.0; }
};

SRB_HD void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
SRB_HD double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// o = A v, A row-major
SRB_HD void mv(const double A[9], const double v[3], double o[3]) {
#pragma unroll
  for (int i = 0; i < 3; i++) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
// o = A^T v
SRB_HD void mtv(const double A[9], const double v[3], double o[3]) {
#pragma unroll
  for (int i = 0; i < 3; i++) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
// o[m][b] = (R^T (v x e_b))_m
SRB_HD void rt_cross(const double R[9], const double v[3], double o[3][3]) {
#pragma unroll
  for (int m = 0; m < 3; m++) {
    o[m][0] = R[3 + m] * v[2] - R[6 + m] * v[1];
    o[m][1] = R[6 + m] * v[0] - R[m] * v[2];
    o[m][2] = R[m] * v[1] - R[3 + m] * v[0];
  }
}
SRB_HD void zcross(const double v[3], double o[3]) {  // z_hat x v
  o[0] = -v[1];
  o[1] = v[0];
  o[2] = 0.0;
}

// SCHED: the fixed-contact-schedule formulation (quadruped_SRBM_NLP.m:145-158) in the same row layout: row leg+0 is
// cs c_z (an equality), rows leg+2..4 are cs (c+ - c) (equalities, linear), rows leg+1 and leg+5..7 are unused (zero);
// the f_z rows 12..15 are unchanged (their bound becomes cs f_max).
// KN: Knot (values; the batched kernels gather them from strided views) or KnotRef (pointers into one scenario's x;
// the solver kernel -- no copy of the 60 values through a local struct).
struct KnotRef {
  const double *X, *c, *f, *Xn, *cn;
  double h, mu, mass, Ib[3], Ibinv[3];
  double csv[4];
};
template <bool LAST, bool WG, bool WJ, bool WH, class Sink, class Lam, bool SCHED = false, class KN = Knot>
SRB_HD void knot_eval(const KN& kn, Sink& out, const Lam& lam) {
  using RW = Rows<LAST>;
  const double h = kn.h, mu = kn.mu;
  const double* r = kn.X;
  const double* om = kn.X + 6;
  const double* v = kn.X + 9;

  double sf, cf, st, ct, sp, cp;
#if defined(__CUDA_ARCH__)
  sincos(kn.X[3], &sf, &cf);
  sincos(kn.X[4], &st, &ct);
  sincos(kn.X[5], &sp, &cp);
#else
  sf = sin(kn.X[3]); cf = cos(kn.X[3]);
  st = sin(kn.X[4]); ct = cos(kn.X[4]);
  sp = sin(kn.X[5]); cp = cos(kn.X[5]);
#endif
  const double ic = 1.0 / ct, tt = st * ic;

  // R (row-major) and its pitch derivatives
  const double R[9] = {cp * ct, -cf * sp + sf * cp * st, sf * sp + cf * cp * st,
                       sp * ct, cf * cp + sf * sp * st,  -sf * cp + cf * sp * st,
                       -st,     sf * ct,                 cf * ct};
  const double Rt[9] = {-cp * st, sf * cp * ct, cf * cp * ct,
                        -sp * st, sf * sp * ct, cf * sp * ct,
                        -ct,      -sf * st,     -cf * st};

  // wrench about the body origin
  double F[3] = {0, 0, 0}, tau[3] = {0, 0, 0}, arm[4][3];
#pragma unroll
  for (int l = 0; l < 4; l++) {
    double t[3];
#pragma unroll
    for (int a = 0; a < 3; a++) arm[l][a] = kn.c[3 * l + a] - r[a];
    cross3(arm[l], kn.f + 3 * l, t);
#pragma unroll
    for (int a = 0; a < 3; a++) { tau[a] += t[a]; F[a] += kn.f[3 * l + a]; }
  }
  double tb[3];  // torque in body frame
  mtv(R, tau, tb);
  const double* Ib = kn.Ib;
  const double w[3] = {om[1] * (Ib[2] * om[2]) - om[2] * (Ib[1] * om[1]),
                       om[2] * (Ib[0] * om[0]) - om[0] * (Ib[2] * om[2]),
                       om[0] * (Ib[1] * om[1]) - om[1] * (Ib[0] * om[0])};
  const double kap[3] = {-h * kn.Ibinv[0], -h * kn.Ibinv[1], -h * kn.Ibinv[2]};
  // Euler rates e = E(roll,pitch) * om
  const double ea = sf * om[1] + cf * om[2], eb = cf * om[1] - sf * om[2];
  const double e[3] = {om[0] + tt * ea, eb, ea * ic};

  double hw[4][3], prel[4][3];
#pragma unroll
  for (int l = 0; l < 4; l++)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      hw[l][a] = hip_x(l) * R[3 * a] + hip_y(l) * R[3 * a + 1];
      prel[l][a] = arm[l][a] - hw[l][a];
    }

  if (WG) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      out.g(a, kn.Xn[a] - r[a] - v[a] * h);
      out.g(3 + a, kn.Xn[3 + a] - kn.X[3 + a] - e[a] * h);
      out.g(6 + a, kn.Xn[9 + a] - v[a] - (F[a] / kn.mass + (a == 2 ? GRAV_Z : 0.0)) * h);
      out.g(9 + a, kn.Xn[6 + a] - om[a] + kap[a] * (tb[a] - w[a]));
    }
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const double fz = kn.f[3 * l + 2], cz = kn.c[3 * l + 2];
      out.g(12 + l, fz);
      out.g(RW::leg(l), SCHED ? kn.csv[l] * cz : cz);
      out.g(RW::leg(l) + 1, SCHED ? 0.0 : fz * cz);
      if (!LAST) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const double ns = (SCHED ? kn.csv[l] : fz) * (kn.cn[3 * l + a] - kn.c[3 * l + a]);
          out.g(RW::leg(l) + 2 + a, ns);
          out.g(RW::leg(l) + 5 + a, SCHED ? 0.0 : ns);
        }
      }
      out.g(RW::kin(l), prel[l][0]);
      out.g(RW::kin(l) + 1, prel[l][1]);
      out.g(RW::kin(l) + 2, prel[l][2] + 0.05);
      out.g(RW::kin(l) + 3, dot3(prel[l], prel[l]));
      out.g(RW::fric + l, kn.f[3 * l] - FRIC * mu * fz);
      out.g(RW::fric + 4 + l, -FRIC * mu * fz - kn.f[3 * l]);
      out.g(RW::fric + 8 + l, kn.f[3 * l + 1] - FRIC * mu * fz);
      out.g(RW::fric + 12 + l, -FRIC * mu * fz - kn.f[3 * l + 1]);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      out.g(RW::state + i, kn.X[i]);
      out.g(RW::state + 6 + i, kn.X[i]);
      out.g(RW::state + 12 + i, kn.X[6 + i]);
      out.g(RW::state + 18 + i, kn.X[6 + i]);
    }
  }
  if (!WJ && !WH) return;

  // d(R hip)/d(roll,pitch,yaw)
  double dh[4][3][3];
#pragma unroll
  for (int l = 0; l < 4; l++) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      dh[l][0][a] = hip_y(l) * R[3 * a + 2];
      dh[l][1][a] = hip_x(l) * Rt[3 * a] + hip_y(l) * Rt[3 * a + 1];
    }
    zcross(hw[l], dh[l][2]);
  }
  const double ic2 = ic * ic;

  if (WJ) {
    int n = 0;
    // rows 0-2: r+ - r - v h
#pragma unroll
    for (int a = 0; a < 3; a++) {
      out.j(n++, a, 36 + a, 1.0);
      out.j(n++, a, a, -1.0);
      out.j(n++, a, 9 + a, -h);
    }
    // rows 3-5: th+ - th - h E om
    const double de[3][3] = {{tt * eb, -ea, eb * ic}, {ea * ic2, 0.0, ea * st * ic2}, {0.0, 0.0, 0.0}};
    const double E[9] = {1.0, tt * sf, tt * cf, 0.0, cf, -sf, 0.0, sf * ic, cf * ic};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      out.j(n++, 3 + a, 39 + a, 1.0);
#pragma unroll
      for (int i = 0; i < 3; i++) out.j(n++, 3 + a, 3 + i, -(a == i ? 1.0 : 0.0) - h * de[i][a]);
#pragma unroll
      for (int jx = 0; jx < 3; jx++) out.j(n++, 3 + a, 6 + jx, -h * E[3 * a + jx]);
    }
    // rows 6-8: v+ - v - h (F/m + g)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      out.j(n++, 6 + a, 45 + a, 1.0);
      out.j(n++, 6 + a, 9 + a, -1.0);
#pragma unroll
      for (int l = 0; l < 4; l++) out.j(n++, 6 + a, 24 + 3 * l + a, -h / kn.mass);
    }
    // rows 9-11: om+ - om + kap (R^T tau - om x Ib om)
    {
      const double dw[3][3] = {{0.0, (Ib[2] - Ib[1]) * om[2], (Ib[2] - Ib[1]) * om[1]},
                               {(Ib[0] - Ib[2]) * om[2], 0.0, (Ib[0] - Ib[2]) * om[0]},
                               {(Ib[1] - Ib[0]) * om[1], (Ib[1] - Ib[0]) * om[0], 0.0}};
      double JF[3][3], dtb[3][3], Jc[4][3][3], Jf[4][3][3];
      rt_cross(R, F, JF);
      // d(R^T tau)/d roll = -x_hat x tb ; pitch: Rt^T tau ; yaw: R^T (tau_y, -tau_x, 0)
      dtb[0][0] = 0.0; dtb[0][1] = tb[2]; dtb[0][2] = -tb[1];
      mtv(Rt, tau, dtb[1]);
      const double tz[3] = {tau[1], -tau[0], 0.0};
      mtv(R, tz, dtb[2]);
#pragma unroll
      for (int l = 0; l < 4; l++) {
        rt_cross(R, kn.f + 3 * l, Jc[l]);  // d/dc = -R^T (f x e_b)
        rt_cross(R, arm[l], Jf[l]);
      }
#pragma unroll
      for (int m = 0; m < 3; m++) {
        out.j(n++, 9 + m, 42 + m, 1.0);
#pragma unroll
        for (int b = 0; b < 3; b++) out.j(n++, 9 + m, 6 + b, -(m == b ? 1.0 : 0.0) - kap[m] * dw[m][b]);
#pragma unroll
        for (int b = 0; b < 3; b++) out.j(n++, 9 + m, b, kap[m] * JF[m][b]);
#pragma unroll
        for (int i = 0; i < 3; i++) {
          if (m == 0 && i == 0) continue;
          out.j(n++, 9 + m, 3 + i, kap[m] * dtb[i][m]);
        }
#pragma unroll
        for (int l = 0; l < 4; l++)
#pragma unroll
          for (int b = 0; b < 3; b++) {
            out.j(n++, 9 + m, 12 + 3 * l + b, -kap[m] * Jc[l][m][b]);
            out.j(n++, 9 + m, 24 + 3 * l + b, kap[m] * Jf[l][m][b]);
          }
      }
    }
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const double fz = kn.f[3 * l + 2], cz = kn.c[3 * l + 2];
      const int L = RW::leg(l), K = RW::kin(l);
      out.j(n++, 12 + l, 24 + 3 * l + 2, 1.0);
      out.j(n++, L, 12 + 3 * l + 2, SCHED ? kn.csv[l] : 1.0);
      out.j(n++, L + 1, 12 + 3 * l + 2, SCHED ? 0.0 : fz);
      out.j(n++, L + 1, 24 + 3 * l + 2, SCHED ? 0.0 : cz);
      if (!LAST) {
#pragma unroll
        for (int rep = 0; rep < 2; rep++)
#pragma unroll
          for (int a = 0; a < 3; a++) {
            const int row = L + 2 + 3 * rep + a;
            const double wgt = SCHED ? (rep == 0 ? kn.csv[l] : 0.0) : fz;
            out.j(n++, row, 12 + 3 * l + a, -wgt);
            out.j(n++, row, 24 + 3 * l + 2, SCHED ? 0.0 : kn.cn[3 * l + a] - kn.c[3 * l + a]);
            out.j(n++, row, 48 + 3 * l + a, wgt);
          }
      }
#pragma unroll
      for (int a = 0; a < 3; a++) {
        out.j(n++, K + a, 12 + 3 * l + a, 1.0);
        out.j(n++, K + a, a, -1.0);
#pragma unroll
        for (int i = 0; i < 3; i++) {
          if (a == 2 && i == 2) continue;
          out.j(n++, K + a, 3 + i, -dh[l][i][a]);
        }
      }
#pragma unroll
      for (int a = 0; a < 3; a++) {
        out.j(n++, K + 3, 12 + 3 * l + a, 2.0 * prel[l][a]);
        out.j(n++, K + 3, a, -2.0 * prel[l][a]);
      }
#pragma unroll
      for (int i = 0; i < 3; i++) out.j(n++, K + 3, 3 + i, -2.0 * dot3(prel[l], dh[l][i]));
      out.j(n++, RW::fric + l, 24 + 3 * l, 1.0);
      out.j(n++, RW::fric + l, 24 + 3 * l + 2, -FRIC * mu);
      out.j(n++, RW::fric + 4 + l, 24 + 3 * l, -1.0);
      out.j(n++, RW::fric + 4 + l, 24 + 3 * l + 2, -FRIC * mu);
      out.j(n++, RW::fric + 8 + l, 24 + 3 * l + 1, 1.0);
      out.j(n++, RW::fric + 8 + l, 24 + 3 * l + 2, -FRIC * mu);
      out.j(n++, RW::fric + 12 + l, 24 + 3 * l + 1, -1.0);
      out.j(n++, RW::fric + 12 + l, 24 + 3 * l + 2, -FRIC * mu);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      out.j(n++, RW::state + i, i, 1.0);
      out.j(n++, RW::state + 6 + i, i, 1.0);
      out.j(n++, RW::state + 12 + i, 6 + i, 1.0);
      out.j(n++, RW::state + 18 + i, 6 + i, 1.0);
    }
  }

  if (WH) {
    int n = 0;
    const double Rtt[9] = {-R[0], sf * Rt[0], cf * Rt[0],
                           -R[3], sf * Rt[3], cf * Rt[3],
                           -R[6], sf * Rt[6], cf * Rt[6]};
    const double lk[3] = {lam(9) * kap[0], lam(10) * kap[1], lam(11) * kap[2]};
    const double lr[3] = {-h * lam(3), -h * lam(4), -h * lam(5)};
    // y = R lk and its angle derivatives
    double y[3], yi[3][3], yy[6][3];
    mv(R, lk, y);
    {
      const double s1[3] = {0.0, -lk[2], lk[1]};
      mv(R, s1, yi[0]);
      mv(Rt, lk, yi[1]);
      zcross(y, yi[2]);
      const double s2[3] = {0.0, -lk[1], -lk[2]};
      mv(R, s2, yy[0]);        // roll-roll
      mv(Rt, s1, yy[1]);       // roll-pitch
      zcross(yi[0], yy[2]);    // roll-yaw
      mv(Rtt, lk, yy[3]);      // pitch-pitch
      zcross(yi[1], yy[4]);    // pitch-yaw
      yy[5][0] = -y[0]; yy[5][1] = -y[1]; yy[5][2] = 0.0;  // yaw-yaw
    }
    double lpp[4], lns[4][3], lpa[4][3], slpp = 0.0;
#pragma unroll
    for (int l = 0; l < 4; l++) {
      lpp[l] = lam(RW::kin(l) + 3);
      slpp += lpp[l];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        lpa[l][a] = lam(RW::kin(l) + a);
        lns[l][a] = (LAST || SCHED) ? 0.0 : lam(RW::leg(l) + 2 + a) + lam(RW::leg(l) + 5 + a);
      }
    }
    // r_a - r_a
#pragma unroll
    for (int a = 0; a < 3; a++) out.h(n++, a, a, 2.0 * slpp);
    // r_a - theta_i
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double t[3];
      cross3(yi[i], F, t);
#pragma unroll
      for (int a = 0; a < 3; a++) {
        double s = t[a];
#pragma unroll
        for (int l = 0; l < 4; l++) s += 2.0 * lpp[l] * dh[l][i][a];
        out.h(n++, a, 3 + i, s);
      }
    }
    // theta_i - theta_j
    {
      const double ic3 = ic2 * ic;
      // lr . d2e : only roll/pitch pairs are non-zero
      const double dde[6] = {lr[0] * (-tt * ea) + lr[1] * (-eb) + lr[2] * (-ea * ic),
                             lr[0] * (eb * ic2) + lr[2] * (eb * st * ic2),
                             0.0,
                             lr[0] * (2.0 * ea * st * ic3) + lr[2] * (ea * (1.0 + st * st) * ic3),
                             0.0,
                             0.0};
      double s6[6];
#pragma unroll
      for (int q = 0; q < 6; q++) s6[q] = dde[q] + dot3(yy[q], tau);
#pragma unroll
      for (int l = 0; l < 4; l++) {
        double ddh[6][3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
          ddh[0][a] = -hip_y(l) * R[3 * a + 1];
          ddh[1][a] = hip_y(l) * Rt[3 * a + 2];
          ddh[3][a] = hip_x(l) * Rtt[3 * a] + hip_y(l) * Rtt[3 * a + 1];
        }
        zcross(dh[l][0], ddh[2]);
        zcross(dh[l][1], ddh[4]);
        ddh[5][0] = -hw[l][0]; ddh[5][1] = -hw[l][1]; ddh[5][2] = 0.0;
        int q = 0;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int jx = i; jx < 3; jx++) {
            s6[q] += -dot3(lpa[l], ddh[q]) +
                     lpp[l] * (2.0 * dot3(dh[l][i], dh[l][jx]) - 2.0 * dot3(prel[l], ddh[q]));
            q++;
          }
      }
      int q = 0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int jx = i; jx < 3; jx++) out.h(n++, 3 + i, 3 + jx, s6[q++]);
    }
    // theta_i - omega_j : lr^T dE/dtheta_i
    {
      const double dEf[9] = {0.0, tt * cf, -tt * sf, 0.0, -sf, -cf, 0.0, cf * ic, -sf * ic};
      const double dEt[9] = {0.0, sf * ic2, cf * ic2, 0.0, 0.0, 0.0, 0.0, sf * st * ic2, cf * st * ic2};
#pragma unroll
      for (int jx = 1; jx < 3; jx++)
        out.h(n++, 3, 6 + jx, lr[0] * dEf[jx] + lr[1] * dEf[3 + jx] + lr[2] * dEf[6 + jx]);
#pragma unroll
      for (int jx = 0; jx < 3; jx++)
        out.h(n++, 4, 6 + jx, lr[0] * dEt[jx] + lr[1] * dEt[3 + jx] + lr[2] * dEt[6 + jx]);
#pragma unroll
      for (int jx = 0; jx < 3; jx++) out.h(n++, 5, 6 + jx, 0.0);
    }
    // omega - omega
    out.h(n++, 6, 7, -lk[2] * (Ib[1] - Ib[0]));
    out.h(n++, 6, 8, -lk[1] * (Ib[0] - Ib[2]));
    out.h(n++, 7, 8, -lk[0] * (Ib[2] - Ib[1]));
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int vc = 12 + 3 * l, vf = 24 + 3 * l;
#pragma unroll
      for (int a = 0; a < 3; a++) out.h(n++, a, vc + a, -2.0 * lpp[l]);
      out.h(n++, 0, vf + 1, -y[2]);
      out.h(n++, 0, vf + 2, y[1]);
      out.h(n++, 1, vf + 0, y[2]);
      out.h(n++, 1, vf + 2, -y[0]);
      out.h(n++, 2, vf + 0, -y[1]);
      out.h(n++, 2, vf + 1, y[0]);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        double t[3];
        cross3(kn.f + 3 * l, yi[i], t);
#pragma unroll
        for (int b = 0; b < 3; b++) out.h(n++, 3 + i, vc + b, t[b] - 2.0 * lpp[l] * dh[l][i][b]);
        cross3(yi[i], arm[l], t);
#pragma unroll
        for (int b = 0; b < 3; b++) out.h(n++, 3 + i, vf + b, t[b]);
      }
#pragma unroll
      for (int a = 0; a < 3; a++) out.h(n++, vc + a, vc + a, 2.0 * lpp[l]);
      out.h(n++, vc + 0, vf + 1, y[2]);
      out.h(n++, vc + 0, vf + 2, -y[1] - lns[l][0]);
      out.h(n++, vc + 1, vf + 0, -y[2]);
      out.h(n++, vc + 1, vf + 2, y[0] - lns[l][1]);
      out.h(n++, vc + 2, vf + 0, y[1]);
      out.h(n++, vc + 2, vf + 1, -y[0]);
      out.h(n++, vc + 2, vf + 2, (SCHED ? 0.0 : lam(RW::leg(l) + 1)) - lns[l][2]);
      if (!LAST) {
#pragma unroll
        for (int a = 0; a < 3; a++) out.h(n++, vf + 2, 48 + 3 * l + a, lns[l][a]);
      }
    }
  }
}

// ---- dealing the entries of one knot to several threads (evaluation kernels, solver): owner part of a row / entry.
// Per-leg rows and entries w.r.t. leg variables go to their leg, the rest round-robin.
template <bool LAST> SRB_HD constexpr int leg_of_row(int row) {  // -1: not a per-leg row
  using RW = Rows<LAST>;
  if (row >= 12 && row < 16) return row - 12;
  if (row >= 16 && row < RW::fric) return (row - 16) / (LAST ? 6 : 12);
  if (row >= RW::fric && row < RW::state) return (row - RW::fric) & 3;
  return -1;
}
SRB_HD constexpr int leg_of_var(int var) {  // X 0-11 | c 12-23 | f 24-35 | X+ 36-47 | c+ 48-59
  if (var >= 12 && var < 36) return ((var - 12) % 12) / 3;
  if (var >= 48) return (var - 48) / 3;
  return -1;
}
template <bool LAST> SRB_HD constexpr int g_owner(int row) {
  const int l = leg_of_row<LAST>(row);
  return l >= 0 ? l : (row & 3);
}
template <bool LAST> SRB_HD constexpr int j_owner(int row, int var) {
  const int lv = leg_of_var(var);
  return lv >= 0 ? lv : g_owner<LAST>(row);
}
SRB_HD constexpr int h_owner(int e, int va, int vb) {
  const int la = leg_of_var(va), lb = leg_of_var(vb);
  return la >= 0 ? la : (lb >= 0 ? lb : (e & 3));
}

}  // namespace srb
