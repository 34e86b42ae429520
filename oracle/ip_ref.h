/*
 * oracle/ip_ref.h -- TEST INFRASTRUCTURE ONLY (CPU oracle / CPU baseline).
 *
 * CPU restatement of the interior-point solve that the reference delegates to IPOPT
 * (un-vendored third party: IPOPT + MUMPS/MA57 inside CasADi 3.5.5's libcasadi_nlpsol_ipopt.so,
 * listed in /root/reference/.MISSING_LARGE_BLOBS:14; version string not recorded in the repo).
 * The reference only fixes the option set (generate_landingCtrller_IPOPT.m:232-263) and the call
 * site (:264-277,:314-327); the algorithm restated here is the published primal-dual
 * filter line-search interior-point method (Waechter & Biegler, Math. Prog. 106, 2006) with the
 * monotone barrier update, specialised to this problem's stage structure (Riccati recursion on
 * the condensed KKT system).  "IPOPT substitute": for the N=21 landingCtrller_IPOPT problem the reference keeps no
 * IPOPT output (no stored solutions, no binary), so converged trajectories of THAT problem are parity unpinned.
 * The restatement itself is pinned -- softly, at the 1e-3 level IPOPT's tol = 1e-4 allows -- against the only real
 * IPOPT outputs in the repository: on the N=41 "CCC" variant (run_Qf, kin_box below) it lands on the stored solution
 * for 25 of the 43 committed runs (same cost to 0.2 %, identical touchdown knots, terminal state within 1e-3, vertical
 * GRFs within ~0.5 N) and in a different local solution of the non-convex problem for the rest
 * (tests/test_reference_solutions.py).  This file is the oracle for the CUDA solver, which implements the same
 * algorithm step by step.
 */
#ifndef IP_REF_H
#define IP_REF_H
#include "srb_ref.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ip_options {
  int max_iter;
  double tol, constr_viol_tol, dual_inf_tol, compl_inf_tol;
  double mu_init, bound_push, bound_frac, bound_relax_factor;
  int max_soc;
  int verbose;
  double jam_alpha; /* watchdog: accepted primal step below this ... */
  int jam_iters;    /* ... for this many consecutive iterations -> re-centre (0 = off) */
  int max_restarts; /* re-centrings allowed per scenario */
  /* Problem-variant data that the reference's parameter vector p (landingCtrller_IPOPT layout) does not carry:
   * running GRF cost sum_k sum_legs (Qf . f^2) dt_k and the kinematic box half-widths of the "CCC" variant
   * (generate_quadruped_SRBM_CCC.m:80-91,169-176; analysis/eval_SRBM_CCC.m:49-52).  Defaults {0,0,0} and
   * {0.15, 0.15, 0.30} give the IPOPT variant (generate_landingCtrller_IPOPT.m:83-85,150-155). */
  double run_Qf[3];
  double kin_box[3];
  /* Fixed-contact-schedule formulation (BASELINE configs[0]; quadruped_SRBM_NLP.m:84-176): formulation = 1 replaces the
   * complementarity / no-slip inequality rows by  f_z <= cs f_max (:148),  cs c_z = 0 (:154),  cs (c+ - c) = 0 (:155-158),
   * drops the terminal rows (:105-108 are commented out) and adds the running state cost sum_k (X_k - Xref_k)' QX
   * (X_k - Xref_k) dt_k (:85-92; Qc = 0 in the reference's parameter set :213 and is not implemented).  cs[4 (N-1)]:
   * cs[4 k + leg] in {0, 1}.  The extra equality rows are handled as IPOPT handles equality rows of a rank-deficient
   * Jacobian (they vanish identically where cs = 0): dual regularisation delta_c, i.e. sigma = 1 / delta_c in the
   * condensed stage matrix and y+ = y + (J dx + c) / delta_c. */
  int formulation;
  const int *cs;
  double QX[12];
  double delta_c; /* 1e-7 */
  /* IPOPT's "acceptable" termination, which the reference sets (generate_landingCtrller_IPOPT.m:233,235: acceptable_tol
   * 1e-4, acceptable_iter 5): stop after acceptable_iter consecutive iterates whose scaled optimality error is within
   * acceptable_tol and whose unscaled errors are within IPOPT's default acceptable levels (constraint violation 1e-2,
   * dual infeasibility 1e10, complementarity 1e-2).  Status 5.  acceptable_iter = 0 switches it off. */
  double acceptable_tol;
  int acceptable_iter;
  double restart_mu; /* barrier parameter a re-centring restarts from (<= 0: mu_init); landing_options.restart_mu */
} ip_options;

typedef struct ip_result {
  int status; /* 0 converged, 1 max_iter, 2 line-search failure, 3 NaN, 4 factorisation failure, 5 acceptable level */
  int iters;
  int n_factor; /* Riccati factorisations (incl. inertia-correction retries) */
  double f, viol, dual_inf, compl_inf, mu;
  int restarts; /* re-centrings (failed line searches / jamming watchdog) */
} ip_result;

void ip_options_default(ip_options *o);

/* one NLP: parameters p, initial guess x0 -> x_out[nx], lam_g_out[m] (optional) */
int ip_solve(const srb_plan *pl, const double *p, const double *x0, const ip_options *opt,
             double *x_out, double *lam_g_out, ip_result *res);

/* sweep: drops[B][12] -> x_out[B][nx], res[B]; OpenMP over scenarios with nthreads threads */
int ip_solve_batch(int N, int B, const double *drops, const srb_problem *pb, const ip_options *opt,
                   double *x_out, ip_result *res, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
