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
 * the condensed KKT system).  "IPOPT substitute": PARITY UNPINNED against real IPOPT output for the
 * N=21 problem (no stored solutions, no binary); this file is the oracle for the CUDA solver,
 * which implements the same algorithm step by step.
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
} ip_options;

typedef struct ip_result {
  int status; /* 0 converged, 1 max_iter, 2 line-search failure, 3 NaN, 4 factorisation failure */
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
