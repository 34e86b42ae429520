/*
 * casadi_symbols.h -- the CasADi-generated-C symbol set exported by liblanding_b200.so so that it
 * drops in for optimizations/landing/codegen_casadi/landingCtrller_IPOPT.so (path baked into the
 * reference's landingCtrller_IPOPT.casadi; loaded with dlopen by importer_internal.cpp:231 and
 * bound by name in external.cpp:63-111,325-362).
 *
 * Each declaration replaces the identically named symbol of
 *   optimizations/landing/codegen_casadi/landingCtrller_IPOPT.c
 * (nlp :67/:10916-10992, nlp_f :10995, nlp_g :11161, nlp_grad :22015, nlp_grad_f :52602,
 *  nlp_hess_l :53527, nlp_jac_g :94014).  casadi_real = double (:19), casadi_int = long long (:23).
 * The number of knots N is not part of this ABI: it is read once from the environment variable
 * LANDING_B200_KNOTS (default 21, the reference's generated size); device from LANDING_B200_DEVICE.
 */
#ifndef LANDING_CASADI_SYMBOLS_H
#define LANDING_CASADI_SYMBOLS_H

#ifdef __cplusplus
extern "C" {
#endif

#define LANDING_CASADI_DECLARE(F)                                                         \
  int F(const double **arg, double **res, long long *iw, double *w, int mem);            \
  int F##_alloc_mem(void);                                                                \
  int F##_init_mem(int mem);                                                              \
  void F##_free_mem(int mem);                                                             \
  int F##_checkout(void);                                                                 \
  void F##_release(int mem);                                                              \
  void F##_incref(void);                                                                  \
  void F##_decref(void);                                                                  \
  long long F##_n_in(void);                                                               \
  long long F##_n_out(void);                                                              \
  double F##_default_in(long long i);                                                     \
  const char *F##_name_in(long long i);                                                   \
  const char *F##_name_out(long long i);                                                  \
  const long long *F##_sparsity_in(long long i);                                          \
  const long long *F##_sparsity_out(long long i);                                         \
  int F##_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w);

LANDING_CASADI_DECLARE(nlp)        /* (x,p) -> (f,g)                                  :67    */
LANDING_CASADI_DECLARE(nlp_f)      /* (x,p) -> (f)                                    :10995 */
LANDING_CASADI_DECLARE(nlp_g)      /* (x,p) -> (g)                                    :11161 */
LANDING_CASADI_DECLARE(nlp_grad)   /* (x,p,lam_f,lam_g) -> (f,g,grad_gamma_x,grad_gamma_p) :22015 */
LANDING_CASADI_DECLARE(nlp_grad_f) /* (x,p) -> (f,grad_f_x)                           :52602 */
LANDING_CASADI_DECLARE(nlp_hess_l) /* (x,p,lam_f,lam_g) -> (hess_gamma_x_x)           :53527 */
LANDING_CASADI_DECLARE(nlp_jac_g)  /* (x,p) -> (g,jac_g_x)                            :94014 */

#ifdef __cplusplus
}
#endif
#endif
