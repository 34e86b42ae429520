/*
 * casadi_abi.c -- host C: the CasADi-generated-function ABI (include/casadi_symbols.h) on top of
 * the batched C API (include/landing_b200.h).  One scenario per call: B = 1, host buffers.
 * Replaces the symbol set of optimizations/landing/codegen_casadi/landingCtrller_IPOPT.c
 * (template :10916-10992).  Re-entrant: every memory object handed out by F_checkout owns a CUDA stream and staging
 * buffers (see g_mem below); incref/decref own the lifetime of the CUDA contexts.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/casadi_symbols.h"
#include "../../include/landing_b200.h"

/* Memory objects ("mem" of the CasADi ABI, mem.h:49-59).  CasADi hands every thread that evaluates a function its own
 * memory object: checkout() under CasADi's mutex, F(..., mem), release() (function_internal.cpp:721-733), and F itself
 * may be entered concurrently.  Here a memory object is one landing_ctx = one CUDA stream + its own staging buffers,
 * created lazily the first time the slot is used, so concurrent calls with different mem run concurrently on the GPU.
 * The reference's generated code ignores mem (always 0, no state); a caller that does the same and never checks out still
 * works: every slot carries its own lock, which is uncontended in the checkout/release protocol. */
#define MAX_MEM 64
static pthread_mutex_t g_mtx = PTHREAD_MUTEX_INITIALIZER; /* slot table, reference count */
static struct {
  landing_ctx *ctx;
  pthread_mutex_t lock;
  int busy; /* checked out */
  int init;
} g_mem[MAX_MEM];
static int g_refs = 0;

static int knots(void) {
  const char *s = getenv("LANDING_B200_KNOTS");
  int n = s ? atoi(s) : 21;
  return n >= 3 ? n : 21;
}

static void slot_init_locked(int i) {
  if (!g_mem[i].init) {
    pthread_mutex_init(&g_mem[i].lock, NULL);
    g_mem[i].init = 1;
  }
}

static int mem_checkout(void) {
  int m = 0;
  pthread_mutex_lock(&g_mtx);
  for (int i = 0; i < MAX_MEM; i++)
    if (!g_mem[i].busy) { m = i; break; }
  slot_init_locked(m);
  g_mem[m].busy = 1; /* (all MAX_MEM slots busy: slot 0 is shared, its lock serialises) */
  pthread_mutex_unlock(&g_mtx);
  return m;
}
static void mem_release(int m) {
  if (m < 0 || m >= MAX_MEM) return;
  pthread_mutex_lock(&g_mtx);
  g_mem[m].busy = 0;
  pthread_mutex_unlock(&g_mtx);
}

static void ref_inc(void) {
  pthread_mutex_lock(&g_mtx);
  g_refs++;
  pthread_mutex_unlock(&g_mtx);
}
static void ref_dec(void) {
  pthread_mutex_lock(&g_mtx);
  if (--g_refs <= 0) { /* last function object gone: tear the CUDA contexts down (external.cpp:114) */
    for (int i = 0; i < MAX_MEM; i++)
      if (g_mem[i].ctx) {
        landing_destroy(g_mem[i].ctx);
        g_mem[i].ctx = NULL;
      }
    g_refs = 0;
  }
  pthread_mutex_unlock(&g_mtx);
}

static int run(const landing_eval_io *io, int mem) {
  int rc = 1;
  if (mem < 0 || mem >= MAX_MEM) mem = 0;
  pthread_mutex_lock(&g_mtx);
  slot_init_locked(mem);
  pthread_mutex_unlock(&g_mtx);
  pthread_mutex_lock(&g_mem[mem].lock);
  if (!g_mem[mem].ctx) {
    const char *d = getenv("LANDING_B200_DEVICE");
    if (landing_create(knots(), d ? atoi(d) : 0, &g_mem[mem].ctx) != LANDING_OK) {
      fprintf(stderr, "landing_b200: %s\n", landing_last_error());
      g_mem[mem].ctx = NULL;
    }
  }
  if (g_mem[mem].ctx) {
    rc = landing_eval_batch(g_mem[mem].ctx, 1, LANDING_HOST, LANDING_AOS, io);
    if (rc) fprintf(stderr, "landing_b200: %s\n", landing_last_error());
  }
  pthread_mutex_unlock(&g_mem[mem].lock);
  return rc;
}

/* sparsity codes of landing_sparsity_for */
enum { SP_JAC = 0, SP_HESS = 1, SP_X = 2, SP_P = 3, SP_ONE = 4, SP_G = 5 };

#define BOILERPLATE(F, NIN, NOUT)                                                          \
  int F##_alloc_mem(void) { return 0; }                                                    \
  int F##_init_mem(int mem) { (void)mem; return 0; }                                       \
  void F##_free_mem(int mem) { (void)mem; }                                                \
  int F##_checkout(void) { return mem_checkout(); }                                        \
  void F##_release(int mem) { mem_release(mem); }                                          \
  void F##_incref(void) { ref_inc(); }                                                     \
  void F##_decref(void) { ref_dec(); }                                                     \
  long long F##_n_in(void) { return NIN; }                                                 \
  long long F##_n_out(void) { return NOUT; }                                               \
  double F##_default_in(long long i) { (void)i; return 0; }                                \
  const char *F##_name_in(long long i) { return (i >= 0 && i < NIN) ? F##_in_names[i] : 0; }   \
  const char *F##_name_out(long long i) { return (i >= 0 && i < NOUT) ? F##_out_names[i] : 0; } \
  const long long *F##_sparsity_in(long long i) {                                          \
    return (i >= 0 && i < NIN) ? landing_sparsity_for(knots(), F##_in_sp[i]) : 0;          \
  }                                                                                        \
  const long long *F##_sparsity_out(long long i) {                                         \
    return (i >= 0 && i < NOUT) ? landing_sparsity_for(knots(), F##_out_sp[i]) : 0;        \
  }                                                                                        \
  int F##_work(long long *sz_arg, long long *sz_res, long long *sz_iw, long long *sz_w) {  \
    if (sz_arg) *sz_arg = NIN;                                                             \
    if (sz_res) *sz_res = NOUT;                                                            \
    if (sz_iw) *sz_iw = 0;                                                                 \
    if (sz_w) *sz_w = 0;                                                                   \
    return 0;                                                                              \
  }

/* nlp: (x,p) -> (f,g) */
static const char *nlp_in_names[] = {"x", "p"}, *nlp_out_names[] = {"f", "g"};
static const int nlp_in_sp[] = {SP_X, SP_P}, nlp_out_sp[] = {SP_ONE, SP_G};
BOILERPLATE(nlp, 2, 2)
int nlp(const double **arg, double **res, long long *iw, double *w, int mem) {
  (void)iw; (void)w;
  landing_eval_io io;
  memset(&io, 0, sizeof io);
  io.x = arg[0]; io.p = arg[1];
  io.f = res[0]; io.g = res[1];
  if (!io.f && !io.g) return 0;
  return run(&io, mem);
}

static const char *nlp_f_in_names[] = {"x", "p"}, *nlp_f_out_names[] = {"f"};
static const int nlp_f_in_sp[] = {SP_X, SP_P}, nlp_f_out_sp[] = {SP_ONE};
BOILERPLATE(nlp_f, 2, 1)
int nlp_f(const double **arg, double **res, long long *iw, double *w, int mem) {
  (void)iw; (void)w;
  landing_eval_io io;
  memset(&io, 0, sizeof io);
  io.x = arg[0]; io.p = arg[1];
  io.f = res[0];
  if (!io.f) return 0;
  return run(&io, mem);
}

static const char *nlp_g_in_names[] = {"x", "p"}, *nlp_g_out_names[] = {"g"};
static const int nlp_g_in_sp[] = {SP_X, SP_P}, nlp_g_out_sp[] = {SP_G};
BOILERPLATE(nlp_g, 2, 1)
int nlp_g(const double **arg, double **res, long long *iw, double *w, int mem) {
  (void)iw; (void)w;
  landing_eval_io io;
  memset(&io, 0, sizeof io);
  io.x = arg[0]; io.p = arg[1];
  io.g = res[0];
  if (!io.g) return 0;
  return run(&io, mem);
}

static const char *nlp_grad_in_names[] = {"x", "p", "lam_f", "lam_g"},
                  *nlp_grad_out_names[] = {"f", "g", "grad_gamma_x", "grad_gamma_p"};
static const int nlp_grad_in_sp[] = {SP_X, SP_P, SP_ONE, SP_G}, nlp_grad_out_sp[] = {SP_ONE, SP_G, SP_X, SP_P};
BOILERPLATE(nlp_grad, 4, 4)
int nlp_grad(const double **arg, double **res, long long *iw, double *w, int mem) {
  (void)iw; (void)w;
  landing_eval_io io;
  memset(&io, 0, sizeof io);
  io.x = arg[0]; io.p = arg[1]; io.lam_f = arg[2]; io.lam_g = arg[3];
  io.f = res[0]; io.g = res[1]; io.grad_x = res[2]; io.grad_p = res[3];
  if (!io.f && !io.g && !io.grad_x && !io.grad_p) return 0;
  return run(&io, mem);
}

static const char *nlp_grad_f_in_names[] = {"x", "p"}, *nlp_grad_f_out_names[] = {"f", "grad_f_x"};
static const int nlp_grad_f_in_sp[] = {SP_X, SP_P}, nlp_grad_f_out_sp[] = {SP_ONE, SP_X};
BOILERPLATE(nlp_grad_f, 2, 2)
int nlp_grad_f(const double **arg, double **res, long long *iw, double *w, int mem) {
  (void)iw; (void)w;
  landing_eval_io io;
  memset(&io, 0, sizeof io);
  io.x = arg[0]; io.p = arg[1];
  io.f = res[0]; io.grad_f = res[1];
  if (!io.f && !io.grad_f) return 0;
  return run(&io, mem);
}

static const char *nlp_hess_l_in_names[] = {"x", "p", "lam_f", "lam_g"},
                  *nlp_hess_l_out_names[] = {"hess_gamma_x_x"};
static const int nlp_hess_l_in_sp[] = {SP_X, SP_P, SP_ONE, SP_G}, nlp_hess_l_out_sp[] = {SP_HESS};
BOILERPLATE(nlp_hess_l, 4, 1)
int nlp_hess_l(const double **arg, double **res, long long *iw, double *w, int mem) {
  (void)iw; (void)w;
  landing_eval_io io;
  memset(&io, 0, sizeof io);
  io.x = arg[0]; io.p = arg[1]; io.lam_f = arg[2]; io.lam_g = arg[3];
  io.hess = res[0];
  if (!io.hess) return 0;
  return run(&io, mem);
}

static const char *nlp_jac_g_in_names[] = {"x", "p"}, *nlp_jac_g_out_names[] = {"g", "jac_g_x"};
static const int nlp_jac_g_in_sp[] = {SP_X, SP_P}, nlp_jac_g_out_sp[] = {SP_G, SP_JAC};
BOILERPLATE(nlp_jac_g, 2, 2)
int nlp_jac_g(const double **arg, double **res, long long *iw, double *w, int mem) {
  (void)iw; (void)w;
  landing_eval_io io;
  memset(&io, 0, sizeof io);
  io.x = arg[0]; io.p = arg[1];
  io.g = res[0]; io.jac = res[1];
  if (!io.g && !io.jac) return 0;
  return run(&io, mem);
}
