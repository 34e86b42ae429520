/*
 * oracle/ref_timing.c -- TEST INFRASTRUCTURE ONLY (CPU baseline of the evaluation rows).
 *
 * Times the REFERENCE's own compiled functions (oracle/_ref/landingCtrller_IPOPT.so = gcc -O3 of
 * optimizations/landing/codegen_casadi/landingCtrller_IPOPT.c, N = 21) on the host cores: B calls of one generated
 * function F(arg, res, iw, w, mem) (landingCtrller_IPOPT.c:10916-10992), one scenario each, OpenMP over scenarios --
 * what a sweep over the reference's evaluation path costs on a CPU.  Used by tools/bench_eval.py next to the GPU
 * kernels' numbers (cpu_baseline kind "reference"); nothing in the product links this file.
 */
#include <dlfcn.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int (*casadi_fn)(const double **arg, double **res, long long *iw, double *w, int mem);
typedef long long (*casadi_n)(void);
typedef const long long *(*casadi_sp)(long long);

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* inputs in[i][B x n_i] (AoS, i < n_in of the function; NULL = zeros as in the generated code), reps passes over the
 * B scenarios; returns 0 and the wall time of the timed passes in *seconds, a checksum of the outputs in *checksum */
int ref_time_function(const char *so_path, const char *name, int B, const double *const *in, int reps, int threads,
                      double *seconds, double *checksum) {
  void *h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  char sym[128];
  casadi_fn f = (casadi_fn)dlsym(h, name);
  strcpy(sym, name); strcat(sym, "_n_in");
  casadi_n n_in = (casadi_n)dlsym(h, sym);
  strcpy(sym, name); strcat(sym, "_n_out");
  casadi_n n_out = (casadi_n)dlsym(h, sym);
  strcpy(sym, name); strcat(sym, "_sparsity_in");
  casadi_sp sp_in = (casadi_sp)dlsym(h, sym);
  strcpy(sym, name); strcat(sym, "_sparsity_out");
  casadi_sp sp_out = (casadi_sp)dlsym(h, sym);
  if (!f || !n_in || !n_out || !sp_in || !sp_out) { dlclose(h); return 2; }
  const int ni = (int)n_in(), no = (int)n_out();
  long long nnz_in[8], nnz_out[8];
  for (int i = 0; i < ni; i++) { const long long *s = sp_in(i); nnz_in[i] = s[2 + s[1]]; }
  for (int i = 0; i < no; i++) { const long long *s = sp_out(i); nnz_out[i] = s[2 + s[1]]; }
  double total = 0.0, t_all = 0.0;
  if (threads < 1) threads = 1;
  for (int rep = -1; rep < reps; rep++) { /* rep -1 = warm-up */
    const double t0 = now_s();
#pragma omp parallel num_threads(threads) reduction(+ : total)
    {
      double *out[8];
      for (int i = 0; i < no; i++) out[i] = (double *)malloc(sizeof(double) * (size_t)(nnz_out[i] > 0 ? nnz_out[i] : 1));
#pragma omp for schedule(static)
      for (int b = 0; b < B; b++) {
        const double *arg[8];
        double *res[8];
        for (int i = 0; i < ni; i++) arg[i] = in[i] ? in[i] + (size_t)b * (size_t)nnz_in[i] : NULL;
        for (int i = 0; i < no; i++) res[i] = out[i];
        f(arg, res, NULL, NULL, 0);
        total += out[no - 1][nnz_out[no - 1] > 0 ? nnz_out[no - 1] - 1 : 0];
      }
      for (int i = 0; i < no; i++) free(out[i]);
    }
    if (rep >= 0) t_all += now_s() - t0;
  }
  *seconds = t_all;
  *checksum = total;
  dlclose(h);
  return 0;
}
