/*
 * oracle/srb_ref.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Generic-N CPU restatement, in plain C, of the CasADi-generated oracle
 * functions of the reference's SRB landing NLP
 *   /root/reference/optimizations/landing/codegen_casadi/landingCtrller_IPOPT.c
 *   (nlp_f :10995, nlp_g :11161, nlp_grad :22015, nlp_grad_f :52602,
 *    nlp_hess_l :53527, nlp_jac_g :94014; sparsity tables casadi_s0..s5 :59-64)
 * following the symbolic model in
 *   optimizations/landing/generate_solver/generate_landingCtrller_IPOPT.m:41-170.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this.  The product (landing_controller_b200/csrc) never does.
 *
 * Parity pin: checked against the compiled reference C (oracle/_ref, N=21) and the
 * committed golden vectors under tests/golden/ (tests/test_oracle_eval.py).
 */
#ifndef SRB_REF_H
#define SRB_REF_H

#ifdef __cplusplus
extern "C" {
#endif

#define SRB_NJ_INT 385   /* Jacobian nz per interior knot   */
#define SRB_NJ_LAST 313  /* ... last knot (no no-slip rows) */
#define SRB_NH_INT 189   /* upper-tri Hessian nz per interior knot */
#define SRB_NH_LAST 177

typedef struct srb_plan {
  int N;            /* knots */
  int nx, np, m;    /* 36N-24, 13N+81, 104N-92 */
  int nnzJ, nnzH;   /* 385N-421, 189(N-1) */
  long long *spJ;   /* CasADi CCS: nrow,ncol,colind[ncol+1],row[nnz]  (mem.h:73-92) */
  long long *spH;   /* upper triangular */
  int *jmap;        /* [(N-1)*385] knot-local emission index -> nz index      */
  int *hmap;        /* [(N-1)*189]                                           */
  int jbnd[36];     /* nz index of the 36 boundary-row entries               */
  int hterm[12];    /* nz index of the terminal-cost diagonal                */
  /* parameter offsets inside p (optistack declaration order)     */
  int o_dt, o_qmin, o_qmax, o_qdmin, o_qdmax, o_qinit, o_qdinit;
  int o_qtmin, o_qtmax, o_qdtmin, o_qdtmax, o_QN, o_mu, o_lleg, o_fmax, o_mass, o_Ib, o_Ibinv;
} srb_plan;

srb_plan *srb_plan_create(int N);
void srb_plan_free(srb_plan *pl);

/* all functions return 0 on success, -1 if a NaN/Inf was produced (oracle_function.cpp:218-266) */
int srb_f(const srb_plan *pl, const double *x, const double *p, double *f);
int srb_g(const srb_plan *pl, const double *x, const double *p, double *g);
int srb_grad_f(const srb_plan *pl, const double *x, const double *p, double *f, double *grad);
int srb_jac_g(const srb_plan *pl, const double *x, const double *p, double *g, double *jac);
int srb_hess_l(const srb_plan *pl, const double *x, const double *p, double lam_f,
               const double *lam_g, double *hess);
int srb_grad(const srb_plan *pl, const double *x, const double *p, double lam_f,
             const double *lam_g, double *f, double *g, double *ggx, double *ggp);

/* knot-level access for the interior-point reference (oracle/ip_ref.c): emission-order lists of
 * one knot's Jacobian / Hessian entries and their (row,var) / (var,var) patterns.
 * knot-local variables: 0-11 X_k, 12-23 c_k, 24-35 f_k, 36-47 X_{k+1}, 48-59 c_{k+1}. */
typedef struct { short r, v; } srb_jpat;
typedef struct { short a, b; } srb_hpat;
void srb_knot_pattern(int last, srb_jpat *jp, srb_hpat *hp);
void srb_knot_lists(const srb_plan *pl, const double *x, const double *p, int k,
                    const double *lam_local, double *gl, double *Jl, double *Hl);

/* lbg(p), ubg(p): the Opti canonicalisation (optistack_internal.cpp:742-870) of
 * generate_landingCtrller_IPOPT.m:90-169; +-HUGE_VAL for one-sided rows. */
void srb_bounds(const srb_plan *pl, const double *p, double *lbg, double *ubg);

/* Shared problem data of a sweep (generate_landingCtrller_IPOPT.m:173-196 defaults
 * or generate_training_data_automated.m:62-102). */
typedef struct srb_problem {
  double T;                 /* horizon; dt = T/(N-1) uniform */
  double q_min[6], q_max[6], qd_min[6], qd_max[6];
  double q_term_min[6], q_term_max[6], qd_term_min[6], qd_term_max[6];
  double q_term_ref[6], qd_term_ref[6];
  double c_ref[12];
  double QN[12];
  double mu, l_leg_max, f_max, mass, Ib[3], Ib_inv[3];
  const double *dt;         /* knot spacings dt[0..N-2], or NULL for the uniform T/(N-1)
                               (generate_training_data_automated.m:28: [0.05 0.02x15 0.05 0.05 0.1 0.2]) */
} srb_problem;

void srb_problem_default(srb_problem *pb);

/* p and x0 from one drop condition (q_init[6], qd_init[6]):
 * Xref = per-row linspace(init, term_ref, N); Uref feet = Xref_pos + c_ref, forces 0;
 * x0 = [Xref(:); Uref(:)]  (generate_landingCtrller_IPOPT.m:199-208,336). */
void srb_build_p_x0(const srb_plan *pl, const srb_problem *pb, const double *q_init,
                    const double *qd_init, double *p, double *x0);

#ifdef __cplusplus
}
#endif
#endif
