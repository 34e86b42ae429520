// solver.cuh -- batched primal-dual interior-point solver (declarations)
#pragma once
#include <string>

#include "kernels.cuh"

namespace srb {

// Host-built, device-resident index tables derived from the knot pattern (interior knot numbering;
// the last knot is remapped onto it).  See solver.cu for the meaning of each table.
struct SolverTables {
  int *dev = nullptr;  // one allocation
  const int *jl_last, *hl_last;            // emission index of the last-knot template -> interior index
  const int *g_e, *g_t; int g_n;           // dynamics entries -> G[state*36 + var]
  const int *h_t;                          // Hessian entry -> M target i*48+j
  const int *t_ptr, *t_ij, *t_terms; int t_n;  // condensing targets: sum sigma_rho J_ea J_eb
  const int *q_ptr, *q_terms;              // stage gradient: per stage variable, sum yhat_rho J_e
  const int *r_ptr, *r_terms;              // per inequality row rho: sum J_e dw[idx]
  const int *c_ptr, *c_terms;              // per local variable (60): sum J_e y_rho  (grad of Lagrangian)
};

struct SolverWorkspace {
  double* scratch = nullptr;  // per-warp-slot scratch (iterate, lists, Riccati factors)
  size_t scratch_bytes = 0;
  void* io = nullptr;         // staging of drops / results for host-buffer calls
  size_t io_bytes = 0;
  int* counter = nullptr;     // work-queue head
  SolverTables tab;
  int n_sm = 0;
};

void solver_free(SolverWorkspace& ws);

int solver_run(SolverWorkspace& ws, const DevicePlan& pl, long long B, int memspace,
               const landing_problem& pb, const landing_options& opt, const landing_solve_io& io,
               cudaStream_t st, int* launches, std::string* err);

}  // namespace srb
