// solver.cuh -- batched primal-dual interior-point solver (declarations)
#pragma once
#include <string>

#include "kernels.cuh"

namespace srb {

struct SolverWorkspace {
  void* dev = nullptr;       // per-warp-slot scratch (iterate, stage blocks, Riccati factors)
  size_t dev_bytes = 0;
  void* io = nullptr;        // staging of drops / results for host-buffer calls
  size_t io_bytes = 0;
  int* counter = nullptr;    // work-queue head
};

void solver_free(SolverWorkspace& ws);

int solver_run(SolverWorkspace& ws, const DevicePlan& pl, long long B, int memspace,
               const landing_problem& pb, const landing_options& opt, const landing_solve_io& io,
               cudaStream_t st, int* launches, std::string* err);

}  // namespace srb
