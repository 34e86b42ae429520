// solver.cu -- placeholder until the interior-point kernel lands (fails loudly; no CPU fallback)
#include "solver.cuh"

namespace srb {
void solver_free(SolverWorkspace& ws) {
  if (ws.dev) cudaFree(ws.dev);
  if (ws.io) cudaFree(ws.io);
  if (ws.counter) cudaFree(ws.counter);
  ws = SolverWorkspace{};
}
int solver_run(SolverWorkspace&, const DevicePlan&, long long, int, const landing_problem&,
               const landing_options&, const landing_solve_io&, cudaStream_t, int*, std::string* err) {
  *err = "landing_solve_batch: interior-point kernel not built yet";
  return LANDING_ERR_ARG;
}
}  // namespace srb
