// capi.cu -- the thin C-ABI layer of liblanding_b200.so (include/landing_b200.h).
// Host side only: context, device staging of host buffers, launches.  No CPU compute path:
// every entry point that produces numbers needs a CUDA device and fails loudly without one.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "solver.cuh"

using namespace srb;

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(LANDING_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)
// restores the calling thread's current device when an entry point returns
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

struct landing_ctx {
  int N = 0, device = 0;
  std::shared_ptr<const HostPlan> plan;
  DevicePlan dpl{};
  cudaStream_t stream = nullptr;
  int* d_maps = nullptr;
  long long launches = 0;
  // grow-only device staging for host-buffer calls
  void* stage = nullptr;
  size_t stage_bytes = 0;
  void* pin = nullptr;  // mapped pinned host buffer of small LANDING_HOST evaluation calls (zero-copy)
  size_t pin_bytes = 0;
  // grow-only scratch for the transposed (SoA) copies of AoS batches
  void* tr = nullptr;
  size_t tr_bytes = 0;
  SolverWorkspace ws;
  // knot spacings of the current call: pinned host staging + device copy (landing_problem.dt or uniform)
  double* h_dt = nullptr;
  double* d_dt = nullptr;
  // contact schedule of the current call (formulation 1): bit mask per knot
  unsigned char* h_cs = nullptr;
  unsigned char* d_cs = nullptr;
  // kino-dynamic NLP (landing_kino_eval_batch): host plan and its device tables, made on first use
  std::shared_ptr<const KinoPlan> kino;
  int* d_kino = nullptr;
};

// dt[0..N-2] of this call on the device (stream-ordered): the caller's vector or the uniform T/(N-1)
static int stage_dt(landing_ctx* c, const landing_problem* pb) {
  const int K = c->N - 1;
  if (!c->h_dt) CU(cudaMallocHost(&c->h_dt, sizeof(double) * K));
  if (!c->d_dt) CU(cudaMalloc(&c->d_dt, sizeof(double) * K));
  CU(cudaStreamSynchronize(c->stream));  // the previous call's copy out of h_dt has completed
  for (int k = 0; k < K; k++) {
    const double h = pb->dt ? pb->dt[k] : pb->T / (double)K;
    if (!(h > 0.0)) return fail(LANDING_ERR_ARG, "landing_problem: knot spacings must be positive");
    c->h_dt[k] = h;
  }
  CU(cudaMemcpyAsync(c->d_dt, c->h_dt, sizeof(double) * K, cudaMemcpyHostToDevice, c->stream));
  if (pb->formulation == 1) {
    if (!pb->cs) return fail(LANDING_ERR_ARG, "landing_problem: formulation 1 needs the contact schedule cs[4 (N-1)]");
    if (!(pb->delta_c > 0.0)) return fail(LANDING_ERR_ARG, "landing_problem: delta_c must be positive");
    if (!c->h_cs) CU(cudaMallocHost(&c->h_cs, K));
    if (!c->d_cs) CU(cudaMalloc(&c->d_cs, K));
    for (int k = 0; k < K; k++) {
      unsigned m = 0;
      for (int l = 0; l < 4; l++) m |= (pb->cs[4 * k + l] != 0) << l;
      c->h_cs[k] = (unsigned char)m;
    }
    CU(cudaMemcpyAsync(c->d_cs, c->h_cs, K, cudaMemcpyHostToDevice, c->stream));
  } else if (pb->formulation != 0) {
    return fail(LANDING_ERR_ARG, "landing_problem: formulation must be 0 or 1");
  }
  return LANDING_OK;
}

static int ensure_tr(landing_ctx* c, size_t bytes) {
  if (bytes <= c->tr_bytes) return LANDING_OK;
  if (c->tr) cudaFree(c->tr);
  c->tr = nullptr;
  c->tr_bytes = 0;
  CU(cudaMalloc(&c->tr, bytes));
  c->tr_bytes = bytes;
  return LANDING_OK;
}

static int ensure_pin(landing_ctx* c, size_t bytes) {
  if (bytes <= c->pin_bytes) return LANDING_OK;
  if (c->pin) cudaFreeHost(c->pin);
  c->pin = nullptr;
  c->pin_bytes = 0;
  if (bytes < (64u << 10)) bytes = 64u << 10;
  CU(cudaHostAlloc(&c->pin, bytes, cudaHostAllocMapped));
  c->pin_bytes = bytes;
  return LANDING_OK;
}

static int ensure_stage(landing_ctx* c, size_t bytes) {
  if (bytes <= c->stage_bytes) return LANDING_OK;
  if (c->stage) cudaFree(c->stage);
  c->stage = nullptr;
  c->stage_bytes = 0;
  CU(cudaMalloc(&c->stage, bytes));
  c->stage_bytes = bytes;
  return LANDING_OK;
}

extern "C" {

void landing_destroy(landing_ctx* c);

const char* landing_last_error(void) { return g_err.c_str(); }

int landing_dims_for(int N, long long d[6]) {
  if (N < 3 || !d) return fail(LANDING_ERR_ARG, "landing_dims_for: need N >= 3");
  d[0] = N;
  d[1] = 36LL * N - 24;
  d[2] = 13LL * N + 81;
  d[3] = 104LL * N - 92;
  d[4] = 385LL * N - 421;
  d[5] = 189LL * (N - 1);
  return LANDING_OK;
}

const long long* landing_sparsity_for(int N, int which) {
  if (N < 3) return nullptr;
  auto pl = get_plan(N);
  switch (which) {
    case 0: return pl->spJ.data();
    case 1: return pl->spH.data();
    case 2: return pl->spDense[0].data();  // x
    case 3: return pl->spDense[1].data();  // p
    case 4: return pl->spDense[2].data();  // scalar
    case 5: return pl->spDense[3].data();  // g / lam_g
    default: return nullptr;
  }
}

int landing_create(int N, int device, landing_ctx** out) {
  if (!out || N < 3) return fail(LANDING_ERR_ARG, "landing_create: need N >= 3 and a ctx pointer");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(LANDING_ERR_NOGPU, "landing_create: no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(LANDING_ERR_ARG, "landing_create: bad device index");
  DeviceGuard guard_(device);
  landing_ctx* c = new landing_ctx();
  c->N = N;
  c->device = device;
  c->plan = get_plan(N);
  const HostPlan& pl = *c->plan;
  const size_t nj = pl.jmap.size(), nh = pl.hmap.size();
  // (any failure below releases what was created: no leaked context, stream or device memory)
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_maps, sizeof(int) * (nj + nh + 36 + 12));
  if (e == cudaSuccess) e = cudaMemcpy(c->d_maps, pl.jmap.data(), sizeof(int) * nj, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(c->d_maps + nj, pl.hmap.data(), sizeof(int) * nh, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(c->d_maps + nj + nh, pl.jbnd, sizeof(int) * 36, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(c->d_maps + nj + nh + 36, pl.hterm, sizeof(int) * 12, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    landing_destroy(c);
    return fail(LANDING_ERR_CUDA, std::string("landing_create: ") + cudaGetErrorString(e));
  }
  c->dpl = DevicePlan{pl.N, pl.nx, pl.np, pl.m, pl.nnzJ, pl.nnzH, pl.off,
                      c->d_maps, c->d_maps + nj, c->d_maps + nj + nh, c->d_maps + nj + nh + 36};
  *out = c;
  return LANDING_OK;
}

void landing_destroy(landing_ctx* c) {
  if (!c) return;
  DeviceGuard guard_(c->device);
  solver_free(c->ws);
  if (c->stage) cudaFree(c->stage);
  if (c->pin) cudaFreeHost(c->pin);
  if (c->tr) cudaFree(c->tr);
  if (c->d_maps) cudaFree(c->d_maps);
  if (c->d_dt) cudaFree(c->d_dt);
  if (c->h_dt) cudaFreeHost(c->h_dt);
  if (c->d_cs) cudaFree(c->d_cs);
  if (c->h_cs) cudaFreeHost(c->h_cs);
  if (c->d_kino) cudaFree(c->d_kino);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int landing_dims(const landing_ctx* c, long long d[6]) {
  if (!c) return fail(LANDING_ERR_ARG, "landing_dims: null ctx");
  return landing_dims_for(c->N, d);
}

const long long* landing_sparsity(const landing_ctx* c, int which) {
  return c ? landing_sparsity_for(c->N, which) : nullptr;
}

long long landing_launch_count(const landing_ctx* c) { return c ? c->launches : 0; }

int landing_synchronize(landing_ctx* c) {
  if (!c) return fail(LANDING_ERR_ARG, "landing_synchronize: null ctx");
  DeviceGuard guard_(c->device);
  CU(cudaStreamSynchronize(c->stream));
  return LANDING_OK;
}

// ---------------------------------------------------------------- kino-dynamic NLP
static std::shared_ptr<const KinoPlan> get_kino_plan(int N) {
  static std::mutex mu;
  static std::map<int, std::shared_ptr<const KinoPlan>> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(N);
  if (it != cache.end()) return it->second;
  auto pl = std::make_shared<const KinoPlan>(make_kino_plan(N));
  cache[N] = pl;
  return pl;
}

int landing_kino_dims(int N, long long d[4]) {
  if (N < 3 || !d) return fail(LANDING_ERR_ARG, "landing_kino_dims: need N >= 3");
  auto pl = get_kino_plan(N);
  d[0] = N; d[1] = pl->nx; d[2] = pl->m; d[3] = pl->nnz;
  return LANDING_OK;
}

const long long* landing_kino_sparsity(int N) { return N < 3 ? nullptr : get_kino_plan(N)->sparsity.data(); }

void landing_kino_setup_default(landing_kino_setup* k) {
  // generate_landingCtrller_KNITRO.m:214-262,325; get_robot_model.m:237-241
  std::memset(k, 0, sizeof(*k));
  const double qtmin[6] = {-10, -10, 0.15, -0.1, -0.1, -10}, qtmax[6] = {10, 10, 5, 0.1, 0.1, 10};
  const double qdtmin[6] = {-10, -10, -10, -.5, -.5, -.5}, qdtmax[6] = {10, 10, 10, .5, .5, .5};
  const double QN[12] = {0, 0, 100, 10, 10, 0, 10, 10, 10, 10, 10, 10};
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < 6; i++) {
    k->q_term_min[i] = qtmin[i]; k->q_term_max[i] = qtmax[i];
    k->qd_term_min[i] = qdtmin[i]; k->qd_term_max[i] = qdtmax[i];
  }
  for (int i = 0; i < 12; i++) k->QN[i] = QN[i];
  k->q_term_ref[2] = 0.25;
  k->z_min = 0.075;
  k->l_leg_max = 0.4;
  for (int l = 0; l < 4; l++) {
    k->jpos_min[3 * l] = -PI / 3; k->jpos_min[3 * l + 1] = -PI / 2; k->jpos_min[3 * l + 2] = 0.0;
    k->jpos_max[3 * l] = PI / 3; k->jpos_max[3 * l + 1] = PI / 2; k->jpos_max[3 * l + 2] = 3 * PI / 4;
  }
  k->tau_max[0] = 18.0; k->tau_max[1] = 18.0; k->tau_max[2] = 27.99;
  k->jpos_guess[0] = 0.0; k->jpos_guess[1] = -PI / 4; k->jpos_guess[2] = PI / 2;
}

// inputs / outputs of the two set-up calls staged through the context's buffer when they live in host memory
struct KinoBuf { const double* in; double* out; long long n; };
static int kino_stage(landing_ctx* c, long long B, int memspace, KinoBuf* bufs, int nb, const double** din, double** dout) {
  size_t tot = 0;
  for (int i = 0; i < nb; i++) if (bufs[i].in || bufs[i].out) tot += sizeof(double) * bufs[i].n * B;
  if (memspace != LANDING_HOST) {
    for (int i = 0; i < nb; i++) { din[i] = bufs[i].in; dout[i] = bufs[i].out; }
    return LANDING_OK;
  }
  int rc = ensure_stage(c, tot + 256);
  if (rc) return rc;
  char* s = (char*)c->stage;
  for (int i = 0; i < nb; i++) {
    din[i] = nullptr; dout[i] = nullptr;
    if (!bufs[i].in && !bufs[i].out) continue;
    const size_t by = sizeof(double) * bufs[i].n * B;
    if (bufs[i].in) { CU(cudaMemcpyAsync(s, bufs[i].in, by, cudaMemcpyHostToDevice, c->stream)); din[i] = (const double*)s; }
    else dout[i] = (double*)s;
    s += by;
  }
  return LANDING_OK;
}
static int kino_unstage(landing_ctx* c, long long B, int memspace, KinoBuf* bufs, int nb, double** dout) {
  if (memspace != LANDING_HOST) return LANDING_OK;
  for (int i = 0; i < nb; i++)
    if (bufs[i].out) CU(cudaMemcpyAsync(bufs[i].out, dout[i], sizeof(double) * bufs[i].n * B, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LANDING_OK;
}

int landing_kino_setup_batch(landing_ctx* c, long long B, int memspace, int layout, const landing_kino_setup* ks,
                             const double* drops, const double* x_srb, double* lbg, double* ubg, double* x0) {
  if (!c || !ks || !drops || B < 0) return fail(LANDING_ERR_ARG, "landing_kino_setup_batch: bad arguments");
  if (B == 0 || (!lbg && !ubg && !x0)) return LANDING_OK;
  DeviceGuard guard_(c->device);
  auto pl = get_kino_plan(c->N);
  const long long nsrb = 36LL * c->N - 24;
  KinoBuf bufs[5] = {{drops, nullptr, 12}, {x_srb, nullptr, nsrb}, {nullptr, lbg, pl->m}, {nullptr, ubg, pl->m}, {nullptr, x0, pl->nx}};
  const double* din[5]; double* dout[5];
  int rc = kino_stage(c, B, memspace, bufs, 5, din, dout);
  if (rc) return rc;
  KinoSetupArgs a{};
  a.N = c->N; a.B = B; a.ks = *ks;
  a.drops = make_cview(din[0], 12, B, layout);
  a.x_srb = make_cview(din[1], nsrb, B, layout);
  a.lbg = make_view(dout[2], pl->m, B, layout);
  a.ubg = make_view(dout[3], pl->m, B, layout);
  a.x0 = make_view(dout[4], pl->nx, B, layout);
  c->launches += launch_kino_setup(a, c->stream);
  CU(cudaGetLastError());
  return kino_unstage(c, B, memspace, bufs, 5, dout);
}

int landing_kino_cost_batch(landing_ctx* c, long long B, int memspace, int layout, const landing_kino_setup* ks,
                            const double* x, double* f, double* grad_f) {
  if (!c || !ks || !x || B < 0) return fail(LANDING_ERR_ARG, "landing_kino_cost_batch: bad arguments");
  if (B == 0 || (!f && !grad_f)) return LANDING_OK;
  DeviceGuard guard_(c->device);
  auto pl = get_kino_plan(c->N);
  KinoBuf bufs[3] = {{x, nullptr, pl->nx}, {nullptr, f, 1}, {nullptr, grad_f, pl->nx}};
  const double* din[3]; double* dout[3];
  int rc = kino_stage(c, B, memspace, bufs, 3, din, dout);
  if (rc) return rc;
  KinoSetupArgs a{};
  a.N = c->N; a.B = B; a.ks = *ks;
  a.x = make_cview(din[0], pl->nx, B, layout);
  a.f = make_view(dout[1], 1, B, layout);
  a.grad_f = make_view(dout[2], pl->nx, B, layout);
  c->launches += launch_kino_cost(a, c->stream);
  CU(cudaGetLastError());
  return kino_unstage(c, B, memspace, bufs, 3, dout);
}

int landing_kino_eval_batch(landing_ctx* c, long long B, int memspace, int layout, const landing_kino_problem* pb,
                            const double* x, double* g, double* jac) {
  if (!c || !pb || !pb->dt || !x || B < 0) return fail(LANDING_ERR_ARG, "landing_kino_eval_batch: bad arguments");
  if (B == 0 || (!g && !jac)) return LANDING_OK;
  DeviceGuard guard_(c->device);
  if (!c->kino) {
    auto pl = get_kino_plan(c->N);
    const size_t n = pl->gpos.size() + pl->bpos.size();
    CU(cudaMalloc(&c->d_kino, sizeof(int) * n));
    CU(cudaMemcpy(c->d_kino, pl->gpos.data(), sizeof(int) * pl->gpos.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_kino + pl->gpos.size(), pl->bpos.data(), sizeof(int) * pl->bpos.size(), cudaMemcpyHostToDevice));
    c->kino = pl;
  }
  const KinoPlan& pl = *c->kino;
  // knot spacings through the same staging as the solver's
  landing_problem tmp{};
  tmp.dt = pb->dt;
  int rc = stage_dt(c, &tmp);
  if (rc) return rc;
  const double* dx = x;
  double *dg = g, *dj = jac;
  if (memspace == LANDING_HOST) {
    const size_t bx = sizeof(double) * pl.nx * B, bg = g ? sizeof(double) * pl.m * B : 0, bj = jac ? sizeof(double) * pl.nnz * B : 0;
    rc = ensure_stage(c, bx + bg + bj + 256);
    if (rc) return rc;
    char* s = (char*)c->stage;
    CU(cudaMemcpyAsync(s, x, bx, cudaMemcpyHostToDevice, c->stream));
    dx = (const double*)s;
    dg = g ? (double*)(s + bx) : nullptr;
    dj = jac ? (double*)(s + bx + bg) : nullptr;
  }
  // AoS batches (one vector per scenario, what landing_solve_batch / landing_kino_setup_batch hand over) are evaluated
  // on transposed copies like the SRB functions (landing_eval_batch): three times the traffic, all of it coalesced.
  double *aos_g = nullptr, *aos_j = nullptr;
  if (layout == LANDING_AOS && B >= 64) {
    rc = ensure_tr(c, sizeof(double) * (pl.nx + (dg ? pl.m : 0) + (dj ? pl.nnz : 0)) * B + 256);
    if (rc) return rc;
    double* cur = (double*)c->tr;
    c->launches += launch_transpose(dx, cur, B, pl.nx, c->stream);  // [B][n_x] -> [n_x][B]
    dx = cur; cur += pl.nx * B;
    if (dg) { aos_g = dg; dg = cur; cur += pl.m * B; }
    if (dj) { aos_j = dj; dj = cur; cur += pl.nnz * B; }
    layout = LANDING_SOA;
  }
  KinoArgs a{};
  a.N = c->N; a.B = B;
  a.x = make_cview(dx, pl.nx, B, layout);
  a.g = make_view(dg, pl.m, B, layout);
  a.jac = make_view(dj, pl.nnz, B, layout);
  a.pr.mu = pb->mu; a.pr.mass = pb->mass;
  for (int i = 0; i < 3; i++) { a.pr.Ib[i] = pb->Ib[i]; a.pr.Ib_inv[i] = pb->Ib_inv[i]; }
  a.dtv = c->d_dt;
  a.gpos = c->d_kino; a.bpos = c->d_kino + pl.gpos.size();
  c->launches += launch_kino(a, dg != nullptr, dj != nullptr, c->stream);
  CU(cudaGetLastError());
  if (aos_g) { c->launches += launch_transpose(dg, aos_g, pl.m, B, c->stream); dg = aos_g; }    // [m][B] -> [B][m]
  if (aos_j) { c->launches += launch_transpose(dj, aos_j, pl.nnz, B, c->stream); dj = aos_j; }
  CU(cudaGetLastError());
  if (memspace == LANDING_HOST) {
    if (g) CU(cudaMemcpyAsync(g, dg, sizeof(double) * pl.m * B, cudaMemcpyDeviceToHost, c->stream));
    if (jac) CU(cudaMemcpyAsync(jac, dj, sizeof(double) * pl.nnz * B, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return LANDING_OK;
}

int landing_fp64_peak(landing_ctx* c, double* tflops) {
  if (!c || !tflops) return fail(LANDING_ERR_ARG, "landing_fp64_peak: null argument");
  DeviceGuard guard_(c->device);
  std::string err;
  const int rc = fp64_peak_run(c->stream, tflops, &err);
  if (rc != LANDING_OK) return fail(rc, err);
  c->launches += 4;
  return LANDING_OK;
}
void* landing_stream(const landing_ctx* c) { return c ? (void*)c->stream : nullptr; }

void landing_problem_default(landing_problem* pb) {
  // generate_landingCtrller_IPOPT.m:173-196
  static const double qmin[6] = {-10, -10, 0.1, -10, -10, -10}, qmax[6] = {10, 10, 1.0, 10, 10, 10};
  static const double qdmin[6] = {-10, -10, -10, -40, -40, -40}, qdmax[6] = {10, 10, 10, 40, 40, 40};
  static const double qtmin[6] = {-10, -10, 0.2, -0.1, -0.1, -10}, qtmax[6] = {10, 10, 5, 0.1, 0.1, 10};
  static const double qtref[6] = {0, 0, 0.275, 0, 0, 0};
  static const double QN[12] = {0, 0, 100, 100, 100, 0, 10, 10, 10, 10, 10, 10};
  static const double sgn[12] = {1, -1, 1, 1, 1, 1, -1, -1, 1, -1, 1, 1};
  static const double cr[3] = {0.2, 0.1, -0.2};
  pb->T = 0.6;
  for (int i = 0; i < 6; i++) {
    pb->q_min[i] = qmin[i]; pb->q_max[i] = qmax[i];
    pb->qd_min[i] = qdmin[i]; pb->qd_max[i] = qdmax[i];
    pb->q_term_min[i] = qtmin[i]; pb->q_term_max[i] = qtmax[i];
    pb->qd_term_min[i] = qdmin[i]; pb->qd_term_max[i] = qdmax[i];
    pb->q_term_ref[i] = qtref[i]; pb->qd_term_ref[i] = 0.0;
  }
  for (int i = 0; i < 12; i++) { pb->QN[i] = QN[i]; pb->c_ref[i] = sgn[i] * cr[i % 3]; }
  pb->mu = 1.0;
  pb->l_leg_max = 0.35;
  pb->f_max = 200.0;
  for (int i = 0; i < 3; i++) pb->Qf[i] = 0.0;
  pb->kin_box[0] = 0.15; pb->kin_box[1] = 0.15; pb->kin_box[2] = 0.30;
  pb->dt = nullptr;  // uniform T/(N-1)
  pb->formulation = 0;
  pb->cs = nullptr;
  for (int i = 0; i < 12; i++) pb->QX[i] = 0.0;
  pb->delta_c = 1e-7;
  // composite rigid-body inertia at q_home (get_mass_matrix.m:19-54, generate_landingCtrller_IPOPT.m:99-104)
  pb->mass = 8.251999999999999;
  pb->Ib[0] = 0.05757729852959269; pb->Ib[1] = 0.23400899479539086; pb->Ib[2] = 0.2796738482657981;
  pb->Ib_inv[0] = 17.37746888893693; pb->Ib_inv[1] = 4.27334000932043; pb->Ib_inv[2] = 3.577551923825657;
}

void landing_options_default(landing_options* o) {
  // generate_landingCtrller_IPOPT.m:232-263
  std::memset(o, 0, sizeof(*o));
  o->max_iter = 3000;
  o->tol = 1e-4;
  o->constr_viol_tol = 1e-3;
  o->dual_inf_tol = 1.0;
  o->compl_inf_tol = 1e-4;
  o->mu_init = 0.1;
  o->bound_push = 0.5;
  o->bound_frac = 0.5;
  o->bound_relax_factor = 1e-6;
  o->max_soc = 4;
  o->jam_iters = 5;
  o->max_restarts = 8;
  o->jam_alpha = 0.02;
  o->restart_mu = 0.0;  // = mu_init
}

int landing_eval_batch(landing_ctx* c, long long B, int memspace, int layout, const landing_eval_io* io) {
  if (!c || !io || B < 0) return fail(LANDING_ERR_ARG, "landing_eval_batch: bad arguments");
  if (B == 0) return LANDING_OK;
  DeviceGuard guard_(c->device);
  const DevicePlan& pl = c->dpl;
  const long long nx = pl.nx, np = pl.np, m = pl.m, nj = pl.nnzJ, nh = pl.nnzH;
  // sizes (doubles) of the 4 inputs and 7 outputs
  const long long in_n[4] = {nx, np, 1, m};
  const double* in_p[4] = {io->x, io->p, io->lam_f, io->lam_g};
  const long long out_n[7] = {1, m, nx, nj, nh, nx, np};
  double* out_p[7] = {io->f, io->g, io->grad_f, io->jac, io->hess, io->grad_x, io->grad_p};
  const double* din[4];
  double* dout[7];
  int* dstatus = io->status;
  // Small host calls -- the CasADi ABI's one scenario per call above all -- go ZERO-COPY: inputs are copied by the CPU
  // into a mapped pinned buffer, the kernels read them and write their results through the device alias of that buffer,
  // and the CPU copies the results out after the stream synchronisation.  One launch + one synchronisation per call
  // instead of up to eleven pageable cudaMemcpyAsync (LANDING_NO_ZEROCOPY=1 restores the staged copies).
  bool zero_copy = false;
  if (memspace == LANDING_HOST) {
    size_t tot = 0;
    for (int i = 0; i < 4; i++) if (in_p[i]) tot += sizeof(double) * in_n[i] * B;
    for (int i = 0; i < 7; i++) if (out_p[i]) tot += sizeof(double) * out_n[i] * B;
    tot += sizeof(int) * B + 256;
    static const bool no_zc = getenv("LANDING_NO_ZEROCOPY") != nullptr;
    // (not for nlp_grad: its parameter sensitivities are accumulated with device-scope atomics)
    zero_copy = !no_zc && tot <= (512u << 10) && B < 64 && !io->grad_x && !io->grad_p;
    int rc = zero_copy ? ensure_pin(c, tot) : ensure_stage(c, tot);
    if (rc) return rc;
    char* cur = (char*)(zero_copy ? c->pin : c->stage);
    for (int i = 0; i < 4; i++) {
      din[i] = nullptr;
      if (in_p[i]) {
        din[i] = (const double*)cur;
        if (zero_copy) std::memcpy(cur, in_p[i], sizeof(double) * in_n[i] * B);
        else CU(cudaMemcpyAsync(cur, in_p[i], sizeof(double) * in_n[i] * B, cudaMemcpyHostToDevice, c->stream));
        cur += sizeof(double) * in_n[i] * B;
      }
    }
    for (int i = 0; i < 7; i++) {
      dout[i] = nullptr;
      if (out_p[i]) { dout[i] = (double*)cur; cur += sizeof(double) * out_n[i] * B; }
    }
    dstatus = io->status ? (int*)cur : nullptr;
  } else {
    for (int i = 0; i < 4; i++) din[i] = in_p[i];
    for (int i = 0; i < 7; i++) dout[i] = out_p[i];
  }
  // AoS batches (one CasADi vector per scenario) would make every warp access 32 different rows: for batches the
  // kernels run on transposed (SoA) copies instead and the results are transposed back through shared-memory tiles
  // (3 x the traffic of a SoA call, but coalesced: DESIGN.md 2.2).  Small batches (the CasADi ABI: B = 1) go direct.
  double* aos_out[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const bool via_soa = (layout == LANDING_AOS && B >= 64);
  if (via_soa) {
    size_t tot = 0;
    for (int i = 0; i < 4; i++) if (din[i]) tot += sizeof(double) * in_n[i] * B;
    for (int i = 0; i < 7; i++) if (dout[i]) tot += sizeof(double) * out_n[i] * B;
    int rc = ensure_tr(c, tot + 256);
    if (rc) return rc;
    double* cur = (double*)c->tr;
    for (int i = 0; i < 4; i++)
      if (din[i]) {
        c->launches += launch_transpose(din[i], cur, B, in_n[i], c->stream);  // [B][n] -> [n][B]
        din[i] = cur;
        cur += in_n[i] * B;
      }
    for (int i = 0; i < 7; i++)
      if (dout[i]) { aos_out[i] = dout[i]; dout[i] = cur; cur += out_n[i] * B; }
    layout = LANDING_SOA;
  }
  EvalArgs a{};
  a.pl = pl;
  a.B = B;
  a.x = make_cview(din[0], nx, B, layout);
  a.p = make_cview(din[1], np, B, layout);
  a.lam_f = make_cview(din[2], 1, B, layout);
  a.lam_g = make_cview(din[3], m, B, layout);
  a.f = make_view(dout[0], 1, B, layout);
  a.g = make_view(dout[1], m, B, layout);
  a.grad_f = make_view(dout[2], nx, B, layout);
  a.jac = make_view(dout[3], nj, B, layout);
  a.hess = make_view(dout[4], nh, B, layout);
  a.grad_x = make_view(dout[5], nx, B, layout);
  a.grad_p = make_view(dout[6], np, B, layout);
  a.status = dstatus;
  c->launches += launch_eval(a, c->stream);
  CU(cudaGetLastError());
  if (via_soa) {
    for (int i = 0; i < 7; i++)
      if (aos_out[i]) {
        c->launches += launch_transpose(dout[i], aos_out[i], out_n[i], B, c->stream);  // [n][B] -> [B][n]
        dout[i] = aos_out[i];
      }
    CU(cudaGetLastError());
  }
  if (memspace == LANDING_HOST && zero_copy) {
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 7; i++)
      if (out_p[i]) std::memcpy(out_p[i], dout[i], sizeof(double) * out_n[i] * B);
    if (io->status) std::memcpy(io->status, dstatus, sizeof(int) * B);
  } else if (memspace == LANDING_HOST) {
    for (int i = 0; i < 7; i++)
      if (out_p[i])
        CU(cudaMemcpyAsync(out_p[i], dout[i], sizeof(double) * out_n[i] * B, cudaMemcpyDeviceToHost, c->stream));
    if (io->status)
      CU(cudaMemcpyAsync(io->status, dstatus, sizeof(int) * B, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return LANDING_OK;
}

void landing_tvlqr_default(landing_tvlqr* q) {
  // quadruped_SRBM_NLP.m:436-475: F = diag(1 1 1, 5 5 5, 4 4 4, 3 3 3, 0...), Q = diag(.25 x3, 1 x3, .5 x3, 1 x3, 0...),
  // R = 90 I, dt = 0.022; inertia at the zero configuration (generateVariationalDynamics.m:4-7; oracle/crba_constants.py)
  static const double f[12] = {1, 1, 1, 5, 5, 5, 4, 4, 4, 3, 3, 3};
  static const double w[12] = {0.25, 0.25, 0.25, 1, 1, 1, 0.5, 0.5, 0.5, 1, 1, 1};
  for (int i = 0; i < 576; i++) q->Q[i] = q->F[i] = 0.0;
  for (int i = 0; i < 12; i++) { q->F[i * 24 + i] = f[i]; q->Q[i * 24 + i] = w[i]; q->R[i] = 90.0; }
  q->T = 0.6;
  q->dt = 0.022;
  q->n_steps = 28;  // T / dt + 1
  static const double Ib[9] = {8.2057376e-02, 0.0, -5.020000000000024e-05, 0.0, 2.46291e-01, 0.0,
                               -5.020000000000024e-05, 0.0, 2.67475776e-01};
  for (int i = 0; i < 9; i++) q->Ib[i] = Ib[i];
  q->mass = 8.251999999999999;
}

int landing_tvlqr_batch(landing_ctx* c, long long B, int memspace, const landing_tvlqr* par, const double* x_star,
                        double* P_out, double* K_out) {
  if (!c || !par || !x_star || B < 0 || par->n_steps < 1) return fail(LANDING_ERR_ARG, "landing_tvlqr_batch: bad arguments");
  if (B == 0) return LANDING_OK;
  DeviceGuard guard_(c->device);
  const long long nx = c->dpl.nx, ns = par->n_steps;
  TvlqrArgs a{};
  a.N = c->N; a.nx = nx; a.par = *par;
  a.x_star = x_star; a.P_out = P_out; a.K_out = K_out;
  if (memspace == LANDING_HOST) {
    const size_t bx = sizeof(double) * nx * B, bp = P_out ? sizeof(double) * 576 * ns * B : 0, bk = K_out ? sizeof(double) * 288 * ns * B : 0;
    int rc = ensure_stage(c, bx + bp + bk + 256);
    if (rc) return rc;
    char* s = (char*)c->stage;
    CU(cudaMemcpyAsync(s, x_star, bx, cudaMemcpyHostToDevice, c->stream));
    a.x_star = (const double*)s;
    a.P_out = P_out ? (double*)(s + bx) : nullptr;
    a.K_out = K_out ? (double*)(s + bx + bp) : nullptr;
  }
  c->launches += launch_tvlqr(a, B, c->stream);
  CU(cudaGetLastError());
  if (memspace == LANDING_HOST) {
    if (P_out) CU(cudaMemcpyAsync(P_out, a.P_out, sizeof(double) * 576 * ns * B, cudaMemcpyDeviceToHost, c->stream));
    if (K_out) CU(cudaMemcpyAsync(K_out, a.K_out, sizeof(double) * 288 * ns * B, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return LANDING_OK;
}

int landing_bounds_batch(landing_ctx* c, long long B, int memspace, int layout, const double* p,
                         double* lbg, double* ubg) {
  if (!c || !p || !lbg || !ubg || B <= 0) return fail(LANDING_ERR_ARG, "landing_bounds_batch: bad arguments");
  DeviceGuard guard_(c->device);
  const DevicePlan& pl = c->dpl;
  const double* dp = p;
  double *dl = lbg, *du = ubg;
  if (memspace == LANDING_HOST) {
    int rc = ensure_stage(c, sizeof(double) * B * (pl.np + 2LL * pl.m));
    if (rc) return rc;
    double* s = (double*)c->stage;
    CU(cudaMemcpyAsync(s, p, sizeof(double) * B * pl.np, cudaMemcpyHostToDevice, c->stream));
    dp = s;
    dl = s + B * pl.np;
    du = dl + B * pl.m;
  }
  c->launches += launch_bounds(pl, B, make_cview(dp, pl.np, B, layout), make_view(dl, pl.m, B, layout),
                               make_view(du, pl.m, B, layout), c->stream);
  CU(cudaGetLastError());
  if (memspace == LANDING_HOST) {
    CU(cudaMemcpyAsync(lbg, dl, sizeof(double) * B * pl.m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(ubg, du, sizeof(double) * B * pl.m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return LANDING_OK;
}

int landing_build_batch(landing_ctx* c, long long B, int memspace, int layout, const landing_problem* pb,
                        const double* drops, double* p, double* x0) {
  if (!c || !pb || !drops || B <= 0) return fail(LANDING_ERR_ARG, "landing_build_batch: bad arguments");
  DeviceGuard guard_(c->device);
  const DevicePlan& pl = c->dpl;
  const double* dd = drops;
  double *dp = p, *dx = x0;
  if (memspace == LANDING_HOST) {
    int rc = ensure_stage(c, sizeof(double) * B * (12 + pl.np + pl.nx));
    if (rc) return rc;
    double* s = (double*)c->stage;
    CU(cudaMemcpyAsync(s, drops, sizeof(double) * B * 12, cudaMemcpyHostToDevice, c->stream));
    dd = s;
    dp = p ? s + 12 * B : nullptr;
    dx = x0 ? s + 12 * B + B * pl.np : nullptr;
  }
  {
    const int rc = stage_dt(c, pb);
    if (rc) return rc;
  }
  c->launches += launch_build(pl, B, *pb, c->d_dt, dd, make_view(dp, pl.np, B, layout), make_view(dx, pl.nx, B, layout),
                              c->stream);
  CU(cudaGetLastError());
  if (memspace == LANDING_HOST) {
    if (p) CU(cudaMemcpyAsync(p, dp, sizeof(double) * B * pl.np, cudaMemcpyDeviceToHost, c->stream));
    if (x0) CU(cudaMemcpyAsync(x0, dx, sizeof(double) * B * pl.nx, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return LANDING_OK;
}

int landing_solve_batch(landing_ctx* c, long long B, int memspace, const landing_problem* pb,
                        const landing_options* opt, const landing_solve_io* io) {
  if (!c || !pb || !io || B < 0) return fail(LANDING_ERR_ARG, "landing_solve_batch: bad arguments");
  if (B == 0) return LANDING_OK;  // an empty sweep is a no-op (the reference's sweep loops simply do not iterate)
  if (!io->drops) return fail(LANDING_ERR_ARG, "landing_solve_batch: drops is NULL");
  DeviceGuard guard_(c->device);
  landing_options o;
  if (opt) o = *opt; else landing_options_default(&o);
  std::string err;
  int launches = 0;
  int rc = stage_dt(c, pb);
  if (rc) return rc;
  rc = solver_run(c->ws, c->dpl, B, memspace, *pb, c->d_dt, pb->formulation == 1 ? c->d_cs : nullptr, o, *io, c->stream,
                  &launches, &err);
  c->launches += launches;
  if (rc) return fail(rc, err);
  return LANDING_OK;
}

// ---------------------------------------------------------------- several GPUs, one process
struct landing_multi {
  int N = 0;
  std::vector<landing_ctx*> ctx;
  struct Shard {  // pinned host staging of one device's shard (grow-only)
    long long cap = 0, cap_lam = 0, cap_x0 = 0;
    double *drops = nullptr, *x = nullptr, *f = nullptr, *viol = nullptr, *lam = nullptr, *x0 = nullptr;
    int *status = nullptr, *iters = nullptr;
  };
  std::vector<Shard> sh;
};

int landing_multi_create(int N, int n_devices, const int* devices, landing_multi** out) {
  if (!out || !devices || n_devices < 1) return fail(LANDING_ERR_ARG, "landing_multi_create: bad arguments");
  std::unique_ptr<landing_multi> m(new landing_multi());
  m->N = N;
  m->sh.resize(n_devices);
  for (int d = 0; d < n_devices; d++) {
    landing_ctx* c = nullptr;
    const int rc = landing_create(N, devices[d], &c);
    if (rc) {
      for (landing_ctx* q : m->ctx) landing_destroy(q);
      return rc;
    }
    m->ctx.push_back(c);
  }
  *out = m.release();
  return LANDING_OK;
}

void landing_multi_destroy(landing_multi* m) {
  if (!m) return;
  for (size_t d = 0; d < m->ctx.size(); d++) {
    DeviceGuard guard_(m->ctx[d]->device);
    landing_multi::Shard& s = m->sh[d];
    for (void* p : {(void*)s.drops, (void*)s.x, (void*)s.f, (void*)s.viol, (void*)s.lam, (void*)s.x0, (void*)s.status, (void*)s.iters})
      if (p) cudaFreeHost(p);
    landing_destroy(m->ctx[d]);
  }
  delete m;
}

int landing_solve_batch_multi(landing_multi* m, long long B, const landing_problem* pb, const landing_options* opt,
                              const landing_solve_io* io) {
  if (!m || !pb || !io || B < 0) return fail(LANDING_ERR_ARG, "landing_solve_batch_multi: bad arguments");
  if (B == 0) return LANDING_OK;
  if (!io->drops || !io->x_star || !io->f_star || !io->status || !io->iters)
    return fail(LANDING_ERR_ARG, "landing_solve_batch_multi: drops, x_star, f_star, status and iters are required");
  const int G = (int)m->ctx.size();
  const long long nx = 36LL * m->N - 24, mr = 104LL * m->N - 92;
  std::vector<int> rcs(G, LANDING_OK);
  std::vector<std::string> errs(G);
  auto work = [&](int d) {
    landing_ctx* c = m->ctx[d];
    landing_multi::Shard& s = m->sh[d];
    const long long Bd = (B - d + G - 1) / G;  // scenarios d, d + G, ...
    if (Bd <= 0) return;
    DeviceGuard guard_(c->device);
    auto grow = [&](auto*& p, long long n) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      return cudaMallocHost((void**)&p, sizeof(*p) * (size_t)n) == cudaSuccess;
    };
    bool ok = true;
    if (Bd > s.cap) {
      ok = grow(s.drops, 12 * Bd) && grow(s.x, nx * Bd) && grow(s.f, Bd) && grow(s.viol, Bd) && grow(s.status, Bd) &&
           grow(s.iters, Bd);
      s.cap = ok ? Bd : 0;
    }
    if (ok && io->lam_g && Bd > s.cap_lam) { ok = grow(s.lam, mr * Bd); s.cap_lam = ok ? Bd : 0; }
    if (ok && io->x0 && Bd > s.cap_x0) { ok = grow(s.x0, nx * Bd); s.cap_x0 = ok ? Bd : 0; }
    if (!ok) { rcs[d] = LANDING_ERR_CUDA; errs[d] = "landing_solve_batch_multi: pinned host allocation failed"; return; }
    for (long long i = 0; i < Bd; i++) {
      const long long b = d + i * G;
      std::memcpy(s.drops + 12 * i, io->drops + 12 * b, sizeof(double) * 12);
      if (io->x0) std::memcpy(s.x0 + nx * i, io->x0 + nx * b, sizeof(double) * nx);
    }
    landing_solve_io sio{};
    sio.drops = s.drops; sio.x0 = io->x0 ? s.x0 : nullptr;
    sio.x_star = s.x; sio.f_star = s.f; sio.lam_g = io->lam_g ? s.lam : nullptr; sio.viol = s.viol;
    sio.status = s.status; sio.iters = s.iters;
    rcs[d] = landing_solve_batch(c, Bd, LANDING_HOST, pb, opt, &sio);
    if (rcs[d]) { errs[d] = landing_last_error(); return; }
    for (long long i = 0; i < Bd; i++) {  // gather: this device's records into the whole sweep's arrays
      const long long b = d + i * G;
      std::memcpy(io->x_star + nx * b, s.x + nx * i, sizeof(double) * nx);
      io->f_star[b] = s.f[i];
      io->status[b] = s.status[i];
      io->iters[b] = s.iters[i];
      if (io->viol) io->viol[b] = s.viol[i];
      if (io->lam_g) std::memcpy(io->lam_g + mr * b, s.lam + mr * i, sizeof(double) * mr);
    }
  };
  std::vector<std::thread> th;
  for (int d = 1; d < G; d++) th.emplace_back(work, d);
  work(0);
  for (auto& t : th) t.join();
  for (int d = 0; d < G; d++)
    if (rcs[d]) return fail(rcs[d], errs[d]);
  return LANDING_OK;
}

}  // extern "C"
