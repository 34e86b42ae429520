// plan.cuh -- problem sizes, parameter offsets and CCS patterns for the N-knot landing NLP.
//
// The reference hard-codes these for N=21 as casadi_s0..s5
// (optimizations/landing/codegen_casadi/landingCtrller_IPOPT.c:59-64).  Here they are generated
// for any N by replaying the knot template (srb_knot.cuh) with a pattern-capturing sink and
// sorting (col,row) as CasADi does; tests check integer equality with the reference at N=21.
#pragma once
#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "srb_knot.cuh"

namespace srb {

struct ParamOff {  // offsets into p: optistack declaration order
  int dt, qmin, qmax, qdmin, qdmax, qinit, qdinit, qtmin, qtmax, qdtmin, qdtmax, QN, mu, lleg, fmax,
      mass, Ib, Ibinv;
};

SRB_HD ParamOff param_offsets(int N) {
  ParamOff o;
  int t = 12 * N;
  o.dt = t; t += N - 1;
  o.qmin = t; t += 6;
  o.qmax = t; t += 6;
  o.qdmin = t; t += 6;
  o.qdmax = t; t += 6;
  o.qinit = t; t += 6;
  o.qdinit = t; t += 6;
  o.qtmin = t; t += 6;
  o.qtmax = t; t += 6;
  o.qdtmin = t; t += 6;
  o.qdtmax = t; t += 6;
  o.QN = t; t += 12;
  o.mu = t++;
  o.lleg = t++;
  o.fmax = t++;
  o.mass = t++;
  o.Ib = t; t += 3;
  o.Ibinv = t;
  return o;
}

// global index of knot-local variable v of knot k in x = [X(:); U(:)]
SRB_HD int global_var(int N, int k, int v) {
  if (v < 12) return 12 * k + v;
  if (v < 36) return 12 * N + 24 * k + (v - 12);
  if (v < 48) return 12 * (k + 1) + (v - 36);
  return 12 * N + 24 * (k + 1) + (v - 48);
}

struct PatternSink {
  std::vector<std::pair<int, int>> jac, hes;
  void g(int, double) {}
  void j(int e, int row, int var, double) {
    if ((int)jac.size() <= e) jac.resize(e + 1);
    jac[e] = {row, var};
  }
  void h(int e, int va, int vb, double) {
    if ((int)hes.size() <= e) hes.resize(e + 1);
    hes[e] = {va, vb};
  }
};

struct HostPlan {
  int N, nx, np, m, nnzJ, nnzH;
  ParamOff off;
  std::vector<long long> spJ, spH;   // CasADi CCS
  std::vector<long long> spDense[5]; // dense column vectors: nx, np, 1, m (and spare)
  std::vector<int> jmap, hmap;       // [(N-1) x NJ_INT], [(N-1) x NH_INT]; -1 = unused slot
  int jbnd[36], hterm[12];
};

namespace detail {
struct Trip { int row, col, id; };

inline std::vector<long long> to_ccs(std::vector<Trip>& t, int nrow, int ncol, std::vector<int>& id2nz) {
  std::sort(t.begin(), t.end(), [](const Trip& a, const Trip& b) {
    return a.col != b.col ? a.col < b.col : a.row < b.row;
  });
  std::vector<long long> sp(2 + ncol + 1 + t.size(), 0);
  sp[0] = nrow;
  sp[1] = ncol;
  for (size_t n = 0; n < t.size(); n++) {
    sp[2 + t[n].col + 1]++;
    sp[2 + ncol + 1 + n] = t[n].row;
    id2nz[t[n].id] = (int)n;
  }
  for (int c = 0; c < ncol; c++) sp[2 + c + 1] += sp[2 + c];
  return sp;
}

inline std::vector<long long> dense_col(int n) {
  std::vector<long long> sp(2 + 2 + n);
  sp[0] = n; sp[1] = 1; sp[2] = 0; sp[3] = n;
  for (int i = 0; i < n; i++) sp[4 + i] = i;
  return sp;
}
}  // namespace detail

inline std::shared_ptr<const HostPlan> make_plan(int N) {
  auto pl = std::make_shared<HostPlan>();
  pl->N = N;
  pl->nx = 36 * N - 24;
  pl->np = 13 * N + 81;
  pl->m = 104 * N - 92;
  pl->nnzJ = 385 * N - 421;
  pl->nnzH = 189 * (N - 1);
  pl->off = param_offsets(N);
  pl->spDense[0] = detail::dense_col(pl->nx);
  pl->spDense[1] = detail::dense_col(pl->np);
  pl->spDense[2] = detail::dense_col(1);
  pl->spDense[3] = detail::dense_col(pl->m);

  Knot z{};
  z.h = 0.03; z.mu = 1; z.mass = 1;
  for (int i = 0; i < 3; i++) { z.Ib[i] = 1; z.Ibinv[i] = 1; }
  PatternSink pi, plst;
  NoLam nl;
  knot_eval<false, false, true, true>(z, pi, nl);
  knot_eval<true, false, true, true>(z, plst, nl);

  using detail::Trip;
  {
    std::vector<Trip> t;
    t.reserve(pl->nnzJ);
    for (int i = 0; i < 12; i++) t.push_back({i, i, i});
    for (int i = 0; i < 6; i++) {
      t.push_back({12 + i, 12 * (N - 1) + i, 12 + i});
      t.push_back({18 + i, 12 * (N - 1) + i, 18 + i});
      t.push_back({24 + i, 12 * (N - 1) + 6 + i, 24 + i});
      t.push_back({30 + i, 12 * (N - 1) + 6 + i, 30 + i});
    }
    for (int k = 0; k < N - 1; k++) {
      const auto& pat = (k == N - 2) ? plst.jac : pi.jac;
      for (int e = 0; e < (int)pat.size(); e++)
        t.push_back({36 + 104 * k + pat[e].first, global_var(N, k, pat[e].second), 36 + k * NJ_INT + e});
    }
    std::vector<int> id2nz(36 + (N - 1) * NJ_INT, -1);
    pl->spJ = detail::to_ccs(t, pl->m, pl->nx, id2nz);
    for (int i = 0; i < 36; i++) pl->jbnd[i] = id2nz[i];
    pl->jmap.assign(id2nz.begin() + 36, id2nz.end());
  }
  {
    std::vector<Trip> t;
    t.reserve(pl->nnzH);
    for (int i = 0; i < 12; i++) t.push_back({12 * (N - 1) + i, 12 * (N - 1) + i, i});
    for (int k = 0; k < N - 1; k++) {
      const auto& pat = (k == N - 2) ? plst.hes : pi.hes;
      for (int e = 0; e < (int)pat.size(); e++)
        t.push_back({global_var(N, k, pat[e].first), global_var(N, k, pat[e].second), 12 + k * NH_INT + e});
    }
    std::vector<int> id2nz(12 + (N - 1) * NH_INT, -1);
    pl->spH = detail::to_ccs(t, pl->nx, pl->nx, id2nz);
    for (int i = 0; i < 12; i++) pl->hterm[i] = id2nz[i];
    pl->hmap.assign(id2nz.begin() + 12, id2nz.end());
  }
  return pl;
}

inline std::shared_ptr<const HostPlan> get_plan(int N) {
  static std::mutex mtx;
  static std::map<int, std::shared_ptr<const HostPlan>> cache;
  std::lock_guard<std::mutex> lk(mtx);
  auto it = cache.find(N);
  if (it != cache.end()) return it->second;
  auto pl = make_plan(N);
  cache[N] = pl;
  return pl;
}

}  // namespace srb
