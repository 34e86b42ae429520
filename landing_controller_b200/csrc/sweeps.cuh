// sweeps.cuh -- the condensing pre-pass and the two sequential passes of the Riccati solve, one stage at a time,
// all 256 threads of the CTA on one stage (included by solver_dev.cuh).
//
// Stage matrix layout: the 48 stage variables are stored in ELIMINATION order
//     m = 0..11 f_k | 12..23 c_{k+1} | 24..35 X_k | 36..47 c_k          (m = (s + 24) mod 48)
// and only the lower triangle M[a][b], a >= b, is kept.  One blocked right-looking partial Cholesky
// (24 pivots, blocks of 4, the stage gradient carried as row 48) then leaves, in place,
//     L (control block) | Yt = M_xu L^-T (rows 24..47, cols 0..23) | P_k = M_xx - Yt Yt' | yv, p_k
// i.e. oracle/ip_ref.c: riccati_factor + the backward half of riccati_solve in one pass.
#pragma once
// (textually included inside namespace srb { namespace { ... } } by solver_dev.cuh)

// ---- static work tables (built once per CTA in shared memory)
//   [0,78)    symmetric G'(Pxx G) tiles (ti >= tj), 3x3, G-column tile coordinates 0..11
//   [78,126)  cross tiles G' Pxc: (ti 0..11, tc 0..3)
//   [ch_off(b), ch_off(b+1))  trailing work items of block step b (see partial_cholesky)
#ifndef SRB_NB
#define SRB_NB 4
#endif
constexpr int NB = SRB_NB;                    // pivot block size
// the Riccati factors are written once per stage and read once by the forward sweep: streaming stores (evict-first) keep
// the L2 for the row arrays and the iterate, which every pass of an iteration touches
#ifdef SRB_NO_STREAM_STORES
#define ST_STREAM(p, v) (*(p) = (v))
#else
#define ST_STREAM(p, v) __stcs((p), (v))
#endif
constexpr int NBLK = NS / NB;                 // block steps
// (row r | column group g << 8): r in [i0, 48] (48 = gradient row), columns i0 + 4g .. i0 + 4g + 3, i0 + 4g <= r;
// ordered group by group: a warp reads ONE column group (broadcast) and consecutive rows (conflict-free)
__host__ __device__ constexpr int ch_count(int b) {
  const int i0 = NB * (b + 1), ng = (NW - i0) / 4;
  int n = 0;
  for (int g = 0; g < ng; g++) n += (NW + 1) - (i0 + 4 * g);
  return n;
}
__host__ __device__ constexpr int ch_off(int b) {
  int o = 126;
  for (int i = 0; i < b; i++) o += ch_count(i);
  return o;
}
constexpr int TL_COUNT = ch_off(NBLK);
static_assert(NBLK <= 6 && NB % 4 == 0, "c_ch_off lists the offsets of up to six block steps; column groups are 4 wide");
__constant__ int c_ch_off[7] = {ch_off(0), ch_off(1), ch_off(2), ch_off(3), ch_off(4), ch_off(5), ch_off(6)};
static_assert(TL_COUNT <= TL_WORDS * 4, "work table does not fit its shared-memory region");

__device__ void build_tile_tables(unsigned short* tl) {
  const int tid = TID;
  if (tid == 0) {
    int n = 0;
    for (int ti = 0; ti < 12; ti++)
      for (int tj = 0; tj <= ti; tj++) tl[n++] = (unsigned short)(ti | (tj << 8));
    for (int ti = 0; ti < 12; ti++)
      for (int tc = 0; tc < 4; tc++) tl[n++] = (unsigned short)(ti | (tc << 8));
  } else if (tid <= NBLK) {
    const int b = tid - 1, i0 = NB * (b + 1);
    int n = ch_off(b);
    const int ng = (NW - i0) / 4;
    for (int g = 0; g < ng; g++)
      for (int r = i0 + 4 * g; r <= NW; r++) tl[n++] = (unsigned short)(r | (g << 8));
  }
}

__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src) {  // 16 bytes, L2 only
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// stage lists of knot k -> list buffer (asynchronously, 16-byte chunks)
__device__ __forceinline__ void prefetch_lists(const Ws& w, int k, double* lb) {
  const int tid = TID, rb = 36 + RK * k;
  const double* Jk = w.JL + (long long)k * NJ_PAD;
  const double* Hk = w.HL + (long long)k * NH_PAD;
  if (tid < NJ_PAD / 2) cp_async16(lb + LB_J + 2 * tid, Jk + 2 * tid);
  else if (tid < NJ_PAD / 2 + RK / 2) { const int i = tid - NJ_PAD / 2; cp_async16(lb + LB_SIG + 2 * i, w.SIG + rb + 2 * i); }
  if (tid < NH_PAD / 2) cp_async16(lb + LB_H + 2 * tid, Hk + 2 * tid);
  else if (tid < NH_PAD / 2 + RK / 2) { const int i = tid - NH_PAD / 2; cp_async16(lb + LB_YH + 2 * i, w.YH + rb + 2 * i); }
  else if (tid < NH_PAD / 2 + RK / 2 + 6) { const int i = tid - NH_PAD / 2 - RK / 2; cp_async16(lb + LB_GD + 2 * i, w.G + rb + 2 * i); }
  cp_async_commit();
}

// G-column tile (0..11: X,c,f) -> tile index in elimination order
__device__ __forceinline__ int rot_tile(int t) { return t < 8 ? t + 8 : t - 8; }

// 1/sqrt(e) for a positive, normal e: MUFU.RSQ64H seed (22 bits) + one third-order correction; 52 cycles on the
// dependent chain instead of the library's 66 (tools/ubench_fp64.cu), max relative error 2.7e-16.
__device__ __forceinline__ double rsqrt_pos(double a) {
#ifdef SRB_LIB_RSQRT
  return rsqrt(a);
#else
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
  const double t = a * y0, r = fma(-t, 0.5 * y0, 0.5);
  const double q = fma(r, 1.5, 1.0), yr = y0 * r;
  return fma(yr, q, y0);
#endif
}

// Blocked partial Cholesky of the lower-stored 48x48 matrix M (+ gradient row qh), 24 pivots.
// Per block step: (1) the NB x NB diagonal block is factored in registers (right-looking: every pivot hangs on a
// chain of ~6 FP64 operations) and, fused with it column by column, each panel row (rows below the block; row 48 =
// gradient) is solved against L_D^T; (2) all threads update the trailing lower triangle.
// false -> a pivot was not positive (wrong inertia).
__device__ __forceinline__ bool partial_cholesky(double* M, double* qh, const unsigned short* tl, int* s_bad, Prof& pf) {
  const int tid = TID;
#ifdef SRB_CHOL_UNROLL
#pragma unroll
#else
#pragma unroll 1  // rolled: the unrolled stage loop body (40 KB of code) does not stay in the instruction cache
#endif
  for (int b = 0; b < NBLK; b++) {
    const int p0 = NB * b, i0 = p0 + NB;
    double L[NB][NB];
    if (tid < 64) {
      // every thread of the two panel warps repeats the small factorisation (cheaper than a broadcast through shared
      // memory); the other six warps skip it, so the FP64 pipes of their sub-partitions stay free for the co-resident CTA
      double x[NB];
#pragma unroll
      for (int i = 0; i < NB; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) L[i][j] = M[(p0 + i) * LDM + p0 + j];
      const bool panel = tid < NW + 1 - i0;
      double* arow = (i0 + tid < NW) ? (M + (i0 + tid) * LDM + p0) : (qh + p0);
#pragma unroll
      for (int j = 0; j < NB; j++) x[j] = panel ? arow[j] : 0.0;
      bool pd = true;
#pragma unroll
      for (int j = 0; j < NB; j++) {
        const double e = L[j][j];
        pd = pd && (e > 1e-14);
        const double r = rsqrt_pos(e);
        L[j][j] = r;  // the diagonal keeps 1/l_jj (what the solves need)
        const double xj = x[j] * r;
        x[j] = xj;
#pragma unroll
        for (int i = j + 1; i < NB; i++) L[i][j] *= r;
#pragma unroll
        for (int i = j + 1; i < NB; i++) {
#pragma unroll
          for (int l = j + 1; l <= i; l++) L[i][l] -= L[i][j] * L[l][j];
          x[i] -= xj * L[i][j];
        }
      }
      if (panel) {
#pragma unroll
        for (int j = 0; j < NB; j++) arow[j] = x[j];
      }
      if (!pd && tid == 0) *s_bad = 1;
    }
    __syncthreads();
    if (*s_bad) return false;  // block-uniform
    pf.lap(PH_C_DIAG);
    if (tid == 63) {  // the block's own factor (nobody reads the diagonal block during the trailing update)
#pragma unroll
      for (int i = 0; i < NB; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) M[(p0 + i) * LDM + p0 + j] = L[i][j];
    }
#ifdef SRB_CHOL_UNROLL
    const int cnt = ch_count(b);
    const unsigned short* list = tl + ch_off(b);
#else
    const int cnt = c_ch_off[b + 1] - c_ch_off[b];
    const unsigned short* list = tl + c_ch_off[b];
#endif
    // M[r][c] -= panel_r . panel_c for c in the item's column group, c <= r
    for (int i = tid; i < cnt; i += NT) {
      const int e = list[i], r = e & 255, c0 = i0 + 4 * (e >> 8);
      const double* xr = (r < NW) ? (M + r * LDM + p0) : (qh + p0);
      const double* xc = M + c0 * LDM + p0;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const double u = xr[q];
        s0 += u * xc[q]; s1 += u * xc[LDM + q]; s2 += u * xc[2 * LDM + q]; s3 += u * xc[3 * LDM + q];
      }
      double* o = (r < NW) ? (M + r * LDM + c0) : (qh + c0);
      o[0] -= s0;
      if (c0 + 1 <= r) o[1] -= s1;
      if (c0 + 2 <= r) o[2] -= s2;
      if (c0 + 3 <= r) o[3] -= s3;
    }
    __syncthreads();
    pf.lap(PH_C_TRAIL);
  }
  return true;
}

// ---------------------------------------------------------------- condensing pre-pass (parallel over stages)
// Everything of a stage that does not depend on P_{k+1}: the lower-triangle sums H + J_I' Sigma J_I per target, the stage
// gradient q = J_I' yhat, the structural entries of G and the defects r.  Done once per iteration for all stages with all
// 256 threads busy (4-deep cp.async ring over the stage lists), instead of inside the sequential sweep (and its retries).
template <int PENDING> __device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory");
}

__device__ __noinline__ void condense_all(const KParams& P, const Ws& w, double* smem) {
  const int K = P.K, tid = TID;
  const int* tbl = reinterpret_cast<const int*>(smem + SM_TBL);
  const SolverTables& tb = P.tab;
  const int* t_g = tbl + tb.o_g;
  const int* t_qptr = tbl + tb.o_qptr;
  const int* t_qterms = tbl + tb.o_qterms;
  const int* t_uabh = tbl + tb.o_uabh;
  const int* t_uptr = tbl + tb.o_uptr;
  const int* t_uterms = tbl + tb.o_uterms;
  const bool run = has_run_cost(P);
  const double hh = 2.0 * P.pb.T / (double)(P.N - 1), rq0 = P.pb.Qf[0] * hh, rq1 = P.pb.Qf[1] * hh, rq2 = P.pb.Qf[2] * hh;
  auto run2h = [&](int j) { return j % 3 == 0 ? rq0 : (j % 3 == 1 ? rq1 : rq2); };  // 2 Qf dt of force component j
  // 4-deep ring of list buffers over regions that are idle before the sweeps: list region | stage matrix (2) | P,G,T
  auto ring = [&](int k) {
    const int j = k & 3;
    return smem + (j == 0 ? SM_LB0 : (j == 1 ? SM_M : (j == 2 ? SM_M + LB_SIZE : SM_P)));
  };
  static_assert(2 * LB_SIZE <= NW * LDM && LB_SIZE <= LB_REGION && LB_SIZE <= NS * LDP + 2 * 12 * LDG, "ring buffers fit");
  for (int k = 0; k < 3; k++) {
    if (k < K) prefetch_lists(w, k, ring(k)); else cp_async_commit();
  }
  for (int k = 0; k < K; k++) {
    if (k + 3 < K) prefetch_lists(w, k + 3, ring(k + 3)); else cp_async_commit();
    cp_async_wait_group<3>();
    __syncthreads();
    const double* lb = ring(k);
    const double* Js = lb + LB_J;
    const double* Hs = lb + LB_H;
    const double* SIGs = lb + LB_SIG;
    const double* YHs = lb + LB_YH;
    double* ct = w.CT + (long long)k * CT_STRIDE;
#pragma unroll
    for (int rnd = 0; rnd < 2; rnd++) {
      // second round in reverse thread order: the threads that had the long targets get the trivial items
      const int it = rnd == 0 ? tid : 2 * NT - 1 - tid;
      if (it < tb.n_u) {
        const int abh = t_uabh[it], h = abh >> 12;
        double a0 = h ? Hs[h - 1] : 0.0, a1 = 0.0;
        const int p0 = t_uptr[it], p1 = t_uptr[it + 1];
        for (int p = p0; p < p1; p += 2) {
          const int u0 = t_uterms[p], u1 = t_uterms[p + 1];
          a0 += SIGs[u0 >> 20] * Js[(u0 >> 10) & 1023] * Js[u0 & 1023];
          a1 += SIGs[u1 >> 20] * Js[(u1 >> 10) & 1023] * Js[u1 & 1023];
        }
        if (run) {  // Hessian diagonal of the running GRF cost (force variables are m = 0..11)
          const int ab = abh & 4095, ta = ab / NW;
          if (ta < 12 && ab == ta * NW + ta) a0 += run2h(ta);
        }
        ST_STREAM(&ct[it], a0 + a1);
      } else if (it < tb.n_u + NW) {
        const int m = it - tb.n_u;
        double a0 = 0.0, a1 = 0.0;
        const int p0 = t_qptr[m], p1 = t_qptr[m + 1];
        for (int p = p0; p < p1; p += 2) {
          const int u0 = t_qterms[p], u1 = t_qterms[p + 1];
          a0 += YHs[u0 >> 10] * Js[u0 & 1023];
          a1 += YHs[u1 >> 10] * Js[u1 & 1023];
        }
        if (run && m < 12) a0 += run2h(m) * w.x[12 * P.N + 24 * k + 12 + m];  // gradient of the running GRF cost
        ST_STREAM(&ct[CT_Q + m], a0 + a1);
      } else if (it < tb.n_u + NW + tb.n_g) {
        const int n = it - tb.n_u - NW;
        ST_STREAM(&ct[CT_G + n], -Js[t_g[n] & 1023]);
      } else if (it < tb.n_u + NW + tb.n_g + 12) {
        const int i = it - tb.n_u - NW - tb.n_g;
        ct[CT_R + dyn_state(i)] = -lb[LB_GD + i];
      }
    }
    __syncthreads();  // the buffer is refilled three stages later
  }
  cp_async_wait_group<0>();
}

// condensed data of stage k -> shared memory (asynchronously, 16-byte chunks)
__device__ __forceinline__ void prefetch_ct(const Ws& w, int k, double* cb) {
  const int tid = TID;
  if (tid < CT_STRIDE / 2) cp_async16(cb + 2 * tid, w.CT + (long long)k * CT_STRIDE + 2 * tid);
  cp_async_commit();
}

// ---------------------------------------------------------------- backward sweep
// Condenses every stage from the entry lists, factors it and propagates P, p.  false -> not PD.
__device__ __noinline__ bool backward_sweep(const KParams& P, const Ws& w, double* smem, double dwreg) {
  const int N = P.N, K = P.K, tid = TID;
  double* M = smem + SM_M;
  double* Pn = smem + SM_P;
  double* Gs = smem + SM_G;
  double* Ts = smem + SM_T;
  double* V = smem + SM_V;
  const unsigned short* tl = reinterpret_cast<const unsigned short*>(smem + SM_TL);
  const int* tbl = reinterpret_cast<const int*>(smem + SM_TBL);
  const SolverTables& tb = P.tab;
  __shared__ int s_ok, s_bad;
  if (TID == 0) s_bad = 0;  // (visible after the barriers below)
  const int* t_g = tbl + tb.o_g;
  const int* t_uabh = tbl + tb.o_uabh;
  // this thread's GEMM tile
  int g_ti = 0, g_tj = 0;
  if (tid < 126) { const int e = tl[tid]; g_ti = e & 255; g_tj = e >> 8; }
  prefetch_ct(w, K - 1, smem + SM_LB0 + ((K - 1) & 1) * CT_STRIDE);
  // terminal block P_K, p_K; G's zero pattern is set once (only its structural entries change)
  for (int i = tid; i < NS * LDP; i += NT) Pn[i] = 0.0;
  for (int i = tid; i < 12 * LDG; i += NT) Gs[i] = 0.0;
  if (tid < NS) V[V_PN + tid] = 0.0;
  __syncthreads();
  if (tid < 12) {
    const int r1 = tid < 6 ? 12 + tid : 24 + (tid - 6), r2 = r1 + 6;
    const double q = w.x[12 * (N - 1) + tid];
    const double ref = tid < 6 ? P.pb.q_term_ref[tid] : P.pb.qd_term_ref[tid - 6];
    Pn[tid * LDP + tid] = 2.0 * P.pb.QN[tid] + w.SIG[r1] + w.SIG[r2] + dwreg;
    V[V_PN + tid] = 2.0 * P.pb.QN[tid] * (q - ref) + w.YH[r1] + w.YH[r2];
  }
  __syncthreads();
  for (int i = tid; i < 288; i += NT) w.PX[(long long)K * 288 + i] = Pn[(i / 24) * LDP + (i % 24)];
  if (tid < NS) w.PV[K * 24 + tid] = V[V_PN + tid];

  Prof pf{P.prof, 0};
  pf.start();
  for (int k = K - 1; k >= 0; k--) {
    pf.count(PH_B_STAGES);
    const double* cb = smem + SM_LB0 + (k & 1) * CT_STRIDE;  // condensed data of this stage (condense_all)
    cp_async_wait_all();
    __syncthreads();  // condensed data of stage k is in shared memory; P_{k+1}, p_{k+1} complete
    pf.lap(PH_B_WAIT);
    if (k > 0) prefetch_ct(w, k - 1, smem + SM_LB0 + ((k - 1) & 1) * CT_STRIDE);
    // P1. dynamics Jacobian G (structural entries), defects r
    if (tid < tb.n_g) {
      const int t = t_g[tid] >> 10;
      Gs[(t / 36) * LDG + (t % 36)] = cb[CT_G + tid];
    } else if (tid >= 224 && tid < 236) {
      V[V_R + tid - 224] = cb[CT_R + tid - 224];
    }
    __syncthreads();
    pf.lap(PH_B_P1);
    // P2. T = Pxx G (12 x 36; thread (i, jj) -> columns jj, jj+9, jj+18, jj+27: conflict free), t = Pxx r + p_x
    if (tid < 108) {
      const int i = tid / 9, jj = tid - i * 9;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
      for (int l = 0; l < 12; l++) {
        const double pv = Pn[i * LDP + l];
        const double* g = Gs + l * LDG + jj;
        s0 += pv * g[0]; s1 += pv * g[9]; s2 += pv * g[18]; s3 += pv * g[27];
      }
      double* t = Ts + i * LDG + jj;
      t[0] = s0; t[9] = s1; t[18] = s2; t[27] = s3;
    } else if (tid < 120) {
      const int i = tid - 108;
      double s = V[V_PN + i];
#pragma unroll
      for (int l = 0; l < 12; l++) s += Pn[i * LDP + l] * V[V_R + l];
      V[V_T + i] = s;
    }
    __syncthreads();
    pf.lap(PH_B_P2);
    // P3. M (lower, elimination order) = [G'PxxG, G'Pxc; ., Pcc]; qh = q + G't (+ p_c + Pcx r)
    if (tid < 126) {
      double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      const double* ga = Gs + 3 * g_ti;
      const bool sym = tid < 78;
      const double* bb = sym ? (Ts + 3 * g_tj) : (Pn + 12 + 3 * g_tj);
      const int ldb = sym ? LDG : LDP;
#pragma unroll
      for (int l = 0; l < 12; l++) {
        const double a0 = ga[l * LDG], a1 = ga[l * LDG + 1], a2 = ga[l * LDG + 2];
        const double b0 = bb[l * ldb], b1 = bb[l * ldb + 1], b2 = bb[l * ldb + 2];
        acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[0][2] += a0 * b2;
        acc[1][0] += a1 * b0; acc[1][1] += a1 * b1; acc[1][2] += a1 * b2;
        acc[2][0] += a2 * b0; acc[2][1] += a2 * b1; acc[2][2] += a2 * b2;
      }
      const int ri = 3 * rot_tile(g_ti), rj = sym ? 3 * rot_tile(g_tj) : 12 + 3 * g_tj;
      if (ri > rj) {
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int c = 0; c < 3; c++) M[(ri + a) * LDM + rj + c] = acc[a][c];
      } else if (ri < rj) {
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int c = 0; c < 3; c++) M[(rj + c) * LDM + ri + a] = acc[a][c];
      } else {
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int c = 0; c <= a; c++) M[(ri + a) * LDM + rj + c] = acc[a][c];
      }
    } else if (tid >= 128 && tid < 206) {  // Pcc, lower triangle
      int i = 0, rem = tid - 128;
      while (rem > i) { rem -= i + 1; i++; }
      M[(12 + i) * LDM + 12 + rem] = Pn[(12 + i) * LDP + 12 + rem];
    } else if (tid >= 208) {
      const int m = tid - 208;  // elimination-order index
      double s = cb[CT_Q + m];
      if (m >= 12 && m < 24) {
        const int j = m - 12;
        s += V[V_PN + 12 + j];
#pragma unroll
        for (int l = 0; l < 12; l++) s += Pn[(12 + j) * LDP + l] * V[V_R + l];
      } else {
        const int g = m < 12 ? m + 24 : m - 24;  // G column of this variable
#pragma unroll
        for (int l = 0; l < 12; l++) s += Gs[l * LDG + g] * V[V_T + l];
      }
      V[V_QH + m] = s;
    }
    __syncthreads();
    pf.lap(PH_B_P3);
    // P4. add the condensed sums (+ delta_w on the (f, X, c) diagonal, dummy c+ of the last stage)
#pragma unroll
    for (int rnd = 0; rnd < 2; rnd++) {
      const int t = tid + NT * rnd;
      if (t < tb.n_u) {
        const int ab = t_uabh[t] & 4095, a = ab / NW, b2 = ab - a * NW;
        double acc = cb[t];
        if (a == b2) acc += (a >= 12 && a < 24) ? (k == K - 1 ? 1.0 : 0.0) : dwreg;
        M[a * LDM + b2] += acc;
      }
    }
    __syncthreads();
    pf.lap(PH_B_P4);
    // P5. eliminate the controls
    if (!partial_cholesky(M, V + V_QH, tl, &s_bad, pf)) {
      cp_async_wait_all();  // no prefetch may still be in flight when the sweep is retried
      __syncthreads();
      return false;
    }
    // P6. P_k, p_k, yv and what the forward sweep needs.  Thread (row = tid / 8, three columns from 3 (tid % 8)):
    // no divisions, every loop fully unrolled, consecutive threads on consecutive addresses.
    {
      const int i = tid >> 3, c = 3 * (tid & 7);
      double* FY = w.FY + (long long)k * 1152;
      if (i < NS) {  // P_k from the lower triangle of the Schur complement (rows 24..47 of M)
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const int j = c + d, a = i < j ? j : i, b2 = i < j ? i : j;
          const double v = M[(24 + a) * LDM + 24 + b2];
          Pn[i * LDP + j] = v;
          if (i < 12) ST_STREAM(&w.PX[(long long)k * 288 + i * NS + j], v);
        }
      }
      // rows 0-23: L (strict lower, 1/l_ii on the diagonal); rows 24-47: Yt
#pragma unroll
      for (int d = 0; d < 3; d++) ST_STREAM(&FY[i * NS + c + d], M[i * LDM + c + d]);
      if (i < 16) {
#pragma unroll
        for (int d = 0; d < 3; d++) ST_STREAM(&FY[(32 + i) * NS + c + d], M[(32 + i) * LDM + c + d]);
      }
      if (tid >= 224 && tid < 224 + NS) {
        const int t = tid - 224;
        V[V_PN + t] = V[V_QH + 24 + t];
        w.PV[k * 24 + t] = V[V_QH + 24 + t];
        w.yvf[k * 24 + t] = V[V_QH + t];
      } else if (tid >= 192 && tid < 204) {
        w.rf[k * 12 + tid - 192] = V[V_R + tid - 192];
      }
    }
    pf.lap(PH_B_P6);
  }
  // free initial foot positions: Cholesky of P_0's (c,c) block (12 x 12) by warp 0, 1/l_ii on the diagonal
  __syncthreads();
  if (tid == 0) s_ok = 1;
  for (int idx = tid; idx < 144; idx += NT) M[(idx / 12) * LDM + (idx % 12)] = Pn[(12 + idx / 12) * LDP + 12 + (idx % 12)];
  __syncthreads();
  if (tid < 32) {
    const int lane = tid;
    bool ok = true;
    for (int j = 0; j < 12 && ok; j++) {
      double v = 0.0;
      if (lane >= j && lane < 12) {
        v = M[lane * LDM + j];
        for (int l = 0; l < j; l++) v -= M[lane * LDM + l] * M[j * LDM + l];
      }
      const double d = __shfl_sync(FULL, v, j);
      if (!(d > 1e-14)) { ok = false; break; }
      const double rs = rsqrt(d);
      if (lane == j) M[j * LDM + j] = rs;
      else if (lane > j && lane < 12) M[lane * LDM + j] = v * rs;
      __syncwarp();
    }
    if (!ok && lane == 0) s_ok = 0;
  }
  __syncthreads();
  if (!s_ok) return false;
  for (int idx = tid; idx < 144; idx += NT) w.L0[idx] = M[(idx / 12) * LDM + (idx % 12)];
  __syncthreads();  // Pn still holds P_0, V_PN p_0 for the forward start
  return true;
}

// ---------------------------------------------------------------- forward sweep
// forward-stage buffer (doubles): FY 48x24 (L | Yt) | J list 388 | yv 24 | r 12 ; G (12x37) is rebuilt per stage
constexpr int FB_FY = 0, FB_J = 1152, FB_YV = FB_J + NJ_PAD, FB_R = FB_YV + NS, FB_SIZE = FB_R + 12;
static_assert(FB_SIZE <= NW * LDM && FB_SIZE <= 2 * 12 * LDG + LB_REGION, "forward buffers alias the backward regions");

__device__ __forceinline__ void prefetch_factors(const Ws& w, int k, double* fb) {
  const int tid = TID;
  const double* FY = w.FY + (long long)k * 1152;
  const double* Jk = w.JL + (long long)k * NJ_PAD;
  for (int i = tid; i < 576; i += NT) cp_async16(fb + FB_FY + 2 * i, FY + 2 * i);
  if (tid < NJ_PAD / 2) cp_async16(fb + FB_J + 2 * tid, Jk + 2 * tid);
  else if (tid < NJ_PAD / 2 + 12) { const int i = tid - NJ_PAD / 2; cp_async16(fb + FB_YV + 2 * i, w.yvf + k * 24 + 2 * i); }
  else if (tid < NJ_PAD / 2 + 18) { const int i = tid - NJ_PAD / 2 - 12; cp_async16(fb + FB_R + 2 * i, w.rf + k * 12 + 2 * i); }
  cp_async_commit();
}

// dx for all stages; equality multipliers of the initial-state rows into YN (dynamics costates: costates())
__device__ __noinline__ void forward_sweep(const KParams& P, const Ws& w, double* smem, const double* drop) {
  const int N = P.N, K = P.K, tid = TID, lane = tid & 31, warp = tid >> 5;
  double* fbuf[2] = {smem + SM_M, smem + SM_G};
  double* Pn = smem + SM_P;  // holds P_0 on entry; afterwards the dense G of the current stage
  double* V = smem + SM_V;
  double* xi = V + V_XI;
  double* u = V + V_U;
  double* rhs = V + V_Q;     // 24
  const int* t_g = reinterpret_cast<const int*>(smem + SM_TBL) + P.tab.o_g;
  // initial state step and free initial feet
  if (tid < 12) xi[tid] = -(w.G[tid] - drop[tid]);
  __syncthreads();
  if (warp == 0) {
    // b = -(p_c + P_cx dX0); solve L0 L0' dc0 = b   (12 x 12; lanes 0..11 own rows, shuffles broadcast)
    double b = 0.0;
    if (lane < 12) {
      b = -V[V_PN + 12 + lane];
      for (int l = 0; l < 12; l++) b -= Pn[(12 + lane) * LDP + l] * xi[l];
    }
    for (int i = 0; i < 12; i++) {  // forward
      const double bi = __shfl_sync(FULL, b, i) * w.L0[i * 12 + i];
      if (lane == i) b = bi;
      else if (lane > i && lane < 12) b -= w.L0[lane * 12 + i] * bi;
    }
    for (int i = 11; i >= 0; i--) {  // backward
      const double bi = __shfl_sync(FULL, b, i) * w.L0[i * 12 + i];
      if (lane == i) b = bi;
      else if (lane < i) b -= w.L0[i * 12 + lane] * bi;
    }
    if (lane < 12) xi[12 + lane] = b;
  }
  __syncthreads();
  if (tid < 12) {  // multipliers of the initial-state rows: -dV0/dX
    double v = V[V_PN + tid];
    for (int l = 0; l < NS; l++) v += Pn[tid * LDP + l] * xi[l];
    w.YN[tid] = -v;
    w.DS[tid] = 0.0;
  }
  __syncthreads();  // P_0 consumed: every backward region may now be overwritten
  prefetch_factors(w, 0, fbuf[0]);
  double* Gs = Pn;
  for (int i = tid; i < 12 * LDG; i += NT) Gs[i] = 0.0;
  for (int k = 0; k < K; k++) {
    const bool last = (k == K - 1);
    const double* fb = fbuf[k & 1];
    const double* Ls = fb + FB_FY;             // 24 x 24
    const double* Ys = fb + FB_FY + NS * NS;   // Yt[i][c]
    cp_async_wait_all();
    __syncthreads();  // factors of stage k in shared memory, xi complete, previous G no longer read
    if (!last) prefetch_factors(w, k + 1, fbuf[(k + 1) & 1]);
    // rhs = -(Y xi + yv): thread (c, part) sums 6 terms, the 4 parts sit in neighbouring lanes | G | dx
    if (tid < 96) {
      const int c = tid >> 2, p = tid & 3;
      double v = 0.0;
#pragma unroll
      for (int ii = 0; ii < 6; ii++) v += Ys[(4 * ii + p) * NS + c] * xi[4 * ii + p];
      v += __shfl_xor_sync(FULL, v, 1);
      v += __shfl_xor_sync(FULL, v, 2);
      if (p == 0) rhs[c] = -(v + fb[FB_YV + c]);
    } else if (tid < 96 + P.tab.n_g) {
      const int e = t_g[tid - 96], t = e >> 10;
      Gs[(t / 36) * LDG + (t % 36)] = -fb[FB_J + (e & 1023)];
    } else if (tid >= 240 && tid < 252) {  // step of this knot's state / foot variables
      const int i = tid - 240;
      w.dx[12 * k + i] = xi[i];
      w.dx[12 * N + 24 * k + i] = xi[12 + i];
    }
    __syncthreads();
    if (warp == 0) {
      // u = L^-T rhs (diagonal of Ls holds 1/l_ii)
      double my = lane < NS ? rhs[lane] : 0.0;
#pragma unroll  // (rolled, the loads of L sit on the dependent chain: +15 % per iteration)
      for (int i = NS - 1; i >= 0; i--) {
        const double ui = __shfl_sync(FULL, my * Ls[i * NS + i], i);
        if (lane == i) my = ui;
        else if (lane < i) my -= Ls[i * NS + lane] * ui;
      }
      if (lane < NS) u[lane] = my;
      if (lane < 12) w.dx[12 * N + 24 * k + 12 + lane] = my;
      __syncwarp();
      // next state: lanes (row, half) sum 18 terms each
      double xn = 0.0;
      if (lane < NS) {
        const int i = lane >> 1, hf = lane & 1;
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < 18; l++) {
          const int j = 18 * hf + l;
          v += Gs[i * LDG + j] * (j < NS ? xi[j] : u[j - NS]);
        }
        v += __shfl_xor_sync(0x00ffffffu, v, 1);
        xn = v + fb[FB_R + i];
      }
      __syncwarp();
      if (lane < NS && (lane & 1) == 0) xi[lane >> 1] = xn;
      if (lane >= 12 && lane < NS) xi[lane] = last ? 0.0 : u[lane];
    }
  }
  __syncthreads();
  if (tid < 12) w.dx[12 * (N - 1) + tid] = xi[tid];
  __syncthreads();
}

// costates = multipliers of the dynamics rows of knot k: -(P_{k+1} [dX_{k+1}; dc_{k+1}] + p_{k+1}), all knots in parallel
__device__ __noinline__ void costates(const KParams& P, const Ws& w) {
  const int N = P.N, K = P.K;
  for (int item = TID; item < K * 12; item += NT) {
    const int k = item / 12, i = item - k * 12;
    const double* PX = w.PX + (long long)(k + 1) * 288 + i * 24;
    double v = w.PV[(k + 1) * 24 + i];
#pragma unroll
    for (int l = 0; l < 12; l++) v += PX[l] * w.dx[12 * (k + 1) + l];
    if (k + 1 < K) {
#pragma unroll
      for (int l = 0; l < 12; l++) v += PX[12 + l] * w.dx[12 * N + 24 * (k + 1) + l];
    }
    const int rho = i < 6 ? i : (i < 9 ? i + 3 : i - 3);
    w.YN[36 + RK * k + rho] = -v;
    w.DS[36 + RK * k + rho] = 0.0;
  }
  __syncthreads();
}

