// vfs_solver.h — device-resident matrix-free Newton-Krylov momentum solve.
//
// Replaces the SNESSolve of Implicit_MatrixFree (Source/implicitsolver.c:4203-4299): SNES trust region
// (SNESTR, :4247) with Eisenstat-Walker version 3 forcing (:4251-4252), a matrix-free Jacobian by forward
// differences of FormFunction_SNES (MatCreateSNESMF, :4242-4243), restarted GMRES without preconditioner
// (KSPGMRES / PCNONE, :4260,4273), tolerances of :4254 and :4277-4279.  The solver itself lives in PETSc 3.1
// (Source/makefile:26-28), which is NOT in the reference tree: its published algorithms are restated here —
// SNESSolve_TR with the More' step-length test on the Krylov iterates, MatMFFD "wp" differencing
// h = error_rel * sqrt(1 + |u|) / |a|, classical Gram-Schmidt GMRES with Givens rotations, the default
// SNES/KSP convergence tests — and the same restatement in numpy (oracle/newton_krylov_ref.py, test
// infrastructure) drives the oracle residual for the parity test.  Parity against PETSc's own iterates is unpinned.
//
// Everything stays in HBM: the Krylov basis, U, F and the work vectors are flat [nzl][my][mx][3] arrays (the layout
// of the reference's global Vec restricted to the slab); a residual evaluation is UnpackX -> the residual kernels ->
// PackAoS, replayed as one CUDA graph; dot products are two-stage deterministic reductions, summed over the ranks
// with ncclAllReduce; only a handful of scalars per Krylov iteration reach the host.
#ifndef VFS_SOLVER_H
#define VFS_SOLVER_H
#include <vector>
#include <math.h>

#define VFS_KS_MAXV 8           // vectors per batched-dot / multi-axpy launch
#define VFS_KS_BLOCKS 592       // 4 blocks per SM on 148 SMs: fixed, so reductions are reproducible

struct VfsSolver {
  long n = 0;                   // doubles per vector (this rank)
  int cap = 0;                  // Krylov vectors allocated (restart + 1)
  std::vector<double *> V;
  double *U = nullptr, *F = nullptr, *G = nullptr, *Y = nullptr, *Yt = nullptr, *W = nullptr, *Up = nullptr;
  double *partial = nullptr, *result = nullptr;     // device: [VFS_KS_MAXV][VFS_KS_BLOCKS], [VFS_KS_MAXV]
  double *h_result = nullptr;                       // pinned host copy of `result`
  long evals = 0;
};

#ifndef VFS_EMU
struct KsPtrs { const double *p[VFS_KS_MAXV]; };
struct KsCoef { double a[VFS_KS_MAXV]; };
// partial[q][block] = sum over the block's grid-stride elements of V_q[t] * w[t].  FORM: w[t] = a x[t] + b y[t] is
// formed (and stored) on the fly — the matrix-free product (F(u + h v) - F(u)) / h never makes a pass of its own.
template <bool FORM> __global__ void __launch_bounds__(256) k_ks_mdot(KsPtrs V, int nv, double *__restrict__ w, long n, double *__restrict__ partial,
                                                                      double a, const double *__restrict__ x, double b, const double *__restrict__ y) {
  double acc[VFS_KS_MAXV];
#pragma unroll
  for (int q = 0; q < VFS_KS_MAXV; q++) acc[q] = 0;
  for (long t = (long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long)gridDim.x * 256) {
    double wv;
    if (FORM) { wv = a * x[t] + b * y[t]; w[t] = wv; } else wv = w[t];
#pragma unroll
    for (int q = 0; q < VFS_KS_MAXV; q++) if (q < nv) acc[q] += V.p[q][t] * wv;
  }
  __shared__ double sm[VFS_KS_MAXV][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < VFS_KS_MAXV; q++) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[q][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < nv) {
    double s = 0;
    for (int q = 0; q < 8; q++) s += sm[threadIdx.x][q];
    partial[(long)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
}
__global__ void k_ks_reduce(const double *__restrict__ partial, int nv, int nb, double *__restrict__ result) {
  // one warp per vector, fixed summation order
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (q >= nv) return;
  double s = 0;
  for (int b = lane; b < nb; b += 32) s += partial[(long)q * nb + b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) result[q] = s;
}
// y += sum_q a_q V_q; NORM: partial[block] = the block's share of |y|^2 after the update
template <bool NORM> __global__ void __launch_bounds__(256) k_ks_maxpy(double *__restrict__ y, KsPtrs V, KsCoef A, int nv, long n, double *__restrict__ partial) {
  double acc = 0;
  for (long t = (long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long)gridDim.x * 256) {
    double v = y[t];
#pragma unroll
    for (int q = 0; q < VFS_KS_MAXV; q++) if (q < nv) v += A.a[q] * V.p[q][t];
    y[t] = v;
    if (NORM) acc += v * v;
  }
  if (NORM) {
    __shared__ double sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int q = 0; q < 8; q++) s += sm[q]; partial[blockIdx.x] = s; }
  }
}
// w = a x + b y   (x or y may alias w)
__global__ void __launch_bounds__(256) k_ks_waxpby(double *w, double a, const double *x, double b, const double *y, long n) {
  for (long t = (long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long)gridDim.x * 256) w[t] = a * x[t] + b * y[t];
}
#endif

static int ks_alloc(vfs_ctx *c, int restart) {
  if (!c->solver) c->solver = new VfsSolver();
  VfsSolver &S = *c->solver;
  S.n = (long)c->d.nzl * c->d.my * c->d.mx * 3;
  const size_t bytes = (size_t)S.n * sizeof(double);
  auto get = [&](double **p) -> int {
    if (*p) return 0;
#ifndef VFS_EMU
    CK(cudaMalloc((void **)p, bytes));
#else
    *p = (double *)malloc(bytes);
#endif
    return 0;
  };
  RUN(get(&S.U)); RUN(get(&S.F)); RUN(get(&S.G)); RUN(get(&S.Y)); RUN(get(&S.Yt)); RUN(get(&S.W)); RUN(get(&S.Up));
  while ((int)S.V.size() < restart + 1) { double *p = nullptr; RUN(get(&p)); S.V.push_back(p); }
  S.cap = (int)S.V.size();
  if (!S.partial) {
#ifndef VFS_EMU
    CK(cudaMalloc((void **)&S.partial, sizeof(double) * VFS_KS_MAXV * VFS_KS_BLOCKS));
    CK(cudaMalloc((void **)&S.result, sizeof(double) * VFS_KS_MAXV));
    CK(cudaMallocHost((void **)&S.h_result, sizeof(double) * VFS_KS_MAXV));
#else
    S.partial = (double *)malloc(sizeof(double)); S.result = (double *)malloc(sizeof(double) * VFS_KS_MAXV); S.h_result = (double *)malloc(sizeof(double) * VFS_KS_MAXV);
#endif
  }
  return 0;
}
static void ks_free(vfs_ctx *c) {
  if (!c->solver) return;
  VfsSolver &S = *c->solver;
#ifndef VFS_EMU
  for (double *p : S.V) cudaFree(p);
  cudaFree(S.U); cudaFree(S.F); cudaFree(S.G); cudaFree(S.Y); cudaFree(S.Yt); cudaFree(S.W); cudaFree(S.Up);
  cudaFree(S.partial); cudaFree(S.result); if (S.h_result) cudaFreeHost(S.h_result);
#else
  for (double *p : S.V) free(p);
  free(S.U); free(S.F); free(S.G); free(S.Y); free(S.Yt); free(S.W); free(S.Up); free(S.partial); free(S.result); free(S.h_result);
#endif
  delete c->solver; c->solver = nullptr;
}

// device results of a reduction -> all ranks' sum on the host
static int ks_collect(vfs_ctx *c, int m, double *out) {
  VfsSolver &S = *c->solver;
#ifndef VFS_EMU
  k_ks_reduce<<<1, 32 * VFS_KS_MAXV, 0, c->stream>>>(S.partial, m, VFS_KS_BLOCKS, S.result);
  c->launches++;
  if (c->prm.nranks > 1) {
    NcclApi &N = nccl_api();
    if (!c->comm || !N.AllReduce) { set_err(c, "vfs_momentum_solve with nranks > 1 needs vfs_nccl_init (dot products are summed with ncclAllReduce)"); return VFS_ERR_HALO; }
    ncclResult_t e = N.AllReduce(S.result, S.result, (size_t)m, ncclDouble, ncclSum, c->comm, c->stream);
    if (e != ncclSuccess) { set_err(c, std::string("ncclAllReduce: ") + N.GetErrorString(e)); return VFS_ERR_HALO; }
  }
  CK(cudaMemcpyAsync(S.h_result, S.result, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int q = 0; q < m; q++) out[q] = S.h_result[q];
#else
  (void)S; (void)m; (void)out;
#endif
  return 0;
}
// out[q] = <V_q, w> over ALL ranks, q < nv (any nv: batches of VFS_KS_MAXV).  x != null: w = a x + b y is formed first
// (inside the first batch's kernel on the device).
static int ks_mdot(vfs_ctx *c, double *const *V, int nv, double *w, double *out, double a = 0, const double *x = nullptr, double b = 0, const double *y = nullptr) {
  VfsSolver &S = *c->solver;
#ifdef VFS_EMU
  if (c->prm.nranks > 1) { set_err(c, "the host emulation of vfs_momentum_solve is single rank"); return VFS_ERR_UNSUPPORTED; }
  if (x) for (long t = 0; t < S.n; t++) w[t] = a * x[t] + b * y[t];
#endif
  for (int q0 = 0; q0 < nv; q0 += VFS_KS_MAXV) {
    const int m = nv - q0 < VFS_KS_MAXV ? nv - q0 : VFS_KS_MAXV;
#ifndef VFS_EMU
    KsPtrs P; for (int q = 0; q < VFS_KS_MAXV; q++) P.p[q] = q < m ? V[q0 + q] : V[q0];
    if (x && q0 == 0) k_ks_mdot<true><<<VFS_KS_BLOCKS, 256, 0, c->stream>>>(P, m, w, S.n, S.partial, a, x, b, y);
    else k_ks_mdot<false><<<VFS_KS_BLOCKS, 256, 0, c->stream>>>(P, m, w, S.n, S.partial, 0., nullptr, 0., nullptr);
    c->launches++;
    RUN(ks_collect(c, m, out + q0));
#else
    for (int q = 0; q < m; q++) { double s = 0; for (long t = 0; t < S.n; t++) s += V[q0 + q][t] * w[t]; out[q0 + q] = s; }
#endif
  }
  return 0;
}
static int ks_norm(vfs_ctx *c, const double *x, double *out) {
  double *v[1] = {const_cast<double *>(x)}; double r = 0;
  RUN(ks_mdot(c, v, 1, const_cast<double *>(x), &r)); *out = sqrt(r); return 0;
}
// y += sum a_q V_q; norm_out != null: also |y| after the update (the same kernel's reduction)
static int ks_maxpy(vfs_ctx *c, double *y, const double *a, double *const *V, int nv, double *norm_out = nullptr) {
  VfsSolver &S = *c->solver;
  for (int q0 = 0; q0 < nv; q0 += VFS_KS_MAXV) {
    const int m = nv - q0 < VFS_KS_MAXV ? nv - q0 : VFS_KS_MAXV;
    const bool last = q0 + m >= nv;
#ifndef VFS_EMU
    KsPtrs P; KsCoef A;
    for (int q = 0; q < VFS_KS_MAXV; q++) { P.p[q] = q < m ? V[q0 + q] : V[q0]; A.a[q] = q < m ? a[q0 + q] : 0.; }
    if (norm_out && last) {
      k_ks_maxpy<true><<<VFS_KS_BLOCKS, 256, 0, c->stream>>>(y, P, A, m, S.n, S.partial);
      c->launches++;
      double r = 0; RUN(ks_collect(c, 1, &r)); *norm_out = sqrt(r);
    } else { k_ks_maxpy<false><<<VFS_KS_BLOCKS * 2, 256, 0, c->stream>>>(y, P, A, m, S.n, nullptr); c->launches++; }
#else
    for (long t = 0; t < S.n; t++) { double v = y[t]; for (int q = 0; q < m; q++) v += a[q0 + q] * V[q0 + q][t]; y[t] = v; }
    if (norm_out && last) { double s = 0; for (long t = 0; t < S.n; t++) s += y[t] * y[t]; *norm_out = sqrt(s); }
#endif
  }
  return 0;
}
static int ks_waxpby(vfs_ctx *c, double *w, double a, const double *x, double b, const double *y) {
  VfsSolver &S = *c->solver;
#ifndef VFS_EMU
  k_ks_waxpby<<<VFS_KS_BLOCKS * 2, 256, 0, c->stream>>>(w, a, x, b, y, S.n);
  c->launches++;
#else
  for (long t = 0; t < S.n; t++) w[t] = a * x[t] + b * y[t];
#endif
  return 0;
}
static int ks_copy(vfs_ctx *c, double *dst, const double *src) {
#ifndef VFS_EMU
  CK(cudaMemcpyAsync(dst, src, (size_t)c->solver->n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
#else
  memcpy(dst, src, (size_t)c->solver->n * sizeof(double));
#endif
  return 0;
}
// G = FormFunction_SNES(Up): always the same two buffers, so the evaluation replays as one CUDA graph
static int ks_residual(vfs_ctx *c) {
  VfsSolver &S = *c->solver;
  S.evals++;
  return run_graphed(c, 2, [&]() -> int {
    { UnpackX f = {c->d, S.Up}; RUN(launch(c, box_owned(c), f)); }
    RUN(snes_core(c));
    { PackAoS f = {c->d, S.G, S_R0, 3}; RUN(launch(c, box_owned(c), f)); }
    return 0;
  });
}

// GMRES(restart) on J y = b with J a = (F(u + h a) - F(u)) / h.  u = S.U, F(u) = S.F, b = S.F, solution in S.Yt.
// delta > 0: the More' test of SNES_TR_KSPConverged_Private — stop once |y_k| >= delta.
static int ks_gmres(vfs_ctx *c, const vfs_solver_params &sp, double rtol, double delta, double unorm, int *its_out, int *reason, double *rnorm_out) {
  VfsSolver &S = *c->solver;
  const int m = sp.restart;
  const double error_rel = 1.490116119384766e-08;       // sqrt(DBL_EPSILON): MatMFFD default
  const double ufact = sqrt(1.0 + unorm);                // "wp": computed once per base vector
  std::vector<double> H((size_t)(m + 1) * m, 0.), cs(m, 0.), sn(m, 0.), g(m + 1, 0.), y(m, 0.), hcol(m + 2, 0.);
  int its = 0; *reason = 0;
  double rnorm0 = 0, ttol = 0;
  bool have_base = false;       // Yt holds the solution of earlier cycles
  for (int cycle = 0;; cycle++) {
    // residual of the cycle: r = b - J Yt  (Yt = 0 in the first cycle)
    double beta;
    if (cycle == 0) { RUN(ks_copy(c, S.V[0], S.F)); }
    else {
      double yn; RUN(ks_norm(c, S.Yt, &yn));
      if (yn == 0) { RUN(ks_copy(c, S.V[0], S.F)); }
      else {
        const double h = error_rel * ufact / yn;
        RUN(ks_waxpby(c, S.Up, 1., S.U, h, S.Yt)); RUN(ks_residual(c));
        RUN(ks_waxpby(c, S.W, 1. / h, S.G, -1. / h, S.F));          // J Yt
        RUN(ks_waxpby(c, S.V[0], 1., S.F, -1., S.W));
      }
    }
    RUN(ks_norm(c, S.V[0], &beta));
    if (cycle == 0) {
      rnorm0 = beta; ttol = fmax(rtol * rnorm0, sp.ksp_atol);
      if (beta <= ttol) { *reason = beta <= sp.ksp_atol ? 3 : 2; break; }       // KSP_CONVERGED_ATOL / RTOL at iteration 0
    }
    if (!(beta == beta)) { *reason = -9; break; }
    RUN(ks_waxpby(c, S.V[0], 1. / beta, S.V[0], 0., S.V[0]));
    for (int q = 0; q <= m; q++) g[q] = 0;
    g[0] = beta;
    int k = 0; double res = beta;
    for (; k < m; k++) {
      // w = J v_k = (F(u + h v_k) - F(u)) / h with |v_k| = 1 (PETSc recomputes the norm of the freshly normalised basis
      // vector: 1 to rounding), formed inside the Gram-Schmidt dot-product kernel
      const double h = error_rel * ufact;
      RUN(ks_waxpby(c, S.Up, 1., S.U, h, S.V[k])); RUN(ks_residual(c));
      // classical Gram-Schmidt; |w| comes out of the update kernel
      RUN(ks_mdot(c, S.V.data(), k + 1, S.W, hcol.data(), 1. / h, S.G, -1. / h, S.F));
      for (int q = 0; q <= k; q++) y[q] = -hcol[q];
      double hn; RUN(ks_maxpy(c, S.W, y.data(), S.V.data(), k + 1, &hn));
      hcol[k + 1] = hn;
      // Givens rotations on the new Hessenberg column
      for (int q = 0; q < k; q++) { const double t = cs[q] * hcol[q] + sn[q] * hcol[q + 1]; hcol[q + 1] = -sn[q] * hcol[q] + cs[q] * hcol[q + 1]; hcol[q] = t; }
      const double den = sqrt(hcol[k] * hcol[k] + hcol[k + 1] * hcol[k + 1]);
      if (den == 0) { *reason = -5; break; }                  // breakdown
      cs[k] = hcol[k] / den; sn[k] = hcol[k + 1] / den;
      hcol[k] = den;
      g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k];
      for (int q = 0; q <= k; q++) H[(size_t)q * m + k] = hcol[q];
      res = fabs(g[k + 1]);
      its++;
      if (hn != 0) RUN(ks_waxpby(c, S.V[k + 1], 1. / hn, S.W, 0., S.W));
      // convergence (KSPDefaultConverged), then the trust-region step-length test on the current iterate
      if (!(res == res)) { *reason = -9; k++; break; }
      if (res <= ttol) { *reason = res <= sp.ksp_atol ? 3 : 2; k++; break; }
      if (res >= sp.ksp_dtol * rnorm0) { *reason = -4; k++; break; }
      if (its >= sp.max_krylov) { *reason = -3; k++; break; }
      if (delta > 0 && !have_base) {
        double yn2 = 0;
        for (int q = k; q >= 0; q--) { double t = g[q]; for (int r = q + 1; r <= k; r++) t -= H[(size_t)q * m + r] * y[r]; y[q] = t / H[(size_t)q * m + q]; yn2 += y[q] * y[q]; }
        if (sqrt(yn2) >= delta) { *reason = 6; k++; break; }        // KSP_CONVERGED_STEP_LENGTH
      }
      if (hn == 0) { *reason = 5; k++; break; }               // happy breakdown
    }
    // solution update: Yt += V_k y
    const int kk = k;
    for (int q = kk - 1; q >= 0; q--) { double t = g[q]; for (int r = q + 1; r < kk; r++) t -= H[(size_t)q * m + r] * y[r]; y[q] = t / H[(size_t)q * m + q]; }
    if (kk > 0) {
      if (!have_base) { RUN(ks_waxpby(c, S.Yt, y[0], S.V[0], 0., S.V[0])); if (kk > 1) RUN(ks_maxpy(c, S.Yt, y.data() + 1, S.V.data() + 1, kk - 1)); }
      else RUN(ks_maxpy(c, S.Yt, y.data(), S.V.data(), kk));
      have_base = true;
    }
    *rnorm_out = res;
    if (*reason) break;
    if (delta > 0) {        // later cycles: the step length of the accumulated iterate
      double yn; RUN(ks_norm(c, S.Yt, &yn));
      if (yn >= delta) { *reason = 6; break; }
    }
  }
  if (!have_base) { RUN(ks_waxpby(c, S.Yt, 0., S.F, 0., S.F)); }
  *its_out = its;
  return 0;
}

extern "C" int vfs_momentum_release(vfs_ctx *c) { if (!c) return VFS_ERR_ARG; RUN(vfs_sync(c)); graph_reset(c); ks_free(c); return 0; }

extern "C" int vfs_solver_defaults(vfs_solver_params *p) {
  if (!p) return VFS_ERR_ARG;
  p->max_newton = 50; p->restart = 30; p->max_krylov = 1000;            // implicitsolver.c:4254,4279; KSPGMRES default restart
  p->snes_atol = 1.e-50; p->snes_rtol = 1.e-8; p->snes_stol = 1.e-8;    // PETSc defaults; the reference passes imp_free_tol as rtol (:4254)
  p->ksp_rtol = 1.e-5; p->ksp_atol = 1.e-50; p->ksp_dtol = 1.e5;
  p->use_ew = 1; p->trust_region = 1;
  return 0;
}

extern "C" int vfs_momentum_solve(vfs_ctx *c, const vfs_solver_params *spp, vfs_solver_info *info) {
  if (!c || !spp || !info) return VFS_ERR_ARG;
  const vfs_solver_params sp = *spp;
  if (sp.restart < 1 || sp.restart > 200 || sp.max_newton < 0) { set_err(c, "bad solver parameters"); return VFS_ERR_ARG; }
  memset(info, 0, sizeof(*info));
  RUN(ks_alloc(c, sp.restart));
  VfsSolver &S = *c->solver;
  S.evals = 0;
  const int saved_graph = c->use_graph;
#ifndef VFS_EMU
  c->use_graph = (c->prm.nranks == 1 || c->comm) ? 1 : 0;
#endif
  auto finish = [&](int r) -> int { c->use_graph = saved_graph; return r; };
#define KS(x) do { int r_ = (x); if (r_) return finish(r_); } while (0)
  // U = Ucont (VecCopy(user->Ucont, U), implicitsolver.c:4297)
  { PackAoS f = {c->d, S.U, S_UC0, 3}; KS(launch(c, box_owned(c), f)); }
  // SNESSolve_TR
  const double mu = 0.25, eta = 0.75, delta0 = 0.2, delta1 = 0.3, delta2 = 0.75, delta3 = 2.0, sigma = 1.e-4, deltatol = 1.e-12;
  double fnorm, xnorm, ynorm = 0, gnorm = 0;
  KS(ks_copy(c, S.Up, S.U)); KS(ks_residual(c)); KS(ks_copy(c, S.F, S.G));
  KS(ks_norm(c, S.F, &fnorm)); KS(ks_norm(c, S.U, &xnorm));
  double delta = delta0 * xnorm;
  info->fnorm0 = fnorm; info->fnorm_history[0] = fnorm; info->n_history = 1;
  const double ttol = fnorm * sp.snes_rtol;
  int reason = 0;
  if (!(fnorm == fnorm)) reason = -4; else if (fnorm < sp.snes_atol) reason = 2;
  // Eisenstat-Walker (version 3, PETSc defaults)
  const double ew_rtol0 = 0.3, ew_rtolmax = 0.9, ew_gamma = 1.0, ew_alpha = 0.5 * (1. + sqrt(5.));
  double ew_rtol_last = 0, ew_norm_last = 0;
  int newton = 0, lits_total = 0;
  for (int it = 0; it < sp.max_newton && !reason; it++) {
    double rtol = sp.ksp_rtol;
    if (sp.use_ew) {
      if (it == 0) rtol = ew_rtol0;
      else {
        rtol = ew_gamma * pow(fnorm / ew_norm_last, ew_alpha);
        double stol = ew_gamma * pow(ew_rtol_last, ew_alpha);
        stol = fmax(rtol, stol); rtol = fmin(ew_rtol0, stol);
        stol = ew_gamma * ttol / fnorm;
        stol = fmax(rtol, stol); rtol = fmin(ew_rtol0, stol);
      }
      rtol = fmin(rtol, ew_rtolmax);
      ew_rtol_last = rtol; ew_norm_last = fnorm;
    }
    int lits = 0, kreason = 0; double krnorm = 0;
    KS(ks_gmres(c, sp, rtol, sp.trust_region ? delta : 0., xnorm, &lits, &kreason, &krnorm));       // J Yt = F
    lits_total += lits;
    if (info->n_history <= 16) info->ksp_its_history[info->n_history - 1] = lits;
    double nrm1; KS(ks_norm(c, S.Yt, &nrm1));
    bool breakout = false;
    if (!sp.trust_region) {
      KS(ks_waxpby(c, S.Y, 1., S.U, -1., S.Yt));
      KS(ks_copy(c, S.Up, S.Y)); KS(ks_residual(c));
      KS(ks_norm(c, S.G, &gnorm));
      ynorm = nrm1;
    } else {
      for (;;) {
        double nrm = nrm1, gpnorm, scale = 1.;
        if (nrm >= delta) { nrm = delta / nrm; gpnorm = (1.0 - nrm) * fnorm; scale = nrm; ynorm = delta; }
        else { gpnorm = 0.0; ynorm = nrm; }
        KS(ks_waxpby(c, S.Y, 1., S.U, -scale, S.Yt));                 // Y <- X - cnorm * Ytmp
        KS(ks_copy(c, S.Up, S.Y)); KS(ks_residual(c));                // G = F(Y)
        KS(ks_norm(c, S.G, &gnorm));
        double rho;
        if (fnorm == gpnorm) rho = 0.0;
        else rho = (fnorm * fnorm - gnorm * gnorm) / (fnorm * fnorm - gpnorm * gpnorm);
        if (rho < mu) delta *= delta1; else if (rho < eta) delta *= delta2; else delta *= delta3;
        if (rho > sigma) break;
        if (delta < xnorm * deltatol) { reason = -8; breakout = true; break; }         // trust region collapsed
        if (S.evals >= 10000) { reason = -2; breakout = true; break; }
      }
    }
    if (breakout) break;
    fnorm = gnorm;
    KS(ks_copy(c, S.F, S.G)); KS(ks_copy(c, S.U, S.Y));
    newton++;
    if (info->n_history < 17) info->fnorm_history[info->n_history++] = fnorm;
    KS(ks_norm(c, S.U, &xnorm));
    // SNES_TR_Converged_Private / SNESDefaultConverged
    if (sp.trust_region && delta < xnorm * deltatol) reason = 7;
    else if (!(fnorm == fnorm)) reason = -4;
    else if (fnorm < sp.snes_atol) reason = 2;
    else if (fnorm <= ttol) reason = 3;
    else if (ynorm < sp.snes_stol * xnorm) reason = 4;
  }
  if (!reason) reason = -5;            // SNES_DIVERGED_MAX_IT
  // Ucont <- U (VecCopy(U, user->Ucont), implicitsolver.c:4307) with lUcont's ghosts refreshed (:4309-4310)
  { UnpackAoS f = {c->d, S.U, S_UC0, 3}; KS(launch(c, box_owned(c), f)); }
  KS(g2l(c, grp(S_UC0, 3)));
  info->newton_iterations = newton; info->krylov_iterations = lits_total; info->residual_evals = (int)S.evals;
  info->reason = reason; info->fnorm = fnorm; info->xnorm = xnorm; info->delta = delta;
#undef KS
  return finish(vfs_sync(c));
}
#endif
