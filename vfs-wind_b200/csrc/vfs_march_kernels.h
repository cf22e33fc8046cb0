// vfs_march_kernels.h — phase-structured k-marching block programs (the performance path).
//
// A block program owns an (i,j) tile of nodes (overlapped tiling: one thread per node of the tile
// INCLUDING its halo, results are produced for the inner nodes only) and marches along k.  Every
// march step is a fixed sequence of phases separated by block-wide barriers; threads talk to their
// i/j neighbours through shared-memory exchange buffers and keep their own column's state in
// registers (struct State).  Written once, compiled twice:
//   * nvcc: k_block_march<P> below, one CUDA thread per tile node;
//   * g++ -DVFS_EMU (tests/emu only): emu_block_march<P> runs the same phases as host loops over
//     the thread index, with the State structs in an array — so the LOGIC of these kernels is
//     checked against the oracle in a GPU-less container.
#ifndef VFS_MARCH_KERNELS_H
#define VFS_MARCH_KERNELS_H
#include "vfs_common.h"
#include "vfs_les_kernels.h"
#include "vfs_rhs_kernels.h"
#include <vector>

struct MarchGrid { int nbx, nby, kbeg, kend, kchunk; };

// k-chunk so that (tiles x chunks) fills whole waves of `nsm` single-block SMs as evenly as possible
static inline int pick_kchunk(int ntiles, int nk, int min_chunk, int nsm = 148) {
  int best = nk; double best_eff = -1;
  for (int nc = 1; nc <= 32; nc++) {
    const int ch = (nk + nc - 1) / nc;
    if (ch < min_chunk && nc > 1) break;
    const long blocks = (long)ntiles * ((nk + ch - 1) / ch);
    const double eff = (double)blocks / (double)(((blocks + nsm - 1) / nsm) * nsm) * (double)ch / (double)(ch + 2);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = ch; }
  }
  return best < 1 ? 1 : best;
}

// ---- LES pass 2 (les.c:308-669): separable Simpson test filters + Germano contraction ---------------
// The reference filters per-node products (w U_a u_b, w |S|S_ij) with the 27-point (1,4,1)^3 Simpson
// stencil, one 27-term sum per product and cell (rhs2.c:499-523).  The stencil is a tensor product,
// so the same sums are formed here as three 3-point passes: along k from the thread's own column
// (global loads, coalesced along i), along i and along j through two shared-memory exchange
// buffers — 4 shared loads + 2 stores per product and cell instead of 27 loads, which is what
// bounded the 27-term form (shared-memory bandwidth, profiles/r01c).  The summation order differs
// from the reference's, i.e. results agree to rounding (~1e-14 relative), not bitwise.  The weight
// sum and sum_weight (les.c:441-468) depend on the grid and mask only and come from LesGeo.
// The tensor algebra that follows the filters (les2_finish_geo) runs in the last phase; its 23
// per-node operands (centre metrics, aj, grid factors, filtered velocity, nvert) of the tile's plane
// are staged by TMA into shared memory while phases 0 and 1 run: the warp that finishes phase 2 last
// issues the copies for the next plane, so nobody waits to refill the single buffer and the ~30
// scattered global loads that stalled the finish (profiles/r01j) become shared-memory reads.
template <int TY_> struct Les2MarchT {
  static constexpr int TX = 32, TY = TY_, NT = TX * TY, NV = 15;
  // finish operands: NOPI scalars read at the node itself (staged for the tile's inner rows only, NTI nodes) and
  // NOPF scalars whose i/j neighbours are read too (whole tile)
  static constexpr int NOPI = 13, NOPF = 4, NOP = NOPI + NOPF, NTI = TX * (TY - 2);
  static constexpr int MINB = TY_ <= 8 ? 2 : 1;                      // resident blocks per SM the tile is sized for
  // operand buffers: two when they fit, so that plane k+1's operands are requested a whole step before they are read
  static constexpr int OPSZ = NOPI * NTI + NOPF * NT;
  static constexpr int NBUF = (2 * NV * NT + 2 * OPSZ) * 8 + 64 <= 227 * 1024 ? 2 : 1;
  static constexpr int NRAW = 13;
  // RAWS: the 13 per-node inputs of the next plane are TMA-staged one step ahead, so the k pass reads shared memory
  // instead of waiting on a burst of global loads (1.53 -> 1.39 ms once the operands were double-buffered; the same
  // staging had been slower, 2.04 vs 1.92 ms, while both streams shared one barrier-bound step)
  static constexpr int RAWS = (NBUF == 2 && (2 * NV * NT + 2 * OPSZ + NRAW * NT) * 8 + 64 <= 227 * 1024) ? 1 : 0;
  static constexpr int OFF_A = NV * NT, OFF_OP = 2 * NV * NT, OFF_OPF = OFF_OP + NOPI * NTI, OFF_RAW = OFF_OP + NBUF * OPSZ, OFF_BAR = OFF_RAW + RAWS * NRAW * NT;
  static constexpr long SMEM_D = OFF_BAR + 4;
  VFS_HD static int raw_sid(int q) { return q == 0 ? S_LW : (q < 4 ? S_U0 + (q - 1) : (q < 7 ? S_LU0 + (q - 4) : S_LSS0 + (q - 7))); }
  // operand slot -> scalar id: 0..9 csi,eta,zet,aj | 10..12 LFINV,LTF2,LF2 || 13..15 UF | 16 nvert
  VFS_HD static int op_sid(int q) { return q < 10 ? S_CSI0 + q : (q < 13 ? S_LFINV + (q - 10) : (q < 16 ? S_UF0 + (q - 13) : S_NV)); }
  struct State { double v[NV]; double ufk[6]; double nvk[2]; double win[2][NRAW]; };
  VfsDev d;
  static int tiles_x(const VfsDev &d) { return (d.mx - 2 + TX - 3) / (TX - 2); }
  static int tiles_y(const VfsDev &d) { return (d.my - 2 + TY - 3) / (TY - 2); }
  VFS_HD static int iorg(int bx) { return bx * (TX - 2); }
  VFS_HD static int jorg(int by) { return by * (TY - 2); }
  struct Ops {        // finish operands: plane k from the staged buffer, planes k-1/k+1 of UF and nvert from registers
    const double *op, *opf; double ufk[6], nvk[2];      // op: inner-row tiles at the node, opf: whole-tile tiles at the node
    VFS_HD double met(int s) const { return op[s * NTI]; }
    VFS_HD double aj() const { return op[9 * NTI]; }
    VFS_HD double geo(int q) const { return op[(10 + q) * NTI]; }
    VFS_HD double u(int a, int di, int dj, int dk) const { return dk == 0 ? opf[a * NT + dj * TX + di] : (dk < 0 ? ufk[a] : ufk[3 + a]); }
    VFS_HD double nv(int di, int dj, int dk) const { return dk == 0 ? opf[3 * NT + dj * TX + di] : (dk < 0 ? nvk[0] : nvk[1]); }
    VFS_HD unsigned char nearv(const VfsDev &dd, long pp) const { return dd.near[pp]; }
  };
  // phase 0: k pass over the thread's own column (+ the k-neighbours of UF / nvert the finish needs).
  // The per-node inputs of planes k-1 and k are carried in registers from the previous step (st.win),
  // so each step fetches one new plane (13 values) instead of three: this kernel's limiter was the
  // L2 -> SM traffic of re-reading every plane three times (profiles/r01l: 12.4 GB through L2 in 2.3 ms).
  VFS_HD static void load_raw(const VfsDev &d, long n, double *r) {
    r[0] = d.s[S_LW][n];
#pragma unroll
    for (int a = 0; a < 3; a++) { r[1 + a] = d.s[S_U0 + a][n]; r[4 + a] = d.s[S_LU0 + a][n]; }
#pragma unroll
    for (int a = 0; a < 6; a++) r[7 + a] = d.s[S_LSS0 + a][n];
  }
  VFS_HD static void add_plane(double *K, const double *r, bool mid) {
    const double sw = mid ? 4. * r[0] : r[0];
    const double U0 = sw * r[4], U1 = sw * r[5], U2 = sw * r[6];
    K[0] += U0 * r[1]; K[1] += U0 * r[2]; K[2] += U0 * r[3];
    K[3] += U1 * r[1]; K[4] += U1 * r[2]; K[5] += U1 * r[3];
    K[6] += U2 * r[1]; K[7] += U2 * r[2]; K[8] += U2 * r[3];
#pragma unroll
    for (int a = 0; a < 6; a++) K[9 + a] += sw * r[7 + a];
  }
  VFS_HD void phase0(State &st, int tid, int bx, int by, int k, bool first, double *sm) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    double K[NV];
#pragma unroll
    for (int a = 0; a < NV; a++) K[a] = 0;
    if (i <= d.mx - 1 && j <= d.my - 1) {
      const long p = d.idx(i, j, k);
      if (first) { load_raw(d, p - d.sk, st.win[0]); load_raw(d, p, st.win[1]); }
      double nw[NRAW];
      load_raw(d, p + d.sk, nw);
#pragma unroll
      for (int a = 0; a < 3; a++) { st.ufk[a] = d.s[S_UF0 + a][p - d.sk]; st.ufk[3 + a] = d.s[S_UF0 + a][p + d.sk]; }
      st.nvk[0] = d.s[S_NV][p - d.sk]; st.nvk[1] = d.s[S_NV][p + d.sk];
      add_plane(K, st.win[0], false); add_plane(K, st.win[1], true); add_plane(K, nw, false);
#pragma unroll
      for (int a = 0; a < NRAW; a++) { st.win[0][a] = st.win[1][a]; st.win[1][a] = nw[a]; }
    }
#pragma unroll
    for (int a = 0; a < NV; a++) { st.v[a] = K[a]; sm[a * NT + tid] = K[a]; }
  }
  // phase 1: i pass (tile-edge columns produce unused values)
  VFS_HD void phase1(State &st, int tid, double *sm) const {
    const int tx = tid % TX;
    const int l = tx > 0 ? tid - 1 : tid, r = tx < TX - 1 ? tid + 1 : tid;
    double *sA = sm + OFF_A;
#pragma unroll
    for (int a = 0; a < NV; a++) {
      const double A = sm[a * NT + l] + 4. * st.v[a] + sm[a * NT + r];
      st.v[a] = A; sA[a * NT + tid] = A;
    }
  }
#if defined(__CUDACC__) && !defined(VFS_EMU)
  // phases 0 and 1 in one, CUDA only: a tile row is exactly one warp (TX = 32), so the i pass takes its two
  // neighbours by warp shuffle instead of through the first exchange buffer — one block barrier per plane
  // instead of two (the barrier was this kernel's largest stall, profiles/r01q).  The lanes at the tile edge
  // get their own value back from the shuffle, which is what phase1's clamped indices read: same bits.
  // The j-pass buffer alternates between the two halves of the exchange area (sA), because a warp may start the
  // next plane while others still read this one's.
  __device__ __forceinline__ void phase01(State &st, int tid, int bx, int by, int k, bool first, double *sm, double *sA) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    double K[NV];
#pragma unroll
    for (int a = 0; a < NV; a++) K[a] = 0;
    if (i <= d.mx - 1 && j <= d.my - 1) {
      const long p = d.idx(i, j, k);
      if (first) { load_raw(d, p - d.sk, st.win[0]); load_raw(d, p, st.win[1]); }
      double nw[NRAW];
      if (RAWS) {
#pragma unroll
        for (int a = 0; a < NRAW; a++) nw[a] = sm[OFF_RAW + a * NT + tid];
      } else load_raw(d, p + d.sk, nw);
#pragma unroll
      for (int a = 0; a < 3; a++) { st.ufk[a] = d.s[S_UF0 + a][p - d.sk]; st.ufk[3 + a] = d.s[S_UF0 + a][p + d.sk]; }
      st.nvk[0] = d.s[S_NV][p - d.sk]; st.nvk[1] = d.s[S_NV][p + d.sk];
      add_plane(K, st.win[0], false); add_plane(K, st.win[1], true); add_plane(K, nw, false);
#pragma unroll
      for (int a = 0; a < NRAW; a++) { st.win[0][a] = st.win[1][a]; st.win[1][a] = nw[a]; }
    }
#pragma unroll
    for (int a = 0; a < NV; a++) {
      const double l = __shfl_up_sync(0xffffffffu, K[a], 1), r = __shfl_down_sync(0xffffffffu, K[a], 1);
      const double A = l + 4. * K[a] + r;
      st.v[a] = A; sA[a * NT + tid] = A;
    }
  }
#endif
  // phase 2: j pass + les.c:470-669 for the inner nodes of the tile
  // sA: the exchange buffer holding the i-pass results of this plane (OFF_A, or the alternating buffer of phase01)
  // sm: here the base of the operand buffer's frame, i.e. shared-memory base + (buffer index) * OPSZ
  VFS_HD void phase2(const State &st, int tid, int bx, int by, int k, const double *sm, const double *sA) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    if (tx < 1 || tx > TX - 2 || ty < 1 || ty > TY - 2 || i > d.mx - 2 || j > d.my - 2) return;
    const long p = d.idx(i, j, k);
    Ops O; O.op = sm + OFF_OP + (tid - TX); O.opf = sm + OFF_OPF + tid;
    if (O.nv(0, 0, 0) > 1.1) { d.s[S_LM][p] = 0; d.s[S_MM][p] = 0; return; }
#pragma unroll
    for (int a = 0; a < 6; a++) O.ufk[a] = st.ufk[a];
    O.nvk[0] = st.nvk[0]; O.nvk[1] = st.nvk[1];
    const int up = tid - TX, dn = tid + TX;
    double f[NV];
#pragma unroll
    for (int a = 0; a < NV; a++) f[a] = sA[a * NT + up] + 4. * st.v[a] + sA[a * NT + dn];
    les2_finish_geo(d, O, i, j, k + d.kofs, p, f);
  }
};
typedef Les2MarchT<16> Les2March;      // one 512-thread block per SM
typedef Les2MarchT<12> Les2March12;    // 384 threads: 170 registers per thread hold the k window without spills
typedef Les2MarchT<8> Les2March8;      // two 256-thread blocks per SM: the phases of one block overlap the other's

#ifndef VFS_EMU
#include "vfs_fused_kernels.h"
// tmap: box (TX, TY) = the whole tile; tmapi: box (TX, TY-2) = its inner rows
template <class M> __global__ void __launch_bounds__(M::NT, M::MINB) k_les2_march(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapi, const M P, int kbeg, int kend, int kchunk) {
  extern __shared__ __align__(128) double vfs_les2_sm[];
  double *sm = vfs_les2_sm;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + M::OFF_BAR);      // bar[0], bar[1]: one per operand buffer
  unsigned *cnt = reinterpret_cast<unsigned *>(bar + 2);
  const int tid = threadIdx.x, bx = blockIdx.x, by = blockIdx.y;
  const int ka = kbeg + blockIdx.z * kchunk, kb = min(kend, ka + kchunk);
  if (ka >= kb) return;
  auto issue = [&](int k) {       // one thread: the operand tiles of plane k -> shared memory (buffer (k-ka) % NBUF)
    const int b = (k - ka) % M::NBUF;
    double *dst = sm + b * M::OPSZ;
    mbar_expect_tx(&bar[b], M::OPSZ * 8);
#pragma unroll 1
    for (int q = 0; q < M::NOPI; q++) tma_load_tile(dst + M::OFF_OP + q * M::NTI, &tmapi, M::iorg(bx) + VFS_G, M::jorg(by) + 1 + VFS_G, k + VFS_G, M::op_sid(q), &bar[b]);
#pragma unroll 1
    for (int q = 0; q < M::NOPF; q++) tma_load_tile(dst + M::OFF_OPF + q * M::NT, &tmap, M::iorg(bx) + VFS_G, M::jorg(by) + VFS_G, k + VFS_G, M::op_sid(M::NOPI + q), &bar[b]);
  };
  unsigned long long *bar_raw = bar + 3;
  auto issue_raw = [&](int k) {
    mbar_expect_tx(bar_raw, M::NRAW * M::NT * 8);
#pragma unroll 1
    for (int q = 0; q < M::NRAW; q++) tma_load_tile(sm + M::OFF_RAW + q * M::NT, &tmap, M::iorg(bx) + VFS_G, M::jorg(by) + VFS_G, k + VFS_G, M::raw_sid(q), bar_raw);
  };
  if (tid == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(bar_raw, 1); *cnt = 0;
    fence_mbar_init();
    issue(ka);
    if (M::RAWS) issue_raw(ka + 1);
  }
  __syncthreads();
  typename M::State st;
  for (int k = ka; k < kb; k++) {
    double *sA = sm + ((k - ka) & 1) * M::OFF_A;       // the two NV*NT halves of the exchange area alternate
    if (M::RAWS) mbar_wait(bar_raw, (k - ka) & 1);
    P.phase01(st, tid, bx, by, k, k == ka, sm, sA);
    __syncthreads();
    if (M::RAWS && tid == 0 && k + 1 < kb) { fence_proxy_async(); issue_raw(k + 2); }
    const int b = (k - ka) % M::NBUF;
    if constexpr (M::NBUF == 2) {
      // everyone is past plane k-1's finish, whose operand buffer is the one plane k+1 goes into: request it now,
      // a whole step before it is read (profiles/r01s: 27 % of this kernel's stall samples sat on the operand wait)
      if (tid == 0 && k + 1 < kb) { fence_proxy_async(); issue(k + 1); }
      mbar_wait(&bar[b], ((k - ka) >> 1) & 1);
      P.phase2(st, tid, bx, by, k, sm + b * M::OPSZ, sA);
    } else {
      mbar_wait(&bar[0], (k - ka) & 1);
      P.phase2(st, tid, bx, by, k, sm, sA);
      // the warp that leaves phase 2 last refills the operand buffer for the next plane
      __syncwarp();
      if ((tid & 31) == 0) {
        __threadfence_block();
        if (atomicAdd(cnt, 1u) == M::NT / 32 - 1) {
          *reinterpret_cast<volatile unsigned *>(cnt) = 0;
          __threadfence_block();
          if (k + 1 < kb) { fence_proxy_async(); issue(k + 1); }
        }
      }
    }
  }
}
template <class M> static inline int run_les2_march(cudaStream_t stream, const CUtensorMap &tmap, const CUtensorMap &tmapi, const M &P, int k0, int k1, long *launches) {
  if (k1 <= k0) return 0;
  static bool attr_set = false;
  const int bytes = (int)(M::SMEM_D * sizeof(double));
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_les2_march<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return -2;
    attr_set = true;
  }
  const int ntx = M::tiles_x(P.d), nty = M::tiles_y(P.d);
  const int kchunk = pick_kchunk(ntx * nty, k1 - k0, 16, 148 * M::MINB);
  dim3 grd(ntx, nty, (k1 - k0 + kchunk - 1) / kchunk);
  k_les2_march<M><<<grd, M::NT, bytes, stream>>>(tmap, tmapi, P, k0, k1, kchunk);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#else
template <class M> static inline int run_les2_march(void *, const M &P, int k0, int k1, long *launches) {
  if (k1 <= k0) return 0;
  std::vector<double> smv(M::SMEM_D);
  std::vector<typename M::State> st(M::NT);
  double *sm = smv.data();
  const VfsDev &d = P.d;
  const int ntx = M::tiles_x(d), nty = M::tiles_y(d);
  const int kchunk = pick_kchunk(ntx * nty, k1 - k0, 4, 7);
  for (int bz = 0; bz * kchunk < k1 - k0; bz++)
    for (int by = 0; by < nty; by++)
      for (int bx = 0; bx < ntx; bx++) {
        const int ka = k0 + bz * kchunk, kb = k1 < ka + kchunk ? k1 : ka + kchunk;
        auto issue = [&](int k) {       // what the TMA unit does: box copy with zero fill outside the padded array
          for (int q = 0; q < M::NOP; q++) {
            const bool inner = q < M::NOPI;       // inner-row box (rows 1 .. TY-2) or the whole tile
            double *dst = inner ? sm + M::OFF_OP + q * M::NTI : sm + M::OFF_OPF + (q - M::NOPI) * M::NT;
            for (int y = 0; y < (inner ? M::TY - 2 : M::TY); y++)
              for (int x = 0; x < M::TX; x++) {
                const int X = M::iorg(bx) + VFS_G + x, Y = M::jorg(by) + VFS_G + y + (inner ? 1 : 0), Z = k + VFS_G;
                const bool in = X >= 0 && X < d.pitch && Y >= 0 && Y < d.ny && Z >= 0 && Z < d.nzt;
                dst[y * M::TX + x] = in ? d.s[M::op_sid(q)][(long)Z * d.sk + (long)Y * d.sj + X] : 0.;
              }
          }
        };
        issue(ka);
        for (int k = ka; k < kb; k++) {
          for (int t = 0; t < M::NT; t++) P.phase0(st[t], t, bx, by, k, k == ka, sm);
          for (int t = 0; t < M::NT; t++) P.phase1(st[t], t, sm);
          // phase 1 reads its neighbours' phase-0 values from the exchange buffer while updating st.v in place
          for (int t = 0; t < M::NT; t++) P.phase2(st[t], t, bx, by, k, sm, sm + M::OFF_A);
          if (k + 1 < kb) issue(k + 1);
        }
      }
  (*launches)++;
  return 0;
}
#endif

#ifndef VFS_EMU
// face fluxes with TMA-staged ucat AND metric planes (k_flux_march, vfs_fused_kernels.h)
static inline int launch_flux_march(cudaStream_t st, const CUtensorMap &tmapA, const CUtensorMap &tmapB, const VfsDev &d, int k0, int k1, long *launches) {
  if (k1 <= k0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_flux_march, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FluxMarch::BYTES) != cudaSuccess) return -2;
    attr_set = true;
  }
  const int ntx = (d.mx - 2 + FluxMarch::TX - 1) / FluxMarch::TX, nty = (d.my - 2 + FluxMarch::TY - 1) / FluxMarch::TY;
  const int kchunk = pick_kchunk(ntx * nty, k1 - k0, 16);
  dim3 grd(ntx, nty, (k1 - k0 + kchunk - 1) / kchunk), blk(FluxMarch::TX, FluxMarch::TY, 1);
  k_flux_march<<<grd, blk, FluxMarch::BYTES, st>>>(tmapA, tmapB, d, k0, k1, kchunk);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#endif

// ---- LES pass 1 (les.c:199-246): grad u, |S|, test-filtered velocity + the per-node products of pass 2 ----
// Same block-program shape: the test filter of u (weights w = 1/aj, 0 where nvert > 0.1) is four separable
// (1,4,1)^3 sums (w, w u_a); the centre-difference stencil of grad u takes its i/j neighbours from a third
// exchange buffer holding the plane's u and nvert and its k neighbours from the thread's own column.
template <int TY_> struct Les1MarchT {
  static constexpr int TX = 32, TY = TY_, NT = TX * TY, NV = 4;
  static constexpr int OFF_A = NV * NT, OFF_U = 2 * NV * NT;
  static constexpr long SMEM_D = 3L * NV * NT;
  static constexpr bool HAS01 = false;
  struct State { double v[NV]; double uk[6], nvk[2], iaj0; };
  VfsDev d;
  static int tiles_x(const VfsDev &d) { return (d.mx - 2 + TX - 3) / (TX - 2); }
  static int tiles_y(const VfsDev &d) { return (d.my - 2 + TY - 3) / (TY - 2); }
  VFS_HD static int iorg(int bx) { return bx * (TX - 2); }
  VFS_HD static int jorg(int by) { return by * (TY - 2); }
  struct Acc {        // u / nvert of plane k from the exchange buffer, planes k-1/k+1 from registers; own metrics from global
    const double *su; double uk[6], nvk[2]; const VfsDev &d; long p;
    VFS_HD double u(int a, int di, int dj, int dk) const { return dk == 0 ? su[a * NT + dj * TX + di] : (dk < 0 ? uk[a] : uk[3 + a]); }
    VFS_HD double nv(int di, int dj, int dk) const { return dk == 0 ? su[3 * NT + dj * TX + di] : (dk < 0 ? nvk[0] : nvk[1]); }
    VFS_HD double met(int s) const { return d.s[S_CSI0 + s][p]; }
    VFS_HD double aj() const { return d.s[S_AJ][p]; }
    VFS_HD unsigned char nearv(const VfsDev &dd, long pp) const { return dd.near[pp]; }
  };
  VFS_HD void phase0(State &st, int tid, int bx, int by, int k, bool, double *sm) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    double K[NV] = {0, 0, 0, 0}, u0[3] = {0, 0, 0}, nv0 = 0;
    st.iaj0 = 0;
    if (i <= d.mx - 1 && j <= d.my - 1) {
      const long p = d.idx(i, j, k);
#pragma unroll
      for (int dk = -1; dk <= 1; dk++) {
        const long n = p + dk * d.sk;
        const double nv = d.s[S_NV][n], ia = d.s[S_IAJ][n];
        const double u[3] = {d.s[S_U0][n], d.s[S_U1][n], d.s[S_U2][n]};
        const double w = nv > 0.1 ? 0. : ia;
        const double sw = dk == 0 ? 4. * w : w;
        K[0] += sw; K[1] += sw * u[0]; K[2] += sw * u[1]; K[3] += sw * u[2];
        if (dk == 0) { u0[0] = u[0]; u0[1] = u[1]; u0[2] = u[2]; nv0 = nv; st.iaj0 = ia; }
        else { const int o = dk < 0 ? 0 : 3; st.uk[o] = u[0]; st.uk[o + 1] = u[1]; st.uk[o + 2] = u[2]; st.nvk[dk < 0 ? 0 : 1] = nv; }
      }
    }
    double *sU = sm + OFF_U;
#pragma unroll
    for (int a = 0; a < NV; a++) { st.v[a] = K[a]; sm[a * NT + tid] = K[a]; }
    sU[tid] = u0[0]; sU[NT + tid] = u0[1]; sU[2 * NT + tid] = u0[2]; sU[3 * NT + tid] = nv0;
  }
  VFS_HD void phase1(State &st, int tid, double *sm) const {
    const int tx = tid % TX;
    const int l = tx > 0 ? tid - 1 : tid, r = tx < TX - 1 ? tid + 1 : tid;
    double *sA = sm + OFF_A;
#pragma unroll
    for (int a = 0; a < NV; a++) {
      const double A = sm[a * NT + l] + 4. * st.v[a] + sm[a * NT + r];
      st.v[a] = A; sA[a * NT + tid] = A;
    }
  }
  VFS_HD void phase2(const State &st, int tid, int bx, int by, int k, const double *sm) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    if (tx < 1 || tx > TX - 2 || ty < 1 || ty > TY - 2 || i > d.mx - 2 || j > d.my - 2) return;
    const long p = d.idx(i, j, k);
    Acc A = {sm + OFF_U + tid, {st.uk[0], st.uk[1], st.uk[2], st.uk[3], st.uk[4], st.uk[5]}, {st.nvk[0], st.nvk[1]}, d, p};
    const double nv0 = A.nv(0, 0, 0);
    const double u0 = A.u(0, 0, 0, 0), u1 = A.u(1, 0, 0, 0), u2 = A.u(2, 0, 0, 0);
    double g[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, S = 0, uf[3] = {0, 0, 0};
    if (!(nv0 > 1.1)) {      // skipped cells keep the zeros of the reference's freshly created work vectors
      grad_center_auto(d, A, i, j, d.kglob(k), p, g);
      S = sabs_of(g);
      const double *sA = sm + OFF_A;
      const int up = tid - TX, dn = tid + TX;
      const double ws = sA[up] + 4. * st.v[0] + sA[dn];
#pragma unroll
      for (int a = 0; a < 3; a++) uf[a] = (sA[(1 + a) * NT + up] + 4. * st.v[1 + a] + sA[(1 + a) * NT + dn]) / ws;
    }
    d.s[S_SABS][p] = S;
#pragma unroll
    for (int a = 0; a < 3; a++) d.s[S_UF0 + a][p] = uf[a];
    // per-node quantities pass 2 filters (les_derive_store)
    d.s[S_LW][p] = nv0 > 0.1 ? 0. : st.iaj0;
    d.s[S_LU0][p] = u0 * A.met(0) + u1 * A.met(1) + u2 * A.met(2);
    d.s[S_LU1][p] = u0 * A.met(3) + u1 * A.met(4) + u2 * A.met(5);
    d.s[S_LU2][p] = u0 * A.met(6) + u1 * A.met(7) + u2 * A.met(8);
    d.s[S_LSS0][p] = (0.5 * (g[0][0] + g[0][0])) * S; d.s[S_LSS1][p] = (0.5 * (g[0][1] + g[1][0])) * S; d.s[S_LSS2][p] = (0.5 * (g[0][2] + g[2][0])) * S;
    d.s[S_LSS3][p] = (0.5 * (g[1][1] + g[1][1])) * S; d.s[S_LSS4][p] = (0.5 * (g[1][2] + g[2][1])) * S; d.s[S_LSS5][p] = (0.5 * (g[2][2] + g[2][2])) * S;
  }
};
typedef Les1MarchT<16> Les1March;
typedef Les1MarchT<8> Les1March8;
#ifndef VFS_EMU
template <class M, int MINB> __global__ void __launch_bounds__(M::NT, MINB) k_filter_march(const M P, int kbeg, int kend, int kchunk) {
  extern __shared__ __align__(16) double vfs_fm_sm[];
  double *sm = vfs_fm_sm;
  const int tid = threadIdx.x, bx = blockIdx.x, by = blockIdx.y;
  const int ka = kbeg + blockIdx.z * kchunk, kb = min(kend, ka + kchunk);
  typename M::State st;
  for (int k = ka; k < kb; k++) {
    if constexpr (M::HAS01) {                   // one barrier per plane, alternating j-pass buffer
      double *sA = sm + ((k - ka) & 1) * M::NV * M::NT;
      P.phase01(st, tid, bx, by, k, k == ka, sm, sA);
      __syncthreads();
      P.phase2(st, tid, bx, by, k, sm, sA);
    } else {
      P.phase0(st, tid, bx, by, k, k == ka, sm);
      __syncthreads();
      P.phase1(st, tid, sm);
      __syncthreads();
      P.phase2(st, tid, bx, by, k, sm);
      if (M::SMEM_D > 2L * M::NV * M::NT) __syncthreads();      // a third buffer written in phase 0 is still read in phase 2
    }
  }
}
template <class M, int MINB> static inline int run_filter_march(cudaStream_t stream, const M &P, int k0, int k1, long *launches) {
  if (k1 <= k0) return 0;
  const int bytes = (int)(M::SMEM_D * sizeof(double));
  static bool attr_set = false;
  if (!attr_set && bytes > 48 * 1024) {
    if (cudaFuncSetAttribute(k_filter_march<M, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return -2;
    attr_set = true;
  }
  const int ntx = M::tiles_x(P.d), nty = M::tiles_y(P.d);
  const int kchunk = pick_kchunk(ntx * nty, k1 - k0, 16, 148 * MINB);
  dim3 grd(ntx, nty, (k1 - k0 + kchunk - 1) / kchunk);
  k_filter_march<M, MINB><<<grd, M::NT, bytes, stream>>>(P, k0, k1, kchunk);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#else
template <class M, int MINB> static inline int run_filter_march(void *, const M &P, int k0, int k1, long *launches) {
  if (k1 <= k0) return 0;
  std::vector<double> smv(M::SMEM_D);
  std::vector<typename M::State> st(M::NT);
  const int ntx = M::tiles_x(P.d), nty = M::tiles_y(P.d);
  const int kchunk = pick_kchunk(ntx * nty, k1 - k0, 4, 7);
  for (int bz = 0; bz * kchunk < k1 - k0; bz++)
    for (int by = 0; by < nty; by++)
      for (int bx = 0; bx < ntx; bx++) {
        const int ka = k0 + bz * kchunk, kb = k1 < ka + kchunk ? k1 : ka + kchunk;
        for (int k = ka; k < kb; k++) {
          for (int t = 0; t < M::NT; t++) P.phase0(st[t], t, bx, by, k, k == ka, smv.data());
          for (int t = 0; t < M::NT; t++) P.phase1(st[t], t, smv.data());
          for (int t = 0; t < M::NT; t++) P.phase2(st[t], t, bx, by, k, smv.data());
        }
      }
  (*launches)++;
  return 0;
}
#endif

// ---- LES pass 3 (les.c:716-796, 967-980): Simpson filter of LM, MM -> Cs, separable ---------------------
// The weight of a neighbour node in this filter depends on that node alone (1/aj, zero where nvert > 1.1 or
// on non-periodic domain ghosts, with the J == 0-only quirk of les.c:756), so sum(s w LM), sum(s w MM) and
// sum(s w) are three separable (1,4,1)^3 filters of per-node products: k pass from the thread's own
// column, i and j passes through two small shared-memory exchange buffers (the 27-term gather it replaces
// was bound by shared-memory bandwidth, profiles/r01m).  Cells next to a periodic plane (ghost-image
// fetches, les.c:738-768) are skipped here and done by the staged kernel on thin slabs.
template <int TY_> struct Les3MarchT {
  static constexpr int TX = 32, TY = TY_, NT = TX * TY, NV = 3;
  static constexpr long SMEM_D = 2L * NV * NT;
  static constexpr bool HAS01 = true;       // phases 0 + 1 fused with a warp-shuffle i pass on the device (see Les2MarchT::phase01)
  // (a register window over k, as in Les2MarchT, was measured slower here: 0.50 vs 0.44 ms at 256^3 — the three
  // planes' twelve loads are independent L2 hits, the window adds a loop-carried chain; writing nu_t from this
  // kernel as well cost what the separate NuT kernel costs, 0.14 ms, at 2 resident blocks per SM — re-measured at 4, see with_nut)
  struct State { double v[NV]; double nvc; };
  VfsDev d;
  int with_nut;      // also nu_t = Cs Delta^2 |S| of the cell (NuT<true>'s arithmetic: |S| from pass 1, Delta^2 = S_LF2 of LesGeo), vfs_rhs_les_fused only
  static int tiles_x(const VfsDev &d) { return (d.mx - 2 + TX - 3) / (TX - 2); }
  static int tiles_y(const VfsDev &d) { return (d.my - 2 + TY - 3) / (TY - 2); }
  VFS_HD static int iorg(int bx) { return bx * (TX - 2); }
  VFS_HD static int jorg(int by) { return by * (TY - 2); }
  VFS_HD void phase0(State &st, int tid, int bx, int by, int k, bool, double *sm) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    double K[NV] = {0, 0, 0};
    st.nvc = 0;
    if (i <= d.mx - 1 && j <= d.my - 1) {
      const long p = d.idx(i, j, k);
      const bool zij = (!d.perx && (i == 0 || i == d.mx - 1)) || (!d.pery && j == 0);
#pragma unroll
      for (int dk = -1; dk <= 1; dk++) {
        const long n = p + dk * d.sk;
        const int K_ = k + dk + d.kofs;
        const double nv = d.s[S_NV][n];
        double w = d.s[S_IAJ][n];
        if (nv > 1.1 || zij || (!d.perz && (K_ == 0 || K_ == d.mz - 1))) w = 0;
        const double sw = dk == 0 ? 4. * w : w;
        K[0] += sw; K[1] += sw * d.s[S_LM][n]; K[2] += sw * d.s[S_MM][n];
        if (dk == 0) st.nvc = nv;
      }
    }
#pragma unroll
    for (int a = 0; a < NV; a++) { st.v[a] = K[a]; sm[a * NT + tid] = K[a]; }
  }
  VFS_HD void phase1(State &st, int tid, double *sm) const {
    const int tx = tid % TX;
    const int l = tx > 0 ? tid - 1 : tid, r = tx < TX - 1 ? tid + 1 : tid;
    double *sA = sm + NV * NT;
#pragma unroll
    for (int a = 0; a < NV; a++) {
      const double A = sm[a * NT + l] + 4. * st.v[a] + sm[a * NT + r];
      st.v[a] = A; sA[a * NT + tid] = A;
    }
  }
#if defined(__CUDACC__) && !defined(VFS_EMU)
  __device__ __forceinline__ void phase01(State &st, int tid, int bx, int by, int k, bool first, double *sm, double *sA) const {
    phase0(st, tid, bx, by, k, first, sA);          // K in st.v (its copy in sA is overwritten below: only this thread wrote the slot)
#pragma unroll
    for (int a = 0; a < NV; a++) {
      const double l = __shfl_up_sync(0xffffffffu, st.v[a], 1), r = __shfl_down_sync(0xffffffffu, st.v[a], 1);
      const double A = l + 4. * st.v[a] + r;
      st.v[a] = A; sA[a * NT + tid] = A;
    }
  }
#endif
  VFS_HD void phase2(const State &st, int tid, int bx, int by, int k, const double *sm) const { phase2(st, tid, bx, by, k, sm, sm + NV * NT); }
  VFS_HD void phase2(const State &st, int tid, int bx, int by, int k, const double *, const double *sA) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    if (tx < 1 || tx > TX - 2 || ty < 1 || ty > TY - 2 || i > d.mx - 2 || j > d.my - 2) return;
    const int kg = k + d.kofs;
    if ((d.perx && (i == 1 || i == d.mx - 2)) || (d.pery && (j == 1 || j == d.my - 2)) || (d.perz && (kg == 1 || kg == d.mz - 2))) return;
    const long p = d.idx(i, j, k);
    if (st.nvc > 1.1) { d.s[S_CS][p] = 0; if (with_nut) d.s[S_NUT][p] = 0; return; }
    const int up = tid - TX, dn = tid + TX;
    const double ws = sA[up] + 4. * st.v[0] + sA[dn];
    const double lm = sA[NT + up] + 4. * st.v[1] + sA[NT + dn];
    const double mmv = sA[2 * NT + up] + 4. * st.v[2] + sA[2 * NT + dn];
    const double C = 0.5 * (lm / ws) / (mmv / ws + 1.e-4);
    double cs = (C < 0) ? 0 : C;                                           // comparisons as in les3_core (NaN-faithful)
    if (st.nvc > 0.1 && st.nvc < 1.1) cs = (0.001 < cs) ? cs : 0.001;      // clip chain, les.c:967-980
    cs = (cs < 0) ? 0 : cs;
    cs = (cs < d.max_cs) ? cs : d.max_cs;
    d.s[S_CS][p] = cs;
    if (with_nut) d.s[S_NUT][p] = nut_value_geo(d, p, cs, d.s[S_SABS][p]);
  }
};
typedef Les3MarchT<16> Les3March;      // (32 x 32 tiles, 1024-thread blocks, 88 % instead of 82 % interior cells: measured slower, 0.43 vs 0.40 ms — the block barrier)
// ---- Fp folded into the projection (momentum.c:1548-1678 + 1687-1735 + 1833-1841 [+ 2297-2331]) ---------
// The staged chain wrote Fp (FpCell), refreshed its ghosts, applied the periodic node copies and read it back four
// times in the projection (at p, p+1, p+sj, p+sk).  Here a block marches an (i,j) tile along k: every step it
// evaluates Fp of plane k+1 once per tile node (+ one halo column and row) from the 18 face-flux scalars into one
// of three rotating shared-memory planes, then projects plane k from the Fp planes k and k+1 — Fp never reaches
// HBM, the Fp ghost refresh / node copies / inter-rank exchange disappear, and the bytes are the 18 flux planes
// read once.  The periodic copies of Fp (node m-1 takes the value of node 1, momentum.c:1687-1713) become an index
// remap of the cell whose Fp is evaluated; non-periodic boundary nodes' Fp is never used (their components are
// masked, momentum.c:1833-1841).  Arithmetic and operand order are FpCell's and project_fp's: bitwise the staged
// result.  Boundary NODES of the slab (mask 7: no projection) are assembled by the staged functor on the shell.
// Fp at node (i, j, k) as the projection reads it: the periodic node copies of Fp (momentum.c:1687-1713: node m-1
// takes the value of node 1) folded into an index remap; zero where the value is never used (non-periodic boundary
// nodes: the component that would read it is masked, momentum.c:1833-1841)
VFS_HD void fp_as_projected(const VfsDev &d, int i, int j, int k, double f[3]) {
  f[0] = f[1] = f[2] = 0;
  bool ok = i <= d.mx - 1 && j <= d.my - 1;
  int kg = k + d.kofs;
  if (ok && i == d.mx - 1) { if (d.perx) i = 1; else ok = false; }
  if (ok && j == d.my - 1) { if (d.pery) j = 1; else ok = false; }
  if (ok && kg == d.mz - 1) {
    // the image of global plane 1: the plane itself on a single rank, the ghost plane two above on the last rank
    if (d.perz) { k = d.single_rank ? 1 : k + 2; kg = 1; } else ok = false;
  }
  if (ok) fp_cell_value(d, i, j, kg, d.idx(i, j, k), f);
}
// One-thread-per-node form of the same fusion: Fp evaluated at the node and its +i, +j, +k neighbours (the re-reads
// hit L1 / L2), no shared memory, no barrier.
struct ProjectFpBox {
  VfsDev d; int mode, s0; double scale;
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    const int m = rhs_mask(d, i, j, kg, p);
    double f[3], fi[3], fj[3], fk[3];
    fp_as_projected(d, i, j, k, f); fp_as_projected(d, i + 1, j, k, fi); fp_as_projected(d, i, j + 1, k, fj); fp_as_projected(d, i, j, k + 1, fk);
    const double ia = d.s[S_IAJ][p];
    double rr[3];
    { const long q = p + 1; const double a = 2. / (ia + d.s[S_IAJ][q]);
      rr[0] = (0.5 * (d.s[S_CSI0][p] * f[0] + d.s[S_CSI1][p] * f[1] + d.s[S_CSI2][p] * f[2]) + 0.5 * (d.s[S_CSI0][q] * fi[0] + d.s[S_CSI1][q] * fi[1] + d.s[S_CSI2][q] * fi[2])) * a; }
    { const long q = p + d.sj; const double a = 2. / (ia + d.s[S_IAJ][q]);
      rr[1] = (0.5 * (d.s[S_ETA0][p] * f[0] + d.s[S_ETA1][p] * f[1] + d.s[S_ETA2][p] * f[2]) + 0.5 * (d.s[S_ETA0][q] * fj[0] + d.s[S_ETA1][q] * fj[1] + d.s[S_ETA2][q] * fj[2])) * a; }
    { const long q = p + d.sk; const double a = 2. / (ia + d.s[S_IAJ][q]);
      rr[2] = (0.5 * (d.s[S_ZET0][p] * f[0] + d.s[S_ZET1][p] * f[1] + d.s[S_ZET2][p] * f[2]) + 0.5 * (d.s[S_ZET0][q] * fk[0] + d.s[S_ZET1][q] * fk[1] + d.s[S_ZET2][q] * fk[2])) * a; }
    if (mode == 0) { for (int a = 0; a < 3; a++) d.s[s0 + a][p] = ((m >> a) & 1) ? 0. : d.s[s0 + a][p] + scale * rr[a]; }
    else { for (int a = 0; a < 3; a++) d.s[S_R0 + a][p] = snes_assemble(d, a, p, (m >> a) & 1, rr[a]); }
  }
};

struct ProjFpMarch {
  static constexpr int TX = 32, TY = 8, NT = TX * TY, FX = TX + 1, FY = TY + 1, FN = FX * FY, NBUF = 3;
  static constexpr long SMEM_D = (long)NBUF * 3 * FN;
  VfsDev d; int mode, s0; double scale;
  static int tiles_x(const VfsDev &d) { return (d.mx - 2 + TX - 1) / TX; }
  static int tiles_y(const VfsDev &d) { return (d.my - 2 + TY - 1) / TY; }
  VFS_HD void fp_slot(double *buf, int slot, int i, int j, int k) const {
    double f[3];
    fp_as_projected(d, i, j, k, f);
    buf[slot] = f[0]; buf[FN + slot] = f[1]; buf[2 * FN + slot] = f[2];
  }
  // phase A: Fp of plane kq -> buffer kq mod 3 (tile nodes + the halo column i0+TX and row j0+TY)
  VFS_HD void phaseA(int tid, int bx, int by, int kq, int kref, double *sm) const {
    const int tx = tid % TX, ty = tid / TX, i0 = 1 + bx * TX, j0 = 1 + by * TY;
    double *buf = sm + ((kq - kref) % NBUF) * 3 * FN;
    fp_slot(buf, ty * FX + tx, i0 + tx, j0 + ty, kq);
    if (tid < TY) fp_slot(buf, tid * FX + TX, i0 + TX, j0 + tid, kq);
    else if (tid < TY + TX) fp_slot(buf, TY * FX + (tid - TY), i0 + (tid - TY), j0 + TY, kq);
  }
  // phase B: projection, masks and assembly of plane k
  VFS_HD void phaseB(int tid, int bx, int by, int k, int kref, const double *sm) const {
    const int tx = tid % TX, ty = tid / TX, i = 1 + bx * TX + tx, j = 1 + by * TY + ty;
    if (i > d.mx - 2 || j > d.my - 2) return;
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    const double *b0 = sm + ((k - kref) % NBUF) * 3 * FN + ty * FX + tx, *b1 = sm + ((k + 1 - kref) % NBUF) * 3 * FN + ty * FX + tx;
    const int m = rhs_mask(d, i, j, kg, p);
    const double f0 = b0[0], f1 = b0[FN], f2 = b0[2 * FN];
    const double ia = d.s[S_IAJ][p];
    double rr[3];
    {
      const long q = p + 1;
      const double iaj = 2. / (ia + d.s[S_IAJ][q]);
      rr[0] = (0.5 * (d.s[S_CSI0][p] * f0 + d.s[S_CSI1][p] * f1 + d.s[S_CSI2][p] * f2) +
               0.5 * (d.s[S_CSI0][q] * b0[1] + d.s[S_CSI1][q] * b0[FN + 1] + d.s[S_CSI2][q] * b0[2 * FN + 1])) * iaj;
    }
    {
      const long q = p + d.sj;
      const double jaj = 2. / (ia + d.s[S_IAJ][q]);
      rr[1] = (0.5 * (d.s[S_ETA0][p] * f0 + d.s[S_ETA1][p] * f1 + d.s[S_ETA2][p] * f2) +
               0.5 * (d.s[S_ETA0][q] * b0[FX] + d.s[S_ETA1][q] * b0[FN + FX] + d.s[S_ETA2][q] * b0[2 * FN + FX])) * jaj;
    }
    {
      const long q = p + d.sk;
      const double kaj = 2. / (ia + d.s[S_IAJ][q]);
      rr[2] = (0.5 * (d.s[S_ZET0][p] * f0 + d.s[S_ZET1][p] * f1 + d.s[S_ZET2][p] * f2) +
               0.5 * (d.s[S_ZET0][q] * b1[0] + d.s[S_ZET1][q] * b1[FN] + d.s[S_ZET2][q] * b1[2 * FN])) * kaj;
    }
    if (mode == 0) {
#pragma unroll
      for (int a = 0; a < 3; a++) d.s[s0 + a][p] = ((m >> a) & 1) ? 0. : d.s[s0 + a][p] + scale * rr[a];
    } else {
#pragma unroll
      for (int a = 0; a < 3; a++) d.s[S_R0 + a][p] = snes_assemble(d, a, p, (m >> a) & 1, rr[a]);
    }
  }
};
#ifndef VFS_EMU
__global__ void __launch_bounds__(ProjFpMarch::NT, 4) k_projfp_march(const ProjFpMarch P, int kbeg, int kend, int kchunk) {
  __shared__ double sm[ProjFpMarch::SMEM_D];
  const int tid = threadIdx.x, bx = blockIdx.x, by = blockIdx.y;
  const int ka = kbeg + blockIdx.z * kchunk, kb = min(kend, ka + kchunk);
  if (ka >= kb) return;
  P.phaseA(tid, bx, by, ka, ka, sm);
  for (int k = ka; k < kb; k++) {
    P.phaseA(tid, bx, by, k + 1, ka, sm);
    __syncthreads();
    P.phaseB(tid, bx, by, k, ka, sm);
  }
}
static inline int run_projfp_march(cudaStream_t stream, const ProjFpMarch &P, int k0, int k1, long *launches) {
  if (k1 <= k0) return 0;
  const int ntx = ProjFpMarch::tiles_x(P.d), nty = ProjFpMarch::tiles_y(P.d);
  const int kchunk = pick_kchunk(ntx * nty, k1 - k0, 16, 148 * 4);
  dim3 grd(ntx, nty, (k1 - k0 + kchunk - 1) / kchunk);
  k_projfp_march<<<grd, ProjFpMarch::NT, 0, stream>>>(P, k0, k1, kchunk);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#else
static inline int run_projfp_march(void *, const ProjFpMarch &P, int k0, int k1, long *launches) {
  if (k1 <= k0) return 0;
  typedef ProjFpMarch M;
  std::vector<double> smv(M::SMEM_D);
  double *sm = smv.data();
  const int ntx = M::tiles_x(P.d), nty = M::tiles_y(P.d);
  const int kchunk = pick_kchunk(ntx * nty, k1 - k0, 4, 7);
  for (int bz = 0; bz * kchunk < k1 - k0; bz++)
    for (int by = 0; by < nty; by++)
      for (int bx = 0; bx < ntx; bx++) {
        const int ka = k0 + bz * kchunk, kb = k1 < ka + kchunk ? k1 : ka + kchunk;
        for (int t = 0; t < M::NT; t++) P.phaseA(t, bx, by, ka, ka, sm);
        for (int k = ka; k < kb; k++) {
          for (int t = 0; t < M::NT; t++) P.phaseA(t, bx, by, k + 1, ka, sm);
          for (int t = 0; t < M::NT; t++) P.phaseB(t, bx, by, k, ka, sm);
        }
      }
  (*launches)++;
  return 0;
}
#endif

// ---- fused residual: face fluxes -> Fp -> projection -> assembly (regular interior) ----------------
// Replaces, for the cells whose whole dependency cone is free of domain-end special cases, the staged
// chain FaceFlux<0,1,2> -> FpCell -> ProjectSNES/ProjectAdd (momentum.c:669-1938, 2297-2331): the 18
// face-flux arrays and Fp never reach HBM.  Cells within 2 (low side) / 3 (high side) layers of a
// domain boundary keep the staged kernels (run on thin slabs by the host), which carry the
// reference's end-of-domain and periodic ghost-copy semantics verbatim.
//
// One thread per node of a TX x TY tile (overlapped tiling), marching k.  At step q a thread computes
// the i- and j-face fluxes of its node on plane q and the k-face flux of face q+1 (one plane ahead:
// the 4th-order divergence of plane q reaches face q+1), all with face_flux_core — the same
// arithmetic as the staged kernels.  ucat and nvert come from a 5-plane shared-memory ring (planes
// q-1..q+3, box (TX+4)x(TY+3)) fed by TMA one step ahead; centre metrics, aj and nu_t of plane q are
// exchanged through shared memory (each node loaded once, coalesced); i-/j-face fluxes and Fp are
// exchanged through shared memory; the k-face flux history (faces q-2..q+1) lives in registers.
// Emission: components x,y of plane q and component z of plane q-1 (needs Fp of planes q-1 and q).
struct RhsMarch {
  static constexpr int TX = 32, TY = 16, NT = TX * TY;
  static constexpr int NXP = TX + 4, NYP = TY + 3, NN = NXP * NYP;
  static constexpr int TILE_D = ((NN * 8 + 127) / 128) * 16, NSC = 4, PLANE_D = NSC * TILE_D, STAGES = 5;
  static constexpr int MXP = TX + 1, MYP = TY + 1, MT_D = ((MXP * MYP + 15) / 16) * 16, NM = 11;
  static constexpr int OFF_M = STAGES * PLANE_D, OFF_F1 = OFF_M + NM * MT_D, OFF_F2 = OFF_F1 + 6 * NT, OFF_FP = OFF_F2 + 6 * NT, OFF_BAR = OFF_FP + 2 * NT;
  static constexpr long SMEM_D = OFF_BAR + 8;
  static constexpr int OX = TX - 4, OY = TY - 4;      // outputs per tile
  struct State { double c3[3][3], v3[2][3], cn[3], vn[3], fp[3], zdot, ajz; };
  VfsDev d; int mode, s0; double scale; Box R;        // R: regular output region, local k, upper bounds exclusive

  static bool region(const VfsDev &d, Box &R) {
    R.i0 = 3; R.i1 = d.mx - 4; R.j0 = 3; R.j1 = d.my - 4;
    int k0 = 3 - d.kofs, k1 = d.mz - 4 - d.kofs;
    R.k0 = k0 < 0 ? 0 : k0; R.k1 = k1 > d.nzl ? d.nzl : k1;
    return R.i1 > R.i0 && R.j1 > R.j0 && R.k1 > R.k0;
  }
  int tiles_x() const { return (R.i1 - R.i0 + OX - 1) / OX; }
  int tiles_y() const { return (R.j1 - R.j0 + OY - 1) / OY; }
  VFS_HD int iorg(int bx) const { return R.i0 - 2 + bx * OX; }
  VFS_HD int jorg(int by) const { return R.j0 - 2 + by * OY; }
  VFS_HD static int slot_off(int kk, int q0) { return ((kk - q0) % STAGES) * PLANE_D; }

  struct RingView {
    const double *ring; int so[4]; int xy;
    VFS_HD double u(int a, int di, int dj, int dk) const { return ring[so[dk + 1] + a * TILE_D + xy + dj * NXP + di]; }
    VFS_HD double nv(int di, int dj, int dk) const { return ring[so[dk + 1] + 3 * TILE_D + xy + dj * NXP + di]; }
  };
  struct AccSM {       // metrics / nu_t of plane q from the exchange buffer
    RingView V; const double *m; int nb; double ucv;
    VFS_HD double u(int a, int di, int dj, int dk) const { return V.u(a, di, dj, dk); }
    VFS_HD double nv(int di, int dj, int dk) const { return V.nv(di, dj, dk); }
    template <int D> VFS_HD double met(int s, int side) const { return m[s * MT_D + side * nb]; }
    template <int D> VFS_HD double iaj(int side) const { return m[9 * MT_D + side * nb]; }
    template <int D> VFS_HD double nut(int side) const { return m[10 * MT_D + side * nb]; }
    template <int D> VFS_HD double uc(int) const { return ucv; }
  };
  struct AccReg {      // metrics / nu_t of the thread's own column on two planes, in registers
    RingView V; double m0[11], m1[11]; double ucv;
    VFS_HD double u(int a, int di, int dj, int dk) const { return V.u(a, di, dj, dk); }
    VFS_HD double nv(int di, int dj, int dk) const { return V.nv(di, dj, dk); }
    template <int D> VFS_HD double met(int s, int side) const { return side ? m1[s] : m0[s]; }
    template <int D> VFS_HD double iaj(int side) const { return side ? m1[9] : m0[9]; }
    template <int D> VFS_HD double nut(int side) const { return side ? m1[10] : m0[10]; }
    template <int D> VFS_HD double uc(int) const { return ucv; }
  };

  VFS_HD void begin(State &st) const {
#pragma unroll
    for (int a = 0; a < 3; a++) { st.c3[0][a] = st.c3[1][a] = st.c3[2][a] = 0; st.v3[0][a] = st.v3[1][a] = 0; st.cn[a] = st.vn[a] = st.fp[a] = 0; }
    st.zdot = 0; st.ajz = 1;
  }
  VFS_HD void load_met(double *sM, int slot, int ii, int jj, int q) const {
    const bool inb = ii < d.mx + VFS_G && jj < d.my + VFS_G;
    const long n = inb ? d.idx(ii, jj, q) : d.idx(0, 0, q);
#pragma unroll
    for (int s = 0; s < 9; s++) sM[s * MT_D + slot] = d.s[S_CSI0 + s][n];
    sM[9 * MT_D + slot] = d.s[S_IAJ][n];
    sM[10 * MT_D + slot] = d.s[S_NUT][n];
  }
  // phase 0: centre metrics, aj, nu_t of plane q -> exchange buffer (tile + one extra column and row)
  VFS_HD void phase0(int tid, int bx, int by, int q, int ka, double *sm) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    if (q < ka) return;
    double *sM = sm + OFF_M;
    load_met(sM, ty * MXP + tx, i, j, q);
    if (tx == TX - 1) load_met(sM, ty * MXP + TX, i + 1, j, q);
    if (ty == TY - 1) load_met(sM, TY * MXP + tx, i, j + 1, q);
  }
  VFS_HD RingView view(const double *sm, int tx, int ty, int kc, int q0, int nlo, int nhi) const {
    RingView V; V.ring = sm; V.xy = (ty + 1) * NXP + tx + 1;
#pragma unroll
    for (int dk = -1; dk <= 2; dk++) V.so[dk + 1] = (dk >= nlo && dk <= nhi) ? slot_off(kc + dk, q0) : 0;
    return V;
  }
  // phase 1a: i- and j-face fluxes of plane q -> exchange buffers
  VFS_HD void phase1a(int tid, int bx, int by, int q, int ka, double *sm) const {
    if (q < ka) return;
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    const bool need1 = ty >= 2 && ty <= TY - 2 && j <= R.j1 && i <= R.i1 + 1;
    const bool need2 = tx >= 2 && tx <= TX - 2 && i <= R.i1 && j <= R.j1 + 1;
    if (!need1 && !need2) return;
    const long p = d.idx(i, j, q);
    const RingView V = view(sm, tx, ty, q, ka - 3, -1, 1);
    const double *m = sm + OFF_M + ty * MXP + tx;
    double fc[3], fv[3];
    if (need1) {
      AccSM A = {V, m, 1, d.s[S_UC0][p]};
      face_flux_core<0, true>(d, A, 0, fc, fv);
      double *sF = sm + OFF_F1;
#pragma unroll
      for (int a = 0; a < 3; a++) { sF[a * NT + tid] = fc[a]; sF[(3 + a) * NT + tid] = fv[a]; }
    }
    if (need2) {
      AccSM A = {V, m, MXP, d.s[S_UC1][p]};
      face_flux_core<1, true>(d, A, 0, fc, fv);
      double *sF = sm + OFF_F2;
#pragma unroll
      for (int a = 0; a < 3; a++) { sF[a * NT + tid] = fc[a]; sF[(3 + a) * NT + tid] = fv[a]; }
    }
  }
  VFS_HD bool need_fp(int tx, int ty, int i, int j) const { return tx >= 2 && tx <= TX - 2 && ty >= 2 && ty <= TY - 2 && i <= R.i1 && j <= R.j1; }
  // phase 1b: k-face flux of face q+1 (between planes q+1 and q+2) -> registers
  VFS_HD void phase1b(State &st, int tid, int bx, int by, int q, int ka, const double *sm) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    if (!need_fp(tx, ty, i, j)) return;
    const long p = d.idx(i, j, q + 1);
    AccReg A;
    A.V = view(sm, tx, ty, q + 1, ka - 3, -1, 2);
#pragma unroll
    for (int s = 0; s < 9; s++) { A.m0[s] = d.s[S_CSI0 + s][p]; A.m1[s] = d.s[S_CSI0 + s][p + d.sk]; }
    A.m0[9] = d.s[S_IAJ][p]; A.m1[9] = d.s[S_IAJ][p + d.sk];
    A.m0[10] = d.s[S_NUT][p]; A.m1[10] = d.s[S_NUT][p + d.sk];
    A.ucv = d.s[S_UC2][p];
    face_flux_core<2, true>(d, A, 0, st.cn, st.vn);
  }
  // phase 2: Fp of plane q (momentum.c:1548-1678, regular branch) -> registers + exchange buffer; shift the k history
  VFS_HD void phase2(State &st, int tid, int bx, int by, int q, int ka, double *sm) const {
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    if (q >= ka && need_fp(tx, ty, i, j)) {
      const RingView V = view(sm, tx, ty, q, ka - 3, -1, 1);
      const double *F1 = sm + OFF_F1 + tid, *F2 = sm + OFF_F2 + tid;
      const double nv0 = V.nv(0, 0, 0);
      const bool c1 = V.nv(-1, 0, 0) + nv0 + V.nv(1, 0, 0) > 0.1, c2 = V.nv(0, -1, 0) + nv0 + V.nv(0, 1, 0) > 0.1, c3 = V.nv(0, 0, -1) + nv0 + V.nv(0, 0, 1) > 0.1;
      double *sFp = sm + OFF_FP;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const double f1 = F1[a * NT], f1m = F1[a * NT - 1], f2 = F2[a * NT], f2m = F2[a * NT - TX];
        const double div = (f1 - f1m + f2 - f2m + st.c3[2][a] - st.c3[1][a]);
        const double vis = (F1[(3 + a) * NT] - F1[(3 + a) * NT - 1] + F2[(3 + a) * NT] - F2[(3 + a) * NT - TX] + st.v3[1][a] - st.v3[0][a]);
        double fp;
        if (!d.second_order) {
          double div4 = 0;
          div4 += c1 ? (f1 - f1m) * 1. : (F1[a * NT + 1] - F1[a * NT - 2]) * (1. / 3.);
          div4 += c2 ? (f2 - f2m) * 1. : (F2[a * NT + TX] - F2[a * NT - 2 * TX]) * (1. / 3.);
          div4 += c3 ? (st.c3[2][a] - st.c3[1][a]) * 1. : (st.cn[a] - st.c3[0][a]) * (1. / 3.);
          fp = (9. / 8.) * div + (-1. / 8.) * div4 + vis;
        } else fp = div + vis;
        st.fp[a] = fp;
      }
      // own half of the projection (momentum.c:1733-1735): 0.5 * (metric . Fp); x and y halves are exchanged
      const double *m = sm + OFF_M + ty * MXP + tx;
      const double f0 = st.fp[0], f1 = st.fp[1], f2 = st.fp[2];
      st.fp[0] = 0.5 * (m[0] * f0 + m[MT_D] * f1 + m[2 * MT_D] * f2);
      st.fp[1] = 0.5 * (m[3 * MT_D] * f0 + m[4 * MT_D] * f1 + m[5 * MT_D] * f2);
      st.fp[2] = 0.5 * (m[6 * MT_D] * f0 + m[7 * MT_D] * f1 + m[8 * MT_D] * f2);
      sFp[tid] = st.fp[0]; sFp[NT + tid] = st.fp[1];
    }
#pragma unroll
    for (int a = 0; a < 3; a++) { st.c3[0][a] = st.c3[1][a]; st.c3[1][a] = st.c3[2][a]; st.c3[2][a] = st.cn[a]; st.v3[0][a] = st.v3[1][a]; st.v3[1][a] = st.vn[a]; }
  }
  // phase 3: projection, masks (momentum.c:1833-1841) and assembly: x,y of plane q, z of plane q-1
  VFS_HD void phase3(State &st, int tid, int bx, int by, int q, int ka, int kb, const double *sm) const {
    if (q < ka) return;
    const int tx = tid % TX, ty = tid / TX, i = iorg(bx) + tx, j = jorg(by) + ty;
    if (!(tx >= 2 && tx <= TX - 3 && ty >= 2 && ty <= TY - 3 && i < R.i1 && j < R.j1)) return;
    const long p = d.idx(i, j, q);
    const bool exy = q < kb, ez = q > ka;
    // all global operands first (independent loads), arithmetic and stores afterwards
    SnesIn in0, in1, in2; double old0 = 0, old1 = 0, old2 = 0;
    in0.uc = in0.uco = in0.ucm = in0.ro = in0.dp = in0.fe = 0; in1 = in0; in2 = in0;
    if (mode == 0) { if (exy) { old0 = d.s[s0][p]; old1 = d.s[s0 + 1][p]; } if (ez) old2 = d.s[s0 + 2][p - d.sk]; }
    else { if (exy) { in0 = snes_inputs(d, 0, p); in1 = snes_inputs(d, 1, p); } if (ez) in2 = snes_inputs(d, 2, p - d.sk); }
    const RingView V = view(sm, tx, ty, q, ka - 3, -1, 1);
    const double *m = sm + OFF_M + ty * MXP + tx, *sFp = sm + OFF_FP + tid;
    const double ia = m[9 * MT_D], nv0 = V.nv(0, 0, 0);
    const double r0 = (st.fp[0] + sFp[1]) * (2. / (ia + m[9 * MT_D + 1]));
    const double r1 = (st.fp[1] + sFp[NT + TX]) * (2. / (ia + m[9 * MT_D + MXP]));
    const double r2 = (st.zdot + st.fp[2]) * (2. / (st.ajz + ia));
    const bool k0 = nv0 + V.nv(1, 0, 0) > 0.1, k1 = nv0 + V.nv(0, 1, 0) > 0.1, k2 = V.nv(0, 0, -1) + nv0 > 0.1;
    st.zdot = st.fp[2]; st.ajz = ia;
    if (mode == 0) {
      if (exy) { d.s[s0][p] = k0 ? 0. : old0 + scale * r0; d.s[s0 + 1][p] = k1 ? 0. : old1 + scale * r1; }
      if (ez) d.s[s0 + 2][p - d.sk] = k2 ? 0. : old2 + scale * r2;
    } else {
      if (exy) { d.s[S_R0][p] = snes_combine(d, in0, k0, r0); d.s[S_R1][p] = snes_combine(d, in1, k1, r1); }
      if (ez) d.s[S_R2][p - d.sk] = snes_combine(d, in2, k2, r2);
    }
  }
};

#ifndef VFS_EMU
__global__ void __launch_bounds__(RhsMarch::NT, 1) k_rhs_march(const __grid_constant__ CUtensorMap tmap, const RhsMarch P, int kchunk) {
  extern __shared__ __align__(128) double vfs_rhs_sm[];
  double *sm = vfs_rhs_sm;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + RhsMarch::OFF_BAR);
  constexpr int STAGES = RhsMarch::STAGES;
  const int tid = threadIdx.x, bx = blockIdx.x, by = blockIdx.y;
  const int ka = P.R.k0 + blockIdx.z * kchunk, kb = min(P.R.k1, ka + kchunk);
  if (ka >= kb) return;
  const int q0 = ka - 3, klast = kb + 3;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int kk) {
    const int slot = (kk - q0) % STAGES;
    double *dst = sm + slot * RhsMarch::PLANE_D;
    mbar_expect_tx(&bars[slot], RhsMarch::NSC * RhsMarch::NN * 8);
    const int sid[4] = {S_U0, S_U1, S_U2, S_NV};
#pragma unroll
    for (int s = 0; s < 4; s++) tma_load_tile(dst + s * RhsMarch::TILE_D, &tmap, P.iorg(bx) - 1 + VFS_G, P.jorg(by) - 1 + VFS_G, kk + VFS_G, sid[s], &bars[slot]);
  };
  auto wait_plane = [&](int kk) { const int n = kk - q0; mbar_wait(&bars[n % STAGES], (n / STAGES) & 1); };
  if (tid == 0) for (int kk = q0; kk <= q0 + 4 && kk <= klast; kk++) issue(kk);
  RhsMarch::State st;
  P.begin(st);
  for (int kk = q0; kk < q0 + 3; kk++) wait_plane(kk);
  for (int q = q0; q <= kb; q++) {
    P.phase0(tid, bx, by, q, ka, sm);
    __syncthreads();
    P.phase1a(tid, bx, by, q, ka, sm);
    wait_plane(q + 3);
    P.phase1b(st, tid, bx, by, q, ka, sm);
    __syncthreads();
    P.phase2(st, tid, bx, by, q, ka, sm);
    __syncthreads();
    P.phase3(st, tid, bx, by, q, ka, kb, sm);
    __syncthreads();
    if (tid == 0 && q > q0 && q + 4 <= klast) { fence_proxy_async(); issue(q + 4); }
  }
}
static inline int run_rhs_march(cudaStream_t stream, const CUtensorMap &tmap, const RhsMarch &P, long *launches) {
  static bool attr_set = false;
  const int bytes = (int)(RhsMarch::SMEM_D * sizeof(double));
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_rhs_march, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return -2;
    attr_set = true;
  }
  const int ntx = P.tiles_x(), nty = P.tiles_y(), nk = P.R.k1 - P.R.k0;
  const int kchunk = pick_kchunk(ntx * nty, nk, 24);
  dim3 grd(ntx, nty, (nk + kchunk - 1) / kchunk);
  k_rhs_march<<<grd, RhsMarch::NT, bytes, stream>>>(tmap, P, kchunk);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#else
static inline int run_rhs_march(void *, const RhsMarch &P, long *launches) {
  typedef RhsMarch M;
  std::vector<double> smv(M::SMEM_D);
  std::vector<M::State> st(M::NT);
  double *sm = smv.data();
  const VfsDev &d = P.d;
  const int ntx = P.tiles_x(), nty = P.tiles_y(), nk = P.R.k1 - P.R.k0;
  const int kchunk = pick_kchunk(ntx * nty, nk, 4, 7);      // small odd chunks: exercise the chunk seams
  for (int bz = 0; bz * kchunk < nk; bz++)
    for (int by = 0; by < nty; by++)
      for (int bx = 0; bx < ntx; bx++) {
        const int ka = P.R.k0 + bz * kchunk, kb = P.R.k1 < ka + kchunk ? P.R.k1 : ka + kchunk;
        const int q0 = ka - 3, klast = kb + 3;
        auto issue = [&](int kk) {      // what the TMA unit does: box copy with zero fill outside the padded array
          double *dst = sm + ((kk - q0) % M::STAGES) * M::PLANE_D;
          const int sid[4] = {S_U0, S_U1, S_U2, S_NV};
          for (int s = 0; s < 4; s++)
            for (int y = 0; y < M::NYP; y++)
              for (int x = 0; x < M::NXP; x++) {
                const int X = P.iorg(bx) - 1 + VFS_G + x, Y = P.jorg(by) - 1 + VFS_G + y, Z = kk + VFS_G;
                const bool in = X >= 0 && X < d.pitch && Y >= 0 && Y < d.ny && Z >= 0 && Z < d.nzt;
                dst[s * M::TILE_D + y * M::NXP + x] = in ? d.s[sid[s]][(long)Z * d.sk + (long)Y * d.sj + X] : 0.;
              }
        };
        for (int kk = q0; kk <= q0 + 4 && kk <= klast; kk++) issue(kk);
        for (int t = 0; t < M::NT; t++) P.begin(st[t]);
        for (int q = q0; q <= kb; q++) {
          for (int t = 0; t < M::NT; t++) P.phase0(t, bx, by, q, ka, sm);
          for (int t = 0; t < M::NT; t++) P.phase1a(t, bx, by, q, ka, sm);
          for (int t = 0; t < M::NT; t++) P.phase1b(st[t], t, bx, by, q, ka, sm);
          for (int t = 0; t < M::NT; t++) P.phase2(st[t], t, bx, by, q, ka, sm);
          for (int t = 0; t < M::NT; t++) P.phase3(st[t], t, bx, by, q, ka, kb, sm);
          if (q > q0 && q + 4 <= klast) issue(q + 4);
        }
      }
  (*launches)++;
  return 0;
}
#endif

#endif
