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

#ifndef VFS_EMU
template <class P, int PH> struct PhaseSeq {
  static __device__ __forceinline__ void run(const P &prog, typename P::State &st, int tid, int bx, int by, int k, double *sm) {
    prog.template phase<PH>(st, tid, bx, by, k, sm);
    if (PH + 1 < P::NPH || P::SYNC_AFTER_LAST) __syncthreads();
    if constexpr (PH + 1 < P::NPH) PhaseSeq<P, PH + 1>::run(prog, st, tid, bx, by, k, sm);
  }
};
template <class P> __global__ void __launch_bounds__(P::NT, 1) k_block_march(const P prog, int kbeg, int kend, int kchunk) {
  extern __shared__ __align__(16) double vfs_march_sm[];
  const int tid = threadIdx.x;
  const int ka = kbeg + blockIdx.z * kchunk, kb = min(kend, ka + kchunk);
  if (ka >= kb) return;
  typename P::State st;
  prog.begin(st, tid, blockIdx.x, blockIdx.y, ka, kb, vfs_march_sm);
  for (int k = ka - P::LEAD; k < kb; k++) PhaseSeq<P, 0>::run(prog, st, tid, blockIdx.x, blockIdx.y, k, vfs_march_sm);
}
template <class P> static inline int run_block_march(cudaStream_t st, const P &prog, const MarchGrid &g, long *launches) {
  if (g.kend <= g.kbeg || g.nbx <= 0 || g.nby <= 0) return 0;
  static bool attr_set = false;
  const int bytes = (int)(P::SMEM_D * sizeof(double));
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_block_march<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return -2;
    attr_set = true;
  }
  dim3 grd(g.nbx, g.nby, (g.kend - g.kbeg + g.kchunk - 1) / g.kchunk), blk(P::NT, 1, 1);
  k_block_march<P><<<grd, blk, bytes, st>>>(prog, g.kbeg, g.kend, g.kchunk);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#else
template <class P, int PH> struct PhaseSeqEmu {
  static void run(const P &prog, std::vector<typename P::State> &st, int bx, int by, int k, double *sm) {
    for (int tid = 0; tid < P::NT; tid++) prog.template phase<PH>(st[tid], tid, bx, by, k, sm);
    if constexpr (PH + 1 < P::NPH) PhaseSeqEmu<P, PH + 1>::run(prog, st, bx, by, k, sm);
  }
};
template <class P> static inline int run_block_march(void *, const P &prog, const MarchGrid &g, long *launches) {
  if (g.kend <= g.kbeg || g.nbx <= 0 || g.nby <= 0) return 0;
  std::vector<double> sm(P::SMEM_D);
  std::vector<typename P::State> st(P::NT);
  const int nz = (g.kend - g.kbeg + g.kchunk - 1) / g.kchunk;
  for (int bz = 0; bz < nz; bz++)
    for (int by = 0; by < g.nby; by++)
      for (int bx = 0; bx < g.nbx; bx++) {
        const int ka = g.kbeg + bz * g.kchunk, kb = g.kend < ka + g.kchunk ? g.kend : ka + g.kchunk;
        for (int tid = 0; tid < P::NT; tid++) prog.begin(st[tid], tid, bx, by, ka, kb, sm.data());
        for (int k = ka - P::LEAD; k < kb; k++) PhaseSeqEmu<P, 0>::run(prog, st, bx, by, k, sm.data());
      }
  (*launches)++;
  return 0;
}
#endif

// ---- LES pass 2 (les.c:308-669): separable Simpson test filters + Germano contraction ---------------
// The reference filters 16 per-node products (w, w U_a u_b, w |S|S_ij) with the 27-point
// (1,4,1)^3 Simpson stencil, one 27-term sum per product and cell (rhs2.c:499-523).  The stencil is
// a tensor product, so the same sums are formed here as three 3-point passes: along k from the
// thread's own column (global loads, coalesced along i), along i and along j through two
// shared-memory exchange buffers — 4 shared loads + 2 stores per product and cell instead of 27
// loads, which is what bounded the 27-term form (shared-memory bandwidth, profiles/r01c).  The
// summation order differs from the reference's, i.e. results agree to rounding (~1e-15 relative),
// not bitwise.  sum_weight (les.c:441-468, coefficients (1,2,1)^3/8) rides along as product 16.
// The tensor algebra that follows the filters (les2_finish) runs in the last phase.
struct Les2Sep {
  static constexpr int TX = 32, TY = 16, NT = TX * TY, NV = 17, NPH = 3, LEAD = 0;
  static constexpr bool SYNC_AFTER_LAST = false;   // phase 0 of the next step does not touch what phase 2 reads
  static constexpr long SMEM_D = 2L * NV * NT;
  struct State { double v[NV]; };
  VfsDev d;
  static int tiles_x(const VfsDev &d) { return (d.mx - 2 + TX - 3) / (TX - 2); }
  static int tiles_y(const VfsDev &d) { return (d.my - 2 + TY - 3) / (TY - 2); }
  VFS_HD void begin(State &, int, int, int, int, int, double *) const {}
  template <int PH> VFS_HD void phase(State &st, int tid, int bx, int by, int k, double *sm) const {
    const int tx = tid % TX, ty = tid / TX;
    const int i = bx * (TX - 2) + tx, j = by * (TY - 2) + ty;      // node of this thread (tile halo included)
    double *sK = sm, *sA = sm + NV * NT;
    if (PH == 0) {            // k pass over the thread's own column
      double K[NV];
#pragma unroll
      for (int a = 0; a < NV; a++) K[a] = 0;
      if (i <= d.mx - 1 && j <= d.my - 1) {
        const long p = d.idx(i, j, k);
#pragma unroll
        for (int dk = -1; dk <= 1; dk++) {
          const long n = p + dk * d.sk;
          const double w = d.s[S_LW][n];
          const double sw = dk == 0 ? 4. * w : w;
          const double u0 = d.s[S_U0][n], u1 = d.s[S_U1][n], u2 = d.s[S_U2][n];
          const double U0 = d.s[S_LU0][n], U1 = d.s[S_LU1][n], U2 = d.s[S_LU2][n];
          K[0] += sw;
          K[1] += sw * (U0 * u0); K[2] += sw * (U0 * u1); K[3] += sw * (U0 * u2);
          K[4] += sw * (U1 * u0); K[5] += sw * (U1 * u1); K[6] += sw * (U1 * u2);
          K[7] += sw * (U2 * u0); K[8] += sw * (U2 * u1); K[9] += sw * (U2 * u2);
#pragma unroll
          for (int a = 0; a < 6; a++) K[10 + a] += sw * d.s[S_LSS0 + a][n];
          K[16] += dk == 0 ? w : 0.5 * w;
        }
      }
#pragma unroll
      for (int a = 0; a < NV; a++) { st.v[a] = K[a]; sK[a * NT + tid] = K[a]; }
    } else if (PH == 1) {     // i pass (tile-edge columns produce unused values)
      const int l = tx > 0 ? tid - 1 : tid, r = tx < TX - 1 ? tid + 1 : tid;
#pragma unroll
      for (int a = 0; a < NV; a++) {
        const double A = a < 16 ? sK[a * NT + l] + 4. * st.v[a] + sK[a * NT + r] : 0.5 * sK[a * NT + l] + st.v[a] + 0.5 * sK[a * NT + r];
        st.v[a] = A; sA[a * NT + tid] = A;
      }
    } else {                  // j pass + les.c:441-669 for the inner nodes of the tile
      if (tx < 1 || tx > TX - 2 || ty < 1 || ty > TY - 2 || i > d.mx - 2 || j > d.my - 2) return;
      const long p = d.idx(i, j, k);
      if (d.s[S_NV][p] > 1.1) { d.s[S_LM][p] = 0; d.s[S_MM][p] = 0; return; }
      const int up = tid - TX, dn = tid + TX;
      double fs[16];
#pragma unroll
      for (int a = 0; a < 16; a++) fs[a] = sA[a * NT + up] + 4. * st.v[a] + sA[a * NT + dn];
      const double sum_weight = 0.5 * sA[16 * NT + up] + st.v[16] + 0.5 * sA[16 * NT + dn];
      les2_finish(d, i, j, k + d.kofs, p, fs, sum_weight);
    }
  }
};
static inline MarchGrid les2_sep_grid(const VfsDev &d, int k0, int k1) {
  MarchGrid g = {Les2Sep::tiles_x(d), Les2Sep::tiles_y(d), k0, k1, 1};
  g.kchunk = pick_kchunk(g.nbx * g.nby, k1 - k0, 16);
  return g;
}

#endif
