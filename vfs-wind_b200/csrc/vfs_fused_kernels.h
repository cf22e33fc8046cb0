// vfs_fused_kernels.h — shared-memory-tiled k-marching kernels (the performance path).
//
// A thread block owns an (i,j) tile and marches along k.  The per-node planes a cell's 27-point
// neighbourhood needs are staged into a ring of shared-memory planes by TMA
// (cp.async.bulk.tensor.4d, one box of (TX+2)x(TY+2) doubles per scalar per plane, completion on
// an mbarrier), STAGES-2 planes ahead of the plane being computed, so HBM/L2 latency is hidden by
// the prefetch distance instead of by occupancy.  Out-of-range box coordinates (tile overhang) are
// zero-filled by the TMA unit.  The arithmetic is the same device functions, in the same summation
// order, as the one-thread-per-cell kernels, so both forms are bitwise identical
// (tests/test_gpu_parity.py::test_tiled_equals_staged).
// CUDA only: the host emulation (tests/emu, -DVFS_EMU) always uses the staged kernels.
#ifndef VFS_FUSED_KERNELS_H
#define VFS_FUSED_KERNELS_H
#include "vfs_common.h"
#include "vfs_les_kernels.h"
#include "vfs_rhs_kernels.h"

#ifndef VFS_EMU
#include <cuda.h>
#include <cuda_runtime.h>

#define VFS_TILE_TX 32
#define VFS_TILE_TY 8

// ---- TMA / mbarrier primitives (raw PTX) -------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "VFS_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra VFS_DONE_%=;\n"
      "bra VFS_WAIT_%=;\n"
      "VFS_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one (box_x, box_y, 1, 1) tile of scalar `w` at padded coordinates (x, y, z) -> shared memory
__device__ __forceinline__ void tma_load_tile(void *dst, const CUtensorMap *map, int x, int y, int z, int w, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
               "l"((unsigned long long)map), "r"(x), "r"(y), "r"(z), "r"(w), "r"(smem_u32(bar))
               : "memory");
}

// host: 4-D tensor map over the whole scalar pool [S_COUNT][nzt][ny][pitch], box (bx, by, 1, 1)
static inline int vfs_make_tensor_map(CUtensorMap *map, void *pool, const VfsDev &d, long scalar_len, int bx, int by) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = 0;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
  cuuint64_t gdim[4] = {(cuuint64_t)d.pitch, (cuuint64_t)d.ny, (cuuint64_t)d.nzt, (cuuint64_t)S_TAIL0};      // the main pool (TMA never touches the tail scalars)
  cuuint64_t gstr[3] = {(cuuint64_t)d.pitch * 8, (cuuint64_t)d.sk * 8, (cuuint64_t)scalar_len * 8};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ((encode_fn)fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, pool, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// ---- generic ring of TMA-staged planes ----------------------------------------------------------------
// NS scalars per plane; tile of (TX+HX) x (TY+HY) nodes whose origin is (OX, OY) nodes before the
// block's first cell; each scalar tile is padded to a 128-byte multiple (TMA destination alignment).
template <int TX_, int TY_, int NS_, int STAGES_, int HX, int HY, int OX_, int OY_, int PB_, int PA_> struct Ring {
  static constexpr int TX = TX_, TY = TY_, NS = NS_, STAGES = STAGES_, OX = OX_, OY = OY_, PB = PB_, PA = PA_;
  static constexpr int NXP = TX + HX, NYP = TY + HY, NN = NXP * NYP;
  static constexpr int TILE_B = ((NN * 8 + 127) / 128) * 128;
  static constexpr int TILE_D = TILE_B / 8;                 // doubles per padded scalar tile
  static constexpr int PLANE_D = TILE_D * NS;
  static constexpr size_t BYTES = (size_t)STAGES * PLANE_D * 8 + STAGES * 8;
  static_assert(STAGES >= PB + PA + 1, "ring too short");
  static_assert((NXP * 8) % 16 == 0, "TMA box inner extent must be a multiple of 16 bytes");
};
struct SidList { int n; int sid[16]; };

// accessor over the ring: scalar slot s at node offset (di,dj,dk) from the thread's cell
template <class R> struct TileAcc {
  const double *sm; int n0, x, y;      // n0: ring index of the cell's own plane; (x,y): tile position incl. halo
  __device__ __forceinline__ double get(int s, int di, int dj, int dk) const {
    return sm[(size_t)((n0 + dk) % R::STAGES) * R::PLANE_D + s * R::TILE_D + (y + dj) * R::NXP + (x + di)];
  }
};

// Block = (i,j) tile, marches k in [ka,kb).  At step k planes k-PB..k+PA are resident; the slot of
// plane k-PB is refilled (plane k-PB+STAGES) as soon as the whole block has finished step k.
template <class R, class Body, int MINB = 2>
__global__ void __launch_bounds__(R::TX *R::TY, MINB)
k_tile_march(const __grid_constant__ CUtensorMap tmap, VfsDev d, int kbeg, int kend, int kchunk, SidList sl, Body body) {
  extern __shared__ __align__(128) unsigned char smraw[];
  double *sm = reinterpret_cast<double *>(smraw);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smraw + (size_t)R::STAGES * R::PLANE_D * 8);
  constexpr int TX = R::TX, TY = R::TY, STAGES = R::STAGES, PB = R::PB, PA = R::PA;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int ka = kbeg + blockIdx.z * kchunk;
  const int kb = min(kend, ka + kchunk);
  if (ka >= kb) return;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = (i <= d.mx - 2) && (j <= d.my - 2);
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int kfirst = ka - PB, klast = kb - 1 + PA;
  // ring index n = kk - kfirst: plane kk in slot n % STAGES, barrier phase (n / STAGES) & 1
  auto issue = [&](int kk) {
    const int n = kk - kfirst, slot = n % STAGES;
    double *dst = sm + (size_t)slot * R::PLANE_D;
    mbar_expect_tx(&bars[slot], R::NS * R::NN * 8);
    for (int s = 0; s < R::NS; s++) tma_load_tile(dst + s * R::TILE_D, &tmap, i0 - R::OX + VFS_G, j0 - R::OY + VFS_G, kk + VFS_G, sl.sid[s], &bars[slot]);
  };
  auto wait_plane = [&](int kk) {
    const int n = kk - kfirst;
    mbar_wait(&bars[n % STAGES], (n / STAGES) & 1);
  };
  if (tid == 0) {
    for (int kk = kfirst; kk < kfirst + STAGES && kk <= klast; kk++) issue(kk);
  }
  for (int kk = kfirst; kk < ka + PA; kk++) wait_plane(kk);
  typename Body::State st;                  // per-thread state a body carries along its k column (e.g. a filter window)
  for (int k = ka; k < kb; k++) {
    // operands the body reads from global memory at its own node are requested before the TMA wait
    const auto pf = body.prefetch(d, active ? i : 1, active ? j : 1, k);
    wait_plane(k + PA);
    if (active) {
      TileAcc<R> A = {sm, k - kfirst, tx + R::OX, ty + R::OY};
      body(d, A, i, j, k, pf, st, k == ka);
    }
    __syncthreads();                         // plane k-PB is no longer needed by anyone
    if (tid == 0) {
      const int kn = k - PB + STAGES;
      if (kn <= klast) { fence_proxy_async(); issue(kn); }
    }
  }
}

template <class R, class Body, int MINB = 2>
static inline int launch_tile_march(cudaStream_t st, const CUtensorMap &tmap, const VfsDev &d, int k0, int k1, int kchunk, const SidList &sl, const Body &body, long *launches) {
  if (k1 <= k0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_tile_march<R, Body, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::BYTES) != cudaSuccess) return -2;
    attr_set = true;
  }
  dim3 grd((d.mx - 2 + R::TX - 1) / R::TX, (d.my - 2 + R::TY - 1) / R::TY, (k1 - k0 + kchunk - 1) / kchunk), blk(R::TX, R::TY, 1);
  k_tile_march<R, Body, MINB><<<grd, blk, R::BYTES, st>>>(tmap, d, k0, k1, kchunk, sl, body);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- LES pass 1 (les.c:199-246): staged ucat(3), aj, nvert; 27-point box ------------------------------
typedef Ring<VFS_TILE_TX, VFS_TILE_TY, 5, 4, 2, 2, 1, 1, 1, 1> RingLes1;
struct Les1Pre { double m[9], aj; unsigned char nearv; };
struct Les1Acc {
  TileAcc<RingLes1> T; const VfsDev &d; long p; const Les1Pre &pf;
  __device__ __forceinline__ double met(int s) const { return pf.m[s]; }
  __device__ __forceinline__ double aj() const { return pf.aj; }
  __device__ __forceinline__ unsigned char nearv(const VfsDev &, long) const { return pf.nearv; }
  __device__ __forceinline__ double u(int a, int di, int dj, int dk) const { return T.get(a, di, dj, dk); }
  __device__ __forceinline__ double iaj(int di, int dj, int dk) const { return T.get(3, di, dj, dk); }
  __device__ __forceinline__ double nv(int di, int dj, int dk) const { return T.get(4, di, dj, dk); }
};
struct Les1Body {
  typedef Les1Win State;
  __device__ __forceinline__ Les1Pre prefetch(const VfsDev &d, int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    Les1Pre f;
#pragma unroll
    for (int s = 0; s < 9; s++) f.m[s] = d.s[S_CSI0 + s][p];
    f.aj = d.s[S_AJ][p]; f.nearv = d.near[p];
    return f;
  }
  __device__ __forceinline__ void operator()(const VfsDev &d, const TileAcc<RingLes1> &T, int i, int j, int k, const Les1Pre &pf, Les1Win &st, bool first) const {
    const long p = d.idx(i, j, k);
    Les1Acc A = {T, d, p, pf};
    les1_core(d, A, i, j, d.kglob(k), p, &st, first);      // (kglob: a ghost plane across the periodic seam is evaluated as the plane it images)
  }
};

struct NoPrefetch {};
struct NoState {};
// ---- face fluxes of Formfunction_2 (momentum.c:669-1451), regular faces ---------------------------------
// One thread per node computes its i-, j- and k-face fluxes (Fc, Fv: 18 doubles) from ucat/nvert
// planes k-1..k+2: box (TX+4) x (TY+3) with origin (i0-1, j0-1), i.e. node offsets -1..TX+2 in i
// (4th-order stencil of the i-face) and -1..TY+1 in j.  No redundant work: every face is computed
// exactly once.  Faces with index 0 or m-2 along their normal (domain-end / periodic-end stencils) are
// left to the staged FaceFlux<D> kernels, which the host runs on those thin slabs only.
typedef Ring<VFS_TILE_TX, VFS_TILE_TY, 4, 5, 4, 3, 1, 1, 1, 2> RingFlux;
// FLUID = true: every nvert the face stencils read is known to be 0 (VfsDev::near), so nv() is the
// literal 0 and the compiler drops the stencil switches, the QUICK / collapse branches and the
// one-sided nu_t picks; the surviving arithmetic is the masked code's with all predicates false.
template <bool FLUID> struct FluxAccT {
  TileAcc<RingFlux> T; const VfsDev &d; long p;
  __device__ __forceinline__ double u(int a, int di, int dj, int dk) const { return T.get(a, di, dj, dk); }
  __device__ __forceinline__ double nv(int di, int dj, int dk) const { return FLUID ? 0. : T.get(3, di, dj, dk); }
  template <int D> __device__ __forceinline__ long sn() const { return D == 0 ? 1 : (D == 1 ? d.sj : d.sk); }
  template <int D> __device__ __forceinline__ double met(int s, int side) const { return d.s[S_CSI0 + s][p + side * sn<D>()]; }
  template <int D> __device__ __forceinline__ double iaj(int side) const { return d.s[S_IAJ][p + side * sn<D>()]; }
  template <int D> __device__ __forceinline__ double nut(int side) const { return d.s[S_NUT][p + side * sn<D>()]; }
  template <int D> __device__ __forceinline__ double uc(int off) const { return d.s[S_UC0 + D][p + off * sn<D>()]; }
};
typedef FluxAccT<false> FluxAcc;
struct FluxBody {
  template <bool FLUID> __device__ __forceinline__ void run(const VfsDev &d, const TileAcc<RingFlux> &T, int i, int j, int kg, long p) const {
    FluxAccT<FLUID> A = {T, d, p};
    double fc[3], fv[3];
    if (i <= d.mx - 3) {
      face_flux_core<0, true>(d, A, i, fc, fv);
#pragma unroll
      for (int a = 0; a < 3; a++) { d.s[S_FC1 + a][p] = fc[a]; d.s[S_FV1 + a][p] = fv[a]; }
    }
    if (j <= d.my - 3) {
      face_flux_core<1, true>(d, A, j, fc, fv);
#pragma unroll
      for (int a = 0; a < 3; a++) { d.s[S_FC2 + a][p] = fc[a]; d.s[S_FV2 + a][p] = fv[a]; }
    }
    if (kg <= d.mz - 3) {
      face_flux_core<2, true>(d, A, kg, fc, fv);
#pragma unroll
      for (int a = 0; a < 3; a++) { d.s[S_FC3 + a][p] = fc[a]; d.s[S_FV3 + a][p] = fv[a]; }
    }
  }
  typedef NoState State;
  __device__ __forceinline__ NoPrefetch prefetch(const VfsDev &, int, int, int) const { return NoPrefetch(); }
  __device__ __forceinline__ void operator()(const VfsDev &d, const TileAcc<RingFlux> &T, int i, int j, int k, const NoPrefetch &, NoState &, bool) const {
    const long p = d.idx(i, j, k);
    if (VFS_WARP_ANY(d.near[p] != 0)) run<false>(d, T, i, j, k + d.kofs, p);
    else run<true>(d, T, i, j, k + d.kofs, p);
  }
};


// ---- face fluxes, second form: metric planes staged too ------------------------------------------------
// profiles/r01p: with the mask-free fast path the tiled flux kernel waits on its ~69 scattered global
// loads per node (centre metrics and 1/aj at p, p+1, p+sj, p+sk: long-scoreboard 6.6 of 12.8 stall
// cycles per issued instruction).  Here the ten metric scalars of planes k and k+1 are a second
// TMA-fed ring (box (TX+2) x (TY+1) from node i0-1: the inner start coordinate of every box stays even, i.e.
// 16-byte aligned, like all other boxes here; three stages: one plane of prefetch), so a node's metric reads
// are shared-memory loads and every metric value crosses L2 -> SM once per tile instead of up to
// four times.  To make room the tile is 32 x 16 (one 512-thread block per SM, the same 16 warps),
// and nvert leaves the ring: only the masked path reads it (from global memory).
struct FluxMarch {
  static constexpr int TX = 32, TY = 16, NT = TX * TY;
  static constexpr int AX = TX + 4, AY = TY + 3, A_NS = 3, A_ST = 5, A_TILE = ((AX * AY * 8 + 127) / 128) * 16, A_PLANE = A_TILE * A_NS;
  static constexpr int BX = TX + 2, BY = TY + 1, B_NS = 10, B_ST = 3, B_TILE = ((BX * BY * 8 + 127) / 128) * 16, B_PLANE = B_TILE * B_NS;
  static constexpr int OFF_B = A_ST * A_PLANE, OFF_BAR = OFF_B + B_ST * B_PLANE;
  static constexpr size_t BYTES = (size_t)(OFF_BAR + A_ST + B_ST) * 8;
  VFS_HD static int b_sid(int q) { return q < 9 ? S_CSI0 + q : S_IAJ; }
};
template <bool FLUID> struct FluxAcc2 {
  const double *sa[4];      // ucat planes k-1, k, k+1, k+2 at the thread's node
  const double *sb[2];      // metric planes k, k+1 at the thread's node
  const VfsDev &d; long p;
  double ucv[3], nutv[4];   // ucont (x, y, z) at p and nu_t at p, p+1, p+sj, p+sk: fetched before the TMA waits (see k_flux_march)
  __device__ __forceinline__ double u(int a, int di, int dj, int dk) const { return sa[dk + 1][a * FluxMarch::A_TILE + dj * FluxMarch::AX + di]; }
  __device__ __forceinline__ double nv(int di, int dj, int dk) const { return FLUID ? 0. : d.s[S_NV][p + di + dj * d.sj + dk * d.sk]; }
  template <int D> __device__ __forceinline__ long sn() const { return D == 0 ? 1 : (D == 1 ? d.sj : d.sk); }
  template <int D> __device__ __forceinline__ double met(int s, int side) const {
    return sb[D == 2 ? side : 0][s * FluxMarch::B_TILE + (D == 1 ? side * FluxMarch::BX : 0) + (D == 0 ? side : 0)];
  }
  template <int D> __device__ __forceinline__ double iaj(int side) const { return met<D>(9, side); }
  template <int D> __device__ __forceinline__ double nut(int side) const { return side ? nutv[1 + D] : nutv[0]; }
  template <int D> __device__ __forceinline__ double uc(int off) const { return off == 0 ? ucv[D] : d.s[S_UC0 + D][p + off * sn<D>()]; }
};
template <bool FLUID, class Acc> __device__ __forceinline__ void flux_node(const VfsDev &d, const Acc &A, int i, int j, int kg, long p) {
  double fc[3], fv[3];
  if (i <= d.mx - 3) {
    face_flux_core<0, true>(d, A, i, fc, fv);
#pragma unroll
    for (int a = 0; a < 3; a++) { d.s[S_FC1 + a][p] = fc[a]; d.s[S_FV1 + a][p] = fv[a]; }
  }
  if (j <= d.my - 3) {
    face_flux_core<1, true>(d, A, j, fc, fv);
#pragma unroll
    for (int a = 0; a < 3; a++) { d.s[S_FC2 + a][p] = fc[a]; d.s[S_FV2 + a][p] = fv[a]; }
  }
  if (kg <= d.mz - 3) {
    face_flux_core<2, true>(d, A, kg, fc, fv);
#pragma unroll
    for (int a = 0; a < 3; a++) { d.s[S_FC3 + a][p] = fc[a]; d.s[S_FV3 + a][p] = fv[a]; }
  }
}
__global__ void __launch_bounds__(FluxMarch::NT, 1)
k_flux_march(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, VfsDev d, int kbeg, int kend, int kchunk) {
  typedef FluxMarch M;
  extern __shared__ __align__(128) unsigned char smraw[];
  double *sm = reinterpret_cast<double *>(smraw);
  unsigned long long *barA = reinterpret_cast<unsigned long long *>(sm + M::OFF_BAR), *barB = barA + M::A_ST;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * M::TX + tx;
  const int i0 = 1 + blockIdx.x * M::TX, j0 = 1 + blockIdx.y * M::TY;
  const int ka = kbeg + blockIdx.z * kchunk, kb = min(kend, ka + kchunk);
  if (ka >= kb) return;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = (i <= d.mx - 2) && (j <= d.my - 2);
  if (tid == 0) {
    for (int s = 0; s < M::A_ST + M::B_ST; s++) mbar_init(&barA[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int kfa = ka - 1, kla = kb + 1;       // ucat planes ka-1 .. kb+1
  const int kfb = ka, klb = kb;               // metric planes ka .. kb
  auto issueA = [&](int kk) {
    const int slot = (kk - kfa) % M::A_ST;
    mbar_expect_tx(&barA[slot], M::A_NS * M::AX * M::AY * 8);
    for (int s = 0; s < M::A_NS; s++) tma_load_tile(sm + slot * M::A_PLANE + s * M::A_TILE, &tmapA, i0 - 1 + VFS_G, j0 - 1 + VFS_G, kk + VFS_G, S_U0 + s, &barA[slot]);
  };
  auto issueB = [&](int kk) {
    const int slot = (kk - kfb) % M::B_ST;
    mbar_expect_tx(&barB[slot], M::B_NS * M::BX * M::BY * 8);
    for (int s = 0; s < M::B_NS; s++) tma_load_tile(sm + M::OFF_B + slot * M::B_PLANE + s * M::B_TILE, &tmapB, i0 - 1 + VFS_G, j0 + VFS_G, kk + VFS_G, M::b_sid(s), &barB[slot]);
  };
  if (tid == 0) {
    for (int kk = kfa; kk < kfa + M::A_ST && kk <= kla; kk++) issueA(kk);
    for (int kk = kfb; kk < kfb + M::B_ST && kk <= klb; kk++) issueB(kk);
  }
  for (int kk = kfa; kk < ka + 2; kk++) mbar_wait(&barA[(kk - kfa) % M::A_ST], ((kk - kfa) / M::A_ST) & 1);
  mbar_wait(&barB[0], 0);
  // the branch flag is fetched one plane ahead (carrying ucont / nu_t the same way costs registers the kernel
  // does not have: 1.37 vs 1.28 ms with the resulting spills)
  unsigned char near_next = d.near[d.idx(active ? i : 1, active ? j : 1, ka)];
  for (int k = ka; k < kb; k++) {
    // the few operands that stay in global memory are requested first: their latency overlaps the TMA waits
    // and the stencil differences (profiles/r01q: issued at their point of use they cost 5.6 of 13 stall cycles)
    const long p = d.idx(active ? i : 1, active ? j : 1, k);
    const unsigned char nearv = near_next;
    near_next = d.near[p + d.sk];
    const double uc0 = d.s[S_UC0][p], uc1 = d.s[S_UC1][p], uc2 = d.s[S_UC2][p];
    const double nt0 = d.s[S_NUT][p], nt1 = d.s[S_NUT][p + 1], nt2 = d.s[S_NUT][p + d.sj], nt3 = d.s[S_NUT][p + d.sk];
    mbar_wait(&barA[(k + 2 - kfa) % M::A_ST], ((k + 2 - kfa) / M::A_ST) & 1);
    mbar_wait(&barB[(k + 1 - kfb) % M::B_ST], ((k + 1 - kfb) / M::B_ST) & 1);
    if (active) {
      const double *ta = sm + (ty + 1) * M::AX + (tx + 1), *tb = sm + M::OFF_B + ty * M::BX + tx + 1;
      const int na = k - kfa, nb = k - kfb;
      if (VFS_WARP_ANY(nearv != 0)) {
        FluxAcc2<false> A = {{ta + ((na - 1) % M::A_ST) * M::A_PLANE, ta + (na % M::A_ST) * M::A_PLANE, ta + ((na + 1) % M::A_ST) * M::A_PLANE, ta + ((na + 2) % M::A_ST) * M::A_PLANE},
                             {tb + (nb % M::B_ST) * M::B_PLANE, tb + ((nb + 1) % M::B_ST) * M::B_PLANE}, d, p, {uc0, uc1, uc2}, {nt0, nt1, nt2, nt3}};
        flux_node<false>(d, A, i, j, k + d.kofs, p);
      } else {
        FluxAcc2<true> A = {{ta + ((na - 1) % M::A_ST) * M::A_PLANE, ta + (na % M::A_ST) * M::A_PLANE, ta + ((na + 1) % M::A_ST) * M::A_PLANE, ta + ((na + 2) % M::A_ST) * M::A_PLANE},
                            {tb + (nb % M::B_ST) * M::B_PLANE, tb + ((nb + 1) % M::B_ST) * M::B_PLANE}, d, p, {uc0, uc1, uc2}, {nt0, nt1, nt2, nt3}};
        flux_node<true>(d, A, i, j, k + d.kofs, p);
      }
    }
    __syncthreads();                         // ucat plane k-1 and metric plane k are no longer needed by anyone
    if (tid == 0) {
      fence_proxy_async();
      const int kna = k - 1 + M::A_ST, knb = k + M::B_ST;
      if (kna <= kla) issueA(kna);
      if (knb <= klb) issueB(knb);
    }
  }
}

static inline SidList sids(int n, const int *v) { SidList s; s.n = n; for (int q = 0; q < n; q++) s.sid[q] = v[q]; return s; }
static inline int launch_les1_tma(cudaStream_t st, const CUtensorMap &tmap, const VfsDev &d, int k0, int k1, long *L) {
  const int v[5] = {S_U0, S_U1, S_U2, S_IAJ, S_NV};
  return launch_tile_march<RingLes1>(st, tmap, d, k0, k1, 64, sids(5, v), Les1Body(), L);
}
static inline int launch_flux_tma(cudaStream_t st, const CUtensorMap &tmap, const VfsDev &d, int k0, int k1, int minb, long *L) {
  const int v[4] = {S_U0, S_U1, S_U2, S_NV};
  if (minb == 3) return launch_tile_march<RingFlux, FluxBody, 3>(st, tmap, d, k0, k1, 64, sids(4, v), FluxBody(), L);
  if (minb == 4) return launch_tile_march<RingFlux, FluxBody, 4>(st, tmap, d, k0, k1, 64, sids(4, v), FluxBody(), L);
  return launch_tile_march<RingFlux, FluxBody, 2>(st, tmap, d, k0, k1, 64, sids(4, v), FluxBody(), L);
}
#define VFS_FLUX_HX 4
#define VFS_FLUX_HY 3
#endif  // !VFS_EMU

#endif
