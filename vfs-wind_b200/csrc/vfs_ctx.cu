// vfs_ctx.cu — context, kernel sequencing and the C ABI of include/vfs_b200.h.
//
// Built by nvcc for sm_100a into libvfs_b200.so (the product).  The same translation unit can be
// compiled by g++ with -DVFS_EMU into a host-loop emulation used ONLY by tests/emu (kernel-logic
// debugging in a GPU-less container); the Python binding never loads that build.
#include "../../include/vfs_b200.h"
#include "vfs_common.h"
#include "vfs_halo_kernels.h"
#include "vfs_metrics_kernels.h"
#include "vfs_c2c_kernels.h"
#include "vfs_rhs_kernels.h"
#include "vfs_les_kernels.h"
#include "vfs_wm_kernels.h"
#include "vfs_fused_kernels.h"
#include "vfs_march_kernels.h"
#include "vfs_actuator_kernels.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#ifndef VFS_EMU
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>      // types and prototypes only: the library is resolved at run time (dlopen), no link dependency
// NCCL entry points used by the k-halo layer.  Resolved from the libnccl.so.2 already loaded in the
// process (torch's bundled copy under Python) or the system one.
struct NcclApi {
  void *h = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
};
static NcclApi &nccl_api() {
  static NcclApi a;
  if (a.h) return a;
  a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a.h) a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!a.h) return a;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
  a.Send = (decltype(a.Send))dlsym(a.h, "ncclSend");
  a.Recv = (decltype(a.Recv))dlsym(a.h, "ncclRecv");
  a.AllReduce = (decltype(a.AllReduce))dlsym(a.h, "ncclAllReduce");
  a.GroupStart = (decltype(a.GroupStart))dlsym(a.h, "ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.h, "ncclGroupEnd");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.GroupStart && a.GroupEnd && a.GetErrorString;
  return a;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(c, std::string(#call) + ": " + cudaGetErrorString(e_)); return VFS_ERR_CUDA; } } while (0)
template <class F> __global__ void __launch_bounds__(256) k_box(F f, Box b) {
  int i = b.i0 + blockIdx.x * blockDim.x + threadIdx.x;
  int j = b.j0 + blockIdx.y * blockDim.y + threadIdx.y;
  int k = b.k0 + blockIdx.z * blockDim.z + threadIdx.z;
  if (i < b.i1 && j < b.j1 && k < b.k1) f(i, j, k);
}
// the same kernel compiled for MINB resident 256-thread blocks per SM (a register cap of 65536 / (256 MINB)): the
// bandwidth-bound functors gain from more loads in flight as long as the cap costs no more than a few spilled words
template <class F, int MINB> __global__ void __launch_bounds__(256, MINB) k_box_occ(F f, Box b) {
  int i = b.i0 + blockIdx.x * blockDim.x + threadIdx.x;
  int j = b.j0 + blockIdx.y * blockDim.y + threadIdx.y;
  int k = b.k0 + blockIdx.z * blockDim.z + threadIdx.z;
  if (i < b.i1 && j < b.j1 && k < b.k1) f(i, j, k);
}
#else
#define CK(call) do { } while (0)
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
#endif

static std::string g_create_err;
struct VfsSolver;

struct vfs_ctx {
  vfs_params prm;
  VfsDev d;
  double *pool = nullptr;        // scalars 0 .. S_TAIL0-1, contiguous
#define VFS_NTAIL 4
  double *tail[VFS_NTAIL] = {nullptr, nullptr, nullptr, nullptr};        // scalars S_TAIL0 .. S_COUNT-1 in four groups, each allocated on first use (ensure_tail)
  double *stage = nullptr;       // AoS staging (device), 3 * nzl*my*mx doubles
  double *stage_x = nullptr;     // staging of vfs_formfunction_snes' X, filled on the upload stream (allocated on first use)
  double *stage_async[2] = {nullptr, nullptr};   // staging of vfs_download_async, allocated on first use (sized for the field)
  size_t stage_async_bytes[2] = {0, 0};
  long scalar_len = 0;
  cudaStream_t stream = 0;
  bool own_stream = false;
  vfs_halo_fn halo_fn = nullptr; void *halo_user = nullptr;
  long launches = 0;
  std::string err;
  cudaEvent_t ev[2 * VFS_T_COUNT] = {0};
  bool ev_valid[VFS_T_COUNT] = {false};
  int fused = 1;                 // use the TMA-staged tiled kernels when applicable
#ifndef VFS_EMU
  CUtensorMap tmap;              // 4-D map over the scalar pool, box (TX+2, TY+2, 1, 1)
  CUtensorMap tmap_flux;         // same pool, box (TX+4, TY+3, 1, 1)
  CUtensorMap tmap_rhs;          // same pool, box of the residual marching kernel
  CUtensorMap tmap_les2;         // same pool, box (32, 16): operand tiles of the LES pass-2 marching kernel
  CUtensorMap tmap_les2_8;       // same pool, box (32, 8)
  CUtensorMap tmap_les2_12;      // same pool, box (32, 12)
  CUtensorMap tmap_les2i, tmap_les2i_8, tmap_les2i_12;   // inner-row boxes (32, TY - 2) of the LES pass-2 tiles
  CUtensorMap tmap_fluxA, tmap_fluxB;   // boxes of k_flux_march: ucat (36, 19), metrics (34, 17)
#endif
  bool tma_ok = false;
#ifndef VFS_EMU
  ncclComm_t comm = nullptr;     // k-neighbour halo exchange inside the library (vfs_nccl_init)
  double *hbuf = nullptr;        // packed send (hi, lo) and receive (lo, hi) staging, VFS_MAXGRP scalars each
  cudaStream_t copy_stream = 0;  // vfs_download_async
  cudaEvent_t ev_pack = 0;
  cudaStream_t up_stream = 0;    // host->device copy of X in vfs_formfunction_snes (overlaps kernels still queued on `stream`)
  cudaEvent_t ev_up = 0;
  cudaStream_t side = 0;         // halo exchanges that overlap interior compute run here (forked / joined by events)
  cudaEvent_t ev_fork = 0, ev_join = 0;
  cudaStream_t side2 = 0;        // vfs_rhs_les_fused: the residual's Contra2Cart + IB_BC run here beside LES pass 3 / nu_t (option 18)
  cudaEvent_t ev_fork2 = 0, ev_join2 = 0;
#endif
  int nut_in_les3 = 1;           // option 21: inside vfs_rhs_les_fused LES pass 3 also writes nu_t (no separate NuT pass over the interior)
  int box_occ = 6;               // option 20: the projection kernel compiled for 6 (default) or 8 resident blocks per SM, 0 = the plain k_box (profiles/r02zc_tune_box_occ.txt)
  int les3_minb = 4;             // option 19: resident 512-thread blocks per SM the LES pass-3 kernel is compiled for (4: 32 registers, 64 warps/SM; 3: 40; 2: 48 registers) — profiles/r02zb_tune_les3_minb.txt
  int unit_overlap = 0;          // option 18 (measured: 6.94 -> 6.92 ms at 256^3, profiles/r02y_tune_unit_overlap.txt; off)
  bool fork_after_les2 = false;  // les_cs records ev_fork2 right after LES pass 2 (the last reader of ucat in the LES block)
  int async_api = 0;             // compute-only entry points return without synchronising (option key 11)
  int overlap = 1;               // overlap the k-face-flux and Fp exchanges with the interior planes of FpCell / Project (option key 9)
  long halo_exchanges = 0, halo_bytes = 0;
  // CUDA graphs of the launch-bound call sequences (single rank only: the halo callback is host code)
  int use_graph = 0; bool capturing = false;
  int graph_calls[3] = {0, 0, 0};
#ifndef VFS_EMU
  cudaGraphExec_t gexec[3] = {0, 0, 0};      // 0: residual, 1: RHS+LES unit, 2: the solver's residual (vfs_solver.h)
#endif
  double *homo_buf = nullptr; size_t homo_doubles = 0;      // line / plane sums of the homogeneous-direction Cs averaging
  void *act_buf = nullptr; size_t act_bytes = 0;      // device copy of the actuator element arrays (vfs_calc_f_eul / vfs_calc_u_lagr)
  VfsSolver *solver = nullptr;   // Krylov vectors of vfs_momentum_solve, allocated on first use
  bool iaj_valid = false;        // S_IAJ = 1/aj is current
  bool sabs_valid = false;       // S_SABS holds |S| of the current ucat (set by les_cs pass 1)
  int flux_minb = 2;             // resident blocks per SM requested for the tiled flux kernel (option key 3)
  int les2_ty = 12;              // tile height of the LES pass-2 block program: 16 (1 block/SM) or 8 (2 blocks/SM) (option key 2)
  int les1_var = 1;              // LES pass 1: 0 = block program 32x16, 1 = TMA tile march, 2 = block program 32x8 x2/SM, 3 = 32x16 x2/SM (option key 4)
  int flux_var = 1;              // regular face fluxes: 1 = k_flux_march (metric planes staged too), 0 = k_tile_march<RingFlux> (option key 7)
  double *wm_table = nullptr;    // Cabot wall law: table of int dy+/(1 + nu_t/nu), built on first use
  bool has_solid = true;         // some node has (int)(nvert + 0.1) == 3 (set when nvert is uploaded; true = unknown)
  bool lesgeo_valid = false;     // S_LFINV..S_LF2 match the current metrics and nvert mask
  bool les_bnd_valid = false;    // the static boundary-node values of the LES intermediates (w, |S|S_ij = 0) are stored
  unsigned char *near = nullptr; // near-solid byte mask (VfsDev::near), one byte per padded node
  bool near_valid = false;
  bool wall_marked = false;      // IB_BC's first-step nvert = 1 marking of wall-function first cells has been applied (momentum.c:2048-2074)
  int fuse_refresh = 1;          // single rank: ghost refresh sequences as one launch (RefreshFused) (option key 8)
  int fastpath = 1;              // mask-free specialisations for warps far from any nvert != 0 (option key 6)
  int fp_pairs = 0;              // FpCell on pairs of cells with 16-byte loads (option key 17): measured SLOWER, 0.82 vs 0.70 ms (profiles/r02u_tune_fp_pairs.txt)
  int les_replay = 1;            // between ranks: LES pass 1 replayed on the ghost planes instead of exchanging its 13 fields (option key 16)
  int box_shape = 0;             // thread-block shape of the one-thread-per-node kernels (option key 15, tuning only)
  int halo_trim = 1;             // exchange only the ghost layers each refresh is read at (option key 14); 0: always G layers
  int cur_lo = VFS_G, cur_hi = VFS_G;   // layers of the exchange in progress (vfs_halo_layers)
  int *d_flag = nullptr;         // device scratch flag (has_solid scan)
  int fp_fused = 0;              // Fp evaluated inside the projection kernel (ProjFpMarch) instead of FpCell + Fp planes in HBM (option key 12;
                                 // bitwise the staged result, measured slower on B200: 2.20 vs 0.72 + 0.76 ms at 256^3, gpurun_out r02a)
};

static void graph_reset(vfs_ctx *c);
static void ks_free(vfs_ctx *c);
static void set_err(vfs_ctx *c, const std::string &m) { if (c) c->err = m; else g_create_err = m; }

template <class F> static int launch(vfs_ctx *c, const Box &b, const F &f) {
  if (b.i1 <= b.i0 || b.j1 <= b.j0 || b.k1 <= b.k0) return 0;
  c->launches++;
#ifndef VFS_EMU
  dim3 blk(128, 2, 1);      // measured best of nine 256-thread shapes for the bandwidth kernels (profiles/r02l_tune_box_shape.txt)
  if (c->box_shape) { static const int S[9][3] = {{64, 2, 2}, {32, 4, 2}, {32, 2, 4}, {32, 8, 1}, {64, 4, 1}, {128, 2, 1}, {64, 1, 4}, {128, 1, 2}, {256, 1, 1}};
    const int q = c->box_shape % 9; blk = dim3(S[q][0], S[q][1], S[q][2]); }
  if (b.i1 - b.i0 == 1) blk = dim3(1, 32, 8);       // a single i plane: no idle lanes (rows are strided either way)
  else if (b.i1 - b.i0 <= 8) blk = dim3(8, 8, 4);
  else if (b.j1 - b.j0 == 1) blk = dim3(64, 1, 4);
  else if (b.k1 - b.k0 == 1) blk = dim3(64, 4, 1);
  dim3 grd((b.i1 - b.i0 + blk.x - 1) / blk.x, (b.j1 - b.j0 + blk.y - 1) / blk.y, (b.k1 - b.k0 + blk.z - 1) / blk.z);
  k_box<F><<<grd, blk, 0, c->stream>>>(f, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(c, std::string("kernel launch: ") + cudaGetErrorString(e)); return VFS_ERR_CUDA; }
#else
  for (int k = b.k0; k < b.k1; k++) for (int j = b.j0; j < b.j1; j++) for (int i = b.i0; i < b.i1; i++) f(i, j, k);
#endif
  return 0;
}
// launch() with the occupancy the kernel is compiled for picked by option 20 (0 = the plain k_box).  Measured at 256^3
// (profiles/r02zc_tune_box_occ.txt): the projection gains from 6 blocks per SM (0.735 -> 0.673 ms: 44 -> 40 registers, 16
// bytes of spills), FpCell loses (0.696 -> 0.749 ms), Contra2Cart's interior is at 40 registers anyway; 8 blocks per SM
// (32 registers) loses everywhere.  Used for the projection only.
template <class F> static int launch_occ(vfs_ctx *c, const Box &b, const F &f) {
#ifndef VFS_EMU
  if (c->box_occ && b.i1 - b.i0 > 8 && b.j1 - b.j0 > 1 && b.k1 - b.k0 > 1) {
    c->launches++;
    dim3 blk(128, 2, 1);
    dim3 grd((b.i1 - b.i0 + blk.x - 1) / blk.x, (b.j1 - b.j0 + blk.y - 1) / blk.y, (b.k1 - b.k0 + blk.z - 1) / blk.z);
    if (c->box_occ >= 8) k_box_occ<F, 8><<<grd, blk, 0, c->stream>>>(f, b);
    else k_box_occ<F, 6><<<grd, blk, 0, c->stream>>>(f, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_err(c, std::string("kernel launch: ") + cudaGetErrorString(e)); return VFS_ERR_CUDA; }
    return 0;
  }
#endif
  return launch(c, b, f);
}
#define RUN(x) do { int r_ = (x); if (r_) return r_; } while (0)

// All domain-boundary nodes of the slab (i, j or global k equal to 0 or m-1) in ONE launch, as three
// disjoint index ranges: the two i planes (j, k full), the two j planes (i inner) and the owned global
// k = 0 / mz-1 planes (i, j inner).  [ka, kb) is the local k range of the i- and j-plane parts (it may
// reach into the k ghost planes, see node_copy).  Replaces 4-6 thin launches per boundary operation.
struct Shell { int mx, my, ka, nk, kp[4], nkp; long nA, nB, nC; int iB, wB, iC, wC, jC, hC; };      // i range of the j planes; i, j ranges of the k planes
#define SHELL_PERIODIC_ONLY 8       // make_shell face mask bit: only the face sets of periodic directions
// images: also visit the ghost planes that image the global k = 0 / mz-1 planes across the periodic seam
// (local plane -1 on the first rank, nzl on the last), for the operations contra2cart replays on ghost planes
static Shell make_shell(const VfsDev &d, int ka, int kb, bool images = false, int faces = 7) {
  Shell s; s.mx = d.mx; s.my = d.my; s.ka = ka; s.nk = kb - ka; s.nkp = 0;
  if (d.kofs == 0) s.kp[s.nkp++] = 0;
  if (d.kofs + d.nzl == d.mz) s.kp[s.nkp++] = d.nzl - 1;
  if (images && d.perz && !d.single_rank) {
    if (d.kofs == 0 && ka <= -1) s.kp[s.nkp++] = -1;
    if (d.kofs + d.nzl == d.mz && kb >= d.nzl + 1) s.kp[s.nkp++] = d.nzl;
  }
  if (faces & SHELL_PERIODIC_ONLY) faces = (d.perx ? 1 : 0) | (d.pery ? 2 : 0) | (d.perz ? 4 : 0);
  // a face set that is left out hands its edge nodes to the sets after it (they widen to the full extent)
  s.iB = (faces & 1) ? 1 : 0; s.wB = (faces & 1) ? d.mx - 2 : d.mx; s.iC = s.iB; s.wC = s.wB;
  s.jC = (faces & 2) ? 1 : 0; s.hC = (faces & 2) ? d.my - 2 : d.my;
  s.nA = (faces & 1) ? 2L * d.my * s.nk : 0; s.nB = (faces & 2) ? 2L * s.wB * s.nk : 0; s.nC = (faces & 4) ? (long)s.nkp * s.wC * s.hC : 0;
  return s;
}
// (32-bit index arithmetic: 64-bit divisions cost more than the copies these kernels do; a shell never has 2^31 nodes)
template <class F> VFS_HD void shell_visit(const F &f, const Shell &s, long tl) {
  unsigned t = (unsigned)tl;
  const unsigned nA = (unsigned)s.nA, nB = (unsigned)s.nB;
  if (t < nA) {
    const unsigned per = (unsigned)s.my * (unsigned)s.nk; const unsigned side = t >= per ? 1u : 0u; const unsigned r = t - side * per;
    const unsigned q = r / (unsigned)s.my;
    f(side ? s.mx - 1 : 0, (int)(r - q * (unsigned)s.my), s.ka + (int)q);
  } else if (t < nA + nB) {
    t -= nA;
    const unsigned w = (unsigned)s.wB, per = w * (unsigned)s.nk; const unsigned side = t >= per ? 1u : 0u; const unsigned r = t - side * per;
    const unsigned q = r / w;
    f(s.iB + (int)(r - q * w), side ? s.my - 1 : 0, s.ka + (int)q);
  } else {
    t -= nA + nB;
    const unsigned w = (unsigned)s.wC, per = w * (unsigned)s.hC; const unsigned qq = t / per; const unsigned r = t - qq * per;
    const unsigned q = r / w;
    f(s.iC + (int)(r - q * w), s.jC + (int)q, s.kp[qq]);
  }
}
#ifndef VFS_EMU
template <class F> __global__ void __launch_bounds__(256) k_shell(F f, Shell s) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < s.nA + s.nB + s.nC) shell_visit(f, s, t);
}
#endif
template <class F> static int launch_shell(vfs_ctx *c, int ka, int kb, const F &f, bool images = false, int faces = 7) {
  const Shell s = make_shell(c->d, ka, kb, images, faces);
  const long n = s.nA + s.nB + s.nC;
  if (n <= 0) return 0;
  c->launches++;
#ifndef VFS_EMU
  k_shell<F><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(f, s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(c, std::string("kernel launch: ") + cudaGetErrorString(e)); return VFS_ERR_CUDA; }
#else
  for (long t = 0; t < n; t++) shell_visit(f, s, t);
#endif
  return 0;
}

static void ev_rec(vfs_ctx *c, int n) {
#ifndef VFS_EMU
  if (c->capturing) return;       // events recorded inside a graph cannot be used for timing
  if (c->ev[n]) cudaEventRecord(c->ev[n], c->stream);
  c->ev_valid[n / 2] = true;
#endif
}

// ---- public field table -----------------------------------------------------------------------
static const struct { int s0, dof; } FIELD[VFS_NFIELDS_PUBLIC] = {
  {S_X, 3}, {S_CSI0, 3}, {S_ETA0, 3}, {S_ZET0, 3}, {S_AJ, 1}, {S_NV, 1}, {S_UC0, 3}, {S_U0, 3}, {S_UO0, 3},
  {S_UCO0, 3}, {S_UCM0, 3}, {S_RO0, 3}, {S_DP0, 3}, {S_FE0, 3}, {S_R0, 3}, {S_CS, 1}, {S_NUT, 1}, {S_USTAR, 1}, {S_CONV0, 3}, {S_VISC0, 3}, {S_P, 1}, {S_PHI, 1}};

static Grp grp(int s0, int n) { Grp g; g.n = n; for (int q = 0; q < n; q++) g.sid[q] = s0 + q; return g; }
static Grp grp_cat(const Grp &a, const Grp &b) { Grp g = a; for (int q = 0; q < b.n; q++) g.sid[g.n++] = b.sid[q]; return g; }

// owned-node boxes (local k)
static Box box_owned(const vfs_ctx *c) { Box b = {0, c->d.mx, 0, c->d.my, 0, c->d.nzl}; return b; }
static int klo(const vfs_ctx *c, int kg) { int k = kg - c->d.kofs; return k < 0 ? 0 : (k > c->d.nzl ? c->d.nzl : k); }
static Box box_interior(const vfs_ctx *c) { Box b = {1, c->d.mx - 1, 1, c->d.my - 1, klo(c, 1), klo(c, c->d.mz - 1)}; return b; }

// ---- ghost refresh primitives -------------------------------------------------------------------
static bool has_wallfn(const VfsDev &d) { for (int q = 0; q < 4; q++) if (d.bc[q] == -1 || d.bc[q] == -2) return true; return false; }
static bool any_per_d(const VfsDev &d) { return d.perx || d.pery || d.perz; }
static int wrap_ij(vfs_ctx *c, const Grp &g, int ka = 0, int kb = -1) {
  const VfsDev &d = c->d;
  if (kb < 0) kb = d.nzl;
  if (d.perx) { WrapFill f = {d, g, 0}; Box b = {0, 2 * VFS_G, 0, d.my, ka, kb}; RUN(launch(c, b, f)); }
  if (d.pery) { WrapFill f = {d, g, 1}; Box b = {-VFS_G, d.mx + VFS_G, 0, 2 * VFS_G, ka, kb}; RUN(launch(c, b, f)); }
  return 0;
}
#ifndef VFS_EMU
// k-halo exchange with the neighbouring slabs.  The G boundary planes of every requested scalar are
// packed into ONE message per neighbour (a few 2 MB messages per peer only reach ~70 GB/s over NVLink,
// one 10-50 MB message is spread over all of NCCL's P2P channels), exchanged with one grouped
// ncclSend/ncclRecv pair per neighbour on the context's stream (stream ordered, graph capturable) and
// unpacked into the ghost planes.  With lo == hi (2 ranks, periodic k) the first send to the peer pairs
// with the peer's first receive: [send hi, send lo, recv lo, recv hi] on both sides is consistent.
struct HaloPtrs { double *s[VFS_MAXGRP]; };
// one message: n scalars x cnt doubles, scalar q's block taken from / stored to P.s[q] + off
__global__ void k_halo_copy(HaloPtrs P, int n, long cntA, long offA, double2 *__restrict__ bufA, long cntB, long offB, double2 *__restrict__ bufB, int unpack) {
  const long a2 = bufA ? cntA / 2 : 0, b2 = bufB ? cntB / 2 : 0, totA = (long)n * a2, tot = totA + (long)n * b2;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
    const bool isA = t < totA;
    const long u = isA ? t : t - totA, c2 = isA ? a2 : b2;
    const int q = (int)(u / c2); const long e = u - (long)q * c2;
    double2 *g = reinterpret_cast<double2 *>(P.s[q] + (isA ? offA : offB)) + e;
    double2 *b = (isA ? bufA : bufB) + u;
    if (unpack) *g = *b; else *b = *g;
  }
}
// nlo / nhi: ghost planes to fill below / above the slab (every rank uses the same pair for a given exchange): the
// hi neighbour receives my top nlo planes into its low ghosts, the lo neighbour my bottom nhi planes into its high ghosts
static int nccl_halo(vfs_ctx *c, const Grp &g, bool seam_only, int nlo, int nhi) {
  NcclApi &N = nccl_api();
  const VfsDev &d = c->d;
  const int r = c->prm.rank, n = c->prm.nranks;
  int lo = r > 0 ? r - 1 : (d.perz ? n - 1 : -1), hi = r < n - 1 ? r + 1 : (d.perz ? 0 : -1);
  if (seam_only) {            // only the periodic seam rank 0 <-> rank n-1 (see g2l_after_copy)
    if (r > 0) lo = -1;
    if (r < n - 1) hi = -1;
    if (lo < 0 && hi < 0) return 0;
  }
  const long cmax = (long)VFS_G * d.sk;                      // doubles per scalar and side at full width (sk is a multiple of 16)
  if (!c->hbuf) {
    CK(cudaMalloc((void **)&c->hbuf, (size_t)4 * VFS_MAXGRP * cmax * sizeof(double)));
  }
  double *sb_hi = c->hbuf, *sb_lo = c->hbuf + (size_t)VFS_MAXGRP * cmax, *rb_lo = c->hbuf + (size_t)2 * VFS_MAXGRP * cmax, *rb_hi = c->hbuf + (size_t)3 * VFS_MAXGRP * cmax;
  HaloPtrs P; for (int q = 0; q < g.n; q++) P.s[q] = c->d.s[g.sid[q]];
  const long cnt_lo = (long)nlo * d.sk, cnt_hi = (long)nhi * d.sk;
  const size_t msg_up = (size_t)g.n * cnt_lo, msg_dn = (size_t)g.n * cnt_hi;      // to the hi neighbour / to the lo neighbour
  const int blocks = 148 * 4;
  // pack: top nlo owned planes -> sb_hi, bottom nhi owned planes -> sb_lo
  k_halo_copy<<<blocks, 256, 0, c->stream>>>(P, g.n, cnt_lo, (long)(VFS_G + d.nzl - nlo) * d.sk, hi >= 0 ? (double2 *)sb_hi : nullptr,
                                             cnt_hi, (long)VFS_G * d.sk, lo >= 0 ? (double2 *)sb_lo : nullptr, 0);
  ncclResult_t e = N.GroupStart();
  if (hi >= 0 && e == ncclSuccess) e = N.Send(sb_hi, msg_up, ncclDouble, hi, c->comm, c->stream);
  if (lo >= 0 && e == ncclSuccess) e = N.Send(sb_lo, msg_dn, ncclDouble, lo, c->comm, c->stream);
  if (lo >= 0 && e == ncclSuccess) e = N.Recv(rb_lo, msg_up, ncclDouble, lo, c->comm, c->stream);
  if (hi >= 0 && e == ncclSuccess) e = N.Recv(rb_hi, msg_dn, ncclDouble, hi, c->comm, c->stream);
  ncclResult_t e2 = N.GroupEnd();
  if (e == ncclSuccess) e = e2;
  if (e != ncclSuccess) { set_err(c, std::string("NCCL halo exchange: ") + N.GetErrorString(e)); return VFS_ERR_HALO; }
  // unpack: rb_lo -> ghost planes [-nlo, 0), rb_hi -> ghost planes [nzl, nzl + nhi)
  k_halo_copy<<<blocks, 256, 0, c->stream>>>(P, g.n, cnt_lo, (long)(VFS_G - nlo) * d.sk, lo >= 0 ? (double2 *)rb_lo : nullptr,
                                             cnt_hi, (long)(VFS_G + d.nzl) * d.sk, hi >= 0 ? (double2 *)rb_hi : nullptr, 1);
  CK(cudaGetLastError());
  c->halo_exchanges++; c->halo_bytes += (long)8 * ((hi >= 0 ? msg_up : 0) + (lo >= 0 ? msg_dn : 0));
  c->launches += 3;
  return 0;
}
#endif
// lo / hi: how many ghost planes below / above the slab the refreshed field is read at before its next refresh
// (1-2 layers for most fields); G when trimming is off.  A host callback learns them from vfs_halo_layers.
static int halo_k(vfs_ctx *c, const Grp &g, bool seam_only = false, int lo = VFS_G, int hi = VFS_G) {
  const VfsDev &d = c->d;
  if (!c->halo_trim) lo = hi = VFS_G;
  if (c->prm.nranks > 1) {
#ifndef VFS_EMU
    if (c->comm) return nccl_halo(c, g, seam_only, lo, hi);
#endif
    if (!c->halo_fn) { set_err(c, "nranks > 1 but neither vfs_nccl_init nor a halo callback was set up"); return VFS_ERR_HALO; }
    c->cur_lo = lo; c->cur_hi = hi;
    int r = c->halo_fn(c->halo_user, g.n, g.sid);
    c->cur_lo = c->cur_hi = VFS_G;
    if (r) { set_err(c, "halo callback failed"); return VFS_ERR_HALO; }
    c->halo_exchanges++;
    return 0;
  }
  if (d.perz) { WrapFill f = {d, g, 2}; Box b = {-VFS_G, d.mx + VFS_G, -VFS_G, d.my + VFS_G, 0, 2 * VFS_G}; RUN(launch(c, b, f)); }
  return 0;
}
extern "C" int vfs_halo_layers(vfs_ctx *c, int *lo, int *hi) { if (!c || !lo || !hi) return VFS_ERR_ARG; *lo = c->cur_lo; *hi = c->cur_hi; return 0; }
#ifndef VFS_EMU
template <class F> __global__ void __launch_bounds__(256) k_linear(F f, long n) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) f(t);
}
#endif
// single rank: the whole refresh (mode 1: g2l, mode 3: g2l + node_copy + g2l) as one launch, see RefreshFused.
// depth: ghost layers refreshed in every periodic direction (how deep the field is read before its next refresh)
static int refresh_fused(vfs_ctx *c, const Grp &g, int mode, int depth = VFS_G) {
  const VfsDev &d = c->d;
  if (!any_per_d(d)) return 0;
  if (!c->halo_trim || depth > VFS_G) depth = VFS_G;
  RefreshSlabs S; S.depth = depth;
  S.ext[0] = d.perx ? d.mx + 2 * depth : d.mx; S.ext[1] = d.pery ? d.my + 2 * depth : d.my; S.ext[2] = d.perz ? d.mz + 2 * depth : d.mz;
  const long W = 2 * depth + 2;
  S.n[0] = d.perx ? W * S.ext[1] * S.ext[2] : 0; S.n[1] = d.pery ? W * S.ext[0] * S.ext[2] : 0; S.n[2] = d.perz ? W * S.ext[0] * S.ext[1] : 0;
  const long n = S.n[0] + S.n[1] + S.n[2];
  RefreshFused f = {d, g, mode, S};
  c->launches++;
#ifndef VFS_EMU
  k_linear<RefreshFused><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(f, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(c, std::string("kernel launch: ") + cudaGetErrorString(e)); return VFS_ERR_CUDA; }
#else
  for (long t = 0; t < n; t++) f(t);
#endif
  return 0;
}
// A k-halo exchange that overlaps compute: ovl_exchange() forks the side stream off the main stream and runs
// the packed NCCL exchange there; the caller queues the work that does not read the ghost planes on the
// main stream and then calls ovl_join(), after which the main stream sees the ghosts.  Event fork/join,
// so the pattern is captured into the step's CUDA graph like everything else.
static bool can_overlap(const vfs_ctx *c) {
#ifndef VFS_EMU
  return c->overlap && c->prm.nranks > 1 && c->comm && c->side;
#else
  (void)c; return false;
#endif
}
static int ovl_exchange(vfs_ctx *c, const Grp &g, int lo = VFS_G, int hi = VFS_G) {
#ifndef VFS_EMU
  CK(cudaEventRecord(c->ev_fork, c->stream));
  CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
  cudaStream_t main_stream = c->stream;
  c->stream = c->side;
  if (!c->halo_trim) lo = hi = VFS_G;
  const int r = nccl_halo(c, g, false, lo, hi);
  c->stream = main_stream;
  return r;
#else
  (void)c; (void)g; (void)lo; (void)hi; return VFS_ERR_UNSUPPORTED;
#endif
}
// Thin-slab launches that are independent of a big kernel queued right before them run on the side stream,
// concurrently with it (same event fork / join): side_begin() ... launches ... side_end(); the caller joins with
// ovl_join() once the big kernel has been queued on the main stream.
struct SideScope { cudaStream_t main_stream; };
static int side_begin(vfs_ctx *c, SideScope *s) {
#ifndef VFS_EMU
  CK(cudaEventRecord(c->ev_fork, c->stream));
  CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
  s->main_stream = c->stream; c->stream = c->side;
#else
  (void)c; (void)s;
#endif
  return 0;
}
static void side_end(vfs_ctx *c, const SideScope *s) {
#ifndef VFS_EMU
  c->stream = s->main_stream;
#else
  (void)c; (void)s;
#endif
}
static int ovl_join(vfs_ctx *c) {
#ifndef VFS_EMU
  CK(cudaEventRecord(c->ev_join, c->side));
  CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
#endif
  (void)c; return 0;
}
// DAGlobalToLocal / DALocalToLocal
static int g2l(vfs_ctx *c, const Grp &g, int lo = VFS_G, int hi = VFS_G) {
  if (c->prm.nranks == 1 && c->fuse_refresh) return refresh_fused(c, g, 1, lo > hi ? lo : hi);
  RUN(wrap_ij(c, g)); return halo_k(c, g, false, lo, hi);
}
// The refresh that follows a node_copy of a field whose ghosts were refreshed just before it: the
// copy has already been applied to the ghost planes of interior slab boundaries (see node_copy), so
// between ranks only the periodic seam (global planes 0 / mz-1 changed by the k copies) must travel.
static int g2l_after_copy(vfs_ctx *c, const Grp &g) { RUN(wrap_ij(c, g)); return c->d.perz ? halo_k(c, g, true) : 0; }
// the "if(periodic) ... f[k][j][i] = f[c][b][a]" loops
static int node_copy(vfs_ctx *c, const Grp &g) {
  const VfsDev &d = c->d;
  NodeCopy f = {d, g, 0};
  // k-ghost planes received from a neighbouring rank across an INTERIOR slab boundary are owned
  // planes in a single-rank run, where this copy updates them; apply it there too so that the
  // result does not depend on the number of ranks (wrap-around ghosts stay stale, as in 1 rank).
  const int ka = d.kofs > 0 ? -VFS_G : 0, kb = d.kofs + d.nzl < d.mz ? d.nzl + VFS_G : d.nzl;
  return launch_shell(c, ka, kb, f, false, SHELL_PERIODIC_ONLY);      // the functor acts on nodes of periodic boundary planes only
}
static bool any_per(const vfs_ctx *c) { return c->d.perx || c->d.pery || c->d.perz; }

// ---- creation -------------------------------------------------------------------------------------
static int check_params(const vfs_params *p, std::string &why) {
  // legacy k_periodic reads the interior plane mz-2 / 1 directly in IB_BC (IbBcBoundary): another rank's plane on a slab
  if (p->k_periodic && !p->kk_periodic && p->nranks > 1) { why = "legacy k_periodic is single-rank only (as in the reference): use kk_periodic between ranks"; return VFS_ERR_UNSUPPORTED; }
  if (p->mx < 6 || p->my < 6 || p->mz < 6) { why = "grid too small (need >= 6 nodes per direction)"; return VFS_ERR_ARG; }
  if (p->nranks < 1 || p->rank < 0 || p->rank >= p->nranks) { why = "bad rank/nranks"; return VFS_ERR_ARG; }
  if (p->kofs < 0 || p->nzl < 1 || p->kofs + p->nzl > p->mz) { why = "bad k-slab"; return VFS_ERR_ARG; }
  if (p->nranks == 1 && (p->kofs != 0 || p->nzl != p->mz)) { why = "single rank must own all k planes"; return VFS_ERR_ARG; }
  if (p->nranks > 1 && p->nzl < VFS_G) { why = "k-slab thinner than the ghost width"; return VFS_ERR_ARG; }
  if (p->levelset || p->rans || p->movefsi || p->rotatefsi) { why = "levelset/rans/movefsi/rotatefsi are outside the hot-path scope"; return VFS_ERR_UNSUPPORTED; }
  if ((p->levelset_weno && p->levelset_weno != 5) || p->freesurface_wallmodel || p->air_flow_levelset) { why = "levelset_weno 1-4 / freesurface_wallmodel / air_flow_levelset key the flux and wall-model code on the level-set field (momentum.c:754,1015,1301) and are not built (levelset_weno = 5, WENO3 everywhere, is)"; return VFS_ERR_UNSUPPORTED; }
  if (p->les < 0 || p->les > 2) { why = "les must be 0, 1 or 2"; return VFS_ERR_UNSUPPORTED; }
  for (int q = 4; q < 6; q++) if (p->bctype[q] == -1 || p->bctype[q] == -2) { why = "wall-function boundary types (-1,-2) on a k side: Contra2Cart_2 has no velocity rule for them (rhs.c:311-440)"; return VFS_ERR_UNSUPPORTED; }
  for (int q = 0; q < 4; q++) if (p->bctype[q] == -2 && !(p->roughness_size > 0)) { why = "bctype -2 (rough-wall log law) needs roughness_size > 0"; return VFS_ERR_ARG; }
  if (!(p->ren > 0) || !(p->dt > 0)) { why = "ren and dt must be positive"; return VFS_ERR_ARG; }
  return 0;
}

static void fill_dev(vfs_ctx *c) {
  const vfs_params &p = c->prm; VfsDev &d = c->d;
  d.mx = p.mx; d.my = p.my; d.mz = p.mz; d.nzl = p.nzl; d.kofs = p.kofs;
  d.pitch = ((p.mx + 2 * VFS_G + 15) / 16) * 16; d.ny = p.my + 2 * VFS_G; d.nzt = p.nzl + 2 * VFS_G;
  d.sj = d.pitch; d.sk = (long)d.ny * d.pitch;
  d.org = (long)VFS_G * d.sk + (long)VFS_G * d.sj + VFS_G;
  d.perx = (p.ii_periodic || p.i_periodic) != 0; d.pery = (p.jj_periodic || p.j_periodic) != 0; d.perz = (p.kk_periodic || p.k_periodic) != 0;
  d.legx = p.i_periodic != 0 && !p.ii_periodic; d.legy = p.j_periodic != 0 && !p.jj_periodic; d.legz = p.k_periodic != 0 && !p.kk_periodic;
  for (int q = 0; q < 6; q++) d.bc[q] = p.bctype[q];
  d.les = p.les; d.second_order = p.second_order; d.laplacian = p.laplacian; d.immersed = p.immersed; d.clark = p.clark;
  d.testfilter_ik = p.testfilter_ik; d.visc_wm = p.viscosity_wallmodel; d.wallfunction = p.wallfunction;
  d.has_feul = (p.rotor_model || p.nacelle_model || p.IB_delta) ? 1 : 0;
  d.ti = p.ti; d.tistart = p.tistart; d.rstart_flg = p.rstart_flg; d.bdf2 = 0; d.single_rank = p.nranks == 1;
  d.inviscid = p.inviscid; d.skew = p.skew; d.weno = p.inviscid || p.levelset_weno == 5;
  d.ren = p.ren; d.dt = p.dt; d.max_cs = p.max_cs; d.roughness = p.roughness_size;
  d.homo = (p.i_homo_filter && p.k_homo_filter) ? 1 : (p.i_homo_filter ? 2 : (p.j_homo_filter ? 3 : (p.k_homo_filter ? 4 : 0)));      // les.c:799,840
}

extern "C" int vfs_create(const vfs_params *p, vfs_ctx **out) {
  if (!p || !out) { g_create_err = "null argument"; return VFS_ERR_ARG; }
  std::string why; int r = check_params(p, why);
  if (r) { g_create_err = why; return r; }
  vfs_ctx *c = new vfs_ctx();
  c->prm = *p; fill_dev(c);
  c->scalar_len = (long)c->d.nzt * c->d.sk;
  size_t bytes = (size_t)c->scalar_len * S_TAIL0 * sizeof(double);
  size_t sbytes = (size_t)p->nzl * p->my * p->mx * 3 * sizeof(double);
#ifndef VFS_EMU
  cudaError_t e = cudaSetDevice(p->device);
  if (e != cudaSuccess) { g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e) + " (no CPU fallback exists)"; delete c; return VFS_ERR_CUDA; }
  e = cudaMalloc((void **)&c->pool, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void **)&c->stage, sbytes);
  if (e == cudaSuccess) e = cudaMalloc((void **)&c->near, (size_t)c->scalar_len);
  if (e != cudaSuccess) { g_create_err = std::string("cudaMalloc: ") + cudaGetErrorString(e); delete c; return VFS_ERR_CUDA; }
  cudaMemset(c->pool, 0, bytes);
  cudaMemset(c->near, 1, (size_t)c->scalar_len);
  cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking); c->own_stream = true;
  cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  cudaStreamCreateWithFlags(&c->side2, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->ev_fork2, cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming);
  for (int q = 0; q < 2 * VFS_T_COUNT; q++) cudaEventCreate(&c->ev[q]);
#else
  c->pool = (double *)calloc(bytes, 1); c->stage = (double *)calloc(sbytes, 1);
  c->near = (unsigned char *)malloc((size_t)c->scalar_len); memset(c->near, 1, (size_t)c->scalar_len);
#endif
  for (int s = 0; s < S_COUNT; s++) c->d.s[s] = s < S_TAIL0 ? c->pool + (long)s * c->scalar_len : nullptr;
  c->d.near = c->near;
#ifndef VFS_EMU
  c->tma_ok = vfs_make_tensor_map(&c->tmap, c->pool, c->d, c->scalar_len, VFS_TILE_TX + 2, VFS_TILE_TY + 2) == 0 &&
              vfs_make_tensor_map(&c->tmap_flux, c->pool, c->d, c->scalar_len, VFS_TILE_TX + VFS_FLUX_HX, VFS_TILE_TY + VFS_FLUX_HY) == 0 &&
              vfs_make_tensor_map(&c->tmap_rhs, c->pool, c->d, c->scalar_len, RhsMarch::NXP, RhsMarch::NYP) == 0 &&
              vfs_make_tensor_map(&c->tmap_les2, c->pool, c->d, c->scalar_len, Les2March::TX, Les2March::TY) == 0 &&
              vfs_make_tensor_map(&c->tmap_les2_8, c->pool, c->d, c->scalar_len, Les2March8::TX, Les2March8::TY) == 0 &&
              vfs_make_tensor_map(&c->tmap_les2_12, c->pool, c->d, c->scalar_len, Les2March12::TX, Les2March12::TY) == 0 &&
              vfs_make_tensor_map(&c->tmap_les2i, c->pool, c->d, c->scalar_len, Les2March::TX, Les2March::TY - 2) == 0 &&
              vfs_make_tensor_map(&c->tmap_les2i_8, c->pool, c->d, c->scalar_len, Les2March8::TX, Les2March8::TY - 2) == 0 &&
              vfs_make_tensor_map(&c->tmap_les2i_12, c->pool, c->d, c->scalar_len, Les2March12::TX, Les2March12::TY - 2) == 0 &&
              vfs_make_tensor_map(&c->tmap_fluxA, c->pool, c->d, c->scalar_len, FluxMarch::AX, FluxMarch::AY) == 0 &&
              vfs_make_tensor_map(&c->tmap_fluxB, c->pool, c->d, c->scalar_len, FluxMarch::BX, FluxMarch::BY) == 0;
#endif
  *out = c;
  return 0;
}

extern "C" void *vfs_host_alloc(unsigned long bytes) {
#ifndef VFS_EMU
  void *p = nullptr; return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
#else
  return malloc(bytes);
#endif
}
extern "C" void vfs_host_free(void *p) {
#ifndef VFS_EMU
  if (p) cudaFreeHost(p);
#else
  free(p);
#endif
}
extern "C" int vfs_device_count(void) {
#ifndef VFS_EMU
  int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
#else
  return 1;
#endif
}

extern "C" int vfs_destroy(vfs_ctx *c) {
  if (!c) return VFS_ERR_ARG;
#ifndef VFS_EMU
  cudaStreamSynchronize(c->stream);
#endif
  ks_free(c);
#ifndef VFS_EMU
  graph_reset(c);            // graph execs hold captured NCCL work: ncclCommDestroy blocks for ever while they exist
  if (c->comm) nccl_api().CommDestroy(c->comm);
  if (c->hbuf) cudaFree(c->hbuf);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->up_stream) { cudaStreamSynchronize(c->up_stream); cudaStreamDestroy(c->up_stream); }
  if (c->ev_up) cudaEventDestroy(c->ev_up);
  if (c->stage_x) cudaFree(c->stage_x);
  if (c->ev_pack) cudaEventDestroy(c->ev_pack);
  for (int q = 0; q < 2; q++) if (c->stage_async[q]) cudaFree(c->stage_async[q]);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->side2) cudaStreamDestroy(c->side2);
  if (c->ev_fork2) cudaEventDestroy(c->ev_fork2);
  if (c->ev_join2) cudaEventDestroy(c->ev_join2);
  if (c->wm_table) cudaFree(c->wm_table);
  cudaFree(c->pool); for (int g = 0; g < VFS_NTAIL; g++) if (c->tail[g]) cudaFree(c->tail[g]); cudaFree(c->stage); cudaFree(c->near); if (c->d_flag) cudaFree(c->d_flag); if (c->act_buf) cudaFree(c->act_buf); if (c->homo_buf) cudaFree(c->homo_buf);
  graph_reset(c);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  for (int q = 0; q < 2 * VFS_T_COUNT; q++) if (c->ev[q]) cudaEventDestroy(c->ev[q]);
#else
  free(c->pool); for (int g = 0; g < VFS_NTAIL; g++) free(c->tail[g]); free(c->stage); free(c->wm_table); free(c->near); free(c->d_flag); free(c->act_buf); free(c->homo_buf);
#endif
  delete c; return 0;
}
extern "C" const char *vfs_last_error(vfs_ctx *c) { return c ? c->err.c_str() : g_create_err.c_str(); }
extern "C" int vfs_set_params(vfs_ctx *c, const vfs_params *p) {
  if (!c || !p) return VFS_ERR_ARG;
  if (p->mx != c->prm.mx || p->my != c->prm.my || p->mz != c->prm.mz || p->nzl != c->prm.nzl || p->kofs != c->prm.kofs || p->nranks != c->prm.nranks) { set_err(c, "geometry cannot change"); return VFS_ERR_ARG; }
  std::string why; int r = check_params(p, why); if (r) { set_err(c, why); return r; }
  graph_reset(c);
  c->prm = *p; double *sv[S_COUNT]; memcpy(sv, c->d.s, sizeof(sv)); fill_dev(c); memcpy(c->d.s, sv, sizeof(sv)); c->d.near = c->near; return 0;
}
extern "C" int vfs_set_stream(vfs_ctx *c, void *s) {
  if (!c) return VFS_ERR_ARG;
#ifndef VFS_EMU
  cudaStreamSynchronize(c->stream);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  c->own_stream = false;
#endif
  graph_reset(c);
  c->stream = (cudaStream_t)s; return 0;
}
extern "C" int vfs_set_halo_callback(vfs_ctx *c, vfs_halo_fn fn, void *user) { if (!c) return VFS_ERR_ARG; c->halo_fn = fn; c->halo_user = user; return 0; }
extern "C" int vfs_nccl_unique_id(char *out128) {
#ifndef VFS_EMU
  NcclApi &N = nccl_api();
  if (!out128) return VFS_ERR_ARG;
  if (!N.ok) { g_create_err = "libnccl.so.2 not found"; return VFS_ERR_HALO; }
  ncclUniqueId id;
  if (N.GetUniqueId(&id) != ncclSuccess) { g_create_err = "ncclGetUniqueId failed"; return VFS_ERR_HALO; }
  memcpy(out128, id.internal, NCCL_UNIQUE_ID_BYTES);
  return 0;
#else
  (void)out128; return VFS_ERR_UNSUPPORTED;
#endif
}
extern "C" int vfs_nccl_init(vfs_ctx *c, const char *id128) {
#ifndef VFS_EMU
  if (!c || !id128) return VFS_ERR_ARG;
  NcclApi &N = nccl_api();
  if (!N.ok) { set_err(c, "libnccl.so.2 not found"); return VFS_ERR_HALO; }
  graph_reset(c);
  if (c->comm) { N.CommDestroy(c->comm); c->comm = nullptr; }
  ncclUniqueId id; memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
  CK(cudaSetDevice(c->prm.device));
  ncclResult_t e = N.CommInitRank(&c->comm, c->prm.nranks, id, c->prm.rank);
  if (e != ncclSuccess) { c->comm = nullptr; set_err(c, std::string("ncclCommInitRank: ") + N.GetErrorString(e)); return VFS_ERR_HALO; }
  return 0;
#else
  (void)c; (void)id128; return VFS_ERR_UNSUPPORTED;
#endif
}
extern "C" long vfs_halo_count(vfs_ctx *c, long *bytes) { if (!c) return 0; if (bytes) *bytes = c->halo_bytes; return c->halo_exchanges; }
extern "C" int vfs_sync(vfs_ctx *c) {
  if (!c) return VFS_ERR_ARG;
#ifndef VFS_EMU
  CK(cudaStreamSynchronize(c->stream));
#endif
  return 0;
}
// end of a compute-only entry point: synchronous by default; with option 11 the call returns as soon as its work is
// queued (errors then surface at the next synchronising call), so a following vfs_formfunction_snes can start
// copying X while these kernels still run
static int api_end(vfs_ctx *c) { return c->async_api ? 0 : vfs_sync(c); }
extern "C" int vfs_layout(vfs_ctx *c, long *L) {
  if (!c || !L) return VFS_ERR_ARG;
  L[0] = VFS_G; L[1] = c->d.pitch; L[2] = c->d.ny; L[3] = c->d.nzt; L[4] = c->d.sk; L[5] = c->scalar_len; L[6] = S_TAIL0; L[7] = 1; return 0;      // (pool scalars: the ones a halo layer can be asked for)
}
extern "C" int vfs_field_scalar_id(vfs_ctx *c, int field, int comp) {
  if (!c || field < 0 || field >= VFS_NFIELDS_PUBLIC || comp < 0 || comp >= FIELD[field].dof) return VFS_ERR_ARG;
  return FIELD[field].s0 + comp;
}
extern "C" void *vfs_scalar_ptr(vfs_ctx *c, int sid) { if (!c || sid < 0 || sid >= S_COUNT) return 0; return c->d.s[sid]; }      // (null for a tail scalar nobody has used yet)
extern "C" long vfs_launch_count(vfs_ctx *c) { return c ? c->launches : 0; }
extern "C" double vfs_last_ms(vfs_ctx *c, int which) {
#ifndef VFS_EMU
  if (!c || which < 0 || which >= VFS_T_COUNT || !c->ev_valid[which]) return 0;
  float ms = 0; if (cudaEventElapsedTime(&ms, c->ev[2 * which], c->ev[2 * which + 1]) != cudaSuccess) return 0; return ms;
#else
  return 0;
#endif
}
static void graph_reset(vfs_ctx *c) {
#ifndef VFS_EMU
  for (int q = 0; q < 3; q++) { if (c->gexec[q]) cudaGraphExecDestroy(c->gexec[q]); c->gexec[q] = 0; c->graph_calls[q] = 0; }
#endif
}
extern "C" int vfs_set_option(vfs_ctx *c, int key, int value) {
  if (!c) return VFS_ERR_ARG;
  if (key == 0) c->fused = value;
  else if (key == 1) c->use_graph = value;
  else if (key == 2) c->les2_ty = value;
  else if (key == 3) c->flux_minb = value;
  else if (key == 4) c->les1_var = value;
  else if (key == 6) { c->fastpath = value; c->near_valid = false; }
  else if (key == 7) c->flux_var = value;
  else if (key == 8) c->fuse_refresh = value;
  else if (key == 9) c->overlap = value;
  else if (key == 11) c->async_api = value;
  else if (key == 12) c->fp_fused = value;
  else if (key == 14) c->halo_trim = value;
  else if (key == 15) c->box_shape = value;
  else if (key == 16) c->les_replay = value;
  else if (key == 17) c->fp_pairs = value;
  else if (key == 18) c->unit_overlap = value;
  else if (key == 19) c->les3_minb = value;
  else if (key == 20) c->box_occ = value;
  else if (key == 21) c->nut_in_les3 = value;
  graph_reset(c);
  return 0;
}
// Run `body` (a pure launch sequence on c->stream) eagerly the first time, capture it into a CUDA
// graph the second time and replay the graph afterwards.  Kernel arguments (VfsDev by value) are
// baked into the graph, so every parameter/stream/option change resets it.
template <class F> static int run_graphed(vfs_ctx *c, int key, F body) {
#ifndef VFS_EMU
  if (c->use_graph && (c->prm.nranks == 1 || c->comm)) {
    if (c->gexec[key]) { CK(cudaGraphLaunch(c->gexec[key], c->stream)); return 0; }
    if (c->graph_calls[key]++ >= 1) {
      cudaGraph_t g = 0;
      CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      c->capturing = true;
      int r = body();
      c->capturing = false;
      cudaError_t e = cudaStreamEndCapture(c->stream, &g);
      if (r) { if (g) cudaGraphDestroy(g); return r; }
      if (e != cudaSuccess || !g) { set_err(c, std::string("graph capture: ") + cudaGetErrorString(e)); return VFS_ERR_CUDA; }
      e = cudaGraphInstantiate(&c->gexec[key], g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) { c->gexec[key] = 0; set_err(c, std::string("graph instantiate: ") + cudaGetErrorString(e)); return VFS_ERR_CUDA; }
      CK(cudaGraphLaunch(c->gexec[key], c->stream));
      return 0;
    }
  }
#endif
  return body();
}

// the rarely used scalars (Ucont_rm1, legacy Conv / Visc) live outside the main pool and exist only once somebody uses them
// four groups, each allocated when its first user shows up: legacy (Ucont_rm1, Conv, Visc), Adv1-3 (skew), Phi
// (UpdatePressure / Projection), the clark gradient planes
static const int TAIL_LO[VFS_NTAIL + 1] = {S_TAIL0, S_ADV1, S_PHI, S_GR0, S_COUNT};
static int tail_group(int sid) { int g = 0; while (g + 1 < VFS_NTAIL && sid >= TAIL_LO[g + 1]) g++; return g; }
static int ensure_tail(vfs_ctx *c, int sid) {
  const int g = tail_group(sid);
  if (c->tail[g]) return 0;
  const size_t bytes = (size_t)c->scalar_len * (TAIL_LO[g + 1] - TAIL_LO[g]) * sizeof(double);
#ifndef VFS_EMU
  if (c->capturing) { set_err(c, "first use of a tail scalar during graph capture"); return VFS_ERR_CUDA; }
  CK(cudaMalloc((void **)&c->tail[g], bytes));
  CK(cudaMemsetAsync(c->tail[g], 0, bytes, c->stream));
#else
  c->tail[g] = (double *)calloc(bytes, 1);
#endif
  for (int s = TAIL_LO[g]; s < TAIL_LO[g + 1]; s++) c->d.s[s] = c->tail[g] + (long)(s - TAIL_LO[g]) * c->scalar_len;
  graph_reset(c);
  return 0;
}
// ---- transfers ------------------------------------------------------------------------------------
struct HasSolid { VfsDev d; int *flag; VFS_HD void operator()(int i, int j, int k) const { if ((int)(d.s[S_NV][d.idx(i, j, k)] + 0.1) == 3) *flag = 1; } };
static int h2d_stage(vfs_ctx *c, const double *host, int dof) {
  size_t n = (size_t)c->d.nzl * c->d.my * c->d.mx * dof * sizeof(double);
#ifndef VFS_EMU
  CK(cudaMemcpyAsync(c->stage, host, n, cudaMemcpyDefault, c->stream));      // `host` may also be a device pointer (unified addressing)
#else
  memcpy(c->stage, host, n);
#endif
  return 0;
}
static int d2h_stage(vfs_ctx *c, double *host, int dof) {
  size_t n = (size_t)c->d.nzl * c->d.my * c->d.mx * dof * sizeof(double);
#ifndef VFS_EMU
  CK(cudaMemcpyAsync(host, c->stage, n, cudaMemcpyDefault, c->stream));
  CK(cudaStreamSynchronize(c->stream));
#else
  memcpy(host, c->stage, n);
#endif
  return 0;
}
extern "C" int vfs_halo_exchange(vfs_ctx *c, int field) {
  if (!c || field < 0 || field >= VFS_NFIELDS_PUBLIC) return VFS_ERR_ARG;
  if (FIELD[field].s0 >= S_TAIL0) return 0;            // Ucont_rm1, Conv, Visc: ghosts never read; Phi's are refreshed by its users
  return g2l(c, grp(FIELD[field].s0, FIELD[field].dof));
}
extern "C" int vfs_upload(vfs_ctx *c, int field, const double *host) {
  if (!c || !host || field < 0 || field >= VFS_NFIELDS_PUBLIC) return VFS_ERR_ARG;
  if (field == VFS_AJ) c->iaj_valid = false;
  if (field == VFS_AJ || field == VFS_CSI || field == VFS_ETA || field == VFS_ZET || field == VFS_NVERT) c->lesgeo_valid = c->les_bnd_valid = false;
  if (field == VFS_UCAT || field == VFS_AJ || field == VFS_CSI || field == VFS_ETA || field == VFS_ZET || field == VFS_NVERT) c->sabs_valid = false;
  if (field == VFS_NVERT) { c->near_valid = false; c->wall_marked = false; }
  const bool tail = FIELD[field].s0 >= S_TAIL0;
  if (tail) RUN(ensure_tail(c, FIELD[field].s0));
  RUN(h2d_stage(c, host, FIELD[field].dof));
  UnpackAoS f = {c->d, c->stage, FIELD[field].s0, FIELD[field].dof};
  RUN(launch(c, box_owned(c), f));
  if (!tail) RUN(vfs_halo_exchange(c, field));          // (the tail fields' ghosts are never read)
  if (field == VFS_NVERT) {
    // solid-cell flag: lets Contra2Cart skip its whole-volume "solid -> 0" sweep.  Scanned on the device over the
    // slab AND the ghost planes Contra2Cart is replayed on (a neighbour's solid cells must be zeroed there too).
#ifndef VFS_EMU
    if (!c->d_flag) CK(cudaMalloc((void **)&c->d_flag, sizeof(int)));
    CK(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
#else
    if (!c->d_flag) c->d_flag = (int *)malloc(sizeof(int));
    *c->d_flag = 0;
#endif
    HasSolid hs = {c->d, c->d_flag};
    Box b = {0, c->d.mx, 0, c->d.my, -VFS_G, c->d.nzl + VFS_G};
    RUN(launch(c, b, hs));
    int any = 1;
#ifndef VFS_EMU
    CK(cudaMemcpyAsync(&any, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
#else
    any = *c->d_flag;
#endif
    c->has_solid = any != 0;
  }
  return vfs_sync(c);
}
extern "C" int vfs_download(vfs_ctx *c, int field, double *host) {
  if (!c || !host || field < 0 || field >= VFS_NFIELDS_PUBLIC) return VFS_ERR_ARG;
  if (FIELD[field].s0 >= S_TAIL0) RUN(ensure_tail(c, FIELD[field].s0));
  PackAoS f = {c->d, c->stage, FIELD[field].s0, FIELD[field].dof};
  RUN(launch(c, box_owned(c), f));
  return d2h_stage(c, host, FIELD[field].dof);
}

extern "C" int vfs_download_async(vfs_ctx *c, int field, double *host, int slot) {
  if (!c || !host || field < 0 || field >= VFS_NFIELDS_PUBLIC || slot < 0 || slot > 1) return VFS_ERR_ARG;
#ifndef VFS_EMU
  const size_t n = (size_t)c->d.nzl * c->d.my * c->d.mx * FIELD[field].dof * sizeof(double);
  if (FIELD[field].s0 >= S_TAIL0) RUN(ensure_tail(c, FIELD[field].s0));
  if (!c->copy_stream) {
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming));
  }
  if (c->stage_async[slot] && c->stage_async_bytes[slot] < n) { CK(cudaStreamSynchronize(c->copy_stream)); cudaFree(c->stage_async[slot]); c->stage_async[slot] = nullptr; }
  if (!c->stage_async[slot]) { CK(cudaMalloc((void **)&c->stage_async[slot], n)); c->stage_async_bytes[slot] = n; }
  PackAoS f = {c->d, c->stage_async[slot], FIELD[field].s0, FIELD[field].dof};
  RUN(launch(c, box_owned(c), f));
  CK(cudaEventRecord(c->ev_pack, c->stream));
  CK(cudaStreamWaitEvent(c->copy_stream, c->ev_pack, 0));
  CK(cudaMemcpyAsync(host, c->stage_async[slot], n, cudaMemcpyDeviceToHost, c->copy_stream));
  return 0;
#else
  (void)slot; return vfs_download(c, field, host);
#endif
}
extern "C" int vfs_download_wait(vfs_ctx *c) {
  if (!c) return VFS_ERR_ARG;
#ifndef VFS_EMU
  if (c->copy_stream) CK(cudaStreamSynchronize(c->copy_stream));
#endif
  return 0;
}

// ---- FormMetrics ------------------------------------------------------------------------------------
extern "C" int vfs_form_metrics(vfs_ctx *c) {
  if (!c) return VFS_ERR_ARG;
  const VfsDev &d = c->d;
  RUN(halo_k(c, grp(S_X, 3)));                       // cell (.,.,k) needs node plane k-1
  { MetricsCenter f = {d}; RUN(launch(c, box_interior(c), f)); }
  for (int side = 0; side < 2; side++) { MetricsMirror f = {d, 0, side}; int i = side ? d.mx - 1 : 0; Box b = {i, i + 1, 0, d.my, 0, d.nzl}; RUN(launch(c, b, f)); }
  for (int side = 0; side < 2; side++) { MetricsMirror f = {d, 1, side}; int j = side ? d.my - 1 : 0; Box b = {0, d.mx, j, j + 1, 0, d.nzl}; RUN(launch(c, b, f)); }
  if (d.kofs == 0) { MetricsMirror f = {d, 2, 0}; Box b = {0, d.mx, 0, d.my, 0, 1}; RUN(launch(c, b, f)); }
  if (d.kofs + d.nzl == d.mz) { MetricsMirror f = {d, 2, 1}; Box b = {0, d.mx, 0, d.my, d.nzl - 1, d.nzl}; RUN(launch(c, b, f)); }
  Grp g = grp(S_CSI0, 10);
  RUN(g2l(c, g));
  if (any_per(c)) { RUN(node_copy(c, g)); RUN(g2l(c, g)); }
  c->iaj_valid = false; c->sabs_valid = false; c->lesgeo_valid = c->les_bnd_valid = false;
  return vfs_sync(c);
}

// ---- Contra2Cart ------------------------------------------------------------------------------------
struct CopyScalar3 { VfsDev d; int from, to; VFS_HD void operator()(int i, int j, int k) const { long p = d.idx(i, j, k); for (int a = 0; a < 3; a++) d.s[to + a][p] = d.s[from + a][p]; } };

static int run_les_derive_boundary(vfs_ctx *c) {
  LesDeriveBoundary f = {c->d, (c->les_bnd_valid && c->halo_trim) ? 1 : 0};
  c->les_bnd_valid = true;
  return launch_shell(c, 0, c->d.nzl, f);
}

// Contra2Cart_2 (rhs.c:65-749).  Between ranks, ucat's ghost planes are not exchanged: every operation
// of the function is REPLAYED on the three ghost planes either side of the slab (C2C_EXT), from the
// ucont / metric / nvert ghosts that are already there, exactly as the owner of those planes runs it
// (same inputs, same code: bitwise the owner's values).  A ghost plane across the periodic seam is
// evaluated as the global plane it images (VfsDev::kglob): images of interior planes are recomputed,
// the image of the boundary plane mz-1 (0) receives the periodic node copy of local plane 1 (nzl-2),
// which is what its owner copies into it.  This removes 4 (8 at the seam ranks) of the 17 inter-rank
// exchanges of one RHS+LES unit; a single rank keeps the wrap-fill path (no ghost planes to replay).
#define C2C_EXT 3
static int ensure_wm_table(vfs_ctx *c);
static int contra2cart(vfs_ctx *c) {
  const VfsDev &d = c->d;
  c->sabs_valid = false;
  Grp gu = grp(S_U0, 3);
  const bool multi = c->prm.nranks > 1;
  const int ka = multi && (d.kofs > 0 || d.perz) ? -C2C_EXT : 0;
  const int kb = multi && (d.kofs + d.nzl < d.mz || d.perz) ? d.nzl + C2C_EXT : d.nzl;
  // refresh of the periodic i/j ghosts (+ the k wrap of a single rank) and the periodic node copies, on [ka, kb)
  auto refresh = [&](bool after_copy) -> int {
    if (!multi) return after_copy ? g2l_after_copy(c, gu) : g2l(c, gu);
    return wrap_ij(c, gu, ka, kb);
  };
  auto copy_nodes = [&](const Grp &g) -> int {
    if (!multi) return node_copy(c, g);
    NodeCopy f = {d, g, 1};
    return launch_shell(c, ka, kb, f, true, SHELL_PERIODIC_ONLY);
  };
  // ghost refresh, periodic node copies, ghost refresh
  auto refresh3 = [&]() -> int {
    if (!multi && c->fuse_refresh) return refresh_fused(c, gu, 3, 3);      // ucat is read at most 3 deep (index -3 / m+2: the reference's DA ghost width)
    RUN(refresh(false));
    if (any_per(c)) { RUN(copy_nodes(gu)); RUN(refresh(true)); }
    return 0;
  };
  if (any_per(c)) RUN(copy_nodes(grp(S_UC0, 3)));                    // rhs.c:129-156
  ev_rec(c, 2 * VFS_T_C2C);
  Box bi = box_interior(c);
  if (multi) { bi.k0 = ka; bi.k1 = kb; }                             // the functors skip planes that are not interior (kglob)
  { C2CInterior f = {d}; RUN(launch(c, bi, f)); }                    // rhs.c:158-247
  ev_rec(c, 2 * VFS_T_C2C + 1);
  RUN(refresh3());                                                   // rhs.c:251-291
  const bool wallfn = has_wallfn(d);
  if (wallfn) {                                                      // wall-function first cells read interior neighbours of the snapshot too
    CopyScalar3 f = {d, S_U0, S_FP0}; Box b = {0, d.mx, 0, d.my, ka, kb}; RUN(launch(c, b, f));
  } else { CopyScalar3 f = {d, S_U0, S_FP0}; RUN(launch_shell(c, ka, kb, f, true)); }     // lUcat snapshot read by the rules
  const bool mirror_top = d.bc[3] == 13 || d.bc[3] == 14;
  { C2CGhostRules f = {d, mirror_top ? 0 : -1}; RUN(launch_shell(c, ka, kb, f, true)); }     // rhs.c:302-682 (boundary nodes)
  if (wallfn) {                                                      // rhs.c:311-440 (first interior cells of the -1 / -2 sides)
    RUN(ensure_wm_table(c));
    C2CWallFn f = {d, c->wm_table};
    const int pl[4] = {1, d.mx - 2, 1, d.my - 2};
    for (int q = 0; q < 4; q++) {
      if (d.bc[q] != -1 && d.bc[q] != -2) continue;
      Box b = bi; b.i0 = 1; b.i1 = d.mx - 1; b.j0 = 1; b.j1 = d.my - 1;
      if (q < 2) { b.i0 = pl[q]; b.i1 = pl[q] + 1; } else { b.j0 = pl[q]; b.j1 = pl[q] + 1; }
      RUN(launch(c, b, f));
    }
  }
  {                                                                  // rhs.c:305-308,676-681 (interior nodes)
    C2CInteriorFix f = {d};
    if (c->has_solid) RUN(launch(c, bi, f));                         // solid cells -> 0: the whole interior
    else {                                                           // only the "corner" lines can change
      const int *bc = d.bc;
      const bool cor[4] = {bc[0] <= 1 && bc[2] <= 1, bc[1] <= 1 && bc[2] <= 1, bc[0] <= 1 && bc[3] <= 1, bc[1] <= 1 && bc[3] <= 1};
      const int ci[4] = {1, d.mx - 2, 1, d.mx - 2}, cj[4] = {1, 1, d.my - 2, d.my - 2};
      for (int q = 0; q < 4; q++) if (cor[q]) { Box b = {ci[q], ci[q] + 1, cj[q], cj[q] + 1, bi.k0, bi.k1}; RUN(launch(c, b, f)); }
    }
  }
  if (mirror_top) { C2CGhostRules f = {d, 1}; RUN(launch_shell(c, ka, kb, f, true)); }      // the j = my-1 plane, after its j-1 neighbours are final
  RUN(refresh3());                                                   // rhs.c:690-748
  return 0;
}
extern "C" int vfs_contra2cart(vfs_ctx *c) { if (!c) return VFS_ERR_ARG; RUN(contra2cart(c)); return api_end(c); }

// ---- IB_BC ---------------------------------------------------------------------------------------------
static int ib_bc(vfs_ctx *c) {
  const VfsDev &d = c->d;
  bool wallfn_any = false;
  for (int q = 0; q < 6; q++) wallfn_any = wallfn_any || d.bc[q] == -1 || d.bc[q] == -2;
  if (wallfn_any && !d.immersed && d.ti == d.tistart && !c->wall_marked) {      // momentum.c:2048-2074
#ifndef VFS_EMU
    if (c->capturing) { set_err(c, "the first-step nvert marking of IB_BC cannot be captured into a graph: run the step eagerly once"); return VFS_ERR_CUDA; }
#endif
    { IbBcMarkWall f = {d}; RUN(launch(c, box_interior(c), f)); }
    RUN(g2l(c, grp(S_NV, 1)));
    c->wall_marked = true; c->near_valid = false; c->lesgeo_valid = c->les_bnd_valid = false; c->sabs_valid = false;
  }
  if (any_per(c)) RUN(node_copy(c, grp(S_U0, 3)));                   // momentum.c:2086-2107
  if (d.immersed) { IbBcFaces f = {d}; RUN(launch(c, box_interior(c), f)); }
  if (has_wallfn(d)) {                                               // momentum.c:2169-2189
    IbBcWallFn f = {d};
    const int pl[4] = {1, d.mx - 2, 1, d.my - 2};
    for (int q = 0; q < 4; q++) {
      if (d.bc[q] != -1 && d.bc[q] != -2) continue;
      Box b = box_interior(c);
      if (q < 2) { b.i0 = pl[q]; b.i1 = pl[q] + 1; } else { b.j0 = pl[q]; b.j1 = pl[q] + 1; }
      RUN(launch(c, b, f));
    }
  }
  IbBcBoundary f = {d};
  const int m[3] = {d.mx, d.my, d.mz};
  const int per[3] = {d.perx, d.pery, d.perz};
  for (int D = 0; D < 3; D++) {
    int planes[3] = {0, m[D] - 2, m[D] - 1};
    bool need[3] = {d.bc[2 * D] == 10 || per[D], d.bc[2 * D + 1] == 10, per[D] != 0};
    for (int q = 0; q < 3; q++) {
      if (!need[q]) continue;
      Box b = box_owned(c);
      if (D == 0) { b.i0 = planes[q]; b.i1 = b.i0 + 1; }
      else if (D == 1) { b.j0 = planes[q]; b.j1 = b.j0 + 1; }
      else { int k = planes[q] - d.kofs; if (k < 0 || k >= d.nzl) continue; b.k0 = k; b.k1 = k + 1; }
      RUN(launch(c, b, f));
    }
  }
  return g2l(c, grp(S_UC0, 3), 2, 2);                                // momentum.c:2231-2232 (read next by the k-face fluxes: uc(-2) at the periodic seam)
}
extern "C" int vfs_ib_bc(vfs_ctx *c) { if (!c) return VFS_ERR_ARG; RUN(ib_bc(c)); return api_end(c); }

// ---- Formfunction_2 ------------------------------------------------------------------------------------
static Box box_clip(Box b, const Box &lim) {
  if (b.i0 < lim.i0) b.i0 = lim.i0; if (b.i1 > lim.i1) b.i1 = lim.i1;
  if (b.j0 < lim.j0) b.j0 = lim.j0; if (b.j1 > lim.j1) b.j1 = lim.j1;
  if (b.k0 < lim.k0) b.k0 = lim.k0; if (b.k1 > lim.k1) b.k1 = lim.k1;
  return b;
}
static bool box_empty(const Box &b) { return b.i1 <= b.i0 || b.j1 <= b.j0 || b.k1 <= b.k0; }
// The staged chain on the owned nodes outside the regular region R of the marching kernel: six
// disjoint projection slabs; Fp on each slab grown by one cell, face fluxes on that grown by
// (-2,+1) along the face normal (the reach of the 4th-order divergence, momentum.c:1565-1637).
struct ShellBoxes { int n; Box proj[6], fp[6], fl[6][3]; };
static ShellBoxes shell_boxes(const vfs_ctx *c, const Box &R) {
  const VfsDev &d = c->d;
  ShellBoxes S; S.n = 0;
  const Box own = {0, d.mx, 0, d.my, 0, d.nzl};
  const Box cells = box_interior(c);
  const int k0f = klo(c, 0), k1f = klo(c, d.mz - 1);
  const Box cand[6] = {{0, d.mx, 0, d.my, 0, R.k0},         {0, d.mx, 0, d.my, R.k1, d.nzl},
                       {0, d.mx, 0, R.j0, R.k0, R.k1},      {0, d.mx, R.j1, d.my, R.k0, R.k1},
                       {0, R.i0, R.j0, R.j1, R.k0, R.k1},   {R.i1, d.mx, R.j0, R.j1, R.k0, R.k1}};
  for (int q = 0; q < 6; q++) {
    const Box pb = box_clip(cand[q], own);
    if (box_empty(pb)) continue;
    const int n = S.n++;
    S.proj[n] = pb;
    const Box grown = {pb.i0 - 1, pb.i1 + 1, pb.j0 - 1, pb.j1 + 1, pb.k0 - 1, pb.k1 + 1};
    const Box fb = box_clip(grown, cells);
    S.fp[n] = fb;
    const Box fi = {fb.i0 - 2, fb.i1 + 1, fb.j0, fb.j1, fb.k0, fb.k1}, li = {0, d.mx - 1, 1, d.my - 1, cells.k0, cells.k1};
    const Box fj = {fb.i0, fb.i1, fb.j0 - 2, fb.j1 + 1, fb.k0, fb.k1}, lj = {1, d.mx - 1, 0, d.my - 1, cells.k0, cells.k1};
    const Box fk = {fb.i0, fb.i1, fb.j0, fb.j1, fb.k0 - 2, fb.k1 + 1}, lk = {1, d.mx - 1, 1, d.my - 1, k0f, k1f};
    S.fl[n][0] = box_clip(fi, li); S.fl[n][1] = box_clip(fj, lj); S.fl[n][2] = box_clip(fk, lk);
  }
  return S;
}

// S_IAJ = 1/aj over the whole padded array (filter weights; harmonic-mean face Jacobians)
static int ensure_iaj(vfs_ctx *c) {
  if (c->iaj_valid) return 0;
  const VfsDev &d = c->d;
  InvAj f = {d}; Box all = {-VFS_G, d.mx + VFS_G, -VFS_G, d.my + VFS_G, -VFS_G, d.nzl + VFS_G};
  RUN(launch(c, all, f)); c->iaj_valid = true;
  return 0;
}

// near-solid byte mask (see NearSolid); with the fast paths switched off every node reads "near"
static int ensure_near(vfs_ctx *c) {
  if (c->near_valid) return 0;
  const VfsDev &d = c->d;
#ifndef VFS_EMU
  if (c->capturing) { set_err(c, "near-solid mask must be built before graph capture"); return VFS_ERR_CUDA; }
  CK(cudaMemsetAsync(c->near, 1, (size_t)c->scalar_len, c->stream));
#else
  memset(c->near, 1, (size_t)c->scalar_len);
#endif
  if (c->fastpath) { NearSolid f = {d, c->near}; Box b = {-2, d.mx + 2, -2, d.my + 2, -2, d.nzl + 2}; RUN(launch(c, b, f)); }
  c->near_valid = true;
  return 0;
}

// Cabot wall model at the j = 0 faces (viscosity_wallmodel, momentum.c:1139-1154): table once, then one
// Newton solve per face of the plane, before the flux kernels read the override
static int ensure_wm_table(vfs_ctx *c) {
  if (!c->wm_table) {
    const size_t bytes = (size_t)(VFS_WM_NYP + 1) * sizeof(double);
#ifndef VFS_EMU
    if (c->capturing) { set_err(c, "wall-model table must be built before graph capture"); return VFS_ERR_CUDA; }
    CK(cudaMalloc((void **)&c->wm_table, bytes));
#else
    c->wm_table = (double *)malloc(bytes);
#endif
    { WmTableIntervals f = {c->wm_table}; Box b = {0, VFS_WM_NYP + 1, 0, 1, 0, 1}; RUN(launch(c, b, f)); }
#ifndef VFS_EMU
    {   // the running sum in index order (the reference's serial loop, wallfunction.c:291-300): 4 MB, once — on the host, where
        // a serial scan costs a millisecond, instead of one GPU thread walking 500 001 entries
      std::vector<double> h(VFS_WM_NYP + 1);
      CK(cudaMemcpyAsync(h.data(), c->wm_table, bytes, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      double acc = 0.;
      for (int i = 1; i <= VFS_WM_NYP; i++) { acc = acc + h[i]; h[i] = acc; }
      CK(cudaMemcpyAsync(c->wm_table, h.data(), bytes, cudaMemcpyHostToDevice, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
#else
    { WmTableScan f = {c->wm_table}; Box b = {0, 1, 0, 1, 0, 1}; RUN(launch(c, b, f)); }
#endif
  }
  return 0;
}
static int wall_model(vfs_ctx *c) {
  const VfsDev &d = c->d;
  RUN(ensure_wm_table(c));
  WallModelPlane f = {d, c->wm_table};
  Box b = {1, d.mx - 1, 0, 1, klo(c, 1), klo(c, d.mz - 1)};
  return launch(c, b, f);
}

// mode 0: rhs[s0] += scale*R, masks (Formfunction_2) ; mode 1: full SNES assembly into S_R0
static int formfunction2(vfs_ctx *c, int mode, int s0, double scale) {
  const VfsDev &d = c->d;
  RUN(ensure_iaj(c));
  RUN(ensure_near(c));
  if (d.visc_wm && d.les) RUN(wall_model(c));
  if (any_per(c)) RUN(node_copy(c, grp(S_UC0, 3)));                  // momentum.c:638-666
  const int k1 = klo(c, 1), k2 = klo(c, d.mz - 1);
  Box R;
  // weno / skew / clark / inviscid: the one-thread-per-face kernels with the variants compiled in (face_flux_core<.., X>)
  const bool ex = d.weno || d.skew || d.clark || d.inviscid;
  if (ex && d.skew) RUN(ensure_tail(c, S_ADV1));
  bool march = !ex && c->fused == 2 && RhsMarch::region(d, R);     // experimental fully fused residual (option 0 = 2)
#ifndef VFS_EMU
  march = march && c->tma_ok;
#endif
  ShellBoxes S; S.n = 1;
  S.proj[0] = box_owned(c); S.fp[0] = box_interior(c);
  { Box f0 = {0, d.mx - 1, 1, d.my - 1, k1, k2}, f1 = {1, d.mx - 1, 0, d.my - 1, k1, k2}, f2 = {1, d.mx - 1, 1, d.my - 1, klo(c, 0), k2};
    S.fl[0][0] = f0; S.fl[0][1] = f1; S.fl[0][2] = f2; }
  ev_rec(c, 2 * VFS_T_FLUX);
  if (march) {
    // regular interior: one marching kernel, fluxes and Fp stay on chip
    RhsMarch P = {d, mode, s0, scale, R};
#ifndef VFS_EMU
    if (run_rhs_march(c->stream, c->tmap_rhs, P, &c->launches)) { set_err(c, "k_rhs_march launch failed"); return VFS_ERR_CUDA; }
#else
    run_rhs_march(c->stream, P, &c->launches);
#endif
    S = shell_boxes(c, R);
  }
  // staged chain: the whole domain, or the boundary slabs around R
#ifndef VFS_EMU
  if (!march && !ex && c->fused && c->tma_ok) {
    // regular faces: TMA-staged tiled kernel; faces 0 and m-2 along their normal: staged kernels on thin slabs
    // the six thin slabs write faces the marching kernel does not: they run beside it on the side stream
    SideScope sc; RUN(side_begin(c, &sc));
    { FaceFlux<0> f = {d}; Box b0 = {0, 1, 1, d.my - 1, k1, k2}, b1 = {d.mx - 2, d.mx - 1, 1, d.my - 1, k1, k2}; RUN(launch(c, b0, f)); RUN(launch(c, b1, f)); }
    { FaceFlux<1> f = {d}; Box b0 = {1, d.mx - 1, 0, 1, k1, k2}, b1 = {1, d.mx - 1, d.my - 2, d.my - 1, k1, k2}; RUN(launch(c, b0, f)); RUN(launch(c, b1, f)); }
    { FaceFlux<2> f = {d};
      Box b0 = {1, d.mx - 1, 1, d.my - 1, klo(c, 0), klo(c, 1)}, b1 = {1, d.mx - 1, 1, d.my - 1, klo(c, d.mz - 2), klo(c, d.mz - 1)};
      RUN(launch(c, b0, f)); RUN(launch(c, b1, f)); }
    side_end(c, &sc);
    if (c->flux_var == 1 ? launch_flux_march(c->stream, c->tmap_fluxA, c->tmap_fluxB, d, k1, k2, &c->launches)
                         : launch_flux_tma(c->stream, c->tmap_flux, d, k1, k2, c->flux_minb, &c->launches)) { set_err(c, "flux kernel launch failed"); return VFS_ERR_CUDA; }
    RUN(ovl_join(c));
  } else
#endif
  for (int n = 0; n < S.n; n++) {
    if (ex) {
      { FaceFluxX<0> f = {d}; RUN(launch(c, S.fl[n][0], f)); }
      { FaceFluxX<1> f = {d}; RUN(launch(c, S.fl[n][1], f)); }
      { FaceFluxX<2> f = {d}; RUN(launch(c, S.fl[n][2], f)); }
      continue;
    }
    { FaceFlux<0> f = {d}; RUN(launch(c, S.fl[n][0], f)); }
    { FaceFlux<1> f = {d}; RUN(launch(c, S.fl[n][1], f)); }
    { FaceFlux<2> f = {d}; RUN(launch(c, S.fl[n][2], f)); }
  }
  ev_rec(c, 2 * VFS_T_FLUX + 1);
  const Grp gi = grp_cat(grp(S_FC1, 3), grp(S_FV1, 3)), gj = grp_cat(grp(S_FC2, 3), grp(S_FV2, 3)), gk = grp_cat(grp(S_FC3, 3), grp(S_FV3, 3));
  const int ka = d.kofs > 0 ? -VFS_G : 0, kb = d.kofs + d.nzl < d.mz ? d.nzl + VFS_G : d.nzl;      // as node_copy()
  if (c->fused == 1 && c->fp_fused && S.n == 1 && !ex) {
    // ---- Fp folded into the projection (ProjFpMarch, vfs_march_kernels.h) ----
    // Flux ghosts as below; between ranks ALL face-flux families travel (Fp of the first ghost plane is evaluated
    // locally from them), in ONE exchange instead of the k-family + Fp exchanges of the staged chain.
    ev_rec(c, 2 * VFS_T_FP);
    if (d.perx) { WrapFill f = {d, gi, 0}; Box b = {0, 2 * VFS_G, 0, d.my, 0, d.nzl}; RUN(launch(c, b, f)); }
    if (d.pery) { WrapFill f = {d, gj, 1}; Box b = {-VFS_G, d.mx + VFS_G, 0, 2 * VFS_G, 0, d.nzl}; RUN(launch(c, b, f)); }
    const bool multi = c->prm.nranks > 1;
    const Grp gall = grp_cat(grp_cat(gi, gj), gk);
    ProjFpMarch P = {d, mode, s0, scale};
    auto shell_nodes = [&]() -> int {          // boundary nodes of the slab: every component masked, no Fp read
      if (mode == 0) { ProjectAdd f = {d, s0, scale}; return launch_shell(c, 0, d.nzl, f); }
      ProjectSNES f = {d}; return launch_shell(c, 0, d.nzl, f);
    };
    auto march = [&](int q0, int q1) -> int {
      q0 = q0 < k1 ? k1 : q0; q1 = q1 > k2 ? k2 : q1;
      if (c->fp_fused == 2) { ProjectFpBox f = {d, mode, s0, scale}; Box b = {1, d.mx - 1, 1, d.my - 1, q0, q1}; return launch(c, b, f); }
      if (run_projfp_march(c->stream, P, q0, q1, &c->launches)) { set_err(c, "projection kernel launch failed"); return VFS_ERR_CUDA; }
      return 0;
    };
    const bool ovl = multi && can_overlap(c) && d.nzl >= 10;
    ev_rec(c, 2 * VFS_T_FP + 1);
    ev_rec(c, 2 * VFS_T_PROJECT);
    if (!ovl) {
      RUN(halo_k(c, multi ? gall : gk, false, 3, 2));
      if (any_per(c)) { NodeCopyFlux f = {d, 3}; RUN(launch_shell(c, ka, kb, f, false, SHELL_PERIODIC_ONLY)); }
      RUN(shell_nodes());
      RUN(march(k1, k2));
    } else {
      // the i/j-plane copies of the owned planes go first (they are part of what the neighbours receive); the cells
      // of local planes 2 .. nzl-4 and the Fp planes up to nzl-3 touch no k ghost plane and no seam copy
      if (any_per(c)) { NodeCopyFlux f = {d, 1}; RUN(launch_shell(c, 0, d.nzl, f, false, SHELL_PERIODIC_ONLY)); }
      RUN(ovl_exchange(c, gall, 3, 2));
      RUN(shell_nodes());
      RUN(march(2, d.nzl - 3));
      RUN(ovl_join(c));
      if (any_per(c)) { NodeCopyFlux f = {d, 3}; RUN(launch_shell(c, ka, kb, f, false, SHELL_PERIODIC_ONLY)); }
      RUN(march(0, 2)); RUN(march(d.nzl - 3, d.nzl));
    }
    ev_rec(c, 2 * VFS_T_PROJECT + 1);
    return 0;
  }
  ev_rec(c, 2 * VFS_T_FP);
  // momentum.c:1458-1496, 1506-1546.  FpCell reads the fluxes of face family D along direction D only, at the
  // cell's own other two indices, so each family is refreshed in its own direction only (a third of the data;
  // between ranks only the k-face family travels).
  {
    if (d.perx) { WrapFill f = {d, gi, 0}; Box b = {0, 2 * VFS_G, 0, d.my, 0, d.nzl}; RUN(launch(c, b, f)); }
    if (d.pery) { WrapFill f = {d, gj, 1}; Box b = {-VFS_G, d.mx + VFS_G, 0, 2 * VFS_G, 0, d.nzl}; RUN(launch(c, b, f)); }
    const bool ovl = can_overlap(c) && S.n == 1 && d.nzl >= 8 && !ex;
    FpCell fp1 = {d};
    FpCell2 fp2 = {d};
    FpCellX fpx = {d};
    if (ex && d.skew) {      // the advective half Adv1-3 travels with the fluxes (momentum.c:1470-1484, 1540-1544)
      const Grp ai = grp(S_ADV1, 3), aj = grp(S_ADV2, 3), ak = grp(S_ADV3, 3);
      if (d.perx) { WrapFill f = {d, ai, 0}; Box b = {0, 2 * VFS_G, 0, d.my, 0, d.nzl}; RUN(launch(c, b, f)); }
      if (d.pery) { WrapFill f = {d, aj, 1}; Box b = {-VFS_G, d.mx + VFS_G, 0, 2 * VFS_G, 0, d.nzl}; RUN(launch(c, b, f)); }
      RUN(halo_k(c, ak, false, 3, 2));
      if (any_per(c)) { NodeCopyAdv f = {d}; RUN(launch_shell(c, ka, kb, f, false, SHELL_PERIODIC_ONLY)); }
    }
    // two cells per thread (16-byte loads) when the box starts at the first interior cell; the functor skips i = 0 / mx-1
    auto fp_launch = [&](const Box &b) -> int {
      if (ex) return launch(c, b, fpx);
      if (c->fp_pairs && b.i0 == 1 && b.i1 == d.mx - 1) { Box h = b; h.i0 = 0; h.i1 = (d.mx + 1) / 2; return launch(c, h, fp2); }
      return launch(c, b, fp1);
    };
    if (!ovl) {
      RUN(halo_k(c, gk, false, 3, 2));      // Fp reads faces k-2 .. k+1 (k-4 / k+3 across the periodic seam = ghost planes -3 / nzl+1)
      if (any_per(c)) { NodeCopyFlux f = {d, 3}; RUN(launch_shell(c, ka, kb, f, false, SHELL_PERIODIC_ONLY)); }
      for (int n = 0; n < S.n; n++) RUN(fp_launch(S.fp[n]));          // momentum.c:1548-1678
    } else {
      // Fp of a cell reads the k-face fluxes of planes k-2 .. k+1 (k-4 / k+3 across the periodic seam): the cells
      // of local planes 2 .. nzl-3 never touch a k ghost plane and run while the exchange is in flight
      RUN(ovl_exchange(c, gk, 3, 2));
      if (any_per(c)) { NodeCopyFlux f = {d, 1}; RUN(launch_shell(c, ka, kb, f, false, SHELL_PERIODIC_ONLY)); }
      Box in = S.fp[0], lo = S.fp[0], hi = S.fp[0];
      in.k0 = in.k0 > 2 ? in.k0 : 2; in.k1 = in.k1 < d.nzl - 2 ? in.k1 : d.nzl - 2;
      lo.k1 = in.k0; hi.k0 = in.k1;
      RUN(fp_launch(in));
      RUN(ovl_join(c));
      if (d.perz) { NodeCopyFlux f = {d, 2}; RUN(launch_shell(c, ka, kb, f, false, SHELL_PERIODIC_ONLY)); }
      RUN(fp_launch(lo)); RUN(fp_launch(hi));
    }
  }
  ev_rec(c, 2 * VFS_T_FP + 1);
  Grp gp = grp(S_FP0, 3);
  const bool ovl_p = can_overlap(c) && S.n == 1 && d.nzl >= 8;
  auto project = [&](const Box &b) -> int {
    if (mode == 0) { ProjectAdd f = {d, s0, scale}; return launch(c, b, f); }
    ProjectSNES f = {d}; return launch_occ(c, b, f);
  };
  if (!ovl_p) {
    RUN(g2l(c, gp, 2, 2));                                            // the projection reads Fp at k+1, the seam copies at k+-2
    if (any_per(c)) RUN(node_copy(c, gp));                            // momentum.c:1687-1713
    ev_rec(c, 2 * VFS_T_PROJECT);
    for (int n = 0; n < S.n; n++) RUN(project(S.proj[n]));
  } else {
    // the projection of a node reads Fp at the node and at its +i, +j, +k neighbours: local planes 2 .. nzl-3 need
    // neither the ghost plane nzl nor the seam copies of planes 0 / nzl-1.  The periodic node copies run twice:
    // before (their i/j part feeds the interior planes) and again after the exchange (fresh k ghosts).
    RUN(wrap_ij(c, gp));
    RUN(ovl_exchange(c, gp, 2, 2));
    if (any_per(c)) RUN(node_copy(c, gp));
    ev_rec(c, 2 * VFS_T_PROJECT);
    Box in = S.proj[0], lo = S.proj[0], hi = S.proj[0];
    in.k0 = 2; in.k1 = d.nzl - 2; lo.k1 = 2; hi.k0 = d.nzl - 2;
    RUN(project(in));
    RUN(ovl_join(c));
    if (any_per(c)) RUN(node_copy(c, gp));
    RUN(project(lo)); RUN(project(hi));
  }
  ev_rec(c, 2 * VFS_T_PROJECT + 1);
  return 0;
}
extern "C" int vfs_formfunction2(vfs_ctx *c, int rhs_field, double scale) {
  if (!c || (rhs_field != VFS_RHS && rhs_field != VFS_RHS_O)) return VFS_ERR_ARG;
  RUN(formfunction2(c, 0, FIELD[rhs_field].s0, scale));
  return vfs_sync(c);
}

static int zero_scalars(vfs_ctx *c, int s0, int n) {
#ifndef VFS_EMU
  CK(cudaMemsetAsync(c->d.s[s0], 0, (size_t)n * c->scalar_len * sizeof(double), c->stream));
#else
  memset(c->d.s[s0], 0, (size_t)n * c->scalar_len * sizeof(double));
#endif
  return 0;
}

// ---- Pressure_Gradient (momentum.c:203-439) ---------------------------------------------------------------
extern "C" int vfs_pressure_gradient(vfs_ctx *c, double k_forcing) {
  if (!c) return VFS_ERR_ARG;
  const VfsDev &d = c->d;
  RUN(ensure_iaj(c));
  const Grp gp = grp(S_P, 1);
  RUN(g2l(c, gp, 2, 2));                                              // momentum.c:247-248
  if (any_per(c)) { RUN(node_copy(c, gp)); RUN(g2l(c, gp, 2, 2)); }    // :250-286
  RUN(zero_scalars(c, S_DP0, 3));                                     // VecSet(dP, 0.), :313
  { PressureGradient f = {d, k_forcing}; RUN(launch(c, box_interior(c), f)); }
  return vfs_sync(c);
}

// ---- UpdatePressure / Projection (poisson.c:3137-3296, 2700-3040; called back to back, solvers.c:662-663) -----------
// VFS_PHI holds the Poisson solver's pressure correction (owned values; its ghosts are refreshed here, as the solver's
// own DAGlobalToLocal does, poisson.c:2661).
extern "C" int vfs_update_pressure(vfs_ctx *c) {
  if (!c) return VFS_ERR_ARG;
  RUN(ensure_tail(c, S_PHI));
  const Grp gp = grp(S_P, 1), gf = grp(S_PHI, 1), gpf = grp_cat(gp, gf);
  RUN(g2l(c, gf, 2, 2));
  { UpdateP f = {c->d}; RUN(launch(c, box_interior(c), f)); }       // :3183-3193
  RUN(g2l(c, gp, 2, 2));                                              // :3242-3243
  if (any_per(c)) RUN(node_copy(c, gpf));                             // :3249-3279: boundary nodes of P and Phi <- their periodic images
  RUN(g2l(c, gpf, 2, 2));                                             // :3288-3293
  return vfs_sync(c);
}
// Ucont -= dt * st * grad(Phi) on the faces and the component-wise periodic copies; the reference's Projection ends with
// Contra2Cart (poisson.c:3049): call vfs_contra2cart next (kept separate because Contra2Cart rewrites lUcont's periodic
// boundary nodes, rhs.c:129-156, which the reference's global Ucont does not see)
extern "C" int vfs_projection(vfs_ctx *c, double st, double poisson_threshold) {
  if (!c) return VFS_ERR_ARG;
  const VfsDev &d = c->d;
  RUN(ensure_tail(c, S_PHI));
  RUN(ensure_iaj(c));
  RUN(g2l(c, grp(S_PHI, 1), 2, 2));
  { ProjectionCorr f = {d, st, poisson_threshold}; RUN(launch(c, box_interior(c), f)); }
  const Grp gu = grp(S_UC0, 3);
  RUN(g2l(c, gu));                                                    // :2981-2982
  if (any_per(c)) {                                                   // :2984-3025
    { PeriodicCompCopy f = {d}; RUN(launch_shell(c, 0, d.nzl, f, false, SHELL_PERIODIC_ONLY)); }
    RUN(g2l(c, gu));
  }
  return vfs_sync(c);
}

// ---- cylinder force diagnostics of Formfunction_2 (momentum.c:570-579, 822-849) ------------------------------------
// out[7] = this rank's lA_cyl, lA_cyl_x, lA_cyl_z, lFpx_cyl, lFpz_cyl, lFvx_cyl, lFvz_cyl from the current UCAT (as the last
// residual evaluation left it) and P; zeros unless bctype[0] == 11 and bctype[1] == 1.  The reference sums them
// over ranks itself (main.c:1269-1277).
extern "C" int vfs_cylinder_forces(vfs_ctx *c, double *out7) {
  if (!c || !out7) return VFS_ERR_ARG;
  const VfsDev &d = c->d;
  for (int q = 0; q < 7; q++) out7[q] = 0;
  if (d.bc[0] != 11 || d.bc[1] != 1) return 0;
  RUN(ensure_iaj(c));
  const long nface = (long)d.nzl * d.my;
  double *buf = 0;
  std::vector<double> h(7 * nface);
#ifndef VFS_EMU
  CK(cudaMalloc((void **)&buf, 7 * nface * sizeof(double)));
  CK(cudaMemsetAsync(buf, 0, 7 * nface * sizeof(double), c->stream));
#else
  buf = (double *)calloc(7 * nface, sizeof(double));
#endif
  int rc = 0;
  { CylinderForce f = {d, buf, nface}; Box b = {d.mx - 2, d.mx - 1, 1, d.my - 1, klo(c, 1), klo(c, d.mz - 1)}; rc = launch(c, b, f); }
#ifndef VFS_EMU
  if (!rc && cudaMemcpyAsync(h.data(), buf, 7 * nface * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = VFS_ERR_CUDA;
  if (!rc) rc = vfs_sync(c);
  cudaFree(buf);
#else
  memcpy(h.data(), buf, 7 * nface * sizeof(double));
  free(buf);
#endif
  if (rc) return rc;
  for (int q = 0; q < 7; q++) { double s = 0; for (long f = 0; f < nface; f++) s += h[q * nface + f]; out7[q] = s; }
  return 0;
}

// ---- actuator forcing: Calc_F_eul / Calc_U_lagr (rotor_model.c:3668, 2937) -------------------------------------
// all objects' element arrays concatenated (reference order: object, element) into one device buffer
static int act_upload(vfs_ctx *c, int nobj, const vfs_actuator *o, ActElems *E, double **ulagr_dev) {
  long n = 0;
  for (int b = 0; b < nobj; b++) n += o[b].n_elmt;
  const size_t bytes = (size_t)n * (7 * sizeof(double) + 6 * sizeof(int) + 3 * sizeof(double)) + 64;
  std::vector<char> h(bytes);
  double *hd = (double *)h.data(); int *hi = (int *)(hd + 10 * n);
  long q = 0;
  for (int b = 0; b < nobj; b++) for (int l = 0; l < o[b].n_elmt; l++, q++) {
    hd[q] = o[b].cent_x[l]; hd[n + q] = o[b].cent_y[l]; hd[2 * n + q] = o[b].cent_z[l]; hd[3 * n + q] = o[b].dA ? o[b].dA[l] : 0.;
    hd[4 * n + q] = o[b].F_lagr_x ? o[b].F_lagr_x[l] : 0.; hd[5 * n + q] = o[b].F_lagr_y ? o[b].F_lagr_y[l] : 0.; hd[6 * n + q] = o[b].F_lagr_z ? o[b].F_lagr_z[l] : 0.;
    hi[q] = o[b].i_min[l]; hi[n + q] = o[b].i_max[l]; hi[2 * n + q] = o[b].j_min[l]; hi[3 * n + q] = o[b].j_max[l]; hi[4 * n + q] = o[b].k_min[l]; hi[5 * n + q] = o[b].k_max[l];
  }
#ifndef VFS_EMU
  if (c->act_bytes < bytes) { if (c->act_buf) cudaFree(c->act_buf); c->act_buf = nullptr; CK(cudaMalloc(&c->act_buf, bytes)); c->act_bytes = bytes; }
  CK(cudaMemcpyAsync(c->act_buf, h.data(), bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
#else
  if (c->act_bytes < bytes) { free(c->act_buf); c->act_buf = malloc(bytes); c->act_bytes = bytes; }
  memcpy(c->act_buf, h.data(), bytes);
#endif
  const double *dd = (const double *)c->act_buf; const int *di = (const int *)(dd + 10 * n);
  E->n = (int)n; E->cx = dd; E->cy = dd + n; E->cz = dd + 2 * n; E->dA = dd + 3 * n; E->fx = dd + 4 * n; E->fy = dd + 5 * n; E->fz = dd + 6 * n;
  E->i0 = di; E->i1 = di + n; E->j0 = di + 2 * n; E->j1 = di + 3 * n; E->k0 = di + 4 * n; E->k1 = di + 5 * n;
  if (ulagr_dev) *ulagr_dev = (double *)c->act_buf + 7 * n;
  return 0;
}
extern "C" int vfs_calc_f_eul(vfs_ctx *c, int nobj, const vfs_actuator *objs, int df, double halfwidth, int widthfixed, const double *dhf, int accumulate) {
  if (!c || nobj < 0 || (nobj > 0 && !objs) || (widthfixed && !dhf)) return VFS_ERR_ARG;
  const VfsDev &d = c->d;
  if (!accumulate) RUN(zero_scalars(c, S_FE0, 3));
  ActElems E;
  RUN(act_upload(c, nobj, objs, &E, nullptr));
  if (E.n > 0) {
    // bounding box of all windows, clipped to the owned interior cells (windows reach ghost nodes on other ranks' behalf)
    Box b = {1 << 30, -(1 << 30), 1 << 30, -(1 << 30), 1 << 30, -(1 << 30)};
    for (int o = 0; o < nobj; o++) for (int l = 0; l < objs[o].n_elmt; l++) {
      b.i0 = std::min(b.i0, objs[o].i_min[l]); b.i1 = std::max(b.i1, objs[o].i_max[l]); b.j0 = std::min(b.j0, objs[o].j_min[l]); b.j1 = std::max(b.j1, objs[o].j_max[l]);
      b.k0 = std::min(b.k0, objs[o].k_min[l] - d.kofs); b.k1 = std::max(b.k1, objs[o].k_max[l] - d.kofs);
    }
    const Box own = {0, d.mx, 0, d.my, 0, d.nzl};
    FEulGather f = {d, E, df, widthfixed, halfwidth, {widthfixed ? dhf[0] : 0., widthfixed ? dhf[1] : 0., widthfixed ? dhf[2] : 0.}};
    RUN(launch(c, box_clip(b, own), f));
  }
  { FEulMask f = {d}; RUN(launch(c, box_owned(c), f)); }                    // rotor_model.c:3832-3903
  RUN(g2l(c, grp(S_FE0, 3), 2, 2));                                          // :3955-3957
  return vfs_sync(c);
}
extern "C" int vfs_calc_u_lagr(vfs_ctx *c, int nobj, const vfs_actuator *objs) {
  if (!c || nobj < 0 || (nobj > 0 && !objs)) return VFS_ERR_ARG;
  const VfsDev &d = c->d;
  ActElems E; double *out = nullptr;
  RUN(act_upload(c, nobj, objs, &E, &out));
  if (E.n == 0) return 0;
  std::vector<double> h(3 * (size_t)E.n);
#ifndef VFS_EMU
  ULagrArgs A = {d, E, out};
  k_ulagr<<<E.n, 256, 0, c->stream>>>(A);
  c->launches++;
  if (c->prm.nranks > 1) {
    NcclApi &N = nccl_api();
    if (!c->comm || !N.AllReduce) { set_err(c, "vfs_calc_u_lagr with nranks > 1 needs vfs_nccl_init (the elements' sums are added with ncclAllReduce)"); return VFS_ERR_HALO; }
    ncclResult_t e = N.AllReduce(out, out, 3 * (size_t)E.n, ncclDouble, ncclSum, c->comm, c->stream);
    if (e != ncclSuccess) { set_err(c, std::string("ncclAllReduce: ") + N.GetErrorString(e)); return VFS_ERR_HALO; }
  }
  CK(cudaMemcpyAsync(h.data(), out, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
#else
  if (c->prm.nranks > 1) { set_err(c, "the host emulation of vfs_calc_u_lagr is single rank"); return VFS_ERR_UNSUPPORTED; }
  for (int l = 0; l < E.n; l++) {
    double acc[3] = {0, 0, 0};
    const int i0 = std::max(E.i0[l], 0), i1 = std::min(E.i1[l], d.mx), j0 = std::max(E.j0[l], 0), j1 = std::min(E.j1[l], d.my);
    const int k0 = std::max(E.k0[l] - d.kofs, 0), k1 = std::min(E.k1[l] - d.kofs, d.nzl);
    for (int k = k0; k < k1; k++) for (int j = j0; j < j1; j++) for (int i = i0; i < i1; i++) ulagr_cell(d, E, l, i, j, k, acc);
    for (int a = 0; a < 3; a++) h[3 * l + a] = acc[a];
  }
#endif
  long q = 0;
  for (int b = 0; b < nobj; b++) for (int l = 0; l < objs[b].n_elmt; l++, q++) {
    if (objs[b].U_lagr_x) objs[b].U_lagr_x[l] = h[3 * q]; if (objs[b].U_lagr_y) objs[b].U_lagr_y[l] = h[3 * q + 1]; if (objs[b].U_lagr_z) objs[b].U_lagr_z[l] = h[3 * q + 2];
  }
  return 0;
}

// ---- legacy Convection / Viscous (rhs.c:751, 1071) --------------------------------------------------------
// Face fluxes on i in [0, mx-2] (j, k interior) and twins, then the flux difference on the interior
// cells; the k-face plane just below an interior slab boundary is computed locally from the ucat /
// nu_t / metric ghost planes, so no exchange is needed (the reference's DALocalToLocal of Fp1-3,
// rhs.c:1471-1478, only refreshes ghosts nobody reads).
template <bool VISC> static int legacy_term(vfs_ctx *c) {
  RUN(ensure_tail(c, S_CONV0));
  const VfsDev &d = c->d;
  RUN(ensure_iaj(c));
  const Box bi = box_interior(c);
  if (box_empty(bi)) return 0;
  { LegacyFlux<0, VISC> f = {d}; Box b = bi; b.i0 = 0; RUN(launch(c, b, f)); }
  { LegacyFlux<1, VISC> f = {d}; Box b = bi; b.j0 = 0; RUN(launch(c, b, f)); }
  { LegacyFlux<2, VISC> f = {d}; Box b = bi; b.k0 = bi.k0 - 1; RUN(launch(c, b, f)); }
  const int so = VISC ? S_VISC0 : S_CONV0;
  RUN(zero_scalars(c, so, 3));                                        // boundary nodes: rhs.c:1506-1560 (Visc), untouched zeros (Conv)
  { LegacyDiv f = {d, VISC ? S_FV1 : S_FC1, so}; RUN(launch(c, bi, f)); }
  return 0;
}
extern "C" int vfs_convection(vfs_ctx *c) { if (!c) return VFS_ERR_ARG; RUN(legacy_term<false>(c)); return vfs_sync(c); }
extern "C" int vfs_viscous(vfs_ctx *c) { if (!c) return VFS_ERR_ARG; RUN(legacy_term<true>(c)); return vfs_sync(c); }

// ---- FormFunction_SNES -----------------------------------------------------------------------------------
struct ZeroNormal {   // wall-normal zeroing applied to an Ucont that is already on the device
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    long p = d.idx(i, j, k);
    double tmp[3] = {d.s[S_UC0][p], d.s[S_UC1][p], d.s[S_UC2][p]};
    const int mx = d.mx, my = d.my, mz = d.mz, kg = d.kglob(k);
    const bool jin = (j != 0 && j != my - 1), kin = (kg != 0 && kg != mz - 1), iin = (i != 0 && i != mx - 1);
    if ((i == 0 && d.bc[0] == 1) || (i == mx - 2 && d.bc[1] == 1)) tmp[0] = 0;
    if (d.bc[0] == 10 && i == 0 && jin && kin) tmp[0] = 0;
    if (d.bc[1] == 10 && i == mx - 2 && jin && kin) tmp[0] = 0;
    if ((j == 0 && d.bc[2] == 1) || (j == my - 2 && d.bc[3] == 1)) tmp[1] = 0;
    if (j == my - 2 && (d.bc[3] == 2 || d.bc[3] == 12)) tmp[1] = 0;
    if (j == 0 && d.bc[2] == 12) tmp[1] = 0;
    if (d.bc[2] == 10 && j == 0 && iin && kin) tmp[1] = 0;
    if ((d.bc[3] == 10 || d.bc[3] == -10) && j == my - 2 && iin && kin) tmp[1] = 0;
    if ((kg == 0 && d.bc[4] == 1) || (kg == mz - 2 && d.bc[5] == 1)) tmp[2] = 0;
    d.s[S_UC0][p] = tmp[0]; d.s[S_UC1][p] = tmp[1]; d.s[S_UC2][p] = tmp[2];
  }
};
// wall-normal flux zeroing (momentum.c:2264-2289) of an Ucont already on the device: only the planes
// whose boundary type asks for it are touched
// ghosts: also on the k ghost planes that image another rank's planes (or the periodic seam), so that the
// ucont ghosts stay what their owners hold and need no exchange afterwards (vfs_rhs_les_fused)
static int zero_normal(vfs_ctx *c, bool ghosts = false) {
  const VfsDev &d = c->d;
  ZeroNormal f = {d};
  const int ka = ghosts && (d.kofs > 0 || d.perz) ? -VFS_G : 0, kb = ghosts && (d.kofs + d.nzl < d.mz || d.perz) ? d.nzl + VFS_G : d.nzl;
  const int *bc = d.bc;
  const bool need[6] = {bc[0] == 1 || bc[0] == 10, bc[1] == 1 || bc[1] == 10, bc[2] == 1 || bc[2] == 12 || bc[2] == 10,
                        bc[3] == 1 || bc[3] == 2 || bc[3] == 12 || bc[3] == 10 || bc[3] == -10, bc[4] == 1, bc[5] == 1};
  const int plane[6] = {0, d.mx - 2, 0, d.my - 2, 0, d.mz - 2};
  for (int q = 0; q < 6; q++) {
    if (!need[q]) continue;
    Box b = box_owned(c);
    b.k0 = ka; b.k1 = kb;
    if (q < 2) { b.i0 = plane[q]; b.i1 = b.i0 + 1; }
    else if (q < 4) { b.j0 = plane[q]; b.j1 = b.j0 + 1; }
    else { const int k = plane[q] - d.kofs; if (k < ka || k >= kb) continue; b.k0 = k; b.k1 = k + 1; }       // (non-periodic k: no seam images)
    RUN(launch(c, b, f));
  }
  return 0;
}
// uc_ghosts_fresh: the k ghost planes of ucont already hold what their owners hold (vfs_rhs_les_fused: exchanged
// at the start of the unit, every change since then replayed on the ghost planes) -> only the local wrap fill
static int snes_core(vfs_ctx *c, bool uc_ghosts_fresh = false) {
  if (uc_ghosts_fresh && c->prm.nranks > 1) RUN(wrap_ij(c, grp(S_UC0, 3), -VFS_G, c->d.nzl + VFS_G));
  else RUN(g2l(c, grp(S_UC0, 3)));                                    // momentum.c:2293-2294
  RUN(contra2cart(c));
  RUN(ib_bc(c));
  return formfunction2(c, 1, S_R0, 0.5);
}
extern "C" int vfs_formfunction_snes_dev(vfs_ctx *c) {
  if (!c) return VFS_ERR_ARG;
  ev_rec(c, 2 * VFS_T_TOTAL);
  RUN(run_graphed(c, 0, [&]() -> int {
    RUN(zero_normal(c));
    return snes_core(c);
  }));
  ev_rec(c, 2 * VFS_T_TOTAL + 1);
  return vfs_sync(c);
}
extern "C" int vfs_formfunction_snes(vfs_ctx *c, const double *x, double *fout) {
  if (!c || !x || !fout) return VFS_ERR_ARG;
  ev_rec(c, 2 * VFS_T_TOTAL);
#ifndef VFS_EMU
  {   // X travels on its own stream into its own staging buffer: the copy does not queue behind kernels of earlier
      // (asynchronous) calls; the main stream waits for it before unpacking
    const size_t n = (size_t)c->d.nzl * c->d.my * c->d.mx * 3 * sizeof(double);
    if (!c->up_stream) {
      CK(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
    }
    if (!c->stage_x) CK(cudaMalloc((void **)&c->stage_x, n));
    CK(cudaMemcpyAsync(c->stage_x, x, n, cudaMemcpyHostToDevice, c->up_stream));
    CK(cudaEventRecord(c->ev_up, c->up_stream));
    CK(cudaStreamWaitEvent(c->stream, c->ev_up, 0));
    UnpackX f = {c->d, c->stage_x}; RUN(launch(c, box_owned(c), f));
  }
#else
  RUN(h2d_stage(c, x, 3));
  { UnpackX f = {c->d, c->stage}; RUN(launch(c, box_owned(c), f)); }
#endif
  RUN(snes_core(c));
  { PackAoS f = {c->d, c->stage, S_R0, 3}; RUN(launch(c, box_owned(c), f)); }
  ev_rec(c, 2 * VFS_T_TOTAL + 1);
  return d2h_stage(c, fout, 3);
}

// ---- LES ---------------------------------------------------------------------------------------------------
static bool les2_march_ok(const vfs_ctx *c) {
#ifndef VFS_EMU
  return c->tma_ok;
#else
  (void)c; return true;
#endif
}
// les.c:798-965: Cs from LM, MM averaged over the homogeneous direction(s); see HomoApply (vfs_les_kernels.h)
static int les_homo(vfs_ctx *c) {
  const VfsDev &d = c->d;
  const int mode = d.homo - 1;
  const Box bi = box_interior(c);
  const int nline = mode == 0 ? d.my : (mode == 1 ? d.nzl * d.my : (mode == 2 ? d.nzl * d.mx : d.my * d.mx));
  const size_t need = 3 * (size_t)nline;
  if (c->homo_doubles < need) {
#ifndef VFS_EMU
    if (c->capturing) { set_err(c, "the buffer of the homogeneous Cs averaging must exist before graph capture: run the step eagerly once"); return VFS_ERR_CUDA; }
    if (c->homo_buf) cudaFree(c->homo_buf);
    c->homo_buf = nullptr;
    CK(cudaMalloc((void **)&c->homo_buf, need * sizeof(double)));
#else
    free(c->homo_buf); c->homo_buf = (double *)malloc(need * sizeof(double));
#endif
    c->homo_doubles = need;
  }
#ifndef VFS_EMU
  if (mode <= 1) k_homo_block<<<nline, 256, 0, c->stream>>>(d, mode, bi.k0, bi.k1, c->homo_buf);
  else k_homo_thread<<<(nline + 255) / 256, 256, 0, c->stream>>>(d, mode, nline, bi.k0, bi.k1, c->homo_buf);
  c->launches++;
  CK(cudaGetLastError());
  if (c->prm.nranks > 1 && (mode == 0 || mode == 3)) {        // the sums run across the k-slabs (les.c:822-824, 911-913)
    NcclApi &N = nccl_api();
    if (!c->comm || !N.AllReduce) { set_err(c, "homogeneous Cs averaging across ranks needs vfs_nccl_init (ncclAllReduce)"); return VFS_ERR_HALO; }
    ncclResult_t e = N.AllReduce(c->homo_buf, c->homo_buf, need, ncclDouble, ncclSum, c->comm, c->stream);
    if (e != ncclSuccess) { set_err(c, std::string("ncclAllReduce: ") + N.GetErrorString(e)); return VFS_ERR_HALO; }
  }
#else
  if (c->prm.nranks > 1 && (mode == 0 || mode == 3)) { set_err(c, "the host emulation has no all-reduce: homogeneous averaging across ranks is CUDA / NCCL only"); return VFS_ERR_UNSUPPORTED; }
  for (int l = 0; l < nline; l++) homo_line_serial(d, mode, l, bi.k0, bi.k1, c->homo_buf + 3 * l);
#endif
  HomoApply f = {d, mode, c->homo_buf};
  return launch(c, bi, f);
}
// defer_refresh: leave the ghost refresh of Cs (les.c:1026-1057) to les_nut(c, true), which does it together
// with nu_t's (nu_t of an interior cell reads the cell's own Cs only, les.c:1206)
static int les_cs(vfs_ctx *c, bool defer_refresh = false, bool *nut_done = nullptr) {
  const VfsDev &d = c->d;
  Box all = {-VFS_G, d.mx + VFS_G, -VFS_G, d.my + VFS_G, -VFS_G, d.nzl + VFS_G};
  c->sabs_valid = false;
  if (d.ti < 2 && d.tistart == 0 && !d.rstart_flg) { FillScalar f = {d, S_CS, 0.0}; return launch(c, all, f); }   // les.c:77-80
  if (d.les == 1) { FillScalar f = {d, S_CS, 0.01}; return launch(c, all, f); }                                  // les.c:82-85
  RUN(ensure_iaj(c));
  RUN(ensure_near(c));
  if (d.clark) RUN(ensure_tail(c, S_GR0));
  ev_rec(c, 2 * VFS_T_LES1);
  // Between ranks the 13 pass-1 fields are NOT exchanged: pass 1 is replayed on the ghost planes pass 2 reads — plane -1
  // / nzl across an interior slab boundary, and across the periodic seam plane -2 / nzl+1 (the images of global
  // planes mz-2 / 1, the sources of the seam's node copies) — from the ucat ghosts Contra2Cart replays, the metric and
  // nvert ghosts: same inputs, same code as the owner, bitwise the owner's values (as for Contra2Cart, section 7).
  const bool multi = c->prm.nranks > 1;
  const bool replay = multi && c->les_replay;
  std::vector<std::pair<int, int> > k_ranges;             // pass-1 plane ranges
  { const Box bi = box_interior(c);
    int k0 = bi.k0, k1 = bi.k1;
    if (replay && d.kofs > 0) k0 = -1;
    if (replay && d.kofs + d.nzl < d.mz) k1 = d.nzl + 1;
    k_ranges.push_back(std::make_pair(k0, k1));
    if (replay && d.perz && d.kofs == 0) k_ranges.push_back(std::make_pair(-2, -1));
    if (replay && d.perz && d.kofs + d.nzl == d.mz) k_ranges.push_back(std::make_pair(d.nzl + 1, d.nzl + 2)); }
  const int ka = replay && (d.kofs > 0 || d.perz) ? -2 : 0, kb = replay && (d.kofs + d.nzl < d.mz || d.perz) ? d.nzl + 2 : d.nzl;
  for (size_t q = 0; q < k_ranges.size(); q++) {
    const int q0 = k_ranges[q].first, q1 = k_ranges[q].second;
    if (d.clark) { LesGradStore f = {d, 0}; Box b = box_interior(c); b.k0 = q0; b.k1 = q1; RUN(launch(c, b, f)); }
    if (c->fused && !d.testfilter_ik) {
      int r;
#ifndef VFS_EMU
      if (c->les1_var == 1 && c->tma_ok) r = launch_les1_tma(c->stream, c->tmap, d, q0, q1, &c->launches);
      else if (c->les1_var == 2) { Les1March8 prog = {d}; r = run_filter_march<Les1March8, 2>(c->stream, prog, q0, q1, &c->launches); }
      else if (c->les1_var == 3) { Les1March prog = {d}; r = run_filter_march<Les1March, 2>(c->stream, prog, q0, q1, &c->launches); }
      else
#endif
      { Les1March prog = {d}; r = run_filter_march<Les1March, 1>(c->stream, prog, q0, q1, &c->launches); }
      if (r) { set_err(c, "les1 march kernel launch failed"); return VFS_ERR_CUDA; }
    } else { LesPass1 f = {d}; Box b = box_interior(c); b.k0 = q0; b.k1 = q1; RUN(launch(c, b, f)); }
  }
  ev_rec(c, 2 * VFS_T_LES1 + 1);
  c->sabs_valid = true;
  Grp g1 = grp_cat(grp(S_UF0, 3), grp(S_LW, 10));
  const Grp g1c = grp_cat(grp(S_UF0, 3), grp(S_LU0, 9));
  const Grp ggr = grp(S_GR0, 9);
  if (replay) {
    { LesDeriveBoundary f = {d, (c->les_bnd_valid && c->halo_trim) ? 1 : 0}; c->les_bnd_valid = true; RUN(launch_shell(c, ka, kb, f, true)); }
    RUN(wrap_ij(c, g1, ka, kb));                                        // les.c:254-267 without the k exchange
    if (any_per(c)) { NodeCopy f = {d, g1c, 1}; RUN(launch_shell(c, ka, kb, f, true, SHELL_PERIODIC_ONLY)); }     // les.c:275-306
    if (d.clark) {
      { LesGradStore f = {d, 1}; RUN(launch_shell(c, ka, kb, f, true)); }
      RUN(wrap_ij(c, ggr, ka, kb));
      if (any_per(c)) { NodeCopy f = {d, ggr, 1}; RUN(launch_shell(c, ka, kb, f, true, SHELL_PERIODIC_ONLY)); }
    }
  } else {
    if (d.clark) {
      { LesGradStore f = {d, 1}; RUN(launch_shell(c, 0, d.nzl, f)); }
      RUN(g2l(c, ggr, 2, 2));
      if (any_per(c)) RUN(node_copy(c, ggr));
    }
    RUN(run_les_derive_boundary(c));
    RUN(g2l(c, g1, 2, 2));                                              // les.c:254-267 (pass 2 reads k+-1, the seam copies k+-2)
    // the weight w is a function of the node's own nvert/aj (get_weight, les.c:31-40) and is NOT
    // periodic-copied by the reference; only the copied fields (ucat_f, grad u, |S|) are
    if (any_per(c)) RUN(node_copy(c, g1c));                             // les.c:275-306
  }
  ev_rec(c, 2 * VFS_T_LES2);
  if (c->fused && !d.testfilter_ik && !d.clark && les2_march_ok(c)) {
    Box bi = box_interior(c);
    if (!c->lesgeo_valid) { LesGeo f = {d}; RUN(launch(c, bi, f)); c->lesgeo_valid = true; }
    int r;
#ifndef VFS_EMU
    if (c->les2_ty == 8) { Les2March8 prog = {d}; r = run_les2_march(c->stream, c->tmap_les2_8, c->tmap_les2i_8, prog, bi.k0, bi.k1, &c->launches); }
    else if (c->les2_ty == 12) { Les2March12 prog = {d}; r = run_les2_march(c->stream, c->tmap_les2_12, c->tmap_les2i_12, prog, bi.k0, bi.k1, &c->launches); }
    else { Les2March prog = {d}; r = run_les2_march(c->stream, c->tmap_les2, c->tmap_les2i, prog, bi.k0, bi.k1, &c->launches); }
#else
    if (c->les2_ty == 8) { Les2March8 prog = {d}; r = run_les2_march(c->stream, prog, bi.k0, bi.k1, &c->launches); }
    else if (c->les2_ty == 12) { Les2March12 prog = {d}; r = run_les2_march(c->stream, prog, bi.k0, bi.k1, &c->launches); }
    else { Les2March prog = {d}; r = run_les2_march(c->stream, prog, bi.k0, bi.k1, &c->launches); }
#endif
    if (r) { set_err(c, "les2 march kernel launch failed"); return VFS_ERR_CUDA; }
  } else
  { LesPass2 f = {d}; RUN(launch(c, box_interior(c), f)); }
  ev_rec(c, 2 * VFS_T_LES2 + 1);
#ifndef VFS_EMU
  if (c->fork_after_les2) CK(cudaEventRecord(c->ev_fork2, c->stream));      // ucat is not read again in the LES block
#endif
  Grp g2 = grp(S_LM, 2);
  RUN(g2l(c, g2, 2, 2));                                              // les.c:675-678
  if (any_per(c)) RUN(node_copy(c, g2));
  ev_rec(c, 2 * VFS_T_LES3);
  if (d.homo) { /* no test filter of LM, MM with homogeneous averaging (les.c:776-779): Cs comes from les_homo below */ }
  else if (c->fused && !d.testfilter_ik) {
    Box bi = box_interior(c);
    SideScope sc; RUN(side_begin(c, &sc));       // the thin slabs next to periodic planes run beside the marching kernel
    LesPass3 f = {d};      // cells next to a periodic plane (ghost-image fetches): thin slabs
    if (d.perx) { Box b0 = bi, b1 = bi; b0.i1 = 2; b1.i0 = d.mx - 2; RUN(launch(c, b0, f)); RUN(launch(c, b1, f)); }
    if (d.pery) { Box b0 = bi, b1 = bi; b0.j1 = 2; b1.j0 = d.my - 2; RUN(launch(c, b0, f)); RUN(launch(c, b1, f)); }
    if (d.perz) {
      Box b0 = bi, b1 = bi; b0.k0 = klo(c, 1); b0.k1 = klo(c, 2); b1.k0 = klo(c, d.mz - 2); b1.k1 = klo(c, d.mz - 1);
      RUN(launch(c, b0, f)); RUN(launch(c, b1, f));
    }
    side_end(c, &sc);
    // vfs_rhs_les_fused (option 21): the marching kernel also writes nu_t of its cells; les_nut then only does the thin slabs
    const bool with_nut = nut_done && c->nut_in_les3 && c->sabs_valid && c->lesgeo_valid;
    if (nut_done) *nut_done = with_nut;
    Les3March prog = {d, with_nut ? 1 : 0};
    if (c->les3_minb == 3 ? run_filter_march<Les3March, 3>(c->stream, prog, bi.k0, bi.k1, &c->launches) : c->les3_minb != 2 ? run_filter_march<Les3March, 4>(c->stream, prog, bi.k0, bi.k1, &c->launches)
                          : run_filter_march<Les3March, 2>(c->stream, prog, bi.k0, bi.k1, &c->launches)) { set_err(c, "les3 march kernel launch failed"); return VFS_ERR_CUDA; }
    RUN(ovl_join(c));
  } else
  { LesPass3 f = {d}; RUN(launch(c, box_interior(c), f)); }
  ev_rec(c, 2 * VFS_T_LES3 + 1);
  if (d.homo) RUN(les_homo(c));                                         // les.c:798-965
  { LesClipBoundary f = {d}; RUN(launch_shell(c, 0, d.nzl, f)); }     // les.c:967-980 (boundary nodes; interior clip is in pass 3)
  if (defer_refresh) return 0;
  Grp g3 = grp(S_CS, 1);
  RUN(g2l(c, g3, 2, 2));                                              // les.c:1026-1027
  if (any_per(c)) RUN(node_copy(c, g3));
  return 0;
}
// interior_done: LES pass 3 of this unit already wrote nu_t of every cell its marching kernel covers (Les3March::with_nut);
// what is left are the cells next to periodic planes, which pass 3 does on thin slabs
static int les_nut(vfs_ctx *c, bool with_cs = false, bool interior_done = false) {
  const VfsDev &d = c->d;
  ev_rec(c, 2 * VFS_T_NUT);
  if (interior_done) {
    const Box bi = box_interior(c);
    NuT<true> f = {d};
    if (d.perx) { Box b0 = bi, b1 = bi; b0.i1 = 2; b1.i0 = d.mx - 2; RUN(launch(c, b0, f)); RUN(launch(c, b1, f)); }
    if (d.pery) { Box b0 = bi, b1 = bi; b0.j1 = 2; b1.j0 = d.my - 2; RUN(launch(c, b0, f)); RUN(launch(c, b1, f)); }
    if (d.perz) {
      Box b0 = bi, b1 = bi; b0.k0 = klo(c, 1); b0.k1 = klo(c, 2); b1.k0 = klo(c, d.mz - 2); b1.k1 = klo(c, d.mz - 1);
      RUN(launch(c, b0, f)); RUN(launch(c, b1, f));
    }
  }
  else if (c->sabs_valid) { NuT<true> f = {d}; RUN(launch(c, box_interior(c), f)); }
  else { NuT<false> f = {d}; RUN(launch(c, box_interior(c), f)); }
  ev_rec(c, 2 * VFS_T_NUT + 1);
  Grp g = with_cs ? grp_cat(grp(S_CS, 1), grp(S_NUT, 1)) : grp(S_NUT, 1);
  RUN(g2l(c, g, 2, 2));                                               // les.c:1320-1321 (+ 1026-1027)
  if (any_per(c)) RUN(node_copy(c, g));
  return 0;
}
extern "C" int vfs_les_cs(vfs_ctx *c) { if (!c) return VFS_ERR_ARG; RUN(les_cs(c)); return api_end(c); }
extern "C" int vfs_les_nut(vfs_ctx *c) { if (!c) return VFS_ERR_ARG; RUN(les_nut(c)); return api_end(c); }

// one cell-update unit: Flow_Solver's LES block (solvers.c:365-371) followed by one
// residual evaluation
extern "C" int vfs_rhs_les_fused(vfs_ctx *c) {
  if (!c) return VFS_ERR_ARG;
  ev_rec(c, 2 * VFS_T_TOTAL);
  RUN(run_graphed(c, 1, [&]() -> int {
    RUN(g2l(c, grp(S_UC0, 3)));
    RUN(contra2cart(c));
#ifndef VFS_EMU
    // Single rank, dynamic model: after LES pass 2 nothing in the LES block reads ucat or ucont any more (pass 3 filters
    // LM / MM, nu_t uses the |S| pass 1 stored), so the residual's wall-normal zeroing + Contra2Cart + IB_BC — which
    // overwrite them — run on a second stream beside pass 3 / nu_t and their ghost refreshes (two branches of the graph),
    // joined before the face fluxes, the first reader of nu_t.  Not between ranks (two NCCL exchanges would be in flight on
    // one communicator), and not when IB_BC's first-step wall marking would rewrite nvert under pass 3.
    const VfsDev &d = c->d;
    bool wallfn_any = false;
    for (int q = 0; q < 6; q++) wallfn_any = wallfn_any || d.bc[q] == -1 || d.bc[q] == -2;
    const bool dynamic = d.les == 2 && !(d.ti < 2 && d.tistart == 0 && !d.rstart_flg);
    if (c->unit_overlap && c->prm.nranks == 1 && dynamic && !wallfn_any && c->side2) {
      c->fork_after_les2 = true;
      bool nd = false;
      int r = les_cs(c, true, &nd);
      c->fork_after_les2 = false;
      if (r) return r;
      RUN(les_nut(c, true, nd));
      CK(cudaStreamWaitEvent(c->side2, c->ev_fork2, 0));
      cudaStream_t main_stream = c->stream;
      c->stream = c->side2;
      r = zero_normal(c, true);
      if (!r) r = g2l(c, grp(S_UC0, 3));
      if (!r) r = contra2cart(c);
      if (!r) r = ib_bc(c);
      c->stream = main_stream;
      if (r) return r;
      CK(cudaEventRecord(c->ev_join2, c->side2));
      CK(cudaStreamWaitEvent(c->stream, c->ev_join2, 0));
      return formfunction2(c, 1, S_R0, 0.5);
    }
#endif
    if (c->d.les) { bool nd = false; RUN(les_cs(c, true, &nd)); RUN(les_nut(c, true, nd)); }     // Cs and nu_t ghosts refreshed together
    RUN(zero_normal(c, true));
    return snes_core(c, true);
  }));
  ev_rec(c, 2 * VFS_T_TOTAL + 1);
  return vfs_sync(c);
}

#include "vfs_solver.h"
