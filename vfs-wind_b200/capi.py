"""ctypes binding of the C ABI in include/vfs_b200.h (libvfs_b200.so, hand-written sm_100a CUDA).

Host-side mirror of the reference's operator interface for the momentum RHS + LES path: method
names follow the reference functions they replace (FormMetrics, Contra2Cart, IB_BC,
Formfunction_2, FormFunction_SNES, Compute_Smagorinsky_Constant_1, Compute_eddy_viscosity_LES;
Source/variables.h:608-609,628,699-700,1196-1197,1229).  There is no CPU fallback: if the CUDA
library is missing or no GPU is usable, construction raises.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvfs_b200.so")

FIELDS = ["COOR", "CSI", "ETA", "ZET", "AJ", "NVERT", "UCONT", "UCAT", "UCAT_OLD", "UCONT_O", "UCONT_RM1", "RHS_O", "DP", "F_EUL",
          "RHS", "CS", "NU_T", "USTAR", "CONV", "VISC", "P", "PHI"]
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}
FIELD_DOF = {"COOR": 3, "CSI": 3, "ETA": 3, "ZET": 3, "AJ": 1, "NVERT": 1, "UCONT": 3, "UCAT": 3, "UCAT_OLD": 3, "UCONT_O": 3,
             "UCONT_RM1": 3, "RHS_O": 3, "DP": 3, "F_EUL": 3, "RHS": 3, "CS": 1, "NU_T": 1, "USTAR": 1, "CONV": 3, "VISC": 3, "P": 1, "PHI": 1}

EXPORTS = ["vfs_create", "vfs_destroy", "vfs_last_error", "vfs_set_params", "vfs_set_stream", "vfs_set_halo_callback", "vfs_sync",
           "vfs_nccl_unique_id", "vfs_nccl_init", "vfs_halo_count", "vfs_layout", "vfs_field_scalar_id", "vfs_scalar_ptr", "vfs_upload", "vfs_download", "vfs_halo_exchange",
           "vfs_form_metrics", "vfs_contra2cart", "vfs_ib_bc", "vfs_les_cs", "vfs_les_nut", "vfs_formfunction2", "vfs_convection", "vfs_viscous", "vfs_pressure_gradient", "vfs_download_async", "vfs_download_wait",
           "vfs_formfunction_snes", "vfs_formfunction_snes_dev", "vfs_rhs_les_fused", "vfs_launch_count", "vfs_last_ms",
           "vfs_calc_f_eul", "vfs_calc_u_lagr", "vfs_solver_defaults", "vfs_momentum_solve", "vfs_momentum_release", "vfs_set_option", "vfs_halo_layers", "vfs_host_alloc", "vfs_host_free", "vfs_device_count", "vfs_cylinder_forces", "vfs_update_pressure", "vfs_projection"]


class VfsParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("mx", "my", "mz", "kofs", "nzl", "rank", "nranks", "device", "ii_periodic", "jj_periodic", "kk_periodic")] + \
               [("bctype", C.c_int * 6)] + \
               [(n, C.c_int) for n in ("les", "second_order", "laplacian", "immersed", "clark", "central", "testfilter_ik",
                                       "viscosity_wallmodel", "wallfunction", "rotor_model", "nacelle_model", "IB_delta",
                                       "ti", "tistart", "rstart_flg", "levelset", "rans", "inviscid", "skew", "movefsi", "rotatefsi",
                                       "i_periodic", "j_periodic", "k_periodic", "i_homo_filter", "j_homo_filter", "k_homo_filter")] + \
               [(n, C.c_double) for n in ("ren", "dt", "max_cs", "roughness_size")] + \
               [(n, C.c_int) for n in ("levelset_weno", "freesurface_wallmodel", "air_flow_levelset")]


class VfsSolverParams(C.Structure):
    _fields_ = [("max_newton", C.c_int), ("restart", C.c_int), ("max_krylov", C.c_int),
                ("snes_atol", C.c_double), ("snes_rtol", C.c_double), ("snes_stol", C.c_double),
                ("ksp_rtol", C.c_double), ("ksp_atol", C.c_double), ("ksp_dtol", C.c_double),
                ("use_ew", C.c_int), ("trust_region", C.c_int)]


class VfsSolverInfo(C.Structure):
    _fields_ = [("newton_iterations", C.c_int), ("krylov_iterations", C.c_int), ("residual_evals", C.c_int), ("reason", C.c_int),
                ("fnorm0", C.c_double), ("fnorm", C.c_double), ("xnorm", C.c_double), ("delta", C.c_double),
                ("n_history", C.c_int), ("fnorm_history", C.c_double * 17), ("ksp_its_history", C.c_int * 16)]


class VfsActuator(C.Structure):
    _fields_ = [("n_elmt", C.c_int)] + [(n, C.POINTER(C.c_double)) for n in ("cent_x", "cent_y", "cent_z", "dA", "F_lagr_x", "F_lagr_y", "F_lagr_z", "U_lagr_x", "U_lagr_y", "U_lagr_z")] + \
               [(n, C.POINTER(C.c_int)) for n in ("i_min", "i_max", "j_min", "j_max", "k_min", "k_max")]


HALO_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int))


class VfsError(RuntimeError):
    pass


def _bind(lib):
    lib.vfs_create.argtypes = [C.POINTER(VfsParams), C.POINTER(C.c_void_p)]
    lib.vfs_destroy.argtypes = [C.c_void_p]
    lib.vfs_last_error.argtypes = [C.c_void_p]
    lib.vfs_last_error.restype = C.c_char_p
    lib.vfs_set_params.argtypes = [C.c_void_p, C.POINTER(VfsParams)]
    lib.vfs_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.vfs_set_halo_callback.argtypes = [C.c_void_p, HALO_FN, C.c_void_p]
    lib.vfs_sync.argtypes = [C.c_void_p]
    lib.vfs_nccl_unique_id.argtypes = [C.c_char_p]
    lib.vfs_nccl_init.argtypes = [C.c_void_p, C.c_char_p]
    lib.vfs_halo_count.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
    lib.vfs_halo_count.restype = C.c_long
    lib.vfs_layout.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
    lib.vfs_field_scalar_id.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.vfs_scalar_ptr.argtypes = [C.c_void_p, C.c_int]
    lib.vfs_scalar_ptr.restype = C.c_void_p
    lib.vfs_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.vfs_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    if hasattr(lib, "vfs_download_async"):
        lib.vfs_download_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.vfs_download_wait.argtypes = [C.c_void_p]
    lib.vfs_halo_exchange.argtypes = [C.c_void_p, C.c_int]
    for f in ("vfs_form_metrics", "vfs_contra2cart", "vfs_ib_bc", "vfs_les_cs", "vfs_les_nut", "vfs_formfunction_snes_dev", "vfs_rhs_les_fused", "vfs_convection", "vfs_viscous"):
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.vfs_formfunction2.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.vfs_pressure_gradient.argtypes = [C.c_void_p, C.c_double]
    lib.vfs_cylinder_forces.argtypes = [C.c_void_p, C.c_void_p]
    lib.vfs_update_pressure.argtypes = [C.c_void_p]
    lib.vfs_projection.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.vfs_formfunction_snes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vfs_launch_count.argtypes = [C.c_void_p]
    lib.vfs_launch_count.restype = C.c_long
    lib.vfs_last_ms.argtypes = [C.c_void_p, C.c_int]
    lib.vfs_last_ms.restype = C.c_double
    if hasattr(lib, "vfs_set_option"):
        lib.vfs_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.vfs_calc_f_eul.argtypes = [C.c_void_p, C.c_int, C.POINTER(VfsActuator), C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double), C.c_int]
    lib.vfs_calc_u_lagr.argtypes = [C.c_void_p, C.c_int, C.POINTER(VfsActuator)]
    lib.vfs_halo_layers.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.vfs_solver_defaults.argtypes = [C.POINTER(VfsSolverParams)]
    lib.vfs_momentum_solve.argtypes = [C.c_void_p, C.POINTER(VfsSolverParams), C.POINTER(VfsSolverInfo)]
    lib.vfs_momentum_release.argtypes = [C.c_void_p]
    return lib


_lib = None


def load():
    """Load the CUDA library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VfsError("libvfs_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        _lib = _bind(C.CDLL(LIB_PATH))
    return _lib


def make_params(mx, my, mz, flags, ren, dt, bctype, kofs=0, nzl=None, rank=0, nranks=1, device=0):
    p = VfsParams()
    p.mx, p.my, p.mz = mx, my, mz
    p.kofs, p.nzl = kofs, (mz if nzl is None else nzl)
    p.rank, p.nranks, p.device = rank, nranks, device
    for q in range(6):
        p.bctype[q] = int(bctype[q])
    p.ti, p.tistart, p.max_cs = 10, 0, 0.5
    for k, v in flags.items():
        if k in ("max_cs", "roughness_size"):
            setattr(p, k, float(v))
        elif hasattr(p, k):
            setattr(p, k, int(v))
        elif v:
            raise VfsError("flag %s is not part of the hot-path contract" % k)
    p.ren, p.dt = float(ren), float(dt)
    return p


def slab_partition(mz, nranks):
    """Contiguous k-slabs: global planes [kofs, kofs+nzl) per rank, remainder to the low ranks."""
    base, rem = divmod(mz, nranks)
    out, k = [], 0
    for r in range(nranks):
        n = base + (1 if r < rem else 0)
        out.append((k, n))
        k += n
    return out


class VfsContext:
    """One device context (one GPU / one k-slab)."""

    def __init__(self, params, lib=None):
        self.lib = lib if lib is not None else load()
        self.p = params
        h = C.c_void_p()
        r = self.lib.vfs_create(C.byref(params), C.byref(h))
        if r != 0:
            raise VfsError("vfs_create failed (%d): %s" % (r, self.lib.vfs_last_error(None).decode()))
        self.h = h
        L = (C.c_long * 8)()
        self.lib.vfs_layout(self.h, L)
        self.G, self.pitch, self.ny, self.nzt, self.plane, self.scalar_len, self.nscalars = [int(x) for x in L[:7]]
        self._cb = None

    def close(self):
        if self.h:
            self.lib.vfs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r):
        if r != 0:
            raise VfsError("vfs call failed (%d): %s" % (r, self.lib.vfs_last_error(self.h).decode()))

    @property
    def shape(self):
        return (self.p.nzl, self.p.my, self.p.mx)

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.p, k, v)
        self._ck(self.lib.vfs_set_params(self.h, C.byref(self.p)))

    def set_stream(self, stream_ptr):
        self._ck(self.lib.vfs_set_stream(self.h, C.c_void_p(stream_ptr)))

    def set_halo_callback(self, pyfn):
        """pyfn(list_of_scalar_ids) -> 0 on success."""
        def tramp(user, n, ids):
            try:
                return int(pyfn([ids[q] for q in range(n)]) or 0)
            except Exception as e:  # noqa
                import traceback
                traceback.print_exc()
                return 1
        self._cb = HALO_FN(tramp)
        self._ck(self.lib.vfs_set_halo_callback(self.h, self._cb, None))

    def nccl_init(self, dist, device=None):
        """Collective: set up the in-library NCCL halo layer.  `dist` is an initialised
        torch.distributed module, used only to broadcast rank 0's 128-byte ncclUniqueId."""
        import torch
        buf = C.create_string_buffer(128)
        if self.p.rank == 0:
            r = self.lib.vfs_nccl_unique_id(buf)
            if r:
                raise VfsError("vfs_nccl_unique_id failed (%d): %s" % (r, self.lib.vfs_last_error(None).decode()))
        t = torch.tensor(list(buf.raw), dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=0)
        self._ck(self.lib.vfs_nccl_init(self.h, bytes(t.cpu().tolist())))

    def halo_layers(self):
        """(lo, hi) ghost planes the exchange in progress must fill (valid inside a halo callback)."""
        lo, hi = C.c_int(0), C.c_int(0)
        self.lib.vfs_halo_layers(self.h, C.byref(lo), C.byref(hi))
        return int(lo.value), int(hi.value)

    def halo_count(self):
        b = C.c_long(0)
        n = self.lib.vfs_halo_count(self.h, C.byref(b))
        return int(n), int(b.value)

    def scalar_ptr(self, sid):
        return self.lib.vfs_scalar_ptr(self.h, sid)

    def scalar_id(self, field, comp=0):
        return self.lib.vfs_field_scalar_id(self.h, FIELD_ID[field], comp)

    def upload(self, field, arr):
        dof = FIELD_DOF[field]
        a = np.ascontiguousarray(arr, dtype=np.float64)
        want = self.shape + ((3,) if dof == 3 else ())
        if a.shape != want:
            raise VfsError("upload %s: shape %s != %s" % (field, a.shape, want))
        self._ck(self.lib.vfs_upload(self.h, FIELD_ID[field], a.ctypes.data_as(C.c_void_p)))

    def download(self, field):
        dof = FIELD_DOF[field]
        out = np.empty(self.shape + ((3,) if dof == 3 else ()), dtype=np.float64)
        self._ck(self.lib.vfs_download(self.h, FIELD_ID[field], out.ctypes.data_as(C.c_void_p)))
        return out

    def upload_ptr(self, field, host_ptr):
        """Upload from a raw host address (e.g. pinned memory) holding [nzl][my][mx][dof] doubles."""
        self._ck(self.lib.vfs_upload(self.h, FIELD_ID[field], C.c_void_p(host_ptr)))

    def download_ptr(self, field, host_ptr):
        self._ck(self.lib.vfs_download(self.h, FIELD_ID[field], C.c_void_p(host_ptr)))

    def download_async(self, field, host_ptr, slot=0):
        """Start an asynchronous download into the (pinned) host address; valid after download_wait()."""
        self._ck(self.lib.vfs_download_async(self.h, FIELD_ID[field], C.c_void_p(host_ptr), int(slot)))

    def download_wait(self):
        self._ck(self.lib.vfs_download_wait(self.h))

    def halo_exchange(self, field):
        self._ck(self.lib.vfs_halo_exchange(self.h, FIELD_ID[field]))

    # --- reference-named entry points -------------------------------------------------------
    def FormMetrics(self):
        self._ck(self.lib.vfs_form_metrics(self.h))

    def Contra2Cart(self):
        self._ck(self.lib.vfs_contra2cart(self.h))

    def IB_BC(self):
        self._ck(self.lib.vfs_ib_bc(self.h))

    def Convection(self):
        """rhs.c:751 on the current UCONT / UCAT; result in field "CONV"."""
        self._ck(self.lib.vfs_convection(self.h))

    def Viscous(self):
        """rhs.c:1071 on the current UCAT / NU_T; result in field "VISC"."""
        self._ck(self.lib.vfs_viscous(self.h))

    def Pressure_Gradient(self, k_forcing=0.0):
        """momentum.c:203 on the current field "P"; result in field "DP"."""
        self._ck(self.lib.vfs_pressure_gradient(self.h, float(k_forcing)))

    def UpdatePressure(self):
        """poisson.c:3137 on the current fields "P", "PHI", "NVERT"."""
        self._ck(self.lib.vfs_update_pressure(self.h))

    def Projection(self, st=1.0, poisson_threshold=0.1):
        """poisson.c:2700-3025: UCONT -= dt * st * grad(PHI), periodic copies (follow with Contra2Cart, :3049)."""
        self._ck(self.lib.vfs_projection(self.h, float(st), float(poisson_threshold)))

    def cylinder_forces(self):
        """momentum.c:822-849: (A_cyl, A_cyl_x, A_cyl_z, Fpx, Fpz, Fvx, Fvz) of this rank's wall faces (bctype[0] == 11)."""
        out = np.zeros(7)
        self._ck(self.lib.vfs_cylinder_forces(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    @staticmethod
    def _actuators(acts):
        """acts: list of dicts with cent (n,3), F_lagr (n,3), dA (n,), win (n,6) int32 -> (array of VfsActuator, keep-alive)."""
        arr = (VfsActuator * len(acts))()
        keep = []
        for q, a in enumerate(acts):
            n = len(a["dA"])
            cols = [np.ascontiguousarray(a["cent"][:, c], dtype=np.float64) for c in range(3)] + [np.ascontiguousarray(a["dA"], dtype=np.float64)] + \
                   [np.ascontiguousarray(a["F_lagr"][:, c], dtype=np.float64) for c in range(3)] + [np.zeros(n) for _ in range(3)]
            wins = [np.ascontiguousarray(a["win"][:, c], dtype=np.int32) for c in range(6)]
            keep.append((cols, wins))
            arr[q].n_elmt = n
            for name, v in zip(("cent_x", "cent_y", "cent_z", "dA", "F_lagr_x", "F_lagr_y", "F_lagr_z", "U_lagr_x", "U_lagr_y", "U_lagr_z"), cols):
                setattr(arr[q], name, v.ctypes.data_as(C.POINTER(C.c_double)))
            for name, v in zip(("i_min", "i_max", "j_min", "j_max", "k_min", "k_max"), wins):
                setattr(arr[q], name, v.ctypes.data_as(C.POINTER(C.c_int)))
        return arr, keep

    def Calc_F_eul(self, acts, df=10, halfwidth_dfunc=4.0, dh_fixed=None, accumulate=False):
        """rotor_model.c:3668: spread the actuators' forces into field "F_EUL"."""
        arr, keep = self._actuators(acts)
        dh = (C.c_double * 3)(*(dh_fixed or (0, 0, 0)))
        self._ck(self.lib.vfs_calc_f_eul(self.h, len(acts), arr, int(df), float(halfwidth_dfunc), int(dh_fixed is not None), dh, int(accumulate)))

    def Calc_U_lagr(self, acts):
        """rotor_model.c:2937: velocity of field "UCAT" interpolated to the actuator elements; list of (n,3) arrays."""
        arr, keep = self._actuators(acts)
        self._ck(self.lib.vfs_calc_u_lagr(self.h, len(acts), arr))
        return [np.stack(k[0][7:10], -1) for k in keep]

    def Compute_Smagorinsky_Constant_1(self):
        self._ck(self.lib.vfs_les_cs(self.h))

    def Compute_eddy_viscosity_LES(self):
        self._ck(self.lib.vfs_les_nut(self.h))

    def Formfunction_2(self, rhs_field, scale):
        self._ck(self.lib.vfs_formfunction2(self.h, FIELD_ID[rhs_field], float(scale)))

    def FormFunction_SNES(self, x, f=None):
        """x: host (nzl,my,mx,3) array; returns F (same shape).  `f` may be a preallocated
        (e.g. pinned) output array; raw integer addresses are accepted for both."""
        if isinstance(x, int):
            self._ck(self.lib.vfs_formfunction_snes(self.h, C.c_void_p(x), C.c_void_p(f)))
            return None
        xa = np.ascontiguousarray(x, dtype=np.float64)
        if f is None:
            f = np.empty_like(xa)
        self._ck(self.lib.vfs_formfunction_snes(self.h, xa.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p)))
        return f

    def FormFunction_SNES_dev(self):
        self._ck(self.lib.vfs_formfunction_snes_dev(self.h))

    def momentum_solve(self, max_newton=None, max_krylov=None, restart=None, rtol=None, atol=None, ksp_rtol=None, use_ew=None, trust_region=None):
        """Implicit_MatrixFree's SNESSolve on the device (Source/implicitsolver.c:4203-4299): VFS_UCONT in/out.
        Defaults are the reference's / PETSc's; returns the vfs_solver_info fields as a dict."""
        sp = VfsSolverParams()
        self.lib.vfs_solver_defaults(C.byref(sp))
        for k, v in (("max_newton", max_newton), ("max_krylov", max_krylov), ("restart", restart), ("snes_rtol", rtol), ("snes_atol", atol),
                     ("ksp_rtol", ksp_rtol), ("use_ew", use_ew), ("trust_region", trust_region)):
            if v is not None:
                setattr(sp, k, v)
        info = VfsSolverInfo()
        self._ck(self.lib.vfs_momentum_solve(self.h, C.byref(sp), C.byref(info)))
        out = {k: getattr(info, k) for k, _ in VfsSolverInfo._fields_ if not k.endswith("history")}
        out["fnorm_history"] = [info.fnorm_history[q] for q in range(info.n_history)]
        out["ksp_its_history"] = [info.ksp_its_history[q] for q in range(max(0, min(16, info.n_history - 1)))]
        return out

    def momentum_release(self):
        self._ck(self.lib.vfs_momentum_release(self.h))

    def rhs_les_fused(self):
        self._ck(self.lib.vfs_rhs_les_fused(self.h))

    def sync(self):
        self._ck(self.lib.vfs_sync(self.h))

    def launch_count(self):
        return int(self.lib.vfs_launch_count(self.h))

    def last_ms(self, which=0):
        return float(self.lib.vfs_last_ms(self.h, which))

    def set_option(self, key, value):
        self.lib.vfs_set_option(self.h, key, value)
