"""ctypes driver for oracle/_ref/libvfsref.so (the reference's own sources behind a PETSc shim).

TEST INFRASTRUCTURE ONLY.  Used by tests/ (as the checker), by tests/golden/make_golden.py (to
generate committed fixtures) and by bench.py's cpu_baseline / --impl reference legs.  The product
path never imports this module.

The reference keeps its run-time switches in ~95 C globals (Source/main.c:21-356); `RefCase`
sets them from a flat dict before creating the single-rank DA, because the DA wrap type depends on
ii/jj/kk_periodic (Source/init.c:145-160).
"""
import ctypes as C
import json
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libvfsref.so")
_GLOBALS_JSON = os.path.join(HERE, "_ref", "globals.json")

_CT = {"int": C.c_int, "PetscInt": C.c_int, "PetscTruth": C.c_int, "double": C.c_double, "PetscReal": C.c_double}
_lib = None
_globals = None


def available():
    return os.path.exists(SO)


def lib():
    global _lib, _globals
    if _lib is None:
        _lib = C.CDLL(SO)
        _globals = json.load(open(_GLOBALS_JSON))
        L = _lib
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int] * 3
        L.ref_vec.restype = C.c_void_p
        L.ref_vec.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_vec_data.restype = C.POINTER(C.c_double)
        L.ref_vec_data.argtypes = [C.c_void_p]
        L.ref_vec_size.restype = C.c_long
        L.ref_vec_size.argtypes = [C.c_void_p]
        L.ref_vec_is_local.argtypes = [C.c_void_p]
        L.ref_vec_dof.argtypes = [C.c_void_p]
        L.ref_local_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ref_set_scalars.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_int)]
        L.ref_global_to_local.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        for f in ("ref_FormMetrics", "ref_Contra2Cart", "ref_IB_BC", "ref_Compute_Smagorinsky_Constant_1",
                  "ref_Compute_eddy_viscosity_LES"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_Formfunction_2.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
        L.ref_FormFunction_SNES.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_Pressure_Gradient.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.ref_actuator_new.restype = C.c_void_p
        L.ref_actuator_new.argtypes = [C.c_int]
        L.ref_actuator_d.restype = C.POINTER(C.c_double)
        L.ref_actuator_d.argtypes = [C.c_void_p, C.c_int]
        L.ref_actuator_i.restype = C.POINTER(C.c_int)
        L.ref_actuator_i.argtypes = [C.c_void_p, C.c_int]
        L.ref_Calc_F_eul.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_Calc_U_lagr.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_UpdatePressure.argtypes = [C.c_void_p]
        L.ref_Projection.argtypes = [C.c_void_p, C.c_double]
        if hasattr(L, "ref_cylinder_forces"):
            L.ref_cylinder_forces.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_Convection.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_Viscous.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_vec_new.restype = C.c_void_p
        L.ref_vec_new.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_vec_free.argtypes = [C.c_void_p]
    return _lib


def set_global(name, value):
    L = lib()
    if name not in _globals:      # a glue build only defines the switches the product reads
        return
    t = _globals[name]
    _CT[t].in_dll(L, name).value = value


def get_global(name):
    L = lib()
    return _CT[_globals[name]].in_dll(L, name).value


# every switch the hot path reads (SURVEY section 5 "Config / flags"); all default 0 as in main.c
FLAG_DEFAULTS = dict(
    les=0, rans=0, levelset=0, levelset_weno=0, inviscid=0, central=0, second_order=0, skew=0, clark=0, laplacian=0,
    immersed=0, i_periodic=0, j_periodic=0, k_periodic=0, ii_periodic=0, jj_periodic=0, kk_periodic=0,
    i_homo_filter=0, j_homo_filter=0, k_homo_filter=0, testfilter_ik=0, max_cs=0.5, wallfunction=0,
    viscosity_wallmodel=0, freesurface_wallmodel=0, movefsi=0, rotatefsi=0, rotor_model=0, nacelle_model=0, IB_delta=0,
    ti=10, tistart=0, rstart_flg=0, wave_momentum_source=0, air_flow_levelset=0, surface_tension=0, lowRe=0,
    roughness_size=0.0, dthick=1.5, forcewidthfixed=0, halfwidth_dfunc=4.0, ii_periodicWT=0, jj_periodicWT=0, kk_periodicWT=0, dpdz_set=0, mean_pressure_gradient=0.0, inletprofile=0, inlet_flux=0.0, my_rank=0, NumberOfBodies=0, block_number=1, averaging=0, poisson_threshold=0.1,
    MoveFrame=0, u_frame=0.0, v_frame=0.0, w_frame=0.0, Nx_WT=1, Ny_WT=1, Nz_WT=1, Sx_WT=1.0, Sy_WT=1.0, Sz_WT=1.0,
)


class RefCase:
    """One single-rank reference context.  Arrays are numpy views into the reference's Vecs with
    shape (nz, ny, nx[, 3]); local (ghosted) Vecs start at index (gzs, gys, gxs) which is -3 in
    DA-wrapped directions: use `view(name)` for the raw ghosted array or `owned(name)` for the
    mx*my*mz owned block."""

    def __init__(self, mx, my, mz, flags, ren, dt, bctype):
        L = lib()
        fl = dict(FLAG_DEFAULTS)
        fl.update(flags)
        fl["periodic"] = sum(fl[k] for k in ("i_periodic", "j_periodic", "k_periodic", "ii_periodic", "jj_periodic", "kk_periodic"))
        for k, v in fl.items():
            set_global(k, v)
        self.flags = fl
        self.mx, self.my, self.mz = mx, my, mz
        self.u = L.ref_create(mx, my, mz)
        bc = (C.c_int * 6)(*bctype)
        L.ref_set_scalars(self.u, ren, dt, bc)
        info = (C.c_int * 6)()
        L.ref_local_info(self.u, info)
        self.gxs, self.gys, self.gzs, self.gxm, self.gym, self.gzm = list(info)
        self._extra = {}

    def vec(self, name):
        if name in self._extra:
            return self._extra[name]
        v = lib().ref_vec(self.u, name.encode())
        if not v:
            raise KeyError(name)
        return v

    def new_vec(self, name, dof, local):
        self._extra[name] = lib().ref_vec_new(self.u, dof, int(local))
        return self._extra[name]

    def view(self, name):
        L = lib()
        v = self.vec(name)
        n = L.ref_vec_size(v)
        dof = L.ref_vec_dof(v)
        loc = L.ref_vec_is_local(v)
        a = np.ctypeslib.as_array(L.ref_vec_data(v), shape=(n,))
        shp = (self.gzm, self.gym, self.gxm) if loc else (self.mz, self.my, self.mx)
        return a.reshape(shp + ((dof,) if dof > 1 else ()))

    def owned(self, name):
        a = self.view(name)
        if lib().ref_vec_is_local(self.vec(name)):
            return a[-self.gzs:-self.gzs + self.mz, -self.gys:-self.gys + self.my, -self.gxs:-self.gxs + self.mx]
        return a

    def set_owned(self, name, arr):
        self.owned(name)[...] = arr

    def wrap_fill(self, name):
        """Fill the DA-wrap ghosts of a local Vec from its owned block (what DALocalToLocal does)."""
        a = self.view(name)
        o = np.array(self.owned(name))
        idx = [np.arange(g, g + n) % m for g, n, m in ((self.gzs, self.gzm, self.mz), (self.gys, self.gym, self.my), (self.gxs, self.gxm, self.mx))]
        a[...] = o[np.ix_(*idx)]

    def set_coords(self, xyz):
        """xyz: (mz, my, mx, 3) node coordinates (entries at index m-1 are unused by the reference)."""
        self.set_owned("coords", xyz)
        self.wrap_fill("coords")

    def global_to_local(self, g, l):
        lib().ref_global_to_local(self.u, g.encode(), l.encode())

    def FormMetrics(self):
        return lib().ref_FormMetrics(self.u)

    def Contra2Cart(self):
        lib().ref_Contra2Cart(self.u)

    def IB_BC(self):
        lib().ref_IB_BC(self.u)

    def Formfunction_2(self, rhs_name, scale):
        return lib().ref_Formfunction_2(self.u, self.vec(rhs_name), scale)

    def FormFunction_SNES(self, x_name, f_name):
        return lib().ref_FormFunction_SNES(self.u, self.vec(x_name), self.vec(f_name))

    def Compute_Smagorinsky_Constant_1(self):
        lib().ref_Compute_Smagorinsky_Constant_1(self.u)

    def Compute_eddy_viscosity_LES(self):
        lib().ref_Compute_eddy_viscosity_LES(self.u)

    def Pressure_Gradient(self, name, mean_k_flux=0.0, mean_k_area=1.0):
        lib().ref_Pressure_Gradient(self.u, self.vec(name), mean_k_flux, mean_k_area)

    def UpdatePressure(self):
        lib().ref_UpdatePressure(self.u)

    def Projection(self, st=1.0):
        lib().ref_Projection(self.u, st)

    def cylinder_forces(self):
        """lA_cyl, lA_cyl_x, lA_cyl_z, lFpx_cyl, lFpz_cyl, lFvx_cyl, lFvz_cyl as the last Formfunction_2 left them."""
        out = np.zeros(7)
        lib().ref_cylinder_forces(self.u, out.ctypes.data_as(C.c_void_p))
        return out

    # actuator elements (IBMNodes subset): `act` = dict of numpy arrays cent (n,3), F_lagr (n,3), dA (n,), win (n,6) int32
    def _actuator(self, act):
        L = lib()
        n = len(act["dA"])
        b = L.ref_actuator_new(n)
        cols = [act["cent"][:, 0], act["cent"][:, 1], act["cent"][:, 2], act["dA"], act["F_lagr"][:, 0], act["F_lagr"][:, 1], act["F_lagr"][:, 2]]
        for q, a in enumerate(cols):
            np.ctypeslib.as_array(L.ref_actuator_d(b, q), shape=(n,))[...] = a
        for q in range(6):
            np.ctypeslib.as_array(L.ref_actuator_i(b, q), shape=(n,))[...] = act["win"][:, q]
        return b, n

    def Calc_F_eul(self, act, df):
        b, n = self._actuator(act)
        return lib().ref_Calc_F_eul(self.u, b, df)

    def Calc_U_lagr(self, act):
        b, n = self._actuator(act)
        lib().ref_Calc_U_lagr(self.u, b)
        return np.stack([np.array(np.ctypeslib.as_array(lib().ref_actuator_d(b, 7 + q), shape=(n,))) for q in range(3)], -1)

    def Calc_U_lagr_multi(self, acts, centres):
        """Several objects (a turbine array) in one call; centres (nobj, 3) -> FSInfo.x_c/y_c/z_c.  Returns a list of (n, 3)."""
        L = lib()
        L.ref_Calc_U_lagr_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        bs = [self._actuator(a) for a in acts]
        arr = (C.c_void_p * len(bs))(*[b for b, _ in bs])
        cen = np.ascontiguousarray(centres, dtype=np.float64)
        L.ref_Calc_U_lagr_multi(self.u, arr, len(bs), cen.ctypes.data_as(C.c_void_p))
        return [np.stack([np.array(np.ctypeslib.as_array(L.ref_actuator_d(b, 7 + q), shape=(n,))) for q in range(3)], -1) for b, n in bs]

    def Convection(self, name):
        return lib().ref_Convection(self.u, self.vec(name))

    def Viscous(self, name):
        return lib().ref_Viscous(self.u, self.vec(name))
