"""Shared driver of the Newton-Krylov parity tests: vfs_momentum_solve (device-resident
GMRES + MFFD + trust-region Newton on the library's residual) against the numpy restatement of the same PETSc 3.1
algorithms (oracle/newton_krylov_ref.py) driving the ORACLE's FormFunction_SNES."""
import numpy as np
import parity_common as pc
import newton_krylov_ref as nk


def run_solver_parity(cfg, refdrv, lib=None, device=0, **kw):
    ref, xyz, fields, met = pc.ref_setup(cfg, refdrv)
    ref.new_vec("X", 3, False); ref.new_vec("F", 3, False)
    # The synthetic RHS_o / dP / F_eul are random everywhere, also on boundary nodes and masked components, where
    # the residual does not depend on U (momentum.c:1833-1938, 2322-2329): a constant part no Newton step can remove.
    # As in a real run they are zeroed there: the mask is the zero pattern of the residual without them.
    for nm in ("RHS_o", "dP", "F_eul"):
        ref.view(nm)[...] = 0
    ref.Contra2Cart(); ref.Compute_Smagorinsky_Constant_1(); ref.Compute_eddy_viscosity_LES()
    ref.view("X")[...] = fields["ucont"]
    ref.FormFunction_SNES("X", "F")
    live = (np.array(ref.view("F")) != 0).astype(float)
    for key, nm in (("rhs_o", "RHS_o"), ("dp", "dP"), ("f_eul", "F_eul")):
        fields[key] = fields[key] * live
        ref.set_owned(nm, fields[key])
    ref.set_owned("Ucont", fields["ucont"]); ref.global_to_local("Ucont", "lUcont")
    ref.set_owned("Ucat", fields["ucat"]); ref.global_to_local("Ucat", "lUcat")
    ctx = pc.dev_setup(cfg, xyz, fields, lib=lib, device=device)
    for d in (ref, ctx):
        d.Contra2Cart(); d.Compute_Smagorinsky_Constant_1(); d.Compute_eddy_viscosity_LES()

    def residual(x):
        ref.view("X")[...] = x
        ref.FormFunction_SNES("X", "F")
        return np.array(ref.view("F"))
    u_ref, info_ref = nk.snes_tr(residual, fields["ucont"], max_newton=kw.get("max_newton", 50), restart=kw.get("restart", 30),
                                 snes_rtol=kw.get("rtol", 1e-8), ksp_rtol=kw.get("ksp_rtol", 1e-5), use_ew=kw.get("use_ew", 1), trust_region=kw.get("trust_region", 1))
    ctx.upload("UCONT", fields["ucont"])           # the global Ucont Vec (Contra2Cart rewrote lUcont's periodic boundary nodes)
    info = ctx.momentum_solve(max_newton=kw.get("max_newton"), restart=kw.get("restart"), rtol=kw.get("rtol"), ksp_rtol=kw.get("ksp_rtol"), use_ew=kw.get("use_ew"), trust_region=kw.get("trust_region"))
    u_dev = ctx.download("UCONT")
    ctx.close()
    return u_dev, info, u_ref, info_ref, fields


def check(u_dev, info, u_ref, info_ref, fields, tol=1e-10):
    assert info["reason"] == info_ref["reason"], (info, info_ref)
    assert info["ksp_its_history"] == info_ref["ksp_its_history"], (info, info_ref)
    assert info["residual_evals"] == info_ref["residual_evals"], (info, info_ref)
    h0, h1 = np.array(info["fnorm_history"]), np.array(info_ref["fnorm_history"])
    assert h0.shape == h1.shape
    # the residual norms fall by many orders of magnitude: each is compared relative to the initial one
    assert np.abs(h0 - h1).max() <= tol * h1[0], (h0, h1)
    assert h1[-1] < 1e-3 * h1[0], h1                 # the solve did something
    # iterates: the update U - U0 to 1e-10 of its own size (and U itself far tighter)
    du_dev, du_ref = u_dev - fields["ucont"], u_ref - fields["ucont"]
    assert pc.relerr(du_dev, du_ref) <= 1e-8, pc.relerr(du_dev, du_ref)
    assert pc.relerr(u_dev, u_ref) <= tol, pc.relerr(u_dev, u_ref)
