#!/usr/bin/env python3
"""Generate the committed golden fixtures from the reference's own sources (oracle/_ref).

Run in the build container (needs /root/reference to have built oracle/_ref/libvfsref.so):
    python tests/golden/make_golden.py
Each fixture stores the seeded case description (inputs are regenerated from the seed by
vfs-wind_b200/cases.py, given the stored metrics) and the reference outputs of every hot-path
function, as float64.  They pin the oracle restatement (tests/test_cpu_oracle_port.py) and the
CUDA path on machines where oracle/_ref is absent."""
import json
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import parity_common as pc  # noqa: E402
import refdrv  # noqa: E402

CASES = {"c2_small": ("c2_box256", (13, 11, 15)), "c3_small": ("c3_turbine", (21, 17, 19)), "c1_small": ("c1_test10", (14, 10, 12))}


def reference_outputs(cfg):
    ref, xyz, fields, met = pc.ref_setup(cfg, refdrv)
    out = {"xyz": xyz, "csi": met["csi"], "eta": met["eta"], "zet": met["zet"], "aj": met["aj"]}
    ref.Contra2Cart()
    out["ucat"] = np.array(ref.owned("Ucat"))
    ref.Compute_Smagorinsky_Constant_1()
    out["cs"] = np.array(ref.owned("lCs"))
    ref.Compute_eddy_viscosity_LES()
    out["nu_t"] = np.array(ref.owned("lNu_t"))
    ref.new_vec("Conv", 3, False); ref.new_vec("Visc", 3, False)       # legacy Convection / Viscous (rhs.c:751, 1071)
    ref.Convection("Conv"); ref.Viscous("Visc")
    out["conv"] = np.array(ref.view("Conv")) * pc.conv_defined(cfg, fields)
    out["visc"] = np.array(ref.view("Visc"))
    ref.IB_BC()
    out["ucont_after_ibbc"] = np.array(ref.owned("lUcont"))
    ref.view("RHS_o")[...] = 0
    ref.Formfunction_2("RHS_o", 1.0)
    out["formfunction2"] = np.array(ref.owned("RHS_o"))
    ref.set_owned("RHS_o", fields["rhs_o"])
    x = pc.krylov_x(fields["ucont"])
    ref.new_vec("X", 3, False); ref.new_vec("F", 3, False)
    ref.view("X")[...] = x
    ref.FormFunction_SNES("X", "F")
    out["snes_f"] = np.array(ref.view("F"))
    out["snes_ucat"] = np.array(ref.owned("Ucat"))
    return out


def main():
    for name, (cfgname, dims) in CASES.items():
        pkg = pc.load_package()
        cfg = pkg.cases.scaled(pkg.cases.CONFIGS[cfgname], *dims)
        out = reference_outputs(cfg)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), cfgname=cfgname, dims=np.array(dims), **out)
        print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
