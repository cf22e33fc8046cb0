"""vfs-wind_b200: B200-native momentum RHS + LES path of VFS-Wind.

The directory name contains a hyphen (it mirrors the reference's repository name), so it is
loaded by path: `from __graft_entry__ import load_package; pkg = load_package()`.
"""
from . import capi, cases, petsc_io, selfcheck  # noqa: F401


def __getattr__(name):
    if name in ("halo", "cases_device"):            # import torch; only needed for multi-GPU runs / large on-device cases
        import importlib
        return importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)
