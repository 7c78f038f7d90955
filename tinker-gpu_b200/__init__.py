"""B200-native AMOEBA polarizable electrostatics: host side of the drop-in for
Tinker-GPU's energy()/induce()/epolar() path.

The directory name carries a hyphen, so import it as
`importlib.import_module("tinker-gpu_b200")` or through the `tinker_gpu_b200`
alias module at the repo root.
"""
from . import tinkerio, params  # noqa: F401
from .tinkerio import read_xyz, read_key, read_prm, find_prm  # noqa: F401
from .params import System, build_system, replicate, save_system, load_system  # noqa: F401


def load_tinker(xyz_path, key_path=None, key_text=None, prm_path=None, prm_dirs=()):
    """.xyz + .key (+ .prm named by the key) -> System.  Mirrors the
    `initial(); getxyz(); mechanic()` preamble of the reference drivers (src/xanalyze.cpp:39)."""
    x = read_xyz(xyz_path)
    k = read_key(key_path, key_text)
    if prm_path is None:
        prm_path = find_prm(k, prm_dirs)
    ff = read_prm(prm_path)
    return build_system(x, k, ff)
