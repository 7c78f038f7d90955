"""ctypes access to oracle/_ref/*.so -- the reference's own arithmetic compiled in place (oracle/Makefile).  TEST
INFRASTRUCTURE ONLY: imported by tests/ (and bench.py's cpu_baseline leg), never by the product path."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


def available(name):
    return os.path.isfile(os.path.join(HERE, "_ref", f"libref_{name}.so"))


def _dp(a):
    return a.ctypes.data_as(_DP)


def valence(system):
    """energy(8 terms), gradient, virial of the reference's dk_bond ... dk_tortor over the lists of system.valence."""
    import importlib
    am = importlib.import_module("tinker-gpu_b200.amoeba")
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_valence.so"))
    lib.ref_valence_eval.argtypes = [C.POINTER(am._ApxValence), _DP, _DP, _DP, _DP]
    st, keep = am.valence_struct(system.valence, system.n)
    x = np.ascontiguousarray(system.xyz, np.float64)
    e8, g, v9 = np.zeros(8), np.zeros((system.n, 3)), np.zeros(9)
    rc = lib.ref_valence_eval(C.byref(st), _dp(x), _dp(e8), _dp(g), _dp(v9))
    if rc != 0:
        raise RuntimeError(f"ref_valence_eval failed ({rc})")
    return dict(energy=e8, grad=g, virial=v9.reshape(3, 3))


def realspace(oracle, ud=None, up=None):
    """Real-space multipole / polarization energies, gradients, torques and the d/p permanent and mutual fields from the
    reference's pair_mpole / pair_polar / pair_dfield / pair_ufield over the oracle's own pair list and scale factors."""
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_realspace.so"))
    lib.ref_realspace_eval.argtypes = [C.c_int, C.c_longlong, _IP, _IP, _DP, _DP, _DP, _DP, _DP, _DP, _DP, C.c_double, C.c_double,
                                       C.c_int, _DP, _DP] + [_DP] * 8
    s = oracle.s
    n = oracle.n
    i, k, R, r = oracle.pairs(s.ewald_cutoff)
    sc = np.ascontiguousarray(oracle._scales(i, k), np.float64)
    pga = np.ascontiguousarray(oracle._pair_params(i, k)[2], np.float64)
    rp = np.ascontiguousarray(oracle._ensure_rpole(), np.float64)
    i32, k32 = np.ascontiguousarray(i, np.int32), np.ascontiguousarray(k, np.int32)
    R = np.ascontiguousarray(R, np.float64)
    pd = np.ascontiguousarray(s.pdamp, np.float64)
    em, ep = C.c_double(), C.c_double()
    out = {nm: np.zeros((n, 3)) for nm in ("gm", "tm", "gp", "tp", "fd", "fp", "ufd", "ufp")}
    u1 = None if ud is None else np.ascontiguousarray(ud, np.float64)
    u2 = None if up is None else np.ascontiguousarray(up, np.float64)
    rc = lib.ref_realspace_eval(n, len(i32), i32.ctypes.data_as(_IP), k32.ctypes.data_as(_IP), _dp(R), _dp(sc), _dp(rp), _dp(pd), _dp(pga),
                                None if u1 is None else _dp(u1), None if u2 is None else _dp(u2), float(oracle.f), float(s.aewald),
                                int(bool(s.use_ewald)), C.byref(em), C.byref(ep), *[_dp(out[nm]) for nm in ("gm", "tm", "gp", "tp", "fd", "fp", "ufd", "ufp")])
    if rc != 0:
        raise RuntimeError(f"ref_realspace_eval failed ({rc})")
    out.update(em=em.value, ep=ep.value, npair=len(i32))
    return out


def bspline5(w):
    """theta[m, 5, 4] (value, 1st, 2nd, 3rd derivative) from the reference's bsplgen<4>."""
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_realspace.so"))
    lib.ref_bspline5.argtypes = [C.c_int, _DP, _DP]
    w = np.ascontiguousarray(w, np.float64)
    out = np.zeros((len(w), 5, 4))
    lib.ref_bspline5(len(w), _dp(w), _dp(out))
    return out


def hal(r, rv, eps, evcut, evoff, ghal, dhal):
    """(e, dE/dr) per pair from the reference's pair_hal_v2."""
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_realspace.so"))
    lib.ref_hal.argtypes = [C.c_longlong, _DP, _DP, _DP, C.c_double, C.c_double, C.c_double, C.c_double, _DP, _DP]
    r, rv, eps = (np.ascontiguousarray(a, np.float64) for a in (r, rv, eps))
    e, de = np.zeros(len(r)), np.zeros(len(r))
    lib.ref_hal(len(r), _dp(r), _dp(rv), _dp(eps), float(evcut), float(evoff), float(ghal), float(dhal), _dp(e), _dp(de))
    return e, de
