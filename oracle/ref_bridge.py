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


def realspace(oracle, ud=None, up=None, threads=1):
    """Real-space multipole / polarization energies, gradients, torques and the d/p permanent and mutual fields from the
    reference's pair_mpole / pair_polar / pair_dfield / pair_ufield over the oracle's own pair list and scale factors.
    threads > 1 (bench.py's CPU legs): the pair list in contiguous slices on that many host threads, sums added; the pairwise
    virial (a process-global accumulator in the shim) is then not collected."""
    from concurrent.futures import ThreadPoolExecutor
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
    names = ("gm", "tm", "gp", "tp", "fd", "fp", "ufd", "ufp")
    u1 = None if ud is None else np.ascontiguousarray(ud, np.float64)
    u2 = None if up is None else np.ascontiguousarray(up, np.float64)
    vm6, vp6 = np.zeros(6), np.zeros(6)
    lib.ref_realspace_virial.argtypes = [_DP, _DP]
    lib.ref_realspace_virial.restype = None

    def sweep(lo, hi):
        em, ep = C.c_double(), C.c_double()
        o = {nm: np.zeros((n, 3)) for nm in names}
        rc = lib.ref_realspace_eval(n, hi - lo, i32[lo:hi].ctypes.data_as(_IP), k32[lo:hi].ctypes.data_as(_IP), _dp(R[lo:hi]), _dp(sc[lo:hi]), _dp(rp),
                                    _dp(pd), _dp(pga[lo:hi]), None if u1 is None else _dp(u1), None if u2 is None else _dp(u2), float(oracle.f),
                                    float(s.aewald), int(bool(s.use_ewald)), C.byref(em), C.byref(ep), *[_dp(o[nm]) for nm in names])
        if rc != 0:
            raise RuntimeError(f"ref_realspace_eval failed ({rc})")
        o.update(em=em.value, ep=ep.value)
        return o

    T = max(1, min(int(threads), len(i32) // 4096 or 1))
    if T == 1:
        lib.ref_realspace_virial(_dp(vm6), _dp(vp6))
        out = sweep(0, len(i32))
        lib.ref_realspace_virial(None, None)
    else:
        lib.ref_realspace_virial(None, None)
        cuts = np.linspace(0, len(i32), T + 1).astype(int)
        with ThreadPoolExecutor(T) as ex:
            parts = list(ex.map(lambda t: sweep(int(cuts[t]), int(cuts[t + 1])), range(T)))
        out = parts[0]
        for q in parts[1:]:
            for nm in names + ("em", "ep"):
                out[nm] = out[nm] + q[nm]

    def sym(v):
        return np.array([[v[0], v[1], v[2]], [v[1], v[3], v[4]], [v[2], v[4], v[5]]])
    out.update(npair=len(i32), vm=sym(vm6), vp=sym(vp6))
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


class RefPME:
    """The reference's host PME operators (src/acc/pme.cpp compiled in place into oracle/_ref/libref_pme.so) set up on the
    coordinates, box and grid of an Oracle; the FFT between the operators is numpy's.  Grids are complex [n3][n2][n1]."""

    def __init__(self, oracle):
        self.lib = lib = C.CDLL(os.path.join(HERE, "_ref", "libref_pme.so"))
        s = oracle.s
        self.n = oracle.n
        self.nfft = tuple(int(v) for v in s.nfft)
        p = oracle.pme_setup()
        b1, b2, b3 = (np.ascontiguousarray(b, np.float64) for b in p["bsmod"])
        xyz = np.ascontiguousarray(oracle.xyz, np.float64)
        lv = np.ascontiguousarray(oracle.lvec, np.float64).ravel()
        rc = np.ascontiguousarray(oracle.recip, np.float64).ravel()
        nf = (C.c_int * 3)(*self.nfft)
        lib.ref_pme_open.argtypes = [C.c_int, _DP, _DP, _DP, C.POINTER(C.c_int), C.c_int, C.c_double, _DP, _DP, _DP, C.c_double, C.c_double]
        lib.ref_pme_open(self.n, _dp(xyz), _dp(lv), _dp(rc), nf, int(s.bsorder), float(s.aewald), _dp(b1), _dp(b2), _dp(b3),
                         float(s.electric), float(s.dielec))
        for name, k in (("ref_pme_qgrid_get", 1), ("ref_pme_qgrid_set", 1), ("ref_pme_rpole_to_cmp", 2), ("ref_pme_cmp_to_fmp", 2),
                        ("ref_pme_cuind_to_fuind", 4), ("ref_pme_fphi_to_cphi", 2), ("ref_pme_grid_mpole", 1), ("ref_pme_grid_uind", 2),
                        ("ref_pme_conv", 2), ("ref_pme_fphi_mpole", 1), ("ref_pme_fphi_uind", 3), ("ref_pme_fphi_uind2", 2)):
            getattr(lib, name).argtypes = [_DP] * k
            getattr(lib, name).restype = None

    def _grid_out(self):
        n1, n2, n3 = self.nfft
        g = np.zeros((n3, n2, n1, 2))
        self.lib.ref_pme_qgrid_get(_dp(g))
        return g[..., 0] + 1j * g[..., 1]

    def _grid_in(self, q):
        g = np.ascontiguousarray(np.stack([q.real, q.imag], -1), np.float64)
        self.lib.ref_pme_qgrid_set(_dp(g))

    def _call(self, name, ins, out_shapes):
        ins = [np.ascontiguousarray(a, np.float64) for a in ins]
        outs = [np.zeros(sh) for sh in out_shapes]
        getattr(self.lib, name)(*[_dp(a) for a in ins + outs])
        return outs[0] if len(outs) == 1 else outs

    def rpole_to_cmp(self, rp):
        return self._call("ref_pme_rpole_to_cmp", [rp], [(self.n, 10)])

    def cmp_to_fmp(self, cmp_):
        return self._call("ref_pme_cmp_to_fmp", [cmp_], [(self.n, 10)])

    def cuind_to_fuind(self, ud, up):
        return self._call("ref_pme_cuind_to_fuind", [ud, up], [(self.n, 3), (self.n, 3)])

    def fphi_to_cphi(self, fphi):
        return self._call("ref_pme_fphi_to_cphi", [fphi], [(self.n, 10)])

    def grid_mpole(self, fmp):
        a = np.ascontiguousarray(fmp, np.float64)
        self.lib.ref_pme_grid_mpole(_dp(a))
        return self._grid_out()

    def grid_uind(self, fud, fup):
        a, b = np.ascontiguousarray(fud, np.float64), np.ascontiguousarray(fup, np.float64)
        self.lib.ref_pme_grid_uind(_dp(a), _dp(b))
        return self._grid_out()

    def convolve(self, qgrid, want_ev=False):
        """fftfront (numpy) + pmeConv (reference) + fftback (numpy): (grid, e, virial 3x3)."""
        self._grid_in(np.fft.fftn(qgrid))
        e, v6 = np.zeros(1), np.zeros(6)
        self.lib.ref_pme_conv(_dp(e) if want_ev else None, _dp(v6) if want_ev else None)
        out = np.fft.ifftn(self._grid_out()) * qgrid.size
        self._grid_in(out)
        v = np.array([[v6[0], v6[1], v6[2]], [v6[1], v6[3], v6[4]], [v6[2], v6[4], v6[5]]])
        return out, (float(e[0]) if want_ev else None), (v if want_ev else None)

    def fphi_mpole(self, grid=None):
        if grid is not None:
            self._grid_in(grid)
        return self._call("ref_pme_fphi_mpole", [], [(self.n, 20)])

    def fphi_uind(self, grid=None):
        if grid is not None:
            self._grid_in(grid)
        return self._call("ref_pme_fphi_uind", [], [(self.n, 10), (self.n, 10), (self.n, 20)])

    def fphi_uind2(self, grid=None):
        if grid is not None:
            self._grid_in(grid)
        return self._call("ref_pme_fphi_uind2", [], [(self.n, 10), (self.n, 10)])


def hal_pairs(vdw_oracle, pairs=None, threads=1):
    """ev, gradient on the reduced sites and virial of the 14-7 term from the reference's pair_hal_v2 over the vdW oracle's own
    pair list (pairs = (i, k) to reuse a list found earlier).  threads > 1 (bench.py's CPU legs only): the pair list is cut into
    that many slices, each prepared and swept on its own host thread (numpy and ctypes release the GIL), partial sums added."""
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_realspace.so"))
    lib.ref_hal_pairs.argtypes = [C.c_int, C.c_longlong, _IP, _IP, _DP, _DP, _DP, C.c_double, C.c_double, C.c_double, C.c_double, _DP, _DP, _DP]
    o, v = vdw_oracle, vdw_oracle.v
    xr = o.reduced()
    i_all, k_all = o.pairs(xr) if pairs is None else pairs
    ex_sorted = ex_scale = None
    if v.vexclude.shape[0]:
        ex = v.vexclude[:, 0].astype(np.int64) * o.n + v.vexclude[:, 1]
        order = np.argsort(ex)
        ex_sorted, ex_scale = ex[order], v.vexclude_scale[order]

    def sweep(lo, hi):
        i, k = i_all[lo:hi], k_all[lo:hi]
        scale = np.ones(i.shape[0])
        if ex_sorted is not None:
            code = i.astype(np.int64) * o.n + k
            pos = np.minimum(np.searchsorted(ex_sorted, code), ex_sorted.shape[0] - 1)
            hit = ex_sorted[pos] == code
            scale[hit] = ex_scale[pos[hit]]
        keep = scale != 0
        i, k, scale = i[keep], k[keep], scale[keep]
        d = np.ascontiguousarray(o.image(xr[i] - xr[k]), np.float64)
        rv = np.ascontiguousarray(v.radmin[v.jvdw[i], v.jvdw[k]], np.float64)
        eps = np.ascontiguousarray(v.epsilon[v.jvdw[i], v.jvdw[k]] * scale, np.float64)
        i32, k32 = np.ascontiguousarray(i, np.int32), np.ascontiguousarray(k, np.int32)
        ev = C.c_double()
        g, v9 = np.zeros((o.n, 3)), np.zeros(9)
        lib.ref_hal_pairs(o.n, len(i32), i32.ctypes.data_as(_IP), k32.ctypes.data_as(_IP), _dp(d), _dp(rv), _dp(eps), float(v.taper),
                          float(v.cutoff), float(v.ghal), float(v.dhal), C.byref(ev), _dp(g), _dp(v9))
        return ev.value, g, v9, len(i32)

    npair = int(i_all.shape[0])
    T = max(1, min(int(threads), npair // 65536 or 1))
    if T == 1:
        ev, g, v9, m = sweep(0, npair)
    else:
        from concurrent.futures import ThreadPoolExecutor
        cuts = np.linspace(0, npair, T + 1).astype(int)
        with ThreadPoolExecutor(T) as exr:
            parts = list(exr.map(lambda t: sweep(int(cuts[t]), int(cuts[t + 1])), range(T)))
        ev = sum(p[0] for p in parts)
        g = sum(p[1] for p in parts)
        v9 = sum(p[2] for p in parts)
        m = sum(p[3] for p in parts)
    return dict(ev=ev, gred=g, virial=v9.reshape(3, 3), npairs=m)


def _frames_lib():
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_pme.so"))
    lib.ref_frames_rotpole.argtypes = [C.c_int, _DP, _IP, _DP, _DP, _DP]
    lib.ref_frames_torque.argtypes = [C.c_int, _DP, _IP, _DP, _DP, _DP]
    return lib


def rotpole(xyz, zaxis, pole):
    """(pole after chkpole, rpole) from the reference's chkpole_acc + rotpole_acc (src/acc/amoeba/rotpole.cpp)."""
    n = len(xyz)
    x = np.ascontiguousarray(xyz, np.float64)
    z = np.ascontiguousarray(zaxis, np.int32)
    p = np.ascontiguousarray(pole, np.float64)
    pc, rp = np.zeros((n, 10)), np.zeros((n, 10))
    _frames_lib().ref_frames_rotpole(n, _dp(x), z.ctypes.data_as(_IP), _dp(p), _dp(pc), _dp(rp))
    return pc, rp


def torque(xyz, zaxis, trq):
    """(gradient on the frame atoms, torque virial 3x3) from the reference's torque_acc (src/acc/amoeba/torque.cpp)."""
    n = len(xyz)
    x = np.ascontiguousarray(xyz, np.float64)
    z = np.ascontiguousarray(zaxis, np.int32)
    t = np.ascontiguousarray(trq, np.float64)
    g, v6 = np.zeros((n, 3)), np.zeros(6)
    _frames_lib().ref_frames_torque(n, _dp(x), z.ctypes.data_as(_IP), _dp(t), _dp(g), _dp(v6))
    return g, np.array([[v6[0], v6[1], v6[2]], [v6[1], v6[3], v6[4]], [v6[2], v6[4], v6[5]]])


_FFT_CB = C.CFUNCTYPE(None, C.c_int)


def _recip_call(P, name, ins):
    """Run one of the reference's reciprocal assembly routines; its fftfront / fftback come back here as a callback that
    transforms the grid held by the library in place with numpy (unnormalised both ways, as FFTW / cuFFT are)."""
    lib = P.lib

    def fft(forward):
        g = P._grid_out()
        P._grid_in(np.fft.fftn(g) if forward else np.fft.ifftn(g) * g.size)
    cb = _FFT_CB(fft)
    fn = getattr(lib, name)
    fn.argtypes = [_FFT_CB] + [_DP] * (len(ins) + 4)
    fn.restype = C.c_int
    n = P.n
    e, g, t, v6 = np.zeros(1), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(6)
    arrs = [np.ascontiguousarray(a, np.float64) for a in ins]
    rc = fn(cb, *[_dp(a) for a in arrs], _dp(e), _dp(g), _dp(t), _dp(v6))
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc})")
    v = np.array([[v6[0], v6[1], v6[2]], [v6[1], v6[3], v6[4]], [v6[2], v6[4], v6[5]]])
    return dict(e=float(e[0]), g=g, t=t, v=v)


def recip_mpole(P, rpole):
    """empoleChgpenEwaldRecip_acc(calc::v1, 0) (src/acc/hippo/empole.cpp:260-388): reciprocal multipole energy, gradient,
    torque, virial.  P: a RefPME set up on the system."""
    return _recip_call(P, "ref_recip_mpole", [rpole])


def recip_polar(P, uind, uinp):
    """epolarEwaldRecipSelf_acc(calc::v1, uind, uinp) (src/acc/amoeba/epolarewald.cpp:358-667); call recip_mpole first."""
    return _recip_call(P, "ref_recip_polar", [uind, uinp])
