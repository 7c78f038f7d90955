"""ctypes access to oracle/_ref/libref_cuda.so -- the reference's own CUDA kernels for the path (oracle/ref_cuda.cu, `make -C
oracle cuda`), run on the same GPU on the same System.  COMPARATOR / TEST INFRASTRUCTURE ONLY: used by tests/ and by bench.py's
`ref_cuda` leg (in a child process), never by the product path.  One system per process (the reference keeps its state in
process globals)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_cuda.so")
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)

ENERGY, GRAD, VIRIAL = 0x10, 0x20, 0x40      # calc::energy / grad / virial (include/ff/calc.h)


def available():
    return os.path.isfile(LIB)


class _RefcuSystem(C.Structure):      # struct refcu_system of oracle/ref_cuda.cu, field for field
    _fields_ = [("n", C.c_int), ("xyz", _DP), ("lvec9", _DP), ("recip9", _DP), ("pole", _DP), ("zaxis", _IP), ("polarity", _DP),
                ("thole", _DP), ("pdamp", _DP), ("jpolar", _IP), ("njpolar", C.c_int), ("thlval", _DP),
                ("nmexclude", C.c_int), ("mexclude", _IP), ("mexclude_scale", _DP),
                ("ndpexclude", C.c_int), ("dpexclude", _IP), ("dpexclude_scale", _DP),
                ("nuexclude", C.c_int), ("uexclude", _IP), ("uexclude_scale", _DP),
                ("nmdpuexclude", C.c_int), ("mdpuexclude", _IP), ("mdpuexclude_scale", _DP),
                ("aewald", C.c_double), ("nfft", C.c_int * 3), ("bsorder", C.c_int), ("bsmod1", _DP), ("bsmod2", _DP), ("bsmod3", _DP),
                ("ewald_cutoff", C.c_double), ("usolve_cutoff", C.c_double), ("list_buffer", C.c_double), ("poleps", C.c_double),
                ("politer", C.c_int), ("pcgprec", C.c_int), ("pcgguess", C.c_int), ("pcgpeek", C.c_double), ("uaccel", C.c_double),
                ("electric", C.c_double), ("dielec", C.c_double)]


class _RefcuVdw(C.Structure):         # struct refcu_vdw of oracle/ref_cuda.cu
    _fields_ = [("ired", _IP), ("kred", _DP), ("jvdw", _IP), ("njvdw", C.c_int), ("radmin", _DP), ("epsilon", _DP),
                ("nvexclude", C.c_int), ("vexclude", _IP), ("vexclude_scale", _DP),
                ("cutoff", C.c_double), ("taper", C.c_double), ("list_buffer", C.c_double)]


class RefCuda:
    """mpoleInit + induce() and the fused energy step (emplar) of the reference's CUDA build on `system`."""

    def __init__(self, system):
        from .amoeba_ref import Oracle
        if not (system.use_ewald and system.use_mpole and system.use_polar and system.poltyp == "MUTUAL"):
            raise ValueError("the comparator drives the Ewald / mutual-polarization path only")
        self.lib = lib = C.CDLL(LIB)
        lib.refcu_last_error.restype = C.c_char_p
        lib.refcu_open.argtypes = [C.POINTER(_RefcuSystem)]
        lib.refcu_set_xyz.argtypes = [_DP, C.c_int]
        lib.refcu_induce.argtypes = [_DP] * 4
        lib.refcu_energy.argtypes = [C.c_int, _DP, _DP, _DP]
        lib.refcu_time.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        s = system
        self.n = n = int(s.n)
        f64 = lambda a: np.ascontiguousarray(a, np.float64)      # noqa: E731
        i32 = lambda a: np.ascontiguousarray(a, np.int32)        # noqa: E731
        pad = lambda a, shape: a if a.size else np.zeros(shape, a.dtype)      # noqa: E731
        keep = dict(xyz=f64(s.xyz), lvec=f64(s.lvec).ravel(), recip=f64(s.recip).ravel(), pole=f64(s.pole), zaxis=i32(s.zaxis),
                    polarity=f64(s.polarity), thole=f64(s.thole), pdamp=f64(s.pdamp), jpolar=i32(s.jpolar), thlval=f64(s.thlval),
                    mex=pad(i32(s.mexclude), (1, 2)), mexs=pad(f64(s.mexclude_scale), (1,)),
                    dpex=pad(i32(s.dpexclude), (1, 2)), dpexs=pad(f64(s.dpexclude_scale), (1, 2)),
                    uex=pad(i32(s.uexclude), (1, 2)), uexs=pad(f64(s.uexclude_scale), (1,)),
                    mdpu=pad(i32(s.mdpuexclude), (1, 2)), mdpus=pad(f64(s.mdpuexclude_scale), (1, 4)),
                    b1=f64(Oracle.bsmod(int(s.nfft[0]), int(s.bsorder))), b2=f64(Oracle.bsmod(int(s.nfft[1]), int(s.bsorder))),
                    b3=f64(Oracle.bsmod(int(s.nfft[2]), int(s.bsorder))))
        self._keep = keep
        dp = lambda k: keep[k].ctypes.data_as(_DP)      # noqa: E731
        ip = lambda k: keep[k].ctypes.data_as(_IP)      # noqa: E731
        st = _RefcuSystem()
        st.n, st.xyz, st.lvec9, st.recip9, st.pole, st.zaxis = n, dp("xyz"), dp("lvec"), dp("recip"), dp("pole"), ip("zaxis")
        st.polarity, st.thole, st.pdamp, st.jpolar = dp("polarity"), dp("thole"), dp("pdamp"), ip("jpolar")
        st.njpolar, st.thlval = int(keep["thlval"].shape[0]), dp("thlval")
        st.nmexclude, st.mexclude, st.mexclude_scale = int(s.mexclude.shape[0]), ip("mex"), dp("mexs")
        st.ndpexclude, st.dpexclude, st.dpexclude_scale = int(s.dpexclude.shape[0]), ip("dpex"), dp("dpexs")
        st.nuexclude, st.uexclude, st.uexclude_scale = int(s.uexclude.shape[0]), ip("uex"), dp("uexs")
        st.nmdpuexclude, st.mdpuexclude, st.mdpuexclude_scale = int(s.mdpuexclude.shape[0]), ip("mdpu"), dp("mdpus")
        st.aewald, st.bsorder = float(s.aewald), int(s.bsorder)
        st.nfft = (C.c_int * 3)(*[int(v) for v in s.nfft])
        st.bsmod1, st.bsmod2, st.bsmod3 = dp("b1"), dp("b2"), dp("b3")
        # System.usolve_cutoff is the range the preconditioner applies = switchOff(USOLVE) + list buffer (precond.cu:30-31)
        st.ewald_cutoff, st.list_buffer = float(s.ewald_cutoff), float(s.list_buffer)
        st.usolve_cutoff = max(float(s.usolve_cutoff) - float(s.list_buffer), 0.0)
        st.poleps, st.politer, st.pcgprec, st.pcgguess = float(s.poleps), int(s.politer), int(bool(s.pcgprec)), int(bool(s.pcgguess))
        st.pcgpeek, st.uaccel, st.electric, st.dielec = float(s.pcgpeek), float(s.uaccel), float(s.electric), float(s.dielec)
        self._check(lib.refcu_open(C.byref(st)), "refcu_open")

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.refcu_last_error().decode(errors='replace')}")

    def set_xyz(self, xyz, rebuild=True):
        a = np.ascontiguousarray(xyz, np.float64)
        self._check(self.lib.refcu_set_xyz(a.ctypes.data_as(_DP), int(rebuild)), "refcu_set_xyz")

    def induce(self):
        out = [np.zeros((self.n, 3)) for _ in range(4)]
        self._check(self.lib.refcu_induce(*[a.ctypes.data_as(_DP) for a in out]), "refcu_induce")
        return dict(uind=out[0], uinp=out[1], udir=out[2], udirp=out[3])

    def energy(self, vers=ENERGY | GRAD | VIRIAL):
        """esum = E_mpole + E_polar: without calc::analyz the reference's terms share one accumulator (empole.cpp:37-39)."""
        es = C.c_double()
        g, v = np.zeros((self.n, 3)), np.zeros(9)
        self._check(self.lib.refcu_energy(int(vers), C.byref(es), g.ctypes.data_as(_DP), v.ctypes.data_as(_DP)), "refcu_energy")
        return dict(esum=es.value, grad=g, virial=v.reshape(3, 3))

    # ---- buffered 14-7 vdW (src/cu/ehal.cu)
    def vdw_open(self, v):
        """v: the System's vdwparams.VdwTerm."""
        lib = self.lib
        lib.refcu_vdw_open.argtypes = [C.POINTER(_RefcuVdw)]
        lib.refcu_ehal.argtypes = [C.c_int, _DP, _DP, _DP]
        lib.refcu_time_ehal.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        nx = int(v.vexclude.shape[0]) if v.vexclude.size else 0
        keep = dict(ired=np.ascontiguousarray(v.ired, np.int32), kred=np.ascontiguousarray(v.kred, np.float64),
                    jvdw=np.ascontiguousarray(v.jvdw, np.int32), radmin=np.ascontiguousarray(v.radmin, np.float64),
                    epsilon=np.ascontiguousarray(v.epsilon, np.float64),
                    vex=np.ascontiguousarray(v.vexclude.reshape(-1, 2) if nx else np.zeros((1, 2)), np.int32),
                    vexs=np.ascontiguousarray(v.vexclude_scale if nx else np.ones(1), np.float64))
        self._keep_vdw = keep
        st = _RefcuVdw()
        st.ired, st.kred, st.jvdw = keep["ired"].ctypes.data_as(_IP), keep["kred"].ctypes.data_as(_DP), keep["jvdw"].ctypes.data_as(_IP)
        st.njvdw, st.radmin, st.epsilon = int(keep["radmin"].shape[0]), keep["radmin"].ctypes.data_as(_DP), keep["epsilon"].ctypes.data_as(_DP)
        st.nvexclude, st.vexclude, st.vexclude_scale = nx, keep["vex"].ctypes.data_as(_IP), keep["vexs"].ctypes.data_as(_DP)
        st.cutoff, st.taper, st.list_buffer = float(v.cutoff), float(v.taper), float(v.list_buffer)
        self._check(lib.refcu_vdw_open(C.byref(st)), "refcu_vdw_open")

    def ehal(self, vers=ENERGY | GRAD | VIRIAL):
        """Pair sum of the 14-7 term (no long-range correction), gradient on the real atoms, pair virial."""
        e = C.c_double()
        g, v = np.zeros((self.n, 3)), np.zeros(9)
        self._check(self.lib.refcu_ehal(int(vers), C.byref(e), g.ctypes.data_as(_DP), v.ctypes.data_as(_DP)), "refcu_ehal")
        return dict(ev=e.value, grad=g, virial=v.reshape(3, 3))

    def time_ehal(self, reps=20, warmup=3, vers=ENERGY | GRAD):
        ms = (C.c_float * reps)()
        self._check(self.lib.refcu_time_ehal(int(vers), int(warmup), int(reps), ms), "refcu_time_ehal")
        return np.array(ms[:], np.float64)

    def time(self, what, reps=20, warmup=3, vers=ENERGY | GRAD | VIRIAL):
        """CUDA-event milliseconds per call: what = 'induce' | 'energy' | 'rebuild'."""
        ms = (C.c_float * reps)()
        code = {"induce": 0, "energy": 1, "rebuild": 2}[what]
        self._check(self.lib.refcu_time(code, int(vers), int(warmup), int(reps), ms), "refcu_time")
        return np.array(ms[:], np.float64)


def main(argv=None):
    """Child-process entry used by bench.py and the GPU test: loads a System blob, checks the reference CUDA path against the
    committed oracle fixture when one is given, times induce() and the energy step, prints one JSON line."""
    import argparse
    import importlib
    import json
    import sys
    ap = argparse.ArgumentParser()
    ap.add_argument("blob")
    ap.add_argument("--fixture", default=None, help="npz with em, ep, uind, grad of the float64 oracle on the same system")
    ap.add_argument("--vdw", default=None, metavar="FIXTURE", help="also run the reference's ehal.cu; npz with ev_pairs, grad of the vdW oracle "
                    "(a SECOND JSON line; the first one is already out if this part fails)")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args(argv)
    sys.path.insert(0, os.path.dirname(HERE))
    tg = importlib.import_module("tinker_gpu_b200")
    s = tg.load_system(a.blob)
    r = RefCuda(s)
    out = dict(n=int(s.n), blob=os.path.basename(a.blob))
    d = r.induce()
    e = r.energy()
    out.update(esum=e["esum"])
    if a.fixture:
        z = np.load(a.fixture)
        debye = 4.803206802
        e0 = float(z["em"]) + float(z["ep"])
        out["parity"] = dict(esum_rel=abs(e["esum"] - e0) / abs(e0),
                             uind_rms_debye=float(np.sqrt(((d["uind"] - z["uind"]) ** 2).mean()) * debye),
                             grad_rms=float(np.sqrt(((e["grad"] - z["grad"]) ** 2).sum(1).mean())),
                             virial_rel=float(np.abs(e["virial"] - z["virial"]).max() / np.abs(z["virial"]).max()))
    for key, what, vers in (("induce_ms", "induce", ENERGY), ("energy_ms", "energy", ENERGY | GRAD),
                            ("energy_v1_ms", "energy", ENERGY | GRAD | VIRIAL), ("rebuild_ms", "rebuild", 0)):
        ms = r.time(what, a.reps, a.warmup, vers)
        out[key] = dict(median=float(np.median(ms)), min=float(ms.min()), max=float(ms.max()), reps=a.reps)
    out["energy_ms"]["vers"] = "energy+grad (calc::v4): zero accumulators, mpoleInit, induce, emplar kernels, recip, torque, reductions"
    print(json.dumps(out), flush=True)
    if a.vdw and s.vdw is not None:
        r.vdw_open(s.vdw)
        w = r.ehal()
        z = np.load(a.vdw)
        vout = dict(ev_pairs=w["ev"], parity=dict(ev_rel=abs(w["ev"] - float(z["ev_pairs"])) / abs(float(z["ev_pairs"])),
                                                  grad_rms=float(np.sqrt(((w["grad"] - z["grad"]) ** 2).sum(1).mean())),
                                                  virial_rel=float(np.abs(w["virial"] - z["virial_pairs"]).max() / np.abs(z["virial_pairs"]).max())))
        ms = r.time_ehal(a.reps, a.warmup)
        vout["ehal_ms"] = dict(median=float(np.median(ms)), min=float(ms.min()), max=float(ms.max()), reps=a.reps,
                               vers="energy+grad: reduced sites, zero accumulators, ehal_cu incl. gradient hand-back, reduction")
        print(json.dumps({"vdw": vout}), flush=True)


if __name__ == "__main__":
    main()
