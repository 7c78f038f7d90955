"""ctypes access to oracle/_ref/libref_dropin.so: the reference's unmodified front-ends (src/amoeba/field.cpp, induce.cpp,
emplar.cpp, mpole.cpp) and energy-buffer reductions (src/energybuffer.cpp) linked
with integration/apx_adapter.cpp and libapx instead of the reference's CUDA kernels (oracle/ref_dropin.cpp, `make -C oracle
dropin`).  TEST INFRASTRUCTURE ONLY.  One system per process."""
import ctypes as C
import importlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_dropin.so")
_DP = C.POINTER(C.c_double)


def available():
    return os.path.isfile(LIB)


def _dp(a):
    return a.ctypes.data_as(_DP)


class DropIn:
    """tinker::induce / dfield / ufield / sparsePrecondApply (the reference's own functions) and emplar_cu on `system`."""

    def __init__(self, system):
        am = importlib.import_module("tinker-gpu_b200.amoeba")
        self.lib = lib = C.CDLL(LIB)
        lib.dropin_last_error.restype = C.c_char_p
        lib.dropin_open.argtypes = [C.POINTER(am._ApxSystem), C.c_double]
        lib.dropin_induce.argtypes = [_DP] * 4
        lib.dropin_dfield.argtypes = [_DP] * 2
        lib.dropin_ufield.argtypes = [_DP] * 4
        lib.dropin_precond.argtypes = [_DP] * 4
        lib.dropin_emplar.argtypes = [C.c_int, C.c_double, _DP, _DP, _DP]
        lib.dropin_set_xyz.argtypes = [_DP]
        lib.dropin_time_induce.argtypes = [C.c_int, _DP]
        self.n = int(system.n)
        st, self._keep = am.system_struct(system)
        self._check(lib.dropin_open(C.byref(st), float(system.list_buffer)), "dropin_open")

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.dropin_last_error().decode(errors='replace')}")

    def _o(self, k):
        return [np.zeros((self.n, 3)) for _ in range(k)]

    def induce(self):
        o = self._o(4)
        self._check(self.lib.dropin_induce(*[_dp(a) for a in o]), "dropin_induce")
        return dict(uind=o[0], uinp=o[1], udir=o[2], udirp=o[3])

    def dfield(self):
        o = self._o(2)
        self._check(self.lib.dropin_dfield(*[_dp(a) for a in o]), "dropin_dfield")
        return o[0], o[1]

    def ufield(self, ud, up):
        i = [np.ascontiguousarray(ud, np.float64), np.ascontiguousarray(up, np.float64)]
        o = self._o(2)
        self._check(self.lib.dropin_ufield(*[_dp(a) for a in i + o]), "dropin_ufield")
        return o[0], o[1]

    def precond(self, rd, rp):
        i = [np.ascontiguousarray(rd, np.float64), np.ascontiguousarray(rp, np.float64)]
        o = self._o(2)
        self._check(self.lib.dropin_precond(*[_dp(a) for a in i + o]), "dropin_precond")
        return o[0], o[1]

    def emplar(self, vers=0x70, preload=0.0):
        """tinker::emplar(vers) between the halves of the reference's energy(): accumulators zeroed, `preload` added to every
        gradient entry and to slot 0 of the energy buffer, then the reference's own energyReduce / virialReduce."""
        e = C.c_double()
        g, v = np.zeros((self.n, 3)), np.zeros(9)
        self._check(self.lib.dropin_emplar(int(vers), float(preload), C.byref(e), _dp(g), _dp(v)), "dropin_emplar")
        return dict(esum=e.value, grad=g, virial=v.reshape(3, 3))

    def set_xyz(self, xyz):
        self._check(self.lib.dropin_set_xyz(_dp(np.ascontiguousarray(xyz, np.float64))), "dropin_set_xyz")

    def time_induce(self, reps=20):
        ms = C.c_double()
        self._check(self.lib.dropin_time_induce(int(reps), C.byref(ms)), "dropin_time_induce")
        return ms.value

    def close(self):
        self.lib.dropin_close()


def main(argv=None):
    """Child-process entry of the GPU test: the reference's front-ends on our kernels vs the float64 oracle fixture and vs
    the same operators called through the C ABI directly.  Prints one JSON line."""
    import argparse
    import json
    import sys
    ap = argparse.ArgumentParser()
    ap.add_argument("blob")
    ap.add_argument("--fixture", default=None)
    a = ap.parse_args(argv)
    sys.path.insert(0, os.path.dirname(HERE))
    tg = importlib.import_module("tinker_gpu_b200")
    am = importlib.import_module("tinker-gpu_b200.amoeba")
    s = tg.load_system(a.blob)
    d = DropIn(s)
    debye = 4.803206802
    out = dict(n=int(s.n), blob=os.path.basename(a.blob))
    u = d.induce()
    fd, fp = d.dfield()
    rng = np.random.default_rng(5)
    pd, pp = rng.normal(size=(s.n, 3)) * 0.05, rng.normal(size=(s.n, 3)) * 0.05
    ufd, ufp = d.ufield(pd, pp)
    zd, zp = d.precond(pd, pp)
    # the reference's emplar() + its own energyReduce / virialReduce, on zeroed accumulators and on pre-loaded ones: the
    # library's contribution must be ADDED (fixed point: the two runs differ by exactly the preload where sums are integers)
    e = d.emplar(0x70)
    e2 = d.emplar(0x70, preload=3.0)
    out["accumulate"] = dict(energy_delta=float(e2["esum"] - e["esum"]), grad_delta_min=float((e2["grad"] - e["grad"]).min()),
                             grad_delta_max=float((e2["grad"] - e["grad"]).max()))
    out["ms_induce_frontend"] = d.time_induce(20)
    if a.fixture:
        z = np.load(a.fixture)
        eref = float(z["em"]) + float(z["ep"])
        out["vs_oracle"] = dict(uind_rms_debye=float(np.sqrt(((u["uind"] - z["uind"]) ** 2).mean()) * debye),
                                udir_rms_debye=float(np.sqrt(((u["udir"] - z["udir"]) ** 2).mean()) * debye),
                                esum_rel=abs(e["esum"] - eref) / abs(eref),
                                grad_rms=float(np.sqrt(((e["grad"] - z["grad"]) ** 2).mean())),
                                virial_rel=float(np.abs(e["virial"] - z["virial"]).max() / np.abs(z["virial"]).max()))
    d.close()
    # the same operators through the C ABI directly: the front-end route may differ only by the float round trip of the globals
    b = am.Amoeba(s, "mixed", device=0)
    b.lib.apx_mpole_init(b.ctx)
    f0, f1 = b.dfield()
    g0, g1 = b.ufield(pd, pp)
    y0, y1 = b.sparsePrecondApply(pd, pp)
    v0, _ = b.induce()
    ms = []
    for _ in range(20):
        b.lib.apx_induce(b.ctx)
        ms.append(b.stats()["ms_induce"])
    out["ms_induce_c_abi"] = float(np.mean(ms))
    b.close()
    out["vs_c_abi"] = dict(dfield=float(np.abs(fd - f0).max() / np.abs(f0).max()), dfieldp=float(np.abs(fp - f1).max() / np.abs(f1).max()),
                           ufield=float(np.abs(ufd - g0).max() / np.abs(g0).max()), precond=float(np.abs(zd - y0).max() / np.abs(y0).max()),
                           uind=float(np.abs(u["uind"] - v0).max() / np.abs(v0).max()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
