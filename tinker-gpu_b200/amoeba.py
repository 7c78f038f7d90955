"""Host-side mirror of the reference's operator interface for the AMOEBA electrostatics path,
bound to the C ABI of libapx (include/apx.h) through ctypes.

Names follow the reference front-ends so that parity tests read like the reference's own:
`energy(vers)` (src/energy.cpp:319), `empole(vers)` (src/amoeba/empole.cpp:81), `epolar(vers)`
(src/amoeba/epolar.cpp:574), `induce()` (src/amoeba/induce.cpp:108), `dfield()` / `ufield()`
(src/amoeba/field.cpp:56,111), `sparsePrecondApply` (src/amoeba/induce.cpp:21), and the `calc::`
version flags (include/tool/rcman.h:107-133).

There is NO CPU fallback: constructing an `Amoeba` without the compiled library or without a
CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .params import System

HERE = os.path.dirname(os.path.abspath(__file__))


class calc:
    energy, grad, virial, analyz = 0x10, 0x20, 0x40, 0x80
    v0 = energy
    v1 = energy + grad + virial
    v3 = energy + analyz
    v4 = energy + grad
    v5 = grad
    v6 = grad + virial


UPRED = {"NONE": 0, "ASPC": 1, "GEAR": 2, "LSQR": 3}     # UPred, include/ff/amoeba/mpole.h:38


class ApxError(RuntimeError):
    """C-ABI image of the reference's FatalError (include/tool/error.h:16-45)."""


class _ApxSystem(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("xyz", C.POINTER(C.c_double)), ("lvec", C.c_double * 9),
        ("pole", C.POINTER(C.c_double)), ("zaxis", C.POINTER(C.c_int)),
        ("polarity", C.POINTER(C.c_double)), ("thole", C.POINTER(C.c_double)), ("pdamp", C.POINTER(C.c_double)),
        ("jpolar", C.POINTER(C.c_int)), ("njpolar", C.c_int), ("thlval", C.POINTER(C.c_double)),
        ("nmdpu", C.c_int), ("mdpu_ik", C.POINTER(C.c_int)), ("mdpu_scale", C.POINTER(C.c_double)),
        ("use_ewald", C.c_int), ("use_mpole", C.c_int), ("use_polar", C.c_int), ("poltyp_mutual", C.c_int),
        ("aewald", C.c_double), ("nfft", C.c_int * 3), ("bsorder", C.c_int),
        ("cutoff", C.c_double), ("usolve_cutoff", C.c_double), ("list_buffer", C.c_double),
        ("poleps", C.c_double), ("politer", C.c_int), ("uaccel", C.c_double),
        ("pcgprec", C.c_int), ("pcgguess", C.c_int), ("pcgpeek", C.c_double),
        ("electric", C.c_double), ("dielec", C.c_double), ("polpred", C.c_int),
    ]


class EnergyResult(C.Structure):
    _fields_ = [("em", C.c_double), ("ep", C.c_double), ("esum", C.c_double), ("virial", C.c_double * 9),
                ("nem", C.c_int), ("nep", C.c_int), ("pcg_iterations", C.c_int), ("pcg_eps", C.c_double),
                ("ev", C.c_double), ("nev", C.c_int),
                ("evalence", C.c_double), ("eval_term", C.c_double * 8), ("nval_term", C.c_int * 8)]


class _ApxVdw(C.Structure):
    _fields_ = [("n", C.c_int), ("ired", C.POINTER(C.c_int)), ("kred", C.POINTER(C.c_double)), ("jvdw", C.POINTER(C.c_int)),
                ("njvdw", C.c_int), ("radmin", C.POINTER(C.c_double)), ("epsilon", C.POINTER(C.c_double)),
                ("nvexclude", C.c_int), ("vexclude", C.POINTER(C.c_int)), ("vexclude_scale", C.POINTER(C.c_double)),
                ("cutoff", C.c_double), ("taper", C.c_double), ("ghal", C.c_double), ("dhal", C.c_double),
                ("elrc_vol", C.c_double), ("vlrc_vol", C.c_double)]


_IP = C.POINTER(C.c_int)


class _ApxValence(C.Structure):
    """include/apx.h: apx_valence."""
    _fields_ = [("n", C.c_int),
                ("nbond", C.c_int), ("ibnd", _IP), ("bk", C.POINTER(C.c_double)), ("bl", C.POINTER(C.c_double)),
                ("nangle", C.c_int), ("iang", _IP), ("ak", C.POINTER(C.c_double)), ("anat", C.POINTER(C.c_double)), ("angtyp", _IP),
                ("nstrbnd", C.c_int), ("isb", _IP), ("sbk", C.POINTER(C.c_double)), ("sb_anat", C.POINTER(C.c_double)),
                ("sb_bl", C.POINTER(C.c_double)),
                ("nurey", C.c_int), ("iury", _IP), ("uk", C.POINTER(C.c_double)), ("ul", C.POINTER(C.c_double)),
                ("nopbend", C.c_int), ("iopb", _IP), ("opbk", C.POINTER(C.c_double)), ("opbtyp", C.c_int),
                ("ntors", C.c_int), ("itors", _IP), ("tors_v", C.POINTER(C.c_double)), ("tors_phase", C.POINTER(C.c_double)),
                ("npitors", C.c_int), ("ipit", _IP), ("kpit", C.POINTER(C.c_double)),
                ("ntortor", C.c_int), ("itt", _IP), ("tt_chk", _IP), ("tt_grid", _IP),
                ("ngrid", C.c_int), ("tnx", _IP), ("tny", _IP), ("tt_off", _IP), ("tt_xoff", _IP), ("tt_yoff", _IP),
                ("ttx", C.POINTER(C.c_double)), ("tty", C.POINTER(C.c_double)), ("tbf", C.POINTER(C.c_double)),
                ("tbx", C.POINTER(C.c_double)), ("tby", C.POINTER(C.c_double)), ("tbxy", C.POINTER(C.c_double)),
                ("consts", C.c_double * 20), ("use", C.c_int * 8)]


class ValenceResult(C.Structure):
    _fields_ = [("e", C.c_double * 8), ("count", C.c_int * 8), ("esum", C.c_double), ("virial", C.c_double * 9)]


def valence_struct(v, n):
    """apx_valence over the arrays of a valparams.ValenceTerms; returns (struct, arrays to keep alive)."""
    keep = []

    def f64(a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        if a.size == 0:
            a = np.zeros(1)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32).ravel()
        if a.size == 0:
            a = np.zeros(1, np.int32)
        keep.append(a)
        return a.ctypes.data_as(_IP)

    s = _ApxValence()
    s.n = int(n)
    s.nbond, s.ibnd, s.bk, s.bl = len(v.ibnd), i32(v.ibnd), f64(v.bk), f64(v.bl)
    s.nangle, s.iang, s.ak, s.anat, s.angtyp = len(v.iang), i32(v.iang), f64(v.ak), f64(v.anat), i32(v.angtyp)
    s.nstrbnd, s.isb, s.sbk, s.sb_anat, s.sb_bl = len(v.isb), i32(v.isb), f64(v.sbk), f64(v.sb_anat), f64(v.sb_bl)
    s.nurey, s.iury, s.uk, s.ul = len(v.iury), i32(v.iury), f64(v.uk), f64(v.ul)
    s.nopbend, s.iopb, s.opbk, s.opbtyp = len(v.iopb), i32(v.iopb), f64(v.opbk), int(v.opbtyp)
    s.ntors, s.itors, s.tors_v, s.tors_phase = len(v.itors), i32(v.itors), f64(v.tors_v), f64(v.tors_phase)
    s.npitors, s.ipit, s.kpit = len(v.ipit), i32(v.ipit), f64(v.kpit)
    s.ntortor, s.itt, s.tt_chk, s.tt_grid = len(v.itt), i32(v.itt), i32(v.tt_chk), i32(v.tt_grid)
    s.ngrid, s.tnx, s.tny = len(v.tnx), i32(v.tnx), i32(v.tny)
    s.tt_off, s.tt_xoff, s.tt_yoff = i32(v.tt_off), i32(v.tt_xoff), i32(v.tt_yoff)
    s.ttx, s.tty, s.tbf, s.tbx, s.tby, s.tbxy = f64(v.ttx), f64(v.tty), f64(v.tbf), f64(v.tbx), f64(v.tby), f64(v.tbxy)
    s.consts = (C.c_double * 20)(*[float(x) for x in v.consts])
    s.use = (C.c_int * 8)(*[int(x) for x in v.use])
    return s, keep


class MdConfig(C.Structure):
    """include/apx.h: apx_md_config."""
    _fields_ = [("dt", C.c_double), ("nrespa", C.c_int), ("thermostat", C.c_int), ("kelvin", C.c_double),
                ("tautemp", C.c_double), ("nfree", C.c_int), ("seed", C.c_ulonglong)]


class MdReport(C.Structure):
    _fields_ = [("epot", C.c_double), ("ekin", C.c_double), ("temp", C.c_double), ("e_valence", C.c_double),
                ("e_nonbonded", C.c_double), ("last_scale", C.c_double), ("steps", C.c_int), ("pcg_iterations", C.c_int),
                ("list_rebuilds", C.c_int), ("total_steps", C.c_longlong), ("ms_device", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("ms_induce", C.c_float), ("ms_energy", C.c_float), ("ms_list", C.c_float), ("ms_ufield_real", C.c_float),
                ("pcg_iterations", C.c_int), ("kernel_launches", C.c_int), ("list_rebuilds", C.c_int),
                ("nverlet", C.c_longlong), ("npairs_m", C.c_longlong), ("npairs_u", C.c_longlong),
                ("ms_ehal", C.c_float), ("nverlet_vdw", C.c_longlong), ("energy_retries", C.c_int)]


_LIBS = {}
_DP = C.POINTER(C.c_double)


def library_path(precision="mixed"):
    # APX_LIBRARY_MIXED: another build of the mixed library (diagnostic variants, tools/diag_precision.py)
    if precision == "mixed" and os.environ.get("APX_LIBRARY_MIXED"):
        return os.environ["APX_LIBRARY_MIXED"]
    return os.path.join(HERE, "libapx.so" if precision == "mixed" else "libapx_f64.so")


def load_library(precision="mixed"):
    """dlopen the C-ABI library; raises (never falls back) when it has not been built."""
    if precision in _LIBS:
        return _LIBS[precision]
    path = library_path(precision)
    if not os.path.isfile(path):
        raise ApxError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback for the CUDA path)")
    lib = C.CDLL(path)
    lib.apx_last_error.restype = C.c_char_p
    lib.apx_version.restype = C.c_char_p
    lib.apx_stream.restype = C.c_void_p
    lib.apx_stream.argtypes = [C.c_void_p]
    lib.apx_create.argtypes = [C.POINTER(_ApxSystem), C.c_int, C.POINTER(C.c_void_p)]
    lib.apx_create_dist.argtypes = [C.POINTER(_ApxSystem), C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_char_p,
                                    C.POINTER(C.c_void_p)]
    lib.apx_nccl_unique_id.argtypes = [C.c_char_p, C.c_void_p]
    lib.apx_local_hub_create.argtypes = [C.c_int]
    lib.apx_local_hub_create.restype = C.c_void_p
    lib.apx_local_hub_destroy.argtypes = [C.c_void_p]
    lib.apx_local_hub_destroy.restype = None
    lib.apx_dist_plan.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_double,
                                  C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.apx_destroy.argtypes = [C.c_void_p]
    lib.apx_destroy.restype = None
    for name, args in {
        "apx_set_positions": [_DP], "apx_set_box": [_DP], "apx_mpole_init": [], "apx_get_rpole": [_DP],
        "apx_dfield": [_DP, _DP], "apx_ufield": [_DP, _DP, _DP, _DP], "apx_precond": [_DP, _DP, _DP, _DP],
        "apx_induce": [], "apx_get_uind": [_DP, _DP], "apx_get_udir": [_DP, _DP],
        "apx_energy": [C.c_int, C.POINTER(EnergyResult)], "apx_empole": [C.c_int, C.POINTER(EnergyResult)],
        "apx_epolar": [C.c_int, C.POINTER(EnergyResult)], "apx_get_gradient": [_DP],
        "apx_pme_mpole_fphi": [_DP], "apx_pme_uind_fphi": [_DP, _DP, _DP, _DP],
        "apx_pme_convolve_grid": [_DP, _DP], "apx_set_native_fft": [C.c_int], "apx_set_pme_fixed_point": [C.c_int],
        "apx_get_stats": [C.POINTER(Stats)], "apx_stats_reset": [], "apx_synchronize": [],
        "apx_get_dist_info": [C.POINTER(C.c_int)],
        "apx_vdw_attach": [C.POINTER(_ApxVdw)], "apx_evdw": [C.c_int, C.POINTER(EnergyResult)],
        "apx_valence_attach": [C.POINTER(_ApxValence)], "apx_evalence": [C.c_int, C.POINTER(ValenceResult)],
        "apx_get_valence_gradient": [_DP],
        "apx_md_init": [_DP, _DP, C.POINTER(MdConfig)], "apx_md_steps": [C.c_int, C.POINTER(MdReport)],
        "apx_md_get_state": [_DP, _DP], "apx_md_set_state": [_DP, _DP, C.c_int],
        "apx_upred_set": [C.c_int], "apx_upred_count": [C.POINTER(C.c_int), C.POINTER(C.c_int)],
        # device-pointer entry points (csrc/devio.cu): pointers as integers, element size, caller's stream
        "apx_set_positions_dev": [C.c_void_p] * 3 + [C.c_int, C.c_void_p],
        "apx_dfield_dev": [C.c_void_p] * 2 + [C.c_int, C.c_void_p],
        "apx_ufield_dev": [C.c_void_p] * 4 + [C.c_int, C.c_void_p],
        "apx_precond_dev": [C.c_void_p] * 4 + [C.c_int, C.c_void_p],
        "apx_induce_dev": [C.c_void_p] * 4 + [C.c_int, C.c_void_p],
        "apx_get_uind_dev": [C.c_void_p] * 4 + [C.c_int, C.c_void_p],
        "apx_add_gradient_dev": [C.c_void_p] * 3 + [C.c_int, C.c_void_p],
        "apx_add_scalars_dev": [C.c_void_p, _DP, C.c_int, C.c_int, C.c_void_p],
    }.items():
        fn = getattr(lib, name)
        fn.argtypes = [C.c_void_p] + args
        fn.restype = C.c_int
    _LIBS[precision] = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(_DP)


def nccl_library_path():
    """libnccl.so.2 bundled with torch (the one torch.distributed has already loaded), else the system one."""
    try:
        import nvidia.nccl as _n
        p = os.path.join(os.path.dirname(_n.__file__), "lib", "libnccl.so.2")
        if os.path.isfile(p):
            return p
    except Exception:
        pass
    return "libnccl.so.2"


def nccl_unique_id(precision="mixed"):
    """128-byte NCCL id made by rank 0; broadcast it to the other ranks and pass it as `handle`."""
    lib = load_library(precision)
    buf = C.create_string_buffer(128)
    if lib.apx_nccl_unique_id(nccl_library_path().encode(), buf) != 0:
        raise ApxError("ncclGetUniqueId failed")
    return buf.raw


class LocalHub:
    """Rendezvous of the in-process transport: `world` ranks = host threads sharing one GPU."""

    def __init__(self, world, precision="mixed"):
        self.lib = load_library(precision)
        self.world = world
        self.handle = self.lib.apx_local_hub_create(world)
        if not self.handle:
            raise ApxError("apx_local_hub_create failed")

    def close(self):
        if self.handle:
            self.lib.apx_local_hub_destroy(self.handle)
            self.handle = None


def dist_plan(w3_sorted, bounds, world, rank, range_frac, precision="mixed"):
    """Host-only halo plan of `rank` (apx_dist_plan): returns (send_idx, send_off, recv_idx, recv_off)."""
    lib = load_library(precision)
    w = np.ascontiguousarray(w3_sorted, dtype=np.float32)
    b = np.ascontiguousarray(bounds, dtype=np.int32)
    n = w.shape[0]
    si, ri = np.zeros(n, np.int32), np.zeros(n, np.int32)
    so, ro = np.zeros(world + 1, np.int32), np.zeros(world + 1, np.int32)
    ip = C.POINTER(C.c_int)
    lib.apx_dist_plan(n, w.ctypes.data_as(C.POINTER(C.c_float)), b.ctypes.data_as(ip), world, rank, float(range_frac),
                      si.ctypes.data_as(ip), so.ctypes.data_as(ip), ri.ctypes.data_as(ip), ro.ctypes.data_as(ip))
    return si[:so[-1]].copy(), so, ri[:ro[-1]].copy(), ro


def system_struct(system):
    """(apx_system filled from a System, list of the numpy arrays it points into) -- what mpoleData / epolarData / pmeData
    upload; also used by the drop-in test scaffolding (oracle/ref_dropin.cpp)."""
    s = _ApxSystem()
    keep = []

    def f64(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return _dp(a)

    def i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_int))

    s.n = system.n
    s.xyz = f64(system.xyz)
    s.lvec = (C.c_double * 9)(*np.asarray(system.lvec, float).ravel())
    s.pole = f64(system.pole)
    s.zaxis = i32(system.zaxis)
    s.polarity = f64(system.polarity)
    s.thole = f64(system.thole)
    s.pdamp = f64(system.pdamp)
    s.jpolar = i32(system.jpolar)
    s.njpolar = int(system.thlval.shape[0])
    s.thlval = f64(system.thlval)
    s.nmdpu = int(system.mdpuexclude.shape[0])
    s.mdpu_ik = i32(system.mdpuexclude.reshape(-1, 2) if s.nmdpu else np.zeros((1, 2)))
    s.mdpu_scale = f64(system.mdpuexclude_scale.reshape(-1, 4) if s.nmdpu else np.ones((1, 4)))
    s.use_ewald, s.use_mpole, s.use_polar = int(system.use_ewald), int(system.use_mpole), int(system.use_polar)
    s.poltyp_mutual = 0 if system.poltyp == "DIRECT" else 1
    s.aewald = system.aewald
    s.nfft = (C.c_int * 3)(*[int(v) for v in system.nfft])
    s.bsorder = system.bsorder
    s.cutoff = float(system.ewald_cutoff)
    s.usolve_cutoff = float(system.usolve_cutoff)
    s.list_buffer = float(system.list_buffer)
    s.poleps, s.politer, s.uaccel = system.poleps, system.politer, system.uaccel
    s.pcgprec, s.pcgguess, s.pcgpeek = int(system.pcgprec), int(system.pcgguess), system.pcgpeek
    s.electric, s.dielec = system.electric, system.dielec
    s.polpred = UPRED[str(getattr(system, "polpred", "NONE") or "NONE").upper()]
    return s, keep


class Amoeba:
    """One electrostatics context on one GPU (the reference's initialize()/finish() pair).

    dist=(rank, world, transport, handle): this context is one rank of a spatially decomposed system on
    `world` GPUs (transport "nccl": handle = nccl_unique_id() of rank 0; "direct": handle = a job id shared by the ranks,
    peer memory without NCCL; "local": handle = LocalHub).
    Every method is then collective (all ranks call it in the same order)."""

    def __init__(self, system: System, precision: str = "mixed", device: int = 0, dist=None, vdw: bool = False,
                 valence: bool = False):
        """vdw=True also attaches the buffered 14-7 term of `system.vdw` (evdwData), valence=True the bonded terms of
        `system.valence` (ebondData ... etortorData): energy() then returns the sum."""
        self.lib = load_library(precision)
        self.system = system
        self.n = system.n
        self.precision = precision
        s, keep = system_struct(system)
        self.ctx = C.c_void_p()
        if dist is None:
            rc = self.lib.apx_create(C.byref(s), device, C.byref(self.ctx))
        else:
            rank, world, transport, handle = dist
            if transport == "local":
                self._hub = handle
                h = C.c_void_p(handle.handle)
                rc = self.lib.apx_create_dist(C.byref(s), device, rank, world, b"local", h, None, C.byref(self.ctx))
            elif transport == "direct":
                # peer memory only (CUDA IPC between the processes of one node), no NCCL: handle = 16+ bytes of job id
                self._idbuf = C.create_string_buffer(bytes(handle).ljust(128, b"\0"), 128)
                rc = self.lib.apx_create_dist(C.byref(s), device, rank, world, b"direct", C.cast(self._idbuf, C.c_void_p),
                                              None, C.byref(self.ctx))
            else:
                self._idbuf = C.create_string_buffer(bytes(handle), 128)
                rc = self.lib.apx_create_dist(C.byref(s), device, rank, world, b"nccl", C.cast(self._idbuf, C.c_void_p),
                                              nccl_library_path().encode(), C.byref(self.ctx))
        if rc != 0:
            msg = self.lib.apx_last_error().decode()
            if self.ctx:
                self.lib.apx_destroy(self.ctx)
                self.ctx = None
            raise ApxError(msg)
        self.last = None
        if vdw:
            self.attach_vdw(system.vdw)
        if valence:
            self.attach_valence(system.valence)

    def attach_vdw(self, v):
        """evdwData(RcOp::ALLOC|INIT), src/evdw.cpp:62-470."""
        if v is None:
            raise ApxError("system has no buffered 14-7 vdW term")
        keep = [np.ascontiguousarray(v.ired, np.int32), np.ascontiguousarray(v.kred, np.float64),
                np.ascontiguousarray(v.jvdw, np.int32), np.ascontiguousarray(v.radmin, np.float64),
                np.ascontiguousarray(v.epsilon, np.float64),
                np.ascontiguousarray(v.vexclude.reshape(-1, 2) if v.vexclude.size else np.zeros((1, 2)), np.int32),
                np.ascontiguousarray(v.vexclude_scale if v.vexclude.size else np.ones(1), np.float64)]
        s = _ApxVdw()
        s.n = self.n
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        s.ired, s.kred, s.jvdw = keep[0].ctypes.data_as(ip), keep[1].ctypes.data_as(dp), keep[2].ctypes.data_as(ip)
        s.njvdw = int(v.radmin.shape[0])
        s.radmin, s.epsilon = keep[3].ctypes.data_as(dp), keep[4].ctypes.data_as(dp)
        s.nvexclude = int(v.vexclude.shape[0]) if v.vexclude.size else 0
        s.vexclude, s.vexclude_scale = keep[5].ctypes.data_as(ip), keep[6].ctypes.data_as(dp)
        s.cutoff, s.taper, s.ghal, s.dhal = float(v.cutoff), float(v.taper), float(v.ghal), float(v.dhal)
        s.elrc_vol, s.vlrc_vol = float(v.elrc_vol), float(v.vlrc_vol)
        self._chk(self.lib.apx_vdw_attach(self.ctx, C.byref(s)))

    def attach_valence(self, v):
        """ebondData ... etortorData (src/bonded/*.cpp): the valence terms of `v` (valparams.ValenceTerms) join energy()."""
        if v is None:
            raise ApxError("system has no valence terms")
        s, keep = valence_struct(v, self.n)
        self._chk(self.lib.apx_valence_attach(self.ctx, C.byref(s)))

    def evalence(self, vers=calc.v1):
        """The bonded terms alone (evalence_cu1, src/cu/evalence.cu): per-term energies and counts, virial_valence; the
        gradient is read with valence_gradient()."""
        r = ValenceResult()
        self._chk(self.lib.apx_evalence(self.ctx, int(vers), C.byref(r)))
        return r

    def valence_gradient(self):
        g = self._out(self.n, 3)
        self._chk(self.lib.apx_get_valence_gradient(self.ctx, _dp(g)))
        return g

    def md_init(self, mass, vel=None, dt=0.002, nrespa=1, thermostat=None, kelvin=298.0, tautemp=0.2, nfree=0, seed=123456789):
        """mdData + integrator kick-off (RespaIntegrator::KickOff, src/md/integrator.cpp:205-222).  dt, tautemp in ps."""
        cfg = MdConfig(float(dt), int(nrespa), {None: 0, "NONE": 0, "BUSSI": 1}[thermostat if thermostat is None else str(thermostat).upper()],
                       float(kelvin), float(tautemp), int(nfree), int(seed))
        m = np.ascontiguousarray(mass, np.float64)
        v = None if vel is None else np.ascontiguousarray(vel, np.float64)
        self._chk(self.lib.apx_md_init(self.ctx, _dp(m), None if v is None else _dp(v), C.byref(cfg)))

    def md_steps(self, nsteps):
        """nsteps x BasicIntegrator::dynamic (src/md/integrator.cpp:70-170); returns the report of the last step."""
        r = MdReport()
        self._chk(self.lib.apx_md_steps(self.ctx, int(nsteps), C.byref(r)))
        return r

    def md_state(self):
        x, v = self._out(self.n, 3), self._out(self.n, 3)
        self._chk(self.lib.apx_md_get_state(self.ctx, _dp(x), _dp(v)))
        return x, v

    def evdw(self, vers=calc.v1):
        """evdw(vers) alone (src/evdw.cpp:472-530)."""
        return self._energy(self.lib.apx_evdw, vers)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.apx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise ApxError(self.lib.apx_last_error().decode())

    def _out(self, *shape):
        return np.zeros(shape, dtype=np.float64)

    # -- state
    def set_positions(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self._chk(self.lib.apx_set_positions(self.ctx, _dp(xyz)))

    def set_box(self, lvec):
        lvec = np.ascontiguousarray(lvec, dtype=np.float64)
        self._chk(self.lib.apx_set_box(self.ctx, _dp(lvec)))

    def mpoleInit(self):
        self._chk(self.lib.apx_mpole_init(self.ctx))

    def rpole(self):
        out = self._out(self.n, 10)
        self._chk(self.lib.apx_get_rpole(self.ctx, _dp(out)))
        return out

    # -- fields / solver
    def dfield(self):
        a, b = self._out(self.n, 3), self._out(self.n, 3)
        self._chk(self.lib.apx_dfield(self.ctx, _dp(a), _dp(b)))
        return a, b

    def ufield(self, uind, uinp):
        uind = np.ascontiguousarray(uind, dtype=np.float64)
        uinp = np.ascontiguousarray(uinp, dtype=np.float64)
        a, b = self._out(self.n, 3), self._out(self.n, 3)
        self._chk(self.lib.apx_ufield(self.ctx, _dp(uind), _dp(uinp), _dp(a), _dp(b)))
        return a, b

    def sparsePrecondApply(self, rsd, rsdp):
        rsd = np.ascontiguousarray(rsd, dtype=np.float64)
        rsdp = np.ascontiguousarray(rsdp, dtype=np.float64)
        a, b = self._out(self.n, 3), self._out(self.n, 3)
        self._chk(self.lib.apx_precond(self.ctx, _dp(rsd), _dp(rsdp), _dp(a), _dp(b)))
        return a, b

    def induce(self):
        self._chk(self.lib.apx_induce(self.ctx))
        return self.uind()

    def upred_set(self, polpred):
        """Select the induced-dipole predictor ("NONE", "ASPC", "GEAR") and empty its history ring."""
        self._chk(self.lib.apx_upred_set(self.ctx, UPRED[str(polpred).upper()]))

    def upred_count(self):
        a, b = C.c_int(), C.c_int()
        self._chk(self.lib.apx_upred_count(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def uind(self):
        a, b = self._out(self.n, 3), self._out(self.n, 3)
        self._chk(self.lib.apx_get_uind(self.ctx, _dp(a), _dp(b)))
        return a, b

    def udir(self):
        a, b = self._out(self.n, 3), self._out(self.n, 3)
        self._chk(self.lib.apx_get_udir(self.ctx, _dp(a), _dp(b)))
        return a, b

    # -- energies
    def _energy(self, fn, vers):
        r = EnergyResult()
        self._chk(fn(self.ctx, int(vers), C.byref(r)))
        out = dict(em=r.em, ep=r.ep, esum=r.esum, virial=np.array(list(r.virial)).reshape(3, 3), nem=r.nem, nep=r.nep,
                   pcg_iterations=r.pcg_iterations, pcg_eps=r.pcg_eps, ev=r.ev, nev=r.nev,
                   evalence=r.evalence, eval_term=np.array(list(r.eval_term)), nval_term=np.array(list(r.nval_term)))
        if vers & calc.grad:
            out["grad"] = self.gradient()
        self.last = out
        return out

    def energy(self, vers=calc.v1):
        return self._energy(self.lib.apx_energy, vers)

    def empole(self, vers=calc.v1):
        return self._energy(self.lib.apx_empole, vers)

    def epolar(self, vers=calc.v1):
        return self._energy(self.lib.apx_epolar, vers)

    def gradient(self):
        g = self._out(self.n, 3)
        self._chk(self.lib.apx_get_gradient(self.ctx, _dp(g)))
        return g

    # -- device-pointer entry points: arguments are torch CUDA tensors (device memory, caller's atom order, float32 or
    #    float64), ordered on torch's current stream -- what the reference's *_cu operators are handed (csrc/devio.cu)
    @staticmethod
    def _dev(*tensors):
        import torch
        eb = None
        for t in tensors:
            if t is None:
                continue
            assert t.is_cuda and t.is_contiguous(), "device-pointer entry points take contiguous CUDA tensors"
            assert t.dtype in (torch.float32, torch.float64)
            assert eb in (None, t.element_size()), "one element type per call"
            eb = t.element_size()
        return eb, torch.cuda.current_stream().cuda_stream

    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()

    def set_positions_dev(self, x, y, z):
        eb, st = self._dev(x, y, z)
        self._chk(self.lib.apx_set_positions_dev(self.ctx, self._p(x), self._p(y), self._p(z), eb, st))

    def dfield_dev(self, field, fieldp):
        eb, st = self._dev(field, fieldp)
        self._chk(self.lib.apx_dfield_dev(self.ctx, self._p(field), self._p(fieldp), eb, st))

    def ufield_dev(self, uind, uinp, field, fieldp):
        eb, st = self._dev(uind, uinp, field, fieldp)
        self._chk(self.lib.apx_ufield_dev(self.ctx, self._p(uind), self._p(uinp), self._p(field), self._p(fieldp), eb, st))

    def precond_dev(self, rsd, rsdp, zrsd, zrsdp):
        eb, st = self._dev(rsd, rsdp, zrsd, zrsdp)
        self._chk(self.lib.apx_precond_dev(self.ctx, self._p(rsd), self._p(rsdp), self._p(zrsd), self._p(zrsdp), eb, st))

    def induce_dev(self, uind, uinp, udir=None, udirp=None):
        eb, st = self._dev(uind, uinp, udir, udirp)
        self._chk(self.lib.apx_induce_dev(self.ctx, self._p(uind), self._p(uinp), self._p(udir), self._p(udirp), eb, st))

    def add_gradient_dev(self, gx, gy, gz):
        """g += dE/dx of the last energy call; int64 tensors = the 2^32 fixed-point grad_prec of the reference's mixed build."""
        import torch
        kind = {torch.int64: 0, torch.float32: 4, torch.float64: 8}[gx.dtype]
        self._chk(self.lib.apx_add_gradient_dev(self.ctx, gx.data_ptr(), gy.data_ptr(), gz.data_ptr(), kind,
                                                torch.cuda.current_stream().cuda_stream))

    def add_scalars_dev(self, dst, vals):
        import torch
        kind = {torch.int64: 0, torch.int32: 1, torch.float32: 4, torch.float64: 8}[dst.dtype]
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self._chk(self.lib.apx_add_scalars_dev(self.ctx, dst.data_ptr(), _dp(vals), len(vals), kind, torch.cuda.current_stream().cuda_stream))

    # -- PME operators (fractional potentials), for parity tests
    def pme_mpole_fphi(self):
        out = self._out(self.n, 20)
        self._chk(self.lib.apx_pme_mpole_fphi(self.ctx, _dp(out)))
        return out

    def pme_uind_fphi(self, uind, uinp):
        uind = np.ascontiguousarray(uind, dtype=np.float64)
        uinp = np.ascontiguousarray(uinp, dtype=np.float64)
        a, b = self._out(self.n, 10), self._out(self.n, 10)
        self._chk(self.lib.apx_pme_uind_fphi(self.ctx, _dp(uind), _dp(uinp), _dp(a), _dp(b)))
        return a, b

    def pme_convolve_grid(self, grid):
        """grid: complex array [nfft3][nfft2][nfft1] -> IFFT(influence * FFT(grid)), unnormalised."""
        g = np.ascontiguousarray(grid, dtype=np.complex128)
        out = np.empty_like(g)
        self._chk(self.lib.apx_pme_convolve_grid(self.ctx, _dp(g.view(np.float64)), _dp(out.view(np.float64))))
        return out

    def set_native_fft(self, on):
        self._chk(self.lib.apx_set_native_fft(self.ctx, int(bool(on))))

    def set_pme_fixed_point(self, on):
        """Deterministic PME spreading (64-bit fixed-point integer sums on the grid); see include/apx.h."""
        self._chk(self.lib.apx_set_pme_fixed_point(self.ctx, int(bool(on))))

    def stats(self):
        s = Stats()
        self._chk(self.lib.apx_get_stats(self.ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def dist_info(self):
        a = (C.c_int * 8)()
        self._chk(self.lib.apx_get_dist_info(self.ctx, a))
        return dict(zip(("rank", "world", "a0", "a1", "halo_atoms", "planes", "halo_lo", "halo_hi"), list(a)))

    def stats_reset(self):
        self._chk(self.lib.apx_stats_reset(self.ctx))

    def synchronize(self):
        self._chk(self.lib.apx_synchronize(self.ctx))
