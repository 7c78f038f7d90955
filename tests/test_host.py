"""CPU-side tests: readers / parameter assignment against the reference's input decks, the C-ABI
library's exported symbols, and the replica aggregation used for N > 1."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, ROOT

needs_ref = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not mounted (GPU box)")


@needs_ref
def test_blobs_match_fresh_parse():
    """The committed System blobs are what our readers produce from the reference decks today."""
    import tinker_gpu_b200 as tg
    d = tg.load_tinker(os.path.join(REFERENCE, "example/dhfr2.xyz"), os.path.join(REFERENCE, "example/dhfr2.key"))
    b = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    assert d.n == b.n == 23558
    assert d.nfft == b.nfft == (64, 64, 64)
    assert abs(d.aewald - 0.5446) < 1e-4
    for f in ("pole", "zaxis", "polarity", "thole", "pdamp", "mdpuexclude", "mdpuexclude_scale", "xyz"):
        assert np.array_equal(getattr(d, f), getattr(b, f)), f
    assert d.poleps == 1e-5 and d.ewald_cutoff == 7.0 and d.usolve_cutoff == 4.5


@needs_ref
def test_parameter_assignment_local_frames():
    """All five local-frame types are assigned on the 22-atom deck (test/file/local_frame)."""
    import tinker_gpu_b200 as tg
    s = tg.load_tinker(os.path.join(REFERENCE, "test/file/local_frame/local_frame.xyz"), key_text="parameters amoeba09\n",
                       prm_path=os.path.join(REFERENCE, "test/file/commit_6fe8e913/amoeba09.prm"))
    kinds = set(s.zaxis[:, 3].tolist())
    assert kinds == {0, 1, 2, 3, 4, 5}
    # water: O bisector of its two H; H z-then-x
    assert s.zaxis[2].tolist() == [3, 4, 0, 3] and s.zaxis[3].tolist() == [2, 4, 0, 2]
    # ammonium-like N: 3-fold, its H: z-bisect
    assert s.zaxis[14, 3] == 5 and s.zaxis[15, 3] == 4
    assert abs(s.pole[:, 0].sum()) < 1e-9            # NaCl + neutral molecules
    assert s.mexclude.shape[0] == s.dpexclude.shape[0] == 33 and s.uexclude.shape[0] == 0


@needs_ref
@pytest.mark.parametrize("extra", ["octahedron\na-axis 30.0\n", "dodecahedron\na-axis 30.0\n", "ewald\n"])
def test_unsupported_cells_are_refused(extra):
    """Truncated-octahedron / dodecahedron cells and non-periodic Ewald exist in the reference (include/ff/image.h:49-65,
    src/cu/pme.cu:966-969) but not here: the reader must say so instead of running them as an ordinary box."""
    import tinker_gpu_b200 as tg
    with pytest.raises(ValueError):
        tg.load_tinker(os.path.join(REFERENCE, "test/file/local_frame/local_frame.xyz"), key_text="parameters amoeba09\n" + extra,
                       prm_path=os.path.join(REFERENCE, "test/file/commit_6fe8e913/amoeba09.prm"))


def test_polarization_groups_and_scales():
    import tinker_gpu_b200 as tg
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    # each water: O-H, O-H (1-2) and H-H (1-3): m = 0, p = 0, d = 0 (same group)
    nwat = int((s.types == 36).sum())                    # amoeba09 water O; the box also holds a few ions
    assert s.mdpuexclude.shape[0] == 3 * nwat
    assert np.all(s.mdpuexclude_scale[:, 0] == 0) and np.all(s.mdpuexclude_scale[:, 1] == 0)
    assert np.all(s.mdpuexclude_scale[:, 2] == 0) and np.all(s.mdpuexclude_scale[:, 3] == 1)
    assert np.allclose(s.pdamp, s.polarity ** (1 / 6))


def test_replicate_box():
    import tinker_gpu_b200 as tg
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    r = tg.replicate(s, (2, 1, 2))
    assert r.n == 4 * s.n and r.nfft == (72, 36, 72)
    assert np.allclose(np.diag(r.lvec), [60, 30, 60])
    za = r.zaxis[s.n:2 * s.n, 0]
    assert za[za >= 0].min() >= s.n and za.max() < 2 * s.n   # frame atoms offset into their own image (ions have none)
    assert r.mdpuexclude.shape[0] == 4 * s.mdpuexclude.shape[0]
    # replicated periodic system: energy per cell identical (checked on the GPU in test_gpu_scaling)


def test_ewald_and_grid_rules():
    from tinker_gpu_b200.params import ewaldcof, pme_grid_default
    assert abs(ewaldcof(7.0) - 0.5445905) < 1e-6
    assert pme_grid_default(62.23) == 80 and pme_grid_default(30.0) == 36 and pme_grid_default(240.0) == 288


def test_c_abi_exports_every_declared_symbol():
    """include/apx.h is the drop-in boundary: every function it declares is exported by both builds."""
    hdr = open(os.path.join(ROOT, "include", "apx.h")).read()
    decl = set(re.findall(r"\b(apx_[a-z0-9_]+)\s*\(", hdr)) - {"apx_system", "apx_ctx"}
    assert len(decl) >= 20
    for name in ("libapx.so", "libapx_f64.so"):
        path = os.path.join(ROOT, "tinker-gpu_b200", name)
        assert os.path.isfile(path), f"{path} missing: run __graft_entry__.build()"
        lib = ctypes.CDLL(path)
        for sym in decl:
            assert hasattr(lib, sym), (name, sym)
    from tinker_gpu_b200.amoeba import load_library
    assert b"float" in load_library("mixed").apx_version()
    assert b"double" in load_library("double").apx_version()


def test_product_path_does_not_import_oracle():
    """The oracle is test infrastructure only (task statement 3): nothing under the package imports,
    includes, loads or executes anything from oracle/."""
    pkg = os.path.join(ROOT, "tinker-gpu_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s*[\"<][^\">]*oracle)|(oracle[/\\.]amoeba_ref)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not pat.search(txt), (dirpath, f)


def test_no_gpu_raises_not_falls_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, ApxError
    s = tg.load_system(os.path.join(GOLDEN, "lf_local_frame_1.npz"))
    with pytest.raises(ApxError, match="no CPU fallback"):
        Amoeba(s)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import bench
dist.init_process_group("gloo")
rank = dist.get_rank()
ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print("RESULT", float(ms), bench.ns_per_day(float(ms), dist.get_world_size()))
dist.destroy_process_group()
"""


def test_replica_aggregation_world2_gloo(tmp_path):
    """N > 1 is independent replicas: step time = max over ranks, value = replicas x per-replica rate."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][0].split()
    assert float(line[1]) == 15.0
    assert abs(float(line[2]) - 2 * 2.0e-6 * 86400e3 / 15.0) < 1e-12


def test_bench_workloads_build():
    """bench.py --workload: the synthetic boxes of BASELINE.md section 4 come out with the stated sizes and grids, and the
    slab decomposition's divisibility rule holds for the GPU counts the 1 M-atom box is meant for."""
    import bench
    w = bench.make_system("water96k")
    assert w.n == 96624 and w.nfft == (108, 108, 144) and w.poleps == 1e-8
    assert w.vdw is not None and w.vdw.ired.max() < w.n
    m = bench.make_system("water1m")
    assert m.n == 1030656 and m.nfft == (288, 288, 216)
    for g in (2, 4, 8):
        assert m.nfft[1] % g == 0 and m.nfft[2] % g == 0
    d = bench.make_system("dhfr424k")
    assert d.n == 424044 and d.nfft == (240, 240, 150)
