"""Host side of DYNAMIC: restart-file and archive writers against the reference's own files, transcript formats of
mdsave.f / xdynamic.cpp, mdinit rules, the kinetic-energy golden of test/kinetic.cpp."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE

needs_ref = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present (GPU box)")


def test_kinetic_golden_arbox():
    """test/kinetic.cpp:34-38: eksum 100446.40376, temp 156008.001336 for the velocities of arbox.dyn_2
    (fixture tests/golden/arbox_dyn2.npz, made by make_valence_golden.py)."""
    from oracle import md_ref
    z = np.load(os.path.join(GOLDEN, "arbox_dyn2.npz"))
    ek, t = md_ref.kinetic(z["vel"], z["mass"], 3 * len(z["mass"]))
    assert abs(ek - 100446.40376) < 1e-4 and abs(t - 156008.001336) < 1e-4


@needs_ref
def test_dyn_writer_reproduces_reference_file(tmp_path):
    from tinker_gpu_b200.tinkerio import read_dyn, write_dyn
    src = os.path.join(REFERENCE, "test/file/arbox/arbox.dyn_2")
    d = read_dyn(src)
    out = str(tmp_path / "a.dyn")
    write_dyn(out, d["title"], d["box"], d["xyz"], d["vel"], d["acc"], d["aalt"])
    assert open(out).read() == open(src).read()


def test_dyn_round_trip(tmp_path):
    from tinker_gpu_b200.tinkerio import read_dyn, write_dyn
    rng = np.random.default_rng(0)
    x, v, a = (rng.normal(size=(7, 3)) * s for s in (10, 50, 1e3))
    x[0, 0], v[0, 1] = 0.0, -1.234e-7
    p = str(tmp_path / "t.dyn")
    write_dyn(p, "seven atoms", [20.0, 21.0, 22.0, 90.0, 90.0, 90.0], x, v, a)
    d = read_dyn(p)
    assert d["n"] == 7 and d["title"] == "seven atoms" and d["box"][:3] == [20.0, 21.0, 22.0]
    for k, ref in (("xyz", x), ("vel", v), ("acc", a)):
        assert np.abs(d[k] - ref).max() <= 1e-15 * np.abs(ref).max()
    assert np.all(d["aalt"] == 0)


@needs_ref
def test_arc_frame_matches_reference_archive(tmp_path):
    """The first frame of test/file/arbox/arbox.arc, rewritten by our writer, is identical text."""
    from tinker_gpu_b200.tinkerio import append_arc_frame, read_xyz
    src = os.path.join(REFERENCE, "test/file/arbox/arbox.arc")
    lines = open(src).read().splitlines()
    n = int(lines[0].split()[0])
    first = "\n".join(lines[:n + 2]) + "\n"
    p = tmp_path / "one.xyz"
    p.write_text(first)
    x = read_xyz(str(p))
    out = str(tmp_path / "o.arc")
    append_arc_frame(out, x)
    append_arc_frame(out, x)
    assert open(out).read() == first + first


def test_mdinit_rules_and_formats():
    from tinker_gpu_b200 import drivers as d
    from tinker_gpu_b200.tinkerio import read_key
    assert d.respa_inner_steps(0.001) == 2          # test/respa.cpp:39
    assert d.respa_inner_steps(0.002) == 4
    o = d.md_options(read_key(None, text="integrator respa\nthermostat bussi\ntau-temperature 0.1\nrespa-inner 8\n"))
    assert o["integrator"] == "RESPA" and o["thermostat"] == "BUSSI" and o["tautemp"] == 0.1 and o["nrespa"] == 8
    assert d.md_options(None)["integrator"] == "BEEMAN"          # xdynamic.cpp:36
    m = np.full(2000, 15.999)
    v = d.maxwell_velocities(m, 298.0, 5)
    from oracle import md_ref
    assert abs(md_ref.kinetic(v, m, 3 * 2000 - 3)[1] - 298.0) < 1e-9
    assert np.abs((m[:, None] * v).sum(0)).max() < 1e-9
    txt = d.format_md_frame(100, 0.002, -1234.5678, 456.789, [62.23] * 3 + [90.0] * 3, 1, "dhfr2.arc")
    assert " Current Time                 0.2000 Picosecond\n" in txt
    assert " Current Potential        -1234.5678 Kcal/mole\n" in txt
    assert " Current Kinetic            456.7890 Kcal/mole\n" in txt
    assert " Lattice Lengths           62.230000     62.230000     62.230000\n" in txt
    assert " Frame Number                      1\n" in txt
    perf = d.format_md_performance(123.4567, 1.5, 1000, 10, 2.0, 23558)
    assert " Performance:  ns/day             123.4567\n" in perf and "               Atoms                 23558\n" in perf


def test_unbuilt_options_are_refused():
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.drivers import dynamic
    s = tg.load_system(os.path.join(GOLDEN, "val_water10.npz"))
    with pytest.raises(NotImplementedError):
        dynamic(s, 1, integrator="BEEMAN")
    with pytest.raises(NotImplementedError):
        dynamic(s, 1, mode=4)
