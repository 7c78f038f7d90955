"""The reference's own CUDA kernels (oracle/_ref/libref_cuda.so = its src/cu/**/*.cu compiled unmodified, oracle/ref_cuda.cu)
run on the GPU on dhfr2 and held to the same float64 oracle fixture as our path -- the comparator of SURVEY section 8(d).
It runs in a child process (the reference keeps its state in process globals and uses the legacy default stream)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.gpu
def test_reference_cuda_kernels_on_dhfr2_match_the_oracle_fixture():
    lib = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
    if not os.path.isfile(lib):
        pytest.skip("oracle/_ref/libref_cuda.so not built (make -C oracle cuda)")
    r = subprocess.run([sys.executable, "-m", "oracle.ref_cuda_bridge", os.path.join(GOLDEN, "dhfr2.npz"), "--fixture",
                        os.path.join(GOLDEN, "dhfr2_oracle.npz"), "--reps", "5", "--warmup", "2"], cwd=ROOT, capture_output=True, text=True,
                       timeout=240)
    assert r.returncode == 0, (r.stderr or r.stdout)[-800:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    print("reference CUDA on dhfr2:", json.dumps(out))
    p = out["parity"]
    # mixed precision (float pair math, fixed-point sums).  First run on a B200 (profiles/r01_refcuda_dhfr2.json): 6.6e-7, 3.9e-7 D,
    # 4.9e-5 kcal/mol/A, 6.1e-7 -- the north-star tolerances, which pins the oracle fixture to the reference's own CUDA build
    assert p["esum_rel"] < 5e-6
    assert p["uind_rms_debye"] < 5e-6
    assert p["grad_rms"] < 5e-4
    assert p["virial_rel"] < 1e-5
    assert out["induce_ms"]["median"] > 0 and out["energy_ms"]["median"] > out["induce_ms"]["median"]


@pytest.mark.gpu
def test_reference_cuda_ehal_on_dhfr2_matches_the_vdw_oracle_fixture():
    lib = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
    if not os.path.isfile(lib):
        pytest.skip("oracle/_ref/libref_cuda.so not built (make -C oracle cuda)")
    r = subprocess.run([sys.executable, "-m", "oracle.ref_cuda_bridge", os.path.join(GOLDEN, "dhfr2.npz"), "--reps", "5", "--warmup", "2",
                        "--vdw", os.path.join(GOLDEN, "dhfr2_vdw_oracle.npz")], cwd=ROOT, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, (r.stderr or r.stdout)[-800:]
    lines = [json.loads(ln) for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    v = lines[-1]["vdw"]
    print("reference CUDA ehal on dhfr2:", json.dumps(v))
    # first run on a B200 (profiles/r01_refcuda_vdw_dhfr2.json): 3.8e-7, 4.8e-5 kcal/mol/A, 3.0e-7; ehal step 0.213 ms
    assert v["parity"]["ev_rel"] < 5e-6
    assert v["parity"]["grad_rms"] < 5e-4
    assert v["parity"]["virial_rel"] < 1e-4
    assert v["ehal_ms"]["median"] > 0


@pytest.mark.gpu
def test_reference_front_ends_run_on_our_kernels_through_the_adapter():
    """tinker::induce / dfield / ufield / sparsePrecondApply / emplar / mpoleInit of the reference's unmodified
    src/amoeba/{induce,field,emplar,mpole}.cpp, linked with the adapter instead of its kernels.  emplar(vers) runs between the
    halves of the reference's energy(): E, virial and gradient come out of the reference's OWN energyReduce / virialReduce and
    gx_elec arrays (src/energy.cpp:345-348,371-374,443-444) and are held against the oracle fixture; a second run on pre-loaded
    accumulators shows the library's contribution is added, not assigned; operators against the C ABI; the front-end induce()
    through the device-pointer entry points costs no more than apx_induce itself (+5 %)."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libref_dropin.so")
    if not os.path.isfile(lib):
        pytest.skip("oracle/_ref/libref_dropin.so not built (make -C oracle dropin)")
    r = subprocess.run([sys.executable, "-m", "oracle.ref_dropin_bridge", os.path.join(GOLDEN, "dhfr2.npz"), "--fixture",
                        os.path.join(GOLDEN, "dhfr2_oracle.npz")], cwd=ROOT, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, (r.stderr or r.stdout)[-800:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    print("reference front-ends on libapx:", json.dumps(out))
    o, c = out["vs_oracle"], out["vs_c_abi"]
    # first run on a B200 (profiles/r01_dropin_dhfr2.json): 7.1e-7 D, 8.0e-7 D, 3.1e-7, 2.5e-9, 7.6e-5 kcal/mol/A; C ABI 4.7e-7
    assert o["uind_rms_debye"] < 2e-6 and o["udir_rms_debye"] < 2e-6      # float round trip of the reference's globals on top of 1e-6
    assert o["esum_rel"] < 1e-6 and o["grad_rms"] < 1e-5 and o["virial_rel"] < 2e-5
    assert max(c.values()) < 1e-6
    acc = out["accumulate"]      # preload 3.0 in slot 0 of eng_buf_elec and in every gx_elec / gy_elec / gz_elec entry
    # two evaluations of -110 556 kcal/mol with float atomics on the PME grid: the delta carries their difference (1e-9 relative)
    assert abs(acc["energy_delta"] - 3.0) < 1e-3
    assert abs(acc["grad_delta_min"] - 3.0) < 1e-4 and abs(acc["grad_delta_max"] - 3.0) < 1e-4     # float atomics in the PME grid
    assert out["ms_induce_frontend"] <= 1.05 * out["ms_induce_c_abi"] + 0.02, out
