"""Scrape the reference's valence-term goldens and build the matching term lists with OUR readers.

    python tests/golden/make_valence_golden.py [/root/reference]

Outputs, all under tests/golden/:
  valence_goldens.json      energy, interaction count, internal virial and per-atom gradient printed in
                            test/ref/{bond,angle.1,strbnd,urey,opbend,torsion,pitors,tortor}.txt
  val_trpcage.npz           Trp-cage System blob (with ValenceTerms) from test/file/commit_6fe8e913/amoebapro13.prm
  val_trpcage_angle.npz     same deck with test/file/commit_291a85c1/amoebapro13.prm (test/angle.cpp:16)
  val_water10.npz           test/file/water10/h2o10.xyz + commit_6fe8e913/water03.prm (test/urey.cpp:14-16)
  arbox_dyn2.npz            velocities of test/file/arbox/arbox.dyn_2 (216 argon atoms) for the kinetic-energy golden
  dhfr2.npz, water30.npz    rewritten with the valence lists attached (other fields unchanged)
(test/ref/angle.2.txt is the Fourier-angle case of another force field and is not built.)
"""
import importlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
import tinker_gpu_b200 as tg  # noqa: E402


def read_ref(path):
    txt = open(path).read()
    m = re.search(r"^ (\S.*?)\s+(-?\d+\.\d+)\s+(\d+)\s*$", txt, re.M)
    vm = re.search(r"Internal Virial Tensor :\s+(.*?)\n\s*\n", txt, re.S)
    g = [[float(a) for a in ln.split()[2:5]] for ln in txt.splitlines() if ln.startswith(" Anlyt")]
    return {"energy": float(m.group(2)), "count": int(m.group(3)),
            "virial": [float(t) for t in vm.group(1).split()], "grad": g}


def blob(xyzfile, prm, keytext, out):
    """Full System (electrostatics + vdW + valence) so that the GPU tests can open a context on it.  The reference
    decks of these cases are gas-phase; the valence terms do not see the box we add for the electrostatics setup."""
    s = tg.load_tinker(os.path.join(REF, xyzfile), key_text=keytext, prm_path=os.path.join(REF, prm))
    assert s.valence is not None
    tg.save_system(os.path.join(HERE, out), s)
    return s


BOX = "a-axis 80.0\newald\newald-cutoff 7.0\n"
CASES = {"bond": ("bond.txt", "val_trpcage.npz"), "angle": ("angle.1.txt", "val_trpcage_angle.npz"),
         "strbnd": ("strbnd.txt", "val_trpcage.npz"), "urey": ("urey.txt", "val_water10.npz"),
         "opbend": ("opbend.txt", "val_trpcage.npz"), "torsion": ("torsion.txt", "val_trpcage.npz"),
         "pitors": ("pitors.txt", "val_trpcage.npz"), "tortor": ("tortor.txt", "val_trpcage.npz")}

if __name__ == "__main__":
    blob("test/file/trpcage/trpcage.xyz", "test/file/commit_6fe8e913/amoebapro13.prm", "parameters amoebapro13\n" + BOX, "val_trpcage.npz")
    blob("test/file/trpcage/trpcage.xyz", "test/file/commit_291a85c1/amoebapro13.prm", "parameters amoebapro13\n" + BOX, "val_trpcage_angle.npz")
    blob("test/file/water10/h2o10.xyz", "test/file/commit_6fe8e913/water03.prm", "parameters water03\n" + BOX, "val_water10.npz")
    d = tg.load_tinker(os.path.join(REF, "example/dhfr2.xyz"), os.path.join(REF, "example/dhfr2.key"))
    tg.save_system(os.path.join(HERE, "dhfr2.npz"), d)
    print("dhfr2:", {t: d.valence.count(t) for t in ("bond", "angle", "strbnd", "urey", "opbend", "torsion", "pitors", "tortor")})
    w = tg.load_tinker(os.path.join(REF, "test/file/tinkernist/water30.xyz"), os.path.join(REF, "test/file/tinkernist/water30.key"),
                       prm_path=os.path.join(REF, "test/file/commit_6fe8e913/amoeba09.prm"))
    tg.save_system(os.path.join(HERE, "water30.npz"), w)
    io = importlib.import_module("tinker-gpu_b200.tinkerio")
    dyn = io.read_dyn(os.path.join(REF, "test/file/arbox/arbox.dyn_2"))
    np.savez_compressed(os.path.join(HERE, "arbox_dyn2.npz"), vel=dyn["vel"], mass=np.full(dyn["n"], 39.948))
    gold = {t: dict(read_ref(os.path.join(REF, "test/ref", f)), blob=b, source="test/ref/" + f) for t, (f, b) in CASES.items()}
    json.dump(gold, open(os.path.join(HERE, "valence_goldens.json"), "w"))
