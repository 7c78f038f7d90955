"""Scrape the reference's valence-term goldens and build the matching term lists with OUR readers.

    python tests/golden/make_valence_golden.py [/root/reference]

Outputs, all under tests/golden/:
  valence_goldens.json      energy, interaction count, internal virial and per-atom gradient printed in
                            test/ref/{bond,angle.1,strbnd,urey,opbend,torsion,pitors,tortor}.txt
  val_trpcage.npz           Trp-cage coordinates + ValenceTerms from test/file/commit_6fe8e913/amoebapro13.prm
  val_trpcage_angle.npz     same deck with test/file/commit_291a85c1/amoebapro13.prm (test/angle.cpp:16)
  val_water10.npz           test/file/water10/h2o10.xyz + commit_6fe8e913/water03.prm (test/urey.cpp:14-16)
  val_dhfr2.npz             example/dhfr2 valence lists (amoebabio09) for the full-size parity / property tests
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
io = importlib.import_module("tinker-gpu_b200.tinkerio")
vp = importlib.import_module("tinker-gpu_b200.valparams")


def read_ref(path):
    txt = open(path).read()
    m = re.search(r"^ (\S.*?)\s+(-?\d+\.\d+)\s+(\d+)\s*$", txt, re.M)
    vm = re.search(r"Internal Virial Tensor :\s+(.*?)\n\s*\n", txt, re.S)
    g = [[float(a) for a in ln.split()[2:5]] for ln in txt.splitlines() if ln.startswith(" Anlyt")]
    return {"energy": float(m.group(2)), "count": int(m.group(3)),
            "virial": [float(t) for t in vm.group(1).split()], "grad": g}


def blob(xyzfile, prm, keytext, out):
    xyz = io.read_xyz(os.path.join(REF, xyzfile))
    ff = io.read_prm(os.path.join(REF, prm))
    key = io.read_key(None, text=keytext)
    v = vp.build_valence(xyz.n, xyz.types, ff.atom_class, ff.atom_atomic, [list(b) for b in xyz.bonds], key, ff)
    mass = np.array([ff.atom_mass.get(int(t), 0.0) for t in xyz.types])
    np.savez_compressed(os.path.join(HERE, out), xyz=xyz.xyz, mass=mass, **vp.valence_to_dict(v))
    return v


CASES = {"bond": ("bond.txt", "val_trpcage.npz"), "angle": ("angle.1.txt", "val_trpcage_angle.npz"),
         "strbnd": ("strbnd.txt", "val_trpcage.npz"), "urey": ("urey.txt", "val_water10.npz"),
         "opbend": ("opbend.txt", "val_trpcage.npz"), "torsion": ("torsion.txt", "val_trpcage.npz"),
         "pitors": ("pitors.txt", "val_trpcage.npz"), "tortor": ("tortor.txt", "val_trpcage.npz")}

if __name__ == "__main__":
    blob("test/file/trpcage/trpcage.xyz", "test/file/commit_6fe8e913/amoebapro13.prm", "parameters amoebapro13\n", "val_trpcage.npz")
    blob("test/file/trpcage/trpcage.xyz", "test/file/commit_291a85c1/amoebapro13.prm", "parameters amoebapro13\n", "val_trpcage_angle.npz")
    blob("test/file/water10/h2o10.xyz", "test/file/commit_6fe8e913/water03.prm", "parameters water03\n", "val_water10.npz")
    v = blob("example/dhfr2.xyz", "params/amoebabio09.prm", open(os.path.join(REF, "example/dhfr2.key")).read(), "val_dhfr2.npz")
    print("dhfr2:", {t: v.count(t) for t in vp.TERMS})
    gold = {t: dict(read_ref(os.path.join(REF, "test/ref", f)), blob=b, source="test/ref/" + f) for t, (f, b) in CASES.items()}
    json.dump(gold, open(os.path.join(HERE, "valence_goldens.json"), "w"))
