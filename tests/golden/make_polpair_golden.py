#!/usr/bin/env python
"""tests/golden/polpair_{ewald,nonewald}.npz + polpair_goldens.json from the reference's test/polpair.cpp: the NaCl pair with a
POLPAIR-specific Thole width (`polpair 7 15 0.05`, test/file/polpair/nacl.key), with and without Ewald -- System blobs built
by OUR readers from the reference's deck, and the literals of test/ref/polpair.{1,2}.txt (total energy, virial, gradient).
Run HERE, never on the GPU box:  python tests/golden/make_polpair_golden.py [/root/reference]"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import tinker_gpu_b200 as tg  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
PRM09 = os.path.join(REF, "test/file/commit_6fe8e913/amoeba09.prm")


def transcript(path):
    t = open(path).read()
    e = float(re.search(r"Total Potential Energy :\s+([-\d.]+)", t).group(1))
    v = [float(x) for x in re.findall(r"[-]?\d+\.\d+", t.split("Internal Virial Tensor :")[1].split("Cartesian")[0])]
    g = [[float(x) for x in m.groups()] for m in re.finditer(r"Anlyt\s+\d+\s+([-\d.]+)\s+([-\d.]+)\s+([-\d.]+)", t)]
    return dict(energy=e, virial=np.array(v).reshape(3, 3).tolist(), gradient=g)


out = {}
key = open(os.path.join(REF, "test/file/polpair/nacl.key")).read()
for name, extra, ref in (("polpair_ewald", "\newald\n", "polpair.1.txt"), ("polpair_nonewald", "", "polpair.2.txt")):
    s = tg.load_tinker(os.path.join(REF, "test/file/polpair/nacl.xyz"), key_text=key + extra, prm_path=PRM09)
    tg.save_system(os.path.join(HERE, name + ".npz"), s)
    out[name] = dict(transcript(os.path.join(REF, "test/ref", ref)), source="test/polpair.cpp, test/ref/" + ref, key=key + extra,
                     tolerance=dict(energy=1e-4, gradient=1e-4, virial=1e-3))
    print(name, "n", s.n, "ewald", s.use_ewald, "thlval", np.unique(s.thlval), out[name]["energy"])
json.dump(out, open(os.path.join(HERE, "polpair_goldens.json"), "w"), indent=1)
