#!/usr/bin/env python
"""tests/golden/arbox.npz + arbox_verlet.npz from the reference's test/verlet.cpp (NVE-Verlet-ArBox): 216 argon atoms, buffered
14-7 vdW only, velocity Verlet at 1 fs from the restart file test/file/arbox/arbox.dyn -- the System blob built by OUR readers,
the initial positions / velocities, and the reference's literal potential and kinetic energies of the following steps.
Run HERE, never on the GPU box:  python tests/golden/make_arbox_verlet_golden.py [/root/reference]"""
import importlib
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import tinker_gpu_b200 as tg  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
io = importlib.import_module("tinker-gpu_b200.tinkerio")
key = open(os.path.join(REF, "test/file/arbox/arbox.key")).read() + "\nintegrator  verlet\n"
s = tg.load_tinker(os.path.join(REF, "test/file/arbox/arbox.xyz"), key_text=key, prm_path=os.path.join(REF, "test/file/commit_6fe8e913/amoeba09.prm"))
d = io.read_dyn(os.path.join(REF, "test/file/arbox/arbox.dyn"))
src = open(os.path.join(REF, "test/verlet.cpp")).read()
kin = [float(x) for x in re.findall(r"-?\d+\.\d+", src.split("arbox_kin[] = {")[1].split("}")[0])]
pot = [float(x) for x in re.findall(r"-?\d+\.\d+", src.split("arbox_pot[] = {")[1].split("}")[0])]
tg.save_system(os.path.join(HERE, "arbox.npz"), s)
np.savez_compressed(os.path.join(HERE, "arbox_verlet.npz"), xyz=d["xyz"], vel=d["vel"], arbox_pot=np.array(pot), arbox_kin=np.array(kin),
                    dt_ps=0.001, nsteps_checked=20, eps=1e-4)
print(s.n, len(pot), len(kin))
