#!/usr/bin/env python
"""Golden fixtures of the buffered 14-7 vdW term from the reference tree (run HERE, not on the GPU box).

    python tests/golden/make_vdw_golden.py [/root/reference]

  vdw_goldens.json      literals of the reference's own tests: NaCl-1 (test/nacl.cpp:36-176: energy,
                        gradient, virial, count; three separations; two vdw-correction variants) and
                        Local-Frame2-1/2 (test/localframe2.cpp:46-98: energy + count, triclinic / monoclinic)
  vdw_<case>.npz        System blobs built by OUR readers from the reference's decks for those cases
  dhfr2.npz             rewritten with the vdW term attached (other fields unchanged)
Only numeric literals of test expectations and systems parsed from input decks are stored.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import tinker_gpu_b200 as tg  # noqa: E402
from make_golden import scrape, PRM09, REF  # noqa: E402


def main():
    gold = {}
    nacl = scrape(os.path.join(REF, "test/nacl.cpp"))["NaCl-1"]
    lf2 = scrape(os.path.join(REF, "test/localframe2.cpp"))
    base_key = open(os.path.join(REF, "test/file/nacl/nacl.key")).read() + "\nvdwterm    only\n"
    xyz_of = {"no switch": "nacl1.xyz", "switch, near cut": "nacl2.xyz", "switch, near off": "nacl3.xyz",
              "evcorr vlambda = 1.0": "nacl1.xyz"}
    for sname, sec in nacl["sections"].items():
        for frag, xf in xyz_of.items():
            if frag in sname:
                key = base_key
                if "evcorr" in frag:
                    key += "b-axis            10.0\nc-axis            10.0\nvdw-correction\n"
                tag = "nacl_" + frag.split(",")[-1].strip().replace(" ", "_").replace("=", "").replace(".", "")
                s = tg.load_tinker(os.path.join(REF, "test/file/nacl", xf), key_text=key, prm_path=PRM09)
                tg.save_system(os.path.join(HERE, "vdw_" + tag + ".npz"), s)
                gold[tag] = {k: sec[k] for k in ("ref_eng", "ref_count", "ref_grad", "ref_v")}
                print(tag, s.n, s.vdw.cutoff, s.vdw.taper, gold[tag]["ref_eng"])
    lfkey = open(os.path.join(REF, "test/file/local_frame/local_frame.key")).read()
    for name in ("Local-Frame2-1", "Local-Frame2-2"):
        c = lf2[name]
        key = lfkey + "\n" + c["key"] + "\nvdwterm  only\n"
        s = tg.load_tinker(os.path.join(REF, "test/file/local_frame/local_frame2.xyz"), key_text=key, prm_path=PRM09)
        tag = name.lower().replace("-", "_")
        tg.save_system(os.path.join(HERE, "vdw_" + tag + ".npz"), s)
        sec = [v for k, v in c["sections"].items() if "ehal" in k][0]
        gold[tag] = {k: sec[k] for k in ("ref_eng", "ref_count")}
        print(tag, s.n, s.lvec.tolist(), gold[tag])
    with open(os.path.join(HERE, "vdw_goldens.json"), "w") as fh:
        json.dump(gold, fh, indent=0, sort_keys=True)
    d = tg.load_tinker(os.path.join(REF, "example/dhfr2.xyz"), os.path.join(REF, "example/dhfr2.key"))
    tg.save_system(os.path.join(HERE, "dhfr2.npz"), d)
    w = tg.load_tinker(os.path.join(REF, "test/file/tinkernist/water30.xyz"),
                       os.path.join(REF, "test/file/tinkernist/water30.key"), prm_path=PRM09)
    tg.save_system(os.path.join(HERE, "water30.npz"), w)
    print("dhfr2 / water30 rewritten with vdw:", d.vdw.radmin.shape, w.vdw.radmin.shape)


if __name__ == "__main__":
    main()
