#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference tree (run HERE, never on the GPU box).

    python tests/golden/make_golden.py [/root/reference]

Outputs, all under tests/golden/:
  reference_goldens.json   literals scraped from the reference's own Catch2 tests
                           (test/localframe.cpp, test/localframe3.cpp) -- energies, fields,
                           induced dipoles, gradients, virials, with the keyfile each case used
  lf_<case>.npz            System blobs for those cases, built by OUR readers from the
                           reference's .xyz/.key/.prm inputs
  tinkernist.npz           2684-atom water box System + the reference's two MD frames
                           (coordinates, uind, udir in Debye) from test/ref/tinkernist.*
  dhfr2.npz, water30.npz   System blobs for the benchmark configurations

Nothing here copies reference source: only numeric literals of its test expectations and
systems parsed from its input decks.
"""
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


def _num_list(body):
    return [float(t) for t in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", body)]


def scrape(cpp_path):
    """TEST_CASE -> {key, arrays/scalars by SECTION} from a Catch2 source file."""
    src = open(cpp_path).read()
    cases = {}
    raw = {m.group(1): m.group(2) for m in re.finditer(r'(\w+)\s*=\s*R"\*\*\((.*?)\)\*\*"', src, re.S)}
    heads = [(m.start(), m.group(1)) for m in re.finditer(r'TEST_CASE\("([^"]+)"', src)]
    heads.append((len(src), None))
    for (a, name), (b, _) in zip(heads[:-1], heads[1:]):
        blk = src[a:b]
        case = {"key": "", "sections": {}}
        for rname, rtext in raw.items():
            if re.search(r"=\s*" + rname + r"\s*;", blk):
                case["key"] += rtext
        # keyfile text: std::string key1 = "..."; key1 += "...";  (first variable wins per case)
        for m in re.finditer(r'(?:std::string\s+)?(key\w*)\s*(\+?=)\s*"((?:[^"\\]|\\.)*)"\s*;', blk):
            txt = m.group(3).encode().decode("unicode_escape")
            case["key"] += txt
        secs = [(m.start(), m.group(1)) for m in re.finditer(r'SECTION\("([^"]+)"\)', blk)]
        bounds = [(0, "")] + secs + [(len(blk), None)]
        for (sa, sname), (sb, _) in zip(bounds[:-1], bounds[1:]):
            part = blk[sa:sb]
            d = case["sections"].setdefault(sname, {})
            for m in re.finditer(r"const\s+double\s+(\w+)\[\]\[3\]\s*=\s*\{(.*?)\};", part, re.S):
                v = _num_list(m.group(2))
                d[m.group(1)] = np.array(v).reshape(-1, 3).tolist()
            for m in re.finditer(r"const\s+(?:double|int)\s+(\w+)\s*=\s*([-+]?\d+\.?\d*(?:[eE][-+]?\d+)?)\s*;", part):
                d[m.group(1)] = float(m.group(2))
        cases[name] = case
    return cases


def read_frames(path, n):
    """Tinker archive-style file -> (nframes, n, 3)."""
    rows = []
    for ln in open(path):
        t = ln.split()
        if len(t) >= 6 and t[0].isdigit() and t[5].isdigit():
            try:
                rows.append([float(t[2]), float(t[3]), float(t[4])])
            except ValueError:
                pass
    a = np.array(rows)
    return a.reshape(-1, n, 3)


def main():
    gold = {}
    gold.update(scrape(os.path.join(REF, "test/localframe.cpp")))
    gold.update(scrape(os.path.join(REF, "test/localframe3.cpp")))
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as fh:
        json.dump(gold, fh, indent=0, sort_keys=True)
    print("scraped", list(gold))

    xyz_of = {"Local-Frame-1": "local_frame.xyz", "Local-Frame-2": "local_frame.xyz", "Local-Frame-3": "local_frame.xyz",
              "Local-Frame-4": "local_frame.xyz", "Local-Frame3-1": "local_frame2.xyz", "Local-Frame3-2": "local_frame2.xyz",
              "Local-Frame3-3": "local_frame2.xyz"}
    for name, xf in xyz_of.items():
        if name not in gold:
            continue
        key = gold[name]["key"]
        if "parameters" not in key:
            key = "parameters  amoeba09\n" + key
        s = tg.load_tinker(os.path.join(REF, "test/file/local_frame", xf), key_text=key, prm_path=PRM09)
        tg.save_system(os.path.join(HERE, "lf_" + name.lower().replace("-", "_") + ".npz"), s)
        print(name, s.n, "ewald" if s.use_ewald else "nonewald", "mpole", s.use_mpole, "polar", s.use_polar)

    # 2684-atom water box with the reference's MD frames
    w = tg.load_tinker(os.path.join(REF, "test/file/tinkernist/water30.xyz"),
                       os.path.join(REF, "test/file/tinkernist/water30.key"), prm_path=PRM09)
    tg.save_system(os.path.join(HERE, "water30.npz"), w)
    fr = {k: read_frames(os.path.join(REF, "test/ref/tinkernist." + k), w.n) for k in ("arc", "uind", "udir")}
    np.savez_compressed(os.path.join(HERE, "tinkernist_frames.npz"), **fr)
    print("tinkernist frames", {k: v.shape for k, v in fr.items()})

    d = tg.load_tinker(os.path.join(REF, "example/dhfr2.xyz"), os.path.join(REF, "example/dhfr2.key"))
    tg.save_system(os.path.join(HERE, "dhfr2.npz"), d)
    print("dhfr2", d.n, d.nfft, d.aewald)


if __name__ == "__main__":
    main()
