"""The ctypes mirrors in tinker-gpu_b200/amoeba.py against include/apx.h as the C compiler lays it out: size of every
struct and offset of every field.  A mismatch here would corrupt arguments silently on the GPU box."""
import ctypes as C
import importlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
am = importlib.import_module("tinker-gpu_b200.amoeba")

PAIRS = {"apx_system": am._ApxSystem, "apx_vdw": am._ApxVdw, "apx_valence": am._ApxValence,
         "apx_energy_result": am.EnergyResult, "apx_valence_result": am.ValenceResult, "apx_stats": am.Stats,
         "apx_md_config": am.MdConfig, "apx_md_report": am.MdReport}


def test_struct_layouts_match_header(tmp_path):
    lines = ['#include "apx.h"', "#include <stdio.h>", "#include <stddef.h>", "int main(void){"]
    for cname, mirror in PAIRS.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe])
    out = subprocess.check_output([exe], text=True)
    seen = 0
    for ln in out.splitlines():
        cname, field, val = ln.split()
        mirror = PAIRS[cname]
        if field == "size":
            assert C.sizeof(mirror) == int(val), cname
        else:
            assert getattr(mirror, field).offset == int(val), f"{cname}.{field}"
        seen += 1
    assert seen == sum(len(m._fields_) + 1 for m in PAIRS.values())


def test_header_structs_all_mirrored():
    hdr = open(os.path.join(ROOT, "include", "apx.h")).read()
    names = set(re.findall(r"typedef struct (\w+) \{", hdr))
    assert names == set(PAIRS), names ^ set(PAIRS)
