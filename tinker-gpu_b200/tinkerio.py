"""Readers for the Tinker input formats the AMOEBA electrostatics path consumes.

Fortran-free restatement of the parts of the Tinker library that the reference
calls before `initialize()`:

* `.xyz`  -- tinker/source/readxyz.f  (optional 2nd-line box record)
* `.key`  -- tinker/source/getkey.f   (keyword lines, case-insensitive)
* `.prm`  -- tinker/source/readprm.f:208 (atom), :1215 (multipole), :1297 (polarize),
             :1340 (polpair) and the `*-scale` force-field headers.

Only the records used by the multipole/polarization path are interpreted; all
other records are kept as raw keyword lines so a caller can inspect them.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class XYZ:
    n: int
    title: str
    names: list
    xyz: np.ndarray          # (n,3) float64, Angstrom
    types: np.ndarray        # (n,) int32 Tinker atom type
    bonds: list              # per-atom list of 0-based bonded atoms (i12)
    box: tuple | None = None  # (a,b,c,alpha,beta,gamma) if 2nd line held a box


def _is_float(tok: str) -> bool:
    try:
        float(tok.replace("D", "E").replace("d", "e"))
        return True
    except ValueError:
        return False


def read_xyz(path: str) -> XYZ:
    """Parse a Tinker Cartesian coordinate file (readxyz.f)."""
    with open(path) as fh:
        lines = [ln.rstrip("\n") for ln in fh]
    k = 0
    while not lines[k].strip():
        k += 1
    head = lines[k].split(None, 1)
    n = int(head[0])
    title = head[1].strip() if len(head) > 1 else ""
    k += 1
    box = None
    # optional periodic box record: six reals and no atom name
    toks = lines[k].split()
    if len(toks) == 6 and all(_is_float(t) for t in toks) and "." in toks[1]:
        box = tuple(float(t) for t in toks)
        k += 1
    names, types, bonds = [], np.zeros(n, np.int32), []
    xyz = np.zeros((n, 3))
    i = 0
    while i < n:
        toks = lines[k].split()
        k += 1
        if not toks:
            continue
        names.append(toks[1])
        xyz[i] = [float(t.replace("D", "E")) for t in toks[2:5]]
        types[i] = int(toks[5])
        bonds.append(sorted(int(t) - 1 for t in toks[6:]))
        i += 1
    return XYZ(n, title, names, xyz, types, bonds, box)


def write_xyz(path: str, sysxyz: XYZ) -> None:
    """Write a Tinker .xyz (prtxyz.f layout); used for synthetic boxes."""
    with open(path, "w") as fh:
        fh.write(f"{sysxyz.n:6d}  {sysxyz.title}\n")
        if sysxyz.box is not None:
            fh.write(" " + "".join(f"{v:12.6f}" for v in sysxyz.box) + "\n")
        for i in range(sysxyz.n):
            x, y, z = sysxyz.xyz[i]
            conn = "".join(f"{b + 1:6d}" for b in sysxyz.bonds[i])
            fh.write(f"{i + 1:6d}  {sysxyz.names[i]:<3s}{x:12.6f}{y:12.6f}{z:12.6f}{sysxyz.types[i]:6d}{conn}\n")


def append_arc_frame(path: str, sysxyz: XYZ) -> None:
    """One more frame at the end of a Tinker archive (mdsave.f:256-266 -> prtxyz): the .xyz record repeated."""
    import tempfile
    with tempfile.NamedTemporaryFile("r", suffix=".xyz") as tmp:
        write_xyz(tmp.name, sysxyz)
        frame = open(tmp.name).read()
    with open(path, "a") as fh:
        fh.write(frame)


def _fortran_d(v: float, width: int = 26, digits: int = 16) -> str:
    """Fortran D<width>.<digits> edit descriptor: 0.dddddD+ee."""
    if v == 0.0:
        body = "0." + "0" * digits + "D+00"
    else:
        m, e = f"{abs(v):.{digits - 1}E}".split("E")
        mant = m.replace(".", "")
        body = ("-" if v < 0 else "") + "0." + mant + "D" + f"{int(e) + 1:+03d}"
    return body.rjust(width)


def write_dyn(path: str, title: str, box6, xyz, vel, acc, aalt=None) -> None:
    """Restart file of a trajectory (tinker/source/prtdyn.f): positions, velocities, accelerations, alternate
    accelerations, each as 3D26.16 per atom."""
    n = len(xyz)
    aalt = np.zeros_like(np.asarray(xyz)) if aalt is None else aalt
    with open(path, "w") as fh:
        fh.write(" Number of Atoms and Title :\n")
        fh.write(f"{n:6d}  {title}\n")
        fh.write(" Periodic Box Dimensions :\n")
        b = list(box6) if box6 is not None else [0.0] * 6
        fh.write("".join(_fortran_d(v) for v in b[:3]) + "\n")
        fh.write("".join(_fortran_d(v) for v in b[3:]) + "\n")
        for label, arr in ((" Current Atomic Positions :", xyz), (" Current Atomic Velocities :", vel),
                           (" Current Atomic Accelerations :", acc), (" Alternate Atomic Accelerations :", aalt)):
            fh.write(label + "\n")
            for row in np.asarray(arr, float):
                fh.write("".join(_fortran_d(v) for v in row) + "\n")


def read_dyn(path: str):
    """-> dict(n, title, box, xyz, vel, acc, aalt) of a .dyn restart file (tinker/source/readdyn.f)."""
    with open(path) as fh:
        lines = fh.read().splitlines()

    def nums(ln):
        return [float(t.replace("D", "E").replace("d", "e")) for t in ln.split()]
    head = lines[1].split(None, 1)
    n = int(head[0])
    out = dict(n=n, title=head[1].strip() if len(head) > 1 else "", box=None)
    k = 2
    if lines[k].strip().startswith("Periodic Box"):
        out["box"] = nums(lines[k + 1]) + nums(lines[k + 2])
        k += 3
    for name in ("xyz", "vel", "acc", "aalt"):
        if k >= len(lines):
            break
        k += 1          # section label
        out[name] = np.array([nums(lines[k + i]) for i in range(n)])
        k += n
    return out


@dataclass
class KeyFile:
    """Keyword lines of a .key (and of the .prm it names), upper-cased keys."""
    lines: list = field(default_factory=list)   # [(KEYWORD, rest-of-line, raw)]
    directory: str = "."

    def has(self, kw: str) -> bool:
        kw = kw.upper()
        return any(k == kw for k, _, _ in self.lines)

    def get(self, kw: str, default=None):
        """Last occurrence wins (Tinker scans all lines, later overrides)."""
        kw = kw.upper()
        val = default
        for k, rest, _ in self.lines:
            if k == kw:
                val = rest
        return val

    def get_float(self, kw, default):
        v = self.get(kw)
        if v is None or not v.split():
            return default
        return float(v.split()[0].replace("D", "E").replace("d", "e"))

    def get_int(self, kw, default):
        v = self.get(kw)
        if v is None or not v.split():
            return default
        return int(v.split()[0])


def _tokenize_keyword_lines(text_lines):
    out = []
    for raw in text_lines:
        s = raw.strip()
        if not s or s[0] in "#!":
            continue
        parts = s.split(None, 1)
        out.append((parts[0].upper(), parts[1] if len(parts) > 1 else "", raw))
    return out


def read_key(path: str | None, text: str | None = None) -> KeyFile:
    """Parse a keyfile from a path or from literal text (tests pass text)."""
    if text is None:
        with open(path) as fh:
            text = fh.read()
    d = os.path.dirname(os.path.abspath(path)) if path else "."
    return KeyFile(_tokenize_keyword_lines(text.splitlines()), d)


def find_prm(key: KeyFile, search_dirs=()) -> str:
    """Resolve the `parameters` keyword the way getprm.f does (adds .prm)."""
    name = key.get("PARAMETERS")
    if name is None:
        raise FileNotFoundError("keyfile has no PARAMETERS keyword")
    name = name.split()[0]
    cands = []
    for d in (key.directory, *search_dirs):
        for ext in ("", ".prm"):
            cands.append(os.path.normpath(os.path.join(d, name + ext)))
            cands.append(os.path.normpath(os.path.join(d, os.path.basename(name) + ext)))
    for c in cands:
        if os.path.isfile(c):
            return c
    raise FileNotFoundError(f"parameter file {name!r} not found in {cands}")


@dataclass
class MultipoleRecord:
    """One `multipole` parameter (readprm.f:1215-1283)."""
    typ: int
    kz: int
    kx: int
    ky: int
    axis: str                # 'None','Z-Only','Z-then-X','Bisector','Z-Bisect','3-Fold'
    pole: np.ndarray         # 13 values, Tinker order c,dx,dy,dz,qxx,qxy,qxz,qyx,qyy,qyz,qzx,qzy,qzz


@dataclass
class PolarizeRecord:
    """One `polarize` parameter (readprm.f:1297-1335)."""
    typ: int
    alpha: float
    thole: float
    dthole: float
    group: list


@dataclass
class ForceField:
    name: str = ""
    atom_class: dict = field(default_factory=dict)     # type -> class
    atom_name: dict = field(default_factory=dict)
    atom_mass: dict = field(default_factory=dict)
    atom_atomic: dict = field(default_factory=dict)    # type -> atomic number (atomid::atomic)
    multipoles: list = field(default_factory=list)      # in file order (order matters in kmpole.f)
    polarize: dict = field(default_factory=dict)
    polpair: list = field(default_factory=list)         # (ia, ib, thole, dthole)
    scales: dict = field(default_factory=dict)          # e.g. 'MPOLE-12-SCALE' -> float
    keywords: KeyFile | None = None


SCALE_DEFAULTS = {
    # tinker/source/initprm.f defaults for AMOEBA-family potentials
    "MPOLE-12-SCALE": 0.0, "MPOLE-13-SCALE": 0.0, "MPOLE-14-SCALE": 1.0, "MPOLE-15-SCALE": 1.0,
    "POLAR-12-SCALE": 0.0, "POLAR-13-SCALE": 0.0, "POLAR-14-SCALE": 1.0, "POLAR-15-SCALE": 1.0,
    "POLAR-12-INTRA": 0.0, "POLAR-13-INTRA": 0.0, "POLAR-14-INTRA": 0.5, "POLAR-15-INTRA": 1.0,
    "DIRECT-11-SCALE": 0.0, "DIRECT-12-SCALE": 1.0, "DIRECT-13-SCALE": 1.0, "DIRECT-14-SCALE": 1.0,
    "MUTUAL-11-SCALE": 1.0, "MUTUAL-12-SCALE": 1.0, "MUTUAL-13-SCALE": 1.0, "MUTUAL-14-SCALE": 1.0,
}


def _axis_type(kz: int, kx: int, ky: int) -> str:
    # readprm.f:1262-1266 / kmpole.f:166-170
    axt = "Z-then-X"
    if kz == 0:
        axt = "None"
    if kz != 0 and kx == 0:
        axt = "Z-Only"
    if kz < 0 or kx < 0:
        axt = "Bisector"
    if kx < 0 and ky < 0:
        axt = "Z-Bisect"
    if max(kz, kx, ky) < 0:
        axt = "3-Fold"
    return axt


def _f(tok):
    return float(tok.replace("D", "E").replace("d", "e"))


def parse_multipole(head_rest: str, next4: list) -> MultipoleRecord:
    toks = head_rest.split()
    ints = [int(t) for t in toks[:-1]]
    c = _f(toks[-1])
    ints += [0] * (4 - len(ints))
    typ, kz, kx, ky = ints[:4]
    axt = _axis_type(kz, kx, ky)
    d = [_f(t) for t in next4[0].split()[:3]]
    qxx = _f(next4[1].split()[0])
    qyx, qyy = [_f(t) for t in next4[2].split()[:2]]
    qzx, qzy, qzz = [_f(t) for t in next4[3].split()[:3]]
    pole = np.array([c, d[0], d[1], d[2], qxx, qyx, qzx, qyx, qyy, qzy, qzx, qzy, qzz])
    return MultipoleRecord(typ, abs(kz), abs(kx), abs(ky), axt, pole)


def parse_polarize(rest: str) -> PolarizeRecord:
    """`polarize type alpha [thole [dthole]] group-types...` -- the Thole and
    direct-Thole fields are present iff they parse as non-integers
    (readprm.f:1305-1327 tests `getnumb` == 0)."""
    toks = rest.split()
    typ = int(toks[0])
    alpha = _f(toks[1])
    rest_t = toks[2:]
    thl = thd = 0.0

    def is_int(t):
        try:
            int(t)
            return True
        except ValueError:
            return False

    if rest_t and not is_int(rest_t[0]):
        thl = _f(rest_t[0])
        rest_t = rest_t[1:]
        if rest_t and not is_int(rest_t[0]):
            thd = _f(rest_t[0])
            rest_t = rest_t[1:]
    grp = [int(t) for t in rest_t if is_int(t) and int(t) != 0]
    return PolarizeRecord(typ, alpha, thl, thd, grp)


def read_prm(path: str) -> ForceField:
    with open(path) as fh:
        raw = fh.read().splitlines()
    ff = ForceField()
    ff.scales = dict(SCALE_DEFAULTS)
    kept = []
    i = 0
    while i < len(raw):
        s = raw[i].strip()
        i += 1
        if not s or s[0] in "#!":
            continue
        parts = s.split(None, 1)
        kw = parts[0].upper()
        rest = parts[1] if len(parts) > 1 else ""
        if kw == "FORCEFIELD":
            ff.name = rest.strip()
        elif kw == "ATOM":
            # atom  type class name "description" atomic mass valence
            t = rest.split('"')
            a = t[0].split()
            typ, cls = int(a[0]), int(a[1])
            ff.atom_class[typ] = cls
            ff.atom_name[typ] = a[2] if len(a) > 2 else ""
            if len(t) >= 3:
                tail = t[2].split()
                if len(tail) >= 2:
                    ff.atom_mass[typ] = _f(tail[1])
                if tail and tail[0].lstrip("-").isdigit():
                    ff.atom_atomic[typ] = int(tail[0])
        elif kw == "MULTIPOLE":
            ff.multipoles.append(parse_multipole(rest, raw[i:i + 4]))
            i += 4
        elif kw == "POLARIZE":
            rec = parse_polarize(rest)
            ff.polarize[rec.typ] = rec
        elif kw == "POLPAIR":
            t = rest.split()
            ff.polpair.append((int(t[0]), int(t[1]), _f(t[2]), _f(t[3]) if len(t) > 3 else 0.0))
        elif kw in SCALE_DEFAULTS:
            ff.scales[kw] = _f(rest.split()[0])
        kept.append((kw, rest, s))
    ff.keywords = KeyFile(kept, os.path.dirname(os.path.abspath(path)))
    return ff
