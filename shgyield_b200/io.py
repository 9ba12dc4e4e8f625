"""
Input/output of the SSHG yield path: susceptibility files, the YAML input of the command line,
and the output table.

 * `loadeps` / `loadshg` follow the reference's loaders (main.py:14-22): chi1 files have three
   columns (energy, Re, Im) and become eps = 1 + 4 pi chi1 * scale; chi2 files have five columns
   (energy, Re 1w, Im 1w, Re 2w, Im 2w) and become scale * ((Re1w + Re2w) + i (Im1w + Im2w)).
 * The YAML schema is RECONSTRUCTED: the reference's README (README.md:49-50) documents
   `python shgyield.py <input.yml>` but neither the script nor an input file is part of the
   checkout; the only surviving fragment is in example/sshg-tutorial.ipynb (cells 4 and 12).  The
   schema below is a superset of that fragment; see `load_input`.
 * `save_spectrum` writes the table of the reference's batch driver (main.py:139-142): columns
   energy, RpP, RsP, RpS, RsS.
"""
from __future__ import annotations

import os

import numpy as np

CHI2_KEYS = ("xxx", "xyy", "xzz", "xyz", "xxz", "xxy",
             "yxx", "yyy", "yzz", "yyz", "yxz", "yxy",
             "zxx", "zyy", "zzz", "zyz", "zxz", "zxy")
CHI2_SCALE = 1e6 * 1e-24          # pm^2/V-type units of the shipped files -> m^2/V (main.py:48)


def loadeps(infile, scale):
    """chi1 file (energy, Re, Im) -> {'energy', 'data': eps = 1 + 4 pi chi1 scale} (main.py:14-17)."""
    energy, real, imag = np.loadtxt(infile, unpack=True)
    return {"energy": energy, "data": 1 + (4 * np.pi * ((real + 1j * imag) * scale))}


def loadshg(infile, scale):
    """chi2 file (energy, Re1w, Im1w, Re2w, Im2w) -> {'energy', 'data'} (main.py:19-22)."""
    energy, re1w, im1w, re2w, im2w = np.loadtxt(infile, unpack=True)
    return {"energy": energy, "data": scale * ((re1w + re2w) + 1j * (im1w + im2w))}


def loadeps_merged(infile, scale):
    """Legacy merged chi1 file `*-chi1-xx_yy_zz` (energy, xx Re, xx Im, yy Re, yy Im, zz Re, zz Im)
    mentioned by the tutorial notebook -> {'xx': ..., 'yy': ..., 'zz': ...}."""
    cols = np.loadtxt(infile, unpack=True)
    if cols.shape[0] != 7:
        raise ValueError("%s: expected 7 columns (energy + xx, yy, zz re/im), found %d" % (infile, cols.shape[0]))
    return {k: {"energy": cols[0], "data": 1 + (4 * np.pi * ((cols[1 + 2 * i] + 1j * cols[2 + 2 * i]) * scale))}
            for i, k in enumerate(("xx", "yy", "zz"))}


def _load_medium(spec, scale, base):
    """chi1 entry of the YAML: a number, a path (3- or 7-column file) or {xx|yy|zz: path | number}."""
    if isinstance(spec, (int, float)):
        return {"xx": float(spec), "yy": float(spec), "zz": float(spec)}
    if isinstance(spec, str):
        path = os.path.join(base, spec)
        ncol = np.loadtxt(path, max_rows=2).shape[-1]
        if ncol == 7:
            return loadeps_merged(path, scale)
        return {"xx": loadeps(path, scale)}
    if isinstance(spec, dict):
        out = {}
        for k, v in spec.items():
            out[k] = float(v) if isinstance(v, (int, float)) else loadeps(os.path.join(base, v), scale)
        return out
    raise ValueError("chi1 entry must be a number, a path or a mapping of components")


def resolve_chi2(spec: dict, base: str, scale: float = CHI2_SCALE) -> dict:
    """chi2 section: `abc: path | 0 | number | abc | -abc`; unlisted components are 0
    (alias resolution of example/sshg-tutorial.ipynb cell 12)."""
    chi2, alias = {}, {}
    for key in CHI2_KEYS:
        if key not in spec:
            chi2[key] = 0
            continue
        val = spec[key]
        if isinstance(val, (int, float)):
            chi2[key] = val if val == 0 else complex(val)
        elif isinstance(val, str) and val.lstrip("-") in CHI2_KEYS:
            alias[key] = val
        elif isinstance(val, str):
            chi2[key] = loadshg(os.path.join(base, val), scale)
        else:
            raise ValueError("chi2.%s: expected a path, 0, a number or another component" % key)
    unknown = set(spec) - set(CHI2_KEYS) - {"scale"}
    if unknown:
        raise ValueError("chi2: unknown components %s" % sorted(unknown))
    for key, val in alias.items():
        src = val.lstrip("-")
        if src not in chi2:
            raise ValueError("chi2.%s refers to %s, which is itself an alias" % (key, src))
        ref = chi2[src]
        if val.startswith("-"):
            ref = {"energy": ref["energy"], "data": -1 * ref["data"]} if isinstance(ref, dict) else -1 * ref
        chi2[key] = ref
    return chi2


def _energy_axis(spec):
    if spec is None:
        return np.linspace(0.01, 10.0, 1000)             # the tutorial's ONEE (notebook cell 8)
    if isinstance(spec, dict):
        lo, hi, step = float(spec["min"]), float(spec["max"]), float(spec["step"])
        n = int(round((hi - lo) / step)) + 1
        return np.linspace(lo, hi, n)
    return np.atleast_1d(np.asarray(spec, dtype=np.float64))


def load_input(path: str) -> dict:
    """Parses the YAML input into the keyword arguments of shg.shgyield plus 'output'.

        parameters: {theta, phi, gamma: 0, energy: {min, max, step} | [..], mode: 3-layer, mref: true,
                     output: spectrum.dat}
        multiref:   {thickness: nm}
        broadening: {eps: 0, chi: 0, out: 5, strict: false}   # sigmas in samples, as in shgyield()
        chi1:       {chib: path | {xx,yy,zz}, chil: path | {xx,yy,zz}, norm: 1.0, vacuum: 1.0}
        chi2:       {scale: 1e-18, abc: path | 0 | abc | -abc, ...}

    Paths are relative to the YAML file.  theta / phi may be lists (a sweep)."""
    import yaml
    with open(path) as f:
        doc = yaml.safe_load(f)
    if not isinstance(doc, dict):
        raise ValueError("%s: not a YAML mapping" % path)
    base = os.path.dirname(os.path.abspath(path))
    par = doc.get("parameters") or {}
    for req in ("theta", "phi"):
        if req not in par:
            raise ValueError("parameters.%s is required" % req)
    chi1 = doc.get("chi1") or {}
    if "chib" not in chi1 or "chil" not in chi1:
        raise ValueError("chi1.chib and chi1.chil are required")
    chi2_spec = dict(doc.get("chi2") or {})
    scale = float(chi2_spec.pop("scale", CHI2_SCALE))
    brd = doc.get("broadening") or {}
    return {
        "energy": _energy_axis(par.get("energy")),
        "eps_m1": _load_medium(chi1.get("vacuum", 1.0), 1.0, base),
        "eps_m2": _load_medium(chi1["chil"], float(chi1.get("norm", 1.0)), base),
        "eps_m3": _load_medium(chi1["chib"], 1.0, base),
        "chi2": resolve_chi2(chi2_spec, base, scale),
        "theta": par["theta"], "phi": par["phi"],
        "thick": (doc.get("multiref") or {}).get("thickness", 0.0),
        "gamma": par.get("gamma", 0),
        "sigma_eps": float(brd.get("eps", 0.0)), "sigma_chi": float(brd.get("chi", 0.0)),
        "sigma_out": float(brd.get("out", 5.0)),
        "strict": bool(brd.get("strict", False)),          # sweeps only: broaden every row itself (see shgyield_grid)
        "mode": par.get("mode", "3-layer"), "mref": bool(par.get("mref", True)),
        "output": par.get("output", "spectrum.dat"),
    }


def save_spectrum(path, result):
    """energy, RpP, RsP, RpS, RsS -- the column order and format of the reference (main.py:139-142)."""
    np.savetxt(path, np.column_stack((result["energy"], result["pp"], result["sp"], result["ps"], result["ss"])),
               fmt="%07.4f  %.8e  %.8e  %.8e  %.8e",
               header="w(eV)  RpP             RsP             RpS             RsS")
