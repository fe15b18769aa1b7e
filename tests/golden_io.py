"""Loader for tests/golden/cvmatrix_golden.npz (written by tests/golden/make_golden.py)."""

import json
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def manifest():
    if "m" not in _cache:
        with open(os.path.join(_DIR, "manifest.json")) as f:
            _cache["m"] = json.load(f)
    return _cache["m"]


def _npz():
    if "z" not in _cache:
        z = np.load(os.path.join(_DIR, "cvmatrix_golden.npz"))
        _cache["z"] = z
        _cache["layout"] = json.loads(str(z["__layout__"]))
    return _cache["z"], _cache["layout"]


def group(name):
    """All arrays / strings stored under one group name -> {relative key: value}."""
    if ("g", name) in _cache:
        return _cache[("g", name)]
    z, layout = _npz()
    out = {}
    flat = z[name] if name in z.files else np.zeros(0)
    off = 0
    for key, dtype, shape in layout[name]:
        if dtype == "str":
            out[key] = shape
            continue
        n = int(np.prod(shape)) if len(shape) else 1
        out[key] = flat[off:off + n].astype(dtype).reshape(shape)
        off += n
    _cache[("g", name)] = out
    return out


def case(name):
    """(spec, inputs, fit, out) of one case; inputs = dict(X, Y|None, w|None, vals=[...])."""
    spec = next(c for c in manifest()["cases"] if c["name"] == name)
    gin = group(spec["inputs"])
    g = group(name)
    inputs = dict(
        X=gin["in/X"],
        Y=gin.get("in/Y") if spec["has_Y"] else None,
        w=gin.get("in/w") if spec["weighted"] else None,
        vals=[gin[f"in/val{i}"].astype(np.int64) for i in range(spec["n_val_sets"])],
    )
    fit = {k[4:]: v for k, v in g.items() if k.startswith("fit/")}
    out = {k[4:]: v for k, v in g.items() if k.startswith("out/")}
    return spec, inputs, fit, out


def case_names(prefix=""):
    return [c["name"] for c in manifest()["cases"] if c["name"].startswith(prefix)]
