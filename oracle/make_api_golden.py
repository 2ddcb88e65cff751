"""The Python operator surface of the reference packages, extracted from their sources with `ast` (nothing is executed or imported:
the packages need their compiled `_C` modules) -> tests/golden/api_surface.json.  Test infrastructure only.

For each of `diff_surfel_rasterization_wet`, `_wet_ch05`, `_wet_ch07` and `diff_surfel_tracing`: the NamedTuple field lists, every
function / method with its argument names and defaults, and the messages of the exceptions the wrappers raise themselves.
tests/test_capi.py compares the drop-in packages of this repository against it (same names, argument order,
defaults, error behaviour).

    python oracle/make_api_golden.py
"""
from __future__ import annotations

import ast
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAST = "/root/reference/submodules/diff-surfel-rasterizations"
SOURCES = {
    "diff_surfel_rasterization_wet": f"{RAST}/diff-surfel-rasterization-wet/diff_surfel_rasterization_wet/__init__.py",
    "diff_surfel_rasterization_wet_ch05": f"{RAST}/diff-surfel-rasterization-wet-ch05/diff_surfel_rasterization_wet_ch05/__init__.py",
    "diff_surfel_rasterization_wet_ch07": f"{RAST}/diff-surfel-rasterization-wet-ch07/diff_surfel_rasterization_wet_ch07/__init__.py",
    "diff_surfel_tracing": "/root/reference/submodules/diff-surfel-tracing/diff_surfel_tracing/__init__.py",
}


def _const(node):
    try:
        return ast.literal_eval(node)
    except Exception:
        return ast.unparse(node)


def _func(fn: ast.FunctionDef) -> dict:
    a = fn.args
    names = [x.arg for x in a.posonlyargs + a.args]
    defaults = [_const(d) for d in a.defaults]
    return {"args": names, "defaults": dict(zip(names[len(names) - len(defaults):], defaults)),
            "staticmethod": any(isinstance(d, ast.Name) and d.id == "staticmethod" for d in fn.decorator_list)}


def _raises(tree) -> list:
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Raise) and isinstance(node.exc, ast.Call) and node.exc.args:
            arg = node.exc.args[0]
            if isinstance(arg, ast.Constant) and isinstance(arg.value, str):
                out.append([ast.unparse(node.exc.func), arg.value])
    return sorted(out)


def surface(path: str) -> dict:
    tree = ast.parse(open(path).read())
    out = {"functions": {}, "classes": {}, "raises": _raises(tree)}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            out["functions"][node.name] = _func(node)
        elif isinstance(node, ast.ClassDef):
            c = {"bases": [ast.unparse(b) for b in node.bases], "fields": [], "methods": {}}
            for item in node.body:
                if isinstance(item, ast.AnnAssign) and isinstance(item.target, ast.Name):
                    c["fields"].append(item.target.id)
                elif isinstance(item, ast.FunctionDef):
                    c["methods"][item.name] = _func(item)
            out["classes"][node.name] = c
    return out


HOST_SOURCES = {   # the C++ entry points' own argument checks (AT_ERROR -> RuntimeError in Python)
    "rasterizer": f"{RAST}/diff-surfel-rasterization-wet-ch05/rasterize_points.cu",
    "tracer": "/root/reference/submodules/diff-surfel-tracing/trace_surfels.cpp",
}


def host_errors(path: str) -> list:
    import re
    return sorted(set(re.findall(r'AT_ERROR\("([^"]+)"\)', open(path).read())))


CONST_SOURCES = [f"{RAST}/diff-surfel-rasterization-wet-ch05/cuda_rasterizer/config.h",
                 f"{RAST}/diff-surfel-rasterization-wet-ch05/cuda_rasterizer/auxiliary.h",
                 "/root/reference/submodules/diff-surfel-tracing/optix_tracer/config.h"]


def constants() -> dict:
    """Numeric `#define NAME value` and `__device__ const float NAME = value;` of the reference's configuration headers (the
    rasterizer's NUM_CHANNELS is per variant and left out; the tracer's is kept as TRACER_NUM_CHANNELS)."""
    import re
    out = {}
    for path in CONST_SOURCES:
        text = open(path).read()
        pairs = re.findall(r"^#define\s+(\w+)\s+([-+0-9.eE]+)f?\s*(?://.*)?$", text, flags=re.M)
        pairs += re.findall(r"^__device__ const float (\w+) = ([-+0-9.eE]+)f?;", text, flags=re.M)
        for name, val in pairs:
            if name == "NUM_CHANNELS":
                if "optix_tracer" not in path:
                    continue
                name = "TRACER_NUM_CHANNELS"
            out[name] = float(val)
    return out


if __name__ == "__main__":
    res = {name: surface(path) for name, path in SOURCES.items()}
    res["_constants"] = constants()
    res["_host_errors"] = {k: host_errors(v) for k, v in HOST_SOURCES.items()}
    with open(os.path.join(ROOT, "tests", "golden", "api_surface.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print(res["_host_errors"])
    for name, s in res.items():
        if name.startswith("_"):
            continue
        print(name, "functions", list(s["functions"]), "classes", {k: list(v["methods"]) for k, v in s["classes"].items()}, "raises", len(s["raises"]))
