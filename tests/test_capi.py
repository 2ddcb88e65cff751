"""CPU: the C-ABI library builds for sm_100a without a GPU, loads, exports every symbol include/envgs_b200.h declares,
and reports errors through return codes + envgs_last_error (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "envgs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(envgs_[a-z_0-9]+)\s*\(", text)) - {"envgs_alloc_fn"})


def test_header_symbols_are_exported(built_lib):
    from envgs_b200 import _lib
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/envgs_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert L.envgs_version() >= 100


def test_only_sm100a_code_is_embedded(built_lib):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_size_queries_and_error_reporting(built_lib):
    from envgs_b200 import _lib
    L = _lib.lib()
    assert L.envgs_raster_geom_bytes(1000) > 1000 * 64
    assert L.envgs_raster_img_bytes(1920, 1080) >= 1920 * 1080 * 20
    assert L.envgs_raster_binning_bytes(0) > 0
    assert L.envgs_tracer_bvh_bytes(262144) > 262144 * 128
    assert L.envgs_trace_cache_bytes(1080, 1920) >= 1080 * 1920 * 136
    args = _lib.RasterArgs()
    args.P, args.C, args.width, args.height = 10, 4, 64, 64          # unsupported channel count
    out = _lib.RasterOut()
    cb = _lib.ALLOC_FN(lambda u, n: None)
    rc = L.envgs_raster_forward(C.byref(args), C.byref(out), cb, cb, cb, None, None)
    assert rc == -1 and b"channel" in L.envgs_last_error()
    args.C = 5                                                          # non-RGB without precomputed colours
    rc = L.envgs_raster_forward(C.byref(args), C.byref(out), cb, cb, cb, None, None)
    assert rc == -1
    with pytest.raises(RuntimeError):
        _lib.check(rc, "envgs_raster_forward")
    assert L.envgs_trace_forward(None, None, None, None) == -1


def test_drop_in_packages_expose_reference_names(built_lib):
    import diff_surfel_rasterization_wet as w3
    import diff_surfel_rasterization_wet_ch05 as w5
    import diff_surfel_rasterization_wet_ch07 as w7
    import diff_surfel_tracing as tr
    assert (w3.NUM_CHANNELS, w5.NUM_CHANNELS, w7.NUM_CHANNELS) == (3, 5, 7)
    for m in (w3, w5, w7):
        assert m.GaussianRasterizationSettings._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg",
                                                           "scale_modifier", "viewmatrix", "projmatrix", "sh_degree", "campos",
                                                           "prefiltered", "debug")
        assert hasattr(m.GaussianRasterizer, "markVisible") and hasattr(m, "rasterize_gaussians")
    assert tr.SurfelTracingSettings._fields[-2:] == ("max_trace_depth", "specular_threshold")
    import inspect
    sig = inspect.signature(tr.SurfelTracer.forward)
    assert list(sig.parameters)[1:6] == ["ray_o", "ray_d", "vertices", "means3D", "grads3D"]
    assert sig.parameters["start_from_first"].default is True
    sig = inspect.signature(w5.GaussianRasterizer.forward)
    assert list(sig.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp"]


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "envgs_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
    for d in ("diff_surfel_rasterization_wet", "diff_surfel_rasterization_wet_ch05", "diff_surfel_rasterization_wet_ch07", "diff_surfel_tracing"):
        src = open(os.path.join(ROOT, d, "__init__.py")).read()
        assert "oracle" not in src


def _header_structs():
    """struct name -> [(kind, field name)] parsed from include/envgs_b200.h; kind is 'ptr', 'i32', 'f32' or 'u64'."""
    import re
    text = open(os.path.join(ROOT, "include", "envgs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            base, names = decl.rsplit(" ", 1)[0], decl
            # "const float* a, *b" does not occur in this header; lists are plain "int32_t P, C, sh_degree"
            typ = decl[:decl.index(decl.split(",")[0].split()[-1])].strip()
            for name in decl[len(typ):].split(","):
                name = name.strip()
                ptr = "*" in typ or name.startswith("*")
                kind = "ptr" if ptr else {"int32_t": "i32", "uint32_t": "i32", "float": "f32", "uint64_t": "u64", "size_t": "u64"}[typ.replace("const ", "")]
                fields.append((kind, name.lstrip("*")))
        out[m.group(1)] = fields
    return out


def test_ctypes_structures_mirror_the_header():
    """The Python shim passes these structs by reference: a field added to the header but not to envgs_b200/_lib.py (or the other way
    round, or in another order) would silently shift every later member."""
    import ctypes as C
    from envgs_b200 import _lib
    pairs = {"envgs_raster_args": _lib.RasterArgs, "envgs_raster_out": _lib.RasterOut, "envgs_raster_grads": _lib.RasterGrads,
             "envgs_trace_args": _lib.TraceArgs, "envgs_trace_out": _lib.TraceOut, "envgs_trace_grads": _lib.TraceGrads,
             "envgs_glue_camera": _lib.GlueCamera, "envgs_densify_params": _lib.DensifyParams, "envgs_row_array": _lib.RowArray}
    structs = _header_structs()
    kind_of = {C.c_void_p: "ptr", C.c_int32: "i32", C.c_int: "i32", C.c_uint32: "i32", C.c_float: "f32", C.c_uint64: "u64", C.c_size_t: "u64"}
    for cname, cls in pairs.items():
        assert cname in structs, cname
        want = structs[cname]
        got = [(kind_of.get(t, "ptr" if hasattr(t, "contents") or t is C.c_char_p else str(t)), n) for n, t in cls._fields_]
        assert [n for _, n in got] == [n for _, n in want], f"{cname}: field names/order differ:\n  header {want}\n  ctypes {got}"
        assert [k for k, _ in got] == [k for k, _ in want], f"{cname}: field kinds differ:\n  header {want}\n  ctypes {got}"


def _check_constants(ref: dict):
    """The reference's configuration constants (config.h / auxiliary.h of the rasterizer, config.h of the tracer) against the
    constexpr block of envgs_b200/csrc/common.cuh and the module constants of envgs_b200/tracer.py."""
    import re
    from envgs_b200 import tracer
    text = open(os.path.join(ROOT, "envgs_b200", "csrc", "common.cuh")).read()
    mine = {}
    for line in re.findall(r"^constexpr (?:int|float) ([^;]+);", text, flags=re.M):
        for name, val in re.findall(r"(\w+) = ([-+0-9.eE]+)f?(?=\s*(?:,|$))", line):
            mine[name] = float(val)
    mapping = {"BLOCK_X": "TILE_X", "BLOCK_Y": "TILE_Y", "near_n": "NEAR_N", "far_n": "FAR_N", "FilterSize": "FILTER_SIZE",
               "FilterInvSquare": "FILTER_INV_SQUARE", "DEPTH_OFFSET": "DEPTH_OFFSET", "ALPHA_OFFSET": "ALPHA_OFFSET",
               "NORMAL_OFFSET": "NORMAL_OFFSET", "MIDDEPTH_OFFSET": "MIDDEPTH_OFFSET", "DISTORTION_OFFSET": "DISTORTION_OFFSET",
               "CHUNK_SIZE": "TR_CHUNK", "MAX_INTERSECTION": "TR_MAX_INTERSECTION", "MID_CHANNELS": "TR_MID_CHANNELS",
               "MAX_TRACE_DEPTH": "TR_MAX_TRACE_DEPTH", "START_OFFSET": "TR_START_OFFSET", "STEP_EPSILON": "TR_STEP_EPSILON",
               "DEPTH_INFINTY": "TR_DEPTH_INF", "RAYO_MID_OFFSET": "MID_RAYO", "RAYD_MID_OFFSET": "MID_RAYD", "DPT_MID_OFFSET": "MID_DPT",
               "ACC_MID_OFFSET": "MID_ACC", "NORM_MID_OFFSET": "MID_NORM", "AUX_MID_OFFSET": "MID_AUX", "RGB_MID_OFFSET": "MID_RGB"}
    for r, m in mapping.items():
        assert r in ref and m in mine, (r, m)
        assert mine[m] == pytest.approx(ref[r], rel=1e-6), (r, ref[r], m, mine[m])
    assert (tracer.NUM_CHANNELS, tracer.AUX_CHANNELS, tracer.MID_CHANNELS, tracer.MAX_TRACE_DEPTH) == \
        (ref["TRACER_NUM_CHANNELS"], ref["AUX_CHANNELS"], ref["MID_CHANNELS"], ref["MAX_TRACE_DEPTH"])


def test_drop_in_packages_match_the_reference_api_surface(built_lib):
    """Names, argument order, defaults, NamedTuple fields and the wrappers' own exception messages of the four reference packages
    (extracted from their sources by oracle/make_api_golden.py -> tests/golden/api_surface.json) against the drop-in packages."""
    import importlib
    import inspect
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "api_surface.json")))
    consts = gold.pop("_constants")
    _check_constants(consts)
    for which, msgs in gold.pop("_host_errors").items():     # AT_ERROR checks of the C++ entry points: same RuntimeError messages
        src = open(os.path.join(ROOT, "envgs_b200", which + ".py")).read()
        assert msgs and all(f'RuntimeError("{m}")' in src for m in msgs), (which, msgs)
    for pkg, ref in gold.items():
        mod = importlib.import_module(pkg)
        for fname, f in ref["functions"].items():
            fn = getattr(mod, fname)
            params = list(inspect.signature(fn).parameters)
            assert params == f["args"], (pkg, fname, params, f["args"])
        for cname, c in ref["classes"].items():
            cls = getattr(mod, cname)
            if "NamedTuple" in c["bases"]:
                assert list(cls._fields) == c["fields"], (pkg, cname)
                continue
            for mname, m in c["methods"].items():
                meth = getattr(cls, mname)
                sig = inspect.signature(meth)
                params = list(sig.parameters)
                want = m["args"]
                if m["staticmethod"] is False and params and params[0] != "self" and want and want[0] == "self":
                    want = want[1:]                      # bound-style introspection of functions defined without `self`
                assert params == want, (pkg, cname, mname, params, want)
                for arg, default in m["defaults"].items():
                    got = sig.parameters[arg].default
                    assert got == default or (got is None and default is None), (pkg, cname, mname, arg, got, default)
        # the messages of the exceptions the reference wrappers raise themselves appear verbatim in the drop-in sources
        src = open(os.path.join(ROOT, "envgs_b200", "rasterizer.py" if "rasterization" in pkg else "tracer.py")).read()
        for exc, msg in ref["raises"]:
            assert msg in src, (pkg, exc, msg)
