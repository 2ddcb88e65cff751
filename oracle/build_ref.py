"""ORACLE tooling (checker only): compile the UNMODIFIED reference rasterizer extensions for sm_100a.

Sources are compiled where they lie under /root/reference (never copied into this repository); outputs go to
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  The resulting modules are used only by
`oracle/make_golden.py` (to generate tests/golden/*.pt on a B200) and by the optional live comparison in
tests/test_gpu_reference.py — never by the product path.

Recipe: torch.utils.cpp_extension.load on the five files of each variant with
`-include cstdint` (gcc 13 no longer provides uint32_t/uintptr_t transitively to rasterizer_impl.h:24,40-61) and the
vendored GLM on the include path.

Tracer: the reference host code needs the OptiX runtime (libnvoptix.so.1), which the GPU boxes of this pool do not ship
(probe log: profiles/round2_optix_probe.txt).  Its DEVICE programs (optix_tracer/forward.cu, backward.cu) hold all of the
surfel arithmetic and use only eight OptiX intrinsics, so they are compiled UNMODIFIED as plain CUDA against the emulation
header oracle/optix_shim/optix.h with the small host driver oracle/ref_trace_driver.cu:
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC \
         -I oracle/optix_shim -I <ref>/optix_tracer -DREF_TRACE_SOURCE='"<ref>/optix_tracer/forward.cu"' -DREF_TRACE_BACKWARD=0 \
         oracle/ref_trace_driver.cu -o oracle/_ref/ref_trace/libref_trace_fwd.so      (and backward.cu / =1 / _bwd.so)
No fast-math flags: the reference's CMake build adds none (CMakeLists.txt, optix_tracer/CMakeLists.txt).
"""
from __future__ import annotations

import os
import sys

REF = "/root/reference/submodules/diff-surfel-rasterizations"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
VARIANTS = {
    "ref_rast_wet": "diff-surfel-rasterization-wet",            # NUM_CHANNELS 3, in-kernel SH
    "ref_rast_wet_ch05": "diff-surfel-rasterization-wet-ch05",  # NUM_CHANNELS 5 (what EnvGS imports)
    "ref_rast_wet_ch07": "diff-surfel-rasterization-wet-ch07",  # NUM_CHANNELS 7 (render() picks it for 7-channel features)
}


REF_TRACE = "/root/reference/submodules/diff-surfel-tracing/optix_tracer"
HERE = os.path.dirname(os.path.abspath(__file__))


def build_tracer(verbose: bool = False) -> dict:
    """Compile the unmodified reference tracer device programs against the OptiX emulation header (see module docstring)."""
    import subprocess
    if not os.path.isdir(REF_TRACE):
        raise FileNotFoundError(f"{REF_TRACE} not present (the reference only exists in the build container)")
    bdir = os.path.join(OUT, "ref_trace")
    os.makedirs(bdir, exist_ok=True)
    deps = [os.path.join(HERE, "ref_trace_driver.cu"), os.path.join(HERE, "optix_shim", "optix.h"), os.path.join(HERE, "ref_trace_args.h")]
    built = {}
    # "_nudge": the same forward program with every emulated hit distance moved by one fp32 ulp (conditioning experiment)
    for src, flag, tag in (("forward.cu", "0", "fwd"), ("backward.cu", "1", "bwd"), ("forward.cu", "0", "fwd_nudge"),
                           ("backward.cu", "1", "bwd_nudge")):
        so = os.path.join(bdir, f"libref_trace_{tag}.so")
        built[f"ref_trace_{tag}"] = so
        if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
            continue
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-w",
               "-I", os.path.join(HERE, "optix_shim"), "-I", REF_TRACE,
               f'-DREF_TRACE_SOURCE="{os.path.join(REF_TRACE, src)}"', f"-DREF_TRACE_BACKWARD={flag}", deps[0], "-o", so]
        if tag.endswith("nudge"):
            cmd.insert(1, "-DOPTIX_SHIM_T_NUDGE=1")
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return built


def build_tracer_optix(verbose: bool = False) -> dict:
    """The reference tracer's device programs under the REAL OptiX runtime: forward.ptx / backward.ptx generated from the
    unmodified optix_tracer/{forward,backward}.cu (what optix_tracer/CMakeLists.txt's CUDA_PTX_COMPILATION does) plus the
    C-ABI host driver oracle/ref_trace_optix_driver.cpp, which repeats the reference host code's OptiX calls
    (optix_wrapper.cpp:33-215, trace_surfels.cpp:47-121, :233-311) against the vendored OptiX 7.7 headers.
    Usable only where the driver's libnvoptix.so.1 can be loaded (/usr/local/nvidia/lib on this pool's GPU boxes, outside the
    default loader path — oracle/ref_trace.py preloads it); otherwise the emulation build is the pin.
    (The reference's own pybind extension — trace_surfels.cpp, ext.cpp, optix_wrapper.cpp, common.cpp — also compiles here
    with torch.utils.cpp_extension, but its OptiXStateWrapper() segfaults on this image; profiles/round2_optix_probe.txt.)"""
    import subprocess
    top = os.path.dirname(REF_TRACE)
    if not os.path.isdir(top):
        raise FileNotFoundError(top)
    bdir = os.path.join(OUT, "ref_trace_optix")
    os.makedirs(bdir, exist_ok=True)
    inc = os.path.join(top, "third_party", "optix", "include")
    so = os.path.join(bdir, "libref_trace_optix.so")
    src = os.path.join(HERE, "ref_trace_optix_driver.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-I", inc, "-I", REF_TRACE, "-I", "/usr/local/cuda/include",
                        "-I", HERE, src, "-o", so, "-L/usr/local/cuda/lib64", "-lcudart", "-ldl"], check=True)
    probe_src = os.path.join(HERE, "optix_probe", "optix_min_probe.cpp")
    probe = os.path.join(bdir, "optix_min_probe")
    if os.path.exists(probe_src) and not os.path.exists(probe):
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I", inc, "-I", "/usr/local/cuda/include", probe_src, "-o", probe,
                        "-L/usr/local/cuda/lib64", "-lcudart", "-ldl"], check=True)
    for arch in ("compute_100", "compute_80"):    # two PTX targets: OptiX JIT-compiles whichever its parser accepts
        pdir = os.path.join(bdir, "ptx_" + arch)
        os.makedirs(pdir, exist_ok=True)
        for name in ("forward", "backward"):
            ptx = os.path.join(pdir, name + ".ptx")
            if not os.path.exists(ptx):
                subprocess.run(["nvcc", "-ptx", "-O3", "-std=c++17", "-arch=" + arch, "-w", "-I", inc, "-I", REF_TRACE,
                                os.path.join(REF_TRACE, name + ".cu"), "-o", ptx], check=True)
    return {"ref_trace_optix": so}


def build(verbose: bool = False) -> dict:
    if not os.path.isdir(REF):
        raise FileNotFoundError(f"{REF} not present (the reference only exists in the build container)")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils import cpp_extension
    built = {}
    for name, sub in VARIANTS.items():
        src = os.path.join(REF, sub)
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        so = os.path.join(bdir, name + ".so")
        if os.path.exists(so):
            built[name] = so
            continue
        sources = [os.path.join(src, "cuda_rasterizer", f) for f in ("rasterizer_impl.cu", "forward.cu", "backward.cu")]
        sources += [os.path.join(src, "rasterize_points.cu"), os.path.join(src, "ext.cpp")]
        cpp_extension.load(
            name=name, sources=sources, extra_include_paths=[os.path.join(REF, "third_party", "glm"), src],
            extra_cflags=["-include", "cstdint", "-O3"],
            extra_cuda_cflags=["-include", "cstdint", "-O3", "-gencode", "arch=compute_100a,code=sm_100a"],
            build_directory=bdir, verbose=verbose, is_python_module=False)
        built[name] = so
    built.update(build_tracer(verbose))
    try:
        built.update(build_tracer_optix(verbose))
    except Exception as ex:   # optional: only a second, stronger pin and a timing baseline
        print(f"[build_ref] real-OptiX tracer build skipped: {ex}")
    return built


def load(name: str):
    """Import a previously built reference module (GPU box: prebuilt files only, nothing is compiled there)."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    so = os.path.join(OUT, name, name + ".so")
    if not os.path.exists(so):
        raise FileNotFoundError(so)
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv))
