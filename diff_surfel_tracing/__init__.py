"""Drop-in for the reference package `diff_surfel_tracing`
(submodules/diff-surfel-tracing/diff_surfel_tracing/__init__.py), imported by easyvolcap/utils/optix_utils.py:7.
OptiX is replaced by envgs_b200's LBVH tracer (sm_100a)."""
from envgs_b200.tracer import SurfelTracer, SurfelTracingSettings, cpu_deep_copy_tuple, trace_surfels, _TraceSurfels

__all__ = ["SurfelTracer", "SurfelTracingSettings", "trace_surfels"]
