"""Drop-in for the reference package `diff_surfel_rasterization_wet_ch05` (NUM_CHANNELS = 5,
submodules/diff-surfel-rasterizations/diff-surfel-rasterization-wet-ch05/diff_surfel_rasterization_wet_ch05/__init__.py).
EasyVolcap imports it at easyvolcap/utils/gaussian2d_utils.py:1013-1015; the kernels are envgs_b200's sm_100a ones."""
from envgs_b200.rasterizer import GaussianRasterizationSettings, cpu_deep_copy_tuple, make_module

NUM_CHANNELS = 5
rasterize_gaussians, _RasterizeGaussians, GaussianRasterizer = make_module(NUM_CHANNELS)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
