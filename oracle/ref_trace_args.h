// ORACLE tooling (checker only): flat argument block shared by the two host drivers around the UNMODIFIED reference tracer
// device programs — oracle/ref_trace_driver.cu (OptiX traversal emulated) and oracle/ref_trace_optix_driver.cpp (real OptiX).
// Mirrored field by field by oracle/ref_trace.py (_Args).  Null pointers stand for the reference's empty tensors.
#pragma once
struct RefTraceArgs
{
    int P, H, W, D, M;
    int training, start_from_first, max_trace_depth;
    float specular_threshold, scale_modifier;
    int num_triangles;
    const float* vertices; const int* triangles;
    float *ray_o, *ray_d, *background, *means3D, *shs, *colors_precomp, *others_precomp, *opacities, *scales, *rotations;
    float *viewmatrix, *projmatrix, *campos;
    float *out_rgb, *out_dpt, *out_acc, *out_dist, *out_norm, *out_aux, *mid_val, *a_weights;
    float *dL_drgb, *dL_ddpt, *dL_dacc, *dL_ddist, *dL_dnorm, *dL_daux;
    float *dL_dray_o, *dL_dray_d, *dL_dmeans3D, *dL_dshs, *dL_dcolors, *dL_dothers, *dL_dopacities, *dL_dscales, *dL_drotations,
          *dL_dtransMat;
};
