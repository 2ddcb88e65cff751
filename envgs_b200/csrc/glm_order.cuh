// glm_order.cuh — small column-major matrix products evaluated in exactly the operation order of the GLM operators the
// reference's per-surfel code goes through (g-truc/glm @ 5c46b9c0, glm/detail/type_mat{3x3,3x4,4x3,4x4}.inl):
//     Result[c][r] = m1[0][r] * m2[c][0] + m1[1][r] * m2[c][1] + ... (k ascending, left-associated)
// INCLUDING the structural zeros and ones of splat2world / ndc2pix.  A zero term cannot change a finite sum, but whether it
// is there decides which product of `a*b + c*d` the compiler keeps as a separately rounded multiply when it contracts the
// sum into FMAs — and that 1-ulp difference in the homography T is amplified to whole pixels in `radii` for edge-on
// surfels (d = 9 Tw.x^2 + 9 Tw.y^2 - Tw.z^2 close to 0).  Writing the same expression trees makes T, the projected centre
// and the radius bit-identical to the reference build (tests/test_gpu_reference.py asserts radii equality at 1 M surfels).
#pragma once

namespace envgs {

template <int COLS, int ROWS>
struct GMat {            // v[c][r], like glm::mat<COLS, ROWS>
    float v[COLS][ROWS];
};

// (K columns x R rows) * (C columns x K rows) -> (C columns x R rows)
template <int K, int R, int C>
__device__ __forceinline__ GMat<C, R> gmul(const GMat<K, R>& a, const GMat<C, K>& b) {
    GMat<C, R> o;
#pragma unroll
    for (int c = 0; c < C; c++)
#pragma unroll
        for (int r = 0; r < R; r++) {
            float acc = a.v[0][r] * b.v[c][0];
#pragma unroll
            for (int k = 1; k < K; k++) acc = acc + a.v[k][r] * b.v[c][k];
            o.v[c][r] = acc;
        }
    return o;
}

template <int COLS, int ROWS>
__device__ __forceinline__ GMat<ROWS, COLS> gtranspose(const GMat<COLS, ROWS>& a) {
    GMat<ROWS, COLS> o;
#pragma unroll
    for (int c = 0; c < COLS; c++)
#pragma unroll
        for (int r = 0; r < ROWS; r++) o.v[r][c] = a.v[c][r];
    return o;
}

}  // namespace envgs
