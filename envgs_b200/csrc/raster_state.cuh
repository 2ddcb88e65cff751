// raster_state.cuh — HBM layout of the rasterizer's three opaque scratch regions.
//
// The reference carves the same three regions (GeometryState / BinningState / ImageState,
// rasterizer_impl.cu:155-194) but stores the per-surfel state as separate AoS arrays (means2D, transMat(9),
// normal_opacity, depths ...).  Here every surfel's render-time state is ONE 64-byte record of four float4
// so the tile kernels fetch it with four 128-bit loads:
//
//   rec[4*i+0] = (Tu.x, Tu.y, Tu.z, Tv.x)
//   rec[4*i+1] = (Tv.y, Tv.z, Tw.x, Tw.y)
//   rec[4*i+2] = (Tw.z, c.x,  c.y,  opacity)       c = projected centre (means2D)
//   rec[4*i+3] = (n.x,  n.y,  n.z,  view depth)    n = view-space normal, flipped to face the camera
//
// plus an 8-byte inclusive pixel bounding box (short4) of the region where the surfel can reach
// alpha >= 1/255 (used for exact, parity-safe culling; see DESIGN.md "culling").
#pragma once
#include "common.cuh"

namespace envgs {

struct GeomState {
    float4* rec;            // 4*P
    short4* pixbox;         // P    inclusive pixel bbox (x0,y0,x1,y1) of the contributing region
    ushort4* tilebox;       // P    tile range [x0,x1) x [y0,y1) actually instanced
    float* depths;          // P
    uint32_t* tiles_touched;  // P
    uint32_t* point_offsets;  // P   inclusive scan of tiles_touched
    float* rgb;             // 3*P  SH -> RGB result (C == 3, shs given)
    uint8_t* clamped;       // 3*P
    int32_t* internal_radii;  // P (used when caller passes radii == NULL)
    char* scan_temp;
    size_t scan_temp_bytes;

    static size_t carve(char* base, int P, GeomState* s, size_t scan_bytes);
};

struct ImageState {
    float* final_T;       // 3*HW : T, M1, M2
    uint32_t* n_contrib;  // 2*HW : last contributor (1-based list position), median contributor
    uint2* ranges;        // tiles
    static size_t carve(char* base, int W, int H, ImageState* s);
};

struct BinningState {
    uint64_t* keys_unsorted;  // N
    uint64_t* keys;           // N
    uint32_t* vals_unsorted;  // N
    uint32_t* point_list;     // N
    // Per-instance render state in SORTED (tile, depth) order, structure-of-arrays: a tile's list is a contiguous index range
    // of every array, so the tile kernels stage a 256-instance batch with a handful of bulk asynchronous copies
    // (cp.async.bulk / UBLKCP, tma.cuh) instead of 256 scattered record gathers.  `cap` = N rounded up (+ slack so that a
    // copy starting at an index rounded down to a multiple of 4 and ending rounded up never leaves the arrays).
    float4* srec[4];          // cap each: the four float4 of the surfel record (raster_state.cuh header)
    short4* sbox;             // cap
    int32_t* sid;             // cap   surfel id (destination row of the gradient / weight reductions)
    float* sfeat;             // C * cap, channel-major
    size_t cap;
    char* sort_temp;
    size_t sort_temp_bytes;
    static size_t carve(char* base, int N, BinningState* s, size_t sort_bytes, int C);
};

// shared-memory image of one staged batch (two of them are double-buffered); SB = 256 + up to 3 leading + 1..3 trailing
constexpr int STAGE_SLOTS = 260;
template <int C>
struct TileStage {
    float4 r0[STAGE_SLOTS], r1[STAGE_SLOTS], r2[STAGE_SLOTS], r3[STAGE_SLOTS];
    short4 box[STAGE_SLOTS];
    int32_t id[STAGE_SLOTS];
    float feat[C][STAGE_SLOTS];
};

template <typename T>
static inline void obtain(char*& p, T*& ptr, size_t count, size_t alignment = 256) {
    size_t off = (reinterpret_cast<uintptr_t>(p) + alignment - 1) & ~(alignment - 1);
    ptr = reinterpret_cast<T*>(off);
    p = reinterpret_cast<char*>(ptr + count);
}

size_t scan_temp_bytes(int P);
size_t sort_temp_bytes(int N);

}  // namespace envgs
