// lbvh.cuh — linear BVH over environment surfel quads (replaces the OptiX GAS of TRACE/trace_surfels.cpp:47-170).
//
// One leaf per surfel (the reference instances two triangles per surfel).  Layout in HBM:
//   BvhNode  nodes[max(P-1,1)]   64 B: both children's AABBs + child links  -> one node visit = 4 x LDG.128
//   QuadPrim prims[P]            64 B: corner v1, edges e_u, e_v, plane normal -> one leaf test = 4 x LDG.128
// Child link >= 0: internal node index; < 0: ~surfel index (leaf).
#pragma once
#include "common.cuh"

namespace envgs {

struct __align__(16) BvhNode {
    float4 a;   // lo0.xyz, hi0.x
    float4 b;   // hi0.yz, lo1.xy
    float4 c;   // lo1.z, hi1.xyz
    int child0, child1, pad0, pad1;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

struct __align__(16) QuadPrim {
    float4 v1;   // corner (-,-) xyz ; w = 1/|e_u|^2
    float4 eu;   // (+,-) - (-,-)    ; w = 1/|e_v|^2
    float4 ev;   // (-,+) - (-,-)    ; w = valid flag (1) / 0 for degenerate or non-finite quads
    float4 n;    // e_u x e_v        ; w unused
};
static_assert(sizeof(QuadPrim) == 64, "QuadPrim must be 64 bytes");

// 4-wide node for the fast traversal kernel: slot j (32 B) = {lo.xyz, hi.x | hi.yz, child link, pad}, so the four lanes
// that cooperate on one ray read one 128-byte line with two LDG.128 each (slot j at byte offset 32*j).
// Child link: >= 0 internal node (same index space as BvhNode), < 0 ~surfel; empty slots have an inverted box.
struct __align__(16) Bvh4Slot {
    float4 a;   // lo.xyz, hi.x
    float4 b;   // hi.y, hi.z, link (int bits), 0
};
struct __align__(128) Bvh4Node { Bvh4Slot s[4]; };
static_assert(sizeof(Bvh4Node) == 128, "Bvh4Node must be 128 bytes");

struct Lbvh {
    int P = 0;              // surfels in the current tree
    int capacity = 0;       // surfels the storage can hold
    char* storage = nullptr;
    size_t storage_bytes = 0;
    bool owns_storage = false;
    // carved views
    BvhNode* nodes = nullptr;
    Bvh4Node* nodes4 = nullptr;     // collapsed 4-wide view (one per binary node; only even-depth ones are reachable)
    QuadPrim* prims = nullptr;
    float* leaf_lo = nullptr;      // 3*P (xyz interleaved as float4? no: 4 floats per leaf -> lo.xyz,_ )
    float* leaf_hi = nullptr;
    uint64_t* keys = nullptr;      // P
    uint64_t* keys_sorted = nullptr;
    uint32_t* vals = nullptr;
    uint32_t* vals_sorted = nullptr;   // sorted position -> surfel id
    int* parent = nullptr;         // 2P-1 : parent of internal node n (index n) / of sorted leaf k (index P-1+k), high bit = which child
    int* flags = nullptr;          // P-1 arrival counters
    float* bounds = nullptr;       // 6 floats: scene lo/hi (as ordered ints during reduction)
    char* sort_temp = nullptr;
    size_t sort_temp_bytes = 0;
    // PLOC builder state (lbvh.cu): two cluster arrays (ping-pong), nearest-neighbour indices, validity flags + their scan
    float4* cl_lo[2] = {nullptr, nullptr};
    float4* cl_hi[2] = {nullptr, nullptr};
    int* cl_link[2] = {nullptr, nullptr};    // child link of the cluster's root (>= 0 node id, < 0 ~surfel id)
    int* cl_slot[2] = {nullptr, nullptr};    // index into parent[] of the cluster's root
    int* cl_nn = nullptr;
    int* cl_valid = nullptr;
    int* cl_scan = nullptr;
    int* cl_counter = nullptr;               // [0] merges done so far, [1] and [2] clusters alive (ping-pong between iterations)
    char* scan_temp = nullptr;
    size_t scan_temp_bytes = 0;
    int full_builds = 0, calls_since_full = 1 << 30;   // topology-reuse policy (lbvh_build)
};

size_t lbvh_storage_bytes(int P);
void lbvh_carve(Lbvh& t, char* base, int P);
int lbvh_build(Lbvh& t, const float* vertices, int P, bool rebuild, cudaStream_t stream);

}  // namespace envgs
