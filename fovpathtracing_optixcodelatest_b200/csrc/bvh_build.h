// bvh_build.h -- host entry of the on-device BVH builder (bvh_build.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fovpt {

constexpr uint32_t kMaxWideDepth = 31;      // 3 deferred children per level must fit the traversal stack (bvh.cuh kStackSize = 96)

struct BuiltBvh {
    const uint4*    nodes = nullptr;       // 4 x uint4 (64 B) per 4-wide inner node (layout: bvh.cuh)
    const float4*   tris = nullptr;        // 3 x float4 per triangle, leaf order
    const float4*   tri_rec = nullptr;     // 3 x float4 per triangle, global-id order
    const uint32_t* sorted_ids = nullptr;  // leaf slot -> global triangle id
    uint32_t        n_nodes = 0, n_tris = 0, max_depth = 0;   // max_depth: levels of the 4-wide tree
    uint32_t        ploc_rounds = 0;
    int             balanced_fallback = 0; // 1: the area-driven tree was too deep for the traversal stack, a balanced Morton tree was built instead
    int             wide = 4;              // branching factor (64-byte nodes, bvh.cuh)
    float           pad = 0.f;
    float           bounds[6] = {0, 0, 0, 0, 0, 0};
};

// d_vtx: float3 soup of all meshes; d_idx: uint3 per triangle, already rebased to the soup; d_tri_mesh: mesh id per triangle.
// Synchronises the stream (the caller is a one-off setup path, like optixAccelBuild + SYNC).
cudaError_t build_bvh(const float* d_vtx, const uint32_t* d_idx, const uint32_t* d_tri_mesh, uint32_t ntri, int leaf_max, int wide, cudaStream_t st,
                      BuiltBvh* out, uint64_t* launches);
void free_bvh(BuiltBvh* b);
// scene assembly (fovpt_scene_commit): idx_global[3t+k] = idx_local[3t+k] + vertex base of t's mesh, tri_mesh[t] = that mesh;
// mesh_table[m] = (first triangle, first vertex) of mesh m, ascending
void launch_rebase_indices(cudaStream_t st, const uint32_t* idx_local, const uint2* mesh_table, uint32_t n_meshes, uint32_t ntri,
                           uint32_t* idx_global, uint32_t* tri_mesh);

}  // namespace fovpt
