// rt_lbvh.h -- GPU BVH builder (rt_lbvh.cu): linear BVH over Morton codes, emitted in the rt_bvh_node layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200rt.h"

namespace b200rt {

struct LbvhResult {              // device arrays, owned by the caller after a successful build (cudaFree)
    rt_bvh_node* d_nodes_abi = nullptr;   // include/b200rt.h layout (what rt_get_bvh returns)
    float4* d_nodes = nullptr;            // traversal layout (code in .a, see rt_device.cuh)
    int n_nodes = 0;
    int* d_prim_index = nullptr;          // slot -> primitive number
    float4* d_prims = nullptr;            // primitive records in leaf order
    int depth = 0;                        // levels of the emitted tree
};

// d_raw: primitives as uploaded, on the device (n x 4 spheres or n x 9 triangles); d_mat_id: per-triangle
// material row (ignored for spheres).  Synchronises `stream`.
cudaError_t lbvh_build(const float* d_raw, const int* d_mat_id, bool is_tri, int n, int sm_count, cudaStream_t stream,
                       LbvhResult* out);

}  // namespace b200rt
