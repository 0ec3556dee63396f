// rt_refit.h -- refit of the current BVH after its primitives have moved (rt_update_geometry).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200rt.h"

namespace b200rt {

// d_nodes: the tree in traversal layout (code in .a, rt_device.cuh), boxes rewritten in place; d_nodes_abi: the same tree
// in include/b200rt.h layout, written from scratch (n_nodes records); d_prims: primitive records in leaf order, rewritten
// from d_raw (primitives as uploaded: n x 4 spheres / n x 9 triangles) through d_slot_prim; d_scratch: n_nodes * 32 bytes.
// Enqueues on `stream`; no synchronisation.
cudaError_t bvh_refit(float4* d_nodes, rt_bvh_node* d_nodes_abi, int n_nodes, const int* d_slot_prim, const float* d_raw,
                      bool is_tri, float4* d_prims, int n, void* d_scratch, int sm_count, cudaStream_t stream);

// *d_out (device) = sum over the internal nodes of the half surface area of their boxes.
cudaError_t bvh_area(const float4* d_nodes, int n_nodes, double* d_out, int sm_count, cudaStream_t stream);

}  // namespace b200rt
