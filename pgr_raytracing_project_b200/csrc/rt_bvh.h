// rt_bvh.h -- host-side BVH construction for libb200rt (reference-order median split).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/b200rt.h"

namespace b200rt {

struct PrimBoxes {          // per primitive, upload order
    std::vector<float> lo, hi;   // n x 3
};

// Axis-aligned boxes of the primitives exactly as uploaded.
void sphere_boxes(const float* center_radius, int64_t n, PrimBoxes& out);
void triangle_boxes(const float* vertices9, int64_t n, PrimBoxes& out);

// Reference builder (cpp_raytracer/raytracer_core.cpp:57-118 / old/bvh copy.cpp:111-174):
// top-down, leaf when span <= 4, split axis = longest extent (x; y if ey > ex; z if ez > ey and
// ez > ex), order by box centre along that axis (ties by primitive number), split at
// start + span/2.  Output in the rt_bvh_node layout: root at 0, pad record at 1, sibling pairs
// allocated depth-first (left subtree before right), leaf primitives ascending by number,
// every box padded outward by 2^-16 * max |coordinate| of the root box.
// Deterministic for any thread count.
void build_median_split(const PrimBoxes& boxes, int64_t n, std::vector<rt_bvh_node>& nodes,
                        std::vector<int32_t>& prim_index, int leaf_size = 4);   // leaf_size 1..4 (4 = the reference's rule)

// Binned surface-area-heuristic builder (builder 2): NOT the reference's tree -- same layout, same closest hits, fewer
// steps per ray.  Deterministic for any thread count.
// trav_cost: cost of one traversal step in primitive tests -- a range of <= leaf_size primitives is cut further while
// cost(split) + trav_cost * area < area * count.
void build_sah(const PrimBoxes& boxes, int64_t n, std::vector<rt_bvh_node>& nodes, std::vector<int32_t>& prim_index, int leaf_size = 4,
               float trav_cost = 3.0f);

// Structural validation of a caller-supplied tree; returns max depth or -1 (msg filled).
int validate_bvh(const rt_bvh_node* nodes, int64_t n_nodes, int64_t n_prims, const char** msg);

}  // namespace b200rt
