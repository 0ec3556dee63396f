// rt_refit.cu -- BVH refit ON THE GPU for scene edits that move primitives but keep their number.
//
// The reference's host calls set_scene -- a full copy and TWO full BVH builds (old/raytracer_core copy.cpp:84-87,
// 162-167) -- on every drag / slider event (interaction.py:906,1169, gui.py:943).  An edit that only moves
// primitives does not need a new tree: the topology stays valid, only the boxes are stale.  bvh_refit keeps the
// topology of whatever tree the context holds (reference-order median split, device LBVH, or one set by rt_set_bvh)
// and recomputes every box bottom-up:
//   k_regather   primitive records in leaf order from the new raw primitives (prim / material words kept)
//   k_parents    parent of every node from the child-pair codes (the traversal layout stores no parent links)
//   k_fit        one thread per LEAF: box of its <= 7 primitives exactly as the builders compute it, then up the
//                tree -- the second arrival at a node (one atomic per node) unions the sibling pair below it
//   k_finish     pad by 2^-16 * scene scale like the builders and write both layouts
// Closest-hit results do not depend on the tree (rt_device.cuh consider()), so a frame over the refitted tree is
// bit-identical to a frame over a tree rebuilt from scratch (asserted in tests/); what a refit can lose is tree
// QUALITY after large moves -- the caller decides when to rebuild (rt_build_bvh).
#include "rt_refit.h"

#include "rt_device.cuh"

namespace b200rt {

namespace {

__global__ void k_regather(const float* __restrict__ raw, const int* __restrict__ slot_prim, int is_tri, int n,
                           float4* __restrict__ prims) {
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const int p = slot_prim[slot];
        if (is_tri) {
            const float* v = raw + 9 * (size_t)p;
            float4* o = prims + kTriStride * (size_t)slot;
            const float mat = o[1].w;                                   // material word stays
            o[0] = make_float4(v[0], v[1], v[2], __int_as_float(p));
            o[1] = make_float4(__fsub_rn(v[3], v[0]), __fsub_rn(v[4], v[1]), __fsub_rn(v[5], v[2]), mat);
            o[2] = make_float4(__fsub_rn(v[6], v[0]), __fsub_rn(v[7], v[1]), __fsub_rn(v[8], v[2]), 0.0f);
        } else {
            const float* s = raw + 4 * (size_t)p;
            prims[slot] = make_float4(s[0], s[1], s[2], s[3]);
        }
    }
}

__global__ void k_parents(const float4* __restrict__ nodes, int n_nodes, int* __restrict__ parent, int* __restrict__ flag) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_nodes; k += gridDim.x * blockDim.x) {
        flag[k] = 0;
        if (k == 0) parent[0] = -1;
        if (k == 1) continue;                                           // pad record
        const int code = node_code(nodes, k);
        if (code >= 0) { parent[code] = k; parent[code + 1] = k; }
    }
}

// box[6 * k ..] = unpadded box of node k
__global__ void k_fit(const float4* __restrict__ nodes, int n_nodes, const int* __restrict__ slot_prim, const float* __restrict__ raw,
                      int is_tri, const int* __restrict__ parent, int* __restrict__ flag, float* __restrict__ box) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_nodes; k += gridDim.x * blockDim.x) {
        if (k == 1) continue;
        const int code = node_code(nodes, k);
        if (code >= 0) continue;                                        // internal: filled in by its second child
        const int first = (~code) >> 3, count = (~code) & 7;
        float b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int q = 0; q < count; ++q) {
            const int p = slot_prim[first + q];
            for (int c = 0; c < 3; ++c) {
                float l, h;
                if (is_tri) {                                           // rt_bvh.cpp triangle_boxes
                    const float a = raw[9 * (size_t)p + c], bb = raw[9 * (size_t)p + 3 + c], d = raw[9 * (size_t)p + 6 + c];
                    l = fminf(a, fminf(bb, d)); h = fmaxf(a, fmaxf(bb, d));
                } else {                                                // rt_bvh.cpp sphere_boxes
                    l = __fsub_rn(raw[4 * (size_t)p + c], raw[4 * (size_t)p + 3]);
                    h = __fadd_rn(raw[4 * (size_t)p + c], raw[4 * (size_t)p + 3]);
                }
                b[c] = fminf(b[c], l); b[3 + c] = fmaxf(b[3 + c], h);
            }
        }
        int at = k;
        for (;;) {
            for (int c = 0; c < 6; ++c) box[6 * (size_t)at + c] = b[c];
            const int p = parent[at];
            if (p < 0) break;
            __threadfence();                                            // this box before the arrival count
            if (atomicAdd(flag + p, 1) == 0) break;                     // the sibling subtree is not done yet
            __threadfence();
            const int sib = at ^ 1;                                     // sibling pairs are adjacent and even-aligned
            for (int c = 0; c < 3; ++c) {
                b[c] = fminf(b[c], __ldcg(box + 6 * (size_t)sib + c));
                b[3 + c] = fmaxf(b[3 + c], __ldcg(box + 6 * (size_t)sib + 3 + c));
            }
            at = p;
        }
    }
}

__global__ void k_finish(const float* __restrict__ box, int n_nodes, float4* __restrict__ nodes, rt_bvh_node* __restrict__ abi) {
    float scale = 0.0f;
    for (int c = 0; c < 6; ++c) scale = fmaxf(scale, fabsf(box[c]));     // root box (rt_bvh.cpp build_median_split)
    const float pad = __fmul_rn(scale, 0x1p-16f);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_nodes; k += gridDim.x * blockDim.x) {
        rt_bvh_node nd;
        if (k == 1) {
            for (int c = 0; c < 3; ++c) { nd.bmin[c] = 0.0f; nd.bmax[c] = 0.0f; }
            nd.a = 0; nd.b = 0;
            abi[1] = nd;
            continue;
        }
        const int code = node_code(nodes, k);
        float lo[3], hi[3];
        for (int c = 0; c < 3; ++c) { lo[c] = __fsub_rn(box[6 * (size_t)k + c], pad); hi[c] = __fadd_rn(box[6 * (size_t)k + 3 + c], pad); }
        node_write(nodes, k, lo, hi, code);
        for (int c = 0; c < 3; ++c) { nd.bmin[c] = lo[c]; nd.bmax[c] = hi[c]; }
        if (code >= 0) { nd.a = code; nd.b = 0; }
        else { nd.a = (~code) >> 3; nd.b = (~code) & 7; }
        abi[k] = nd;
    }
}

// sum of the half surface areas of the internal nodes' boxes: the tree-quality figure a refit can only make worse
__global__ void k_area(const float4* __restrict__ nodes, int n_nodes, double* __restrict__ out) {
    double part = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_nodes; k += gridDim.x * blockDim.x) {
        if (k == 1) continue;
        float lo[3], hi[3];
        int code;
        node_read(nodes, k, lo, hi, code);
        if (code < 0) continue;
        const double ex = (double)hi[0] - lo[0], ey = (double)hi[1] - lo[1], ez = (double)hi[2] - lo[2];
        part += ex * ey + ey * ez + ez * ex;
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0 && part != 0.0) atomicAdd(out, part);
}

inline int grid_of(int64_t n, int sm_count) {
    int64_t g = (n + 255) / 256, cap = (int64_t)sm_count * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

cudaError_t bvh_refit(float4* d_nodes, rt_bvh_node* d_nodes_abi, int n_nodes, const int* d_slot_prim, const float* d_raw,
                      bool is_tri, float4* d_prims, int n, void* d_scratch, int sm_count, cudaStream_t stream) {
    if (n <= 0 || n_nodes <= 0) return cudaSuccess;
    int* parent = static_cast<int*>(d_scratch);                          // n_nodes ints
    int* flag = parent + n_nodes;                                        // n_nodes ints
    float* box = reinterpret_cast<float*>(flag + n_nodes);               // 6 * n_nodes floats  (total 32 bytes per node)
    k_regather<<<grid_of(n, sm_count), 256, 0, stream>>>(d_raw, d_slot_prim, is_tri ? 1 : 0, n, d_prims);
    k_parents<<<grid_of(n_nodes, sm_count), 256, 0, stream>>>(d_nodes, n_nodes, parent, flag);
    k_fit<<<grid_of(n_nodes, sm_count), 256, 0, stream>>>(d_nodes, n_nodes, d_slot_prim, d_raw, is_tri ? 1 : 0, parent, flag, box);
    k_finish<<<grid_of(n_nodes, sm_count), 256, 0, stream>>>(box, n_nodes, d_nodes, d_nodes_abi);
    return cudaGetLastError();
}

cudaError_t bvh_area(const float4* d_nodes, int n_nodes, double* d_out, int sm_count, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(double), stream);
    if (e != cudaSuccess) return e;
    if (n_nodes > 0) k_area<<<grid_of(n_nodes, sm_count), 256, 0, stream>>>(d_nodes, n_nodes, d_out);
    return cudaGetLastError();
}

}  // namespace b200rt
