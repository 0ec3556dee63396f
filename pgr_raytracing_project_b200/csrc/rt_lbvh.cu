// rt_lbvh.cu -- BVH construction ON THE GPU (rt_build_bvh builder 1).
//
// The reference rebuilds its BVH on the host on every scene edit -- Scene::build_bvh, twice per
// RayTracer::set_scene (old/raytracer_core copy.cpp:84-87,162-167), 4.2 s for 1M spheres -- and the host
// calls set_scene on every drag / slider event (interaction.py:906,1169, gui.py:943).  This builder makes
// that path interactive: a linear BVH (Karras 2012: "Maximizing Parallelism in the Construction of BVHs,
// Octrees, and k-d Trees") over 63-bit Morton codes of the box centres, emitted directly in the layout the
// traversal kernels read (rt_bvh_node: root at 0, pad record at 1, sibling pairs adjacent, leaves of <= 4
// primitives, boxes padded by 2^-16 * scene scale exactly like the reference-order builder).
//
//   k_boxes        primitive boxes + centre bounds (ordered-int atomics)
//   k_morton       21 bits per axis, interleaved; key = code, value = primitive number
//   cub radix sort keys + values
//   k_hierarchy    one thread per internal node of the binary radix tree: range, split, children, parents
//   k_fit          one thread per leaf walks up; the second arrival at a node unions the children's boxes
//   cub scan       rank of the internal nodes that stay internal (range > 4 primitives)
//   k_emit         every such node writes its two child records at its pair slot: a child is an internal node
//                  (code = its pair slot) or a leaf (the <= 4 primitives of its range, sorted by number)
//   k_gather       primitive records in leaf order
// Closest-hit results do not depend on the tree (ties go to the lower primitive number, boxes are padded),
// so frames rendered over this tree are bit-identical to frames over the reference-order tree (asserted in
// tests/); only the traversal counters, i.e. speed, differ.
#include <cub/cub.cuh>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "rt_device.cuh"
#include "rt_lbvh.h"

namespace b200rt {

namespace {

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// boxes exactly as the host builder computes them (rt_bvh.cpp sphere_boxes / triangle_boxes)
__global__ void k_boxes(const float* __restrict__ raw, int is_tri, int n, float* __restrict__ lo, float* __restrict__ hi,
                        int* __restrict__ cbounds /* 6 ordered ints: centre min xyz, max xyz */) {
    float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        for (int c = 0; c < 3; ++c) {
            float l, h;
            if (is_tri) {
                float a = raw[9 * (size_t)i + c], b = raw[9 * (size_t)i + 3 + c], d = raw[9 * (size_t)i + 6 + c];
                l = fminf(a, fminf(b, d)); h = fmaxf(a, fmaxf(b, d));
            } else {
                l = __fsub_rn(raw[4 * (size_t)i + c], raw[4 * (size_t)i + 3]);
                h = __fadd_rn(raw[4 * (size_t)i + c], raw[4 * (size_t)i + 3]);
            }
            lo[3 * (size_t)i + c] = l; hi[3 * (size_t)i + c] = h;
            float ctr = __fmul_rn(__fadd_rn(l, h), 0.5f);
            cmin[c] = fminf(cmin[c], ctr); cmax[c] = fmaxf(cmax[c], ctr);
        }
    }
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            cmin[c] = fminf(cmin[c], __shfl_xor_sync(0xffffffffu, cmin[c], o));
            cmax[c] = fmaxf(cmax[c], __shfl_xor_sync(0xffffffffu, cmax[c], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(cbounds + c, f2ord(cmin[c])); atomicMax(cbounds + 3 + c, f2ord(cmax[c])); }
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {   // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float* __restrict__ lo, const float* __restrict__ hi, const int* __restrict__ cbounds, int n,
                         unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long code = 0;
        for (int c = 0; c < 3; ++c) {
            float mn = ord2f(cbounds[c]), mx = ord2f(cbounds[3 + c]);
            float ctr = (lo[3 * (size_t)i + c] + hi[3 * (size_t)i + c]) * 0.5f;
            float ext = mx - mn;
            float u = ext > 0.0f ? (ctr - mn) / ext : 0.0f;
            long long q = (long long)(u * 2097152.0f);
            q = q < 0 ? 0 : (q > 2097151 ? 2097151 : q);
            code |= spread21((unsigned long long)q) << c;
        }
        keys[i] = code; vals[i] = i;
    }
}

// common-prefix length of sorted keys i and j (ties broken by position), -1 outside the array
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// child encoding inside the radix tree: >= 0 internal node index, < 0 leaf ~k
__global__ void k_hierarchy(const unsigned long long* __restrict__ keys, int n, int* __restrict__ child_l, int* __restrict__ child_r,
                            int* __restrict__ first, int* __restrict__ last, int* __restrict__ parent_int,
                            int* __restrict__ parent_leaf) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
        const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = delta(keys, n, i, i - d);
        int lmax = 2;
        while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
        int l = 0;
        for (int t = lmax >> 1; t >= 1; t >>= 1)
            if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
        const int j = i + l * d;
        const int dnode = delta(keys, n, i, j);
        int s = 0, t = l;
        do {
            t = (t + 1) >> 1;
            if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        const int gamma = i + s * d + (d < 0 ? -1 : 0);
        const int lo_ = min(i, j), hi_ = max(i, j);
        const int cl = (lo_ == gamma) ? ~gamma : gamma;
        const int cr = (hi_ == gamma + 1) ? ~(gamma + 1) : gamma + 1;
        child_l[i] = cl; child_r[i] = cr; first[i] = lo_; last[i] = hi_;
        if (cl >= 0) parent_int[cl] = i; else parent_leaf[~cl] = i;
        if (cr >= 0) parent_int[cr] = i; else parent_leaf[~cr] = i;
        if (i == 0) parent_int[0] = -1;
    }
}

__global__ void k_fit(const float* __restrict__ lo, const float* __restrict__ hi, const int* __restrict__ sorted, int n,
                      const int* __restrict__ child_l, const int* __restrict__ child_r, const int* __restrict__ parent_int,
                      const int* __restrict__ parent_leaf, float* __restrict__ box /* 6 per internal */, int* __restrict__ flag) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        int p = parent_leaf[k];
        while (p >= 0) {
            if (atomicAdd(flag + p, 1) == 0) break;            // the sibling subtree is not done yet
            float b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
            const int ch[2] = {child_l[p], child_r[p]};
            for (int q = 0; q < 2; ++q) {
                const float* l; const float* h;
                if (ch[q] >= 0) { l = box + 6 * (size_t)ch[q]; h = l + 3; }
                else { int prim = sorted[~ch[q]]; l = lo + 3 * (size_t)prim; h = hi + 3 * (size_t)prim; }
                for (int c = 0; c < 3; ++c) { b[c] = fminf(b[c], __ldcg(l + c)); b[3 + c] = fmaxf(b[3 + c], __ldcg(h + c)); }
            }
            for (int c = 0; c < 6; ++c) box[6 * (size_t)p + c] = b[c];
            __threadfence();
            p = parent_int[p];
        }
    }
}

// levels of the EMITTED tree (internal nodes spanning > 4 primitives stay internal) = the traversal stack bound
__global__ void k_depth(int n, const int* __restrict__ parent_int, const int* __restrict__ parent_leaf, const int* __restrict__ first,
                        const int* __restrict__ last, int* __restrict__ max_depth) {
    int best = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        int depth = 1;
        for (int p = parent_leaf[k]; p >= 0; p = parent_int[p])
            if (last[p] - first[p] + 1 > 4) ++depth;
        best = max(best, depth);
    }
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, best);
}

__global__ void k_kept(const int* __restrict__ first, const int* __restrict__ last, int n_int, int* __restrict__ kept) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_int; i += gridDim.x * blockDim.x)
        kept[i] = (last[i] - first[i] + 1 > 4) ? 1 : 0;
}

__device__ __forceinline__ void write_node(rt_bvh_node* out, const float* b, float pad, int a, int cnt) {
    rt_bvh_node nd;
    for (int c = 0; c < 3; ++c) { nd.bmin[c] = __fsub_rn(b[c], pad); nd.bmax[c] = __fadd_rn(b[3 + c], pad); }
    nd.a = a; nd.b = cnt;
    *out = nd;
}

__global__ void k_emit(const float* __restrict__ lo, const float* __restrict__ hi, int* __restrict__ sorted, int n,
                       const int* __restrict__ child_l, const int* __restrict__ child_r, const int* __restrict__ first,
                       const int* __restrict__ last, const float* __restrict__ box, const int* __restrict__ kept,
                       const int* __restrict__ rank, rt_bvh_node* __restrict__ out) {
    float scale = 0.0f;
    for (int c = 0; c < 6; ++c) scale = fmaxf(scale, fabsf(box[c]));     // root box = internal node 0
    const float pad = __fmul_rn(scale, 0x1p-16f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
        if (!kept[i]) continue;
        const int pair = 2 + 2 * rank[i];
        if (i == 0) {
            write_node(out, box, pad, pair, 0);
            rt_bvh_node z; for (int c = 0; c < 3; ++c) { z.bmin[c] = 0.0f; z.bmax[c] = 0.0f; } z.a = 0; z.b = 0;
            out[1] = z;
        }
        const int ch[2] = {child_l[i], child_r[i]};
        for (int q = 0; q < 2; ++q) {
            const int c = ch[q];
            if (c >= 0 && kept[c]) { write_node(out + pair + q, box + 6 * (size_t)c, pad, 2 + 2 * rank[c], 0); continue; }
            int f, cnt;
            float b[6];
            if (c >= 0) {
                f = first[c]; cnt = last[c] - first[c] + 1;
                for (int k = 0; k < 6; ++k) b[k] = box[6 * (size_t)c + k];
            } else {
                f = ~c; cnt = 1;
                const int prim = sorted[f];
                for (int k = 0; k < 3; ++k) { b[k] = lo[3 * (size_t)prim + k]; b[3 + k] = hi[3 * (size_t)prim + k]; }
            }
            int v[4];                                              // canonical leaf order: ascending primitive number
            for (int k = 0; k < cnt; ++k) v[k] = sorted[f + k];
            for (int x = 1; x < cnt; ++x) { int key = v[x], y = x - 1; while (y >= 0 && v[y] > key) { v[y + 1] = v[y]; --y; } v[y + 1] = key; }
            for (int k = 0; k < cnt; ++k) sorted[f + k] = v[k];
            write_node(out + pair + q, b, pad, f, cnt);
        }
    }
}

// n <= 4: the root is the only node (a leaf)
__global__ void k_emit_tiny(const float* __restrict__ lo, const float* __restrict__ hi, int* __restrict__ sorted, int n,
                            rt_bvh_node* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    float b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) { b[c] = fminf(b[c], lo[3 * i + c]); b[3 + c] = fmaxf(b[3 + c], hi[3 * i + c]); }
    float scale = 0.0f;
    for (int c = 0; c < 6; ++c) scale = fmaxf(scale, fabsf(b[c]));
    for (int i = 0; i < n; ++i) sorted[i] = i;
    write_node(out, b, __fmul_rn(scale, 0x1p-16f), 0, n);
    rt_bvh_node z; for (int c = 0; c < 3; ++c) { z.bmin[c] = 0.0f; z.bmax[c] = 0.0f; } z.a = 0; z.b = 0;
    out[1] = z;
}

// ABI records -> device records (code in .a, see rt_device.cuh) and primitive records in leaf order
__global__ void k_device_nodes(const rt_bvh_node* __restrict__ abi, int n_nodes, rt_bvh_node* __restrict__ dev) {
    float4* nodes = reinterpret_cast<float4*>(dev);                    // sibling pairs interleaved (rt_device.cuh node_slot)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_nodes; k += gridDim.x * blockDim.x) {
        const rt_bvh_node nd = abi[k];
        node_write(nodes, k, nd.bmin, nd.bmax, nd.b == 0 ? nd.a : ~((nd.a << 3) | nd.b));
        if (k & 1) { float* f = reinterpret_cast<float*>(nodes) + (size_t)(k >> 1) * 16; f[14] = 0.0f; f[15] = 0.0f; }
    }
}

__global__ void k_gather(const float* __restrict__ raw, const int* __restrict__ mat_id, int is_tri, const int* __restrict__ prim_index,
                         int n, float4* __restrict__ prims) {
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const int p = prim_index[slot];
        if (is_tri) {
            const float* v = raw + 9 * (size_t)p;
            prims[kTriStride * (size_t)slot + 0] = make_float4(v[0], v[1], v[2], __int_as_float(p));
            prims[kTriStride * (size_t)slot + 1] = make_float4(__fsub_rn(v[3], v[0]), __fsub_rn(v[4], v[1]), __fsub_rn(v[5], v[2]),
                                                      __int_as_float(mat_id[p]));
            prims[kTriStride * (size_t)slot + 2] = make_float4(__fsub_rn(v[6], v[0]), __fsub_rn(v[7], v[1]), __fsub_rn(v[8], v[2]), 0.0f);
        } else {
            const float* s = raw + 4 * (size_t)p;
            prims[slot] = make_float4(s[0], s[1], s[2], s[3]);
        }
    }
}

inline int grid_of(int64_t n, int sm_count) {
    int64_t g = (n + 255) / 256;
    int64_t cap = (int64_t)sm_count * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

#define LB(call)                                    \
    do {                                            \
        cudaError_t e_ = (call);                    \
        if (e_ != cudaSuccess) { free_all(); cudaFree(abi); cudaFree(dev); cudaFree(prims); cudaFree(vals_sorted); return e_; } \
    } while (0)

cudaError_t lbvh_build(const float* d_raw, const int* d_mat_id, bool is_tri, int n, int sm_count, cudaStream_t stream,
                       LbvhResult* out) {
    *out = LbvhResult{};
    if (n <= 0) return cudaSuccess;
    float *lo = nullptr, *hi = nullptr, *box = nullptr;
    int *cbounds = nullptr, *vals = nullptr, *vals_sorted = nullptr, *child_l = nullptr, *child_r = nullptr, *first = nullptr,
        *last = nullptr, *parent_int = nullptr, *parent_leaf = nullptr, *flag = nullptr, *kept = nullptr, *rank = nullptr,
        *scalars = nullptr;
    unsigned long long *keys = nullptr, *keys_sorted = nullptr;
    void* tmp = nullptr;
    rt_bvh_node* abi = nullptr;
    rt_bvh_node* dev = nullptr;
    float4* prims = nullptr;
    auto free_all = [&]() {
        cudaFree(lo); cudaFree(hi); cudaFree(box); cudaFree(cbounds); cudaFree(vals); cudaFree(child_l); cudaFree(child_r);
        cudaFree(first); cudaFree(last); cudaFree(parent_int); cudaFree(parent_leaf); cudaFree(flag); cudaFree(kept);
        cudaFree(rank); cudaFree(scalars); cudaFree(keys); cudaFree(keys_sorted); cudaFree(tmp);
    };
    const int g = grid_of(n, sm_count);
    const size_t nn = (size_t)n, ni = (size_t)(n > 1 ? n - 1 : 1);
    static const bool trace = std::getenv("B200RT_TRACE") != nullptr;
    const auto t_host0 = std::chrono::steady_clock::now();
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (trace) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); }
    LB(cudaMalloc(&lo, nn * 12)); LB(cudaMalloc(&hi, nn * 12));
    LB(cudaMalloc(&cbounds, 6 * sizeof(int)));
    LB(cudaMalloc(&vals_sorted, nn * sizeof(int)));
    {
        const int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
        LB(cudaMemcpyAsync(cbounds, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    }
    if (trace) cudaEventRecord(ev0, stream);
    k_boxes<<<g, 256, 0, stream>>>(d_raw, is_tri ? 1 : 0, n, lo, hi, cbounds);
    int n_nodes = 2, depth = 1;
    LB(cudaMalloc(&scalars, 2 * sizeof(int)));
    if (n <= 4) {
        LB(cudaMalloc(&abi, 2 * sizeof(rt_bvh_node)));
        k_emit_tiny<<<1, 32, 0, stream>>>(lo, hi, vals_sorted, n, abi);
    } else {
        LB(cudaMalloc(&keys, nn * 8)); LB(cudaMalloc(&keys_sorted, nn * 8)); LB(cudaMalloc(&vals, nn * sizeof(int)));
        k_morton<<<g, 256, 0, stream>>>(lo, hi, cbounds, n, keys, vals);
        size_t tmp_bytes = 0, scan_bytes = 0;
        LB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_sorted, vals, vals_sorted, n, 0, 63, stream));
        LB(cudaMalloc(&kept, ni * sizeof(int))); LB(cudaMalloc(&rank, ni * sizeof(int)));
        LB(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, kept, rank, n - 1, stream));
        if (scan_bytes > tmp_bytes) tmp_bytes = scan_bytes;
        LB(cudaMalloc(&tmp, tmp_bytes));
        LB(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_sorted, vals, vals_sorted, n, 0, 63, stream));
        LB(cudaMalloc(&child_l, ni * sizeof(int))); LB(cudaMalloc(&child_r, ni * sizeof(int)));
        LB(cudaMalloc(&first, ni * sizeof(int))); LB(cudaMalloc(&last, ni * sizeof(int)));
        LB(cudaMalloc(&parent_int, ni * sizeof(int))); LB(cudaMalloc(&parent_leaf, nn * sizeof(int)));
        LB(cudaMalloc(&flag, ni * sizeof(int))); LB(cudaMalloc(&box, ni * 24));
        LB(cudaMemsetAsync(flag, 0, ni * sizeof(int), stream));
        LB(cudaMemsetAsync(scalars, 0, 2 * sizeof(int), stream));
        k_hierarchy<<<g, 256, 0, stream>>>(keys_sorted, n, child_l, child_r, first, last, parent_int, parent_leaf);
        k_fit<<<g, 256, 0, stream>>>(lo, hi, vals_sorted, n, child_l, child_r, parent_int, parent_leaf, box, flag);
        k_depth<<<g, 256, 0, stream>>>(n, parent_int, parent_leaf, first, last, scalars);
        k_kept<<<g, 256, 0, stream>>>(first, last, n - 1, kept);
        LB(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, kept, rank, n - 1, stream));
        int last_rank = 0, last_kept = 0;
        LB(cudaMemcpyAsync(&last_rank, rank + (n - 2), sizeof(int), cudaMemcpyDeviceToHost, stream));
        LB(cudaMemcpyAsync(&last_kept, kept + (n - 2), sizeof(int), cudaMemcpyDeviceToHost, stream));
        LB(cudaMemcpyAsync(&depth, scalars, sizeof(int), cudaMemcpyDeviceToHost, stream));
        LB(cudaStreamSynchronize(stream));
        n_nodes = 2 + 2 * (last_rank + last_kept);
        LB(cudaMalloc(&abi, (size_t)n_nodes * sizeof(rt_bvh_node)));
        k_emit<<<g, 256, 0, stream>>>(lo, hi, vals_sorted, n, child_l, child_r, first, last, box, kept, rank, abi);
    }
    cudaError_t e = cudaMalloc(&dev, (size_t)n_nodes * sizeof(rt_bvh_node));
    if (e == cudaSuccess) e = cudaMalloc(&prims, nn * (is_tri ? kTriStride : 1) * sizeof(float4));
    if (e != cudaSuccess) { cudaFree(abi); cudaFree(dev); cudaFree(prims); cudaFree(vals_sorted); free_all(); return e; }
    k_device_nodes<<<grid_of(n_nodes, sm_count), 256, 0, stream>>>(abi, n_nodes, dev);
    k_gather<<<g, 256, 0, stream>>>(d_raw, d_mat_id, is_tri ? 1 : 0, vals_sorted, n, prims);
    if (trace) cudaEventRecord(ev1, stream);
    e = cudaStreamSynchronize(stream);
    if (trace) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
        std::fprintf(stderr, "lbvh_build: n=%d nodes=%d depth=%d  first kernel..last kernel %.3f ms (incl. allocations in between), host wall %.3f ms\n",
                     n, n_nodes, depth, ms, host_ms);
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    free_all();
    if (e != cudaSuccess) { cudaFree(abi); cudaFree(dev); cudaFree(prims); cudaFree(vals_sorted); return e; }
    out->d_nodes_abi = abi; out->d_nodes = reinterpret_cast<float4*>(dev); out->n_nodes = n_nodes;
    out->d_prim_index = vals_sorted; out->d_prims = prims; out->depth = depth;
    return cudaSuccess;
}

}  // namespace b200rt
