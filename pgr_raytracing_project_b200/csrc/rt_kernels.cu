// rt_kernels.cu -- sm_100a kernels of the render hot path and their launchers.
// Compile with: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo (see build.py).
//
// Kernel family (DESIGN.md "Kernels"):
//   k_trace_primary  camera ray through each pixel centre -> closest hit (prim, t)      [AOV]
//   k_trace_rays     arbitrary rays -> closest hit
//   k_render         megakernel: spp jittered camera samples per pixel, path tracing, resolve
//   k_untile / k_accumulate / k_tonemap_u8   framebuffer plumbing (HBM-bound elementwise)
// All tracing kernels are persistent: grid = SMs x resident CTAs, each warp pulls one 8x4 pixel
// block at a time from a global counter (one atomic per warp, broadcast by shuffle), so rays of a
// warp stay spatially coherent while load imbalance between hit/miss regions is evened out.
#include "rt_kernels.h"
#include "rt_kernel_common.cuh"

namespace b200rt {

namespace {

template <bool TRI, bool STATS>
__global__ void __launch_bounds__(kThreads)
k_trace_primary(const __grid_constant__ SceneView sc, const __grid_constant__ CameraBlock cam,
                const __grid_constant__ TileMap tm, int n_work, int32_t* __restrict__ d_prim,
                float* __restrict__ d_t, unsigned int* counter, unsigned long long* d_stats) {
    const int lane = threadIdx.x & 31;
    const double inv_w = __ddiv_rn(1.0, (double)tm.width), inv_h = __ddiv_rn(1.0, (double)tm.height);
    Counters cnt = {0, 0, 0};
    unsigned long long rays = 0;
    for (;;) {
        int w = next_work(counter, lane);
        if (w >= n_work) break;
        PixelWork p = decode_work(tm, w, lane);
        if (!p.active) continue;
        Ray r = camera_ray(cam, p.i, p.j, 0.5f, 0.5f, inv_w, inv_h);
        Hit h;
        intersect<TRI, STATS>(sc, r, h, cnt, true);
        d_prim[p.out_index] = h.prim;
        d_t[p.out_index] = h.prim >= 0 ? h.t : 0.0f;
        if (STATS) { rays += 1; cnt.segments += 1; }
    }
    if (STATS) flush_stats(d_stats, rays, cnt);
}

template <bool TRI, bool STATS>
__global__ void __launch_bounds__(kThreads)
k_trace_rays(const __grid_constant__ SceneView sc, const float* __restrict__ org, const float* __restrict__ dir,
             int64_t n, int32_t* __restrict__ d_prim, float* __restrict__ d_t, unsigned long long* d_stats) {
    Counters cnt = {0, 0, 0};
    unsigned long long rays = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        float dx = dir[3 * k], dy = dir[3 * k + 1], dz = dir[3 * k + 2];
        normalize3(dx, dy, dz);
        Ray r = make_ray(org[3 * k], org[3 * k + 1], org[3 * k + 2], dx, dy, dz);
        Hit h;
        intersect<TRI, STATS>(sc, r, h, cnt, false);
        d_prim[k] = h.prim;
        d_t[k] = h.prim >= 0 ? h.t : 0.0f;
        if (STATS) { rays += 1; cnt.segments += 1; }
    }
    if (STATS) flush_stats(d_stats, rays, cnt);
}

template <bool TRI, bool STATS>
__global__ void __launch_bounds__(kThreads)
k_render(const __grid_constant__ SceneView sc, const __grid_constant__ CameraBlock cam,
         const __grid_constant__ TileMap tm, int n_work, int spp, int max_depth, int integrator, uint32_t k0,
         uint32_t k1, uint32_t sample_offset, int resolve, float* __restrict__ d_out, unsigned int* counter,
         unsigned long long* d_stats) {
    const int lane = threadIdx.x & 31;
    const double inv_w = __ddiv_rn(1.0, (double)tm.width), inv_h = __ddiv_rn(1.0, (double)tm.height);
    const float inv_spp = __fdiv_rn(1.0f, (float)spp);
    Counters cnt = {0, 0, 0};
    unsigned long long rays = 0;
    for (;;) {
        int w = next_work(counter, lane);
        if (w >= n_work) break;
        PixelWork p = decode_work(tm, w, lane);
        if (!p.active) continue;
        uint32_t pixel = (uint32_t)(p.j * tm.width + p.i);
        float sr = 0.0f, sg = 0.0f, sb = 0.0f;
        for (int s = 0; s < spp; ++s) {
            float cr, cg, cb;
            radiance<TRI, STATS>(sc, cam, p.i, p.j, pixel, sample_offset + (uint32_t)s, max_depth, integrator, k0,
                                 k1, inv_w, inv_h, cr, cg, cb, cnt);
            sr = __fadd_rn(sr, cr); sg = __fadd_rn(sg, cg); sb = __fadd_rn(sb, cb);
        }
        if (STATS) rays += (unsigned long long)spp;
        float* o = d_out + 3 * (size_t)p.out_index;
        if (resolve) { o[0] = resolve1(sr, inv_spp); o[1] = resolve1(sg, inv_spp); o[2] = resolve1(sb, inv_spp); }
        else { o[0] = sr; o[1] = sg; o[2] = sb; }
    }
    if (STATS) flush_stats(d_stats, rays, cnt);
}

// ------------------------------------------------------------------------------------------------
// k_path: persistent-thread path tracer with lane-level continuation (the default kernel).
// Work unit = one pixel (all of its samples, in order, so the sample sum is deterministic).  Each
// lane is a small state machine {NONE, START, TRAV, SHADE}.  A warp traverses until fewer than
// `refill_below` of its lanes still traverse; the finished lanes are then shaded (next bounce or
// next sample) or refilled with new pixels from the global counter -- one atomic per refill, the
// requesting lanes ranked with ballot/popc -- and the warp re-enters the traversal loop with the
// unfinished lanes resuming exactly where they were.  AOV = true is the primary-hit query
// (pixel-centre ray, writes prim/t instead of radiance).
enum { PH_NONE = 0, PH_START = 1, PH_TRAV = 2, PH_SHADE = 3 };

template <bool TRI, bool STATS, bool AOV>
__global__ void __launch_bounds__(kThreads)
k_path(const __grid_constant__ SceneView sc, const __grid_constant__ CameraBlock cam,
       const __grid_constant__ TileMap tm, int n_tasks, int spp, int max_depth, int integrator, uint32_t k0,
       uint32_t k1, uint32_t sample_offset, int resolve, int refill_below, int leaf_vote, float* __restrict__ d_out,
       int32_t* __restrict__ d_prim, float* __restrict__ d_t, unsigned int* counter, unsigned long long* d_stats) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const double inv_w = __ddiv_rn(1.0, (double)tm.width), inv_h = __ddiv_rn(1.0, (double)tm.height);
    const float inv_spp = __fdiv_rn(1.0f, (float)spp);
    int stack_code[kStackDepth];
    float stack_tn[kStackDepth];
    Trav tv;
    tv.cur = kDone; tv.sp = 0; tv.h.t = kTMax; tv.h.prim = -1; tv.h.slot = -1;
    Counters cnt = {0, 0, 0};
    unsigned long long rays = 0;
    int phase = PH_NONE;
    int pi = 0, pj = 0, out_index = 0, s = 0, b = 0;
    uint32_t pixel = 0, ctl_z = 0, ctl_w = 0;
    Ray r = make_ray(0.f, 0.f, 0.f, 0.f, 0.f, 1.f);
    float tr = 1.f, tg = 1.f, tb = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
    bool pool_empty = false;

    for (;;) {
        // ---- A: lanes outside the traversal loop: start a sample / shade a finished segment
        while (phase == PH_START || phase == PH_SHADE) {
            if (phase == PH_START) {
                float jx = 0.5f, jy = 0.5f;
                if (!AOV) {
                    uint4 ctl = philox4x32_10(pixel, sample_offset + (uint32_t)s, 0u, 0u, k0, k1);
                    jx = u01(ctl.x); jy = u01(ctl.y); ctl_z = ctl.z; ctl_w = ctl.w;
                }
                r = camera_ray(cam, pi, pj, jx, jy, inv_w, inv_h);
                tr = tg = tb = 1.0f; cr = cg = cb = 0.0f; b = 0;
                if (STATS) rays += 1;
                trav_begin<STATS>(sc, r, tv, cnt);
                phase = tv.cur != kDone ? PH_TRAV : PH_SHADE;
            } else {
                if (STATS) cnt.segments += 1;
                if (AOV) {
                    d_prim[out_index] = tv.h.prim;
                    d_t[out_index] = tv.h.prim >= 0 ? tv.h.t : 0.0f;
                    phase = PH_NONE;
                    break;
                }
                bool path_end = true;
                if (tv.h.prim < 0) {
                    cr = __fmaf_rn(tr, sc.bg_r, cr); cg = __fmaf_rn(tg, sc.bg_g, cg); cb = __fmaf_rn(tb, sc.bg_b, cb);
                } else {
                    const float4* mp = sc.mats + 2 * (size_t)material_row<TRI>(sc, tv.h);
                    float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
                    cr = __fmaf_rn(tr, m1.y, cr); cg = __fmaf_rn(tg, m1.z, cg); cb = __fmaf_rn(tb, m1.w, cb);
                    if (b + 1 < max_depth) {
                        uint32_t sample = sample_offset + (uint32_t)s;
                        uint4 ctl = make_uint4(0u, 0u, ctl_z, ctl_w);
                        if (b > 0) ctl = philox4x32_10(pixel, sample, (uint32_t)b, 0u, k0, k1);
                        if (scatter<TRI>(sc, tv.h, r, integrator, b, max_depth, ctl, m0, m1, pixel, sample, k0, k1, tr, tg, tb)) {
                            ++b;
                            trav_begin<STATS>(sc, r, tv, cnt);
                            phase = tv.cur != kDone ? PH_TRAV : PH_SHADE;
                            path_end = false;
                        }
                    }
                }
                if (path_end) {
                    sr = __fadd_rn(sr, cr); sg = __fadd_rn(sg, cg); sb = __fadd_rn(sb, cb);
                    if (++s < spp) phase = PH_START;
                    else {
                        float* o = d_out + 3 * (size_t)out_index;
                        if (resolve) { o[0] = resolve1(sr, inv_spp); o[1] = resolve1(sg, inv_spp); o[2] = resolve1(sb, inv_spp); }
                        else { o[0] = sr; o[1] = sg; o[2] = sb; }
                        phase = PH_NONE;
                    }
                }
            }
        }
        // ---- B: refill idle lanes with new pixels (warp-aggregated atomic)
        unsigned need = __ballot_sync(0xffffffffu, phase == PH_NONE);
        if (need != 0u && !pool_empty) {
            int n_need = __popc(need);
            int leader = __ffs(need) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(counter, (unsigned)n_need);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (phase == PH_NONE) {
                unsigned idx = base + (unsigned)__popc(need & lt_mask);
                if (idx < (unsigned)n_tasks) {
                    PixelWork p = decode_work(tm, (int)(idx >> 5), (int)(idx & 31u));
                    if (p.active) {
                        pi = p.i; pj = p.j; out_index = p.out_index;
                        pixel = (uint32_t)(p.j * tm.width + p.i);
                        s = 0; sr = sg = sb = 0.0f;
                        phase = PH_START;
                    }
                }
            }
            if (base + (unsigned)n_need >= (unsigned)n_tasks) pool_empty = true;
            continue;
        }
        // ---- C: traverse
        unsigned act = __ballot_sync(0xffffffffu, phase == PH_TRAV);
        if (act == 0u) break;                      // every lane idle and the pool is empty
        // leave the loop again once a quarter (32 - refill_below in 32) of the lanes that entered are done
        int min_active = (__popc(act) * refill_below) >> 5;
        trav_run<TRI, STATS, 2>(sc, r, tv, LocalStack{stack_code, stack_tn}, min_active < 1 ? 1 : min_active, leaf_vote, cnt, b == 0);
        if (phase == PH_TRAV && tv.cur == kDone) phase = PH_SHADE;
    }
    if (STATS) flush_stats(d_stats, rays, cnt);
}

// ------------------------------------------------------------------------------------------------
// k_cam_tris: per-frame table of camera-relative triangle records (rt_device.cuh cam_tri_record) for
// the shared origin of all camera rays: 48 B read + 48 B written per triangle, HBM-streaming.
__global__ void __launch_bounds__(256)
k_cam_tris(const float4* __restrict__ prims, float4* __restrict__ cam_prims, int n, float ox, float oy, float oz) {
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const float4* p = prims + kTriStride * (size_t)slot;
        float4 r0, r1, r2, A, B, C;
        cam_tri_record(__ldg(p), __ldg(p + 1), __ldg(p + 2), ox, oy, oz, r0, r1, r2);
        cam_tri_pack(r0, r1, r2, A, B, C);
        float4* o = cam_prims + 3 * (size_t)slot;
        o[0] = A; o[1] = B; o[2] = C;
    }
}

// k_packet: camera rays only (max_depth 1 renders and the primary-hit AOV).  One warp = one 8x4
// pixel block = one packet walking the BVH with packet_intersect().  Work distribution keeps the
// packets that are in flight on one SM next to each other in the image, so that they share the node
// and triangle lines they pull into that SM's L1: a CTA (8 warps) owns one CHUNK of kChunk
// consecutive blocks at a time (a 32x8-pixel strip); its warps take blocks from the chunk through a
// shared-memory word (chunk << 8 | next offset, one ATOMS per block), and the warp that takes the
// last offset + 1 fetches the CTA's next chunk from the global counter (one ATOMG per chunk) while
// the others that run out wait for the word to change.  No CTA barrier, no fixed assignment: load
// stays balanced to within one block per warp.
// One CTA builds this frame's chunk order from last frame's costs (ChunkSchedule): latest-start-time
// first.  A chunk's 8 blocks run side by side on the CTA's 8 warps, so its duration is its largest block
// cost; the frame's length is T = (sum of all block costs) / resident warps; a chunk is due at T, or --
// when finished regions are being copied out while the kernel runs (BandSignal) -- at the share of T at
// which its region should be complete (regions in raster order).  key = due - duration, counting-sorted into 256 classes (order
// inside a class is whatever the atomics give: pure scheduling, pixels do not depend on it).  With one
// common due date this is longest-processing-time-first.  No history (all zero): raster order.
// The costs may have been accumulated over n_frames frames (the order is not rebuilt every frame).
__global__ void __launch_bounds__(1024)
k_chunk_order(unsigned int* __restrict__ cost_sum, unsigned int* __restrict__ cost_max, int* __restrict__ order, int n,
              int n_warps, int n_frames, const __grid_constant__ BandSignal band) {
    __shared__ unsigned long long s_sum;
    __shared__ unsigned int s_max;
    __shared__ int s_hist[256];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) { s_sum = 0ull; s_max = 0u; }
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    unsigned long long part = 0;
    unsigned int pm = 0;
    for (int k = tid; k < n; k += 1024) { part += cost_sum[k]; pm = max(pm, cost_max[k]); }
    for (int o = 16; o > 0; o >>= 1) { part += __shfl_xor_sync(0xffffffffu, part, o); pm = max(pm, __shfl_xor_sync(0xffffffffu, pm, o)); }
    if (lane == 0 && part) { atomicAdd(&s_sum, part); atomicMax(&s_max, pm); }
    __syncthreads();
    if (s_sum == 0ull) {
        for (int k = tid; k < n; k += 1024) order[k] = k;
        return;
    }
    const float T = (float)s_sum / ((float)(n_warps > 0 ? n_warps : 1) * (float)(n_frames > 0 ? n_frames : 1));   // costs of n_frames frames
    const float kmin = -(float)s_max, scale = 255.0f / (T - kmin);
    auto cls_of = [&](int k) {
        float due = T;
        if (band.cnt != nullptr) {
            const int n_regions = ((band.tiles_y + band.band_rows - 1) / band.band_rows) * band.n_groups;
            due = T * (float)(region_of_tile(band, (k * kChunk) >> 5) + 1) / (float)n_regions;
        }
        float key = due - (float)cost_max[k];
        int c = (int)((key - kmin) * scale);
        return c < 0 ? 0 : (c > 255 ? 255 : c);
    };
    for (int k = tid; k < n; k += 1024) atomicAdd(&s_hist[cls_of(k)], 1);
    __syncthreads();
    if (tid < 32) {                                          // exclusive scan of 256 counts: 8 per lane
        int v[8], t = 0;
        for (int q = 0; q < 8; ++q) { v[q] = s_hist[tid * 8 + q]; t += v[q]; }
        int x = t;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        int base = x - t;
        for (int q = 0; q < 8; ++q) { s_hist[tid * 8 + q] = base; base += v[q]; }
    }
    __syncthreads();
    for (int k = tid; k < n; k += 1024) order[atomicAdd(&s_hist[cls_of(k)], 1)] = k;
    __syncthreads();
    for (int k = tid; k < n; k += 1024) { cost_sum[k] = 0u; cost_max[k] = 0u; }
}

// Tile push (BandSignal::host_fb): one warp copies the finished 32x32-pixel tile (ty, tx) of the device frame `fb`
// into the caller's page-locked frame `host` (its device alias): rows of 384 contiguous bytes, read from L2 (other
// SMs wrote them), 16 bytes per lane when the row pitch allows it.  Kept out of line: k_packet's register budget.
__device__ __noinline__ void push_tile(const float* __restrict__ fb, float* __restrict__ host, int width, int height, int ty,
                                       int tx, int lane) {
    const int x0 = tx * 32, y0 = ty * 32;
    const int rows = min(32, height - y0), cols = min(32, width - x0);
    const size_t base = ((size_t)y0 * width + x0) * 3;
    if (cols == 32 && (width & 3) == 0) {
        const size_t pitch4 = (size_t)width * 3 / 4;
        const float4* s = reinterpret_cast<const float4*>(fb + base) + lane;
        float4* d = reinterpret_cast<float4*>(host + base) + lane;
        if (lane < 24) {
#pragma unroll 4
            for (int r = 0; r < rows; ++r) d[(size_t)r * pitch4] = __ldcg(s + (size_t)r * pitch4);
        }
    } else {
        const size_t pitch = (size_t)width * 3;
        for (int r = 0; r < rows; ++r)
            for (int k = lane; k < cols * 3; k += 32) host[base + (size_t)r * pitch + k] = __ldcg(fb + base + (size_t)r * pitch + k);
    }
}

// Top treelet for shared-memory staging: the first `levels` levels below the root as a HEAP of sibling pairs (pair 0 = the
// root's children; the children of pair s's left / right node are pairs 2s + 1 / 2s + 2), 4 float4 per pair in the device
// node format, child codes rewritten for a kernel that holds the heap in shared memory: code < two_t = node record of the
// heap (2 x pair), code >= two_t = global node record code - two_t; leaf codes unchanged.  Pairs below a leaf stay empty.
__global__ void __launch_bounds__(256)
k_build_treelet(const float4* __restrict__ nodes, int n_pairs, float4* __restrict__ out) {
    const int two_t = 2 * n_pairs;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_pairs; s += gridDim.x * blockDim.x) {
        int g = node_code(nodes, 0);                           // the root's child pair (the root is internal: checked by the host)
        const unsigned path = (unsigned)s + 1u;                // 1-based heap index: the bits below the leading one, top down
        bool exists = true;
        for (int bit = 30 - __clz(path); bit >= 0 && exists; --bit) {
            const int side = (path >> bit) & 1u;
            const int c = node_code(nodes, g + side);
            if (c < 0) exists = false; else g = c;
        }
        // a pair record as the traversal reads it (rt_device.cuh: interleaved; the two codes are .z / .w of the second float4)
        float4 rec[4] = {make_float4(0, 0, 0, 0), make_float4(0, 0, __int_as_float(-1), __int_as_float(-1)), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
        if (exists) {
            for (int k = 0; k < 4; ++k) rec[k] = nodes[2 * (size_t)g + k];
            int code[2] = {__float_as_int(rec[1].z), __float_as_int(rec[1].w)};
            for (int side = 0; side < 2; ++side) {
                if (code[side] >= 0) {
                    const int child = 2 * s + 1 + side;
                    code[side] = child < n_pairs ? 2 * child : code[side] + two_t;
                }
            }
            rec[1].z = __int_as_float(code[0]); rec[1].w = __int_as_float(code[1]);
        }
        for (int k = 0; k < 4; ++k) out[4 * (size_t)s + k] = rec[k];
    }
}

// rt_render_tiles_host: one warp per tile of this launch's tile map (32x32 tiles): the tile's rows from the device frame
// (L2) into the shared page-locked host frame, 384 contiguous bytes per row; the last warp to finish publishes `epoch`.
__global__ void __launch_bounds__(256)
k_push_tiles(const __grid_constant__ TileMap tm, const float* __restrict__ fb, float* __restrict__ host, unsigned int* flag,
             unsigned int epoch, unsigned int* cnt) {
    const int lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    for (int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < tm.n_local_tiles; k += n_warps) {
        const int tile = tm.first_tile + k * tm.tile_stride;
        const int ty = tile / tm.tiles_x;
        int tx = tile - ty * tm.tiles_x;
        if (tm.skew) tx = (tx + tm.skew * ty) % tm.tiles_x;
        push_tile(fb, host, tm.width, tm.height, ty, tx, lane);
    }
    __threadfence_system();
    __syncwarp();
    if (lane == 0 && atomicAdd(cnt, 1u) + 1u == (unsigned)n_warps) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(flag) = epoch;
    }
}

// Item mode, single batch: the warp that finishes the LAST sample of a block adds the block's sample planes in sample order
// (the other samples' planes come from other SMs: read through L2), resolves and stores the pixels -- k_plane_accumulate
// folded into the render kernel, so a multi-sample frame is one launch and its pixels leave (to a peer GPU's frame, over
// NVLink) while the rest of the frame is still being traced.  Out of line: k_packet's register budget.
__device__ __noinline__ void fold_block(const float4* __restrict__ planes, size_t plane_stride, int slot, int batch, int spp, int resolve,
                                        float* __restrict__ o, bool active, int lane) {
    float sr = 0.0f, sg = 0.0f, sb = 0.0f;
    if (active) {
        for (int s = 0; s < batch; ++s) {
            const float4 c = __ldcg(planes + (size_t)s * plane_stride + slot);
            sr = __fadd_rn(sr, c.x); sg = __fadd_rn(sg, c.y); sb = __fadd_rn(sb, c.z);
        }
        if (resolve) {
            const float inv_spp = __fdiv_rn(1.0f, (float)spp);
            sr = resolve1(sr, inv_spp); sg = resolve1(sg, inv_spp); sb = resolve1(sb, inv_spp);
        }
    }
    warp_store_rgb(o, active, sr, sg, sb, lane);
}

// ptxas settles at 48 registers = 5 CTAs of 256 threads per SM, the measured optimum: forcing 40 / 32 registers (6 / 8
// CTAs) spills and is 3 % / 18 % slower, and anything that pushes the kernel to 64 registers (4 CTAs) costs 4-5 % --
// which is why the final pixel store here is the plain one and not warp_store_rgb
#ifndef PACKET_TREELET_MINB
#define PACKET_TREELET_MINB 5
#endif
#ifndef PACKET_MINB
#define PACKET_MINB 0        // 0: ptxas' own choice (48 registers = 5 CTAs of 256 threads per SM, the measured optimum)
#endif
template <bool TRI, bool STATS, bool AOV, bool ITEM, bool TREELET = false>
__global__ void __launch_bounds__(kPacketThreads, TREELET ? PACKET_TREELET_MINB : PACKET_MINB)
k_packet(const __grid_constant__ SceneView sc, const float4* __restrict__ cam_prims, const __grid_constant__ CameraBlock cam,
         const __grid_constant__ TileMap tm, int n_work, int spp, uint32_t k0, uint32_t k1, uint32_t sample_offset,
         int resolve, float* __restrict__ d_out, int32_t* __restrict__ d_prim, float* __restrict__ d_t,
         unsigned int* counter, unsigned long long* d_stats, const __grid_constant__ BandSignal band,
         const __grid_constant__ ChunkSchedule sched, unsigned long long* d_block_times, float4* __restrict__ planes,
         int plane_batch, int plane_sample0, unsigned int* fold_cnt) {
    // planes != nullptr: ITEM MODE for multi-sample frames -- the work item is (block, sample) instead of a block
    // with a sample loop inside: item = block * plane_batch + sb, sample plane_sample0 + sb, radiance written to
    // planes[sb][block * 32 + lane]; k_plane_accumulate then adds the planes in sample order (same bits as the
    // loop) and resolves.  A frame of few blocks and many samples (one rank's tiles of a multi-GPU frame) thus
    // splits into enough items to balance; the item ORDER is described where items are decoded below.
    __shared__ uint2 s_stack[kPacketThreads / 32][kStackDepth];
    __shared__ unsigned s_word;
    extern __shared__ float4 s_tree[];                         // TREELET: the top levels of the tree (SceneView::treelet)
    const int lane = threadIdx.x & 31;
    uint2* stack = s_stack[threadIdx.x >> 5];
    if (TREELET) {
        for (int k = threadIdx.x; k < 2 * sc.treelet_two_t; k += kPacketThreads) s_tree[k] = __ldg(sc.treelet + k);
    }
    constexpr bool item_mode = ITEM && !AOV;
    const int n_items = item_mode ? n_work * plane_batch : n_work;
    const int n_chunks = (n_items + kChunk - 1) / kChunk;
    if (threadIdx.x == 0) s_word = take_chunk(counter, n_chunks, sched.order);
    __syncthreads();
    const double inv_w = __ddiv_rn(1.0, (double)tm.width), inv_h = __ddiv_rn(1.0, (double)tm.height);
    const float inv_spp = __fdiv_rn(1.0f, (float)spp);
    Counters cnt = {0, 0, 0};
    unsigned long long rays = 0;
    for (;;) {
        const int item = chunk_next_block(&s_word, counter, n_chunks, sched.order, lane);
        if (item < 0) break;
        if (item >= n_items) continue;
#ifndef B200RT_ITEM_SAMPLES_ADJACENT
        // Item order: a chunk (8 consecutive items = what the 8 warps of a CTA take together) is 8 DIFFERENT neighbouring blocks at
        // one sample, the next chunk the same blocks at the next sample.  With the samples of ONE block side by side (the first
        // form, -DB200RT_ITEM_SAMPLES_ADJACENT) the 8 warps walk the same nodes at the same time and stall TOGETHER on every line
        // the first of them misses (long-scoreboard 5.1 -> 7.7 cycles per issue, ncu): one rank's share of an 8-GPU frame 0.385 ->
        // 0.348 ms, the 8-spp frame on one GPU 2.83 -> 2.51 ms (profiles/r02z_exp_item_order.txt).  Same planes, same pixels.
        int w = item, item_sb = 0;
        if (item_mode) {
            const int per_group = kChunk * plane_batch, g = item / per_group, r = item - g * per_group;
            const int gs = min(kChunk, n_work - g * kChunk);
            w = g * kChunk + r % gs; item_sb = r / gs;
        }
#else
        const int w = item_mode ? item / plane_batch : item;
        const int item_sb = item_mode ? item - w * plane_batch : 0;
#endif
        PixelWork p = decode_work(tm, w, lane);
        const uint32_t pixel = (uint32_t)(p.j * tm.width + p.i);
        float sr = 0.0f, sg = 0.0f, sb = 0.0f;
        int work = 0;
        unsigned long long t_start = 0;
        if (STATS && d_block_times) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
        const int s_begin = item_mode ? plane_sample0 + item_sb : 0, s_end = item_mode ? s_begin + 1 : spp;
        for (int s = s_begin; s < s_end; ++s) {
            float jx = 0.5f, jy = 0.5f;
            if (!AOV) {
                uint4 ctl = philox4x32_10(pixel, sample_offset + (uint32_t)s, 0u, 0u, k0, k1);
                jx = u01(ctl.x); jy = u01(ctl.y);
            }
            Ray r = camera_ray(cam, p.i, p.j, jx, jy, inv_w, inv_h);
            Hit h;
            packet_intersect<TRI, STATS, TREELET>(sc, cam_prims, r, p.active, lane, stack, h, cnt, work, s_tree);
            if (STATS && p.active) { rays += 1; cnt.segments += 1; }
            if (AOV) {
                if (p.active) { d_prim[p.out_index] = h.prim; d_t[p.out_index] = h.prim >= 0 ? h.t : 0.0f; }
            } else {
                float cr = 0.0f, cg = 0.0f, cb = 0.0f;
                if (h.prim < 0) {
                    cr = __fmaf_rn(1.0f, sc.bg_r, cr); cg = __fmaf_rn(1.0f, sc.bg_g, cg); cb = __fmaf_rn(1.0f, sc.bg_b, cb);
                } else {
                    float4 m1 = __ldg(sc.mats + 2 * (size_t)material_row<TRI>(sc, h) + 1);
                    cr = __fmaf_rn(1.0f, m1.y, cr); cg = __fmaf_rn(1.0f, m1.z, cg); cb = __fmaf_rn(1.0f, m1.w, cb);
                }
                if (item_mode) { if (p.active) planes[(size_t)item_sb * ((size_t)n_work * 32) + (size_t)w * 32 + lane] = make_float4(cr, cg, cb, 0.0f); }
                else { sr = __fadd_rn(sr, cr); sg = __fadd_rn(sg, cg); sb = __fadd_rn(sb, cb); }
            }
        }
        if (!AOV && !item_mode && p.active) {              // (plain stores: the warp transpose costs this kernel registers)
            float* o = d_out + 3 * (size_t)p.out_index;
            if (resolve) { o[0] = resolve1(sr, inv_spp); o[1] = resolve1(sg, inv_spp); o[2] = resolve1(sb, inv_spp); }
            else { o[0] = sr; o[1] = sg; o[2] = sb; }
        }
        if (item_mode && fold_cnt != nullptr) {              // single batch: the block's last sample folds the planes into the pixels
            __threadfence();                                 // release: this sample's plane entries before the count
            __syncwarp();
            unsigned last = 0u;
            if (lane == 0) last = atomicAdd(fold_cnt + w, 1u) + 1u == (unsigned)plane_batch ? 1u : 0u;
            if (__shfl_sync(0xffffffffu, last, 0)) {
                __threadfence();                             // acquire: the other samples' plane entries
                fold_block(planes, (size_t)n_work * 32, w * 32 + lane, plane_batch, spp, resolve, d_out + 3 * (size_t)p.out_index, p.active, lane);
            }
        }
        if (sched.order != nullptr && lane == 0) {
            atomicAdd(sched.cost_sum + item / kChunk, (unsigned)work);
            atomicMax(sched.cost_max + item / kChunk, (unsigned)work);
        }
        if (STATS && d_block_times && lane == 0) {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            d_block_times[2 * (size_t)w] = t_start; d_block_times[2 * (size_t)w + 1] = t_end;
        }
        if (!AOV && !item_mode && band.cnt != nullptr) {     // see BandSignal
            __threadfence();                                 // release: this block's pixels before the count
            __syncwarp();
            if (band.host_fb == nullptr) {
                if (lane == 0) {
                    const int b = region_of_tile(band, w >> 5);
                    if (atomicAdd(band.cnt + b, 1u) + 1u == (unsigned)region_blocks(band, b)) { __threadfence_system(); band.flags[b] = 1u; }
                }
            } else {                                         // tile push: regions are tiles (32 blocks each)
                const int tile = w >> 5;
                unsigned last = 0u;
                if (lane == 0) last = atomicAdd(band.cnt + tile, 1u) + 1u == 32u ? 1u : 0u;
                if (__shfl_sync(0xffffffffu, last, 0)) {
                    __threadfence();                         // acquire: the other 31 blocks' pixels
                    push_tile(d_out, band.host_fb, tm.width, tm.height, tile / band.tiles_x, tile % band.tiles_x, lane);
                    if (band.push_times != nullptr && lane == 0) {
                        unsigned long long t;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                        band.push_times[tile] = t;
                    }
                }
            }
        }
    }
    if (STATS) flush_stats(d_stats, rays, cnt);
}

// Item mode of k_packet, second half: per pixel, add the batch's sample planes IN SAMPLE ORDER to the running sum
// (kept raw in d_out between batches), resolve after the last batch.  d_out may be a peer GPU's frame.
__global__ void __launch_bounds__(256)
k_plane_accumulate(const __grid_constant__ TileMap tm, int n_tasks, int batch, int sample0, int spp, int resolve, int last,
                   const float4* __restrict__ planes, float* __restrict__ d_out) {
    const float inv_spp = __fdiv_rn(1.0f, (float)spp);
    if (tm.tile_w == 32 && tm.tile_h == 32 && !tm.compact && (tm.width & 3) == 0) {
        // frame layout, 32x32 tiles: one warp per TILE ROW (lane = pixel of the row), the row's 96 floats staged in shared memory
        // and stored as 24 x 16 bytes = 384 contiguous, 128-byte aligned bytes -- full lines instead of the 96-byte runs of an
        // 8x4 block's rows.  That matters when the frame is another GPU's (peer stores over NVLink: a rank's second pass took
        // 30 us longer than the display rank's with the block-wise stores).
        __shared__ float s_row[8][96];
        const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
        const int n_rows = tm.n_local_tiles * 32, n_warps = gridDim.x * 8;
        for (int row = blockIdx.x * 8 + wib; row < n_rows; row += n_warps) {
            const int kt = row >> 5, y = row & 31;
            const int tile = tm.first_tile + kt * tm.tile_stride;
            const int ty = tile / tm.tiles_x;
            int tx = tile - ty * tm.tiles_x;
            if (tm.skew) tx = (tx + tm.skew * ty) % tm.tiles_x;
            const int i = tx * 32 + lane, j = ty * 32 + y;
            if (j >= tm.height) continue;                                // warp-uniform
            const bool active = i < tm.width;
            const int k = ((kt * 32 + (y >> 2) * 4 + (lane >> 3)) << 5) + ((y & 3) << 3) + (lane & 7);
            float* o = d_out + 3 * ((size_t)j * tm.width + i);
            float sr = 0.0f, sg = 0.0f, sb = 0.0f;
            if (active) {
                if (sample0 > 0) { sr = o[0]; sg = o[1]; sb = o[2]; }
                for (int s = 0; s < batch; ++s) {
                    const float4 c = planes[(size_t)s * n_tasks + k];
                    sr = __fadd_rn(sr, c.x); sg = __fadd_rn(sg, c.y); sb = __fadd_rn(sb, c.z);
                }
                if (last && resolve) { sr = resolve1(sr, inv_spp); sg = resolve1(sg, inv_spp); sb = resolve1(sb, inv_spp); }
            }
            if (tx * 32 + 32 <= tm.width) {                             // whole row inside the frame (warp-uniform)
                s_row[wib][3 * lane] = sr; s_row[wib][3 * lane + 1] = sg; s_row[wib][3 * lane + 2] = sb;
                __syncwarp();
                if (lane < 24) reinterpret_cast<float4*>(o - 3 * lane)[lane] = reinterpret_cast<const float4*>(s_row[wib])[lane];
                __syncwarp();
            } else if (active) { o[0] = sr; o[1] = sg; o[2] = sb; }
        }
        return;
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_tasks; k += gridDim.x * blockDim.x) {   // n_tasks % 32 == 0
        PixelWork p = decode_work(tm, k >> 5, k & 31);
        float* o = d_out + 3 * (size_t)p.out_index;
        float sr = 0.0f, sg = 0.0f, sb = 0.0f;
        if (p.active) {
            if (sample0 > 0) { sr = o[0]; sg = o[1]; sb = o[2]; }
            for (int s = 0; s < batch; ++s) {
                const float4 c = planes[(size_t)s * n_tasks + k];
                sr = __fadd_rn(sr, c.x); sg = __fadd_rn(sg, c.y); sb = __fadd_rn(sb, c.z);
            }
            if (last && resolve) { sr = resolve1(sr, inv_spp); sg = resolve1(sg, inv_spp); sb = resolve1(sb, inv_spp); }
        }
        warp_store_rgb(o, p.active, sr, sg, sb, k & 31);
    }
}

// rt_frame_sync: one thread per process.  Everything this process stored into the shared frame before this kernel
// (stream order) is released system-wide, the process arrives (one atomic on the frame owner's memory, over NVLink
// for the peers) and spins until all `world` processes of this epoch have arrived: target = world * epoch, the
// counter only ever grows, so nothing is reset between frames.  A peer that has not arrived after ~20 s is a lost
// process: the error word is set and the kernel traps, so that this process fails loudly at its next CUDA call
// instead of hanging the GPU or handing out an incomplete frame.
__global__ void k_frame_sync(unsigned long long* words, unsigned long long target) {
    __threadfence_system();
    atomicAdd_system(words, 1ull);
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned long long*>(words) < target) {
        __nanosleep(200);
        if (clock64() - t0 > 40000000000ll) { atomicExch_system(words + 1, 1ull); __trap(); }
    }
    __threadfence_system();
}

__global__ void k_untile(int width, int height, int tile_w, int tile_h, int tiles_x, int n_ranks, int tiles_per_rank,
                         const float* __restrict__ tiles, float* __restrict__ frame) {
    int64_t n = (int64_t)width * height;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(p / width), i = (int)(p - (int64_t)j * width);
        int tx = i / tile_w, ty = j / tile_h;
        int tile = ty * tiles_x + tx;
        int rank = tile % n_ranks, k = tile / n_ranks;
        int64_t src = ((((int64_t)rank * tiles_per_rank + k) * tile_h + (j - ty * tile_h)) * tile_w + (i - tx * tile_w)) * 3;
        frame[3 * p] = tiles[src]; frame[3 * p + 1] = tiles[src + 1]; frame[3 * p + 2] = tiles[src + 2];
    }
}

__global__ void k_resolve(const float* __restrict__ sum, float* __restrict__ out, int64_t n, float inv_spp) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        out[k] = resolve1(sum[k], inv_spp);
}

// Sum of n_planes partial-sum frames IN PLANE ORDER ((p0 + p1) + p2 ...: deterministic), then resolve.
__global__ void k_resolve_planes(const float* __restrict__ planes, int n_planes, int64_t plane_stride, float* __restrict__ out,
                                 int64_t n, float inv_spp) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        float s = planes[k];
        for (int p = 1; p < n_planes; ++p) s = __fadd_rn(s, planes[(int64_t)p * plane_stride + k]);
        out[k] = resolve1(s, inv_spp);
    }
}

// interaction.py:1311-1325 in float32: accum*w_old + batch*w_new, each product rounded.
__global__ void k_accumulate(const float* __restrict__ batch, float* __restrict__ accum, int64_t n, float w_old,
                             float w_new, int first) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        float b = batch[k];
        accum[k] = first ? b : __fadd_rn(__fmul_rn(accum[k], w_old), __fmul_rn(b, w_new));
    }
}

// interaction.py:1435-1439 then gui.py:73: x*e/(1+x*e) -> clip -> *255 -> uint8 (truncation)
__global__ void k_tonemap_u8(const float* __restrict__ accum, uint8_t* __restrict__ out, int64_t n, float exposure) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        float x = __fmul_rn(accum[k], exposure);
        x = __fdiv_rn(x, __fadd_rn(1.0f, x));
        x = fminf(fmaxf(x, 0.0f), 1.0f);
        out[k] = (uint8_t)__fmul_rn(x, 255.0f);
    }
}

}  // namespace

namespace {

template <bool TRI, bool STATS, bool AOV>
cudaError_t launch_path(const SceneView& sc, const CameraBlock& cam, const TileMap& tm, int spp, int max_depth,
                        int integrator, uint64_t seed, uint32_t sample_offset, int resolve, float* d_out,
                        int32_t* d_prim, float* d_t, const LaunchCfg& cfg) {
    int n_work = work_items(tm);
    int grid = resident_grid(k_path<TRI, STATS, AOV>, cfg.sm_count);
    int need = (n_work + (kThreads / 32) - 1) / (kThreads / 32);
    if (grid > need) grid = need;
    k_path<TRI, STATS, AOV><<<grid, kThreads, 0, cfg.stream>>>(
        sc, cam, tm, n_work * 32, spp, max_depth, integrator, (uint32_t)seed, (uint32_t)(seed >> 32), sample_offset,
        resolve, cfg.refill_below, cfg.leaf_vote, d_out, d_prim, d_t, cfg.d_work_counter, cfg.d_stats);
    return cudaGetLastError();
}

template <bool TRI, bool STATS, bool AOV>
cudaError_t launch_packet(const SceneView& sc, const CameraBlock& cam, const TileMap& tm, int spp, uint64_t seed,
                          uint32_t sample_offset, int resolve, float* d_out, int32_t* d_prim, float* d_t,
                          const LaunchCfg& cfg) {
    int n_work = work_items(tm);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_packet<TRI, STATS, AOV, false>, kPacketThreads, 0);
    int grid = cfg.sm_count * (per_sm < 1 ? 1 : per_sm);
    const bool item_mode = !AOV && spp > 1 && cfg.d_planes != nullptr && cfg.plane_batch >= 1;
    const int batch_all = item_mode ? (spp < cfg.plane_batch ? spp : cfg.plane_batch) : 1;
    int need = (n_work * batch_all + kChunk - 1) / kChunk;
    if (grid > need) grid = need;
    if (cfg.sched.order != nullptr && cfg.sched.reorder_frames > 0) {
        k_chunk_order<<<1, 1024, 0, cfg.stream>>>(cfg.sched.cost_sum, cfg.sched.cost_max, cfg.sched.order, need,
                                                  grid * (kPacketThreads / 32), cfg.sched.reorder_frames, cfg.band);
    }
    if (!item_mode) {
        if (!STATS && sc.treelet != nullptr && sc.treelet_two_t > 0) {        // top treelet staged in shared memory (option "treelet")
            const size_t smem = (size_t)sc.treelet_two_t * 32;
            auto kern = k_packet<TRI, false, AOV, false, true>;
            if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            int per = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, kPacketThreads, smem);
            int g2 = cfg.sm_count * (per < 1 ? 1 : per);
            if (g2 > need) g2 = need;
            kern<<<g2, kPacketThreads, smem, cfg.stream>>>(
                sc, cfg.d_cam_prims, cam, tm, n_work, spp, (uint32_t)seed, (uint32_t)(seed >> 32), sample_offset, resolve, d_out,
                d_prim, d_t, cfg.d_work_counter, nullptr, cfg.band, cfg.sched, cfg.d_block_times, nullptr, 1, 0, nullptr);
            return cudaGetLastError();
        }
        k_packet<TRI, STATS, AOV, false><<<grid, kPacketThreads, 0, cfg.stream>>>(
            sc, cfg.d_cam_prims, cam, tm, n_work, spp, (uint32_t)seed, (uint32_t)(seed >> 32), sample_offset, resolve, d_out,
            d_prim, d_t, cfg.d_work_counter, cfg.d_stats, cfg.band, cfg.sched, cfg.d_block_times, nullptr, 1, 0, nullptr);
        return cudaGetLastError();
    }
    // multi-sample frame: (block, sample) items, batches of <= plane_batch samples (the schedule is attached by the
    // caller only when one batch covers all samples)
    if (batch_all >= spp && cfg.d_fold_cnt != nullptr) {           // one batch: planes folded by the kernel itself (fold_block)
        cudaError_t e = cudaMemsetAsync(cfg.d_fold_cnt, 0, (size_t)n_work * sizeof(unsigned int), cfg.stream);
        if (e != cudaSuccess) return e;
        k_packet<TRI, STATS, false, true><<<grid, kPacketThreads, 0, cfg.stream>>>(
            sc, cfg.d_cam_prims, cam, tm, n_work, spp, (uint32_t)seed, (uint32_t)(seed >> 32), sample_offset, resolve, d_out,
            d_prim, d_t, cfg.d_work_counter, cfg.d_stats, cfg.band, cfg.sched, cfg.d_block_times, cfg.d_planes, spp, 0, cfg.d_fold_cnt);
        return cudaGetLastError();
    }
    for (int s0 = 0; s0 < spp; s0 += batch_all) {
        const int batch = spp - s0 < batch_all ? spp - s0 : batch_all;
        if (s0 > 0) {
            cudaError_t e = cudaMemsetAsync(cfg.d_work_counter, 0, sizeof(unsigned int), cfg.stream);
            if (e != cudaSuccess) return e;
        }
        k_packet<TRI, STATS, false, true><<<grid, kPacketThreads, 0, cfg.stream>>>(
            sc, cfg.d_cam_prims, cam, tm, n_work, spp, (uint32_t)seed, (uint32_t)(seed >> 32), sample_offset, resolve, d_out,
            d_prim, d_t, cfg.d_work_counter, cfg.d_stats, cfg.band, cfg.sched, cfg.d_block_times, cfg.d_planes, batch, s0, nullptr);
        const int n_tasks = n_work * 32;
        int ag = (n_tasks + 255) / 256, acap = cfg.sm_count * 8;
        k_plane_accumulate<<<ag > acap ? acap : ag, 256, 0, cfg.stream>>>(tm, n_tasks, batch, s0, spp, resolve, s0 + batch >= spp ? 1 : 0,
                                                                           cfg.d_planes, d_out);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_cam_tris(const SceneView& sc, const CameraBlock& cam, const LaunchCfg& cfg) {
    if (sc.n_prims <= 0) return cudaSuccess;
    int64_t g = ((int64_t)sc.n_prims + 255) / 256;
    int cap = cfg.sm_count * 8;
    k_cam_tris<<<(int)(g > cap ? cap : g), 256, 0, cfg.stream>>>(sc.prims, cfg.d_cam_prims, sc.n_prims, cam.px, cam.py, cam.pz);
    return cudaGetLastError();
}

int packet_chunks(const TileMap& tm, int items_per_block) { return (work_items(tm) * items_per_block + kChunk - 1) / kChunk; }

cudaError_t launch_trace_primary(const SceneView& sc, bool is_tri, const CameraBlock& cam, const TileMap& tm,
                                 int32_t* d_prim, float* d_t, const LaunchCfg& cfg) {
    int n_work = work_items(tm);
    if (n_work == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(cfg.d_work_counter, 0, sizeof(unsigned int), cfg.stream);
    if (e != cudaSuccess) return e;
    if (is_tri && !cfg.cam_table_valid && (e = launch_cam_tris(sc, cam, cfg)) != cudaSuccess) return e;   // camera rays use the table
    bool st = cfg.d_stats != nullptr;
    if (cfg.variant == 3) {
        if (is_tri) return st ? launch_packet<true, true, true>(sc, cam, tm, 1, 0, 0, 0, nullptr, d_prim, d_t, cfg)
                              : launch_packet<true, false, true>(sc, cam, tm, 1, 0, 0, 0, nullptr, d_prim, d_t, cfg);
        return st ? launch_packet<false, true, true>(sc, cam, tm, 1, 0, 0, 0, nullptr, d_prim, d_t, cfg)
                  : launch_packet<false, false, true>(sc, cam, tm, 1, 0, 0, 0, nullptr, d_prim, d_t, cfg);
    }
    if (cfg.variant == 0) {
        if (is_tri) return st ? launch_path<true, true, true>(sc, cam, tm, 1, 1, 0, 0, 0, 0, nullptr, d_prim, d_t, cfg)
                              : launch_path<true, false, true>(sc, cam, tm, 1, 1, 0, 0, 0, 0, nullptr, d_prim, d_t, cfg);
        return st ? launch_path<false, true, true>(sc, cam, tm, 1, 1, 0, 0, 0, 0, nullptr, d_prim, d_t, cfg)
                  : launch_path<false, false, true>(sc, cam, tm, 1, 1, 0, 0, 0, 0, nullptr, d_prim, d_t, cfg);
    }
#define LAUNCH(T, S)                                                                                         \
    {                                                                                                        \
        int grid = resident_grid(k_trace_primary<T, S>, cfg.sm_count);                                       \
        int need = (n_work + (kThreads / 32) - 1) / (kThreads / 32);                                         \
        if (grid > need) grid = need;                                                                        \
        k_trace_primary<T, S><<<grid, kThreads, 0, cfg.stream>>>(sc, cam, tm, n_work, d_prim, d_t,           \
                                                                 cfg.d_work_counter, cfg.d_stats);           \
    }
    if (is_tri) { if (st) LAUNCH(true, true) else LAUNCH(true, false) }
    else { if (st) LAUNCH(false, true) else LAUNCH(false, false) }
#undef LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_trace_rays(const SceneView& sc, bool is_tri, const float* d_org, const float* d_dir, int64_t n,
                              int32_t* d_prim, float* d_t, const LaunchCfg& cfg) {
    if (n == 0) return cudaSuccess;
    bool st = cfg.d_stats != nullptr;
    int64_t need = (n + kThreads - 1) / kThreads;
#define LAUNCH(T, S)                                                                                         \
    {                                                                                                        \
        int grid = resident_grid(k_trace_rays<T, S>, cfg.sm_count);                                          \
        if (grid > need) grid = (int)need;                                                                   \
        k_trace_rays<T, S><<<grid, kThreads, 0, cfg.stream>>>(sc, d_org, d_dir, n, d_prim, d_t, cfg.d_stats); \
    }
    if (is_tri) { if (st) LAUNCH(true, true) else LAUNCH(true, false) }
    else { if (st) LAUNCH(false, true) else LAUNCH(false, false) }
#undef LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_render(const SceneView& sc, bool is_tri, const CameraBlock& cam, const TileMap& tm, int spp,
                          int max_depth, int integrator, uint64_t seed, uint32_t sample_offset, int resolve,
                          float* d_out, const LaunchCfg& cfg) {
    int n_work = work_items(tm);
    if (n_work == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(cfg.d_work_counter, 0, sizeof(unsigned int), cfg.stream);
    if (e != cudaSuccess) return e;
    if (cfg.variant == 5)                                           // the tiny-scene kernel builds its own table in shared memory
        return launch_tiny(sc, is_tri, sc.n_mats, cam, tm, spp, max_depth, integrator, seed, sample_offset, resolve, d_out, cfg);
    if (is_tri && !cfg.cam_table_valid && (e = launch_cam_tris(sc, cam, cfg)) != cudaSuccess) return e;   // bounce 0 uses the table
    bool st = cfg.d_stats != nullptr;
    if (cfg.variant == 3 && max_depth == 1) {
        if (is_tri) return st ? launch_packet<true, true, false>(sc, cam, tm, spp, seed, sample_offset, resolve, d_out, nullptr, nullptr, cfg)
                              : launch_packet<true, false, false>(sc, cam, tm, spp, seed, sample_offset, resolve, d_out, nullptr, nullptr, cfg);
        return st ? launch_packet<false, true, false>(sc, cam, tm, spp, seed, sample_offset, resolve, d_out, nullptr, nullptr, cfg)
                  : launch_packet<false, false, false>(sc, cam, tm, spp, seed, sample_offset, resolve, d_out, nullptr, nullptr, cfg);
    }
    if (cfg.variant == 0 || cfg.variant == 3) {
        if (is_tri) return st ? launch_path<true, true, false>(sc, cam, tm, spp, max_depth, integrator, seed, sample_offset, resolve, d_out, nullptr, nullptr, cfg)
                              : launch_path<true, false, false>(sc, cam, tm, spp, max_depth, integrator, seed, sample_offset, resolve, d_out, nullptr, nullptr, cfg);
        return st ? launch_path<false, true, false>(sc, cam, tm, spp, max_depth, integrator, seed, sample_offset, resolve, d_out, nullptr, nullptr, cfg)
                  : launch_path<false, false, false>(sc, cam, tm, spp, max_depth, integrator, seed, sample_offset, resolve, d_out, nullptr, nullptr, cfg);
    }
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#define LAUNCH(T, S)                                                                                         \
    {                                                                                                        \
        int grid = resident_grid(k_render<T, S>, cfg.sm_count);                                              \
        int need = (n_work + (kThreads / 32) - 1) / (kThreads / 32);                                         \
        if (grid > need) grid = need;                                                                        \
        k_render<T, S><<<grid, kThreads, 0, cfg.stream>>>(sc, cam, tm, n_work, spp, max_depth, integrator,   \
                                                          k0, k1, sample_offset, resolve, d_out,             \
                                                          cfg.d_work_counter, cfg.d_stats);                  \
    }
    if (is_tri) { if (st) LAUNCH(true, true) else LAUNCH(true, false) }
    else { if (st) LAUNCH(false, true) else LAUNCH(false, false) }
#undef LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_frame_sync(unsigned long long* words, unsigned long long target, cudaStream_t stream) {
    k_frame_sync<<<1, 1, 0, stream>>>(words, target);
    return cudaGetLastError();
}

cudaError_t launch_untile(int width, int height, int tile_w, int tile_h, int n_ranks, const float* d_tiles,
                          float* d_frame, cudaStream_t stream) {
    int tiles_x = (width + tile_w - 1) / tile_w, tiles_y = (height + tile_h - 1) / tile_h;
    int n_tiles = tiles_x * tiles_y;
    int tiles_per_rank = (n_tiles + n_ranks - 1) / n_ranks;
    int64_t n = (int64_t)width * height;
    if (n == 0) return cudaSuccess;
    k_untile<<<elementwise_grid(n), 256, 0, stream>>>(width, height, tile_w, tile_h, tiles_x, n_ranks, tiles_per_rank,
                                                      d_tiles, d_frame);
    return cudaGetLastError();
}

cudaError_t launch_resolve(const float* d_sum, float* d_out, int64_t n, int spp_total, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_resolve<<<elementwise_grid(n), 256, 0, stream>>>(d_sum, d_out, n, 1.0f / (float)spp_total);
    return cudaGetLastError();
}

cudaError_t launch_resolve_planes(const float* d_planes, int n_planes, int64_t plane_stride, float* d_out, int64_t n,
                                  int spp_total, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_resolve_planes<<<elementwise_grid(n), 256, 0, stream>>>(d_planes, n_planes, plane_stride, d_out, n, 1.0f / (float)spp_total);
    return cudaGetLastError();
}

cudaError_t launch_accumulate(const float* d_batch, float* d_accum, int64_t n, int n_old, int n_batch,
                              cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    double total = (double)n_old + (double)n_batch;
    float w_old = (float)((double)n_old / total), w_new = (float)((double)n_batch / total);
    k_accumulate<<<elementwise_grid(n), 256, 0, stream>>>(d_batch, d_accum, n, w_old, w_new, n_old == 0 ? 1 : 0);
    return cudaGetLastError();
}

// Option "qnodes": compressed copy of the tree for the incoherent bounces (rt_device.cuh pair_hit_q).  One thread per sibling
// pair; the grid is the root box enlarged by 1/256 of its extent below and ~1 % above (so that no plane of the tree comes
// near the ends of the 15-bit range), every thread derives it from node 0 with the same float operations and thread 0
// publishes it.  lo planes are rounded DOWN and hi planes UP, then moved one more cell outward (covers the rounding of the
// double division below); 16-bit plane = 0x8000 | cell.
// out_of_range: a plane that does not fit the grid (a caller-supplied tree whose child sticks out of the root box, NaN boxes):
// clamping it would SHRINK the box, so the host drops the compressed copy for such a tree.
__device__ __forceinline__ unsigned quantise_axis(float lo, float hi, float g0, float e, bool& out_of_range) {
    double ul = floor(((double)lo - (double)g0) / (double)e * 32768.0) - 1.0;
    double uh = ceil(((double)hi - (double)g0) / (double)e * 32768.0) + 1.0;
    if (!(ul >= 0.0 && uh <= 32767.0)) out_of_range = true;                       // (NaN compares false)
    ul = fmin(fmax(ul, 0.0), 32767.0); uh = fmin(fmax(uh, 0.0), 32767.0);          // NaN -> 0
    return (0x8000u | (unsigned)(int)ul) | ((0x8000u | (unsigned)(int)uh) << 16);
}
// quality[0] / quality[1] (zeroed by the launcher): number of LEAF boxes / sum over them of (half-area as the grid renders the
// box) / (half-area as stored), each ratio capped at 1000 -- the host keeps the compressed copy only while the mean ratio stays
// close to 1 (a scene whose detail is finer than a grid cell would enter many more leaves).  quality[2] != 0: some plane did
// not fit the grid.
__device__ __forceinline__ double qplane_value(unsigned s16, float g0, float e) { return (double)g0 + (double)(s16 & 0x7fffu) * (1.0 / 32768.0) * (double)e; }
__global__ void __launch_bounds__(256)
k_quantize_pairs(const float4* __restrict__ nodes, int n_pairs, uint4* __restrict__ out, float* __restrict__ grid, double* __restrict__ quality) {
    float rlo[3], rhi[3], g0[3], e[3];
    int rcode;
    node_read(nodes, 0, rlo, rhi, rcode);
    for (int c = 0; c < 3; ++c) {
        const float ext = __fsub_rn(rhi[c], rlo[c]);
        g0[c] = __fsub_rn(rlo[c], __fmul_rn(ext, 0x1p-8f));
        e[c] = __fadd_rn(__fmul_rn(ext, 1.015625f), 1e-30f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int c = 0; c < 3; ++c) { grid[c] = g0[c]; grid[3 + c] = e[c]; }
    }
    double area = 0.0, qarea = 0.0;
    bool out_of_range = false;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n_pairs; m += gridDim.x * blockDim.x) {
        float lo[3], hi[3];
        int code[2];
        unsigned w[6];
        for (int side = 0; side < 2; ++side) {
            node_read(nodes, 2 * m + side, lo, hi, code[side]);
            double d[3], q[3];
            for (int c = 0; c < 3; ++c) {
                bool oor = false;
                const unsigned ww = quantise_axis(lo[c], hi[c], g0[c], e[c], oor);
                if (oor && !(m == 0 && side == 1)) out_of_range = true;          // (record 1 is the pad record next to the root)
                w[3 * side + c] = ww;
                d[c] = (double)hi[c] - (double)lo[c];
                q[c] = qplane_value(ww >> 16, g0[c], e[c]) - qplane_value(ww & 0xffffu, g0[c], e[c]);
            }
            if (code[side] <= -2 && d[0] >= 0.0 && d[1] >= 0.0 && d[2] >= 0.0) {      // a leaf (the pad record of pair 0 has code 0)
                const double a0 = d[0] * d[1] + d[1] * d[2] + d[2] * d[0], a1 = q[0] * q[1] + q[1] * q[2] + q[2] * q[0];
                area += 1.0;
                qarea += a0 > 0.0 ? fmin(a1 / a0, 1000.0) : 1000.0;
            }
        }
        out[2 * (size_t)m] = make_uint4(w[0], w[1], w[2], w[3]);
        out[2 * (size_t)m + 1] = make_uint4(w[4], w[5], (unsigned)code[0], (unsigned)code[1]);
    }
    for (int o = 16; o > 0; o >>= 1) { area += __shfl_xor_sync(0xffffffffu, area, o); qarea += __shfl_xor_sync(0xffffffffu, qarea, o); }
    if ((threadIdx.x & 31) == 0 && (area != 0.0 || qarea != 0.0)) { atomicAdd(quality, area); atomicAdd(quality + 1, qarea); }
    if (out_of_range) quality[2] = 1.0;
}
// Option "qnodes" bit 1: the 48-byte triangle records as a 32-byte part (v0|prim, e1|material: one 256-bit load) and a 16-byte part (e2|0)
__global__ void __launch_bounds__(256)
k_split_tris(const float4* __restrict__ prims, int n, float4* __restrict__ tri_a, float4* __restrict__ tri_b) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        tri_a[2 * (size_t)k] = prims[kTriStride * (size_t)k];
        tri_a[2 * (size_t)k + 1] = prims[kTriStride * (size_t)k + 1];
        tri_b[k] = prims[kTriStride * (size_t)k + 2];
    }
}

cudaError_t launch_quantize_pairs(const float4* d_nodes, int n_pairs, uint4* d_qnodes, float* d_qgrid, double* d_quality, cudaStream_t stream) {
    if (n_pairs <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_quality, 0, 3 * sizeof(double), stream);
    if (e != cudaSuccess) return e;
    k_quantize_pairs<<<elementwise_grid(n_pairs), 256, 0, stream>>>(d_nodes, n_pairs, d_qnodes, d_qgrid, d_quality);
    return cudaGetLastError();
}
cudaError_t launch_split_tris(const float4* d_prims, int n, float4* d_tri_a, float4* d_tri_b, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    k_split_tris<<<elementwise_grid(n), 256, 0, stream>>>(d_prims, n, d_tri_a, d_tri_b);
    return cudaGetLastError();
}

cudaError_t launch_build_treelet(const float4* d_nodes, int n_pairs, float4* d_treelet, cudaStream_t stream) {
    if (n_pairs <= 0) return cudaSuccess;
    k_build_treelet<<<(n_pairs + 255) / 256, 256, 0, stream>>>(d_nodes, n_pairs, d_treelet);
    return cudaGetLastError();
}

cudaError_t launch_push_tiles(const TileMap& tm, const float* d_fb, float* d_host, unsigned int* d_flag, unsigned int epoch,
                              unsigned int* d_cnt, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    int grid = (tm.n_local_tiles + 7) / 8;
    if (grid < 1) grid = 1;
    if (grid > 148 * 4) grid = 148 * 4;
    k_push_tiles<<<grid, 256, 0, stream>>>(tm, d_fb, d_host, d_flag, epoch, d_cnt);
    return cudaGetLastError();
}

cudaError_t launch_tonemap_u8(const float* d_accum, uint8_t* d_rgb8, int64_t n, float exposure, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_tonemap_u8<<<elementwise_grid(n), 256, 0, stream>>>(d_accum, d_rgb8, n, exposure);
    return cudaGetLastError();
}

}  // namespace b200rt
