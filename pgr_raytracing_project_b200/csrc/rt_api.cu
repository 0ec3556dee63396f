// rt_api.cu -- the C ABI of libb200rt.so (include/b200rt.h): context, scene upload, BVH
// management, camera, and the launches of the render hot path.  No torch types, no CPU render
// fallback: every tracing entry point runs the sm_100a kernels of rt_kernels.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200rt.h"
#include "rt_bvh.h"
#include "rt_kernels.h"
#include "rt_display.h"
#include "rt_lbvh.h"
#include "rt_refit.h"

using namespace b200rt;

struct rt_ctx {
    int device = 0;
    int sm_count = 148;
    std::recursive_mutex mu;
    std::string err;

    // host copy of the scene (RayTracer::set_scene deep-copies, old/raytracer_core copy.cpp:162-167)
    bool is_tri = false;
    int64_t n = 0;
    std::vector<float> prim_data;      // spheres: n x 4; triangles: n x 9 (as uploaded)
    std::vector<float> mats;           // m x 8
    std::vector<int32_t> mat_id;       // triangles only
    std::vector<int32_t> object_id;
    int m = 0;
    float bg[3] = {0.1f, 0.1f, 0.1f};  // Scene::Scene(), old/raytracer_core copy.cpp:54

    // host BVH
    std::vector<rt_bvh_node> nodes;
    std::vector<int32_t> prim_index;
    bool bvh_valid = false;
    int bvh_depth = 0;
    int64_t n_nodes = 0;                 // node records of the current tree (host vector `nodes` may be a stale mirror)
    float root_extent = 0.0f;            // max |coordinate| of the root box
    int treelet_levels = 0;              // option "treelet": levels of the tree the packet kernel stages in shared memory (0: none)
    float4* d_treelet = nullptr;         // 2^levels - 1 sibling pairs, heap order (k_build_treelet); rebuilt when the tree changes
    int treelet_pairs = 0;
    bool treelet_valid = false;
    // option "qnodes": compressed copies for the incoherent bounces (k_quantize_pairs / k_split_tris); rebuilt when the tree changes
    int qmode = -1;                      // -1 auto (bit 0 when the grid is fine enough for the scene), 0 off; forced: bit 0 = 32-byte sibling pairs,
                                         // bit 1 = split triangle records (measured: no gain on top of bit 0)
    int qmode_used = 0;                  // what the current tree's launches use
    int64_t q_area_pct = 0;              // mean over the leaves of (half-area on the grid / half-area as stored), in % (k_quantize_pairs)
    int q_area_limit = 115;              // auto keeps the compressed pairs up to this
    uint4* d_qnodes = nullptr;
    float* d_qgrid = nullptr;
    float4* d_tri_a = nullptr;
    float4* d_tri_b = nullptr;
    int64_t q_pairs = 0, q_tris = 0;     // capacities
    bool qnodes_valid = false;
    int leaf_size = 4;                   // option "leaf_size": primitives per leaf of builder 0 (4 = the reference's rule)
    int sah_cost = 30;                   // option "sah_cost": builder 2's cost of a traversal step, in tenths of a primitive test
    int builder = 0;                     // option "builder": what set_scene -> render builds with (0 host median split, 1 device LBVH)
    rt_bvh_node* d_nodes_abi = nullptr;  // device-built tree in ABI layout, until the host mirror is asked for (rt_get_bvh)
    bool host_bvh_stale = false;

    // device copies
    float4* d_nodes = nullptr;
    float4* d_prims = nullptr;
    float4* d_cam_prims = nullptr;           // triangles: camera-relative records (k_cam_tris), valid for cam_table_pos
    float cam_table_pos[3] = {0, 0, 0};
    cudaStream_t cam_table_stream = nullptr; // the stream the table was built on (another stream rebuilds: no cross-stream order)
    bool cam_table_ok = false;               // reset by every scene upload; checked against the camera position per launch
    int* d_slot_prim = nullptr;
    float4* d_mats = nullptr;
    bool device_valid = false;

    // camera (Camera::Camera(), old/raytracer_core copy.h:158)
    double pos[3] = {0, 2, 3}, target[3] = {0, 0, -3}, up[3] = {0, 1, 0};
    double fov = 45.0, aspect = 1.333;

    // options
    int integrator = 0, stats = 0, kernel = -1, refill = 8, leaf_vote = 8, tiny_threads = 256, tiny_mode = 0, wf_rays_per_lane = 0;
    bool refill_set = false, leaf_vote_set = false;   // set by the caller: also used by the cooperative-leaf launches (whose own optimum is 24 / 7)
    int kernel_used = -1;                    // variant picked by the most recent tracing launch
    // auto choice for multi-bounce renders of tiny scenes (<= 64 primitives): which of the lock-step megakernel
    // (open scenes, short paths: the reference's default scene) and the wavefront (closed scenes, long paths: a
    // Cornell box) wins depends on the scene, so the first two such renders after a scene upload are timed, one
    // with each (all variants give the same pixels), and the faster one is kept
    int tune_state = 0;                      // 0, 1: candidate to time next; 2: decided
    int tune_pending = -1;                   // candidate whose timing events are in flight
    double tune_pending_samples = 0, tune_rate[2] = {0, 0};
    cudaEvent_t tune_ev0 = nullptr, tune_ev1 = nullptr;

    unsigned int* d_work_counter = nullptr;
    unsigned long long* d_stats = nullptr;   // rays, segments, node_records, prim_tests
    uint64_t launches = 0;

    WaveBuffers wave = {{nullptr, nullptr}, {nullptr, nullptr}, nullptr, nullptr, nullptr, nullptr, 0};
    int wave_depth = 0;                      // counters sized for this max_depth
    // two waves in flight (option "wf_streams" 2): second buffer set, two internal streams, fork / join / accumulate-order events
    WaveBuffers wave2 = {{nullptr, nullptr}, {nullptr, nullptr}, nullptr, nullptr, nullptr, nullptr, 0};
    int wf_streams = 2;
    WavePipe pipe = {nullptr, {nullptr, nullptr}, nullptr, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};

    float* d_fb = nullptr;                   // rt_render_host framebuffer
    size_t fb_floats = 0;
    // rt_render_host copy/compute overlap (BandSignal): streams, per-band counters and host-mapped flags
    cudaStream_t render_stream = nullptr, copy_stream = nullptr, band_stream = nullptr;
    cudaEvent_t band_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned int* d_band_cnt = nullptr;
    unsigned int* h_band_flags = nullptr;    // cudaHostAlloc mapped
    unsigned int* d_band_flags = nullptr;    // device alias of h_band_flags
    int host_path = -1;                      // read-only option "host_path": how the last rt_render_host moved its frame:
                                             // 0 render, then one copy; 1 region flags + DMA copies; 2 tile push; 3 bands (render / copy overlapped)
    int overlap = 2;                         // option "overlap": 0 render then copy, 1 region flags + DMA copies, 2 tile push
    unsigned int* d_tile_cnt = nullptr;      // tile push: one completion counter per 32x32 tile
    int tile_cnt_cap = 0;
    // cost-aware chunk order of the packet kernel (ChunkSchedule): history of the last frame of this tile map
    int* d_chunk_order = nullptr;
    unsigned int* d_chunk_cost = nullptr;
    int chunk_cap = 0;
    long long chunk_key = -1;
    int chunk_frames = 0;                    // frames whose costs are in d_chunk_cost since the last rebuild of the order
    int chunk_age = 0;                       // frames rendered with the current order
    CameraBlock chunk_cam;                   // camera the current order was built for
    int schedule = 1;                        // option "schedule"
    unsigned long long* d_block_times = nullptr;   // debug option "block_times" (device pointer supplied by the caller)
    unsigned int* d_fold_cnt = nullptr;      // item mode, one batch: per-block sample counters (fold_block)
    size_t fold_cap = 0;
    int fold = 0;                            // option "fold": 1 = the packet kernel folds the sample planes itself; 0 (default, measured faster:
                                             // 2.81 vs 2.86 ms at 8 spp, 0.694 vs 0.743 ms at 2 spp on C3) = separate k_plane_accumulate pass
    float4* d_planes = nullptr;              // item mode of the packet kernel: sample planes (grow-only)
    size_t planes_cap = 0;                   // in float4
    void* d_display = nullptr;               // rt_display_u8 scratch (tone-mapped copy, sorted copy, sort workspace)
    size_t display_bytes = 0;
    int32_t* d_pick = nullptr;               // rt_select_object scratch: org3 dir3 | prim | t
    void* d_edit = nullptr;                  // rt_update_geometry scratch: raw primitives + refit workspace (grow-only)
    size_t edit_bytes = 0;
    // Context-owned scratch (work counter, wave buffers, sample planes, chunk schedule, camera table, stats) is shared by
    // every launch.  Launches on ONE stream are ordered by the stream; a launch on a DIFFERENT stream than the previous one
    // first waits for that one's completion event (ScratchOrder), so two streams can never advance each other's counters.
    void* h_stage[2] = {nullptr, nullptr};   // staged_upload: page-locked ring
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    cudaStream_t stage_stream = nullptr;
    cudaEvent_t scratch_ev = nullptr;
    cudaStream_t scratch_stream = nullptr;
    bool scratch_pending = false;
    double built_area = 0.0;                 // tree-quality figure (bvh_area) of the tree as BUILT; 0 = not measured yet
    int refit_limit = 200;                   // option "refit_limit": rebuild when a refit leaves more than this % of built_area (0: never)
    int64_t refits = 0, refit_rebuilds = 0, area_pct = 100;
};

namespace {

std::string g_create_error;

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int fail(rt_ctx* c, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return 1;
}
int cuda_fail(rt_ctx* c, const char* what, cudaError_t e) {
    return fail(c, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return cuda_fail(ctx, #call, e_); \
    } while (0)

void free_device_scene(rt_ctx* c) {
    cudaFree(c->d_nodes); cudaFree(c->d_prims); cudaFree(c->d_cam_prims); cudaFree(c->d_nodes_abi); c->d_nodes_abi = nullptr; cudaFree(c->d_slot_prim); cudaFree(c->d_mats);
    c->d_nodes = c->d_prims = c->d_cam_prims = c->d_mats = nullptr; c->d_slot_prim = nullptr;
    c->device_valid = false; c->cam_table_ok = false; c->treelet_valid = false; c->qnodes_valid = false;
}

// Camera basis exactly as Camera::get_ray builds it (old/raytracer_core copy.h:160-184): forward
// from target, right = forward x world-up (0,1,0) with the (1,0,0) fallback, up = right x forward,
// tan(fov * 3.14159 / 360).  Like the reference, Camera::up is carried but not read.
CameraBlock camera_block(const rt_ctx* c, double aspect) {
    auto norm = [](double* v) {
        double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (l > 0) { v[0] /= l; v[1] /= l; v[2] /= l; }
    };
    double f[3] = {c->target[0] - c->pos[0], c->target[1] - c->pos[1], c->target[2] - c->pos[2]};
    norm(f);
    double r[3] = {f[1] * 0.0 - f[2] * 1.0, f[2] * 0.0 - f[0] * 0.0, f[0] * 1.0 - f[1] * 0.0};
    norm(r);
    if (std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) < 0.001) { r[0] = 1; r[1] = 0; r[2] = 0; }
    double u[3] = {r[1] * f[2] - r[2] * f[1], r[2] * f[0] - r[0] * f[2], r[0] * f[1] - r[1] * f[0]};
    norm(u);
    double tan_fov = std::tan(c->fov * 3.14159 / 360.0);
    CameraBlock b;
    b.px = (float)c->pos[0]; b.py = (float)c->pos[1]; b.pz = (float)c->pos[2]; b.pad_ = 0.0f;
    for (int k = 0; k < 3; ++k) { b.fwd[k] = f[k]; b.right[k] = r[k]; b.up[k] = u[k]; }
    b.sx = aspect * tan_fov;
    b.sy = tan_fov;
    return b;
}

int ensure_bvh(rt_ctx* ctx);

// RAII around every group of launches that uses context-owned scratch on `stream`: on entry, if the previous such group
// ran on another stream, make `stream` wait for it; on exit, record the completion event the next group may have to wait on.
int ensure_treelet(rt_ctx* ctx, cudaStream_t stream);
int ensure_qnodes(rt_ctx* ctx, cudaStream_t stream);
struct ScratchOrder {
    rt_ctx* c; cudaStream_t st;
    ScratchOrder(rt_ctx* ctx, void* stream) : c(ctx), st((cudaStream_t)stream) {
        if (!c->scratch_ev) cudaEventCreateWithFlags(&c->scratch_ev, cudaEventDisableTiming);
        if (c->scratch_pending && c->scratch_stream != st) cudaStreamWaitEvent(st, c->scratch_ev, 0);
        ensure_treelet(c, st);                                 // the staged top treelet follows the tree (option "treelet")
        ensure_qnodes(c, st);                                  // and so do the compressed copies (option "qnodes")
    }
    ~ScratchOrder() {
        if (c->scratch_ev && cudaEventRecord(c->scratch_ev, st) == cudaSuccess) { c->scratch_stream = st; c->scratch_pending = true; }
    }
};

void free_wave_set(WaveBuffers& w) {
    for (int k = 0; k < 2; ++k) { cudaFree(w.ray_o[k]); cudaFree(w.ray_d[k]); w.ray_o[k] = w.ray_d[k] = nullptr; }
    cudaFree(w.hit); cudaFree(w.path_thr); cudaFree(w.path_rad); cudaFree(w.counters);
    w.hit = w.path_thr = w.path_rad = nullptr; w.counters = nullptr;
    w.capacity = 0;
}
void free_wave(rt_ctx* c) {
    free_wave_set(c->wave); free_wave_set(c->wave2);
    c->wave_depth = 0;
}
int alloc_wave_set(rt_ctx* ctx, WaveBuffers& w, int64_t cap, int depth) {
    size_t b = (size_t)cap * sizeof(float4);
    for (int k = 0; k < 2; ++k) { CK(cudaMalloc(&w.ray_o[k], b)); CK(cudaMalloc(&w.ray_d[k], b)); }
    CK(cudaMalloc(&w.hit, b)); CK(cudaMalloc(&w.path_thr, b)); CK(cudaMalloc(&w.path_rad, b));
    CK(cudaMalloc(&w.counters, sizeof(unsigned int) * 2 * (depth + 2)));
    w.capacity = (int)cap;
    return 0;
}
// the pipe of a wavefront launch: nullptr members = one wave at a time
const WavePipe* wave_pipe(rt_ctx* ctx) {
    WavePipe& p = ctx->pipe;
    p.wave2 = nullptr;
    if (ctx->wf_streams < 2 || ctx->wave2.capacity == 0) return &p;
    if (!p.streams[0]) {
        for (int k = 0; k < 2; ++k) {
            if (cudaStreamCreateWithFlags(&p.streams[k], cudaStreamNonBlocking) != cudaSuccess) return &p;
            cudaEventCreateWithFlags(&p.join[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&p.acc[k], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming);
    }
    p.counters[0] = ctx->d_work_counter; p.counters[1] = ctx->d_work_counter + 4;
    p.wave2 = &ctx->wave2;
    return &p;
}

// Wavefront buffers: grow-only, sized for min(paths of the call, 8 Mi) paths per wave.
int ensure_wave(rt_ctx* ctx, int64_t n_tasks, int spp, int max_depth) {
    int64_t want = n_tasks * (int64_t)spp;
    const int64_t kCap = (int64_t)8 << 20;
    if (want > kCap) want = n_tasks > kCap ? kCap : (kCap / n_tasks) * n_tasks;
    if (want < 1024) want = 1024;
    const bool want2 = ctx->wf_streams >= 2;
    if (want <= ctx->wave.capacity && max_depth <= ctx->wave_depth && (!want2 || ctx->wave2.capacity == ctx->wave.capacity)) return 0;
    int64_t cap = want > ctx->wave.capacity ? want : ctx->wave.capacity;
    int depth = max_depth > ctx->wave_depth ? max_depth : ctx->wave_depth;
    cudaDeviceSynchronize();
    free_wave(ctx);
    if (int rc = alloc_wave_set(ctx, ctx->wave, cap, depth)) return rc;
    if (want2) { if (int rc = alloc_wave_set(ctx, ctx->wave2, cap, depth)) return rc; }
    ctx->wave_depth = depth;
    return 0;
}

// Host array -> device through a ring of two page-locked staging buffers: the host cores copy chunk k + 1 into a staging
// buffer (and, if asked, into the context's own host copy of the scene) while the DMA engine sends chunk k.  A pageable
// cudaMemcpy of the 36 MB of a 1M-triangle edit took 3.9 ms; this is bounded by the parallel host copy (~1 ms).
int staged_upload(rt_ctx* ctx, const void* src, size_t bytes, void* d_dst, void* host_copy) {
    constexpr size_t kChunk = (size_t)4 << 20;
    if (!ctx->h_stage[0]) {
        for (int k = 0; k < 2; ++k) {
            CK(cudaHostAlloc(&ctx->h_stage[k], kChunk, cudaHostAllocDefault));
            CK(cudaEventCreateWithFlags(&ctx->stage_ev[k], cudaEventDisableTiming));
        }
        CK(cudaStreamCreateWithFlags(&ctx->stage_stream, cudaStreamNonBlocking));
    }
    const char* s = static_cast<const char*>(src);
    int k = 0;
    for (size_t off = 0; off < bytes; off += kChunk, k ^= 1) {
        const size_t len = std::min(kChunk, bytes - off);
        CK(cudaEventSynchronize(ctx->stage_ev[k]));                       // the DMA that last read this buffer is done
        char* stage = static_cast<char*>(ctx->h_stage[k]);
        char* keep = host_copy ? static_cast<char*>(host_copy) + off : nullptr;
        const long pieces = (long)((len + 65535) / 65536);
#pragma omp parallel for schedule(static)
        for (long p = 0; p < pieces; ++p) {
            const size_t o = (size_t)p * 65536, l = std::min((size_t)65536, len - o);
            std::memcpy(stage + o, s + off + o, l);
            if (keep) std::memcpy(keep + o, s + off + o, l);
        }
        CK(cudaMemcpyAsync(static_cast<char*>(d_dst) + off, stage, len, cudaMemcpyHostToDevice, ctx->stage_stream));
        CK(cudaEventRecord(ctx->stage_ev[k], ctx->stage_stream));
    }
    CK(cudaStreamSynchronize(ctx->stage_stream));
    return 0;
}

int64_t task_count(const TileMap& tm) { return (int64_t)tm.n_local_tiles * (tm.tile_w >> 3) * (tm.tile_h >> 2) * 32; }

// Host scene + BVH -> device arrays in leaf order.
int ensure_device(rt_ctx* ctx) {
    if (ctx->device_valid) return 0;
    if (int rc = ensure_bvh(ctx)) return rc;
    if (ctx->device_valid) return 0;                     // the implicit build was the DEVICE builder (option "builder" 1): its arrays are the scene
    free_device_scene(ctx);
    const int64_t n = ctx->n;
    const int64_t n_nodes = (int64_t)ctx->nodes.size();
    if (n > 0) {
        const int per = ctx->is_tri ? kTriStride : 1;
        std::vector<float4> prims((size_t)n * per);
        for (int64_t slot = 0; slot < n; ++slot) {
            const int32_t p = ctx->prim_index[slot];
            if (ctx->is_tri) {
                const float* v = &ctx->prim_data[9 * (size_t)p];
                int32_t mid = ctx->mat_id[p];
                float pw, mw;
                std::memcpy(&pw, &p, 4); std::memcpy(&mw, &mid, 4);
                prims[kTriStride * slot + 0] = make_float4(v[0], v[1], v[2], pw);
                prims[kTriStride * slot + 1] = make_float4(v[3] - v[0], v[4] - v[1], v[5] - v[2], mw);   // e1 = v1 - v0
                prims[kTriStride * slot + 2] = make_float4(v[6] - v[0], v[7] - v[1], v[8] - v[2], 0.0f); // e2 = v2 - v0
            } else {
                const float* s = &ctx->prim_data[4 * (size_t)p];
                prims[slot] = make_float4(s[0], s[1], s[2], s[3]);
            }
        }
        CK(cudaMalloc(&ctx->d_prims, prims.size() * sizeof(float4)));
        CK(cudaMemcpy(ctx->d_prims, prims.data(), prims.size() * sizeof(float4), cudaMemcpyHostToDevice));
        if (ctx->is_tri) CK(cudaMalloc(&ctx->d_cam_prims, (size_t)n * 3 * sizeof(float4)));
        CK(cudaMalloc(&ctx->d_slot_prim, (size_t)n * sizeof(int)));
        CK(cudaMemcpy(ctx->d_slot_prim, ctx->prim_index.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        // device nodes: box + code (code >= 0 child-pair index, code <= -2 leaf ~((first << 3) | count)), sibling pairs interleaved
        // (rt_device.cuh node_slot)
        const int64_t n_rec = (n_nodes + 1) & ~(int64_t)1;
        std::vector<float> dn((size_t)n_rec * 8, 0.0f);
        for (int64_t k = 0; k < n_nodes; ++k) {
            const rt_bvh_node& nd = ctx->nodes[k];
            const int32_t code = nd.b == 0 ? nd.a : ~((nd.a << 3) | nd.b);
            for (int c = 0; c < 3; ++c) { dn[node_slot((int)k, c)] = nd.bmin[c]; dn[node_slot((int)k, 4 + c)] = nd.bmax[c]; }
            std::memcpy(&dn[node_slot((int)k, 3)], &code, 4);
        }
        CK(cudaMalloc(&ctx->d_nodes, dn.size() * sizeof(float)));
        CK(cudaMemcpy(ctx->d_nodes, dn.data(), dn.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    const int m = ctx->m;
    if (m > 0) {
        std::vector<float4> mats((size_t)m * 2);
        for (int k = 0; k < m; ++k) {
            const float* s = &ctx->mats[8 * (size_t)k];
            mats[2 * k + 0] = make_float4(s[0], s[1], s[2], s[3]);   // albedo | metallic
            mats[2 * k + 1] = make_float4(s[4], s[5], s[6], s[7]);   // roughness | emission
        }
        CK(cudaMalloc(&ctx->d_mats, mats.size() * sizeof(float4)));
        CK(cudaMemcpy(ctx->d_mats, mats.data(), mats.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    ctx->device_valid = true;
    return 0;
}

int ensure_bvh(rt_ctx* ctx) {
    if (ctx->bvh_valid) return 0;
    return rt_build_bvh(ctx, ctx->builder);
}

// (Re)build the staged treelet for the current tree when option "treelet" asks for one.
int ensure_treelet(rt_ctx* ctx, cudaStream_t stream) {
    if (ctx->treelet_levels <= 0 || ctx->n_nodes <= 2 || !ctx->d_nodes) { ctx->treelet_pairs = 0; return 0; }
    if (ctx->treelet_valid) return 0;
    const int pairs = (1 << ctx->treelet_levels) - 1;
    cudaFree(ctx->d_treelet); ctx->d_treelet = nullptr;
    CK(cudaMalloc(&ctx->d_treelet, (size_t)pairs * 64));
    CK(launch_build_treelet(ctx->d_nodes, pairs, ctx->d_treelet, stream));
    ctx->treelet_pairs = pairs; ctx->treelet_valid = true;
    return 0;
}

// (Re)build the compressed pairs / split triangle records for the current tree when option "qnodes" asks for them.
// Scenes of <= 64 primitives never reach the wavefront's incoherent-bounce kernel and are skipped.
int ensure_qnodes(rt_ctx* ctx, cudaStream_t stream) {
    if (ctx->qmode == 0 || ctx->n <= 64 || ctx->n_nodes <= 2 || !ctx->d_nodes || !ctx->device_valid) { ctx->qmode_used = 0; return 0; }
    if (ctx->qnodes_valid) return 0;
    ctx->qmode_used = 0;
    if (!(ctx->root_extent < 0x1p40f)) return 0;
    const int64_t pairs = (ctx->n_nodes + 1) / 2;
    if (pairs > ctx->q_pairs) {
        cudaFree(ctx->d_qnodes); ctx->d_qnodes = nullptr; ctx->q_pairs = 0;
        CK(cudaMalloc(&ctx->d_qnodes, (size_t)pairs * 32));
        ctx->q_pairs = pairs;
    }
    if (!ctx->d_qgrid) CK(cudaMalloc(&ctx->d_qgrid, 64));              // 6 floats of grid, then (at byte 32) 3 doubles of quality
    double* d_quality = reinterpret_cast<double*>(ctx->d_qgrid + 8);
    CK(launch_quantize_pairs(ctx->d_nodes, (int)pairs, ctx->d_qnodes, ctx->d_qgrid, d_quality, stream));
    double quality[3] = {0.0, 0.0, 0.0};
    CK(cudaMemcpyAsync(quality, d_quality, sizeof(quality), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));                                  // once per tree (build / refit)
    ctx->q_area_pct = quality[0] > 0.0 ? (int64_t)(100.0 * quality[1] / quality[0] + 0.5) : 100;
    int mode = ctx->qmode > 0 ? ctx->qmode : (ctx->q_area_pct <= ctx->q_area_limit ? 5 : 4);   // auto: cooperative leaves, pairs when the grid is fine enough
    if (quality[2] != 0.0) { mode &= ~1; ctx->q_area_pct = -1; }                        // a box outside the root box (caller-supplied tree): full records only
    if ((mode & 4) && (!ctx->is_tri || ctx->n >= ((int64_t)1 << 25))) mode &= ~4;      // the pair table packs the slot in 25 bits
    if ((mode & 2) && ctx->is_tri) {
        if (ctx->n > ctx->q_tris) {
            cudaFree(ctx->d_tri_a); cudaFree(ctx->d_tri_b); ctx->d_tri_a = ctx->d_tri_b = nullptr; ctx->q_tris = 0;
            CK(cudaMalloc(&ctx->d_tri_a, (size_t)ctx->n * 32));
            CK(cudaMalloc(&ctx->d_tri_b, (size_t)ctx->n * 16));
            ctx->q_tris = ctx->n;
        }
        CK(launch_split_tris(ctx->d_prims, (int)ctx->n, ctx->d_tri_a, ctx->d_tri_b, stream));
        ctx->launches += 1;
    }
    ctx->launches += 1;
    ctx->qmode_used = ctx->is_tri ? mode : (mode & 1);
    ctx->qnodes_valid = true;
    return 0;
}

SceneView scene_view(const rt_ctx* c) {
    SceneView v;
    const bool q = c->qmode_used != 0 && c->qnodes_valid;
    v.qnodes = q ? c->d_qnodes : nullptr; v.qgrid = q ? c->d_qgrid : nullptr;
    v.tri_a = q && (c->qmode_used & 2) ? c->d_tri_a : nullptr; v.tri_b = q && (c->qmode_used & 2) ? c->d_tri_b : nullptr;
    v.treelet = (c->treelet_levels > 0 && c->treelet_valid && c->treelet_pairs > 0) ? c->d_treelet : nullptr;
    v.treelet_two_t = v.treelet ? 2 * c->treelet_pairs : 0;
    v.nodes = c->d_nodes; v.prims = c->d_prims; v.cam_prims = c->d_cam_prims; v.slot_prim = c->d_slot_prim; v.mats = c->d_mats;
    v.n_prims = (int)c->n; v.n_nodes = (int)c->n_nodes; v.n_mats = c->m;
    v.sane_extent = 0;
    if (c->n_nodes > 0) v.sane_extent = c->root_extent < 0x1p40f ? 1 : 0;      // NaN compares false
    v.bg_r = c->bg[0]; v.bg_g = c->bg[1]; v.bg_b = c->bg[2];
    return v;
}

// Kernel choice when option "kernel" is -1 (auto), from B200 measurements (DESIGN.md "Kernel choice"):
// tiny scenes are shading-bound and favour the lock-step megakernel; multi-bounce paths favour the
// wavefront queues, with the coherent bounce 0 walked by packets; single-segment (camera-ray) work favours
// the packet kernel.
bool tiny_ok(const rt_ctx* c, int max_depth) {
    return c->n > 0 && c->n <= kTinyMaxPrims && c->m > 0 && c->m <= kTinyMaxMats && max_depth >= 1 && max_depth <= kTinyMaxDepth;
}
int pick_kernel(const rt_ctx* c, int max_depth) {
    if (max_depth == 0) return 1;                                  // RayTracer::trace_ray(depth <= 0): no segment, black frame -- one kernel for it
    if (c->kernel >= 0 && !(c->kernel == 5 && !tiny_ok(c, max_depth))) {
        if (c->kernel == 3 && max_depth != 1) return 0;        // packets handle camera rays only
        if (c->kernel == 4 && max_depth == 1) return 3;
        return c->kernel;
    }
    if (c->n > 0 && c->n <= 64 && max_depth >= 2 && c->tune_state >= 2)       // tiny scene, both candidates timed (tunes()): the faster one
        return c->tune_rate[1] < c->tune_rate[0] ? (tiny_ok(c, max_depth) ? 5 : 2) : 1;
    if (tiny_ok(c, max_depth) && max_depth >= 2) return 5;     // whole scene in shared memory (timed against variant 1, see tunes())
    if (c->n <= 64) return 1;
    return max_depth >= 2 ? 4 : 3;
}
bool is_wavefront(int variant) { return variant == 2 || variant == 4; }

// candidates: the lock-step megakernel against the tiny-scene kernel (whole scene in shared memory, rt_tiny.cu) when the
// scene is one it takes, else against the wavefront
bool tunes(const rt_ctx* c, int max_depth) { return c->kernel < 0 && c->n > 0 && c->n <= 64 && max_depth >= 2; }
int tune_candidate(const rt_ctx* c, int max_depth, int k) { return k == 0 ? 1 : (tiny_ok(c, max_depth) ? 5 : 2); }

// Before a tuned render: collect the timing of the previous one, then say which variant runs now.
int tune_begin(rt_ctx* c, cudaStream_t stream, double samples, int max_depth) {
    if (!c->tune_ev0) { cudaEventCreate(&c->tune_ev0); cudaEventCreate(&c->tune_ev1); }
    if (c->tune_pending >= 0) {
        float ms = 0.0f;
        if (cudaEventSynchronize(c->tune_ev1) == cudaSuccess && cudaEventElapsedTime(&ms, c->tune_ev0, c->tune_ev1) == cudaSuccess) {
            c->tune_rate[c->tune_pending] = ms / c->tune_pending_samples;
            c->tune_state = c->tune_pending + 1;
        }
        c->tune_pending = -1;
    }
    if (c->tune_state >= 2) return tune_candidate(c, max_depth, c->tune_rate[1] < c->tune_rate[0] ? 1 : 0);
    c->tune_pending = c->tune_state;
    c->tune_pending_samples = samples;
    cudaEventRecord(c->tune_ev0, stream);
    return tune_candidate(c, max_depth, c->tune_state);
}
void tune_end(rt_ctx* c, cudaStream_t stream) { if (c->tune_pending >= 0) cudaEventRecord(c->tune_ev1, stream); }

LaunchCfg launch_cfg(rt_ctx* c, void* stream, int max_depth = 1, int variant = -1) {
    LaunchCfg cfg;
    cfg.stream = (cudaStream_t)stream;
    cfg.sm_count = c->sm_count;
    cfg.d_work_counter = c->d_work_counter;
    cfg.d_stats = c->stats ? c->d_stats : nullptr;
    cfg.variant = variant >= 0 ? variant : pick_kernel(c, max_depth);
    c->kernel_used = cfg.variant;
    cfg.d_cam_prims = c->d_cam_prims;
    cfg.d_planes = nullptr; cfg.plane_batch = 0; cfg.d_fold_cnt = nullptr;
    cfg.cam_table_valid = 0;
    cfg.band = BandSignal{nullptr, nullptr, nullptr, nullptr, 0, 0, 1, 1, 1};
    cfg.sched = ChunkSchedule{nullptr, nullptr, nullptr, 0};
    cfg.d_block_times = c->stats ? c->d_block_times : nullptr;
    cfg.tiny_threads = c->tiny_threads; cfg.tiny_mode = c->tiny_mode; cfg.wf_rays_per_lane = c->wf_rays_per_lane;
    cfg.refill_below = c->refill;
    cfg.leaf_vote = c->leaf_vote;
    cfg.qmode = c->qnodes_valid ? c->qmode_used : 0;
    // the cooperative leaf step serves 8 leaf-holding lanes per pass: measured best with a leaf phase from 7 lanes on and a refill below 24 of 32
    // traversing lanes -- idle lanes still work in the leaf passes, so refilling early costs little (C3 8 spp depth 4: 12.75 ms with the
    // per-lane kernel's 8 / 8, 12.43 with 7 / 16, 12.22 with 7 / 24, 12.55 with 7 / 28; profiles/r02y_exp_qnodes_coop.txt)
    cfg.coop_leaf_vote = c->leaf_vote_set ? c->leaf_vote : 7;
    cfg.coop_refill = c->refill_set ? c->refill : 24;
    return cfg;
}

// Attach the chunk-cost history to a packet launch over tile map `tm` (ChunkSchedule).  The history is
// only meaningful for the tile map it was recorded on; any other map starts from raster order.
// Multi-sample camera-ray frames run the packet kernel in item mode ((block, sample) work items, rt_kernels.cu):
// give the launch a planes buffer of up to 512 MB, i.e. plane_batch samples per pass.
int attach_planes(rt_ctx* ctx, LaunchCfg& cfg, const TileMap& tm, int spp) {
    if (cfg.variant != 3 || spp <= 1) return 0;
    const int64_t n_tasks = task_count(tm);
    if (n_tasks <= 0) return 0;
    int64_t batch = ((int64_t)512 << 20) / (n_tasks * (int64_t)sizeof(float4));
    if (batch > spp) batch = spp;
    if (batch < 1) return 0;                                // frame too large: sample loop inside the block
    const size_t need = (size_t)(batch * n_tasks);
    if (need > ctx->planes_cap) {
        cudaDeviceSynchronize();
        cudaFree(ctx->d_planes); ctx->d_planes = nullptr; ctx->planes_cap = 0;
        CK(cudaMalloc(&ctx->d_planes, need * sizeof(float4)));
        ctx->planes_cap = need;
    }
    cfg.d_planes = ctx->d_planes; cfg.plane_batch = (int)batch;
    if (ctx->fold && batch >= spp) {                          // one batch: the kernel folds the planes itself
        const size_t blocks = (size_t)(n_tasks / 32);
        if (blocks > ctx->fold_cap) {
            cudaDeviceSynchronize();
            cudaFree(ctx->d_fold_cnt); ctx->d_fold_cnt = nullptr; ctx->fold_cap = 0;
            CK(cudaMalloc(&ctx->d_fold_cnt, blocks * sizeof(unsigned int)));
            ctx->fold_cap = blocks;
        }
        cfg.d_fold_cnt = ctx->d_fold_cnt;
    }
    return 0;
}

int attach_schedule(rt_ctx* ctx, LaunchCfg& cfg, const TileMap& tm, const CameraBlock& cam, int spp = 1) {
    cfg.sched = ChunkSchedule{nullptr, nullptr, nullptr, 0};
    if (!ctx->schedule || cfg.variant != 3) return 0;
    if (cfg.d_planes && cfg.plane_batch < spp) return 0;     // several sample batches per frame: raster order
    const int items_per_block = cfg.d_planes ? spp : 1;
    const int n_chunks = packet_chunks(tm, items_per_block);
    if (n_chunks <= 0) return 0;
    if (n_chunks > ctx->chunk_cap) {
        cudaFree(ctx->d_chunk_order); cudaFree(ctx->d_chunk_cost);
        ctx->d_chunk_order = nullptr; ctx->d_chunk_cost = nullptr; ctx->chunk_cap = 0; ctx->chunk_key = -1;
        CK(cudaMalloc(&ctx->d_chunk_order, (size_t)n_chunks * sizeof(int)));
        CK(cudaMalloc(&ctx->d_chunk_cost, (size_t)n_chunks * 2 * sizeof(unsigned int)));
        ctx->chunk_cap = n_chunks;
    }
    long long key = ((((long long)tm.width * 65537 + tm.height) * 257 + tm.tile_w) * 257 + tm.tile_h) * 1031 + tm.first_tile;
    key = key * 1031 + tm.tile_stride + 7919LL * tm.compact + 104729LL * tm.n_local_tiles + 15485863LL * tm.skew + (cfg.band.cnt != nullptr ? 32452843LL : 0LL) + (cfg.band.host_fb != nullptr ? 86028121LL : 0LL) + 49979687LL * items_per_block;
    // the order is rebuilt (k_chunk_order, ~12 us) for a new tile map, whenever the camera has changed since it was
    // built, on the 2nd frame (first one with costs) and then every 8th frame; in between the costs accumulate
    const bool new_map = key != ctx->chunk_key;
    if (new_map) {
        CK(cudaMemsetAsync(ctx->d_chunk_cost, 0, (size_t)ctx->chunk_cap * 2 * sizeof(unsigned int), cfg.stream));
        ctx->chunk_key = key; ctx->chunk_frames = 0; ctx->chunk_age = 0;
    }
    const bool cam_moved = std::memcmp(&ctx->chunk_cam, &cam, sizeof(CameraBlock)) != 0;
    if (new_map || cam_moved || ctx->chunk_age == 1 || ctx->chunk_age >= 8) {
        cfg.sched.reorder_frames = ctx->chunk_frames > 0 ? ctx->chunk_frames : 1;     // no costs yet: raster order
        ctx->chunk_frames = 0; ctx->chunk_age = new_map ? 0 : 1; ctx->chunk_cam = cam;
        if (new_map) ctx->chunk_age = 0;
    }
    ctx->chunk_frames += 1; ctx->chunk_age += 1;
    cfg.sched.order = ctx->d_chunk_order;
    cfg.sched.cost_sum = ctx->d_chunk_cost;
    cfg.sched.cost_max = ctx->d_chunk_cost + ctx->chunk_cap;
    return 0;
}

// The camera-relative triangle table depends on the scene and on the camera POSITION only: a launch whose camera
// sits where the table was built for (progressive batches, a turning camera) reuses it.
void claim_cam_table(rt_ctx* ctx, LaunchCfg& cfg, const CameraBlock& cam) {
    if (!ctx->is_tri || cfg.variant == 5) return;              // the tiny-scene kernel keeps its own table in shared memory
    const bool same = ctx->cam_table_ok && ctx->cam_table_stream == cfg.stream && ctx->cam_table_pos[0] == cam.px && ctx->cam_table_pos[1] == cam.py && ctx->cam_table_pos[2] == cam.pz;
    cfg.cam_table_valid = same ? 1 : 0;
    ctx->cam_table_pos[0] = cam.px; ctx->cam_table_pos[1] = cam.py; ctx->cam_table_pos[2] = cam.pz;
    ctx->cam_table_ok = true; ctx->cam_table_stream = cfg.stream;
}

// launches of one packet-kernel call: k_chunk_order (if scheduled) + k_cam_tris (triangles) + k_packet
// kernels of one launch_render / launch_trace_primary call (every variant builds the camera table when it is stale)
int render_launches(const rt_ctx* ctx, const LaunchCfg& cfg, int spp = 1) {
    if (cfg.variant == 5) return 1;
    const int table = (ctx->is_tri && !cfg.cam_table_valid) ? 1 : 0;
    if (cfg.variant != 3) return 1 + table;
    const int order = (cfg.sched.order && cfg.sched.reorder_frames > 0) ? 1 : 0;
    if (cfg.d_planes && spp > 1 && cfg.d_fold_cnt) return 1 + table + order;
    if (cfg.d_planes && spp > 1) return 2 * ((spp + cfg.plane_batch - 1) / cfg.plane_batch) + table + order;
    return 1 + table + order;
}

TileMap full_frame_map(int width, int height) {
    TileMap tm;
    tm.width = width; tm.height = height;
    tm.tile_w = 32; tm.tile_h = 32;                       // TILE_SIZE, old/raytracer_core copy.cpp:264
    tm.tiles_x = (width + 31) / 32;
    tm.n_tiles = tm.tiles_x * ((height + 31) / 32);
    tm.first_tile = 0; tm.tile_stride = 1; tm.n_local_tiles = tm.n_tiles;
    tm.compact = 0; tm.skew = 0;
    return tm;
}

int check_frame(rt_ctx* ctx, int width, int height) {
    if (width <= 0 || height <= 0 || (int64_t)width * height > (int64_t)1 << 30) return fail(ctx, "invalid frame size");
    return 0;
}

}  // namespace

extern "C" {

int rt_abi_version(void) { return B200RT_ABI_VERSION; }

int rt_create(int device, rt_ctx** out) {
    if (!out) return fail(nullptr, "rt_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, std::string("rt_create: no CUDA device (") + cudaGetErrorString(e) +
                                 "); libb200rt has no CPU fallback");
    if (device < 0 || device >= count) return fail(nullptr, "rt_create: device index out of range");
    rt_ctx* ctx = new rt_ctx();
    ctx->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { delete ctx; return cuda_fail(nullptr, "cudaGetDeviceProperties", e); }
    if (prop.major < 10) {
        delete ctx;
        return fail(nullptr, "rt_create: device is not Blackwell (sm_100a) -- this library ships sm_100a code only");
    }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaMalloc(&ctx->d_work_counter, 64)) != cudaSuccess || (e = cudaMalloc(&ctx->d_stats, 64)) != cudaSuccess ||
        (e = cudaMemset(ctx->d_stats, 0, 64)) != cudaSuccess || (e = cudaMalloc(&ctx->d_pick, 64)) != cudaSuccess) {
        delete ctx;
        return cuda_fail(nullptr, "rt_create: cudaMalloc", e);
    }
    *out = ctx;
    return 0;
}

void rt_destroy(rt_ctx* ctx) {
    if (!ctx) return;
    {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        DeviceGuard g(ctx->device);
        cudaDeviceSynchronize();
        free_device_scene(ctx);
        free_wave(ctx);
        cudaFree(ctx->d_treelet); cudaFree(ctx->d_fold_cnt);
        cudaFree(ctx->d_qnodes); cudaFree(ctx->d_qgrid); cudaFree(ctx->d_tri_a); cudaFree(ctx->d_tri_b);
        cudaFree(ctx->d_work_counter); cudaFree(ctx->d_stats); cudaFree(ctx->d_fb); cudaFree(ctx->d_pick); cudaFree(ctx->d_edit); cudaFree(ctx->d_display); cudaFree(ctx->d_planes);
        cudaFree(ctx->d_band_cnt); cudaFree(ctx->d_tile_cnt); cudaFree(ctx->d_chunk_order); cudaFree(ctx->d_chunk_cost);
        if (ctx->h_band_flags) cudaFreeHost(ctx->h_band_flags);
        if (ctx->render_stream) cudaStreamDestroy(ctx->render_stream);
        if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
        if (ctx->band_stream) cudaStreamDestroy(ctx->band_stream);
        for (auto& e : ctx->band_ev) if (e) cudaEventDestroy(e);
        if (ctx->tune_ev0) { cudaEventDestroy(ctx->tune_ev0); cudaEventDestroy(ctx->tune_ev1); }
        if (ctx->scratch_ev) cudaEventDestroy(ctx->scratch_ev);
        for (int k = 0; k < 2; ++k) {
            if (ctx->pipe.streams[k]) cudaStreamDestroy(ctx->pipe.streams[k]);
            if (ctx->pipe.join[k]) cudaEventDestroy(ctx->pipe.join[k]);
            if (ctx->pipe.acc[k]) cudaEventDestroy(ctx->pipe.acc[k]);
        }
        if (ctx->pipe.fork) cudaEventDestroy(ctx->pipe.fork);
        for (int k = 0; k < 2; ++k) { if (ctx->h_stage[k]) cudaFreeHost(ctx->h_stage[k]); if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]); }
        if (ctx->stage_stream) cudaStreamDestroy(ctx->stage_stream);
    }
    delete ctx;
}

const char* rt_last_error(rt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rt_set_spheres(rt_ctx* ctx, const float* cr, const float* mat8, const int32_t* object_id, int64_t n) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || n > (int64_t)1 << 28) return fail(ctx, "rt_set_spheres: invalid count");
    if (n > 0 && (!cr || !mat8)) return fail(ctx, "rt_set_spheres: NULL array");
    ctx->is_tri = false; ctx->n = n; ctx->m = (int)n;
    ctx->prim_data.assign(cr, cr + 4 * n);
    ctx->mats.assign(mat8, mat8 + 8 * n);
    ctx->mat_id.clear();
    ctx->object_id.resize(n);
    for (int64_t i = 0; i < n; ++i) ctx->object_id[i] = object_id ? object_id[i] : (int32_t)i;
    ctx->bvh_valid = false; ctx->device_valid = false; ctx->tune_state = 0; ctx->tune_pending = -1;
    return 0;
}

int rt_set_triangles(rt_ctx* ctx, const float* v, const int32_t* material_id, int64_t n, const float* materials, int m) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || n > (int64_t)1 << 28 || m < 0) return fail(ctx, "rt_set_triangles: invalid count");
    if (n > 0 && (!v || !materials || m == 0)) return fail(ctx, "rt_set_triangles: NULL array / no material");
    ctx->is_tri = true; ctx->n = n; ctx->m = m;
    ctx->prim_data.assign(v, v + 9 * n);
    ctx->mats.assign(materials, materials + 8 * (size_t)m);
    ctx->mat_id.resize(n);
    ctx->object_id.resize(n);
    for (int64_t i = 0; i < n; ++i) {
        int32_t id = material_id ? material_id[i] : 0;
        if (id < 0 || id >= m) return fail(ctx, "rt_set_triangles: material_id out of range");
        ctx->mat_id[i] = id;
        ctx->object_id[i] = (int32_t)i;
    }
    ctx->bvh_valid = false; ctx->device_valid = false; ctx->tune_state = 0; ctx->tune_pending = -1;
    return 0;
}

int rt_update_geometry(rt_ctx* ctx, const float* h_prims, int64_t n) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n != ctx->n) return fail(ctx, "rt_update_geometry: primitive count differs from the uploaded scene (use rt_set_spheres / rt_set_triangles)");
    if (n == 0) return 0;
    if (!h_prims) return fail(ctx, "rt_update_geometry: NULL array");
    const size_t per = ctx->is_tri ? 9 : 4;
    ctx->cam_table_ok = false;
    if (!ctx->bvh_valid || !ctx->device_valid) {                           // nothing built yet: the next launch builds
        ctx->prim_data.assign(h_prims, h_prims + per * (size_t)n);
        ctx->bvh_valid = false; ctx->device_valid = false;
        return 0;
    }
    ctx->prim_data.resize(per * (size_t)n);                                // (filled by the staged upload below)
    DeviceGuard g(ctx->device);
    const size_t raw_bytes = per * (size_t)n * sizeof(float), raw_pad = (raw_bytes + 255) & ~(size_t)255;
    const size_t need = raw_pad + (size_t)ctx->n_nodes * 32;
    if (need > ctx->edit_bytes) {
        cudaFree(ctx->d_edit); ctx->d_edit = nullptr; ctx->edit_bytes = 0;
        CK(cudaMalloc(&ctx->d_edit, need));
        ctx->edit_bytes = need;
    }
    if (!ctx->d_nodes_abi) CK(cudaMalloc(&ctx->d_nodes_abi, (size_t)ctx->n_nodes * sizeof(rt_bvh_node)));
    CK(cudaDeviceSynchronize());                                           // renders in flight on other streams still read the old boxes
    double* d_area = reinterpret_cast<double*>(ctx->d_stats + 6);         // spare words of the 64-byte stats block
    if (ctx->built_area == 0.0) {                                          // first edit of this tree: how good was it as built?
        CK(bvh_area(ctx->d_nodes, (int)ctx->n_nodes, d_area, ctx->sm_count, nullptr));
        CK(cudaMemcpy(&ctx->built_area, d_area, sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (int rc = staged_upload(ctx, h_prims, raw_bytes, ctx->d_edit, ctx->prim_data.data())) return rc;
    CK(bvh_refit(ctx->d_nodes, ctx->d_nodes_abi, (int)ctx->n_nodes, ctx->d_slot_prim, static_cast<const float*>(ctx->d_edit), ctx->is_tri,
                 ctx->d_prims, (int)n, static_cast<char*>(ctx->d_edit) + raw_pad, ctx->sm_count, nullptr));
    CK(bvh_area(ctx->d_nodes, (int)ctx->n_nodes, d_area, ctx->sm_count, nullptr));
    ctx->launches += 5;
    ctx->refits += 1;
    ctx->treelet_valid = false; ctx->qnodes_valid = false;
    double area = 0.0;
    CK(cudaMemcpy(&area, d_area, sizeof(double), cudaMemcpyDeviceToHost));            // synchronises
    ctx->area_pct = ctx->built_area > 0.0 ? (int64_t)(100.0 * area / ctx->built_area) : 100;
    if (ctx->refit_limit > 0 && ctx->area_pct > ctx->refit_limit) {                   // the edit wrecked the tree: build a new one
        ctx->refit_rebuilds += 1;                                                     // (lazily, with option "builder", at the next launch)
        ctx->bvh_valid = false; ctx->device_valid = false;
        return 0;
    }
    rt_bvh_node root;
    CK(cudaMemcpy(&root, ctx->d_nodes_abi, sizeof(root), cudaMemcpyDeviceToHost));
    ctx->root_extent = 0.0f;
    for (int k = 0; k < 3; ++k) ctx->root_extent = std::fmax(ctx->root_extent, std::fmax(std::fabs(root.bmin[k]), std::fabs(root.bmax[k])));
    ctx->host_bvh_stale = true;                                            // rt_get_bvh downloads the refitted boxes
    return 0;
}

int rt_update_materials(rt_ctx* ctx, const float* h_material8, int m) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (m != ctx->m) return fail(ctx, "rt_update_materials: material count differs from the uploaded scene");
    if (m == 0) return 0;
    if (!h_material8) return fail(ctx, "rt_update_materials: NULL array");
    ctx->mats.assign(h_material8, h_material8 + 8 * (size_t)m);
    ctx->tune_state = 0; ctx->tune_pending = -1;
    if (!ctx->device_valid || !ctx->d_mats) return 0;                        // uploaded with the scene later
    DeviceGuard g(ctx->device);
    std::vector<float4> mats((size_t)m * 2);
    for (int k = 0; k < m; ++k) {
        const float* q = &ctx->mats[8 * (size_t)k];
        mats[2 * k + 0] = make_float4(q[0], q[1], q[2], q[3]);
        mats[2 * k + 1] = make_float4(q[4], q[5], q[6], q[7]);
    }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ctx->d_mats, mats.data(), mats.size() * sizeof(float4), cudaMemcpyHostToDevice));
    return 0;
}

int rt_set_background(rt_ctx* ctx, const float rgb[3]) {
    if (!ctx || !rgb) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    std::memcpy(ctx->bg, rgb, 12);
    return 0;
}

// the host vectors hold the tree: refresh the summary fields
static void note_host_tree(rt_ctx* ctx) {
    ctx->n_nodes = (int64_t)ctx->nodes.size();
    ctx->host_bvh_stale = false;
    ctx->root_extent = 0.0f;
    if (!ctx->nodes.empty())
        for (int k = 0; k < 3; ++k)
            ctx->root_extent = std::fmax(ctx->root_extent, std::fmax(std::fabs(ctx->nodes[0].bmin[k]), std::fabs(ctx->nodes[0].bmax[k])));
}

// builder 1: everything on the device (rt_lbvh.cu); the host mirror of the tree (rt_get_bvh, the stack bound) is
// downloaded afterwards.  Materials are uploaded like ensure_device does.
static int build_bvh_device(rt_ctx* ctx) {
    DeviceGuard g(ctx->device);
    free_device_scene(ctx);
    ctx->nodes.clear(); ctx->prim_index.clear(); ctx->n_nodes = 0; ctx->root_extent = 0.0f; ctx->host_bvh_stale = false;
    const int64_t n = ctx->n;
    if (n > 0) {
        float* d_raw = nullptr;
        int* d_mid = nullptr;
        const size_t raw_bytes = ctx->prim_data.size() * sizeof(float);
        CK(cudaMalloc(&d_raw, raw_bytes));
        cudaError_t e = cudaMemcpy(d_raw, ctx->prim_data.data(), raw_bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && ctx->is_tri) {
            e = cudaMalloc(&d_mid, (size_t)n * sizeof(int));
            if (e == cudaSuccess) e = cudaMemcpy(d_mid, ctx->mat_id.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice);
        }
        LbvhResult r;
        if (e == cudaSuccess) e = lbvh_build(d_raw, d_mid, ctx->is_tri, (int)n, ctx->sm_count, nullptr, &r);
        cudaFree(d_raw); cudaFree(d_mid);
        if (e != cudaSuccess) return cuda_fail(ctx, "rt_build_bvh(device)", e);
        ctx->d_nodes = r.d_nodes; ctx->d_prims = r.d_prims; ctx->d_slot_prim = r.d_prim_index;
        ctx->d_nodes_abi = r.d_nodes_abi; ctx->host_bvh_stale = true;
        ctx->n_nodes = r.n_nodes;
        rt_bvh_node root;
        e = cudaMemcpy(&root, r.d_nodes_abi, sizeof(root), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return cuda_fail(ctx, "rt_build_bvh(device): download", e);
        ctx->root_extent = 0.0f;
        for (int k = 0; k < 3; ++k) ctx->root_extent = std::fmax(ctx->root_extent, std::fmax(std::fabs(root.bmin[k]), std::fabs(root.bmax[k])));
        ctx->bvh_depth = r.depth;
        if (ctx->bvh_depth > kStackDepth - 2) {
            free_device_scene(ctx);
            return fail(ctx, "rt_build_bvh: device-built tree deeper than the traversal stack (duplicate primitives?); use builder 0");
        }
        if (ctx->is_tri) CK(cudaMalloc(&ctx->d_cam_prims, (size_t)n * 3 * sizeof(float4)));
    }
    if (ctx->m > 0) {
        std::vector<float4> mats((size_t)ctx->m * 2);
        for (int k = 0; k < ctx->m; ++k) {
            const float* s = &ctx->mats[8 * (size_t)k];
            mats[2 * k + 0] = make_float4(s[0], s[1], s[2], s[3]);
            mats[2 * k + 1] = make_float4(s[4], s[5], s[6], s[7]);
        }
        CK(cudaMalloc(&ctx->d_mats, mats.size() * sizeof(float4)));
        CK(cudaMemcpy(ctx->d_mats, mats.data(), mats.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    ctx->bvh_valid = true; ctx->device_valid = true; ctx->built_area = 0.0;
    return 0;
}

int rt_build_bvh(rt_ctx* ctx, int builder) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (builder < 0 || builder > 2) return fail(ctx, "rt_build_bvh: builder must be 0 (reference median split, host), 1 (LBVH, device) or 2 (binned SAH, host)");
    ctx->tune_state = 0; ctx->tune_pending = -1;
    if (builder == 1) return build_bvh_device(ctx);
    PrimBoxes boxes;
    if (ctx->is_tri) triangle_boxes(ctx->prim_data.data(), ctx->n, boxes);
    else sphere_boxes(ctx->prim_data.data(), ctx->n, boxes);
    const char* msg;
    if (builder == 2) {
        build_sah(boxes, ctx->n, ctx->nodes, ctx->prim_index, ctx->leaf_size, 0.1f * (float)ctx->sah_cost);
        if (validate_bvh(ctx->nodes.data(), (int64_t)ctx->nodes.size(), ctx->n, &msg) > kStackDepth - 2) builder = 0;   // degenerate input: the balanced tree
    }
    if (builder == 0) build_median_split(boxes, ctx->n, ctx->nodes, ctx->prim_index, ctx->leaf_size);
    ctx->bvh_depth = validate_bvh(ctx->nodes.data(), (int64_t)ctx->nodes.size(), ctx->n, &msg);
    if (ctx->bvh_depth < 0) return fail(ctx, msg);
    if (ctx->bvh_depth > kStackDepth - 2) return fail(ctx, "rt_build_bvh: tree deeper than the traversal stack");
    note_host_tree(ctx);
    ctx->bvh_valid = true; ctx->device_valid = false; ctx->built_area = 0.0;
    return 0;
}

int rt_build_bvh_host(const float* h_prims, int is_triangles, int64_t n, rt_bvh_node* h_nodes, int64_t* n_nodes,
                      int32_t* h_prim_index) {
    return rt_build_bvh_host_ex(h_prims, is_triangles, n, 0, 4, h_nodes, n_nodes, h_prim_index);
}

int rt_build_bvh_host_ex(const float* h_prims, int is_triangles, int64_t n, int builder, int leaf_size, rt_bvh_node* h_nodes,
                         int64_t* n_nodes, int32_t* h_prim_index) {
    if (n < 0 || (n > 0 && !h_prims) || !n_nodes) return fail(nullptr, "rt_build_bvh_host: bad arguments");
    if ((builder != 0 && builder != 2) || leaf_size < 1 || leaf_size > 4)
        return fail(nullptr, "rt_build_bvh_host_ex: builder must be 0 (reference median split) or 2 (binned SAH), leaf_size 1..4");
    PrimBoxes boxes;
    if (is_triangles) triangle_boxes(h_prims, n, boxes); else sphere_boxes(h_prims, n, boxes);
    std::vector<rt_bvh_node> nodes;
    std::vector<int32_t> prim_index;
    if (builder == 2) {
        const char* msg;
        build_sah(boxes, n, nodes, prim_index, leaf_size);
        if (validate_bvh(nodes.data(), (int64_t)nodes.size(), n, &msg) > kStackDepth - 2) builder = 0;
    }
    if (builder == 0) build_median_split(boxes, n, nodes, prim_index, leaf_size);
    *n_nodes = (int64_t)nodes.size();
    if (h_nodes && !nodes.empty()) std::memcpy(h_nodes, nodes.data(), nodes.size() * sizeof(rt_bvh_node));
    if (h_prim_index && n) std::memcpy(h_prim_index, prim_index.data(), (size_t)n * sizeof(int32_t));
    return 0;
}

int rt_get_bvh(rt_ctx* ctx, rt_bvh_node* nodes, int64_t* n_nodes, int32_t* prim_index) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (int rc = ensure_bvh(ctx)) return rc;
    if (ctx->host_bvh_stale) {                              // device-built tree: fetch the host mirror now
        DeviceGuard g(ctx->device);
        ctx->nodes.resize(ctx->n_nodes); ctx->prim_index.resize(ctx->n);
        CK(cudaMemcpy(ctx->nodes.data(), ctx->d_nodes_abi, (size_t)ctx->n_nodes * sizeof(rt_bvh_node), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ctx->prim_index.data(), ctx->d_slot_prim, (size_t)ctx->n * sizeof(int), cudaMemcpyDeviceToHost));
        cudaFree(ctx->d_nodes_abi); ctx->d_nodes_abi = nullptr;
        ctx->host_bvh_stale = false;
    }
    if (n_nodes) *n_nodes = (int64_t)ctx->nodes.size();
    if (nodes && !ctx->nodes.empty()) std::memcpy(nodes, ctx->nodes.data(), ctx->nodes.size() * sizeof(rt_bvh_node));
    if (prim_index && ctx->n) std::memcpy(prim_index, ctx->prim_index.data(), (size_t)ctx->n * sizeof(int32_t));
    return 0;
}

int rt_set_bvh(rt_ctx* ctx, const rt_bvh_node* nodes, int64_t n_nodes, const int32_t* prim_index) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n_nodes < 0 || (n_nodes > 0 && (!nodes || !prim_index))) return fail(ctx, "rt_set_bvh: NULL array");
    if (ctx->n > 0 && n_nodes == 0) return fail(ctx, "rt_set_bvh: empty tree for a non-empty scene");
    const char* msg;
    int depth = validate_bvh(nodes, n_nodes, ctx->n, &msg);
    if (depth < 0) return fail(ctx, msg);
    if (depth > kStackDepth - 2) return fail(ctx, "rt_set_bvh: tree deeper than the traversal stack");
    std::vector<char> seen((size_t)ctx->n, 0);
    for (int64_t k = 0; k < ctx->n; ++k) {
        int32_t p = prim_index[k];
        if (p < 0 || p >= ctx->n || seen[p]) return fail(ctx, "rt_set_bvh: prim_index is not a permutation");
        seen[p] = 1;
    }
    ctx->nodes.assign(nodes, nodes + n_nodes);
    ctx->prim_index.assign(prim_index, prim_index + ctx->n);
    ctx->bvh_depth = depth;
    note_host_tree(ctx);
    ctx->bvh_valid = true; ctx->device_valid = false; ctx->built_area = 0.0;
    return 0;
}

int rt_set_camera(rt_ctx* ctx, const double pos[3], const double target[3], const double up[3], double fov_deg, double aspect) {
    if (!ctx || !pos || !target) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    for (int k = 0; k < 3; ++k) { ctx->pos[k] = pos[k]; ctx->target[k] = target[k]; if (up) ctx->up[k] = up[k]; }
    ctx->fov = fov_deg;
    ctx->aspect = aspect;
    return 0;
}

int rt_get_camera_block(rt_ctx* ctx, int width, int height, double out[14]) {
    if (!ctx || !out) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    double aspect = (width > 0 && height > 0) ? (double)width / height : ctx->aspect;
    CameraBlock b = camera_block(ctx, aspect);
    out[0] = b.px; out[1] = b.py; out[2] = b.pz;
    for (int k = 0; k < 3; ++k) { out[3 + k] = b.fwd[k]; out[6 + k] = b.right[k]; out[9 + k] = b.up[k]; }
    out[12] = b.sx; out[13] = b.sy;
    return 0;
}

int rt_trace_primary(rt_ctx* ctx, int width, int height, int32_t* d_prim, float* d_t, void* stream) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (int rc = check_frame(ctx, width, height)) return rc;
    if (!d_prim || !d_t) return fail(ctx, "rt_trace_primary: NULL output");
    DeviceGuard g(ctx->device);
    if (int rc = ensure_device(ctx)) return rc;
    ctx->aspect = (double)width / height;            // RayTracer::render, old/raytracer_core copy.cpp:259
    CameraBlock cam = camera_block(ctx, ctx->aspect);
    TileMap tm = full_frame_map(width, height);
    if (is_wavefront(pick_kernel(ctx, 1))) { if (int rc = ensure_wave(ctx, task_count(tm), 1, 1)) return rc; }
    ScratchOrder order(ctx, stream);
    if (is_wavefront(pick_kernel(ctx, 1))) {
        int nl = 0;
        LaunchCfg wcfg = launch_cfg(ctx, stream);
        claim_cam_table(ctx, wcfg, cam);
        CK(launch_wavefront(scene_view(ctx), ctx->is_tri, true, cam, tm, 1, 1, 0, 0, 0, 0, nullptr, d_prim, d_t, wcfg, ctx->wave, &nl, wave_pipe(ctx)));
        ctx->launches += nl;
        return 0;
    }
    LaunchCfg cfg = launch_cfg(ctx, stream);
    if (cfg.variant == 5) { cfg.variant = 1; ctx->kernel_used = 1; }   // the primary-hit query of a tiny scene: one ray per thread
    if (int rc = attach_schedule(ctx, cfg, tm, cam)) return rc;
    claim_cam_table(ctx, cfg, cam);
    CK(launch_trace_primary(scene_view(ctx), ctx->is_tri, cam, tm, d_prim, d_t, cfg));
    ctx->launches += render_launches(ctx, cfg);
    return 0;
}

int rt_trace_rays(rt_ctx* ctx, const float* d_origin, const float* d_direction, int64_t n, int32_t* d_prim, float* d_t,
                  void* stream) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || (n > 0 && (!d_origin || !d_direction || !d_prim || !d_t))) return fail(ctx, "rt_trace_rays: bad arguments");
    DeviceGuard g(ctx->device);
    if (int rc = ensure_device(ctx)) return rc;
    ScratchOrder order(ctx, stream);
    CK(launch_trace_rays(scene_view(ctx), ctx->is_tri, d_origin, d_direction, n, d_prim, d_t, launch_cfg(ctx, stream)));
    if (n) ctx->launches += 1;
    return 0;
}

int rt_select_object(rt_ctx* ctx, double x, double y, int width, int height, int32_t* out_object_id) {
    if (!ctx || !out_object_id) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (int rc = ensure_device(ctx)) return rc;
    (void)width; (void)height;                       // the reference ignores them too (camera.aspect_ratio is used)
    CameraBlock b = camera_block(ctx, ctx->aspect);
    double vx = (x - 0.5) * 2.0 * b.sx, vy = (0.5 - y) * 2.0 * b.sy;
    float host[8];
    host[0] = b.px; host[1] = b.py; host[2] = b.pz;
    double d[3];
    for (int k = 0; k < 3; ++k) d[k] = b.fwd[k] + b.right[k] * vx + b.up[k] * vy;
    double l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int k = 0; k < 3; ++k) host[3 + k] = (float)(d[k] / l);
    float* dv = reinterpret_cast<float*>(ctx->d_pick);
    CK(cudaMemcpy(dv, host, 24, cudaMemcpyHostToDevice));
    LaunchCfg cfg = launch_cfg(ctx, nullptr);
    cfg.d_stats = nullptr;
    CK(launch_trace_rays(scene_view(ctx), ctx->is_tri, dv, dv + 3, 1, ctx->d_pick + 6, dv + 7, cfg));
    ctx->launches += 1;
    int32_t prim; float t;
    CK(cudaMemcpy(&prim, ctx->d_pick + 6, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&t, dv + 7, 4, cudaMemcpyDeviceToHost));
    // cast_ray_for_selection(ray, 0.001, 1000.0), old/raytracer_core copy.cpp:247
    *out_object_id = (prim >= 0 && t <= 1000.0f) ? ctx->object_id[prim] : -1;
    return 0;
}

// layout 0: compact per-rank tile buffer; 1: frame layout, tiles dealt skewed; 2: frame layout, plain row-major tiles
// (tile_cap >= 0: at most that many tiles -- a band of the frame)
static int render_tiles(rt_ctx* ctx, int width, int height, int tile_w, int tile_h, int first_tile, int tile_stride, int spp,
                        int max_depth, uint64_t seed, uint32_t sample_offset, int resolve, float* d_out, void* stream, int layout,
                        int tile_cap = -1) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (int rc = check_frame(ctx, width, height)) return rc;
    if (tile_w <= 0 || tile_h <= 0 || (tile_w & 7) || (tile_h & 3)) return fail(ctx, "rt_render_tiles: tile must be a multiple of 8x4");
    if (first_tile < 0 || tile_stride <= 0 || spp <= 0 || max_depth < 0 || !d_out) return fail(ctx, "rt_render_tiles: bad arguments");
    DeviceGuard g(ctx->device);
    if (int rc = ensure_device(ctx)) return rc;
    ctx->aspect = (double)width / height;
    CameraBlock cam = camera_block(ctx, ctx->aspect);
    TileMap tm;
    tm.width = width; tm.height = height; tm.tile_w = tile_w; tm.tile_h = tile_h;
    tm.tiles_x = (width + tile_w - 1) / tile_w;
    tm.n_tiles = tm.tiles_x * ((height + tile_h - 1) / tile_h);
    tm.first_tile = first_tile; tm.tile_stride = tile_stride;
    tm.n_local_tiles = first_tile < tm.n_tiles ? (tm.n_tiles - first_tile + tile_stride - 1) / tile_stride : 0;
    if (tile_cap >= 0 && tm.n_local_tiles > tile_cap) tm.n_local_tiles = tile_cap;
    tm.compact = layout == 0 ? 1 : 0; tm.skew = layout == 1 ? 1 : 0;
    if (is_wavefront(pick_kernel(ctx, max_depth)) && tm.n_local_tiles) { if (int rc = ensure_wave(ctx, task_count(tm), spp, max_depth)) return rc; }
    ScratchOrder order(ctx, stream);
    if (is_wavefront(pick_kernel(ctx, max_depth)) && tm.n_local_tiles) {
        int nl = 0;
        LaunchCfg wcfg = launch_cfg(ctx, stream, max_depth);
        claim_cam_table(ctx, wcfg, cam);
        CK(launch_wavefront(scene_view(ctx), ctx->is_tri, false, cam, tm, spp, max_depth, ctx->integrator, seed, sample_offset,
                            resolve, d_out, nullptr, nullptr, wcfg, ctx->wave, &nl, wave_pipe(ctx)));
        ctx->launches += nl;
        return 0;
    }
    LaunchCfg cfg = launch_cfg(ctx, stream, max_depth);
    if (int rc = attach_planes(ctx, cfg, tm, spp)) return rc;
    if (int rc = attach_schedule(ctx, cfg, tm, cam, spp)) return rc;
    claim_cam_table(ctx, cfg, cam);
    CK(launch_render(scene_view(ctx), ctx->is_tri, cam, tm, spp, max_depth, ctx->integrator, seed, sample_offset, resolve,
                     d_out, cfg));
    if (tm.n_local_tiles) ctx->launches += render_launches(ctx, cfg, spp);
    return 0;
}

int rt_render_tiles(rt_ctx* ctx, int width, int height, int tile_w, int tile_h, int first_tile, int tile_stride, int spp,
                    int max_depth, uint64_t seed, uint32_t sample_offset, int resolve, float* d_out, void* stream) {
    return render_tiles(ctx, width, height, tile_w, tile_h, first_tile, tile_stride, spp, max_depth, seed, sample_offset, resolve,
                        d_out, stream, 0);
}

int rt_render_tiles_frame(rt_ctx* ctx, int width, int height, int tile_w, int tile_h, int rank, int world, int spp,
                          int max_depth, uint64_t seed, uint32_t sample_offset, int resolve, float* d_frame, void* stream) {
    if (ctx && (rank < 0 || world <= 0 || rank >= world)) return fail(ctx, "rt_render_tiles_frame: bad rank / world");
    return render_tiles(ctx, width, height, tile_w, tile_h, rank, world, spp, max_depth, seed, sample_offset, resolve, d_frame,
                        stream, 1);
}

// ---- frames shared between the processes of one node (CUDA IPC): the display rank allocates, the others map
// it and render their tiles straight into it over NVLink (rt_render_tiles_frame), so the frame exchange is
// the render kernel's own stores.
int rt_frame_alloc(rt_ctx* ctx, int width, int height, int planes, float** d_frame, unsigned char handle[64]) {
    if (!ctx || !d_frame || !handle) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (int rc = check_frame(ctx, width, height)) return rc;
    if (planes < 1 || planes > 64) return fail(ctx, "rt_frame_alloc: planes must be in 1..64");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(ctx->device);
    float* p = nullptr;
    const size_t frame_bytes = ((size_t)planes * width * height * 3 * sizeof(float) + 255) & ~(size_t)255;
    CK(cudaMalloc(&p, frame_bytes + 256));                               // + the sync words of rt_frame_sync
    CK(cudaMemset(reinterpret_cast<char*>(p) + frame_bytes, 0, 256));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(ctx, "cudaIpcGetMemHandle", e); }
    std::memcpy(handle, &h, 64);
    *d_frame = p;
    return 0;
}

int rt_frame_free(rt_ctx* ctx, float* d_frame) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaFree(d_frame));
    return 0;
}

int rt_frame_open(rt_ctx* ctx, const unsigned char handle[64], float** d_peer_frame) {
    if (!ctx || !handle || !d_peer_frame) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_peer_frame = static_cast<float*>(p);
    return 0;
}

int rt_frame_close(rt_ctx* ctx, float* d_peer_frame) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaIpcCloseMemHandle(d_peer_frame));
    return 0;
}

int rt_frame_sync(rt_ctx* ctx, float* d_frame, int width, int height, int planes, int world, uint64_t epoch, void* stream) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (int rc = check_frame(ctx, width, height)) return rc;
    if (!d_frame || planes < 1 || planes > 64 || world < 1 || epoch == 0) return fail(ctx, "rt_frame_sync: bad arguments");
    DeviceGuard g(ctx->device);
    const size_t frame_bytes = ((size_t)planes * width * height * 3 * sizeof(float) + 255) & ~(size_t)255;
    unsigned long long* words = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d_frame) + frame_bytes);
    CK(launch_frame_sync(words, (unsigned long long)world * epoch, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

// ---- host frames shared between the processes of one node (see b200rt.h)
int rt_host_register(rt_ctx* ctx, void* h_ptr, uint64_t bytes, void** d_alias) {
    if (!ctx || !h_ptr || !d_alias || bytes == 0) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaHostRegister(h_ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    void* d = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&d, h_ptr, 0);
    if (e != cudaSuccess) { cudaHostUnregister(h_ptr); return cuda_fail(ctx, "cudaHostGetDevicePointer", e); }
    *d_alias = d;
    return 0;
}

int rt_host_unregister(rt_ctx* ctx, void* h_ptr) {
    if (!ctx || !h_ptr) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaDeviceSynchronize());
    CK(cudaHostUnregister(h_ptr));
    return 0;
}

int rt_render_tiles_host(rt_ctx* ctx, int width, int height, int rank, int world, int spp, int max_depth, uint64_t seed,
                         uint32_t sample_offset, float* d_host_frame, uint32_t* d_flag, uint32_t epoch, void* stream) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (int rc = check_frame(ctx, width, height)) return rc;
    if (rank < 0 || world <= 0 || rank >= world || !d_host_frame || !d_flag) return fail(ctx, "rt_render_tiles_host: bad arguments");
    if (((uintptr_t)d_host_frame & 15u) != 0) return fail(ctx, "rt_render_tiles_host: the host frame must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    const size_t need = (size_t)width * height * 3;
    if (need > ctx->fb_floats) {
        CK(cudaDeviceSynchronize());
        cudaFree(ctx->d_fb); ctx->d_fb = nullptr; ctx->fb_floats = 0;
        CK(cudaMalloc(&ctx->d_fb, need * sizeof(float)));
        ctx->fb_floats = need;
    }
    if (int rc = render_tiles(ctx, width, height, 32, 32, rank, world, spp, max_depth, seed, sample_offset, 1, ctx->d_fb, stream, 1)) return rc;
    TileMap tm = full_frame_map(width, height);
    tm.first_tile = rank; tm.tile_stride = world; tm.skew = 1;
    tm.n_local_tiles = rank < tm.n_tiles ? (tm.n_tiles - rank + world - 1) / world : 0;
    CK(launch_push_tiles(tm, ctx->d_fb, d_host_frame, d_flag, epoch, ctx->d_work_counter + 8, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

int rt_host_wait(const volatile uint32_t* h_flags, int n, uint32_t epoch, double timeout_s) {
    if (!h_flags || n <= 0) return 1;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0;; ++spins) {
        bool all = true;
        for (int k = 0; k < n; ++k) all = all && h_flags[k] == epoch;
        if (all) { std::atomic_thread_fence(std::memory_order_acquire); return 0; }
        if ((spins & 1023u) == 1023u) {                       // frames take fractions of a millisecond: spin, but look at the clock now and then,
            const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (waited > timeout_s) return 2;
            if (waited > 0.002) std::this_thread::yield();     // and stop hogging the core once the wait is not a frame's any more
        }
    }
}

static int render_frame(rt_ctx* ctx, int width, int height, int spp, int max_depth, uint64_t seed, uint32_t sample_offset,
                        int resolve, float* d_out, void* stream) {
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (int rc = check_frame(ctx, width, height)) return rc;
    if (spp <= 0 || max_depth < 0 || !d_out) return fail(ctx, "rt_render: bad arguments");
    DeviceGuard g(ctx->device);
    if (int rc = ensure_device(ctx)) return rc;
    ctx->aspect = (double)width / height;
    CameraBlock cam = camera_block(ctx, ctx->aspect);
    TileMap tm = full_frame_map(width, height);
    const bool tuned = tunes(ctx, max_depth);
    if (tuned || is_wavefront(pick_kernel(ctx, max_depth))) { if (int rc = ensure_wave(ctx, task_count(tm), spp, max_depth)) return rc; }   // allocate outside the timed events
    ScratchOrder order(ctx, stream);
    const int variant = tuned ? tune_begin(ctx, (cudaStream_t)stream, (double)width * height * spp, max_depth) : pick_kernel(ctx, max_depth);
    if (is_wavefront(variant)) {
        int nl = 0;
        LaunchCfg wcfg = launch_cfg(ctx, stream, max_depth, variant);
        claim_cam_table(ctx, wcfg, cam);
        CK(launch_wavefront(scene_view(ctx), ctx->is_tri, false, cam, tm, spp, max_depth, ctx->integrator, seed, sample_offset,
                            resolve, d_out, nullptr, nullptr, wcfg, ctx->wave, &nl, wave_pipe(ctx)));
        ctx->launches += nl;
        if (tuned) tune_end(ctx, (cudaStream_t)stream);
        return 0;
    }
    LaunchCfg cfg = launch_cfg(ctx, stream, max_depth, variant);
    if (int rc = attach_planes(ctx, cfg, tm, spp)) return rc;
    if (int rc = attach_schedule(ctx, cfg, tm, cam, spp)) return rc;
    claim_cam_table(ctx, cfg, cam);
    CK(launch_render(scene_view(ctx), ctx->is_tri, cam, tm, spp, max_depth, ctx->integrator, seed, sample_offset, resolve,
                     d_out, cfg));
    if (tuned) tune_end(ctx, (cudaStream_t)stream);
    ctx->launches += render_launches(ctx, cfg, spp);
    return 0;
}

int rt_render(rt_ctx* ctx, int width, int height, int spp, int max_depth, uint64_t seed, uint32_t sample_offset,
              float* d_out, void* stream) {
    if (!ctx) return 1;
    return render_frame(ctx, width, height, spp, max_depth, seed, sample_offset, 1, d_out, stream);
}

int rt_render_sum(rt_ctx* ctx, int width, int height, int spp, int max_depth, uint64_t seed, uint32_t sample_offset,
                  float* d_out, void* stream) {
    if (!ctx) return 1;
    return render_frame(ctx, width, height, spp, max_depth, seed, sample_offset, 0, d_out, stream);
}

int rt_resolve(rt_ctx* ctx, const float* d_sum, float* d_out, int64_t n, int spp_total, void* stream) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || spp_total <= 0 || (n > 0 && (!d_sum || !d_out))) return fail(ctx, "rt_resolve: bad arguments");
    DeviceGuard g(ctx->device);
    CK(launch_resolve(d_sum, d_out, n, spp_total, (cudaStream_t)stream));
    if (n) ctx->launches += 1;
    return 0;
}

int rt_resolve_planes(rt_ctx* ctx, const float* d_planes, int n_planes, int64_t plane_stride, float* d_out, int64_t n,
                      int spp_total, void* stream) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || n_planes < 1 || plane_stride < n || spp_total <= 0 || (n > 0 && (!d_planes || !d_out)))
        return fail(ctx, "rt_resolve_planes: bad arguments");
    DeviceGuard g(ctx->device);
    CK(launch_resolve_planes(d_planes, n_planes, plane_stride, d_out, n, spp_total, (cudaStream_t)stream));
    if (n) ctx->launches += 1;
    return 0;
}

int rt_untile(rt_ctx* ctx, int width, int height, int tile_w, int tile_h, int n_ranks, const float* d_tiles, float* d_frame,
              void* stream) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (int rc = check_frame(ctx, width, height)) return rc;
    if (tile_w <= 0 || tile_h <= 0 || n_ranks <= 0 || !d_tiles || !d_frame) return fail(ctx, "rt_untile: bad arguments");
    DeviceGuard g(ctx->device);
    CK(launch_untile(width, height, tile_w, tile_h, n_ranks, d_tiles, d_frame, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

constexpr int kMaxBands = 64;

// Single-launch render with the device->host copy of finished regions overlapped (camera-ray packet
// kernel only): see BandSignal in rt_kernels.h.  A region is a rectangle of 32x32 tiles, copied with one
// cudaMemcpy2DAsync on a second stream as soon as the kernel raises the region's flag.
static int render_host_overlapped(rt_ctx* ctx, int width, int height, int spp, uint64_t seed, uint32_t sample_offset,
                                  float* h_out) {
    if (!ctx->render_stream) CK(cudaStreamCreateWithFlags(&ctx->render_stream, cudaStreamNonBlocking));
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaMalloc(&ctx->d_band_cnt, kMaxBands * sizeof(unsigned int)));
        CK(cudaHostAlloc(&ctx->h_band_flags, kMaxBands * sizeof(unsigned int), cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(&ctx->d_band_flags, ctx->h_band_flags, 0));
    }
    ctx->aspect = (double)width / height;
    CameraBlock cam = camera_block(ctx, ctx->aspect);
    TileMap tm = full_frame_map(width, height);
    ScratchOrder order(ctx, ctx->render_stream);
    LaunchCfg cfg = launch_cfg(ctx, ctx->render_stream, 1);
    BandSignal& bs = cfg.band;
    bs.cnt = ctx->d_band_cnt;
    bs.flags = ctx->d_band_flags;
    bs.tiles_x = tm.tiles_x; bs.tiles_y = tm.n_tiles / tm.tiles_x;
    int want_rows = 4, want_cols = 3;                              // 4 x 3 regions (each copy costs ~3.5 us of DMA idle; measured best on C3)
    if (const char* e = std::getenv("B200RT_REGIONS")) std::sscanf(e, "%d,%d", &want_rows, &want_cols);
    want_rows = std::max(1, std::min(want_rows, 16)); want_cols = std::max(1, std::min(want_cols, 4));
    bs.band_rows = (bs.tiles_y + want_rows - 1) / want_rows;
    bs.group_cols = (bs.tiles_x + want_cols - 1) / want_cols;
    bs.n_groups = (bs.tiles_x + bs.group_cols - 1) / bs.group_cols;
    const int n_regions = ((bs.tiles_y + bs.band_rows - 1) / bs.band_rows) * bs.n_groups;   // <= 48
    if (int rc = attach_schedule(ctx, cfg, tm, cam)) return rc;
    claim_cam_table(ctx, cfg, cam);
    volatile unsigned int* flags = ctx->h_band_flags;
    for (int b = 0; b < n_regions; ++b) flags[b] = 0u;
    CK(cudaMemsetAsync(ctx->d_band_cnt, 0, kMaxBands * sizeof(unsigned int), ctx->render_stream));
    CK(launch_render(scene_view(ctx), ctx->is_tri, cam, tm, spp, 1, ctx->integrator, seed, sample_offset, 1, ctx->d_fb, cfg));
    ctx->launches += render_launches(ctx, cfg);
    static const bool trace = std::getenv("B200RT_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
    const size_t pitch = (size_t)width * 3 * sizeof(float);
    bool copied[kMaxBands] = {false};
    int n_copied = 0;
    bool kernel_over = false;
    while (n_copied < n_regions) {
        bool any = false;
        for (int b = 0; b < n_regions; ++b) {
            if (copied[b] || (!kernel_over && flags[b] == 0u)) continue;
            const int by = b / bs.n_groups, gx = b - by * bs.n_groups;
            const int y0 = by * bs.band_rows * 32, y1 = std::min(height, (by + 1) * bs.band_rows * 32);
            const int x0 = gx * bs.group_cols * 32, x1 = std::min(width, (gx + 1) * bs.group_cols * 32);
            const size_t off = ((size_t)y0 * width + x0) * 3;
            CK(cudaMemcpy2DAsync(h_out + off, pitch, ctx->d_fb + off, pitch, (size_t)(x1 - x0) * 3 * sizeof(float), (size_t)(y1 - y0),
                                 cudaMemcpyDeviceToHost, ctx->copy_stream));
            copied[b] = true; ++n_copied; any = true;
            if (trace) std::fprintf(stderr, "region %d copy issued at %.0f us\n", b, us());
        }
        if (!any && !kernel_over) {
            cudaError_t q = cudaStreamQuery(ctx->render_stream);    // kernel over (or failed): stop polling flags
            if (q != cudaErrorNotReady) {
                if (q != cudaSuccess) return cuda_fail(ctx, "rt_render_host: render", q);
                kernel_over = true;
            }
        }
    }
    CK(cudaStreamSynchronize(ctx->render_stream));
    if (trace) std::fprintf(stderr, "render stream done at %.0f us\n", us());
    CK(cudaStreamSynchronize(ctx->copy_stream));
    if (trace) std::fprintf(stderr, "copy stream done at %.0f us\n", us());
    return 0;
}

// Single-launch render whose kernel stores every finished 32x32 tile straight into the caller's page-locked frame
// (BandSignal tile push, rt_kernels.h): no polling, no DMA; returns when the stream has drained.
// d_host = device alias of h_out.
static int render_host_push(rt_ctx* ctx, int width, int height, int spp, uint64_t seed, uint32_t sample_offset, float* d_host) {
    if (!ctx->render_stream) CK(cudaStreamCreateWithFlags(&ctx->render_stream, cudaStreamNonBlocking));
    ctx->aspect = (double)width / height;
    CameraBlock cam = camera_block(ctx, ctx->aspect);
    TileMap tm = full_frame_map(width, height);
    if (tm.n_tiles > ctx->tile_cnt_cap) {
        cudaFree(ctx->d_tile_cnt); ctx->d_tile_cnt = nullptr; ctx->tile_cnt_cap = 0;
        CK(cudaMalloc(&ctx->d_tile_cnt, (size_t)tm.n_tiles * sizeof(unsigned int)));
        ctx->tile_cnt_cap = tm.n_tiles;
    }
    ScratchOrder order(ctx, ctx->render_stream);
    LaunchCfg cfg = launch_cfg(ctx, ctx->render_stream, 1);
    BandSignal& bs = cfg.band;
    bs.cnt = ctx->d_tile_cnt; bs.flags = nullptr; bs.host_fb = d_host;
    static const char* times_path = std::getenv("B200RT_PUSH_TIMES");
    static unsigned long long* d_times = nullptr;
    if (times_path) {
        if (!d_times) CK(cudaMalloc(&d_times, (size_t)1 << 20));
        CK(cudaMemsetAsync(d_times, 0, (size_t)tm.n_tiles * 8, ctx->render_stream));
        bs.push_times = d_times;
    }
    bs.tiles_x = tm.tiles_x; bs.tiles_y = tm.n_tiles / tm.tiles_x;
    bs.band_rows = 1; bs.group_cols = 1; bs.n_groups = tm.tiles_x;          // one region per tile, raster due order
    if (int rc = attach_schedule(ctx, cfg, tm, cam)) return rc;
    claim_cam_table(ctx, cfg, cam);
    const auto t0 = std::chrono::steady_clock::now();
    CK(cudaMemsetAsync(ctx->d_tile_cnt, 0, (size_t)tm.n_tiles * sizeof(unsigned int), ctx->render_stream));
    CK(launch_render(scene_view(ctx), ctx->is_tri, cam, tm, spp, 1, ctx->integrator, seed, sample_offset, 1, ctx->d_fb, cfg));
    ctx->launches += render_launches(ctx, cfg);
    CK(cudaStreamSynchronize(ctx->render_stream));
    if (times_path) {                                                        // debug dump: tile, ns since the first push
        std::vector<unsigned long long> t((size_t)tm.n_tiles);
        cudaMemcpy(t.data(), d_times, t.size() * 8, cudaMemcpyDeviceToHost);
        if (FILE* f = std::fopen(times_path, "w")) {
            unsigned long long t_min = ~0ull;
            for (auto v : t) if (v && v < t_min) t_min = v;
            for (size_t k = 0; k < t.size(); ++k) std::fprintf(f, "%zu %lld\n", k, t[k] ? (long long)(t[k] - t_min) : -1ll);
            std::fclose(f);
        }
    }
    static const bool trace = std::getenv("B200RT_TRACE") != nullptr;
    if (trace) std::fprintf(stderr, "tile push: launch..drain %.0f us\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    return 0;
}

// Any other frame into a PAGE-LOCKED host buffer: rendered as a few bands of tile rows (separate launches, frame layout),
// each band's rows copied by the DMA engine on a second stream while the next band renders -- what the tile push does for
// camera-ray frames, at band granularity, for the call the reference's host really makes (render(W, H, 8, 4),
// interaction.py:1294-1298: multi-bounce, several samples).
static int render_host_banded(rt_ctx* ctx, int width, int height, int spp, int max_depth, uint64_t seed, uint32_t sample_offset,
                              float* h_out) {
    if (!ctx->render_stream) CK(cudaStreamCreateWithFlags(&ctx->render_stream, cudaStreamNonBlocking));
    if (!ctx->band_stream) CK(cudaStreamCreateWithFlags(&ctx->band_stream, cudaStreamNonBlocking));
    constexpr int kBands = 4;
    for (int b = 0; b < kBands; ++b) if (!ctx->band_ev[b]) CK(cudaEventCreateWithFlags(&ctx->band_ev[b], cudaEventDisableTiming));
    const int tiles_x = (width + 31) / 32, tiles_y = (height + 31) / 32;
    const int rows = (tiles_y + kBands - 1) / kBands;
    for (int b = 0; b * rows < tiles_y; ++b) {
        const int ty0 = b * rows, nrows = std::min(rows, tiles_y - ty0);
        if (int rc = render_tiles(ctx, width, height, 32, 32, ty0 * tiles_x, 1, spp, max_depth, seed, sample_offset, 1, ctx->d_fb,
                                  ctx->render_stream, 2, nrows * tiles_x)) return rc;
        CK(cudaEventRecord(ctx->band_ev[b], ctx->render_stream));
        CK(cudaStreamWaitEvent(ctx->band_stream, ctx->band_ev[b], 0));
        const int y0 = ty0 * 32, y1 = std::min(height, (ty0 + nrows) * 32);
        const size_t off = (size_t)y0 * width * 3;
        CK(cudaMemcpyAsync(h_out + off, ctx->d_fb + off, (size_t)(y1 - y0) * width * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->band_stream));
    }
    CK(cudaStreamSynchronize(ctx->band_stream));
    return 0;
}

static bool is_page_locked(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// device alias of a page-locked host buffer the SMs can store into with 16-byte vectors, or nullptr
static float* pushable_alias(float* h_out) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h_out) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
    if (((uintptr_t)at.devicePointer & 15u) != 0) return nullptr;
    return static_cast<float*>(at.devicePointer);
}

int rt_render_host(rt_ctx* ctx, int width, int height, int spp, int max_depth, uint64_t seed, uint32_t sample_offset,
                   float* h_out) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (int rc = check_frame(ctx, width, height)) return rc;
    if (!h_out) return fail(ctx, "rt_render_host: NULL output");
    if (spp <= 0 || max_depth < 0) return fail(ctx, "rt_render_host: bad arguments");
    DeviceGuard g(ctx->device);
    size_t need = (size_t)width * height * 3;
    if (need > ctx->fb_floats) {
        cudaFree(ctx->d_fb); ctx->d_fb = nullptr; ctx->fb_floats = 0;
        CK(cudaMalloc(&ctx->d_fb, need * sizeof(float)));
        ctx->fb_floats = need;
    }
    if (int rc = ensure_device(ctx)) return rc;
    if (ctx->overlap && pick_kernel(ctx, max_depth) == 3 && max_depth == 1 && spp == 1 && ctx->n > 0 && need >= ((size_t)1 << 18)) {
        CK(cudaStreamSynchronize(nullptr));               // order after earlier work of the legacy stream
        if (ctx->overlap >= 2)
            if (float* alias = pushable_alias(h_out)) { ctx->host_path = 2; return render_host_push(ctx, width, height, spp, seed, sample_offset, alias); }
        ctx->host_path = 1;
        return render_host_overlapped(ctx, width, height, spp, seed, sample_offset, h_out);
    }
    // frames of >= 4 MB into page-locked memory: bands, copy overlapped (tiny scenes first let rt_render time its two kernel
    // candidates on whole frames)
    if (ctx->overlap && max_depth >= 1 && ctx->n > 0 && need * sizeof(float) >= ((size_t)4 << 20) && height >= 256 &&
        !(tunes(ctx, max_depth) && ctx->tune_state < 2) && is_page_locked(h_out)) {
        CK(cudaStreamSynchronize(nullptr));
        ctx->host_path = 3;
        return render_host_banded(ctx, width, height, spp, max_depth, seed, sample_offset, h_out);
    }
    ctx->host_path = 0;
    if (int rc = rt_render(ctx, width, height, spp, max_depth, seed, sample_offset, ctx->d_fb, nullptr)) return rc;
    CK(cudaMemcpy(h_out, ctx->d_fb, need * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

int rt_accumulate(rt_ctx* ctx, const float* d_batch, float* d_accum, int64_t n, int n_old, int n_batch, void* stream) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || n_old < 0 || n_batch <= 0 || (n > 0 && (!d_batch || !d_accum))) return fail(ctx, "rt_accumulate: bad arguments");
    DeviceGuard g(ctx->device);
    CK(launch_accumulate(d_batch, d_accum, n, n_old, n_batch, (cudaStream_t)stream));
    if (n) ctx->launches += 1;
    return 0;
}

int rt_tonemap_u8(rt_ctx* ctx, const float* d_accum, uint8_t* d_rgb8, int64_t n, float exposure, void* stream) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || (n > 0 && (!d_accum || !d_rgb8))) return fail(ctx, "rt_tonemap_u8: bad arguments");
    DeviceGuard g(ctx->device);
    CK(launch_tonemap_u8(d_accum, d_rgb8, n, exposure, (cudaStream_t)stream));
    if (n) ctx->launches += 1;
    return 0;
}

int rt_display_u8(rt_ctx* ctx, const float* d_accum, uint8_t* d_rgb8, int64_t n, float exposure, int enhance, void* stream) {
    if (!ctx) return 1;
    if (!enhance) return rt_tonemap_u8(ctx, d_accum, d_rgb8, n, exposure, stream);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || n > 0x7fffffff || (n > 0 && (!d_accum || !d_rgb8))) return fail(ctx, "rt_display_u8: bad arguments");
    DeviceGuard g(ctx->device);
    const size_t need = display_scratch_bytes(n);
    if (need > ctx->display_bytes) {
        cudaFree(ctx->d_display); ctx->d_display = nullptr; ctx->display_bytes = 0;
        CK(cudaMalloc(&ctx->d_display, need));
        ctx->display_bytes = need;
    }
    int nl = 0;
    ScratchOrder order(ctx, stream);
    CK(launch_display_u8(d_accum, d_rgb8, n, exposure, ctx->d_display, ctx->display_bytes, (cudaStream_t)stream, &nl));
    ctx->launches += nl;
    return 0;
}

int rt_set_option(rt_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    std::string k(name);
    if (k == "integrator") { if (value != 0 && value != 1) return fail(ctx, "integrator must be 0 (v1) or 1 (v2)"); ctx->integrator = (int)value; }
    else if (k == "stats") ctx->stats = value != 0;
    else if (k == "kernel") { if (value < -1 || value > 5) return fail(ctx, "kernel must be -1 (auto), 0 (k_path), 1 (simple megakernel), 2 (wavefront), 3 (camera-ray packets), 4 (wavefront with packet bounce 0) or 5 (tiny scenes: whole scene in shared memory; other scenes fall back to auto)"); ctx->kernel = (int)value; }
    else if (k == "leaf_vote") { if (value < 1 || value > 32) return fail(ctx, "leaf_vote must be in 1..32"); ctx->leaf_vote = (int)value; ctx->leaf_vote_set = true; }
    else if (k == "overlap") { if (value < 0 || value > 2) return fail(ctx, "overlap must be 0 (render, then copy), 1 (region flags + DMA copies) or 2 (tile push)"); ctx->overlap = (int)value; }
    else if (k == "refit_limit") { if (value < 0 || value > 100000) return fail(ctx, "refit_limit must be 0 (never rebuild) or a percentage"); ctx->refit_limit = (int)value; }
    else if (k == "builder") { if (value < 0 || value > 2) return fail(ctx, "builder must be 0 (reference median split, host), 1 (LBVH, device) or 2 (binned SAH, host)"); ctx->builder = (int)value; }
    else if (k == "fold") ctx->fold = value != 0;
    else if (k == "treelet") { if (value < 0 || value > 10) return fail(ctx, "treelet must be 0 (off) or 1..10 levels"); ctx->treelet_levels = (int)value; ctx->treelet_valid = false; }
    else if (k == "qnodes") { if (value < -1 || value > 5) return fail(ctx, "qnodes must be -1 (auto), 0 (off), 1 (compressed sibling pairs), 2 (split triangle records), 3 (both), 4 (cooperative leaf step) or 5 (compressed pairs + cooperative leaf step)"); ctx->qmode = (int)value; ctx->qnodes_valid = false; }
    else if (k == "qnodes_area_limit") { if (value < 100 || value > 100000) return fail(ctx, "qnodes_area_limit must be a percentage >= 100"); ctx->q_area_limit = (int)value; ctx->qnodes_valid = false; }
    else if (k == "sah_cost") { if (value < 0 || value > 1000) return fail(ctx, "sah_cost must be in 0..1000 (tenths of a primitive test)"); ctx->sah_cost = (int)value; }
    else if (k == "leaf_size") { if (value < 1 || value > 4) return fail(ctx, "leaf_size must be in 1..4"); ctx->leaf_size = (int)value; }
    else if (k == "schedule") { ctx->schedule = value != 0; ctx->chunk_key = -1; }
    else if (k == "block_times") ctx->d_block_times = reinterpret_cast<unsigned long long*>((uintptr_t)value);
    else if (k == "wf_streams") { if (value != 1 && value != 2) return fail(ctx, "wf_streams must be 1 (one wave at a time) or 2 (two waves in flight)"); ctx->wf_streams = (int)value; }
    else if (k == "wf_rays_per_lane") { if (value < 0 || value > 1024) return fail(ctx, "wf_rays_per_lane must be in 0..1024"); ctx->wf_rays_per_lane = (int)value; }
    else if (k == "tiny_mode") { if (value != 0 && value != 1) return fail(ctx, "tiny_mode must be 0 (CTA-local wavefront) or 1 (lock step)"); ctx->tiny_mode = (int)value; }
    else if (k == "tiny_threads") { if (value != 128 && value != 256) return fail(ctx, "tiny_threads must be 128 or 256"); ctx->tiny_threads = (int)value; }
    else if (k == "refill") { if (value < 1 || value > 32) return fail(ctx, "refill must be in 1..32"); ctx->refill = (int)value; ctx->refill_set = true; }
    else return fail(ctx, "rt_set_option: unknown option '" + k + "'");
    return 0;
}

int rt_get_option(rt_ctx* ctx, const char* name, int64_t* value) {
    if (!ctx || !name || !value) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    std::string k(name);
    if (k == "integrator") *value = ctx->integrator;
    else if (k == "stats") *value = ctx->stats;
    else if (k == "kernel") *value = ctx->kernel;
    else if (k == "kernel_used") *value = ctx->kernel_used;
    else if (k == "refill") *value = ctx->refill;
    else if (k == "tiny_threads") *value = ctx->tiny_threads;
    else if (k == "tiny_mode") *value = ctx->tiny_mode;
    else if (k == "wf_rays_per_lane") *value = ctx->wf_rays_per_lane;
    else if (k == "wf_streams") *value = ctx->wf_streams;
    else if (k == "overlap") *value = ctx->overlap;
    else if (k == "host_path") *value = ctx->host_path;
    else if (k == "schedule") *value = ctx->schedule;
    else if (k == "leaf_vote") *value = ctx->leaf_vote;
    else if (k == "sm_count") *value = ctx->sm_count;
    else if (k == "bvh_depth") *value = ctx->bvh_depth;
    else if (k == "n_prims") *value = ctx->n;
    else if (k == "n_nodes") *value = ctx->n_nodes;
    else if (k == "builder") *value = ctx->builder;
    else if (k == "leaf_size") *value = ctx->leaf_size;
    else if (k == "sah_cost") *value = ctx->sah_cost;
    else if (k == "treelet") *value = ctx->treelet_levels;
    else if (k == "qnodes") *value = ctx->qmode;
    else if (k == "qnodes_used") *value = ctx->qnodes_valid ? ctx->qmode_used : 0;
    else if (k == "qnodes_area_pct") *value = ctx->q_area_pct;
    else if (k == "qnodes_area_limit") *value = ctx->q_area_limit;
    else if (k == "fold") *value = ctx->fold;
    else if (k == "refit_limit") *value = ctx->refit_limit;
    else if (k == "refits") *value = ctx->refits;
    else if (k == "refit_rebuilds") *value = ctx->refit_rebuilds;
    else if (k == "refit_area_pct") *value = ctx->area_pct;
    else return fail(ctx, "rt_get_option: unknown option '" + k + "'");
    return 0;
}

int rt_get_stats(rt_ctx* ctx, rt_stats* out) {
    if (!ctx || !out) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    unsigned long long v[4];
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(v, ctx->d_stats, sizeof(v), cudaMemcpyDeviceToHost));
    out->rays = v[0]; out->segments = v[1]; out->node_records = v[2]; out->prim_tests = v[3];
    out->launches = ctx->launches;
    return 0;
}

int rt_reset_stats(rt_ctx* ctx) {
    if (!ctx) return 1;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    CK(cudaMemset(ctx->d_stats, 0, 64));
    ctx->launches = 0;
    return 0;
}

}  // extern "C"
