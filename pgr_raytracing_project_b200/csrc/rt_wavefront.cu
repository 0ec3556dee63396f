// rt_wavefront.cu -- wavefront form of the render hot path (kernel variant 2, sm_100a).
//
// A frame is processed in waves of paths (pixel task x sample).  Per wave:
//   k_wf_generate     camera ray per path            -> ray queue 0 (compacted, warp-aggregated append)
//   per bounce b:
//     k_wf_trace      persistent threads; every lane pulls the next ray of queue b as soon as its own
//                     traversal is finished (one atomic per refill, lanes ranked by ballot/popc), so the
//                     warps stay full while the walk per ray is exactly intersect()'s -> hit records; the first
//                     kWfSmemLevels levels of every lane's traversal stack live in shared memory (HybridStack)
//     k_wf_shade      one thread per ray of queue b: emission / background, Russian roulette, scatter;
//                     surviving paths are appended to queue b+1 (ballot/popc compaction)
//   k_wf_accumulate   per pixel: add the wave's samples IN SAMPLE ORDER to the running sum (so the
//                     frame is bit-identical to the megakernels'), resolve after the last wave
// Ray / hit / path state live in HBM as float4 SoA (112 B per path in flight); at 2 M paths that is
// ~0.25 GB per bounce of streaming traffic, i.e. tens of microseconds at B200 HBM bandwidth, in
// exchange for full warps in both the traversal and the shading kernels.
#include "rt_kernel_common.cuh"

namespace b200rt {

namespace {

#ifndef B200RT_WF_SMEM_LEVELS
#define B200RT_WF_SMEM_LEVELS 16
#endif
#ifndef B200RT_QTRACE_MIN_CTAS
#define B200RT_QTRACE_MIN_CTAS 10    // resident CTAs per SM the "qnodes" instances of k_wf_trace are compiled for: a 48-register cap (left alone they take 56-63 and run slower than the full records)
#endif
constexpr int kWfSmemLevels = B200RT_WF_SMEM_LEVELS;   // traversal-stack levels of k_wf_trace kept in shared memory (HybridStack)

struct WaveArgs {
    TileMap tm;
    int task0, n_tasks_wave;      // pixel tasks [task0, task0 + n_tasks_wave) of the launch's enumeration
    int sample0, batch;           // samples [sample0, sample0 + batch) of every pixel in this wave
    int n_paths;                  // n_tasks_wave * batch
    int spp, max_depth, integrator;
    uint32_t k0, k1, sample_offset;
    int resolve, last_wave;
};

__device__ __forceinline__ unsigned warp_append(unsigned int* counter, bool pred, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return 0u;
    int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (unsigned)__popc(m & ((1u << lane) - 1u));
}

template <bool STATS, bool AOV>
__global__ void __launch_bounds__(256)
k_wf_generate(const __grid_constant__ CameraBlock cam, const __grid_constant__ WaveArgs wa, const __grid_constant__ WaveBuffers wb,
              unsigned long long* d_stats) {
    const int lane = threadIdx.x & 31;
    const double inv_w = __ddiv_rn(1.0, (double)wa.tm.width), inv_h = __ddiv_rn(1.0, (double)wa.tm.height);
    unsigned long long rays = 0;
    const int n_round = (wa.n_paths + 31) & ~31;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_round; slot += gridDim.x * blockDim.x) {
        bool valid = slot < wa.n_paths;
        Ray r;
        if (valid) {
            int task = wa.task0 + slot / wa.batch, sb = slot - (slot / wa.batch) * wa.batch;
            PixelWork p = decode_work(wa.tm, task >> 5, task & 31);
            valid = p.active;
            if (valid) {
                float jx = 0.5f, jy = 0.5f;
                if (!AOV) {
                    uint32_t pixel = (uint32_t)(p.j * wa.tm.width + p.i);
                    uint4 ctl = philox4x32_10(pixel, wa.sample_offset + (uint32_t)(wa.sample0 + sb), 0u, 0u, wa.k0, wa.k1);
                    jx = u01(ctl.x); jy = u01(ctl.y);
                }
                r = camera_ray(cam, p.i, p.j, jx, jy, inv_w, inv_h);
                if (STATS) rays += 1;
            }
        }
        unsigned q = warp_append(wb.counters, valid, lane);
        if (valid) {
            wb.ray_o[0][q] = make_float4(r.ox, r.oy, r.oz, __int_as_float(slot));
            wb.ray_d[0][q] = make_float4(r.dx, r.dy, r.dz, 0.0f);
            if (!AOV) {
                wb.path_thr[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
                wb.path_rad[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
    }
    if (STATS) { Counters c = {0, 0, 0}; flush_stats(d_stats, rays, c); }
}

// Bounce 0 of a wave for the hybrid variant 4: camera rays are coherent, so generation and the first
// trace are ONE kernel that walks each 8x4 pixel block as a packet (packet_intersect, the k_packet
// machinery incl. the CTA-chunked work distribution) -- once per sample of the wave's batch -- and writes
// ray + closest hit of every path to queue 0, ready for k_wf_shade.  Bounces >= 1 are incoherent and
// stay with the per-lane k_wf_trace.  Same bits as k_wf_generate + k_wf_trace(0).
template <bool TRI, bool STATS>
__global__ void __launch_bounds__(kPacketThreads)
k_wf_packet0(const __grid_constant__ SceneView sc, const float4* __restrict__ cam_prims, const __grid_constant__ CameraBlock cam,
             const __grid_constant__ WaveArgs wa, const __grid_constant__ WaveBuffers wb, unsigned int* chunk_counter,
             unsigned long long* d_stats) {
    __shared__ uint2 s_stack[kPacketThreads / 32][kStackDepth];
    __shared__ unsigned s_word;
    const int lane = threadIdx.x & 31;
    uint2* stack = s_stack[threadIdx.x >> 5];
    // work item = (block, sample of the wave's batch), the samples of a block being consecutive items: the 8 warps of a CTA
    // trace one block's samples side by side (shared L1 lines), and a wave of few blocks and many samples -- one rank's
    // tiles of a multi-GPU frame -- still splits into enough items to balance (a block with its sample loop inside left
    // 1.4 items per resident warp at 8 ranks)
    const int n_blocks = wa.n_tasks_wave >> 5, block0 = wa.task0 >> 5;
    const int n_items = n_blocks * wa.batch;
    const int n_chunks = (n_items + kChunk - 1) / kChunk;
    if (threadIdx.x == 0) s_word = take_chunk(chunk_counter, n_chunks, nullptr);
    __syncthreads();
    const double inv_w = __ddiv_rn(1.0, (double)wa.tm.width), inv_h = __ddiv_rn(1.0, (double)wa.tm.height);
    Counters cnt = {0, 0, 0};
    unsigned long long rays = 0;
    for (;;) {
        const int item = chunk_next_block(&s_word, chunk_counter, n_chunks, nullptr, lane);
        if (item < 0) break;
        if (item >= n_items) continue;
        const int wl = item / wa.batch, sb = item - wl * wa.batch;
        PixelWork p = decode_work(wa.tm, block0 + wl, lane);
        const uint32_t pixel = (uint32_t)(p.j * wa.tm.width + p.i);
        uint4 ctl = philox4x32_10(pixel, wa.sample_offset + (uint32_t)(wa.sample0 + sb), 0u, 0u, wa.k0, wa.k1);
        Ray r = camera_ray(cam, p.i, p.j, u01(ctl.x), u01(ctl.y), inv_w, inv_h);
        Hit h;
        int work = 0;
        packet_intersect<TRI, STATS>(sc, cam_prims, r, p.active, lane, stack, h, cnt, work);
        if (STATS && p.active) rays += 1;
        const int slot = (wl * 32 + lane) * wa.batch + sb;
        const unsigned q = warp_append(wb.counters, p.active, lane);
        if (p.active) {
            wb.ray_o[0][q] = make_float4(r.ox, r.oy, r.oz, __int_as_float(slot));
            wb.ray_d[0][q] = make_float4(r.dx, r.dy, r.dz, 0.0f);
            wb.hit[q] = make_float4(h.t, __int_as_float(h.prim), __int_as_float(h.slot), 0.0f);
            wb.path_thr[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            wb.path_rad[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
    if (STATS) flush_stats(d_stats, rays, cnt);
}

// CAM: the rays of this queue are camera rays (bounce 0 of variant 2: triangle test from the per-camera table)
// QM: option "qnodes" (bounces >= 1 only): bit 0 = compressed sibling pairs, bit 1 = split triangle records, bit 2 = cooperative leaf step (trav_run)
template <bool TRI, bool STATS, bool CAM, bool TREELET = false, int QM = 0>
__global__ void __launch_bounds__(kThreads, QM ? B200RT_QTRACE_MIN_CTAS : 0)
k_wf_trace(const __grid_constant__ SceneView sc, const __grid_constant__ WaveBuffers wb, int bounce, int max_depth, int refill_below,
           int leaf_vote, unsigned long long* d_stats, int rays_per_lane) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float4* __restrict__ ray_o = wb.ray_o[bounce & 1];
    const float4* __restrict__ ray_d = wb.ray_d[bounce & 1];
    const unsigned count = wb.counters[bounce];
    // Small queues (one rank's share of a multi-GPU frame, deep bounces): the grid is sized for the machine, not for the
    // queue, and with one or two rays per lane the refill has nothing to even out -- the kernel then lasts as long as its
    // unluckiest warp.  CTAs beyond what gives every lane `rays_per_lane` rays leave at once (never fewer than gridDim / 8).
    if (rays_per_lane > 0) {
        const unsigned want = (count + (unsigned)(kThreads * rays_per_lane) - 1u) / (unsigned)(kThreads * rays_per_lane);
        const unsigned keep = want > gridDim.x / 8u ? want : gridDim.x / 8u;
        if (blockIdx.x >= keep) return;
    }
    unsigned int* fetch = wb.counters + (max_depth + 1) + bounce;
    __shared__ uint2 s_stack[kWfSmemLevels][kThreads];
    __shared__ CoopWarp s_coop[(QM & 4) ? kThreads / 32 : 1];   // cooperative leaf step: per-warp ray copies, pair table, best keys
    unsigned coop_sa = (unsigned)__cvta_generic_to_shared(&s_coop[(QM & 4) ? (threadIdx.x >> 5) : 0]);
    if (QM & 4) asm volatile("" : "+r"(coop_sa));            // one live register instead of re-deriving the address at every use
    extern __shared__ float4 s_tree[];                         // TREELET: top levels of the tree (SceneView::treelet), option "treelet"
    const int two_t = TREELET ? sc.treelet_two_t : 0;
    if (TREELET) {
        for (int k = threadIdx.x; k < 2 * two_t; k += kThreads) s_tree[k] = __ldg(sc.treelet + k);
        __syncthreads();
    }
    int stack_code[kStackDepth - kWfSmemLevels];
    float stack_tn[kStackDepth - kWfSmemLevels];
    typedef HybridStack<kWfSmemLevels, kThreads> StackT;
    const StackT stack{&s_stack[0][threadIdx.x], stack_code, stack_tn};
    Trav tv;
    tv.cur = kDone; tv.sp = 0; tv.h.t = kTMax; tv.h.prim = -1; tv.h.slot = -1;
    Counters cnt = {0, 0, 0};
    Ray r = make_ray(0.f, 0.f, 0.f, 0.f, 0.f, 1.f);
    QRay qr = {};
    int q = -1;
    bool pool_empty = false;
    for (;;) {
        bool idle = tv.cur == kDone;
        if (idle && q >= 0) {                       // publish the finished ray's closest hit
            wb.hit[q] = make_float4(tv.h.t, __int_as_float(tv.h.prim), __int_as_float(tv.h.slot), 0.0f);
            q = -1;
        }
        unsigned need = __ballot_sync(0xffffffffu, idle);
        if (need != 0u && !pool_empty) {
            int n_need = __popc(need), leader = __ffs(need) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(fetch, (unsigned)n_need);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (idle) {
                unsigned idx = base + (unsigned)__popc(need & lt_mask);
                if (idx < count) {
                    q = (int)idx;
                    float4 o = __ldg(ray_o + idx), d = __ldg(ray_d + idx);
                    r = make_ray(o.x, o.y, o.z, d.x, d.y, d.z);
                    if (QM & 1) qr = make_qray(sc.qgrid, r);
                    if (QM & 4) coop_store_ray(coop_sa, lane, o, d);
                    trav_begin<STATS>(sc, r, tv, cnt, two_t);
                }
            }
            if (base + (unsigned)n_need >= count) pool_empty = true;
        }
        unsigned act = __ballot_sync(0xffffffffu, tv.cur != kDone);
        if (act == 0u) {
            if (pool_empty && __ballot_sync(0xffffffffu, q >= 0) == 0u) break;
            continue;                               // root misses waiting to be published / more to fetch
        }
        int min_active = pool_empty ? 1 : (__popc(act) * refill_below) >> 5;
        trav_run<TRI, STATS, CAM ? 1 : 0, StackT, TREELET, QM>(sc, r, tv, stack, min_active < 1 ? 1 : min_active, leaf_vote,
                                                                                              cnt, CAM, s_tree, two_t, &qr, coop_sa, lane);
    }
    if (STATS) flush_stats(d_stats, 0, cnt);
}

template <bool TRI, bool STATS, bool AOV>
__global__ void __launch_bounds__(256)
k_wf_shade(const __grid_constant__ SceneView sc, const __grid_constant__ WaveArgs wa, const __grid_constant__ WaveBuffers wb, int bounce,
           int32_t* __restrict__ d_prim, float* __restrict__ d_t, unsigned long long* d_stats) {
    const int lane = threadIdx.x & 31;
    const float4* __restrict__ ray_o = wb.ray_o[bounce & 1];
    const float4* __restrict__ ray_d = wb.ray_d[bounce & 1];
    float4* __restrict__ next_o = wb.ray_o[(bounce + 1) & 1];
    float4* __restrict__ next_d = wb.ray_d[(bounce + 1) & 1];
    const unsigned count = wb.counters[bounce];
    const unsigned n_round = (count + 31u) & ~31u;
    Counters cnt = {0, 0, 0};
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n_round; q += gridDim.x * blockDim.x) {
        bool alive = false;
        Ray r;
        int slot = 0;
        if (q < count) {
            float4 o = ray_o[q], d = ray_d[q], hh = wb.hit[q];
            slot = __float_as_int(o.w);
            Hit h;
            h.t = hh.x; h.prim = __float_as_int(hh.y); h.slot = __float_as_int(hh.z);
            int task = wa.task0 + slot / wa.batch, sb = slot - (slot / wa.batch) * wa.batch;
            PixelWork p = decode_work(wa.tm, task >> 5, task & 31);
            if (STATS) cnt.segments += 1;
            if (AOV) {
                d_prim[p.out_index] = h.prim;
                d_t[p.out_index] = h.prim >= 0 ? h.t : 0.0f;
            } else {
                float4 thr = wb.path_thr[slot], rad = wb.path_rad[slot];
                if (h.prim < 0) {
                    rad.x = __fmaf_rn(thr.x, sc.bg_r, rad.x); rad.y = __fmaf_rn(thr.y, sc.bg_g, rad.y); rad.z = __fmaf_rn(thr.z, sc.bg_b, rad.z);
                } else {
                    const float4* mp = sc.mats + 2 * (size_t)material_row<TRI>(sc, h);
                    float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
                    rad.x = __fmaf_rn(thr.x, m1.y, rad.x); rad.y = __fmaf_rn(thr.y, m1.z, rad.y); rad.z = __fmaf_rn(thr.z, m1.w, rad.z);
                    if (bounce + 1 < wa.max_depth) {
                        uint32_t pixel = (uint32_t)(p.j * wa.tm.width + p.i);
                        uint32_t sample = wa.sample_offset + (uint32_t)(wa.sample0 + sb);
                        uint4 ctl = philox4x32_10(pixel, sample, (uint32_t)bounce, 0u, wa.k0, wa.k1);
                        r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                        alive = scatter<TRI>(sc, h, r, wa.integrator, bounce, wa.max_depth, ctl, m0, m1, pixel, sample,
                                             wa.k0, wa.k1, thr.x, thr.y, thr.z);
                        if (alive) wb.path_thr[slot] = thr;
                    }
                }
                wb.path_rad[slot] = rad;
            }
        }
        unsigned nq = warp_append(wb.counters + bounce + 1, alive, lane);
        if (alive) {
            next_o[nq] = make_float4(r.ox, r.oy, r.oz, __int_as_float(slot));
            next_d[nq] = make_float4(r.dx, r.dy, r.dz, 0.0f);
        }
    }
    if (STATS) flush_stats(d_stats, 0, cnt);
}

__global__ void __launch_bounds__(256)
k_wf_accumulate(const __grid_constant__ WaveArgs wa, const __grid_constant__ WaveBuffers wb, float* __restrict__ d_out) {
    const float inv_spp = __fdiv_rn(1.0f, (float)wa.spp);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < wa.n_tasks_wave; k += gridDim.x * blockDim.x) {   // whole warps
        int task = wa.task0 + k;
        PixelWork p = decode_work(wa.tm, task >> 5, task & 31);
        float* o = d_out + 3 * (size_t)p.out_index;
        float sr = 0.0f, sg = 0.0f, sb = 0.0f;
        if (p.active) {
            if (wa.sample0 > 0) { sr = o[0]; sg = o[1]; sb = o[2]; }
            for (int s = 0; s < wa.batch; ++s) {
                float4 rad = wb.path_rad[(size_t)k * wa.batch + s];
                sr = __fadd_rn(sr, rad.x); sg = __fadd_rn(sg, rad.y); sb = __fadd_rn(sb, rad.z);
            }
            if (wa.last_wave && wa.resolve) { sr = resolve1(sr, inv_spp); sg = resolve1(sg, inv_spp); sb = resolve1(sb, inv_spp); }
        }
        warp_store_rgb(o, p.active, sr, sg, sb, task & 31);
    }
}

int grid_for(int64_t n, int threads, int sm_count, int per_sm) {
    int64_t g = (n + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count * per_sm;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// Option "qnodes": the k_wf_trace instance for the mode the context settled on (LaunchCfg::qmode; rt_api.cu ensure_qnodes).
// false = no such instance applies (the caller launches the full-record kernel).
template <bool TRI, int M>
void launch_qtrace_mode(const SceneView& sc, const WaveBuffers& wb, int b, int max_depth, const LaunchCfg& cfg, cudaStream_t st) {
    auto kern = k_wf_trace<TRI, false, false, false, M>;
    static const int per_sm = [] {
        int n = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_wf_trace<TRI, false, false, false, M>, kThreads, 0);
        return n < 1 ? 1 : n;
    }();
    kern<<<cfg.sm_count * per_sm, kThreads, 0, st>>>(sc, wb, b, max_depth, (M & 4) ? cfg.coop_refill : cfg.refill_below,
                                                   (M & 4) ? cfg.coop_leaf_vote : cfg.leaf_vote, nullptr, cfg.wf_rays_per_lane);
}
template <bool TRI>
bool launch_qtrace(const SceneView& sc, const WaveBuffers& wb, int b, int max_depth, const LaunchCfg& cfg, cudaStream_t st) {
    if (cfg.qmode == 0 || sc.qnodes == nullptr) return false;
    int qm = TRI ? cfg.qmode : (cfg.qmode & 1);              // spheres: the pairs only
    if (sc.tri_a == nullptr) qm &= ~2;
    if (qm == 1) launch_qtrace_mode<TRI, 1>(sc, wb, b, max_depth, cfg, st);
    else if constexpr (TRI) {
        if (qm == 2) launch_qtrace_mode<true, 2>(sc, wb, b, max_depth, cfg, st);
        else if (qm == 3) launch_qtrace_mode<true, 3>(sc, wb, b, max_depth, cfg, st);
        else if (qm == 4) launch_qtrace_mode<true, 4>(sc, wb, b, max_depth, cfg, st);
        else if (qm == 5) launch_qtrace_mode<true, 5>(sc, wb, b, max_depth, cfg, st);
        else return false;
    } else return false;
    return true;
}

template <bool TRI, bool STATS, bool AOV>
cudaError_t run_wavefront(const SceneView& sc, const CameraBlock& cam, const TileMap& tm, int spp, int max_depth,
                          int integrator, uint64_t seed, uint32_t sample_offset, int resolve, float* d_out,
                          int32_t* d_prim, float* d_t, const LaunchCfg& cfg, const WaveBuffers& wb0, int* n_launches, const WavePipe* pipe) {
    const bool packet0 = !AOV && cfg.variant == 4;             // bounce 0 by camera-ray packets
    if (TRI && !cfg.cam_table_valid) {                         // bounce 0 (camera rays) reads the camera-relative table
        cudaError_t e = launch_cam_tris(sc, cam, cfg);
        if (e != cudaSuccess) return e;
        *n_launches += 1;
    }
    int packet_grid = 0;
    if (packet0) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wf_packet0<TRI, STATS>, kPacketThreads, 0);
        packet_grid = cfg.sm_count * (per_sm < 1 ? 1 : per_sm);
    }
    const int n_tasks = work_items(tm) * 32;
    const int cap = wb0.capacity;
    const bool two = pipe != nullptr && pipe->wave2 != nullptr && !AOV && (int64_t)n_tasks * spp >= 65536;
    int chunk = n_tasks < cap ? n_tasks : (cap & ~31);
    if (two && spp == 1 && chunk == n_tasks) chunk = (((n_tasks + 1) / 2) + 31) & ~31;         // at least two waves: split the pixels ...
    const int trace_grid = resident_grid(k_wf_trace<TRI, STATS, false>, cfg.sm_count);
    if (two) {                                                 // fork: both internal streams start after what precedes on cfg.stream
        cudaError_t e = cudaEventRecord(pipe->fork, cfg.stream);
        if (e != cudaSuccess) return e;
        cudaStreamWaitEvent(pipe->streams[0], pipe->fork, 0);
        cudaStreamWaitEvent(pipe->streams[1], pipe->fork, 0);
    }
    int wave = 0;
    bool acc_pending[2] = {false, false};
    for (int task0 = 0; task0 < n_tasks; task0 += chunk) {
        int nt = n_tasks - task0 < chunk ? n_tasks - task0 : chunk;
        int batch_max = cap / nt;
        if (batch_max < 1) batch_max = 1;
        if (two && spp >= 2 && batch_max > (spp + 1) / 2) batch_max = (spp + 1) / 2;             // ... or the samples
        for (int s0 = 0; s0 < spp; ++wave) {
            const int set = two ? (wave & 1) : 0;
            const WaveBuffers& wb = set ? *pipe->wave2 : wb0;
            cudaStream_t st = two ? pipe->streams[set] : cfg.stream;
            unsigned int* chunk_counter = two ? pipe->counters[set] : cfg.d_work_counter;
            int batch = spp - s0 < batch_max ? spp - s0 : batch_max;
            WaveArgs wa;
            wa.tm = tm; wa.task0 = task0; wa.n_tasks_wave = nt; wa.sample0 = s0; wa.batch = batch;
            wa.n_paths = nt * batch; wa.spp = spp; wa.max_depth = max_depth; wa.integrator = integrator;
            wa.k0 = (uint32_t)seed; wa.k1 = (uint32_t)(seed >> 32); wa.sample_offset = sample_offset;
            wa.resolve = resolve; wa.last_wave = s0 + batch >= spp;
            cudaError_t e = cudaMemsetAsync(wb.counters, 0, sizeof(unsigned int) * 2 * (max_depth + 2), st);
            if (e != cudaSuccess) return e;
            if (packet0) {
                e = cudaMemsetAsync(chunk_counter, 0, sizeof(unsigned int), st);
                if (e != cudaSuccess) return e;
                int need = ((nt >> 5) * batch + kChunk - 1) / kChunk;
                k_wf_packet0<TRI, STATS><<<packet_grid < need ? packet_grid : need, kPacketThreads, 0, st>>>(
                    sc, cfg.d_cam_prims, cam, wa, wb, chunk_counter, cfg.d_stats);
            } else {
                k_wf_generate<STATS, AOV><<<grid_for(wa.n_paths, 256, cfg.sm_count, 8), 256, 0, st>>>(cam, wa, wb, cfg.d_stats);
            }
            *n_launches += 1;
            int depth = AOV ? 1 : max_depth;
            for (int b = 0; b < depth; ++b) {
                if (b == 0 && !packet0)
                    k_wf_trace<TRI, STATS, true><<<trace_grid, kThreads, 0, st>>>(sc, wb, b, max_depth, cfg.refill_below, cfg.leaf_vote, cfg.d_stats, cfg.wf_rays_per_lane);
                else if (b > 0 && !STATS && sc.treelet != nullptr && sc.treelet_two_t > 0) {      // option "treelet": top levels staged in shared memory
                    const size_t smem = (size_t)sc.treelet_two_t * 32;
                    auto kern = k_wf_trace<TRI, false, false, true>;
                    if (smem > 30 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    int per = 0;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, kThreads, smem);
                    kern<<<cfg.sm_count * (per < 1 ? 1 : per), kThreads, smem, st>>>(sc, wb, b, max_depth, cfg.refill_below, cfg.leaf_vote, nullptr, cfg.wf_rays_per_lane);
                } else if (b > 0 && !STATS && launch_qtrace<TRI>(sc, wb, b, max_depth, cfg, st)) {   // option "qnodes": compressed pairs / cooperative leaves
                } else if (b > 0)
                    k_wf_trace<TRI, STATS, false><<<trace_grid, kThreads, 0, st>>>(sc, wb, b, max_depth, cfg.refill_below, cfg.leaf_vote, cfg.d_stats, cfg.wf_rays_per_lane);
                k_wf_shade<TRI, STATS, AOV><<<grid_for(wa.n_paths, 256, cfg.sm_count, 8), 256, 0, st>>>(sc, wa, wb, b, d_prim, d_t,
                                                                                                      cfg.d_stats);
                *n_launches += (packet0 && b == 0) ? 1 : 2;
            }
            if (!AOV) {
                // per-pixel sums are taken in SAMPLE ORDER: this wave's accumulate follows the previous wave's, whichever stream that ran on
                if (two && acc_pending[set ^ 1]) cudaStreamWaitEvent(st, pipe->acc[set ^ 1], 0);
                k_wf_accumulate<<<grid_for(nt, 256, cfg.sm_count, 8), 256, 0, st>>>(wa, wb, d_out);
                *n_launches += 1;
                if (two) { cudaEventRecord(pipe->acc[set], st); acc_pending[set] = true; }
            }
            s0 += batch;
        }
    }
    if (two) {                                                 // join
        for (int k = 0; k < 2; ++k) {
            cudaEventRecord(pipe->join[k], pipe->streams[k]);
            cudaStreamWaitEvent(cfg.stream, pipe->join[k], 0);
        }
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_wavefront(const SceneView& sc, bool is_tri, bool aov, const CameraBlock& cam, const TileMap& tm,
                             int spp, int max_depth, int integrator, uint64_t seed, uint32_t sample_offset, int resolve,
                             float* d_out, int32_t* d_prim, float* d_t, const LaunchCfg& cfg, const WaveBuffers& wb,
                             int* n_launches, const WavePipe* pipe) {
    if (work_items(tm) == 0) return cudaSuccess;
    bool st = cfg.d_stats != nullptr;
#define RUN(T, S, A) return run_wavefront<T, S, A>(sc, cam, tm, spp, max_depth, integrator, seed, sample_offset, resolve, \
                                                    d_out, d_prim, d_t, cfg, wb, n_launches, pipe)
    if (aov) {
        if (is_tri) { if (st) RUN(true, true, true); else RUN(true, false, true); }
        else { if (st) RUN(false, true, true); else RUN(false, false, true); }
    } else {
        if (is_tri) { if (st) RUN(true, true, false); else RUN(true, false, false); }
        else { if (st) RUN(false, true, false); else RUN(false, false, false); }
    }
#undef RUN
}

}  // namespace b200rt
