// rt_tiny.cu -- the render hot path for TINY scenes: the whole scene lives in shared memory (kernel variant 5, sm_100a).
//
// north_star: "shared-memory staging of the top treelets".  For a scene of a few dozen primitives (the reference's own
// default scene: 9 spheres, interaction.py:294-355; BASELINE config 2: a 36-triangle Cornell box) the top treelet IS the
// scene: every CTA stages all primitive records, the per-camera triangle table (computed in place, no k_cam_tris launch)
// and the materials in shared memory once -- 8 KB at most -- and never touches global memory again except for the pixels
// it writes.  With everything on chip, what limited the general kernels on such scenes was SIMT divergence (the
// megakernel k_render ran the Cornell box with 9.8 of 32 threads active per instruction: divergent BVH walks, dead paths
// idling until the warp's longest path ends) and, for the wavefront, 112 bytes of path state per segment streamed through
// HBM for a scene that fits in 4 KB.  This kernel removes both:
//
//   * closest hit by BRUTE FORCE over the <= 64 primitives, in two passes.  Pass 1, convergent: every lane runs the INSIDE
//     test of the same primitive at the same time (broadcast LDS.128 x 3, no stack, no branch) and notes the primitives its
//     ray passes through in a 64-bit mask.  Pass 2, per lane: the full test -- distance (one IEEE division), closest-hit
//     update -- only for the 2-3 primitives of the mask.  (One pass with the division under a branch executed that branch
//     for the whole warp at 9 of 10 primitives: some lane almost always is inside.)  Closest-hit selection is order
//     independent (ties go to the lower primitive number, consider()), so the result is the BVH walk's bit for bit (the
//     oracle's MODE_BRUTE == MODE_NEAR_FIRST, asserted on CPU);
//   * a CTA-LOCAL WAVEFRONT: a CTA (8 warps) owns 256 paths at a time -- PB pixel blocks (8x4 pixels) x G samples,
//     PB * G = 8 -- and walks them bounce by bounce in lock step: bounce 0 generates the camera rays and tests them through
//     the per-camera table (all 256 threads), every later bounce first COMPACTS the surviving paths through a queue in
//     shared memory (ballot / popc ranks + one shared atomic per warp) so that the survivors fill whole warps, then
//     tests them with the any-ray triangle test.  Path state never leaves the SM: 40 bytes per path per bounce of shared
//     memory instead of 112 bytes of HBM;
//   * per-pixel sample sums are taken IN SAMPLE ORDER by one thread per pixel (radiance of a group's samples is parked in
//     shared memory), so frames are bit-identical to every other variant's and to the oracle's.
//
// Reference lines restated: the integrator is rt_device.cuh's scatter() (old/raytracer_core copy.cpp:211-243 /
// cpp_raytracer/raytracer_core.cpp:291-351), the pixel loop + resolve old/raytracer_core copy.cpp:257-318.
#include "rt_kernel_common.cuh"

// resident CTAs per SM the kernel is compiled for: 4 x 256 / 8 x 128 threads = 64 registers (measured on the Cornell box:
// 17.9 ms against 19.1 ms with 3 CTAs of 80 registers -- the CTAs wait for each other's warps at the phase barriers, so
// more independent CTAs per SM pay more than fewer spills)
#ifndef TINY_MINB_256
#define TINY_MINB_256 4
#endif
#ifndef TINY_MINB_128
#define TINY_MINB_128 8
#endif

namespace b200rt {

namespace {

struct TinyArgs {
    TileMap tm;
    int n_work;                   // 8x4 pixel blocks of the tile map
    int pb, g, g_log2;            // pixel blocks x samples per item, pb * g == warps per CTA, g a power of two
    int spp, max_depth, integrator;
    uint32_t k0, k1, sample_offset;
    int resolve;
    int n, m;                     // primitives (<= kTinyMaxPrims), materials (<= kTinyMaxMats)
};

__device__ __forceinline__ unsigned smem_append(unsigned* counter, bool pred, int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return 0u;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (unsigned)__popc(m & ((1u << lane) - 1u));
}

// Closest hit of ray r over all n primitives in shared memory (see the header: inside pass, then the full test on the
// primitives of the mask).  CAM: camera ray (triangles: per-camera table route).
// Pass 1 for triangles, TWO TRIANGLES PER INSTRUCTION: the records of triangles 2j and 2j + 1 are staged interleaved
// (s_pair, 6 float4 per pair: v0, e1, -e1, e2 for the any-ray test; s_cpair, 5 float4: the three table vectors for the camera-ray
// test), so every product / fma of the inside test is one packed FFMA2 / FMUL2 over the pair (the ray's components broadcast) --
// the arithmetic of tri_mt_inside / cam_tri_inside operation by operation, half the issue slots.
constexpr int kPairF4 = 6, kCPairF4 = 5;

__device__ __forceinline__ float2 bc(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 dot3_2(const float2& ax, const float2& ay, const float2& az, const float2& bx, const float2& by, const float2& bz) {
    return __ffma2_rn(az, bz, __ffma2_rn(ay, by, __fmul2_rn(ax, bx)));
}

// tri_inside() of both triangles of the pair, the sign flips and the u + v sum packed: bit 0 = first, bit 1 = second triangle
__device__ __forceinline__ unsigned tri_inside2(float2 det, float2 un, float2 vn) {
    const float2 sg = make_float2(det.x < 0.0f ? -1.0f : 1.0f, det.y < 0.0f ? -1.0f : 1.0f);
    det = __fmul2_rn(det, sg); un = __fmul2_rn(un, sg); vn = __fmul2_rn(vn, sg);
    const float2 uv = __fadd2_rn(un, vn);
    return ((det.x > 0.0f && un.x >= 0.0f && vn.x >= 0.0f && uv.x <= det.x) ? 1u : 0u) |
           ((det.y > 0.0f && un.y >= 0.0f && vn.y >= 0.0f && uv.y <= det.y) ? 2u : 0u);
}

// any-ray inside test of the pair: bit 0 = triangle 2j, bit 1 = triangle 2j + 1
__device__ __forceinline__ unsigned pair_mt_inside(const float4* p, const Ray& r) {
    const float4 P0 = p[0], P1 = p[1], P2 = p[2], P3 = p[3], P4 = p[4], P5 = p[5];
    const float2 v0x = make_float2(P0.x, P0.y), v0y = make_float2(P0.z, P0.w), v0z = make_float2(P1.x, P1.y);
    const float2 e1x = make_float2(P1.z, P1.w), e1y = make_float2(P2.x, P2.y), e1z = make_float2(P2.z, P2.w);
    const float2 n1x = make_float2(P3.x, P3.y), n1y = make_float2(P3.z, P3.w), n1z = make_float2(P4.x, P4.y);     // -e1
    const float2 e2x = make_float2(P4.z, P4.w), e2y = make_float2(P5.x, P5.y), e2z = make_float2(P5.z, P5.w);
    // p = d x e2 (cross3: fma(ay, bz, -(az * by)), ...), the negated product as (-az) * by
    const float2 px = __ffma2_rn(e2z, bc(r.dy), __fmul2_rn(e2y, bc(-r.dz)));
    const float2 py = __ffma2_rn(e2x, bc(r.dz), __fmul2_rn(e2z, bc(-r.dx)));
    const float2 pz = __ffma2_rn(e2y, bc(r.dx), __fmul2_rn(e2x, bc(-r.dy)));
    const float2 det = dot3_2(e1x, e1y, e1z, px, py, pz);
    const float2 sx = __ffma2_rn(v0x, bc(-1.0f), bc(r.ox)), sy = __ffma2_rn(v0y, bc(-1.0f), bc(r.oy)), sz = __ffma2_rn(v0z, bc(-1.0f), bc(r.oz));
    const float2 un = dot3_2(sx, sy, sz, px, py, pz);
    // q = s x e1, the negated product as az * (-by)
    const float2 qx = __ffma2_rn(sy, e1z, __fmul2_rn(sz, n1y));
    const float2 qy = __ffma2_rn(sz, e1x, __fmul2_rn(sx, n1z));
    const float2 qz = __ffma2_rn(sx, e1y, __fmul2_rn(sy, n1x));
    const float2 vn = dot3_2(bc(r.dx), bc(r.dy), bc(r.dz), qx, qy, qz);
    return tri_inside2(det, un, vn);
}

// camera-ray inside test of the pair (table vectors r0, r1, r2 of both triangles interleaved)
__device__ __forceinline__ unsigned pair_cam_inside(const float4* p, const Ray& r) {
    const float4 P0 = p[0], P1 = p[1], P2 = p[2], P3 = p[3], P4 = p[4];
    const float2 dx = bc(r.dx), dy = bc(r.dy), dz = bc(r.dz);
    const float2 det = dot3_2(dx, dy, dz, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y));
    const float2 un = dot3_2(dx, dy, dz, make_float2(P1.z, P1.w), make_float2(P2.x, P2.y), make_float2(P2.z, P2.w));
    const float2 vn = dot3_2(dx, dy, dz, make_float2(P3.x, P3.y), make_float2(P3.z, P3.w), make_float2(P4.x, P4.y));
    return tri_inside2(det, un, vn);
}

// interleave component c of two float4-record triangles into a pair array (staging)
__device__ __forceinline__ void stage_pair_mt(const float4* s_prims, int n, int j, float4* out) {
    const int a = 2 * j, b = 2 * j + 1;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 av0 = s_prims[3 * a], ae1 = s_prims[3 * a + 1], ae2 = s_prims[3 * a + 2];
    const float4 bv0 = b < n ? s_prims[3 * b] : z, be1 = b < n ? s_prims[3 * b + 1] : z, be2 = b < n ? s_prims[3 * b + 2] : z;   // odd n: a null triangle
    out[0] = make_float4(av0.x, bv0.x, av0.y, bv0.y);
    out[1] = make_float4(av0.z, bv0.z, ae1.x, be1.x);
    out[2] = make_float4(ae1.y, be1.y, ae1.z, be1.z);
    out[3] = make_float4(-ae1.x, -be1.x, -ae1.y, -be1.y);
    out[4] = make_float4(-ae1.z, -be1.z, ae2.x, be2.x);
    out[5] = make_float4(ae2.y, be2.y, ae2.z, be2.z);
}
__device__ __forceinline__ void stage_pair_cam(const float4* s_cam, int n, int j, float4* out) {
    const int a = 2 * j, b = 2 * j + 1;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a0 = s_cam[3 * a], a1 = s_cam[3 * a + 1], a2 = s_cam[3 * a + 2];
    const float4 b0 = b < n ? s_cam[3 * b] : z, b1 = b < n ? s_cam[3 * b + 1] : z, b2 = b < n ? s_cam[3 * b + 2] : z;
    out[0] = make_float4(a0.x, b0.x, a0.y, b0.y);
    out[1] = make_float4(a0.z, b0.z, a1.x, b1.x);
    out[2] = make_float4(a1.y, b1.y, a1.z, b1.z);
    out[3] = make_float4(a2.x, b2.x, a2.y, b2.y);
    out[4] = make_float4(a2.z, b2.z, 0.f, 0.f);
}

template <bool TRI, bool CAM>
__device__ __forceinline__ void tiny_closest(const float4* s_prims, const float4* s_cam, const int* s_slot_prim, const float4* s_pair,
                                             const float4* s_cpair, int n, const Ray& r, Hit& h) {
    h.t = kTMax; h.prim = -1; h.slot = -1;
    unsigned m[2] = {0u, 0u};
    if (TRI) {
        const int n_pairs = (n + 1) >> 1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int j0 = half * 16, j1 = n_pairs < j0 + 16 ? n_pairs : j0 + 16;
            unsigned mm = 0u;
            int sh = 0;
#pragma unroll 2
            for (int j = j0; j < j1; ++j, sh += 2) {
                const unsigned in = CAM ? pair_cam_inside(s_cpair + kCPairF4 * j, r) : pair_mt_inside(s_pair + kPairF4 * j, r);
                mm |= in << sh;
            }
            m[half] = mm;
        }
    } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int k0 = half * 32, k1 = n < k0 + 32 ? n : k0 + 32;
            unsigned bit = 1u, mm = 0u;
#pragma unroll 4
            for (int k = k0; k < k1; ++k, bit <<= 1)
                if (sphere_maybe(s_prims[k], r)) mm |= bit;
            m[half] = mm;
        }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        unsigned mm = m[half];
        while (mm != 0u) {
            const int k = half * 32 + __ffs(mm) - 1;
            mm &= mm - 1u;
            if (TRI) {
                if (CAM) test_cam_tri_records(s_cam[3 * k], s_cam[3 * k + 1], s_cam[3 * k + 2], k, r, h);
                else test_tri_mt_records(s_prims[3 * k], s_prims[3 * k + 1], s_prims[3 * k + 2], k, r, h);
            } else test_sphere_record(s_prims[k], s_slot_prim[k], k, r, h);
        }
    }
}

template <bool TRI, bool STATS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_tiny(const __grid_constant__ SceneView sc, const __grid_constant__ CameraBlock cam, const __grid_constant__ TinyArgs a,
       float* __restrict__ d_out, unsigned int* counter, unsigned long long* d_stats) {
    __shared__ float4 s_prims[kTinyMaxPrims * (TRI ? kTriStride : 1)];
    __shared__ float4 s_cam[TRI ? kTinyMaxPrims * 3 : 1];
    __shared__ int s_slot_prim[kTinyMaxPrims];
    __shared__ float4 s_mats[kTinyMaxMats * 2];
    __shared__ float4 s_pair[TRI ? (kTinyMaxPrims / 2) * kPairF4 : 1];     // triangle pairs, interleaved (pair_mt_inside)
    __shared__ float4 s_cpair[TRI ? (kTinyMaxPrims / 2) * kCPairF4 : 1];   // camera-table pairs (pair_cam_inside)
    __shared__ float4 q_o[2][THREADS];            // origin | path id
    __shared__ float4 q_d[2][THREADS];            // direction | throughput.r
    __shared__ float2 q_t[2][THREADS];            // throughput.g, throughput.b
    __shared__ float s_rad[3][THREADS];           // radiance of the path with this id
    __shared__ unsigned s_cnt[kTinyMaxDepth + 1];
    __shared__ uint32_t s_pixel[THREADS];         // pixel number (RNG counter word) of pixel block b, lane l at [b * 32 + l]
    __shared__ int s_item;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- stage the scene
    const int n = a.n;
    for (int k = tid; k < n * (TRI ? kTriStride : 1); k += THREADS) s_prims[k] = __ldg(sc.prims + k);
    for (int k = tid; k < n; k += THREADS) s_slot_prim[k] = TRI ? k : __ldg(sc.slot_prim + k);
    for (int k = tid; k < a.m * 2; k += THREADS) s_mats[k] = __ldg(sc.mats + k);
    __syncthreads();
    if (TRI) {
        for (int k = tid; k < n; k += THREADS)
            cam_tri_record(s_prims[3 * k], s_prims[3 * k + 1], s_prims[3 * k + 2], cam.px, cam.py, cam.pz, s_cam[3 * k], s_cam[3 * k + 1],
                           s_cam[3 * k + 2]);
        __syncthreads();
        for (int j = tid; j < (n + 1) / 2; j += THREADS) {
            stage_pair_mt(s_prims, n, j, s_pair + kPairF4 * j);
            stage_pair_cam(s_cam, n, j, s_cpair + kCPairF4 * j);
        }
    }
    SceneView ss = sc;                            // the shading helpers read prims / mats through this view: shared memory
    ss.prims = s_prims; ss.mats = s_mats; ss.slot_prim = s_slot_prim;

    const double inv_w = __ddiv_rn(1.0, (double)a.tm.width), inv_h = __ddiv_rn(1.0, (double)a.tm.height);
    const float inv_spp = __fdiv_rn(1.0f, (float)a.spp);
    const int n_items = (a.n_work + a.pb - 1) / a.pb;
    const int my_pb = warp >> a.g_log2, my_s = warp - (my_pb << a.g_log2);   // this thread's camera path: pixel block, sample within the group
    unsigned long long st_rays = 0, st_seg = 0;

    for (;;) {
        __syncthreads();                          // previous item fully consumed (s_item, s_rad)
        if (tid == 0) s_item = (int)atomicAdd(counter, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        // camera paths: thread = (pixel block my_pb, sample my_s, lane); the pixel sums live in the warps < pb (block = warp)
        const int w = item * a.pb + my_pb;
        PixelWork p = decode_work(a.tm, w < a.n_work ? w : a.n_work - 1, lane);
        if (w >= a.n_work) p.active = false;
        const uint32_t pixel = (uint32_t)(p.j * a.tm.width + p.i);
        const int ws = item * a.pb + warp;
        PixelWork ps = decode_work(a.tm, ws < a.n_work ? ws : a.n_work - 1, lane);
        if (ws >= a.n_work || warp >= a.pb) ps.active = false;
        if (warp < a.pb) s_pixel[tid] = (uint32_t)(ps.j * a.tm.width + ps.i);    // visible after the barrier below
        float sum_r = 0.0f, sum_g = 0.0f, sum_b = 0.0f;
        for (int g0 = 0; g0 < a.spp; g0 += a.g) {
            if (tid <= a.max_depth) s_cnt[tid] = 0u;
            __syncthreads();
            // ---- bounce 0: camera ray, table route, shade
            {
                const int s = g0 + my_s;
                const bool act = p.active && s < a.spp;
                float cr = 0.0f, cg = 0.0f, cb = 0.0f, tr = 1.0f, tg = 1.0f, tb = 1.0f;
                bool alive = false;
                Ray r;
                if (act) {
                    const uint32_t sample = a.sample_offset + (uint32_t)s;
                    const uint4 ctl = philox4x32_10(pixel, sample, 0u, 0u, a.k0, a.k1);
                    r = camera_ray(cam, p.i, p.j, u01(ctl.x), u01(ctl.y), inv_w, inv_h);
                    Hit h;
                    tiny_closest<TRI, true>(s_prims, s_cam, s_slot_prim, s_pair, s_cpair, n, r, h);
                    if (STATS) { st_rays += 1; st_seg += 1; }
                    if (h.prim < 0) {
                        cr = __fmaf_rn(tr, sc.bg_r, cr); cg = __fmaf_rn(tg, sc.bg_g, cg); cb = __fmaf_rn(tb, sc.bg_b, cb);
                    } else {
                        const float4* mp = s_mats + 2 * material_row<TRI, true>(ss, h);
                        const float4 m0 = mp[0], m1 = mp[1];
                        cr = __fmaf_rn(tr, m1.y, cr); cg = __fmaf_rn(tg, m1.z, cg); cb = __fmaf_rn(tb, m1.w, cb);
                        if (1 < a.max_depth)
                            alive = scatter<TRI, true>(ss, h, r, a.integrator, 0, a.max_depth, ctl, m0, m1, pixel, sample, a.k0, a.k1, tr, tg, tb);
                    }
                }
                s_rad[0][tid] = cr; s_rad[1][tid] = cg; s_rad[2][tid] = cb;
                const unsigned q = smem_append(&s_cnt[1], alive, lane);
                if (alive) {
                    q_o[1][q] = make_float4(r.ox, r.oy, r.oz, __int_as_float(tid));
                    q_d[1][q] = make_float4(r.dx, r.dy, r.dz, tr);
                    q_t[1][q] = make_float2(tg, tb);
                }
            }
            // ---- bounces >= 1: the survivors, compacted, any-ray route, shade
            for (int b = 1; b < a.max_depth; ++b) {
                __syncthreads();
                const int n_alive = (int)s_cnt[b];
                if (n_alive == 0) break;                              // CTA-uniform
                const int cur = b & 1, nxt = cur ^ 1;
                if (warp * 32 < n_alive) {                            // warp-uniform: warps without a survivor go straight to the barrier
                    bool alive = false;
                    Ray r;
                    float tr = 0.0f, tg = 0.0f, tb = 0.0f;
                    int pid = 0;
                    if (tid < n_alive) {
                        const float4 o = q_o[cur][tid], d = q_d[cur][tid];
                        const float2 t2 = q_t[cur][tid];
                        pid = __float_as_int(o.w);
                        tr = d.w; tg = t2.x; tb = t2.y;
                        r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                        Hit h;
                        tiny_closest<TRI, false>(s_prims, s_cam, s_slot_prim, s_pair, s_cpair, n, r, h);
                        if (STATS) st_seg += 1;
                        float cr = s_rad[0][pid], cg = s_rad[1][pid], cb = s_rad[2][pid];
                        if (h.prim < 0) {
                            cr = __fmaf_rn(tr, sc.bg_r, cr); cg = __fmaf_rn(tg, sc.bg_g, cg); cb = __fmaf_rn(tb, sc.bg_b, cb);
                        } else {
                            const float4* mp = s_mats + 2 * material_row<TRI, true>(ss, h);
                            const float4 m0 = mp[0], m1 = mp[1];
                            cr = __fmaf_rn(tr, m1.y, cr); cg = __fmaf_rn(tg, m1.z, cg); cb = __fmaf_rn(tb, m1.w, cb);
                            if (b + 1 < a.max_depth) {
                                // path id -> its pixel and sample
                                const int pw = pid >> 5, ppb = pw >> a.g_log2, psm = pw - (ppb << a.g_log2);
                                const uint32_t ppix = s_pixel[ppb * 32 + (pid & 31)];
                                const uint32_t sample = a.sample_offset + (uint32_t)(g0 + psm);
                                const uint4 ctl = philox4x32_10(ppix, sample, (uint32_t)b, 0u, a.k0, a.k1);
                                alive = scatter<TRI, true>(ss, h, r, a.integrator, b, a.max_depth, ctl, m0, m1, ppix, sample, a.k0, a.k1, tr, tg, tb);
                            }
                        }
                        s_rad[0][pid] = cr; s_rad[1][pid] = cg; s_rad[2][pid] = cb;
                    }
                    const unsigned q = smem_append(&s_cnt[b + 1], alive, lane);
                    if (alive) {
                        q_o[nxt][q] = make_float4(r.ox, r.oy, r.oz, __int_as_float(pid));
                        q_d[nxt][q] = make_float4(r.dx, r.dy, r.dz, tr);
                        q_t[nxt][q] = make_float2(tg, tb);
                    }
                }
            }
            __syncthreads();
            // ---- the group's samples, in sample order, into the pixel sums
            if (warp < a.pb) {
                const int ns = a.spp - g0 < a.g ? a.spp - g0 : a.g;
                for (int s = 0; s < ns; ++s) {
                    const int id = (warp * a.g + s) * 32 + lane;
                    sum_r = __fadd_rn(sum_r, s_rad[0][id]); sum_g = __fadd_rn(sum_g, s_rad[1][id]); sum_b = __fadd_rn(sum_b, s_rad[2][id]);
                }
            }
        }
        if (warp < a.pb && ws < a.n_work) {                           // warp-uniform
            if (a.resolve) { sum_r = resolve1(sum_r, inv_spp); sum_g = resolve1(sum_g, inv_spp); sum_b = resolve1(sum_b, inv_spp); }
            warp_store_rgb(d_out + 3 * (size_t)ps.out_index, ps.active, sum_r, sum_g, sum_b, lane);
        }
    }
    if (STATS) {
        Counters c = {0, (unsigned long long)n * st_seg, st_seg};      // brute force: no node records, n inside tests per segment
        flush_stats(d_stats, st_rays, c);
    }
}

// Lock-step form (option "tiny_mode" 1): one pixel per thread, samples and bounces in lock step inside the warp, no
// compaction, no barriers, path state in registers -- k_render's structure with the scene in shared memory and the
// two-pass brute-force closest hit.  Dead paths idle until the warp's last path of the sample ends (about 77 % of the
// lanes busy on the Cornell box, against ~90 % after compaction), but warps never wait for each other.
template <bool TRI, bool STATS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_tiny_lockstep(const __grid_constant__ SceneView sc, const __grid_constant__ CameraBlock cam, const __grid_constant__ TinyArgs a,
                float* __restrict__ d_out, unsigned int* counter, unsigned long long* d_stats) {
    __shared__ float4 s_prims[kTinyMaxPrims * (TRI ? kTriStride : 1)];
    __shared__ float4 s_cam[TRI ? kTinyMaxPrims * 3 : 1];
    __shared__ int s_slot_prim[kTinyMaxPrims];
    __shared__ float4 s_mats[kTinyMaxMats * 2];
    __shared__ float4 s_pair[TRI ? (kTinyMaxPrims / 2) * kPairF4 : 1];
    __shared__ float4 s_cpair[TRI ? (kTinyMaxPrims / 2) * kCPairF4 : 1];
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = a.n;
    for (int k = tid; k < n * (TRI ? kTriStride : 1); k += THREADS) s_prims[k] = __ldg(sc.prims + k);
    for (int k = tid; k < n; k += THREADS) s_slot_prim[k] = TRI ? k : __ldg(sc.slot_prim + k);
    for (int k = tid; k < a.m * 2; k += THREADS) s_mats[k] = __ldg(sc.mats + k);
    __syncthreads();
    if (TRI) {
        for (int k = tid; k < n; k += THREADS)
            cam_tri_record(s_prims[3 * k], s_prims[3 * k + 1], s_prims[3 * k + 2], cam.px, cam.py, cam.pz, s_cam[3 * k], s_cam[3 * k + 1],
                           s_cam[3 * k + 2]);
        __syncthreads();
        for (int j = tid; j < (n + 1) / 2; j += THREADS) {
            stage_pair_mt(s_prims, n, j, s_pair + kPairF4 * j);
            stage_pair_cam(s_cam, n, j, s_cpair + kCPairF4 * j);
        }
    }
    __syncthreads();
    SceneView ss = sc;
    ss.prims = s_prims; ss.mats = s_mats; ss.slot_prim = s_slot_prim;
    const double inv_w = __ddiv_rn(1.0, (double)a.tm.width), inv_h = __ddiv_rn(1.0, (double)a.tm.height);
    const float inv_spp = __fdiv_rn(1.0f, (float)a.spp);
    unsigned long long st_rays = 0, st_seg = 0;
    for (;;) {
        const int w = next_work(counter, lane);
        if (w >= a.n_work) break;
        const PixelWork p = decode_work(a.tm, w, lane);
        const uint32_t pixel = (uint32_t)(p.j * a.tm.width + p.i);
        float sum_r = 0.0f, sum_g = 0.0f, sum_b = 0.0f;
        for (int s = 0; s < a.spp; ++s) {
            const uint32_t sample = a.sample_offset + (uint32_t)s;
            float cr = 0.0f, cg = 0.0f, cb = 0.0f, tr = 1.0f, tg = 1.0f, tb = 1.0f;
            bool alive = p.active;
            Ray r;
            uint4 ctl = make_uint4(0u, 0u, 0u, 0u);
            if (alive) {
                ctl = philox4x32_10(pixel, sample, 0u, 0u, a.k0, a.k1);
                r = camera_ray(cam, p.i, p.j, u01(ctl.x), u01(ctl.y), inv_w, inv_h);
                if (STATS) st_rays += 1;
            }
            for (int b = 0; b < a.max_depth; ++b) {
                if (!__any_sync(0xffffffffu, alive)) break;
                if (alive) {
                    Hit h;
                    if (b == 0) tiny_closest<TRI, true>(s_prims, s_cam, s_slot_prim, s_pair, s_cpair, n, r, h);
                    else tiny_closest<TRI, false>(s_prims, s_cam, s_slot_prim, s_pair, s_cpair, n, r, h);
                    if (STATS) st_seg += 1;
                    alive = false;
                    if (h.prim < 0) {
                        cr = __fmaf_rn(tr, sc.bg_r, cr); cg = __fmaf_rn(tg, sc.bg_g, cg); cb = __fmaf_rn(tb, sc.bg_b, cb);
                    } else {
                        const float4* mp = s_mats + 2 * material_row<TRI, true>(ss, h);
                        const float4 m0 = mp[0], m1 = mp[1];
                        cr = __fmaf_rn(tr, m1.y, cr); cg = __fmaf_rn(tg, m1.z, cg); cb = __fmaf_rn(tb, m1.w, cb);
                        if (b + 1 < a.max_depth) {
                            if (b > 0) ctl = philox4x32_10(pixel, sample, (uint32_t)b, 0u, a.k0, a.k1);
                            alive = scatter<TRI, true>(ss, h, r, a.integrator, b, a.max_depth, ctl, m0, m1, pixel, sample, a.k0, a.k1, tr, tg, tb);
                        }
                    }
                }
            }
            sum_r = __fadd_rn(sum_r, cr); sum_g = __fadd_rn(sum_g, cg); sum_b = __fadd_rn(sum_b, cb);
        }
        if (a.resolve) { sum_r = resolve1(sum_r, inv_spp); sum_g = resolve1(sum_g, inv_spp); sum_b = resolve1(sum_b, inv_spp); }
        warp_store_rgb(d_out + 3 * (size_t)p.out_index, p.active, sum_r, sum_g, sum_b, lane);
    }
    if (STATS) {
        Counters c = {0, (unsigned long long)n * st_seg, st_seg};
        flush_stats(d_stats, st_rays, c);
    }
}

}  // namespace

bool tiny_eligible(const SceneView& sc, int n_mats, int max_depth) {
    return sc.n_prims > 0 && sc.n_prims <= kTinyMaxPrims && n_mats > 0 && n_mats <= kTinyMaxMats && max_depth >= 1 && max_depth <= kTinyMaxDepth;
}

cudaError_t launch_tiny(const SceneView& sc, bool is_tri, int n_mats, const CameraBlock& cam, const TileMap& tm, int spp, int max_depth,
                        int integrator, uint64_t seed, uint32_t sample_offset, int resolve, float* d_out, const LaunchCfg& cfg) {
    TinyArgs a;
    a.tm = tm; a.n_work = work_items(tm);
    if (a.n_work == 0) return cudaSuccess;
    const int threads = cfg.tiny_threads == 128 ? 128 : 256, warps = threads / 32;
    a.g = 1; a.g_log2 = 0;
    while (a.g * 2 <= warps && a.g * 2 <= spp) { a.g *= 2; a.g_log2 += 1; }   // samples per group: the largest power of two <= min(spp, warps)
    a.pb = warps / a.g;
    a.spp = spp; a.max_depth = max_depth; a.integrator = integrator;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32); a.sample_offset = sample_offset;
    a.resolve = resolve; a.n = sc.n_prims; a.m = n_mats;
    cudaError_t e = cudaMemsetAsync(cfg.d_work_counter, 0, sizeof(unsigned int), cfg.stream);
    if (e != cudaSuccess) return e;
    const int n_items = (a.n_work + a.pb - 1) / a.pb;
    const bool st = cfg.d_stats != nullptr;
    if (cfg.tiny_mode == 1) {
#define LAUNCH_LS(T, S)                                                                                      \
    {                                                                                                        \
        int per_sm = 0;                                                                                      \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tiny_lockstep<T, S, 128, 8>, 128, 0);       \
        int grid = cfg.sm_count * (per_sm < 1 ? 1 : per_sm);                                                 \
        const int need = (a.n_work + 3) / 4;                                                                 \
        if (grid > need) grid = need;                                                                        \
        k_tiny_lockstep<T, S, 128, 8><<<grid, 128, 0, cfg.stream>>>(sc, cam, a, d_out, cfg.d_work_counter, cfg.d_stats); \
    }
        if (is_tri) { if (st) LAUNCH_LS(true, true) else LAUNCH_LS(true, false) }
        else { if (st) LAUNCH_LS(false, true) else LAUNCH_LS(false, false) }
#undef LAUNCH_LS
        return cudaGetLastError();
    }
#define LAUNCH(T, S, TH, MB)                                                                                 \
    {                                                                                                        \
        int per_sm = 0;                                                                                      \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tiny<T, S, TH, MB>, TH, 0);                 \
        int grid = cfg.sm_count * (per_sm < 1 ? 1 : per_sm);                                                 \
        if (grid > n_items) grid = n_items;                                                                  \
        k_tiny<T, S, TH, MB><<<grid, TH, 0, cfg.stream>>>(sc, cam, a, d_out, cfg.d_work_counter, cfg.d_stats); \
    }
#define LAUNCH_TS(T, S) { if (threads == 128) LAUNCH(T, S, 128, TINY_MINB_128) else LAUNCH(T, S, 256, TINY_MINB_256) }
    if (is_tri) { if (st) LAUNCH_TS(true, true) else LAUNCH_TS(true, false) }
    else { if (st) LAUNCH_TS(false, true) else LAUNCH_TS(false, false) }
#undef LAUNCH_TS
#undef LAUNCH
    return cudaGetLastError();
}

}  // namespace b200rt
