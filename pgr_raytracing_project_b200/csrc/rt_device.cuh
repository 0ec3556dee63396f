// rt_device.cuh -- device-side building blocks of the render hot path (sm_100a).
//
// Arithmetic contract: IEEE float32/float64, this translation unit is compiled
// with -fmad=false so nothing is contracted implicitly; every fused multiply-add below is an
// explicit __fmaf_rn.  sqrt and division are the IEEE-rounded ones (nvcc defaults
// -prec-sqrt=true -prec-div=true; never build this with --use_fast_math).  The CPU checker in
// oracle/ follows the same contract, written independently, which is what makes hit ids, hit
// distances and whole images comparable bit for bit.
//
// Reference lines restated here (all under /root/reference):
//   camera ray ............ old/raytracer_core copy.h:160-184, old/raytracer_core copy.cpp:288-289
//   ray setup ............. cpp_raytracer/raytracer_core.h:107-121
//   slab test ............. cpp_raytracer/raytracer_core.h:132-153 + swap of old/bvh copy.cpp:15-17
//   sphere test ........... old/raytracer_core copy.cpp:21-52
//   traversal ............. cpp_raytracer/raytracer_core.cpp:198-243 (closest hit, shrinking tmax)
//   integrators ........... old/raytracer_core copy.cpp:211-243 (0), raytracer_core.cpp:291-351 (1)
//   resolve ............... cpp_raytracer/raytracer_core.cpp:398-409
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200rt {

constexpr float kTMin = 0.001f;   // old/raytracer_core copy.cpp:217
constexpr float kTMax = 1e10f;
constexpr int kStackDepth = 64;   // cpp_raytracer/raytracer_core.cpp:200
constexpr int kTinyMaxPrims = 64, kTinyMaxMats = 64, kTinyMaxDepth = 8;   // scenes the tiny-scene kernel (rt_tiny.cu) takes
constexpr int kTriStride = 3;     // float4 per triangle record: v0|prim, e1|material, e2|0 (48 B).  Measured and dropped: 64-byte
                                  // records read as two 256-bit loads (7.82 vs 7.78 ms on the C3 4-spp depth-4 frame, +33 % bytes)

struct CameraBlock {              // basis in double: primary directions are rounded to f32 once
    float px, py, pz;
    float pad_;
    double fwd[3], right[3], up[3];
    double sx, sy;                // aspect * tan(fov/2), tan(fov/2)
};

struct SceneView {
    const float4* __restrict__ nodes;      // 2 x float4 per 32-byte node: bmin | code, bmax | 0 (see intersect())
    const float4* __restrict__ prims;      // leaf order; triangle: v0|prim, e1|material, e2|0 (kTriStride x float4) ; sphere: c|r
    const float4* __restrict__ cam_prims;  // triangles: camera-relative records for the current camera position (k_cam_tris)
    const int* __restrict__ slot_prim;     // slot -> primitive number (upload order)
    const float4* __restrict__ mats;       // 2 x float4 per material: albedo|metallic, roughness|emission
    const float4* __restrict__ treelet;    // top levels of the tree as a heap of sibling pairs (k_build_treelet), or nullptr
    int treelet_two_t;                     // 2 x pairs in the treelet = node records a CTA stages in shared memory (0: none)
    int n_prims;
    int n_nodes;
    int n_mats;
    int sane_extent;                       // every |coordinate| of the root box < 2^40 (octant slab test usable)
    float bg_r, bg_g, bg_b;
    // option "qnodes" (incoherent bounces, k_wf_trace): compressed copies of the tree / the triangle records, or nullptr
    const uint4* __restrict__ qnodes;      // 32 bytes per sibling pair: planes on a 15-bit grid over the root box (k_quantize_pairs)
    const float* __restrict__ qgrid;       // grid origin xyz, grid extent xyz (written by k_quantize_pairs)
    const float4* __restrict__ tri_a;      // triangles, leaf order: v0|prim, e1|material (32 bytes, ONE 256-bit load) ...
    const float4* __restrict__ tri_b;      // ... and e2|0 (16 bytes)
};

struct Counters { unsigned long long nodes, prims, segments; };

struct Ray { float ox, oy, oz, dx, dy, dz, ix, iy, iz, ax, ay, az; };   // a = o * inv (slab offsets)

// ------------------------------------------------------------------------------- math
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
}
__device__ __forceinline__ void cross3(float ax, float ay, float az, float bx, float by, float bz,
                                       float& rx, float& ry, float& rz) {
    rx = __fmaf_rn(ay, bz, -__fmul_rn(az, by));
    ry = __fmaf_rn(az, bx, -__fmul_rn(ax, bz));
    rz = __fmaf_rn(ax, by, -__fmul_rn(ay, bx));
}
// exactly antisymmetric cross product (both products rounded, then subtracted): used by the triangle
// test so that the shared edge of the two halves of a quad is watertight (oracle: cross_as)
__device__ __forceinline__ void cross_as(float ax, float ay, float az, float bx, float by, float bz,
                                         float& rx, float& ry, float& rz) {
    rx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
    ry = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
    rz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
}
__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {   // raytracer_core.h:91-94
    float len = __fsqrt_rn(dot3(x, y, z, x, y, z));
    if (len > 0.0f) {
        float inv = __fdiv_rn(1.0f, len);
        x = __fmul_rn(x, inv); y = __fmul_rn(y, inv); z = __fmul_rn(z, inv);
    } else { x = 0.0f; y = 0.0f; z = 1.0f; }
}

// ------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t x) { return __fmul_rn((float)(x >> 8), 0x1p-24f); }

// ------------------------------------------------------------------------------- rays
// 1/d of the slab test.  A direction component that is zero (an axis-parallel ray) or too small for 1/d to stay finite
// gets +-2^80 instead: a power of two, so o * inv and the slab distances fmaf(plane, inv, -o * inv) = (plane - o) * 2^80 are
// exact, finite and sign-correct -- the ray is inside the slab iff lo <= o <= hi, as in the reference's (plane - o) * (1/d)
// form (old/bvh copy.cpp:9-25) -- where inf would make o * inv - o * inv a NaN and lose the box.  (oracle: safe_inv)
__device__ __forceinline__ float safe_inv(float d) { return fabsf(d) < 0x1p-80f ? copysignf(0x1p80f, d) : __fdiv_rn(1.0f, d); }

__device__ __forceinline__ Ray make_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
    Ray r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
    r.ix = safe_inv(dx); r.iy = safe_inv(dy); r.iz = safe_inv(dz);
    r.ax = __fmul_rn(ox, r.ix); r.ay = __fmul_rn(oy, r.iy); r.az = __fmul_rn(oz, r.iz);
    return r;
}

// Camera ray through pixel (i + jx, j + jy); direction evaluated in double, rounded once.
__device__ __forceinline__ Ray camera_ray(const CameraBlock& c, int i, int j, float jx, float jy,
                                          double inv_w, double inv_h) {
    double u = __dmul_rn(__dadd_rn((double)i, (double)jx), inv_w);
    double v = __dmul_rn(__dadd_rn((double)j, (double)jy), inv_h);
    double ndc_x = __dmul_rn(__dsub_rn(u, 0.5), 2.0);
    double ndc_y = __dmul_rn(__dsub_rn(0.5, v), 2.0);
    double vx = __dmul_rn(ndc_x, c.sx), vy = __dmul_rn(ndc_y, c.sy);
    double dx = __dadd_rn(__dadd_rn(c.fwd[0], __dmul_rn(c.right[0], vx)), __dmul_rn(c.up[0], vy));
    double dy = __dadd_rn(__dadd_rn(c.fwd[1], __dmul_rn(c.right[1], vx)), __dmul_rn(c.up[1], vy));
    double dz = __dadd_rn(__dadd_rn(c.fwd[2], __dmul_rn(c.right[2], vx)), __dmul_rn(c.up[2], vy));
    double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    double il = __ddiv_rn(1.0, len);
    return make_ray(c.px, c.py, c.pz, (float)__dmul_rn(dx, il), (float)__dmul_rn(dy, il), (float)__dmul_rn(dz, il));
}

// ------------------------------------------------------------------------------- intersection
struct Hit { float t; int prim; int slot; };

// explicit shared-memory accesses through a 32-bit shared-window address (HybridStackA, leaf_phase_coop)
__device__ __forceinline__ void sts_f32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_v2(unsigned a, unsigned x, unsigned y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ uint2 lds_v2(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void reds_min_u64(unsigned a, unsigned long long v) { asm volatile("red.shared.min.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory"); }


// A sibling pair (64 bytes, 64-byte aligned) as TWO 256-bit loads (LDG.E.256.CONSTANT, sm_100: ld.global.nc.v8.f32)
// instead of four 128-bit ones.  The per-lane traversal of incoherent rays is bound by the L1TEX tag stage: every
// load instruction costs one pass per DISTINCT 128-byte line among the lanes, however many bytes each lane takes
// from its line, so halving the instructions halves the passes (ncu: l1tex__throughput 87 % with 4 x LDG.128).
__device__ __forceinline__ void ldg_node(const float4* __restrict__ p, float4& lo, float4& hi) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "l"(p));
}
// ---- device node records (written by rt_api.cu ensure_device, rt_lbvh.cu k_device_nodes, rt_refit.cu k_finish).
// Node k: box lo / hi + a code: code >= 0 = internal, index of the child pair (children code, code + 1); code <= -2 = leaf,
// ~code = (first_slot << 3) | count.  Node 0 is the root, node 1 a pad record, siblings are nodes 2m, 2m + 1.
// A sibling pair is stored INTERLEAVED in 64 aligned bytes (two 256-bit loads):
//   float4 0: L.lo.x R.lo.x L.lo.y R.lo.y     float4 1: L.lo.z R.lo.z L.code R.code
//   float4 2: L.hi.x R.hi.x L.hi.y R.hi.y     float4 3: L.hi.z R.hi.z 0      0
// so that the same plane of the LEFT and the RIGHT child sit in one 64-bit register pair: the slab products of both children
// are packed FFMA2 (fma.rn.f32x2, sm_100: two IEEE fmas per instruction, scalar operands broadcast) -- 6 instead of 12 issue
// slots per traversal step, the same bits.  Everything that is not the traversal reads / writes nodes through node_slot().
__host__ __device__ inline size_t node_slot(int k, int c) {           // c: 0..2 = lo.xyz, 3 = code, 4..6 = hi.xyz -> float index
    return (size_t)(k >> 1) * 16 + (size_t)(c < 4 ? 2 * c : 8 + 2 * (c - 4)) + (size_t)(k & 1);
}
__device__ __forceinline__ int node_code(const float4* __restrict__ nodes, int k) {
    return __float_as_int(reinterpret_cast<const float*>(nodes)[node_slot(k, 3)]);
}
__device__ __forceinline__ void node_read(const float4* __restrict__ nodes, int k, float lo[3], float hi[3], int& code) {
    const float* f = reinterpret_cast<const float*>(nodes);
    for (int c = 0; c < 3; ++c) { lo[c] = f[node_slot(k, c)]; hi[c] = f[node_slot(k, 4 + c)]; }
    code = __float_as_int(f[node_slot(k, 3)]);
}
__device__ __forceinline__ void node_write(float4* __restrict__ nodes, int k, const float lo[3], const float hi[3], int code) {
    float* f = reinterpret_cast<float*>(nodes);
    for (int c = 0; c < 3; ++c) { f[node_slot(k, c)] = lo[c]; f[node_slot(k, 4 + c)] = hi[c]; }
    f[node_slot(k, 3)] = __int_as_float(code);
}
// the root (node 0 = left record of pair 0) as a plain box: lo | code, hi | 0
__device__ __forceinline__ void root_node(const float4* __restrict__ nodes, float4& lo, float4& hi) {
    const float4 a = __ldg(nodes), b = __ldg(nodes + 1), c = __ldg(nodes + 2), d = __ldg(nodes + 3);
    lo = make_float4(a.x, a.z, b.x, b.z);
    hi = make_float4(c.x, c.z, d.x, 0.0f);
}

struct PairRec { float4 a, b, c, d; };                                 // the four float4 of a sibling pair (layout above)
__device__ __forceinline__ void ldg_pair(const float4* __restrict__ p, PairRec& q) {
    ldg_node(p, q.a, q.b);
    ldg_node(p + 2, q.c, q.d);
}
__device__ __forceinline__ int pair_lc(const PairRec& q) { return __float_as_int(q.b.z); }
__device__ __forceinline__ int pair_rc(const PairRec& q) { return __float_as_int(q.b.w); }

__device__ __forceinline__ bool box_hit(const float4& lo, const float4& hi, const Ray& r, float tlo,
                                        float thi, float& tn) {
    float x1 = __fmaf_rn(lo.x, r.ix, -r.ax), x2 = __fmaf_rn(hi.x, r.ix, -r.ax);
    float y1 = __fmaf_rn(lo.y, r.iy, -r.ay), y2 = __fmaf_rn(hi.y, r.iy, -r.ay);
    float z1 = __fmaf_rn(lo.z, r.iz, -r.az), z2 = __fmaf_rn(hi.z, r.iz, -r.az);
    float n = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fmaxf(fminf(z1, z2), tlo));
    float f = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fminf(fmaxf(z1, z2), thi));
    tn = n;
    return n <= f;
}

// Both children of a pair against one ray: box_hit() of the left and of the right child, the six slab products of each plane pair
// as ONE packed fma (x = left child, y = right child; the very values box_hit computes).
__device__ __forceinline__ void pair_hit(const PairRec& q, const Ray& r, float tlo, float thi, bool& hl, bool& hr, float& tl, float& tr) {
    const float2 ix = make_float2(r.ix, r.ix), iy = make_float2(r.iy, r.iy), iz = make_float2(r.iz, r.iz);
    const float2 ax = make_float2(-r.ax, -r.ax), ay = make_float2(-r.ay, -r.ay), az = make_float2(-r.az, -r.az);
    const float2 x1 = __ffma2_rn(make_float2(q.a.x, q.a.y), ix, ax), x2 = __ffma2_rn(make_float2(q.c.x, q.c.y), ix, ax);
    const float2 y1 = __ffma2_rn(make_float2(q.a.z, q.a.w), iy, ay), y2 = __ffma2_rn(make_float2(q.c.z, q.c.w), iy, ay);
    const float2 z1 = __ffma2_rn(make_float2(q.b.x, q.b.y), iz, az), z2 = __ffma2_rn(make_float2(q.d.x, q.d.y), iz, az);
    const float nl = fmaxf(fmaxf(fminf(x1.x, x2.x), fminf(y1.x, y2.x)), fmaxf(fminf(z1.x, z2.x), tlo));
    const float fl = fminf(fminf(fmaxf(x1.x, x2.x), fmaxf(y1.x, y2.x)), fminf(fmaxf(z1.x, z2.x), thi));
    const float nr = fmaxf(fmaxf(fminf(x1.y, x2.y), fminf(y1.y, y2.y)), fmaxf(fminf(z1.y, z2.y), tlo));
    const float fr = fminf(fminf(fmaxf(x1.y, x2.y), fmaxf(y1.y, y2.y)), fminf(fmaxf(z1.y, z2.y), thi));
    tl = nl; tr = nr;
    hl = nl <= fl; hr = nr <= fr;
}

// ------------------------------------------------------------------------------- compressed sibling pairs (option "qnodes")
// The per-lane traversal of incoherent rays pays one L1 data-pipe pass per LANE and load instruction (ncu:
// l1tex__data_pipe_lsu_wavefronts 82 % of peak in k_wf_trace, about 1 wavefront per touched sector), so a sibling pair read as
// 2 x LDG.256 costs two passes per ray and step.  The compressed copy holds a pair in 32 bytes = ONE 256-bit load:
//   word 0..2: left child,  per axis  lo16 | hi16 << 16        word 3..5: right child, the same
//   word 6, 7: the two child codes (as in the full records)
// A 16-bit plane is 0x8000 | q, q in 0..32767 a cell index of a grid over the (slightly enlarged) root box, rounded OUTWARD
// by the quantiser (k_quantize_pairs), so that bytes (0x3F, hi byte, lo byte, 0x00) ARE the float f = 1 + q / 32768 and one
// PRMT (byte permute) decodes a plane; its selector is per lane and picks lo or hi by the sign of the ray's direction, so the
// decode also replaces the slab test's per-axis min / max.  With A = extent * (1/d) and B = (origin - extent - o) * (1/d)
// the plane's distance along the ray is fma(f, A, B).  The boxes are conservative (outward rounding + the slack below for
// the float evaluation), closest hits do not depend on the order or the number of boxes entered (consider()), so the
// hits are the exact tree's bit for bit; what changes is a few more boxes entered (grid cell = extent / 32768).
struct QRay { float ax, ay, az, bnx, bny, bnz, bfx, bfy, bfz; unsigned snx, sny, snz; };

__device__ __forceinline__ void ldg_u8(const uint4* __restrict__ p, uint4& a, uint4& b) {
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
constexpr unsigned kQSelLo = 0x7104u, kQSelHi = 0x7324u, kQSelFlip = 0x0220u, kQExp = 0x3F000000u;

// one axis of the ray against the grid: distance of plane f is fma(f, a, b); bn / bf = b minus / plus a slack that covers
// the rounding of a, b and of the fma itself (each 2^-24 relative to 2|a| + |b|; 2^-21 leaves a factor 4), so that the near
// value never exceeds and the far value never falls below the exact distance of the (already enlarged) plane
__device__ __forceinline__ void qray_axis(float g0, float e, float o, float inv, float& a, float& bn, float& bf, unsigned& sn) {
    a = __fmul_rn(e, inv);
    const double c = __dsub_rn(__dsub_rn((double)g0, (double)e), (double)o);
    const float b = __double2float_rn(__dmul_rn(c, (double)inv));
    const float slack = __fmul_rn(0x1p-21f, __fmaf_rn(2.0f, fabsf(a), fabsf(b)));
    bn = __fsub_rn(b, slack); bf = __fadd_rn(b, slack);
    sn = inv < 0.0f ? kQSelHi : kQSelLo;
}
__device__ __forceinline__ QRay make_qray(const float* __restrict__ grid, const Ray& r) {
    QRay q;
    qray_axis(__ldg(grid + 0), __ldg(grid + 3), r.ox, r.ix, q.ax, q.bnx, q.bfx, q.snx);
    qray_axis(__ldg(grid + 1), __ldg(grid + 4), r.oy, r.iy, q.ay, q.bny, q.bfy, q.sny);
    qray_axis(__ldg(grid + 2), __ldg(grid + 5), r.oz, r.iz, q.az, q.bnz, q.bfz, q.snz);
    return q;
}
// (prmt.b32 itself: __byte_perm() masks its selector with 0x7777 first, one more instruction per axis and step)
__device__ __forceinline__ float qplane(unsigned w, unsigned sel) {
    unsigned f;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(w), "r"(kQExp), "r"(sel));
    return __uint_as_float(f);
}

// both children of a compressed pair (words w0 = Lx Ly Lz Rx, w1 = Ry Rz Lcode Rcode) against one ray
__device__ __forceinline__ void pair_hit_q(const uint4& w0, const uint4& w1, const QRay& q, float tlo, float thi,
                                           bool& hl, bool& hr, float& tl, float& tr) {
    const unsigned fx = q.snx ^ kQSelFlip, fy = q.sny ^ kQSelFlip, fz = q.snz ^ kQSelFlip;
    const float2 nx = __ffma2_rn(make_float2(qplane(w0.x, q.snx), qplane(w0.w, q.snx)), make_float2(q.ax, q.ax), make_float2(q.bnx, q.bnx));
    const float2 ny = __ffma2_rn(make_float2(qplane(w0.y, q.sny), qplane(w1.x, q.sny)), make_float2(q.ay, q.ay), make_float2(q.bny, q.bny));
    const float2 nz = __ffma2_rn(make_float2(qplane(w0.z, q.snz), qplane(w1.y, q.snz)), make_float2(q.az, q.az), make_float2(q.bnz, q.bnz));
    const float2 gx = __ffma2_rn(make_float2(qplane(w0.x, fx), qplane(w0.w, fx)), make_float2(q.ax, q.ax), make_float2(q.bfx, q.bfx));
    const float2 gy = __ffma2_rn(make_float2(qplane(w0.y, fy), qplane(w1.x, fy)), make_float2(q.ay, q.ay), make_float2(q.bfy, q.bfy));
    const float2 gz = __ffma2_rn(make_float2(qplane(w0.z, fz), qplane(w1.y, fz)), make_float2(q.az, q.az), make_float2(q.bfz, q.bfz));
    const float nl = fmaxf(fmaxf(nx.x, ny.x), fmaxf(nz.x, tlo)), fl = fminf(fminf(gx.x, gy.x), fminf(gz.x, thi));
    const float nr = fmaxf(fmaxf(nx.y, ny.y), fmaxf(nz.y, tlo)), fr = fminf(fminf(gx.y, gy.y), fminf(gz.y, thi));
    tl = nl; tr = nr;
    hl = nl <= fl; hr = nr <= fr;
}

// closest-hit update; ties go to the lower primitive number
__device__ __forceinline__ void consider(Hit& h, float t, int prim, int slot) {
    if (!(t >= kTMin && t <= h.t)) return;
    if (t < h.t || h.prim < 0 || prim < h.prim) { h.t = t; h.prim = prim; h.slot = slot; }
}

// Ray/triangle tests (an extension: the reference has no triangles).  Two routes, the same rule on both sides
// of the parity check (oracle/rt_oracle.c test_tri_cam / test_tri_mt):
//
// CAMERA RAYS (bounce 0 of every path, the primary-hit query; one shared origin): Moller-Trumbore written as scalar
// triple products of the direction d with vectors that depend only on the triangle and the ORIGIN (s = o - v0):
//   det = d.(e2 x e1)   u*det = d.(e2 x s)   v*det = d.(s x e1)   t*det = e2.(s x e1)
// The three vectors and the scalar are a per-camera table (cam_tri_record, written by k_cam_tris), so the test is
// three dot products; inside test division-free against det made positive (negation is exact), distance = one IEEE
// division.  The cross products are exactly antisymmetric (cross_as), so the two halves of a quad (shared v0 and
// edge vector) have u*det of one equal to -v*det of the other: no cracks along quad diagonals in what the camera
// sees.
//
// ANY OTHER RAY (bounces >= 1, rt_trace_rays): classic Moller-Trumbore, p = d x e2, q = s x e1, det = e1.p,
// u*det = s.p, v*det = d.q, t*det = e2.q, with the same division-free inside test -- 15 fewer operations per
// test than evaluating the table vectors per ray, and the triangle test is half of the incoherent-bounce kernel.
__device__ __forceinline__ void tri_accept(Hit& h, float det, float un, float vn, float c, int prim, int slot) {
    const float sg = det < 0.0f ? -1.0f : 1.0f;              // sign flip as a multiplication by +-1 (exact; FFMA pipe)
    det = __fmul_rn(det, sg); un = __fmul_rn(un, sg); vn = __fmul_rn(vn, sg);
    if (det > 0.0f && un >= 0.0f && vn >= 0.0f && __fadd_rn(un, vn) <= det)
        consider(h, __fdiv_rn(__fmul_rn(c, sg), det), prim, slot);
}

// The inside test of tri_accept alone (no distance): exactly its predicate, so that a kernel may first collect the
// triangles a ray passes THROUGH (cheap, convergent) and run the full test -- division, closest-hit update -- only on those.
__device__ __forceinline__ bool tri_inside(float det, float un, float vn) {
    const float sg = det < 0.0f ? -1.0f : 1.0f;
    det = __fmul_rn(det, sg); un = __fmul_rn(un, sg); vn = __fmul_rn(vn, sg);
    return det > 0.0f && un >= 0.0f && vn >= 0.0f && __fadd_rn(un, vn) <= det;
}

// (e2 x e1 | e2.(s x e1)), (e2 x s | prim), (s x e1 | material) for origin (ox,oy,oz)
__device__ __forceinline__ void cam_tri_record(float4 v0, float4 e1, float4 e2, float ox, float oy, float oz,
                                               float4& r0, float4& r1, float4& r2) {
    float sx = __fsub_rn(ox, v0.x), sy = __fsub_rn(oy, v0.y), sz = __fsub_rn(oz, v0.z);
    cross_as(e2.x, e2.y, e2.z, e1.x, e1.y, e1.z, r0.x, r0.y, r0.z);
    cross_as(e2.x, e2.y, e2.z, sx, sy, sz, r1.x, r1.y, r1.z);
    cross_as(sx, sy, sz, e1.x, e1.y, e1.z, r2.x, r2.y, r2.z);
    r0.w = dot3(e2.x, e2.y, e2.z, r2.x, r2.y, r2.z);
    r1.w = v0.w; r2.w = e1.w;
}

// camera route: record `slot` of the per-camera table; all three loads issued up front, one combined predicate
__device__ __forceinline__ void test_cam_tri_records(const float4& r0, const float4& r1, const float4& r2, int slot, const Ray& r, Hit& h) {
    const float det = dot3(r.dx, r.dy, r.dz, r0.x, r0.y, r0.z);
    const float un = dot3(r.dx, r.dy, r.dz, r1.x, r1.y, r1.z);
    const float vn = dot3(r.dx, r.dy, r.dz, r2.x, r2.y, r2.z);
    tri_accept(h, det, un, vn, r0.w, __float_as_int(r1.w), slot);
}
__device__ __forceinline__ bool cam_tri_inside(const float4& r0, const float4& r1, const float4& r2, const Ray& r) {
    return tri_inside(dot3(r.dx, r.dy, r.dz, r0.x, r0.y, r0.z), dot3(r.dx, r.dy, r.dz, r1.x, r1.y, r1.z), dot3(r.dx, r.dy, r.dz, r2.x, r2.y, r2.z));
}
// The per-camera TABLE in global memory (k_cam_tris) holds the record re-packed so that det and u*det -- two dot products with
// the same direction -- are one chain of packed fmas (FFMA2: 3 issue slots instead of 6; the same bits as two dot3):
//   A = (r0.x, r1.x, r0.y, r1.y)   B = (r0.z, r1.z, r2.x, r2.y)   C = (r2.z, e2.(s x e1), prim, material)
__device__ __forceinline__ void cam_tri_pack(const float4& r0, const float4& r1, const float4& r2, float4& A, float4& B, float4& C) {
    A = make_float4(r0.x, r1.x, r0.y, r1.y);
    B = make_float4(r0.z, r1.z, r2.x, r2.y);
    C = make_float4(r2.z, r0.w, r1.w, r2.w);
}
__device__ __forceinline__ void test_cam_tri_packet(const float4* __restrict__ cam_prims, int slot, const Ray& r, Hit& h) {
    const float4* p = cam_prims + 3 * (size_t)slot;
    const float4 A = __ldg(p), B = __ldg(p + 1), C = __ldg(p + 2);
    float2 du = __fmul2_rn(make_float2(A.x, A.y), make_float2(r.dx, r.dx));
    du = __ffma2_rn(make_float2(A.z, A.w), make_float2(r.dy, r.dy), du);
    du = __ffma2_rn(make_float2(B.x, B.y), make_float2(r.dz, r.dz), du);
    float vn = __fmaf_rn(r.dz, C.x, __fmaf_rn(r.dy, B.w, __fmul_rn(r.dx, B.z)));
    // tri_accept() with det and u*det still packed: the sign flip of the pair is one instruction
    const float sg = du.x < 0.0f ? -1.0f : 1.0f;
    du = __fmul2_rn(du, make_float2(sg, sg)); vn = __fmul_rn(vn, sg);
    if (du.x > 0.0f && du.y >= 0.0f && vn >= 0.0f && __fadd_rn(du.y, vn) <= du.x)
        consider(h, __fdiv_rn(__fmul_rn(C.y, sg), du.x), __float_as_int(C.z), slot);
}

// any-ray route (records v0 | prim, e1 | material, e2 | 0 already in registers)
__device__ __forceinline__ void test_tri_mt_records(const float4& v0, const float4& e1, const float4& e2, int slot, const Ray& r, Hit& h) {
    float px, py, pz, qx, qy, qz;
    cross3(r.dx, r.dy, r.dz, e2.x, e2.y, e2.z, px, py, pz);
    const float det = dot3(e1.x, e1.y, e1.z, px, py, pz);
    const float sx = __fsub_rn(r.ox, v0.x), sy = __fsub_rn(r.oy, v0.y), sz = __fsub_rn(r.oz, v0.z);
    const float un = dot3(sx, sy, sz, px, py, pz);
    cross3(sx, sy, sz, e1.x, e1.y, e1.z, qx, qy, qz);
    const float vn = dot3(r.dx, r.dy, r.dz, qx, qy, qz);
    const float c = dot3(e2.x, e2.y, e2.z, qx, qy, qz);
    tri_accept(h, det, un, vn, c, __float_as_int(v0.w), slot);
}
// inside test of test_tri_mt_records alone (the same det, u*det, v*det)
__device__ __forceinline__ bool tri_mt_inside(const float4& v0, const float4& e1, const float4& e2, const Ray& r) {
    float px, py, pz, qx, qy, qz;
    cross3(r.dx, r.dy, r.dz, e2.x, e2.y, e2.z, px, py, pz);
    const float det = dot3(e1.x, e1.y, e1.z, px, py, pz);
    const float sx = __fsub_rn(r.ox, v0.x), sy = __fsub_rn(r.oy, v0.y), sz = __fsub_rn(r.oz, v0.z);
    const float un = dot3(sx, sy, sz, px, py, pz);
    cross3(sx, sy, sz, e1.x, e1.y, e1.z, qx, qy, qz);
    const float vn = dot3(r.dx, r.dy, r.dz, qx, qy, qz);
    return tri_inside(det, un, vn);
}
__device__ __forceinline__ void test_tri_mt(const SceneView& sc, int slot, const Ray& r, Hit& h) {
    const float4* p = sc.prims + kTriStride * (size_t)slot;
    test_tri_mt_records(__ldg(p), __ldg(p + 1), __ldg(p + 2), slot, r, h);
}

// the discriminant test of test_sphere_record alone
__device__ __forceinline__ bool sphere_maybe(const float4& s, const Ray& r) {
    double ocx = __dsub_rn((double)r.ox, (double)s.x), ocy = __dsub_rn((double)r.oy, (double)s.y),
           ocz = __dsub_rn((double)r.oz, (double)s.z);
    double dx = r.dx, dy = r.dy, dz = r.dz, rad = s.w;
    double a = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    double half_b = __dadd_rn(__dadd_rn(__dmul_rn(ocx, dx), __dmul_rn(ocy, dy)), __dmul_rn(ocz, dz));
    double c = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(ocx, ocx), __dmul_rn(ocy, ocy)), __dmul_rn(ocz, ocz)),
                         __dmul_rn(rad, rad));
    return !(__dsub_rn(__dmul_rn(half_b, half_b), __dmul_rn(a, c)) < 0.0);
}

// v1 Sphere::hit in double on the float32 ray / sphere (centre | radius), roots rounded to float32; `prim` is the
// primitive number of slot `slot`
__device__ __forceinline__ void test_sphere_record(const float4& s, int prim, int slot, const Ray& r, Hit& h) {
    double ocx = __dsub_rn((double)r.ox, (double)s.x), ocy = __dsub_rn((double)r.oy, (double)s.y),
           ocz = __dsub_rn((double)r.oz, (double)s.z);
    double dx = r.dx, dy = r.dy, dz = r.dz, rad = s.w;
    double a = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    double half_b = __dadd_rn(__dadd_rn(__dmul_rn(ocx, dx), __dmul_rn(ocy, dy)), __dmul_rn(ocz, dz));
    double c = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(ocx, ocx), __dmul_rn(ocy, ocy)), __dmul_rn(ocz, ocz)),
                         __dmul_rn(rad, rad));
    double disc = __dsub_rn(__dmul_rn(half_b, half_b), __dmul_rn(a, c));
    if (disc < 0.0) return;
    double sq = __dsqrt_rn(disc);
    float t = (float)__ddiv_rn(__dsub_rn(-half_b, sq), a);
    if (!(t >= kTMin && t <= h.t)) t = (float)__ddiv_rn(__dadd_rn(-half_b, sq), a);
    if (!(t >= kTMin && t <= h.t)) return;
    consider(h, t, prim, slot);
}

// cam: this ray is a camera ray (table route for triangles)
template <bool TRI>
__device__ __forceinline__ void test_prim(const SceneView& sc, int slot, const Ray& r, Hit& h, bool cam) {
    if (TRI) {
        if (cam) test_cam_tri_packet(sc.cam_prims, slot, r, h);
        else test_tri_mt(sc, slot, r, h);
    } else {
        // v1 Sphere::hit in double on the float32 ray / sphere, roots rounded to float32
        float4 s = __ldg(sc.prims + slot);
        double ocx = __dsub_rn((double)r.ox, (double)s.x), ocy = __dsub_rn((double)r.oy, (double)s.y),
               ocz = __dsub_rn((double)r.oz, (double)s.z);
        double dx = r.dx, dy = r.dy, dz = r.dz, rad = s.w;
        double a = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        double half_b = __dadd_rn(__dadd_rn(__dmul_rn(ocx, dx), __dmul_rn(ocy, dy)), __dmul_rn(ocz, dz));
        double c = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(ocx, ocx), __dmul_rn(ocy, ocy)), __dmul_rn(ocz, ocz)),
                             __dmul_rn(rad, rad));
        double disc = __dsub_rn(__dmul_rn(half_b, half_b), __dmul_rn(a, c));
        if (disc < 0.0) return;
        double sq = __dsqrt_rn(disc);
        float t = (float)__ddiv_rn(__dsub_rn(-half_b, sq), a);
        if (!(t >= kTMin && t <= h.t)) t = (float)__ddiv_rn(__dadd_rn(-half_b, sq), a);
        if (!(t >= kTMin && t <= h.t)) return;
        consider(h, t, __ldg(sc.slot_prim + slot), slot);
    }
}

// Device node records (written by the upload in rt_api.cu from the host rt_bvh_node array):
// float4 (bmin | code), float4 (bmax | 0) with code >= 0: internal, index of the child pair (children at
// code, code + 1; a pair is 64-byte aligned); code <= -2: leaf, ~code = (first_slot << 3) | count.
constexpr int kDone = -1;

// Closest hit over the flattened BVH: sibling pairs fetched as 2 x LDG.256 (ldg_pair: ld.global.nc.v8.f32),
// nearer child first, farther child pushed with its entry distance, dropped on pop when that
// distance exceeds the closest hit.  (Same visiting order as oracle MODE_NEAR_FIRST, so the
// node / primitive counters agree exactly with the CPU checker.)
template <bool TRI, bool STATS>
__device__ __forceinline__ void intersect(const SceneView& sc, const Ray& r, Hit& h, Counters& cnt, bool cam) {
    h.t = kTMax; h.prim = -1; h.slot = -1;
    if (sc.n_nodes == 0) return;
    float4 lo, hi;
    root_node(sc.nodes, lo, hi);
    float tn;
    if (STATS) cnt.nodes += 1;
    if (!box_hit(lo, hi, r, kTMin, h.t, tn)) return;
    int cur = __float_as_int(lo.w);
    int stack_code[kStackDepth];
    float stack_tn[kStackDepth];
    int sp = 0;
    for (;;) {
        if (cur >= 0) {
            PairRec q;
            ldg_pair(sc.nodes + 2 * (size_t)cur, q);
            if (STATS) cnt.nodes += 2;
            float tl, tr;
            bool hl, hr;
            pair_hit(q, r, kTMin, h.t, hl, hr, tl, tr);
            int lc = pair_lc(q), rc = pair_rc(q);
            if (hl && hr) {
                if (tr < tl) { int c = lc; lc = rc; rc = c; float tf = tl; tl = tr; tr = tf; }
                stack_code[sp] = rc; stack_tn[sp] = tr; ++sp;
                cur = lc;
                continue;
            } else if (hl) { cur = lc; continue; }
            else if (hr) { cur = rc; continue; }
        } else {
            const int code = ~cur;
            const int first = code >> 3, count = code & 7;
            for (int k = 0; k < count; ++k) {
                if (STATS) cnt.prims += 1;
                test_prim<TRI>(sc, first + k, r, h, cam);
            }
        }
        bool found = false;
        while (sp > 0) {
            --sp;
            if (stack_tn[sp] <= h.t) { cur = stack_code[sp]; found = true; break; }
        }
        if (!found) break;
    }
}

// ------------------------------------------------------------------------------- resumable traversal
// The same walk as intersect(), cut into a per-lane state (Trav + a local-memory stack) so that a
// warp can leave the traversal loop when too few of its lanes still have work, shade / refill the
// finished lanes, and come back.  Node codes as above, kDone = -1 = no node.  The per-ray visiting
// order is exactly intersect()'s.

struct Trav { int cur; int sp; Hit h; };

// Per-lane traversal stacks.  LocalStack: 64 entries of local memory (two word-interleaved arrays).  HybridStack<D>:
// the first D levels live in SHARED memory (uint2 entry, [level][thread]: conflict-free whatever level each lane is
// at), deeper levels fall back to local memory.  Local-memory stack traffic goes through the L1 tag stage like any
// global access -- in the incoherent-bounce kernel it was 18 % of all L1 sectors at a 32 % hit rate (the stack
// lines and the node lines evict each other), and every pop sits on the ray's critical path -- shared memory has
// neither problem.
struct LocalStack {
    int* code; float* tn;
    __device__ __forceinline__ void put(int i, int c, float t) const { code[i] = c; tn[i] = t; }
    __device__ __forceinline__ void get(int i, int& c, float& t) const { t = tn[i]; c = code[i]; }
};
template <int D, int THREADS>
struct HybridStack {
    uint2* s;                      // &smem[0][threadIdx.x]; level stride = THREADS entries
    int* code; float* tn;          // levels >= D
    __device__ __forceinline__ void put(int i, int c, float t) const {
        if (i < D) s[i * THREADS] = make_uint2((unsigned)c, __float_as_uint(t));
        else { code[i - D] = c; tn[i - D] = t; }
    }
    __device__ __forceinline__ void get(int i, int& c, float& t) const {
        if (i < D) { const uint2 e = s[i * THREADS]; c = (int)e.x; t = __uint_as_float(e.y); }
        else { t = tn[i - D]; c = code[i - D]; }
    }
};
template <class STACK>
__device__ __forceinline__ void trav_pop(Trav& tv, const STACK& st) {
    while (tv.sp > 0) {
        --tv.sp;
        int c; float t;
        st.get(tv.sp, c, t);
        if (t <= tv.h.t) { tv.cur = c; return; }
    }
    tv.cur = kDone;
}

// two_t > 0: the kernel stages the top treelet in shared memory (see packet_walk): the root's children are node records 0, 1 of it
template <bool STATS>
__device__ __forceinline__ void trav_begin(const SceneView& sc, const Ray& r, Trav& tv, Counters& cnt, int two_t = 0) {
    tv.h.t = kTMax; tv.h.prim = -1; tv.h.slot = -1;
    tv.sp = 0; tv.cur = kDone;
    if (sc.n_nodes == 0) return;
    float4 lo, hi;
    root_node(sc.nodes, lo, hi);
    float tn;
    if (STATS) cnt.nodes += 1;
    if (box_hit(lo, hi, r, kTMin, tv.h.t, tn)) tv.cur = two_t > 0 ? 0 : __float_as_int(lo.w);
}

// Runs the warp's traversals until fewer than `min_active` lanes still have work (warp-uniform call).
// Phase voting: in every iteration the warp executes EITHER one internal-node step (sibling pair
// fetch + two slab tests) OR one leaf step (its <= 4 primitives), whichever the vote picks, and only
// the lanes currently in that phase run it; the others wait for their phase's turn.  Lanes thereby
// regroup by phase (software re-convergence) instead of idling through each other's inner loops --
// the per-ray visiting order, and with it every counter, stays exactly intersect()'s.
// leaf_vote: the leaf phase runs when at least that many lanes hold a leaf (or no lane holds an
// internal node).
// A lane that has finished a leaf, or whose internal step found no child to enter, does not pop on the spot: it
// parks in state kNeedPop and all such lanes pop TOGETHER at the top of the next iteration (one convergent copy
// of the pop loop instead of two divergent ones that ran for 1-4 lanes at a time: -11 % instructions).
// CAM: 0 = no ray of this warp is a camera ray, 1 = all are, 2 = per lane (`cam`).
constexpr int kNeedPop = (int)0x80000000;   // not a leaf code: would mean first slot 2^28 - 1, count 7

// ------------------------------------------------------------------------------- cooperative leaf step (option "qnodes" bit 2)
// In the leaf phase of trav_run only the lanes that hold a leaf work -- about 10 of 32 in the incoherent bounces -- and each
// loops over its own <= 4 triangles.  Here the first EIGHT leaf-holding lanes (by lane number; the others keep their leaf for
// the next phase) hand their <= 4 triangles to the whole warp: owner number j posts (lane, count, first slot) in a per-warp
// table in shared memory, lane p works for owner p / 4 on its triangle p % 4: it reads the owner's ray from shared memory
// (every lane keeps a copy of its ray there, k_wf_trace writes it at refill), runs the Moller-Trumbore test of
// test_tri_mt_records / tri_accept operation by operation, and on a hit folds (distance bits, primitive number, index in the
// leaf) into the owner's 64-bit best key with one shared-memory atomicMin -- distances are positive floats, so the smallest
// key is the closest hit, ties to the lower primitive number: consider()'s own rule, hence the same closest hit bit for bit
// whatever the order.  The owner then applies that one candidate to its closest hit.  One pass of ~24 busy lanes replaces
// 3-4 passes of ~10.
struct CoopWarp { float ray[6][32]; uint2 map[8]; unsigned long long best[32]; };   // per warp, shared memory
// The area is addressed through ONE register, its 32-bit shared-memory address (`sa`, kept live by k_wf_trace), with explicit
// ld / st / red.shared: through a generic pointer the compiler re-derives the address from %tid and the shared window at every
// use (3 x 8 instructions per leaf phase at 32 lanes, two of them S2R).
constexpr unsigned kCoopMap = 768u, kCoopBest = 832u;                                // byte offsets of map / best in CoopWarp
__device__ __forceinline__ void coop_store_ray(unsigned sa, int lane, const float4& o, const float4& d) {
    const unsigned a = sa + 4u * (unsigned)lane;
    sts_f32(a, o.x); sts_f32(a + 128u, o.y); sts_f32(a + 256u, o.z); sts_f32(a + 384u, d.x); sts_f32(a + 512u, d.y); sts_f32(a + 640u, d.z);
}

__device__ __forceinline__ void leaf_phase_coop(const SceneView& sc, Trav& tv, bool is_leaf, int lane, unsigned sa) {
    const unsigned lb = __ballot_sync(0xffffffffu, is_leaf);
    const int rank = __popc(lb & ((1u << lane) - 1u));
    const bool mine = is_leaf && rank < 8;
    const int code = ~tv.cur;
    const int cnt = code & 7, first = code >> 3;
    if (mine) {
        sts_v2(sa + kCoopMap + 8u * (unsigned)rank, (unsigned)lane | ((unsigned)cnt << 5), (unsigned)first);
        sts_v2(sa + kCoopBest + 8u * (unsigned)lane, 0xffffffffu, 0xffffffffu);
    }
    __syncwarp();
    const int n_own = min(__popc(lb), 8), k = lane & 3;
    if ((lane >> 2) < n_own) {
        const uint2 m = lds_v2(sa + kCoopMap + 8u * (unsigned)(lane >> 2));
        if (k < (int)(m.x >> 5)) {
            const unsigned owner = m.x & 31u;
            const int slot = (int)m.y + k;
            const float4* p = sc.prims + kTriStride * (size_t)slot;
            const float4 v0 = __ldg(p), e1 = __ldg(p + 1), e2 = __ldg(p + 2);
            const unsigned ra = sa + 4u * owner;
            const float ox = lds_f32(ra), oy = lds_f32(ra + 128u), oz = lds_f32(ra + 256u);
            const float dx = lds_f32(ra + 384u), dy = lds_f32(ra + 512u), dz = lds_f32(ra + 640u);
            float px, py, pz, qx, qy, qz;
            cross3(dx, dy, dz, e2.x, e2.y, e2.z, px, py, pz);
            float det = dot3(e1.x, e1.y, e1.z, px, py, pz);
            const float sx = __fsub_rn(ox, v0.x), sy = __fsub_rn(oy, v0.y), sz = __fsub_rn(oz, v0.z);
            float un = dot3(sx, sy, sz, px, py, pz);
            cross3(sx, sy, sz, e1.x, e1.y, e1.z, qx, qy, qz);
            float vn = dot3(dx, dy, dz, qx, qy, qz);
            const float c = dot3(e2.x, e2.y, e2.z, qx, qy, qz);
            const float sg = det < 0.0f ? -1.0f : 1.0f;
            det = __fmul_rn(det, sg); un = __fmul_rn(un, sg); vn = __fmul_rn(vn, sg);
            if (det > 0.0f && un >= 0.0f && vn >= 0.0f && __fadd_rn(un, vn) <= det) {
                const float t = __fdiv_rn(__fmul_rn(c, sg), det);
                if (t >= kTMin)
                    reds_min_u64(sa + kCoopBest + 8u * owner, ((unsigned long long)__float_as_uint(t) << 32) |
                                                              (unsigned long long)(((unsigned)__float_as_int(v0.w) << 2) | (unsigned)k));
            }
        }
    }
    __syncwarp();
    if (mine) {
        const uint2 key = lds_v2(sa + kCoopBest + 8u * (unsigned)lane);                 // .x = primitive << 2 | index in leaf, .y = distance bits
        if ((key.x & key.y) != 0xffffffffu)
            consider(tv.h, __uint_as_float(key.y), (int)(key.x >> 2), first + (int)(key.x & 3u));
        if (cnt > 4) {                                                          // leaves of more than 4 (no builder here makes them)
            const unsigned a = sa + 4u * (unsigned)lane;
            Ray own;
            own.ox = lds_f32(a); own.oy = lds_f32(a + 128u); own.oz = lds_f32(a + 256u);
            own.dx = lds_f32(a + 384u); own.dy = lds_f32(a + 512u); own.dz = lds_f32(a + 640u);
            for (int j = 4; j < cnt; ++j) test_tri_mt(sc, first + j, own, tv.h);
        }
        tv.cur = kNeedPop;
    }
}

// QM (option "qnodes"): bit 0 = internal steps read the compressed pairs (SceneView::qnodes, `qr` = the ray against their grid),
// bit 1 = triangle records as one 256-bit + one 128-bit load (SceneView::tri_a / tri_b) instead of three 128-bit ones,
// bit 2 = cooperative leaf step (leaf_phase_coop; triangles, no camera rays; `coop_sa` = shared-memory address of this warp's CoopWarp, `lane`).
template <bool TRI, bool STATS, int CAM, class STACK, bool TREELET = false, int QM = 0>
__device__ __forceinline__ void trav_run(const SceneView& sc, const Ray& r, Trav& tv, const STACK& st, int min_active,
                                         int leaf_vote, Counters& cnt, bool cam, const float4* s_tree = nullptr, int two_t = 0,
                                         const QRay* qr = nullptr, unsigned coop_sa = 0u, int lane = 0) {
    for (;;) {
        if (tv.cur == kNeedPop) trav_pop(tv, st);
        const bool is_int = tv.cur >= 0, is_leaf = tv.cur < kDone;
        const int n_int = __popc(__ballot_sync(0xffffffffu, is_int));
        const int n_leaf = __popc(__ballot_sync(0xffffffffu, is_leaf));
        if (n_int + n_leaf < min_active) break;
        if (TRI && CAM == 0 && !STATS && (QM & 4) && (n_int == 0 || n_leaf >= leaf_vote)) {
            leaf_phase_coop(sc, tv, is_leaf, lane, coop_sa);
        } else if (n_int == 0 || n_leaf >= leaf_vote) {
            if (is_leaf) {
                int code = ~tv.cur;
                int first = code >> 3, count = code & 7;
                for (int k = 0; k < count; ++k) {
                    if (STATS) cnt.prims += 1;
                    if (TRI && CAM == 0 && (QM & 2)) {
                        float4 v0, e1;
                        ldg_node(sc.tri_a + 2 * (size_t)(first + k), v0, e1);
                        test_tri_mt_records(v0, e1, __ldg(sc.tri_b + first + k), first + k, r, tv.h);
                    } else test_prim<TRI>(sc, first + k, r, tv.h, CAM == 2 ? cam : CAM == 1);
                }
                tv.cur = kNeedPop;
            }
        } else if (is_int && (QM & 1)) {
            uint4 w0, w1;
            ldg_u8(sc.qnodes + tv.cur, w0, w1);                 // pair m = nodes 2m, 2m + 1 = 2 x uint4 at index 2m = cur
            int lc = (int)w1.z, rc = (int)w1.w;
            float tl, tr;
            bool hl, hr;
            pair_hit_q(w0, w1, *qr, kTMin, tv.h.t, hl, hr, tl, tr);
            if (hl && hr) {
                if (tr < tl) { int c = lc; lc = rc; rc = c; float tf = tl; tl = tr; tr = tf; }
                st.put(tv.sp, rc, tr); ++tv.sp;
                tv.cur = lc;
            } else tv.cur = hl ? lc : (hr ? rc : kNeedPop);
        } else if (is_int) {
            PairRec q;
            int lc, rc;
            if (TREELET && tv.cur < two_t) {                  // per lane: top of the tree out of shared memory
                const float4* p = s_tree + 2 * tv.cur;
                q.a = p[0]; q.b = p[1]; q.c = p[2]; q.d = p[3];
                lc = pair_lc(q); rc = pair_rc(q);
            } else {
                ldg_pair(sc.nodes + 2 * (size_t)(TREELET ? tv.cur - two_t : tv.cur), q);
                lc = pair_lc(q); rc = pair_rc(q);
                if (TREELET) { lc = lc >= 0 ? lc + two_t : lc; rc = rc >= 0 ? rc + two_t : rc; }
            }
            if (STATS) cnt.nodes += 2;
            float tl, tr;
            bool hl, hr;
            pair_hit(q, r, kTMin, tv.h.t, hl, hr, tl, tr);
            if (hl && hr) {
                if (tr < tl) { int c = lc; lc = rc; rc = c; float tf = tl; tl = tr; tr = tf; }
                st.put(tv.sp, rc, tr); ++tv.sp;
                tv.cur = lc;
            } else tv.cur = hl ? lc : (hr ? rc : kNeedPop);
        }
    }
}

// ------------------------------------------------------------------------------- packet traversal
// One warp = one packet of 32 camera rays (an 8x4 pixel block, one shared origin).  The warp walks
// the UNION of its lanes' traversals with a single stack: at an internal node every lane slab-tests
// both children against its own closest hit, ballots decide which children are entered at all and
// which first (each lane votes for its nearer hit child), the other child is pushed with the
// smallest entry distance of any lane (CREDUX.MIN on the float bits, distances are positive) and is
// dropped on pop when that exceeds every lane's closest hit (CREDUX.MAX).  Node and primitive
// addresses are warp-uniform, so each fetch is one broadcast request, no lane ever idles in another
// lane's phase, and no per-lane stack exists: the warp's stack is 64 x 8 bytes of shared memory that
// every lane writes / reads at the same address (no bank conflict, no synchronisation needed: a lane
// reads back what it wrote itself).
// Results are the per-ray traversal's bit for bit: every lane tests a superset of the primitives its
// own walk would test, and closest-hit selection is order independent (consider()).
// cnt.nodes / cnt.prims count what the PACKET fetched (lane 0 only), not per-ray visits; `work` is the
// packet's step count (internal steps + primitives tested), the scheduler's cost measure.
// inactive lanes (pixels outside the frame) carry closest hit 0 and never enter a box.
//
// OCT in 0..7: every lane's direction has the sign pattern OCT (bit k set = component k negative) and
// no component is tiny, so the near / far plane of each slab is known at compile time and the slab
// test needs no per-axis min / max (4 instead of 10 FMNMX-pipe instructions per box; that pipe, not
// FFMA, limits the traversal).  fmaf is monotonic, so picking the plane by sign gives exactly the
// value fminf / fmaxf would pick: same bits as box_hit().  OCT = 8: generic test.  Both children of a pair are tested at once
// (pair_hit_oct): their slab products are packed fmas over the interleaved pair record.
template <int OCT>
__device__ __forceinline__ void pair_hit_oct(const PairRec& q, const Ray& r, float tlo, float thi, bool& hl, bool& hr, float& tl, float& tr) {
    if (OCT == 8) { pair_hit(q, r, tlo, thi, hl, hr, tl, tr); return; }
    const float2 ix = make_float2(r.ix, r.ix), iy = make_float2(r.iy, r.iy), iz = make_float2(r.iz, r.iz);
    const float2 ax = make_float2(-r.ax, -r.ax), ay = make_float2(-r.ay, -r.ay), az = make_float2(-r.az, -r.az);
    const float2 lox = make_float2(q.a.x, q.a.y), loy = make_float2(q.a.z, q.a.w), loz = make_float2(q.b.x, q.b.y);
    const float2 hix = make_float2(q.c.x, q.c.y), hiy = make_float2(q.c.z, q.c.w), hiz = make_float2(q.d.x, q.d.y);
    const float2 nx = __ffma2_rn((OCT & 1) ? hix : lox, ix, ax), fx = __ffma2_rn((OCT & 1) ? lox : hix, ix, ax);
    const float2 ny = __ffma2_rn((OCT & 2) ? hiy : loy, iy, ay), fy = __ffma2_rn((OCT & 2) ? loy : hiy, iy, ay);
    const float2 nz = __ffma2_rn((OCT & 4) ? hiz : loz, iz, az), fz = __ffma2_rn((OCT & 4) ? loz : hiz, iz, az);
    const float nl = fmaxf(fmaxf(nx.x, ny.x), fmaxf(nz.x, tlo)), fl = fminf(fminf(fx.x, fy.x), fminf(fz.x, thi));
    const float nr = fmaxf(fmaxf(nx.y, ny.y), fmaxf(nz.y, tlo)), fr = fminf(fminf(fx.y, fy.y), fminf(fz.y, thi));
    tl = nl; tr = nr;
    hl = nl <= fl; hr = nr <= fr;
}

// TREELET (shared-memory staging of the top treelet, north_star): s_tree = the top levels of the tree staged in shared
// memory by the CTA (SceneView::treelet: a heap of sibling pairs whose child codes are already in the kernel's code space:
// code < two_t = node record of the staged heap, code >= two_t = global node record code - two_t).  The packet's node
// address is warp-uniform, so the shared / global choice is a uniform branch.
template <bool TRI, bool STATS, int OCT, bool TREELET = false>
__device__ __forceinline__ void packet_walk(const SceneView& sc, const float4* __restrict__ cam_prims, const Ray& r, int lane,
                                            uint2* __restrict__ stack, int cur, Hit& h, Counters& cnt, int& work,
                                            const float4* s_tree = nullptr, int two_t = 0) {
    int sp = 0;
    for (;;) {
        if (cur >= 0) {
            PairRec q;
            int lc, rc;
            if (TREELET && cur < two_t) {
                const float4* p = s_tree + 2 * cur;
                q.a = p[0]; q.b = p[1]; q.c = p[2]; q.d = p[3];
                lc = pair_lc(q); rc = pair_rc(q);
            } else {
                ldg_pair(sc.nodes + 2 * (size_t)(TREELET ? cur - two_t : cur), q);
                lc = pair_lc(q); rc = pair_rc(q);
                if (TREELET) { lc = lc >= 0 ? lc + two_t : lc; rc = rc >= 0 ? rc + two_t : rc; }
            }
            if (STATS && lane == 0) cnt.nodes += 2;
            work += 1;
            float tl, tr;
            bool hl, hr;
            pair_hit_oct<OCT>(q, r, kTMin, h.t, hl, hr, tl, tr);
            const unsigned bl = __ballot_sync(0xffffffffu, hl), br = __ballot_sync(0xffffffffu, hr);
            if (bl != 0u && br != 0u) {
                // entry distances with +inf for a missed child: a lane votes "right first" iff tr' < tl'
                const float tl2 = hl ? tl : __int_as_float(0x7f800000), tr2 = hr ? tr : __int_as_float(0x7f800000);
                const unsigned vr = __ballot_sync(0xffffffffu, tr2 < tl2);
                const bool right_first = 2 * __popc(vr) > __popc(bl | br);
                const unsigned tf = __reduce_min_sync(0xffffffffu, __float_as_uint(right_first ? tl2 : tr2));
                stack[sp] = make_uint2((unsigned)(right_first ? lc : rc), tf);
                ++sp;
                cur = right_first ? rc : lc;
                continue;
            } else if (bl != 0u) { cur = lc; continue; }
            else if (br != 0u) { cur = rc; continue; }
        } else {
            const int code = ~cur;
            const int first = code >> 3, count = code & 7;
            if (STATS && lane == 0) cnt.prims += count;
            work += count;
            for (int k = 0; k < count; ++k) {
                if (TRI) test_cam_tri_packet(cam_prims, first + k, r, h);
                else test_prim<false>(sc, first + k, r, h, true);
            }
        }
        const unsigned max_t = __reduce_max_sync(0xffffffffu, __float_as_uint(h.t));
        bool found = false;
        while (sp > 0) {
            --sp;
            const uint2 e = stack[sp];
            if (e.y <= max_t) { cur = (int)e.x; found = true; break; }
        }
        if (!found) break;
    }
}

template <bool TRI, bool STATS, bool TREELET = false>
__device__ __forceinline__ void packet_intersect(const SceneView& sc, const float4* __restrict__ cam_prims, const Ray& r,
                                                 bool active, int lane, uint2* __restrict__ stack, Hit& h, Counters& cnt,
                                                 int& work, const float4* s_tree = nullptr) {
    h.t = active ? kTMax : 0.0f; h.prim = -1; h.slot = -1;
    if (sc.n_nodes == 0) return;
    float4 lo, hi;
    root_node(sc.nodes, lo, hi);
    float tn;
    if (STATS && lane == 0) cnt.nodes += 1;
    const bool hit = box_hit(lo, hi, r, kTMin, h.t, tn);
    if (!__any_sync(0xffffffffu, hit)) return;
    const int two_t = TREELET ? sc.treelet_two_t : 0;
    const int root = (TREELET && two_t > 0) ? 0 : __float_as_int(lo.w);          // staged heap: the root's children are pair 0
    // octant of the packet: sign bits of the direction; usable when all lanes agree, no component is
    // tiny (so 1/d, o/d and every slab product stay finite) and the scene is of sane extent
    const int oct = (r.dx < 0.0f ? 1 : 0) | (r.dy < 0.0f ? 2 : 0) | (r.dz < 0.0f ? 4 : 0);
    const bool tiny = !(fabsf(r.dx) >= 0x1p-60f && fabsf(r.dy) >= 0x1p-60f && fabsf(r.dz) >= 0x1p-60f);
    const int oct0 = __shfl_sync(0xffffffffu, oct, 0);
    const bool uniform = sc.sane_extent && __all_sync(0xffffffffu, oct == oct0 && !tiny);
    if (!uniform) { packet_walk<TRI, STATS, 8, TREELET>(sc, cam_prims, r, lane, stack, root, h, cnt, work, s_tree, two_t); return; }
    switch (oct0) {
        case 0: packet_walk<TRI, STATS, 0, TREELET>(sc, cam_prims, r, lane, stack, root, h, cnt, work, s_tree, two_t); break;
        case 1: packet_walk<TRI, STATS, 1, TREELET>(sc, cam_prims, r, lane, stack, root, h, cnt, work, s_tree, two_t); break;
        case 2: packet_walk<TRI, STATS, 2, TREELET>(sc, cam_prims, r, lane, stack, root, h, cnt, work, s_tree, two_t); break;
        case 3: packet_walk<TRI, STATS, 3, TREELET>(sc, cam_prims, r, lane, stack, root, h, cnt, work, s_tree, two_t); break;
        case 4: packet_walk<TRI, STATS, 4, TREELET>(sc, cam_prims, r, lane, stack, root, h, cnt, work, s_tree, two_t); break;
        case 5: packet_walk<TRI, STATS, 5, TREELET>(sc, cam_prims, r, lane, stack, root, h, cnt, work, s_tree, two_t); break;
        case 6: packet_walk<TRI, STATS, 6, TREELET>(sc, cam_prims, r, lane, stack, root, h, cnt, work, s_tree, two_t); break;
        default: packet_walk<TRI, STATS, 7, TREELET>(sc, cam_prims, r, lane, stack, root, h, cnt, work, s_tree, two_t); break;
    }
}

// ------------------------------------------------------------------------------- shading
// SMEM: the SceneView's prims / mats pointers point into SHARED memory (the tiny-scene kernel stages the whole scene
// there, rt_tiny.cu): plain loads (LDS) instead of ld.global.nc.
template <bool SMEM>
__device__ __forceinline__ float4 ld4(const float4* p) { return SMEM ? *p : __ldg(p); }

template <bool TRI, bool SMEM = false>
__device__ __forceinline__ void shading_normal(const SceneView& sc, const Hit& h, const Ray& r, float px,
                                               float py, float pz, float& nx, float& ny, float& nz) {
    if (TRI) {
        const float4* p = sc.prims + kTriStride * (size_t)h.slot;
        float4 e1 = ld4<SMEM>(p + 1), e2 = ld4<SMEM>(p + 2);
        cross3(e1.x, e1.y, e1.z, e2.x, e2.y, e2.z, nx, ny, nz);
        normalize3(nx, ny, nz);
    } else {
        float4 s = ld4<SMEM>(sc.prims + h.slot);
        float inv = __fdiv_rn(1.0f, s.w);                           // raytracer_core.h:210
        nx = __fmul_rn(__fsub_rn(px, s.x), inv); ny = __fmul_rn(__fsub_rn(py, s.y), inv); nz = __fmul_rn(__fsub_rn(pz, s.z), inv);
    }
    if (!(dot3(r.dx, r.dy, r.dz, nx, ny, nz) < 0.0f)) { nx = -nx; ny = -ny; nz = -nz; }   // old/..core copy.h:132-135
}

__device__ __forceinline__ void unit_sphere(uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t k0,
                                            uint32_t k1, float& x, float& y, float& z) {
    for (uint32_t j = 0;; ++j) {                                      // old/raytracer_core copy.cpp:170-178
        uint4 o = philox4x32_10(pixel, sample, bounce, 1u + j, k0, k1);
        x = __fmaf_rn(2.0f, u01(o.x), -1.0f); y = __fmaf_rn(2.0f, u01(o.y), -1.0f); z = __fmaf_rn(2.0f, u01(o.z), -1.0f);
        if (dot3(x, y, z, x, y, z) < 1.0f || j == 255u) return;
    }
}

// Scatter at a hit (both integrators).  Returns false when the path ends here.
// ctl = the control draws of this bounce: .z Russian roulette, .w metal selection.
template <bool TRI, bool SMEM = false>
__device__ __forceinline__ bool scatter(const SceneView& sc, const Hit& h, Ray& r, int integrator, int b,
                                        int max_depth, uint4 ctl, float4 m0, float4 m1, uint32_t pixel,
                                        uint32_t sample, uint32_t k0, uint32_t k1, float& tr, float& tg, float& tb) {
    bool metal;
    if (integrator == 0) {
        int remaining = max_depth - b;                                  // v1 counts depth down
        if (!(remaining < 3 || u01(ctl.z) < 0.8f)) return false;
        metal = u01(ctl.w) < m0.w;
    } else {
        int depth = b + 1;                                              // v2 counts depth up
        if (depth > 3) {
            float mc = (tr > tg) ? (tr > tb ? tr : tb) : (tg > tb ? tg : tb);
            float p = (mc > 0.95f) ? 0.95f : mc;
            if (p < 0.1f) p = 0.1f;
            if (u01(ctl.z) >= p) return false;
            float ip = __fdiv_rn(1.0f, p);
            tr = __fmul_rn(tr, ip); tg = __fmul_rn(tg, ip); tb = __fmul_rn(tb, ip);
        }
        metal = m0.w > 0.0f;
    }
    float px = __fmaf_rn(r.dx, h.t, r.ox), py = __fmaf_rn(r.dy, h.t, r.oy), pz = __fmaf_rn(r.dz, h.t, r.oz);
    float nx, ny, nz;
    shading_normal<TRI, SMEM>(sc, h, r, px, py, pz, nx, ny, nz);
    float ux, uy, uz;
    unit_sphere(pixel, sample, (uint32_t)b, k0, k1, ux, uy, uz);
    float dx, dy, dz;
    if (metal) {
        float k2 = __fmul_rn(2.0f, dot3(r.dx, r.dy, r.dz, nx, ny, nz));
        float rx = __fmaf_rn(-k2, nx, r.dx), ry = __fmaf_rn(-k2, ny, r.dy), rz = __fmaf_rn(-k2, nz, r.dz);
        dx = __fmaf_rn(ux, m1.x, rx); dy = __fmaf_rn(uy, m1.x, ry); dz = __fmaf_rn(uz, m1.x, rz);
    } else {
        if (!(dot3(ux, uy, uz, nx, ny, nz) > 0.0f)) { ux = -ux; uy = -uy; uz = -uz; }
        dx = __fadd_rn(nx, ux); dy = __fadd_rn(ny, uy); dz = __fadd_rn(nz, uz);
    }
    tr = __fmul_rn(tr, m0.x); tg = __fmul_rn(tg, m0.y); tb = __fmul_rn(tb, m0.z);
    normalize3(dx, dy, dz);
    r = make_ray(px, py, pz, dx, dy, dz);
    return true;
}

template <bool TRI, bool SMEM = false>
__device__ __forceinline__ int material_row(const SceneView& sc, const Hit& h) {
    if (TRI) return __float_as_int(ld4<SMEM>(sc.prims + kTriStride * (size_t)h.slot + 1).w);
    return h.prim;
}

// One camera sample of pixel (i,j): radiance added into (cr,cg,cb).
template <bool TRI, bool STATS>
__device__ __forceinline__ void radiance(const SceneView& sc, const CameraBlock& cam, int i, int j,
                                         uint32_t pixel, uint32_t sample, int max_depth, int integrator,
                                         uint32_t k0, uint32_t k1, double inv_w, double inv_h, float& out_r,
                                         float& out_g, float& out_b, Counters& cnt) {
    uint4 ctl = philox4x32_10(pixel, sample, 0u, 0u, k0, k1);
    Ray r = camera_ray(cam, i, j, u01(ctl.x), u01(ctl.y), inv_w, inv_h);
    float cr = 0.0f, cg = 0.0f, cb = 0.0f, tr = 1.0f, tg = 1.0f, tb = 1.0f;
    for (int b = 0; b < max_depth; ++b) {
        Hit h;
        intersect<TRI, STATS>(sc, r, h, cnt, b == 0);
        if (STATS) cnt.segments += 1;
        if (h.prim < 0) {
            cr = __fmaf_rn(tr, sc.bg_r, cr); cg = __fmaf_rn(tg, sc.bg_g, cg); cb = __fmaf_rn(tb, sc.bg_b, cb);
            break;
        }
        const float4* mp = sc.mats + 2 * (size_t)material_row<TRI>(sc, h);
        float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
        cr = __fmaf_rn(tr, m1.y, cr); cg = __fmaf_rn(tg, m1.z, cg); cb = __fmaf_rn(tb, m1.w, cb);
        if (b + 1 == max_depth) break;
        if (b > 0) ctl = philox4x32_10(pixel, sample, (uint32_t)b, 0u, k0, k1);
        if (!scatter<TRI>(sc, h, r, integrator, b, max_depth, ctl, m0, m1, pixel, sample, k0, k1, tr, tg, tb)) break;
    }
    out_r = cr; out_g = cg; out_b = cb;
}

__device__ __forceinline__ float resolve1(float sum, float inv_spp) {   // raytracer_core.cpp:398-409
    float c = __fsqrt_rn(__fmul_rn(sum, inv_spp));
    return c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
}

}  // namespace b200rt
