/* oracle/rt_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, float32 CPU restatement of the reference's render hot path
 * (camera ray generation -> BVH traversal + ray/primitive intersection -> path-traced
 * shading -> sample accumulation / resolve).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product library
 * (libb200rt.so) never links, imports or calls anything in this directory.
 *
 * PARITY PINNING.  The reference ships no tests and no golden vectors, and
 * its current generation ("v2", cpp_raytracer/raytracer_core.{h,cpp}) neither compiles nor
 * renders.  This restatement is therefore pinned against outputs of the
 * reference itself: the runnable v1 generation (old/ *), compiled unmodified by
 * oracle/build_ref.sh into oracle/_ref/, whose primary-hit ids / distances / images are frozen
 * in tests/golden/ by tests/golden/make_golden.py.  Triangle scenes have no reference
 * counterpart (the reference only has spheres): for those, parity is "unpinned by the
 * reference" and rests on brute-force == BVH self-consistency plus a float64 numpy
 * Moller-Trumbore check in tests/.
 *
 * What follows which reference lines (all under /root/reference):
 *   camera basis + get_ray ........ old/raytracer_core copy.h:160-184 (v1 convention: target,
 *                                   ndc in [-1,1], aspect = W/H, pi = 3.14159, world-up basis)
 *                                   evaluated in float like cpp_raytracer/raytracer_core.h:266-271
 *   ray setup (normalise, 1/d) .... cpp_raytracer/raytracer_core.h:107-121
 *   AABB slab test ................ cpp_raytracer/raytracer_core.h:132-153 WITH the near/far
 *                                   ordering of old/bvh copy.cpp:9-25 (v2 forgot the swap)
 *   sphere test ................... cpp_raytracer/raytracer_core.h:192-215 / old/raytracer_core
 *                                   copy.cpp:21-52 (nearer root in [tmin,tmax], else farther),
 *                                   quadratic evaluated in double on the float32 ray so that the
 *                                   1e-5 bar against v1's doubles is met at grazing angles
 *   face-forward normal ........... old/raytracer_core copy.h:132-135
 *   BVH build ..................... cpp_raytracer/raytracer_core.cpp:57-118 (leaf <= 4, longest
 *                                   axis, median by centre) with consistent child links and a
 *                                   leaf flag that does not alias a child index
 *   BVH traversal (REF order) ..... cpp_raytracer/raytracer_core.cpp:198-243
 *   integrator 0 (v1) ............. old/raytracer_core copy.cpp:211-243
 *   integrator 1 (v2) ............. cpp_raytracer/raytracer_core.cpp:291-351
 *   unit-sphere / hemisphere ...... old/raytracer_core copy.cpp:170-192
 *   pixel loop, mean, sqrt, clamp . cpp_raytracer/raytracer_core.cpp:381-409
 * Deliberate, documented departures (DESIGN.md "Arithmetic contract"): Philox4x32-10
 * counter RNG instead of PCG32/mt19937; closest-hit ties go to the lower primitive index;
 * triangle primitive (Moller-Trumbore, see test_tri_cam / test_tri_mt) added; node boxes padded by 2^-16 * scene scale.
 *
 * Arithmetic contract (shared with the CUDA kernels, written independently there): IEEE
 * float32, no implicit contraction (compile with -ffp-contract=off, no -ffast-math), explicit
 * fmaf() exactly where written below, IEEE sqrtf and division.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } v3;

typedef struct {            /* shared 32-byte flattened node (include/b200rt.h: rt_bvh_node) */
    float bmin[3]; int32_t a;   /* internal: a = index of left child (right = a+1); leaf: first slot */
    float bmax[3]; int32_t b;   /* internal: 0; leaf: primitive count (1..4) */
} node_t;

typedef struct {
    /* primitives (upload order) */
    int is_tri;
    int64_t n;
    float* cr;        /* spheres: n x 4 */
    float* v0e;       /* triangles: n x 9 = v0, e1, e2 */
    float* v9;        /* triangles: n x 9 as uploaded (v0, v1, v2), for the boxes */
    int32_t* mat_id;  /* per primitive material row */
    int32_t* object_id;
    float* mats;      /* m x 8 */
    int m;
    float bg[3];
    /* camera block */
    float pos[3]; double fwd[3], right[3], up[3], sx, sy;   /* basis kept in double, see camera_ray */
    /* bvh */
    node_t* nodes;
    int64_t n_nodes;
    int32_t* prim_index;   /* slot -> primitive */
} scene_t;

/* ---------------------------------------------------------------- small math ---------- */
static inline float dot3(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline v3 cross3(v3 a, v3 b) {
    v3 r;
    r.x = fmaf(a.y, b.z, -(a.z * b.y));
    r.y = fmaf(a.z, b.x, -(a.x * b.z));
    r.z = fmaf(a.x, b.y, -(a.y * b.x));
    return r;
}
/* exactly antisymmetric cross product (both products rounded, then subtracted): cross_as(a,b) ==
 * -cross_as(b,a) bit for bit, which keeps the shared edge of two triangles of a fan watertight */
static inline v3 cross_as(v3 a, v3 b) {
    v3 r;
    r.x = a.y * b.z - a.z * b.y;
    r.y = a.z * b.x - a.x * b.z;
    r.z = a.x * b.y - a.y * b.x;
    return r;
}
static inline v3 sub3(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline v3 scale3(v3 a, float s) { v3 r = {a.x * s, a.y * s, a.z * s}; return r; }
static inline v3 normalize3(v3 a) {          /* raytracer_core.h:91-94 */
    float len = sqrtf(dot3(a, a));
    if (len > 0.0f) return scale3(a, 1.0f / len);
    v3 z = {0.0f, 0.0f, 1.0f};
    return z;
}
/* IEEE minNum/maxNum (a NaN operand is ignored), spelled out so the compiler cannot pick a
 * different NaN rule than the GPU's FMNMX. */
static inline float fmin_n(float a, float b) { return (a < b) ? a : ((b != b) ? a : b); }
static inline float fmax_n(float a, float b) { return (a > b) ? a : ((b != b) ? a : b); }

/* ---------------------------------------------------------------- Philox4x32-10 ------- */
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline float u01(uint32_t x) { return (float)(x >> 8) * 0x1p-24f; }   /* [0,1) */

void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

/* ---------------------------------------------------------------- scene --------------- */
scene_t* orc_scene_new(void) {
    scene_t* s = (scene_t*)calloc(1, sizeof(scene_t));
    s->bg[0] = s->bg[1] = s->bg[2] = 0.1f;   /* Scene::Scene(), old/raytracer_core copy.cpp:54 */
    return s;
}
static void free_prims(scene_t* s) {
    free(s->cr); free(s->v0e); free(s->v9); free(s->mat_id); free(s->object_id); free(s->mats);
    s->cr = s->v0e = s->v9 = s->mats = NULL; s->mat_id = s->object_id = NULL; s->n = 0; s->m = 0;
}
static void free_bvh(scene_t* s) {
    free(s->nodes); free(s->prim_index); s->nodes = NULL; s->prim_index = NULL; s->n_nodes = 0;
}
void orc_scene_free(scene_t* s) { if (!s) return; free_prims(s); free_bvh(s); free(s); }

void orc_set_spheres(scene_t* s, const float* cr, const float* mat8, const int32_t* object_id, int64_t n) {
    free_prims(s); free_bvh(s);
    s->is_tri = 0; s->n = n; s->m = (int)n;
    s->cr = (float*)malloc(sizeof(float) * 4 * (n ? n : 1));
    s->mats = (float*)malloc(sizeof(float) * 8 * (n ? n : 1));
    s->mat_id = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    s->object_id = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    memcpy(s->cr, cr, sizeof(float) * 4 * n);
    memcpy(s->mats, mat8, sizeof(float) * 8 * n);
    for (int64_t i = 0; i < n; ++i) { s->mat_id[i] = (int32_t)i; s->object_id[i] = object_id ? object_id[i] : (int32_t)i; }
}

void orc_set_triangles(scene_t* s, const float* v9, const int32_t* mat_id, int64_t n, const float* mats, int m) {
    free_prims(s); free_bvh(s);
    s->is_tri = 1; s->n = n; s->m = m;
    s->v0e = (float*)malloc(sizeof(float) * 9 * (n ? n : 1));
    s->v9 = (float*)malloc(sizeof(float) * 9 * (n ? n : 1));
    memcpy(s->v9, v9, sizeof(float) * 9 * n);
    s->mats = (float*)malloc(sizeof(float) * 8 * (m ? m : 1));
    s->mat_id = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    s->object_id = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    memcpy(s->mats, mats, sizeof(float) * 8 * m);
    for (int64_t i = 0; i < n; ++i) {
        const float* v = v9 + 9 * i;
        float* o = s->v0e + 9 * i;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
        o[3] = v[3] - v[0]; o[4] = v[4] - v[1]; o[5] = v[5] - v[2];   /* e1 = v1 - v0 */
        o[6] = v[6] - v[0]; o[7] = v[7] - v[1]; o[8] = v[8] - v[2];   /* e2 = v2 - v0 */
        s->mat_id[i] = mat_id ? mat_id[i] : 0;
        s->object_id[i] = (int32_t)i;
    }
}

void orc_set_background(scene_t* s, const float* rgb) { memcpy(s->bg, rgb, sizeof(float) * 3); }

/* Camera basis once per camera, in double, rounded to float (v1 recomputes it per ray in
 * double: old/raytracer_core copy.h:160-184).  cam = pos3 target3 up3 fov aspect.  Like the
 * reference, the basis is built from world up (0,1,0); cam.up is carried but not read. */
void orc_set_camera(scene_t* s, const double* cam) {
    double fx = cam[3] - cam[0], fy = cam[4] - cam[1], fz = cam[5] - cam[2];
    double fl = sqrt(fx * fx + fy * fy + fz * fz);
    if (fl > 0) { fx /= fl; fy /= fl; fz /= fl; }
    /* right = forward x (0,1,0) */
    double rx = fy * 0.0 - fz * 1.0, ry = fz * 0.0 - fx * 0.0, rz = fx * 1.0 - fy * 0.0;
    double rl = sqrt(rx * rx + ry * ry + rz * rz);
    if (rl > 0) { rx /= rl; ry /= rl; rz /= rl; }
    if (sqrt(rx * rx + ry * ry + rz * rz) < 0.001) { rx = 1; ry = 0; rz = 0; }
    /* up = right x forward */
    double ux = ry * fz - rz * fy, uy = rz * fx - rx * fz, uz = rx * fy - ry * fx;
    double ul = sqrt(ux * ux + uy * uy + uz * uz);
    if (ul > 0) { ux /= ul; uy /= ul; uz /= ul; }
    double tan_fov = tan(cam[9] * 3.14159 / 360.0);
    s->pos[0] = (float)cam[0]; s->pos[1] = (float)cam[1]; s->pos[2] = (float)cam[2];
    s->fwd[0] = fx; s->fwd[1] = fy; s->fwd[2] = fz;
    s->right[0] = rx; s->right[1] = ry; s->right[2] = rz;
    s->up[0] = ux; s->up[1] = uy; s->up[2] = uz;
    s->sx = cam[10] * tan_fov;
    s->sy = tan_fov;
}
/* out[14] = pos fwd right up sx sy (as doubles) */
void orc_get_camera_block(const scene_t* s, double* out) {
    for (int c = 0; c < 3; ++c) { out[c] = s->pos[c]; out[3 + c] = s->fwd[c]; out[6 + c] = s->right[c]; out[9 + c] = s->up[c]; }
    out[12] = s->sx; out[13] = s->sy;
}

/* ---------------------------------------------------------------- BVH build ----------- */
static void prim_box(const scene_t* s, int64_t i, float* lo, float* hi) {
    if (s->is_tri) {
        const float* p = s->v9 + 9 * i;                /* min/max over the uploaded vertices */
        for (int c = 0; c < 3; ++c) {
            float a = p[c], b = p[3 + c], d = p[6 + c];
            lo[c] = fminf(a, fminf(b, d)); hi[c] = fmaxf(a, fmaxf(b, d));
        }
    } else {
        const float* p = s->cr + 4 * i;
        for (int c = 0; c < 3; ++c) { lo[c] = p[c] - p[3]; hi[c] = p[c] + p[3]; }
    }
}

typedef struct { float key; int32_t idx; } sort_item;
static int cmp_item(const void* pa, const void* pb) {
    const sort_item* a = (const sort_item*)pa; const sort_item* b = (const sort_item*)pb;
    if (a->key < b->key) return -1;
    if (a->key > b->key) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx);
}
static int cmp_i32(const void* pa, const void* pb) {
    int32_t a = *(const int32_t*)pa, b = *(const int32_t*)pb; return (a > b) - (a < b);
}

typedef struct {
    const scene_t* s; float* lo; float* hi; float* ctr;   /* per primitive boxes / box centres */
    int32_t* idx; sort_item* tmp; node_t* nodes; int64_t next_pair;
} build_t;

static void range_box(const build_t* b, int64_t start, int64_t end, float* lo, float* hi) {
    for (int c = 0; c < 3; ++c) { lo[c] = INFINITY; hi[c] = -INFINITY; }
    for (int64_t k = start; k < end; ++k) {
        const float* l = b->lo + 3 * b->idx[k]; const float* h = b->hi + 3 * b->idx[k];
        for (int c = 0; c < 3; ++c) { if (l[c] < lo[c]) lo[c] = l[c]; if (h[c] > hi[c]) hi[c] = h[c]; }
    }
}

/* Fills nodes[at] for [start,end) and recurses; children pairs are allocated depth-first. */
static void build_rec(build_t* b, int64_t at, int64_t start, int64_t end) {
    node_t* nd = &b->nodes[at];
    range_box(b, start, end, nd->bmin, nd->bmax);
    int64_t span = end - start;
    if (span <= 4) {                                   /* raytracer_core.cpp:86 */
        qsort(b->idx + start, (size_t)span, sizeof(int32_t), cmp_i32);   /* canonical leaf order */
        nd->a = (int32_t)start; nd->b = (int32_t)span;
        return;
    }
    float ex = nd->bmax[0] - nd->bmin[0], ey = nd->bmax[1] - nd->bmin[1], ez = nd->bmax[2] - nd->bmin[2];
    int axis = 0;                                      /* raytracer_core.cpp:96-99 */
    if (ey > ex) axis = 1;
    if (ez > ey && ez > ex) axis = 2;
    for (int64_t k = start; k < end; ++k) { b->tmp[k].key = b->ctr[3 * b->idx[k] + axis]; b->tmp[k].idx = b->idx[k]; }
    qsort(b->tmp + start, (size_t)span, sizeof(sort_item), cmp_item);    /* raytracer_core.cpp:102-103 */
    for (int64_t k = start; k < end; ++k) b->idx[k] = b->tmp[k].idx;
    int64_t mid = start + span / 2;                    /* raytracer_core.cpp:105 */
    int64_t pair = b->next_pair; b->next_pair += 2;
    nd->a = (int32_t)pair; nd->b = 0;
    build_rec(b, pair, start, mid);
    build_rec(b, pair + 1, mid, end);
}

/* Builds the oracle's own BVH (returns node count incl. the pad record at index 1). */
int64_t orc_build_bvh(scene_t* s) {
    free_bvh(s);
    int64_t n = s->n;
    if (n == 0) return 0;
    build_t b; b.s = s;
    b.lo = (float*)malloc(sizeof(float) * 3 * n); b.hi = (float*)malloc(sizeof(float) * 3 * n);
    b.ctr = (float*)malloc(sizeof(float) * 3 * n);
    b.idx = (int32_t*)malloc(sizeof(int32_t) * n); b.tmp = (sort_item*)malloc(sizeof(sort_item) * n);
    b.nodes = (node_t*)calloc((size_t)(2 * n + 2), sizeof(node_t));
    for (int64_t i = 0; i < n; ++i) {
        prim_box(s, i, b.lo + 3 * i, b.hi + 3 * i);
        for (int c = 0; c < 3; ++c) b.ctr[3 * i + c] = (b.lo[3 * i + c] + b.hi[3 * i + c]) * 0.5f;
        b.idx[i] = (int32_t)i;
    }
    b.next_pair = 2;
    build_rec(&b, 0, 0, n);
    /* pad every box outward by 2^-16 * (largest |coordinate| of the root box) */
    float scale = 0.0f;
    for (int c = 0; c < 3; ++c) { scale = fmaxf(scale, fabsf(b.nodes[0].bmin[c])); scale = fmaxf(scale, fabsf(b.nodes[0].bmax[c])); }
    float pad = scale * 0x1p-16f;
    for (int64_t k = 0; k < b.next_pair; ++k) {
        if (k == 1) continue;
        for (int c = 0; c < 3; ++c) { b.nodes[k].bmin[c] -= pad; b.nodes[k].bmax[c] += pad; }
    }
    s->nodes = b.nodes; s->n_nodes = b.next_pair; s->prim_index = b.idx;
    free(b.lo); free(b.hi); free(b.ctr); free(b.tmp);
    return s->n_nodes;
}

/* Adopt a BVH built elsewhere (the product library's rt_get_bvh output) so both sides walk
 * the very same tree. */
void orc_set_bvh(scene_t* s, const void* nodes, int64_t n_nodes, const int32_t* prim_index) {
    free_bvh(s);
    s->nodes = (node_t*)malloc(sizeof(node_t) * (n_nodes ? n_nodes : 1));
    memcpy(s->nodes, nodes, sizeof(node_t) * n_nodes);
    s->n_nodes = n_nodes;
    s->prim_index = (int32_t*)malloc(sizeof(int32_t) * (s->n ? s->n : 1));
    memcpy(s->prim_index, prim_index, sizeof(int32_t) * s->n);
}
int64_t orc_get_bvh(const scene_t* s, void* nodes, int32_t* prim_index) {
    if (nodes) memcpy(nodes, s->nodes, sizeof(node_t) * s->n_nodes);
    if (prim_index && s->prim_index) memcpy(prim_index, s->prim_index, sizeof(int32_t) * s->n);
    return s->n_nodes;
}

/* ---------------------------------------------------------------- intersection -------- */
typedef struct { v3 o, d, inv, ood; } ray_t;
typedef struct { float t; int32_t prim; uint64_t n_node, n_prim; } hit_t;

/* 1/d of the slab test.  A direction component that is zero (an axis-parallel ray) or too small for 1/d to stay finite
 * gets +-2^80 instead: a power of two, so o * inv and the slab distances fmaf(plane, inv, -o * inv) = (plane - o) * 2^80
 * are exact, finite and sign-correct -- the ray is inside the slab iff lo <= o <= hi, as in the reference's
 * (plane - o) * (1/d) form (old/bvh copy.cpp:9-25) -- where inf would make o * inv - o * inv a NaN and lose the box. */
static inline float safe_inv(float d) { return fabsf(d) < 0x1p-80f ? copysignf(0x1p80f, d) : 1.0f / d; }

static inline ray_t make_ray(v3 o, v3 dir_unit) {     /* raytracer_core.h:113-115 */
    ray_t r; r.o = o; r.d = dir_unit;
    r.inv.x = safe_inv(dir_unit.x); r.inv.y = safe_inv(dir_unit.y); r.inv.z = safe_inv(dir_unit.z);
    r.ood.x = o.x * r.inv.x; r.ood.y = o.y * r.inv.y; r.ood.z = o.z * r.inv.z;
    return r;
}

/* slab test on [tlo, thi]; returns entry distance through *tn */
static inline int box_hit(const node_t* nd, const ray_t* r, float tlo, float thi, float* tn) {
    float x1 = fmaf(nd->bmin[0], r->inv.x, -r->ood.x), x2 = fmaf(nd->bmax[0], r->inv.x, -r->ood.x);
    float y1 = fmaf(nd->bmin[1], r->inv.y, -r->ood.y), y2 = fmaf(nd->bmax[1], r->inv.y, -r->ood.y);
    float z1 = fmaf(nd->bmin[2], r->inv.z, -r->ood.z), z2 = fmaf(nd->bmax[2], r->inv.z, -r->ood.z);
    float n = fmax_n(fmax_n(fmin_n(x1, x2), fmin_n(y1, y2)), fmax_n(fmin_n(z1, z2), tlo));
    float f = fmin_n(fmin_n(fmax_n(x1, x2), fmax_n(y1, y2)), fmin_n(fmax_n(z1, z2), thi));
    *tn = n;
    return n <= f;
}

/* candidate (t, prim) replaces the current closest hit? ties go to the lower primitive index */
static inline void consider(hit_t* h, float t, int32_t prim, float tmin) {
    if (!(t >= tmin && t <= h->t)) return;
    if (t < h->t || h->prim < 0 || prim < h->prim) { h->t = t; h->prim = prim; }
}

static inline void test_sphere(const scene_t* s, int32_t prim, const ray_t* r, float tmin, hit_t* h) {
    /* old/raytracer_core copy.cpp:21-43, evaluated in double exactly as v1 writes it, on the
     * float32 ray and float32 sphere; each root is rounded to float32 before the range test.
     * (A pure float32 quadratic cannot hold 1e-5 relative on the radius-100 ground sphere near
     * its horizon: |oc|^2 - r^2 cancels at ulp(1e4) ~ 1e-3.) */
    const float* p = s->cr + 4 * (int64_t)prim;
    double ocx = (double)r->o.x - (double)p[0], ocy = (double)r->o.y - (double)p[1], ocz = (double)r->o.z - (double)p[2];
    double dx = r->d.x, dy = r->d.y, dz = r->d.z, rad = p[3];
    double a = dx * dx + dy * dy + dz * dz;
    double half_b = ocx * dx + ocy * dy + ocz * dz;
    double c = (ocx * ocx + ocy * ocy + ocz * ocz) - rad * rad;
    double disc = half_b * half_b - a * c;
    if (disc < 0.0) return;
    double sqrtd = sqrt(disc);
    float t = (float)((-half_b - sqrtd) / a);
    if (!(t >= tmin && t <= h->t)) t = (float)((-half_b + sqrtd) / a);
    consider(h, t, prim, tmin);
}

/* Ray/triangle tests (an extension: the reference has no triangle primitive).  Two routes, the same rule as the
 * CUDA kernels (rt_device.cuh):
 *
 * CAMERA RAYS (bounce 0 of every path, orc_trace_primary): Moller-Trumbore written as scalar triple products of
 * the ray direction d with three vectors that depend only on the triangle and the ray ORIGIN:
 *   det = d.(e2 x e1),  u*det = d.(e2 x s),  v*det = d.(s x e1),  t*det = e2.(s x e1),  s = o - v0.
 * The kernels read the three vectors and the scalar from a per-camera table computed with these very operations.
 * The cross products are the exactly antisymmetric cross_as(): for two triangles that share v0 and an edge
 * vector (the two halves of a quad) u*det of one is exactly -v*det of the other, so a camera ray through the shared
 * edge is accepted by at least one of them (no cracks along quad diagonals).
 *
 * ANY OTHER RAY (bounces >= 1, orc_trace_rays): classic Moller-Trumbore, p = d x e2, q = s x e1, det = e1.p,
 * u*det = s.p, v*det = d.q, t*det = e2.q.
 *
 * Both: inside test division-free (compare against det after making det positive; multiplying by -1 is exact),
 * hit distance one IEEE division. */
static inline void tri_accept(hit_t* h, float det, float un, float vn, float c, int32_t prim, float tmin) {
    float sg = det < 0.0f ? -1.0f : 1.0f;
    det = det * sg; un = un * sg; vn = vn * sg;
    if (det > 0.0f && un >= 0.0f && vn >= 0.0f && un + vn <= det) consider(h, (c * sg) / det, prim, tmin);
}

static inline void test_tri_cam(const scene_t* s, int32_t prim, const ray_t* r, float tmin, hit_t* h) {
    const float* p = s->v0e + 9 * (int64_t)prim;
    v3 v0 = {p[0], p[1], p[2]}, e1 = {p[3], p[4], p[5]}, e2 = {p[6], p[7], p[8]};
    v3 sv = sub3(r->o, v0);
    v3 nn = cross_as(e2, e1);
    v3 av = cross_as(e2, sv);
    v3 bv = cross_as(sv, e1);
    tri_accept(h, dot3(r->d, nn), dot3(r->d, av), dot3(r->d, bv), dot3(e2, bv), prim, tmin);
}

static inline void test_tri_mt(const scene_t* s, int32_t prim, const ray_t* r, float tmin, hit_t* h) {
    const float* p = s->v0e + 9 * (int64_t)prim;
    v3 v0 = {p[0], p[1], p[2]}, e1 = {p[3], p[4], p[5]}, e2 = {p[6], p[7], p[8]};
    v3 pv = cross3(r->d, e2);
    float det = dot3(e1, pv);
    v3 sv = sub3(r->o, v0);
    float un = dot3(sv, pv);
    v3 qv = cross3(sv, e1);
    tri_accept(h, det, un, dot3(r->d, qv), dot3(e2, qv), prim, tmin);
}

static inline void test_prim(const scene_t* s, int32_t prim, const ray_t* r, float tmin, hit_t* h, int cam) {
    h->n_prim++;
    if (!s->is_tri) test_sphere(s, prim, r, tmin, h);
    else if (cam) test_tri_cam(s, prim, r, tmin, h);
    else test_tri_mt(s, prim, r, tmin, h);
}

enum { MODE_BRUTE = 0, MODE_REF_ORDER = 1, MODE_NEAR_FIRST = 2 };

/* cam: the ray is a camera ray (triangle route, see above) */
static void intersect(const scene_t* s, const ray_t* r, float tmin, float tmax, int mode, hit_t* h, int cam) {
    h->t = tmax; h->prim = -1;
    if (s->n == 0) return;
    if (mode == MODE_BRUTE || !s->nodes) {           /* Scene::hit brute force, old/..core copy.cpp:117-130 */
        for (int64_t i = 0; i < s->n; ++i) test_prim(s, (int32_t)i, r, tmin, h, cam);
        return;
    }
    if (mode == MODE_REF_ORDER) {
        /* raytracer_core.cpp:198-243: pop; box-test against [tmin, closest]; leaf -> test
         * primitives with the shrinking closest; internal -> push left then right. */
        int32_t stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp > 0) {
            const node_t* nd = &s->nodes[stack[--sp]];
            float tn; h->n_node++;
            if (!box_hit(nd, r, tmin, h->t, &tn)) continue;
            if (nd->b > 0) { for (int k = 0; k < nd->b; ++k) test_prim(s, s->prim_index[nd->a + k], r, tmin, h, cam); }
            else { stack[sp++] = nd->a; stack[sp++] = nd->a + 1; }
        }
        return;
    }
    /* MODE_NEAR_FIRST: the traversal order the CUDA kernel uses (DESIGN.md "Traversal"):
     * sibling pairs fetched together, nearer child first, farther child pushed with its
     * entry distance and dropped on pop when that distance exceeds the closest hit. */
    {
        const node_t* root = &s->nodes[0];
        float tn; h->n_node++;
        if (!box_hit(root, r, tmin, h->t, &tn)) return;
        struct { int32_t a, b; float tn; } stack[64]; int sp = 0;
        int32_t ca = root->a, cb = root->b;
        for (;;) {
            if (cb == 0) {
                const node_t* L = &s->nodes[ca]; const node_t* R = L + 1;
                float tl, tr; h->n_node += 2;
                int hl = box_hit(L, r, tmin, h->t, &tl), hr = box_hit(R, r, tmin, h->t, &tr);
                if (hl && hr) {
                    const node_t* nr = L; const node_t* fr = R; float tf = tr;
                    if (tr < tl) { nr = R; fr = L; tf = tl; }
                    stack[sp].a = fr->a; stack[sp].b = fr->b; stack[sp].tn = tf; sp++;
                    ca = nr->a; cb = nr->b; continue;
                } else if (hl) { ca = L->a; cb = L->b; continue; }
                else if (hr) { ca = R->a; cb = R->b; continue; }
            } else {
                for (int k = 0; k < cb; ++k) test_prim(s, s->prim_index[ca + k], r, tmin, h, cam);
            }
            int found = 0;
            while (sp > 0) { --sp; if (stack[sp].tn <= h->t) { ca = stack[sp].a; cb = stack[sp].b; found = 1; break; } }
            if (!found) break;
        }
    }
}

/* ---------------------------------------------------------------- camera rays --------- */
static inline ray_t camera_ray(const scene_t* s, int i, int j, float jx, float jy, double inv_w, double inv_h) {
    /* old/raytracer_core copy.cpp:288-289 + old/raytracer_core copy.h:160-184: the direction is
     * evaluated in double like v1 does and rounded to float32 once, so that primary rays carry
     * half-ulp directions (needed for the 1e-5 distance bar on the radius-100 ground sphere). */
    double u = ((double)i + (double)jx) * inv_w;
    double v = ((double)j + (double)jy) * inv_h;
    double ndc_x = (u - 0.5) * 2.0;
    double ndc_y = (0.5 - v) * 2.0;
    double vx = ndc_x * s->sx, vy = ndc_y * s->sy;
    double dx = s->fwd[0] + s->right[0] * vx + s->up[0] * vy;
    double dy = s->fwd[1] + s->right[1] * vx + s->up[1] * vy;
    double dz = s->fwd[2] + s->right[2] * vx + s->up[2] * vy;
    double len = sqrt(dx * dx + dy * dy + dz * dz);
    double il = 1.0 / len;
    v3 d = {(float)(dx * il), (float)(dy * il), (float)(dz * il)};
    v3 o = {s->pos[0], s->pos[1], s->pos[2]};
    return make_ray(o, d);
}

#define T_MIN 0.001f
#define T_MAX 1e10f

/* Primary-hit AOV at pixel centres.  prim = primitive index (upload order) or -1;
 * t = hit distance or 0.  stats[0..2] += rays, node records fetched, primitives tested. */
void orc_trace_primary(const scene_t* s, int W, int H, int mode, int32_t* prim, float* t, uint64_t* stats) {
    double inv_w = 1.0 / (double)W, inv_h = 1.0 / (double)H;
    uint64_t nn = 0, np = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : nn, np)
    for (int j = 0; j < H; ++j) {
        for (int i = 0; i < W; ++i) {
            ray_t r = camera_ray(s, i, j, 0.5f, 0.5f, inv_w, inv_h);
            hit_t h; h.n_node = 0; h.n_prim = 0;
            intersect(s, &r, T_MIN, T_MAX, mode, &h, 1);
            int64_t p = (int64_t)j * W + i;
            prim[p] = h.prim; t[p] = h.prim >= 0 ? h.t : 0.0f;
            nn += h.n_node; np += h.n_prim;
        }
    }
    if (stats) { stats[0] += (uint64_t)W * H; stats[1] += nn; stats[2] += np; }
}

/* Arbitrary rays (origin, direction; direction is normalised here like the Ray ctor). */
void orc_trace_rays(const scene_t* s, const float* org, const float* dir, int64_t n, int mode,
                    int32_t* prim, float* t, uint64_t* stats) {
    uint64_t nn = 0, np = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nn, np)
    for (int64_t k = 0; k < n; ++k) {
        v3 o = {org[3 * k], org[3 * k + 1], org[3 * k + 2]}, d = {dir[3 * k], dir[3 * k + 1], dir[3 * k + 2]};
        ray_t r = make_ray(o, normalize3(d));
        hit_t h; h.n_node = 0; h.n_prim = 0;
        intersect(s, &r, T_MIN, T_MAX, mode, &h, 0);
        prim[k] = h.prim; t[k] = h.prim >= 0 ? h.t : 0.0f;
        nn += h.n_node; np += h.n_prim;
    }
    if (stats) { stats[0] += (uint64_t)n; stats[1] += nn; stats[2] += np; }
}

/* ---------------------------------------------------------------- shading ------------- */
static inline v3 shading_normal(const scene_t* s, int32_t prim, const ray_t* r, v3 point) {
    v3 n;
    if (s->is_tri) {
        const float* p = s->v0e + 9 * (int64_t)prim;
        v3 e1 = {p[3], p[4], p[5]}, e2 = {p[6], p[7], p[8]};
        n = normalize3(cross3(e1, e2));
    } else {
        const float* p = s->cr + 4 * (int64_t)prim;
        v3 c = {p[0], p[1], p[2]};
        n = scale3(sub3(point, c), 1.0f / p[3]);       /* raytracer_core.h:210 */
    }
    if (dot3(r->d, n) < 0.0f) return n;                /* old/raytracer_core copy.h:132-135 */
    return scale3(n, -1.0f);
}

static inline v3 unit_sphere(uint32_t pixel, uint32_t sample, uint32_t bounce, const uint32_t* key) {
    for (uint32_t j = 0;; ++j) {                       /* old/raytracer_core copy.cpp:170-178 */
        uint32_t ctr[4] = {pixel, sample, bounce, 1u + j}, o[4];
        philox4x32_10(ctr, key, o);
        v3 p = {fmaf(2.0f, u01(o[0]), -1.0f), fmaf(2.0f, u01(o[1]), -1.0f), fmaf(2.0f, u01(o[2]), -1.0f)};
        if (dot3(p, p) < 1.0f || j == 255u) return p;
    }
}

/* One camera sample; returns radiance and the number of path segments traced. */
static v3 radiance(const scene_t* s, int i, int j, int W, uint32_t sample, int max_depth, int integrator,
                   int mode, const uint32_t* key, double inv_w, double inv_h, uint64_t* segments,
                   uint64_t* n_node, uint64_t* n_prim) {
    uint32_t pixel = (uint32_t)((int64_t)j * W + i);
    uint32_t ctr0[4] = {pixel, sample, 0u, 0u}, ctl[4];
    philox4x32_10(ctr0, key, ctl);
    ray_t r = camera_ray(s, i, j, u01(ctl[0]), u01(ctl[1]), inv_w, inv_h);
    v3 color = {0, 0, 0}, thr = {1, 1, 1};
    for (int b = 0; b < max_depth; ++b) {
        hit_t h; h.n_node = 0; h.n_prim = 0;
        intersect(s, &r, T_MIN, T_MAX, mode, &h, b == 0);
        (*segments)++; *n_node += h.n_node; *n_prim += h.n_prim;
        if (h.prim < 0) {                              /* miss: background */
            color.x = fmaf(thr.x, s->bg[0], color.x); color.y = fmaf(thr.y, s->bg[1], color.y); color.z = fmaf(thr.z, s->bg[2], color.z);
            break;
        }
        const float* m = s->mats + 8 * (int64_t)s->mat_id[h.prim];
        color.x = fmaf(thr.x, m[5], color.x); color.y = fmaf(thr.y, m[6], color.y); color.z = fmaf(thr.z, m[7], color.z);
        if (b + 1 == max_depth) break;                 /* the next call would return black */
        if (b > 0) { uint32_t c[4] = {pixel, sample, (uint32_t)b, 0u}; philox4x32_10(c, key, ctl); }
        int metal;
        if (integrator == 0) {
            /* v1, old/raytracer_core copy.cpp:219-238: depth counts down from max_depth */
            int remaining = max_depth - b;
            if (!(remaining < 3 || u01(ctl[2]) < 0.8f)) break;
            metal = u01(ctl[3]) < m[3];
        } else {
            /* v2, raytracer_core.cpp:317-329: depth counts up from 1 */
            int depth = b + 1;
            if (depth > 3) {
                float mc = (thr.x > thr.y) ? (thr.x > thr.z ? thr.x : thr.z) : (thr.y > thr.z ? thr.y : thr.z);
                float p = (mc > 0.95f) ? 0.95f : mc;
                if (p < 0.1f) p = 0.1f;
                if (u01(ctl[2]) >= p) break;
                float ip = 1.0f / p;
                thr = scale3(thr, ip);
            }
            metal = m[3] > 0.0f;
        }
        v3 point = {fmaf(r.d.x, h.t, r.o.x), fmaf(r.d.y, h.t, r.o.y), fmaf(r.d.z, h.t, r.o.z)};
        v3 n = shading_normal(s, h.prim, &r, point);
        v3 us = unit_sphere(pixel, sample, (uint32_t)b, key);
        v3 nd;
        if (metal) {
            float k2 = 2.0f * dot3(r.d, n);            /* reflect: v - n*(2 v.n) */
            v3 refl = {fmaf(-k2, n.x, r.d.x), fmaf(-k2, n.y, r.d.y), fmaf(-k2, n.z, r.d.z)};
            nd.x = fmaf(us.x, m[4], refl.x); nd.y = fmaf(us.y, m[4], refl.y); nd.z = fmaf(us.z, m[4], refl.z);
        } else {
            if (!(dot3(us, n) > 0.0f)) us = scale3(us, -1.0f);
            nd.x = n.x + us.x; nd.y = n.y + us.y; nd.z = n.z + us.z;
        }
        thr.x *= m[0]; thr.y *= m[1]; thr.z *= m[2];
        r = make_ray(point, normalize3(nd));
    }
    return color;
}

static inline float resolve1(float sum, float inv_spp) {   /* raytracer_core.cpp:398-409 */
    float c = sqrtf(sum * inv_spp);
    return c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
}

/* Full render of the pixel rectangle [x0,x0+w) x [y0,y0+h) of a W x H frame.
 * out = h*w*3 floats (row-major, RGB).  resolve != 0: mean -> sqrt -> clamp (what
 * RayTracer::render returns); resolve == 0: raw radiance sums.
 * stats[0..3] += camera samples, node records, primitive tests, path segments. */
void orc_render(const scene_t* s, int W, int H, int x0, int y0, int w, int h, int spp, int max_depth,
                uint64_t seed, uint32_t sample_offset, int integrator, int mode, int resolve,
                float* out, uint64_t* stats) {
    double inv_w = 1.0 / (double)W, inv_h = 1.0 / (double)H;
    float inv_spp = 1.0f / (float)spp;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint64_t seg = 0, nn = 0, np = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : seg, nn, np)
    for (int jj = 0; jj < h; ++jj) {
        for (int ii = 0; ii < w; ++ii) {
            int i = x0 + ii, j = y0 + jj;
            v3 sum = {0, 0, 0};
            for (int sidx = 0; sidx < spp; ++sidx) {
                v3 c = radiance(s, i, j, W, sample_offset + (uint32_t)sidx, max_depth, integrator, mode, key,
                                inv_w, inv_h, &seg, &nn, &np);
                sum.x += c.x; sum.y += c.y; sum.z += c.z;
            }
            float* o = out + 3 * ((int64_t)jj * w + ii);
            if (resolve) { o[0] = resolve1(sum.x, inv_spp); o[1] = resolve1(sum.y, inv_spp); o[2] = resolve1(sum.z, inv_spp); }
            else { o[0] = sum.x; o[1] = sum.y; o[2] = sum.z; }
        }
    }
    if (stats) { stats[0] += (uint64_t)w * h * spp; stats[1] += nn; stats[2] += np; stats[3] += seg; }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's reference arm: a launcher (torch.distributed.run) exports OMP_NUM_THREADS=1 to its
 * workers; the CPU arm is to use every host core whatever the launcher says. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
