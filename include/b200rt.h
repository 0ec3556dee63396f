/* b200rt.h -- C ABI of libb200rt.so, the B200-native (sm_100a CUDA) replacement for the render
 * hot path of Samuel-2000/PGR-Raytracing-Project: camera ray generation -> BVH traversal with
 * ray/primitive intersection -> path-traced shading -> sample accumulation / resolve.
 *
 * This is exactly what a binding for the reference's native module `cpp_raytracer.raytracer_cpp`
 * has to bind for that path (INTEGRATION.md shows the pybind11 and ctypes stubs).  Each entry
 * point names the reference interface it replaces; paths are relative to the reference root,
 * "v1" = old/ * (the generation that compiles and runs, and the API interaction.py is written
 * against), "v2" = cpp_raytracer/raytracer_core.{h,cpp}.
 *
 * Conventions: every call returns 0 on success, non-zero on failure (rt_last_error gives the
 * text); nothing throws across the boundary.  Pointers prefixed d_ are DEVICE pointers owned by
 * the caller (e.g. a torch CUDA tensor's data_ptr()); h_ are HOST pointers; scene / BVH device
 * memory is owned by the context.  `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream); calls that take a stream only enqueue work on it.  A context is
 * internally locked: calls on one context from different host threads serialise
 * (gui.py:943,957,981 call set_scene while the render worker may be inside render).
 * There is NO CPU fallback: without a CUDA device rt_create fails.
 */
#ifndef B200RT_H
#define B200RT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RT_ABI_VERSION 1

typedef struct rt_ctx rt_ctx;

/* Flattened BVH node, 32 bytes, read on the device as two 16-byte vector loads.
 * Replaces v2 BVHNodeFlat (cpp_raytracer/raytracer_core.h:221-235, 96 bytes, leaf flag aliased
 * with a child index) and v1's pointer BVHNode (old/bvh copy.h:26-39).
 *   internal node: b == 0, a = index of the left child; the right child is a + 1 (siblings are
 *                  adjacent and 64-byte aligned: a is even);
 *   leaf node:     b = primitive count (1..4), a = first slot into the leaf-ordered primitive
 *                  arrays (prim_index[slot] = primitive number in upload order).
 * Node 0 is the root, node 1 is an unused pad record. */
typedef struct rt_bvh_node {
    float bmin[3];
    int32_t a;
    float bmax[3];
    int32_t b;
} rt_bvh_node;

/* Counters for the roofline; filled only while option "stats" is 1. */
typedef struct rt_stats {
    uint64_t rays;          /* camera samples (render) or rays (trace_*) */
    uint64_t segments;      /* path segments traced (== rays for primary-only work) */
    uint64_t node_records;  /* 32-byte node records fetched */
    uint64_t prim_tests;    /* primitives tested (48 B per triangle, 16 B per sphere) */
    uint64_t launches;      /* kernels launched by this context since creation / reset */
} rt_stats;

/* ---- context: replaces RayTracer::RayTracer / ~RayTracer (old/raytracer_core copy.cpp:148-160) */
int rt_create(int device, rt_ctx** out);
void rt_destroy(rt_ctx* ctx);
const char* rt_last_error(rt_ctx* ctx);          /* ctx may be NULL (errors of rt_create) */
int rt_abi_version(void);

/* ---- scene upload: replaces RayTracer::set_scene (old/raytracer_core copy.cpp:162-167; v2
 * RayTracerWrapper::set_scene raytracer_core.cpp:542-547) which deep-copies the sphere list.
 * center_radius: n x 4 (cx cy cz r); material8: n x 8 (albedo3 metallic roughness emission3,
 * i.e. Material old/raytracer_core copy.h:110-119 without the never-read ior); object_id may
 * be NULL (ids = 0..n-1).  Uploading invalidates the BVH. */
int rt_set_spheres(rt_ctx* ctx, const float* h_center_radius, const float* h_material8,
                   const int32_t* h_object_id, int64_t n);
/* Triangle extension (BASELINE.json configs 2-4; the reference has spheres only).
 * vertices: n x 9 (v0 v1 v2); material_id: n (may be NULL = 0); materials: m x 8. */
int rt_set_triangles(rt_ctx* ctx, const float* h_vertices, const int32_t* h_material_id, int64_t n,
                     const float* h_materials, int m);
/* Scene EDITS without a rebuild.  The reference's host re-sends the whole scene through RayTracer::set_scene -- a deep
 * copy and two BVH builds (old/raytracer_core copy.cpp:84-87,162-167) -- on every drag / slider event
 * (interaction.py:906,1169, gui.py:943).  rt_update_geometry replaces the positions of the n primitives of the current
 * scene (same kind, same count, same order; n x 4 spheres or n x 9 triangles) and REFITS the current tree on the device
 * (same topology, every box recomputed bottom-up; ~0.1 ms for 1M triangles + the upload) instead of rebuilding it.
 * Pixels are those of a tree rebuilt from scratch, bit for bit; after large moves a refitted tree gets slower, not
 * wrong, so every refit measures the summed surface area of the internal boxes against the tree as built and, above
 * option "refit_limit" percent (default 200; 0 = never), drops the tree: the next launch builds a new one with option
 * "builder".  Read-only options "refits", "refit_rebuilds", "refit_area_pct" report what happened.
 * rt_update_materials replaces the m material rows (no BVH work at all). */
int rt_update_geometry(rt_ctx* ctx, const float* h_prims, int64_t n);
int rt_update_materials(rt_ctx* ctx, const float* h_material8, int m);
/* Scene::background_color (old/raytracer_core copy.h:226; interaction.py:297). */
int rt_set_background(rt_ctx* ctx, const float rgb[3]);

/* ---- BVH: replaces Scene::build_bvh / BVH::build (old/raytracer_core copy.cpp:104-110,
 * old/bvh copy.cpp:111-199) and v2 SceneIntersector::build_bvh + BVHBuilder::build
 * (raytracer_core.cpp:57-141,165-189).
 * builder 0 = the reference's top-down median split (leaf <= 4, longest axis, median by box
 *             centre), built on the host, deterministic;
 * builder 1 = LBVH built on the device (Morton order), for interactive edits.
 * builder 2 = binned surface-area heuristic on the host (16 centroid bins per axis; leaves of <= "leaf_size" primitives, cut
 * further while that is cheaper): a better tree than the reference's rule gives, same pixels (see rt_build_bvh_host_ex).
 * rt_get_bvh / rt_set_bvh let a checker walk the very same tree (nodes may be NULL to query
 * the count).  prim_index has n entries. */
int rt_build_bvh(rt_ctx* ctx, int builder);
int rt_get_bvh(rt_ctx* ctx, rt_bvh_node* h_nodes, int64_t* n_nodes, int32_t* h_prim_index);
int rt_set_bvh(rt_ctx* ctx, const rt_bvh_node* h_nodes, int64_t n_nodes, const int32_t* h_prim_index);
/* Context-free host build with builder 0 (no device needed): for offline BVH caching and for
 * checking the builder on machines without a GPU.  h_prims = n x 4 spheres or n x 9 triangles.
 * Returns the node count through *n_nodes (call with h_nodes == NULL to size the arrays:
 * at most 2*n + 2 records). */
int rt_build_bvh_host(const float* h_prims, int is_triangles, int64_t n, rt_bvh_node* h_nodes,
                      int64_t* n_nodes, int32_t* h_prim_index);
/* The same with a choice of host builder and leaf size: builder 0 = the reference's median split (leaf_size 4 = its rule),
 * 2 = binned surface-area heuristic -- NOT the reference's tree (cpp_raytracer/raytracer_core.cpp:57-118 knows one rule only):
 * the same layout and, because closest hits do not depend on the tree, the same hits; fewer nodes entered and primitives
 * tested per ray.  A tree deeper than the traversal stack (degenerate input) falls back to builder 0.  At most 2*n + 2 records. */
int rt_build_bvh_host_ex(const float* h_prims, int is_triangles, int64_t n, int builder, int leaf_size,
                         rt_bvh_node* h_nodes, int64_t* n_nodes, int32_t* h_prim_index);

/* ---- camera: replaces RayTracer::set_camera (old/raytracer_core copy.h:266) + the basis that
 * Camera::get_ray recomputes per ray (old/raytracer_core copy.h:160-184).  aspect <= 0 means
 * "use W/H of each render call" (what RayTracer::render does, old/raytracer_core copy.cpp:259).
 * rt_get_camera_block returns pos3 fwd3 right3 up3 sx sy as doubles. */
int rt_set_camera(rt_ctx* ctx, const double pos[3], const double target[3], const double up[3],
                  double fov_deg, double aspect);
int rt_get_camera_block(rt_ctx* ctx, int width, int height, double out[14]);

/* ---- primary-hit AOV: the parity hook.  Replaces the per-pixel use of Scene::hit
 * (old/raytracer_core copy.cpp:112-131) / SceneIntersector::intersect (raytracer_core.cpp:191-273)
 * for camera rays through pixel centres u=(i+.5)/W, v=(j+.5)/H, t in [0.001, 1e10].
 * d_prim: W*H int32 primitive number in upload order (-1 = miss); d_t: W*H float (0 on miss). */
int rt_trace_primary(rt_ctx* ctx, int width, int height, int32_t* d_prim, float* d_t, void* stream);
/* Arbitrary rays (origin, direction; normalised like the Ray ctor raytracer_core.h:113). */
int rt_trace_rays(rt_ctx* ctx, const float* d_origin, const float* d_direction, int64_t n,
                  int32_t* d_prim, float* d_t, void* stream);
/* RayTracer::select_object (old/raytracer_core copy.cpp:245-248): closest object id under the
 * normalised screen position (x,y), t in [0.001, 1000], or -1.  Synchronous. */
int rt_select_object(rt_ctx* ctx, double x, double y, int width, int height, int32_t* out_object_id);

/* ---- render: replaces RayTracer::render (old/raytracer_core copy.cpp:257-318) /
 * PathTracer::render (raytracer_core.cpp:354-416): spp jittered camera samples per pixel,
 * path tracing to max_depth, mean -> sqrt gamma -> clamp [0,1], RGB float32, row-major, top
 * row first.  d_out: H*W*3 floats.  Samples use Philox4x32-10 keyed by `seed`, counter
 * (pixel, sample_offset + s, bounce, draw): successive progressive batches pass successive
 * sample_offset values (v1 keeps thread-local generator state across calls instead). */
int rt_render(rt_ctx* ctx, int width, int height, int spp, int max_depth, uint64_t seed,
              uint32_t sample_offset, float* d_out, void* stream);
/* Same, for the interleaved tiles {first_tile + k*tile_stride} of the frame cut into
 * tile_w x tile_h tiles (row-major tile numbering) -- the multi-GPU partition.  d_out is the
 * compact buffer [k][tile_h][tile_w][3]; resolve = 0 writes raw radiance sums instead. */
int rt_render_tiles(rt_ctx* ctx, int width, int height, int tile_w, int tile_h, int first_tile,
                    int tile_stride, int spp, int max_depth, uint64_t seed, uint32_t sample_offset,
                    int resolve, float* d_out, void* stream);
/* The same partition written in FRAME layout: rank's tiles of the W x H frame go to their final place in
 * d_frame (H*W*3 floats), which may be another GPU's frame mapped with rt_frame_open -- the render
 * kernel's own stores are then the frame exchange (NVLink peer writes), and the G-GPU frame is bit-identical
 * to the 1-GPU one.  Tiles are dealt out skewed (logical tile L = rank, rank + world, ... sits in tile row
 * ty = L / tiles_x, column (L % tiles_x + ty) % tiles_x) so that every rank gets a share of every tile
 * row and every tile column. */
int rt_render_tiles_frame(rt_ctx* ctx, int width, int height, int tile_w, int tile_h, int rank, int world,
                          int spp, int max_depth, uint64_t seed, uint32_t sample_offset, int resolve,
                          float* d_frame, void* stream);
/* Frames shared between the processes of one node (CUDA IPC).  rt_frame_alloc: library-owned device buffer
 * of `planes` H*W*3 float frames plus its 64-byte handle (send it to the other ranks by any means);
 * rt_frame_open maps such a buffer into this process (peer access over NVLink); close / free undo them. */
int rt_frame_alloc(rt_ctx* ctx, int width, int height, int planes, float** d_frame, unsigned char handle[64]);
int rt_frame_free(rt_ctx* ctx, float* d_frame);
int rt_frame_open(rt_ctx* ctx, const unsigned char handle[64], float** d_peer_frame);
int rt_frame_close(rt_ctx* ctx, float* d_peer_frame);
/* Barrier between the `world` processes that render into one shared frame, THROUGH that frame's memory (rt_frame_alloc
 * reserves the sync words behind the planes): enqueued on `stream` after the render, it publishes this process's
 * stores, counts the process in with one atomic on the owner's memory and returns (in stream order) once all `world`
 * processes of frame number `epoch` (1, 2, 3, ... per shared frame, the same on every process) have arrived -- a few
 * microseconds over NVLink instead of a collective launch.  A process that has not arrived after about 20 s is taken
 * for lost: the waiting kernels trap, so every waiting process fails at its next CUDA call rather than reading an
 * incomplete frame.  width / height / planes as given to rt_frame_alloc. */
int rt_frame_sync(rt_ctx* ctx, float* d_frame, int width, int height, int planes, int world, uint64_t epoch, void* stream);
/* HOST frames shared between the processes of one node -- the multi-GPU form of rt_render_host.  The frame lives in
 * host memory that every process maps (POSIX shared memory, a file under /dev/shm ...); each process page-locks it
 * and gets a device alias (rt_host_register: cudaHostRegister portable + mapped), and rt_render_tiles_host renders
 * this rank's (skew-dealt, see rt_render_tiles_frame) 32x32 tiles and lets the GPU store them straight into the
 * shared host frame over ITS OWN PCIe link: no gather on a display GPU, no serial device->host copy; N GPUs move the
 * frame N times as fast.  When the last tile of the rank has been stored the GPU writes `epoch` into the rank's flag
 * word (host memory as well, d_flag = its device alias); the consumer waits for all ranks' words with rt_host_wait
 * (a spin on host memory, no CUDA call).  The call only enqueues work on `stream`. */
int rt_host_register(rt_ctx* ctx, void* h_ptr, uint64_t bytes, void** d_alias);
int rt_host_unregister(rt_ctx* ctx, void* h_ptr);
int rt_render_tiles_host(rt_ctx* ctx, int width, int height, int rank, int world, int spp, int max_depth, uint64_t seed,
                         uint32_t sample_offset, float* d_host_frame, uint32_t* d_flag, uint32_t epoch, void* stream);
/* 0 when all n flag words equal `epoch`; 2 after timeout_s seconds. */
int rt_host_wait(const volatile uint32_t* h_flags, int n, uint32_t epoch, double timeout_s);
/* Full frame, raw radiance sums (no mean / gamma / clamp): the per-rank partial of a sample-range
 * partition; sum the partials (e.g. ncclReduce) and finish with rt_resolve. */
int rt_render_sum(rt_ctx* ctx, int width, int height, int spp, int max_depth, uint64_t seed,
                  uint32_t sample_offset, float* d_out, void* stream);
/* mean over spp_total -> sqrt gamma -> clamp [0,1] (cpp_raytracer/raytracer_core.cpp:398-409). */
int rt_resolve(rt_ctx* ctx, const float* d_sum, float* d_out, int64_t n_floats, int spp_total, void* stream);
/* Sample-range partition with the exchange done by the render kernels themselves: rank g renders
 * rt_render_sum straight into plane g of the display rank's shared buffer (rt_frame_alloc / rt_frame_open);
 * after a barrier the display rank sums the planes in plane order (deterministic) and resolves. */
int rt_resolve_planes(rt_ctx* ctx, const float* d_planes, int n_planes, int64_t plane_stride_floats, float* d_out,
                      int64_t n_floats, int spp_total, void* stream);
/* Scatter gathered compact tile buffers [rank][k][tile_h][tile_w][3] back into a H*W*3 frame. */
int rt_untile(rt_ctx* ctx, int width, int height, int tile_w, int tile_h, int n_ranks,
              const float* d_tiles, float* d_frame, void* stream);
/* Host-buffer call behind the reference-facing RayTracer::render (old/raytracer_core copy.cpp:257, result in
 * host memory): render into a context-owned device framebuffer and bring it to h_out (H*W*3 floats).  With
 * page-locked h_out (cudaHostAlloc / cudaHostRegister / torch pin_memory) a camera-ray frame crosses PCIe WHILE it
 * is rendered (option "overlap"); pageable memory works, slower.  Synchronous: returns when h_out is complete. */
int rt_render_host(rt_ctx* ctx, int width, int height, int spp, int max_depth, uint64_t seed,
                   uint32_t sample_offset, float* h_out);

/* ---- progressive accumulation: replaces the numpy running mean of gamma'd batches in
 * interaction.py:1311-1325: accum = accum * n_old/(n_old+n_batch) + batch * n_batch/(n_old+n_batch)
 * (n_old == 0: accum = batch). */
int rt_accumulate(rt_ctx* ctx, const float* d_batch, float* d_accum, int64_t n_floats, int n_old,
                  int n_batch, void* stream);
/* Display chain of interaction.py:1435-1439 + gui.py:73: x*e/(1+x*e) -> clip -> *255 -> uint8. */
int rt_tonemap_u8(rt_ctx* ctx, const float* d_accum, uint8_t* d_rgb8, int64_t n_floats, float exposure,
                  void* stream);

/* The whole display chain of the host on the device: tone map as above, then (enhance != 0) the contrast
 * stretch of interaction.py:1441-1449 -- (x - p2) / (p98 - p2) clipped, p2 / p98 = numpy's percentile(2) /
 * percentile(98) over all n_floats tone-mapped values (exact order statistics, numpy >= 2 float32 rules) --
 * then * 255 -> uint8.  enhance == 0 is rt_tonemap_u8. */
int rt_display_u8(rt_ctx* ctx, const float* d_accum, uint8_t* d_rgb8, int64_t n_floats, float exposure, int enhance,
                  void* stream);

/* ---- options and counters.  Options: "integrator" 0 = v1 semantics (default; the generation
 * that runs: RR `depth<3 || rand<0.8` unweighted, metal chosen with probability metallic),
 * 1 = v2 semantics (raytracer_core.cpp:317-347); "stats" 0/1; "kernel" -1 = auto (default: picks by
 * scene size and max_depth), 0 = persistent path kernel with lane-level continuation, 1 = simple
 * one-pixel-per-thread megakernel, 2 = wavefront (generate / trace / shade / accumulate kernels over
 * compacted ray queues), 3 = camera-ray packets (max_depth 1 and the primary-hit query: a warp walks
 * the BVH as one 8x4-pixel packet; falls back to 0 for max_depth > 1), 4 = wavefront whose bounce 0 (the
 * coherent camera rays) is generated and traced by packets (falls back to 3 for max_depth 1) -- all give
 * bit-identical results; "refill" 1..32 = share (in 32nds)
 * of a warp's traversing lanes below which it leaves the traversal loop to shade / refill (default
 * 8); "leaf_vote" 1..32 = lanes holding a leaf at which the warp runs the leaf step (default 8);
 * "builder" 0/1/2 = what the implicit build of the first render after a scene upload uses (rt_build_bvh's
 * argument; default 0); "schedule" 0/1 = cost-aware work order of the packet kernel (default 1);
 * "overlap" = how rt_render_host moves a camera-ray frame (max_depth 1, 1 spp) to the host: 2 (default) = TILE
 * PUSH, the render kernel itself stores every finished 32x32 tile into h_out (needs page-locked, 16-byte aligned
 * h_out; otherwise mode 1 is used), 1 = finished regions are copied by the DMA engine while the kernel renders the
 * rest, 0 = render, then one copy; other frames of >= 4 MB into page-locked memory go out as bands of tile rows, each copied while
 * the next renders.  Read-only "host_path" says which way the last rt_render_host took (0 copy, 1 regions, 2 tile push, 3 bands).
 * "kernel" 5 = tiny scenes (<= 64 primitives, max_depth <= 8): the whole scene staged in shared memory (rt_tiny.cu), "tiny_mode"
 * 0 / 1 = its CTA-local wavefront / lock-step form, "tiny_threads" 128 / 256; "treelet" 0..10 = levels of the tree staged in shared
 * memory by the packet / incoherent-ray kernels (0 = off, the measured optimum); "sah_cost" = builder 2's cost of a traversal step in tenths of a primitive test (default 30, the measured optimum on the 1M-triangle benchmark scene); "leaf_size" 1..4 = primitives per leaf of builders 0 and 2
 * (4 = the reference's rule); "wf_streams" 1 / 2 = waves of a wavefront frame in flight (2: kernel tails overlap); "fold",
 * "wf_rays_per_lane" = measured-and-dropped variants kept for A/B runs.
 * "qnodes" = how the incoherent bounces of the wavefront (bounces >= 1, scenes of > 64 primitives) walk the tree; a sum of
 * 1 = COMPRESSED sibling pairs: a pair in 32 bytes (one 256-bit load instead of two), planes on a 15-bit grid over the root box, rounded
 * outward, so the closest hits -- and every pixel -- stay those of the full records; 2 = triangle records read as a 32-byte + a 16-byte
 * part (measured: no gain, kept for A/B runs); 4 = COOPERATIVE leaf step (triangles): the triangles of up to eight leaf-holding lanes are
 * tested by all 32 lanes of the warp, closest hit folded by a shared-memory atomic min on (distance, primitive) -- the same hit bit for
 * bit.  -1 (default) = 4, plus 1 while the grid is fine enough for the scene (mean over the leaves of half-area on the grid / half-area
 * as stored <= "qnodes_area_limit" %, default 115; read-only "qnodes_area_pct", "qnodes_used"); 0 = off.  A tree with a box outside the root box
 * (possible only with rt_set_bvh) never uses the compressed pairs ("qnodes_area_pct" reads -1).  Instrumented launches
 * ("stats" 1) always walk the full records, so their counters stay the per-ray walk's. */
int rt_set_option(rt_ctx* ctx, const char* name, int64_t value);
int rt_get_option(rt_ctx* ctx, const char* name, int64_t* value);
int rt_get_stats(rt_ctx* ctx, rt_stats* out);   /* synchronises the device */
int rt_reset_stats(rt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
