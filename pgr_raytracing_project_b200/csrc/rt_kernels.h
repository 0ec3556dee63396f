// rt_kernels.h -- host-callable launchers of the sm_100a kernels (rt_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "rt_device.cuh"

namespace b200rt {

// How a launch enumerates pixels: the frame is cut into tile_w x tile_h tiles (row-major tile
// numbering); this launch covers tiles first_tile + k * tile_stride, k in [0, n_local_tiles).
// Inside a tile every warp takes one 8x4 pixel block at a time from a global work counter.
struct TileMap {
    int width, height;
    int tile_w, tile_h;         // multiples of 8 and 4
    int tiles_x, n_tiles;
    int first_tile, tile_stride, n_local_tiles;
    int compact;                // 0: write frame layout; 1: write [k][tile_h][tile_w][...] layout
    int skew;                   // logical tile L = row ty, column (L % tiles_x + skew * ty) % tiles_x: with skew != 0 a
                                // rank's tiles L = rank, rank + world, ... hit every tile row AND every tile column
};

// Region-completion signalling of k_packet for rt_render_host's copy/compute overlap: the frame's
// 32x32 tiles are grouped into rectangular regions (band_rows tile rows x group_cols tile columns,
// numbered row-major); a warp that finishes a block fences its pixels (gpu scope) and bumps the
// region's counter, and the warp that completes a region fences system-wide (fences are cumulative)
// and raises flags[region] in mapped pinned host memory, which the calling host thread polls to start
// that region's device->host copy (a 2-D copy) while the kernel is still rendering the rest of the
// frame.  Regions rather than full-width bands, because a frame has a few very slow packets (see
// ChunkSchedule) and a band that contains one cannot complete before it does; small regions confine
// the wait to a small share of the bytes.
//
// TILE PUSH (host_fb != nullptr; regions are single 32x32 tiles): no flags, no host polling, no DMA -- the warp that
// completes a tile reads its 32 rows back from L2 and stores them straight into the caller's page-locked frame
// (host_fb = that buffer's device alias), 16 bytes per lane, 384 contiguous bytes per row.  SM-issued coalesced
// stores into mapped host memory run at 51.5 GB/s on this box (DMA: 56.5; tools/exp/zc_copy.cu), so the frame
// crosses PCIe WHILE it is rendered and the call ends one tile (12 KB) after the kernel's last block.
struct BandSignal {
    unsigned int* cnt;               // device, one counter per region, zeroed on the launch stream
    volatile unsigned int* flags;    // mapped pinned host memory (device pointer), zeroed by the host
    float* host_fb;                  // tile push: device alias of the caller's page-locked H x W x 3 frame, else nullptr
    unsigned long long* push_times;  // debug (B200RT_PUSH_TIMES): globaltimer when a tile's push has been issued, or nullptr
    int tiles_x, tiles_y;            // tile grid of the frame (full-frame tile map only)
    int band_rows, group_cols;       // region size in tiles
    int n_groups;                    // regions per band row
};

__host__ __device__ inline int region_of_tile(const BandSignal& b, int tile) {
    const int ty = tile / b.tiles_x, tx = tile - ty * b.tiles_x;
    return (ty / b.band_rows) * b.n_groups + tx / b.group_cols;
}
__host__ __device__ inline int region_blocks(const BandSignal& b, int region) {   // 32 8x4 blocks per tile
    const int by = region / b.n_groups, gx = region - by * b.n_groups;
    const int rows = min(b.band_rows, b.tiles_y - by * b.band_rows), cols = min(b.group_cols, b.tiles_x - gx * b.group_cols);
    return rows * cols * 32;
}

// Cost-aware chunk order of k_packet.  Packets differ in cost by more than 10x (rays that run along a
// BVH split plane), and a slow packet that starts late is the kernel's tail.  Every frame records the
// step counts of each chunk; the next frame of the same tile map starts chunks in latest-start-time
// order (k_chunk_order in rt_kernels.cu).  Pure scheduling: pixels do not depend on it.
// order == nullptr: raster order, nothing recorded.
struct ChunkSchedule {
    int* order;                      // n_chunks entries, written by k_chunk_order, read by k_packet
    unsigned int* cost_sum;          // n_chunks: sum of the chunk's block costs (last frame in, this frame out)
    unsigned int* cost_max;          // n_chunks: largest block cost of the chunk
    int reorder_frames;              // > 0: rebuild `order` now from the costs of that many frames; 0: keep the order
};

struct LaunchCfg {
    cudaStream_t stream;
    int sm_count;
    unsigned int* d_work_counter;            // zeroed by the launcher on `stream`
    unsigned long long* d_stats;             // [rays, segments, node_records, prim_tests] or nullptr
    int variant;                             // 0 = k_path (lane continuation), 1 = simple per-pixel megakernel, 2 = wavefront,
                                             // 3 = k_packet (camera rays: warp = packet with one shared stack),
                                             // 4 = wavefront with bounce 0 by packets (k_wf_packet0),
                                             // 5 = k_tiny (whole scene in shared memory, CTA-local wavefront; rt_tiny.cu)
    float4* d_cam_prims;                     // camera-relative triangle records (3 x float4 per slot)
    int cam_table_valid;                     // the table already holds this scene + camera position: skip k_cam_tris
    BandSignal band;                         // cnt == nullptr: no signalling
    ChunkSchedule sched;                     // order == nullptr: raster order, no cost recording
    float4* d_planes;                        // item mode of k_packet (multi-sample frames): sample planes, plane_batch x tasks
    int plane_batch;                         // samples per batch the planes buffer holds for this tile map (0: loop mode)
    unsigned int* d_fold_cnt;                // item mode, one batch: per-block sample counters (zeroed by the launcher); nullptr: k_plane_accumulate pass
    unsigned long long* d_block_times;       // debug (instrumented k_packet only): [2 * work item] = globaltimer start, end; or nullptr
    int wf_rays_per_lane;                    // k_wf_trace: CTAs beyond queue / (128 x this) leave at once (0: the whole grid works)
    int tiny_threads;                        // k_tiny: threads per CTA (256 or 128)
    int tiny_mode;                           // 0 = CTA-local wavefront with compaction (k_tiny), 1 = lock step per warp (k_tiny_lockstep)
    int refill_below;                        // k_path: leave the traversal loop below this many of 32 lanes
    int leaf_vote;                           // phase voting: leaf step when >= this many lanes hold a leaf
    int qmode;                               // option "qnodes": k_wf_trace of bounces >= 1 reads compressed pairs (bit 0) / split triangle records (bit 1),
                                             // runs the cooperative leaf step (bit 2)
    int coop_leaf_vote, coop_refill;         // leaf_vote / refill_below of the cooperative-leaf launches
};

// Wavefront state in HBM (allocated by the context, float4 SoA).
struct WaveBuffers {
    float4* ray_o[2];        // origin | path slot (int bits); double-buffered by bounce parity
    float4* ray_d[2];        // unit direction | 0
    float4* hit;             // t | prim | slot | 0
    float4* path_thr;        // throughput
    float4* path_rad;        // radiance accumulated along the path
    unsigned int* counters;  // [0..max_depth] queue sizes, then [0..max_depth] fetch cursors
    int capacity;            // paths per wave
};

// Two waves in flight (LaunchCfg::wave2 != nullptr): the frame's waves alternate between two buffer sets on two internal streams,
// so that one wave's kernel tails (the last, longest rays of a persistent trace kernel) overlap the other wave's kernels.
struct WavePipe {
    const WaveBuffers* wave2;        // second buffer set, or nullptr: one wave at a time on cfg.stream
    cudaStream_t streams[2];
    cudaEvent_t fork, join[2], acc[2];
    unsigned int* counters[2];       // chunk counters of k_wf_packet0, one per set
};
cudaError_t launch_wavefront(const SceneView& sc, bool is_tri, bool aov, const CameraBlock& cam, const TileMap& tm,
                             int spp, int max_depth, int integrator, uint64_t seed, uint32_t sample_offset, int resolve,
                             float* d_out, int32_t* d_prim, float* d_t, const LaunchCfg& cfg, const WaveBuffers& wb,
                             int* n_launches, const WavePipe* pipe = nullptr);
// tiny scenes (rt_tiny.cu): the whole scene staged in shared memory, brute-force closest hit, CTA-local wavefront
bool tiny_eligible(const SceneView& sc, int n_mats, int max_depth);
cudaError_t launch_tiny(const SceneView& sc, bool is_tri, int n_mats, const CameraBlock& cam, const TileMap& tm, int spp, int max_depth,
                        int integrator, uint64_t seed, uint32_t sample_offset, int resolve, float* d_out, const LaunchCfg& cfg);
cudaError_t launch_cam_tris(const SceneView& sc, const CameraBlock& cam, const LaunchCfg& cfg);   // per-frame camera-relative triangle table
int packet_chunks(const TileMap& tm, int items_per_block);   // chunks k_packet cuts this tile map into (ChunkSchedule sizes)
cudaError_t launch_trace_primary(const SceneView& sc, bool is_tri, const CameraBlock& cam, const TileMap& tm,
                                 int32_t* d_prim, float* d_t, const LaunchCfg& cfg);
cudaError_t launch_trace_rays(const SceneView& sc, bool is_tri, const float* d_org, const float* d_dir, int64_t n,
                              int32_t* d_prim, float* d_t, const LaunchCfg& cfg);
cudaError_t launch_render(const SceneView& sc, bool is_tri, const CameraBlock& cam, const TileMap& tm, int spp,
                          int max_depth, int integrator, uint64_t seed, uint32_t sample_offset, int resolve,
                          float* d_out, const LaunchCfg& cfg);
cudaError_t launch_untile(int width, int height, int tile_w, int tile_h, int n_ranks, const float* d_tiles,
                          float* d_frame, cudaStream_t stream);
cudaError_t launch_resolve(const float* d_sum, float* d_out, int64_t n, int spp_total, cudaStream_t stream);
cudaError_t launch_resolve_planes(const float* d_planes, int n_planes, int64_t plane_stride, float* d_out, int64_t n,
                                  int spp_total, cudaStream_t stream);
cudaError_t launch_accumulate(const float* d_batch, float* d_accum, int64_t n, int n_old, int n_batch,
                              cudaStream_t stream);
// Barrier between the processes that share a frame (rt_frame_sync): words[0] = arrival counter, words[1] = error flag.
cudaError_t launch_frame_sync(unsigned long long* words, unsigned long long target, cudaStream_t stream);
// Top `n_pairs` sibling pairs of the tree (heap order) with child codes for shared-memory staging (rt_kernels.cu k_build_treelet).
cudaError_t launch_build_treelet(const float4* d_nodes, int n_pairs, float4* d_treelet, cudaStream_t stream);
// Option "qnodes": the tree's sibling pairs compressed to 32 bytes on a 15-bit grid over the root box (d_qgrid: 6 floats), and the
// triangle records split into a 32-byte and a 16-byte part (rt_device.cuh pair_hit_q, trav_run QM).
// d_quality[0], [1]: number of leaves, sum of their (half-area on the grid / half-area as stored) (the host's keep-or-drop figure);
// [2] != 0: a plane did not fit the grid (a child box outside the root box: the compressed copy must not be used).
cudaError_t launch_quantize_pairs(const float4* d_nodes, int n_pairs, uint4* d_qnodes, float* d_qgrid, double* d_quality, cudaStream_t stream);
cudaError_t launch_split_tris(const float4* d_prims, int n, float4* d_tri_a, float4* d_tri_b, cudaStream_t stream);
// rt_render_tiles_host: this launch's 32x32 tiles of the device frame `fb` stored into the page-locked frame `host` (device
// alias); the last warp to finish writes `epoch` to *flag (system scope).  cnt: one device word, zeroed here.
cudaError_t launch_push_tiles(const TileMap& tm, const float* d_fb, float* d_host, unsigned int* d_flag, unsigned int epoch,
                              unsigned int* d_cnt, cudaStream_t stream);
cudaError_t launch_tonemap_u8(const float* d_accum, uint8_t* d_rgb8, int64_t n, float exposure, cudaStream_t stream);

}  // namespace b200rt
