// rt_kernel_common.cuh -- helpers shared by the tracing kernels (rt_kernels.cu, rt_wavefront.cu).
#pragma once
#include "rt_kernels.h"

namespace b200rt {

constexpr int kThreads = 128;

struct PixelWork { int i, j, out_index; bool active; };

// Decode work item `w` (one 8x4 pixel block) for this lane.
__device__ __forceinline__ PixelWork decode_work(const TileMap& tm, int w, int lane) {
    int k, sy, sx;                        // local tile number, block row / column inside the tile
    if (tm.tile_w == 32 && tm.tile_h == 32) {             // the library's own tiles: shifts instead of two integer divisions
        k = w >> 5; sy = (w >> 2) & 7; sx = w & 3;
    } else {
        int bx = tm.tile_w >> 3;
        int per_tile = bx * (tm.tile_h >> 2);
        k = w / per_tile;
        int sub = w - k * per_tile;
        sy = sub / bx; sx = sub - sy * bx;
    }
    int tile = tm.first_tile + k * tm.tile_stride;
    int ty = tile / tm.tiles_x, tx = tile - ty * tm.tiles_x;
    if (tm.skew) tx = (tx + tm.skew * ty) % tm.tiles_x;
    int lx = (sx << 3) + (lane & 7), ly = (sy << 2) + (lane >> 3);
    PixelWork p;
    p.i = tx * tm.tile_w + lx;
    p.j = ty * tm.tile_h + ly;
    p.active = p.i < tm.width && p.j < tm.height;
    p.out_index = tm.compact ? (k * tm.tile_h + ly) * tm.tile_w + lx : p.j * tm.width + p.i;
    return p;
}

// One RGB value per lane for an 8x4 pixel block (lane = row * 8 + column).  The three component stores of
// 12-byte-strided pixels become, per row, three stores of 8 CONSECUTIVE floats (full 32-byte sectors; 9 shuffles):
// it matters when the frame lives in another GPU's memory (peer stores over NVLink), where partial sectors are
// expensive.  Whole warp must call; rows with an inactive pixel (frame edge) fall back to per-pixel stores.
__device__ __forceinline__ void warp_store_rgb(float* o, bool active, float r, float g, float b, int lane) {
    if (__all_sync(0xffffffffu, active)) {
        float* row = o - 3 * (lane & 7);
        const int q = lane & 7, base = lane & ~7;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int f = q + 8 * k, px = f / 3, c = f - 3 * px;
            const float vr = __shfl_sync(0xffffffffu, r, base + px), vg = __shfl_sync(0xffffffffu, g, base + px),
                        vb = __shfl_sync(0xffffffffu, b, base + px);
            row[f] = c == 0 ? vr : (c == 1 ? vg : vb);
        }
    } else if (active) { o[0] = r; o[1] = g; o[2] = b; }
}

__device__ __forceinline__ int next_work(unsigned int* counter, int lane) {
    unsigned int w = 0;
    if (lane == 0) w = atomicAdd(counter, 1u);
    return (int)__shfl_sync(0xffffffffu, w, 0);
}

__device__ __forceinline__ void flush_stats(unsigned long long* d_stats, unsigned long long rays,
                                            const Counters& c) {
    unsigned long long v[4] = {rays, c.segments, c.nodes, c.prims};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned long long x = v[q];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(d_stats + q, x);
    }
}


// ---- chunked work distribution of the packet kernels (k_packet, k_wf_packet0): see k_packet.
constexpr int kPacketThreads = 256;
constexpr int kChunk = 8;
constexpr unsigned kNoChunk = 0xFFFFFFu;

static __device__ __forceinline__ unsigned take_chunk(unsigned int* global_counter, int n_chunks, const int* __restrict__ order) {
    const unsigned c = atomicAdd(global_counter, 1u);
    if (c >= (unsigned)n_chunks) return kNoChunk << 8;
    return (unsigned)(order ? order[c] : (int)c) << 8;
}

static __device__ __forceinline__ int chunk_next_block(unsigned* word, unsigned int* global_counter, int n_chunks,
                                                const int* __restrict__ order, int lane) {
    unsigned v = 0xFFFFFFFFu;
    if (lane == 0) {
        for (;;) {
            const unsigned old = atomicAdd(word, 1u);
            const unsigned chunk = old >> 8, off = old & 0xFFu;
            if (chunk == kNoChunk) break;
            if (off < (unsigned)kChunk) { v = chunk * kChunk + off; break; }
            if (off == (unsigned)kChunk) {                       // this warp installs the next chunk
                atomicExch(word, take_chunk(global_counter, n_chunks, order));
                continue;
            }
            while ((*(volatile unsigned*)word >> 8) == chunk) __nanosleep(32);
        }
    }
    return (int)__shfl_sync(0xffffffffu, v, 0);
}

template <typename K>
inline int resident_grid(K kernel, int sm_count) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
    return sm_count * per_sm;
}

inline int work_items(const TileMap& tm) { return tm.n_local_tiles * (tm.tile_w >> 3) * (tm.tile_h >> 2); }

inline int elementwise_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}


}  // namespace b200rt
