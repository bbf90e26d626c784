// render.hpp — launch interface between the C ABI (capi.cu) and the CUDA kernels (render.cu, wavefront.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"
#include "mort_scene_format.h"

namespace mort {

struct FrameParams {
    DeviceScene sc;
    CameraParams cam;
    uint32_t seed, frame;
    int32_t sj_mod, sj_rem;         // sample-split: strata rows s_j % sj_mod == sj_rem
    int32_t n_rows;                 // strata rows this call renders
    int32_t n_subset;               // samples per pixel this call renders = n_rows * sqrt_spp
    int32_t lanes_per_pixel;        // megakernel: pixels per warp task (1..16)
    int32_t min_task_px;            // megakernel: smallest task the guided tail hands out (== lanes_per_pixel: fixed-size tasks)
    int32_t n_pixels;               // pixels THIS call renders (all of them, or the rank's 8-row bands)
    int32_t tile_mod, tile_rem;     // tile split: local pixel index -> global pixel through bands b = k * tile_mod + tile_rem
    int32_t band_px;                // pixels per band = band rows (8 unless opts.tile_rows says otherwise) x width
    int32_t n_staged;               // BVH nodes copied to shared memory per block
    int32_t accumulate;             // exact frames only: add this call's sums to what accum_exact already holds
    float4* accum;                  // W*H (null when accum_exact is used)
    unsigned long long* accum_exact;// W*H x 4 (r, g, b fixed point, flags) or null
    unsigned long long* counters;   // [0] segments, [1] samples, [2] scheduler watchdog trips (an error)
    unsigned int* work_counter;     // persistent-warp work queue head
    // block wavefront (pool.cu)
    int32_t pool_paths;             // paths per block pool
    int32_t pool_overlap;           // tree scenes: classify starts on traced paths while other warps still trace (no barrier between the two)
    int32_t pool_refill;            // trace phase: lanes refill mid-traversal when at least this many of a warp's lanes are idle (0 = never)
    unsigned long long* work64;     // head of the call's sample index space [0, total_samples)
    unsigned long long total_samples;   // n_frames * n_pixels * n_subset
    unsigned long long frame_samples;   // n_pixels * n_subset
    unsigned long long frame_words;     // exact-frame words between consecutive frames of a batch (W * H * 4)
    uint32_t pix_mask; int32_t frame_shift;   // batch of frames in one launch: a path's pixel word = pixel | frame-in-batch << frame_shift (0: single frame)
};

struct LaunchShape { int threads, blocks, smem_bytes; };

// megakernel (render.cu)
cudaError_t mega_query(int threads, int n_staged, int min_blocks, bool linear, int* max_blocks_per_sm, int* regs);
cudaError_t mega_launch(const FrameParams& p, const LaunchShape& shape, int min_blocks, cudaStream_t st);

// block wavefront (pool.cu): one persistent kernel, a wavefront per thread block over a pool of paths in shared memory
struct PoolShape { int threads, min_blocks, pool_paths, tree; };
cudaError_t pool_query(const PoolShape& want, int* blocks_per_sm, int* regs, int* smem_bytes);
cudaError_t pool_launch(const FrameParams& p, const PoolShape& shape, int blocks, cudaStream_t st);
// the same kernel compiled with motion-aware node boxes (motion.cu): p.sc.nodes = boxes at time 0, p.sc.node_dt = change to time 1
cudaError_t pool_query_motion(const PoolShape& want, int* blocks_per_sm, int* regs, int* smem_bytes);
cudaError_t pool_launch_motion(const FrameParams& p, const PoolShape& shape, int blocks, cudaStream_t st);
// the same kernel (one shape: 512 threads x 2 blocks) compiled with media anywhere in world::hit's visit order (stages.cu)
cudaError_t pool_query_stages(const PoolShape& want, int* blocks_per_sm, int* regs, int* smem_bytes);
cudaError_t pool_launch_stages(const FrameParams& p, const PoolShape& shape, int blocks, cudaStream_t st);
// zero / resolve the exact frame over the pixels THIS call renders (all of them, or the rank's 8-row bands)
cudaError_t zero_exact_launch(unsigned long long* d_exact, int n_pixels_local, int band_px, int tile_mod, int tile_rem, cudaStream_t st);
cudaError_t resolve_exact_tiles_launch(const unsigned long long* d_exact, int n_pixels_local, int band_px, int tile_mod, int tile_rem, float4* d_accum, cudaStream_t st);

// wavefront (wavefront.cu)
struct WavefrontBuffers;            // opaque SoA queues
cudaError_t wavefront_alloc(WavefrontBuffers** out, int n_paths);
void wavefront_free(WavefrontBuffers* b);
size_t wavefront_bytes(int n_paths);
cudaError_t wavefront_render(const FrameParams& p, WavefrontBuffers* buf, int n_paths, int sm_count, cudaStream_t st, uint64_t* launches);

// parity hook + tone pipeline (render.cu)
cudaError_t trace_launch(const DeviceScene& sc, const float* d_rays, int n, mhit_record* d_out, mhit_medium_probe* d_probes,
                         int brute_force, const int32_t* d_mat_offsets, cudaStream_t st);
cudaError_t trace_launch_motion(const DeviceScene& sc, const float* d_rays, int n, mhit_record* d_out, mhit_medium_probe* d_probes,
                                int brute_force, const int32_t* d_mat_offsets, cudaStream_t st);
cudaError_t accumulate_exact_launch(unsigned long long* d_sum, const unsigned long long* d_frame, int n_pixels, cudaStream_t st);
cudaError_t resolve_exact_launch(const unsigned long long* d_exact, int n_pixels, float4* d_accum, cudaStream_t st);
cudaError_t tonemap_launch(const float4* d_accum, int n_pixels, float scale, uint8_t* d_rgba8, cudaStream_t st);

}  // namespace mort
