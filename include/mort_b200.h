/* mort_b200.h — C ABI of libmort_b200.so, the B200-native drop-in for the render path of lgleznah/mort.
 *
 * The reference has no FFI / plugin boundary: it is one translation unit whose observable surface is
 *   (1) the scene-builder calls scene code makes          /root/reference/world.cuh:27-102, object / material /
 *                                                          texture constructors (cited per function below)
 *   (2) the Camera's public fields + initialize()         /root/reference/camera.cuh:12-84
 *   (3) one render call per frame that fills an RGBA8      /root/reference/mort.cu:44-47,99-106;
 *       bottom-up image                                    /root/reference/camera.cuh:178-208
 *   (4) `mort <scene 1-10>`                                /root/reference/mort.cu:633-689
 * Each entry point below replaces the piece of that surface it cites.  Plain pointers and sizes only; every
 * call returns 0 on success or a negative mort_status, never exits the process (the reference exit()s on
 * CUDA errors, /root/reference/include/book.h:21-30) and never falls back to the CPU: without a CUDA
 * device mort_create fails.  A context is single-threaded, owns all device memory it allocates, and the
 * caller owns every buffer it passes in.
 *
 * Handles are the reference's: (type tag, array slot) pairs (tags in include/mort_scene_format.h).
 */
#ifndef MORT_B200_H
#define MORT_B200_H
#include <stddef.h>
#include <stdint.h>
#include "mort_scene_format.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mort_ctx mort_ctx;
typedef struct { int32_t type, idx; } mort_handle;

enum mort_status {
    MORT_OK = 0, MORT_ERR_ARG = -1, MORT_ERR_CUDA = -2, MORT_ERR_SCENE = -3, MORT_ERR_STATE = -4, MORT_ERR_IO = -5
};

/* ---- lifetime ---------------------------------------------------------------------------------------- */
/* replaces main()'s implicit device 0 + cudaDeviceSetLimit (mort.cu:695-709) */
int mort_create(int cuda_device, mort_ctx** out);
int mort_destroy(mort_ctx* ctx);
const char* mort_last_error(const mort_ctx* ctx);
/* launches go to this cudaStream_t.  Default: the context's own stream, a BLOCKING stream — it is ordered after
 * everything the host queued on the legacy default stream (e.g. a framework's fill of a buffer handed to
 * mort_render_device) and before what it queues there afterwards.  A host that works on other (non-blocking) streams
 * passes the stream its buffers are ready on; every device pointer given to a mort_* call must be ready on that stream. */
int mort_set_stream(mort_ctx* ctx, void* cuda_stream);

/* ---- scenes ------------------------------------------------------------------------------------------- */
/* the switch of mort.cu:649-689: scene_id 1..10 = the shipped scenes, anything else = empty world */
int mort_build_scene(mort_ctx* ctx, int scene_id, const char* asset_dir);
/* BASELINE.json config 4: scene-1 recipe over cells [-G,G)^2 (no reference counterpart: objects.cuh:746 caps at 1100) */
int mort_build_sphere_field(mort_ctx* ctx, int G, uint64_t seed, int camera_kind);
int mort_load_scene(mort_ctx* ctx, const char* mscn_path, const char* asset_dir);
int mort_dump_scene(mort_ctx* ctx, const char* mscn_path);
/* Scene text (grammar: mort_b200/csrc/scene_text.cpp): one builder call per statement, so user scenes and the synthetic
 * fields need no recompile (the reference's scenes are C++ functions, mort.cu:129-631).  Dump writes the journal of the
 * builder calls this context received + the camera; loading it back rebuilds the same arrays slot for slot. */
int mort_load_scene_text(mort_ctx* ctx, const char* path, const char* asset_dir);
int mort_dump_scene_text(mort_ctx* ctx, const char* path);
int mort_clear_scene(mort_ctx* ctx);                                        /* world::clear, world.cuh:92-96 */

/* textures.cuh:20,42,79,164 + world.cuh:76-90 */
int mort_add_solid(mort_ctx* ctx, float r, float g, float b, mort_handle* out);
int mort_add_checker(mort_ctx* ctx, float scale, mort_handle even, mort_handle odd, mort_handle* out);
int mort_add_image(mort_ctx* ctx, const uint8_t* rgb8_rows_top_down, int width, int height, mort_handle* out);
int mort_add_noise(mort_ctx* ctx, float scale, mort_handle* out);           /* draws from the context's host rand() stream */
/* materials.cuh:36,71,104-105,149,180 + world.cuh:56-74 */
int mort_add_lambertian(mort_ctx* ctx, mort_handle tex, mort_handle* out);
int mort_add_metal(mort_ctx* ctx, float r, float g, float b, float fuzz, mort_handle* out);
int mort_add_dielectric(mort_ctx* ctx, float ior, mort_handle* out);
int mort_add_diffuse_light(mort_ctx* ctx, mort_handle tex, mort_handle* out);
int mort_add_isotropic(mort_ctx* ctx, mort_handle tex, mort_handle* out);
/* objects.cuh:38,46,170,258,296,384,459-469,529 + world.cuh:27-54; skip = "reachable only through a parent" */
int mort_add_sphere(mort_ctx* ctx, const float center[3], float radius, mort_handle mat, int skip, mort_handle* out);
int mort_add_moving_sphere(mort_ctx* ctx, const float center0[3], const float center1[3], float radius, mort_handle mat, int skip, mort_handle* out);
int mort_add_quad(mort_ctx* ctx, const float Q[3], const float u[3], const float v[3], mort_handle mat, int skip, mort_handle* out);
int mort_add_translate(mort_ctx* ctx, mort_handle obj, const float offset[3], int skip, mort_handle* out);
int mort_add_rotate_y(mort_ctx* ctx, mort_handle obj, float degrees, int skip, mort_handle* out);
int mort_add_constant_medium(mort_ctx* ctx, mort_handle boundary, float density, mort_handle mat, int skip, mort_handle* out);
int mort_add_list(mort_ctx* ctx, int skip, mort_handle* out);
int mort_list_add(mort_ctx* ctx, mort_handle list, mort_handle obj);
int mort_add_bvh(mort_ctx* ctx, mort_handle list, int skip, mort_handle* out);   /* marks the world bvh_mode (world.cuh:51-54) */
int mort_add_box(mort_ctx* ctx, const float a[3], const float b[3], mort_handle mat);                                 /* utils.h:51-67 */
int mort_add_rotated_box(mort_ctx* ctx, const float size[3], const float translation[3], float degrees, mort_handle mat, mort_handle* out);   /* utils.h:69-96 */
/* rand() of the context's host stream (glibc TYPE_3, unseeded) for scene code that draws like mort.cu does */
int mort_host_rand(mort_ctx* ctx);

/* ---- camera (camera.cuh:12-84) ------------------------------------------------------------------------ */
typedef struct {
    float aspect_ratio; int32_t image_width, samples_per_pixel, bounce_limit, vfov;
    float background[3]; float lookfrom[3], lookat[3], vup[3]; float defocus_angle, focus_dist;
    int32_t light_obj_type, light_obj_idx;      /* light_obj_type = -1: no light sampling (mort.cu:215) */
} mort_camera_desc;
int mort_get_camera(mort_ctx* ctx, mort_camera_desc* out);
int mort_set_camera(mort_ctx* ctx, const mort_camera_desc* desc);           /* runs Camera::initialize */
/* the overrides BASELINE.json's configs need: <= 0 keeps the scene's value */
int mort_override_camera(mort_ctx* ctx, int image_width, float aspect_ratio, int samples_per_pixel, int bounce_limit);
int mort_get_camera_record(mort_ctx* ctx, mscn_camera* out);               /* every derived field, for parity checks */

/* ---- commit = world::toDevice (world.cuh:98-102): flatten, SAH-build the 4-wide BVH, upload ------------ */
int mort_commit(mort_ctx* ctx);

/* ---- tree build on the GPU, refit, motion-aware bounds ----------------------------
 * The reference builds its BVH on the host with a median split and a bubble sort, at most 1024 nodes (objects.cuh:521,528-661),
 * and bounds a moving sphere by the union of its end boxes (objects.cuh:46-55).  Here the binned-SAH build also runs on
 * the GPU (gpu_build.cu) and gives the SAME tree as the host builder, node for node; MORT_BUILD_AUTO uses it from 16384
 * leaves up.  Options apply to the next mort_commit. */
enum { MORT_BUILD_AUTO = 0, MORT_BUILD_HOST = 1, MORT_BUILD_GPU = 2 };
typedef struct {
    int32_t builder;        /* MORT_BUILD_* */
    int32_t max_leaf;       /* primitives per leaf, 1..4 (0 = 4) */
    float k_trav;           /* SAH cost of a node step relative to one sphere test (0 = 1.0) */
    int32_t host_threads;   /* host builder: 0 = all cores */
    int32_t gpu_small;      /* GPU builder: subtrees of at most this many primitives are built by one thread each (0 = 64) */
    int32_t gpu_flags;      /* GPU builder: bit 0 = plain per-thread atomics (no warp / block aggregation) */
    int32_t motion_bounds;  /* 1 = node boxes of subtrees with moving spheres are stored at time 0 and time 1 and interpolated at the ray's
                               time (tighter than the reference's union box; hits are unchanged: the sphere test itself is exact) */
    int32_t reserved;
} mort_build_opts;
typedef struct {
    int32_t built_on_gpu, gpu_levels, gpu_small_subtrees, bvh2_nodes;
    int32_t motion_nodes, refits, reserved[2];
    double flatten_ms, build_ms, gpu_stream_ms, refit_ms;   /* commit: host flattening without the build; the build; its stream time; last mort_refit */
    uint64_t gpu_workspace_bytes;
} mort_build_info;
int mort_set_build_opts(mort_ctx* ctx, const mort_build_opts* opts);         /* NULL = defaults */
int mort_get_build_info(mort_ctx* ctx, mort_build_info* out);
/* Dynamic scenes: move / resize a sphere of the committed scene (center1 = NULL: a static sphere), then mort_refit: the
 * topology of the tree is kept, the primitive boxes and every node box are recomputed bottom-up on the GPU from the device
 * records (one kernel per tree level).  Hits after a refit equal those of a fresh commit of the edited scene (the tree is
 * merely less tight).  MORT_ERR_STATE without a committed scene. */
int mort_update_sphere(mort_ctx* ctx, mort_handle sphere, const float center0[3], const float center1[3], float radius);
int mort_refit(mort_ctx* ctx);

/* ---- render (renderKernel, mort.cu:44-47,99-106; Camera::render, camera.cuh:178-208) ------------------- */
/* Three schedulers over the same per-ray code, Philox stream and estimator:
 *   MEGAKERNEL  persistent warps, a lane owns a path from camera to termination (render.cu)
 *   WAVEFRONT   one kernel per stage over SoA queues in HBM (wavefront.cu)
 *   POOL        block wavefront: one persistent kernel, every thread block alternates trace / (classify /) shade phases over a
 *               pool of paths in its shared memory, shading sorted by material class (pool.cu)
 * MEGAKERNEL and POOL accumulate exactly (integers) and render bit-identical frames.
 * A scene in which a constant_medium is reached through translate / rotate_y / list wrappers (hitDispatch allows it, objects.cuh:875-877;
 * no shipped scene does it) is always rendered by the block wavefront's general-media build (stages.cu), whatever `mode` says. */
enum { MORT_MODE_MEGAKERNEL = 0, MORT_MODE_WAVEFRONT = 1, MORT_MODE_POOL = 2,
       MORT_MODE_AUTO = 3 /* MEGAKERNEL or POOL, whichever measured faster for the committed scene's class (the `mort` CLI's default) */ };
typedef struct {
    uint32_t seed, frame;          /* Philox key; the reference's seed is 69420 (mort.cu:707) */
    int32_t mode;                  /* MORT_MODE_* (default: MORT_MODE_POOL) */
    int32_t sample_mod, sample_rem;/* sample-split across GPUs: this call renders strata rows s_j % mod == rem (1,0 = all) */
    int32_t stage_nodes;           /* megakernel: BVH nodes staged in shared memory: <= 0 none (default), N first N (breadth-first) */
    int32_t threads_per_block;     /* 0 = default */
    int32_t blocks_per_sm;         /* 0 = default */
    int32_t wavefront_paths;       /* paths in flight for the wavefront mode, 0 = default */
    int32_t exact_accum;           /* 1: d_accum receives W*H x 4 uint64 (exact fixed-point sums, see mort_resolve_exact_device) */
    int32_t tile_mod, tile_rem;    /* tile-split across GPUs (megakernel): this call renders the 8-row bands b with b % mod == rem and leaves
                                      every other pixel of d_accum untouched (0,0 or 1,0 = whole frame) */
    int32_t accumulate;            /* exact_accum only: 1 = ADD this call's sums to the contents of d_accum (progressive rendering:
                                      render frame 0 with 0, frames 1.. with 1, a new `frame` each time), 0 = overwrite */
    int32_t pool_paths;            /* MORT_MODE_POOL: paths per thread-block pool (0 = default 1024; 88 B of shared memory each) */
    int32_t pool_refill;           /* MORT_MODE_POOL, tree scenes: lanes whose ray left the tree take a new one as soon as this many lanes of
                                      their warp are idle (1..31; 0 = default, -1 = off: fixed 32-ray chunks) */
    int32_t n_frames;              /* MORT_MODE_POOL, mort_render_device: > 1 = a batch: frame keys frame .. frame + n - 1 in ONE launch into n consecutive
                                      frames of d_accum (the reference's frame loop, mort.cu:93-120, without a kernel tail per frame); <= 256 */
    int32_t tile_rows;             /* tile split: rows per band, 0 = 8.  Thin bands balance the ranks better (sky vs geometry): mort_group_render uses 2 */
    int32_t pool_flags;            /* MORT_MODE_POOL experiments: bit 0 = barrier between trace and classify, bit 1 = force the overlapped form (0 = by scene class) */
} mort_render_opts;
void mort_default_render_opts(mort_render_opts* o);

/* Accumulation image: W*H float4, rows bottom-up like the reference's frame (camera.cuh:70-78):
 * xyz = sum of the sample colours (IEEE: a NaN sample poisons the channel exactly as in camera.cuh:190-198),
 * w = number of samples that contained a NaN. */
/* Device-resident frame: d_accum is a device pointer owned by the caller (e.g. a torch tensor). */
int mort_render_device(mort_ctx* ctx, const mort_render_opts* opts, void* d_accum);
/* Exact accumulation image (opts.exact_accum = 1, megakernel only): W*H x 4 uint64 per pixel = Q39.24 fixed-point sums of
 * r, g, b and a word of NaN / +inf sample counts.  Integer sums are associative, so partial frames of a sample split can be
 * added in ANY order (e.g. by an int64 SUM all-reduce) and the N-GPU frame equals the 1-GPU frame bit for bit.
 * mort_resolve_exact_device turns it into the float4 accumulation image described above. */
int mort_resolve_exact_device(mort_ctx* ctx, const void* d_exact, void* d_accum);
/* ---- progressive accumulation + checkpoint / resume --------------------------------------------------------
 * The reference renders the same frame again on every idle callback and shows the latest one (mort.cu:99-119,
 * gpu_anim.h) — nothing accumulates and nothing survives the process.  Here frames with different `frame` keys are
 * independent sample sets of the same image, and their exact sums add in any order, so a long render can be split
 * into frames, stopped, saved and resumed without changing a bit of the result.
 * Limits of the flag word when many frames are added: 2^20-1 NaN samples and 2^14-1 +inf samples per channel per pixel. */
/* d_sum += d_frame (both W*H x 4 uint64, device).  Also merges partial frames rendered elsewhere. */
int mort_accumulate_exact_device(mort_ctx* ctx, void* d_sum, const void* d_frame);
/* 64-bit fingerprint of the committed scene + camera (what a checkpoint is only valid for) */
int mort_scene_fingerprint(mort_ctx* ctx, uint64_t* out);
/* Checkpoint file = {magic, version, W, H, samples per pixel per frame, frames_done, seed, fingerprint} + the exact sums.
 * Load verifies W, H, samples per frame and the fingerprint against the committed scene (MORT_ERR_STATE on mismatch). */
int mort_save_checkpoint(mort_ctx* ctx, const char* path, const void* d_sum, uint32_t seed, uint32_t frames_done);
int mort_load_checkpoint(mort_ctx* ctx, const char* path, void* d_sum, uint32_t* seed, uint32_t* frames_done);
/* Host-buffer progressive render: adds n_frames more frames (frame keys frames_done .. frames_done + n_frames - 1) to a
 * context-owned exact image and returns the mean over ALL frames so far, tone-mapped (rgba8_out, may be NULL) and as float
 * sums (accum_out, may be NULL; divide by samples_per_pixel * *frames_total).  checkpoint_path (may be NULL): with
 * resume != 0 an existing file is loaded first (its seed must equal opts->seed), and the file is rewritten after the last
 * frame.  Rendering k frames, stopping, and resuming for the rest gives the same bits as rendering them in one call. */
int mort_render_progressive(mort_ctx* ctx, const mort_render_opts* opts, int n_frames, const char* checkpoint_path, int resume,
                            uint8_t* rgba8_out, float* accum_out, uint32_t* frames_total);
/* forget the running image (it also restarts by itself when the scene, the camera, the frame size or the seed changes) */
int mort_reset_progressive(mort_ctx* ctx);

/* Tone pipeline of camera.cuh:194-207 on device: mean over n_samples, NaN flush, gamma 2, quantise to RGBA8
 * (w=255).  d_rgba8 is a device pointer to W*H*4 bytes. */
int mort_tonemap_device(mort_ctx* ctx, const void* d_accum, int samples_per_pixel_total, void* d_rgba8);
/* Host-buffer frame (the reference-facing call): renders, tone-maps and copies back.  rgba8_out: W*H*4 bytes,
 * bottom-up (may be NULL); accum_out: W*H*4 floats (may be NULL). */
int mort_render(mort_ctx* ctx, const mort_render_opts* opts, uint8_t* rgba8_out, float* accum_out);

/* ---- image files (host buffers; no context, no device) ------------------------------------------------------------------
 * rgba8: the frame as mort_render returns it (bottom-up rows, camera.cuh:70-78).  The extension picks the format: ".ppm" (binary
 * P6, top-down) or ".png" (8-bit RGB, top-down).  mort_write_pfm: accum4 = W*H float4 sums, written as scale * rgb, rows bottom-up
 * (PFM's own order), NaN-poisoned pixels stay NaN. */
int mort_write_image(const char* path, const uint8_t* rgba8, int width, int height);
int mort_write_pfm(const char* path, const float* accum4, int width, int height, float scale);

/* ---- multi-GPU: shard by samples or tiles, ONE collective per frame -----------------------------
 * The reference is single-GPU (one renderKernel launch per frame, mort.cu:99-106).  Here every GPU renders its share of
 * the frame independently into an exact partial frame (opts.exact_accum) and the partial frames are combined once per
 * frame by one NCCL sum-reduce of 64-bit integers over NVLink; integer sums are associative, so the N-GPU frame equals
 * the 1-GPU frame bit for bit.  NCCL is bound at run time (dlopen libnccl.so.2); single-GPU use needs no NCCL.
 *
 * (a) one PROCESS per GPU: rank 0 obtains a 128-byte id, the launcher hands it to all ranks (a file, MPI, torch.distributed ...),
 *     every rank attaches its context, renders with sample_mod/sample_rem (or tile_mod/tile_rem) = world/rank and calls
 *     mort_comm_reduce_exact on the ctx's stream; the root then resolves / tone-maps as usual. */
int mort_comm_unique_id(void* id128);
int mort_comm_attach(mort_ctx* ctx, const void* id128, int world, int rank);
int mort_comm_detach(mort_ctx* ctx);
int mort_comm_reduce_exact(mort_ctx* ctx, void* d_exact /* W*H x 4 uint64, in place */, int root);
/* (b) one process drives n GPUs of the box (`mort <scene> --gpus N --split sample|tile`): a context + a host thread per device.
 *     Build / load / commit the SAME scene on every mort_group_ctx(g, r) with the calls above, then mort_group_render =
 *     shares rendered concurrently + one grouped ncclReduce to rank 0 + resolve + tone map + copy to the host buffers. */
typedef struct mort_group mort_group;
enum { MORT_SPLIT_SAMPLE = 0, MORT_SPLIT_TILE = 1 };
typedef struct {
    int32_t n_gpus, split;
    double kernel_ms_max, kernel_ms_min;   /* render kernels, CUDA events per rank: slowest and fastest rank */
    double collective_ms;                  /* the ncclReduce on rank 0's stream */
    uint64_t collective_bytes;             /* partial frame each rank contributes */
    uint64_t segments, samples;            /* summed over ranks */
} mort_group_stats;
int mort_group_create(int n_devices, const int* devices /* NULL: 0..n-1 */, mort_group** out);
int mort_group_destroy(mort_group* g);
int mort_group_size(const mort_group* g);
mort_ctx* mort_group_ctx(mort_group* g, int rank);
const char* mort_group_last_error(const mort_group* g);
int mort_group_render(mort_group* g, const mort_render_opts* opts, int split, uint8_t* rgba8_out, float* accum_out);
int mort_group_get_stats(mort_group* g, mort_group_stats* out);

/* ---- parity hook: closest hit of arbitrary rays (world::hit with the medium loop disabled) -------------- */
enum { MORT_TRACE_BVH = 0, MORT_TRACE_BRUTE_FORCE = 1 };
/* rays7: n x {ox,oy,oz,dx,dy,dz,time} (host).  out: n records (host).  probes: n * mort_stats.n_media (host) or NULL — the boundary
 * probes of the media world::hit's own loop reaches (top level, not hidden), in array order. */
int mort_trace(mort_ctx* ctx, const float* rays7, int n, mhit_record* out, mhit_medium_probe* probes, int flags);

/* ---- stats -------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t width, height, sqrt_spp, bounce_limit;
    int32_t n_leaves, n_spheres, n_quads, n_nodes, bvh_depth, n_media /* top-level, visible */, n_instances, n_materials, n_textures;
    double sah_cost, build_ms, upload_ms;
    double last_render_ms;             /* CUDA-event time of the render kernels of the last mort_render* call */
    uint64_t last_segments;            /* path segments (top-level closest-hit queries) of that call */
    uint64_t last_samples;             /* camera paths of that call */
    uint64_t last_kernel_launches;     /* kernels launched by that call */
    int32_t sm_count, staged_nodes, threads_per_block, blocks_per_sm;
    int32_t regs_per_thread, reserved0;
    uint64_t device_bytes;
} mort_stats;
int mort_get_stats(mort_ctx* ctx, mort_stats* out);

#ifdef __cplusplus
}
#endif
#endif
