/* mort_oracle.h — CPU restatement of the reference's render path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (mort_b200/csrc, libmort_b200.so) never includes, links or calls it.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its own, so
 * this oracle is pinned against outputs of the reference itself, executed on a B200 through
 * oracle/ref_harness.cu and committed under tests/golden/ (scene + camera dumps, primary-hit records,
 * noisy and converged images); see tests/test_oracle_golden.py.
 */
#ifndef MORT_ORACLE_H
#define MORT_ORACLE_H
#include <stdint.h>
#include "mort_scene_format.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_scene oracle_scene;

/* Loads a scene dump (include/mort_scene_format.h).  `rgb` (may be NULL) are the texels of image
 * texture 0, RGB8 rows top-down. */
oracle_scene* oracle_scene_load(const char* mscn_path, const uint8_t* rgb, int rgb_w, int rgb_h);
void oracle_scene_free(oracle_scene* s);
void oracle_scene_camera(const oracle_scene* s, mscn_camera* out);
int oracle_scene_counts(const oracle_scene* s, mscn_header* out);
/* Same overrides as the harness (width / aspect / spp / depth; <= 0 keeps the value), followed by
 * Camera::initialize (camera.cuh:47-84). */
void oracle_camera_override(oracle_scene* s, int width, float aspect, int spp, int depth);

/* world::hit with the medium loop disabled + the two boundary probes of every constant_medium.
 * probes may be NULL; otherwise n * n_medium records. */
int oracle_trace(const oracle_scene* s, const float* rays7, int n, mhit_record* out, mhit_medium_probe* probes);

/* Full frame.  hdr: W*H*4 floats = (sum r, sum g, sum b over samples without NaN, number of samples with a
 * NaN), rows bottom-up like the reference; rgba8: W*H*4 bytes exactly as camera.cuh:194-207 would write
 * them; either may be NULL.  Samples are the strata rows s_j with s_j % sj_mod == sj_rem.  counters[0] =
 * path segments (top-level closest-hit queries), counters[1] = samples. */
int oracle_render(const oracle_scene* s, uint32_t seed, uint32_t frame, int sj_mod, int sj_rem, int n_threads,
                  float* hdr, uint8_t* rgba8, uint64_t counters[2]);

/* Camera::get_ray for one (pixel, stratum) with the canonical Philox stream; out = 7 floats. */
void oracle_camera_ray(const oracle_scene* s, uint32_t seed, uint32_t frame, int x, int y, int s_i, int s_j, float* ray7);

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* first n uniforms of the canonical stream of (pixel, sample) */
void oracle_stream_uniforms(uint32_t seed, uint32_t frame, uint32_t pixel, uint32_t sample, int n, float* out);

#ifdef __cplusplus
}
#endif
#endif
