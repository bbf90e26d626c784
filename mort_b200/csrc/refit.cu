// refit.cu — bottom-up bounds pass over a committed tree on the GPU: refit after edits, and motion-aware node boxes
// (bodies and the why: refit_core.cuh).  Streaming kernels over the records (32-96 B in, 24 B out each) and the nodes
// (128 B each), one launch per tree level from the deepest up: HBM- / launch-bound, a few hundred microseconds for 10^6 spheres.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "refit.hpp"
#include "refit_core.cuh"

namespace mort {
namespace {

constexpr int kT = 256;
__global__ void __launch_bounds__(kT) k_raw(rf::Ctx c, int n, int motion) { const int i = blockIdx.x * kT + threadIdx.x; if (i < n) rf::body_raw(c, i, motion != 0); }
__global__ void __launch_bounds__(kT) k_pad(rf::Ctx c, int n, int motion) { const int i = blockIdx.x * kT + threadIdx.x; if (i < n) rf::body_pad(c, i, motion != 0); }
__global__ void __launch_bounds__(kT) k_level(rf::Ctx c, int first, int count, int motion) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i < count * 4) rf::body_level(c, first + (i >> 2), i & 3, motion != 0);
}
__global__ void __launch_bounds__(kT) k_delta(rf::Ctx c, int n_nodes) { const int i = blockIdx.x * kT + threadIdx.x; if (i < n_nodes * 4) rf::body_delta(c, i >> 2, i & 3); }

}  // namespace

cudaError_t refit_run(const RefitArgs& a, cudaStream_t st, float* pad_out, float* extent_out) {
    rf::Ctx c; memset(&c, 0, sizeof(c));
    c.nodes = a.nodes; c.node_t1 = a.motion ? a.node_t1 : nullptr;
    c.spheres = a.spheres; c.quads = a.quads; c.instances = a.instances; c.n_spheres = a.n_spheres; c.n_quads = a.n_quads;
    for (int t = 0; t < 2; t++) { c.sphere_box[t] = reinterpret_cast<rf::Box*>(a.sphere_box[t]); c.quad_box[t] = reinterpret_cast<rf::Box*>(a.quad_box[t]); }
    c.extent_key = a.extent_key;
    const int n = a.n_spheres + a.n_quads, motion = a.motion ? 1 : 0;
    if (n == 0 || a.level_first.size() < 2) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(a.extent_key, 0, 4, st);
    if (e != cudaSuccess) return e;
    k_raw<<<(n + kT - 1) / kT, kT, 0, st>>>(c, n, motion);
    float ext = 0.f;
    e = cudaMemcpyAsync(&ext, a.extent_key, 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    float M = ext;
    for (int k = 0; k < 3; k++) M = std::max(M, std::fabs(a.cam_center[k]));
    c.pad = 2e-6f * M;
    k_pad<<<(n + kT - 1) / kT, kT, 0, st>>>(c, n, motion);
    for (int L = (int)a.level_first.size() - 2; L >= 0; L--) {
        const int first = a.level_first[L], count = a.level_first[L + 1] - first;
        if (count > 0) k_level<<<(count * 4 + kT - 1) / kT, kT, 0, st>>>(c, first, count, motion);
    }
    if (motion) k_delta<<<(a.n_nodes * 4 + kT - 1) / kT, kT, 0, st>>>(c, a.n_nodes);
    if (pad_out) *pad_out = c.pad;
    if (extent_out) *extent_out = M;
    return cudaGetLastError();
}

}  // namespace mort
