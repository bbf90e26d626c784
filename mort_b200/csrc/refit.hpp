// refit.hpp — entry point of the bottom-up bounds pass (refit.cu); all pointers are device pointers of one committed scene.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "device_types.h"

namespace mort {

struct RefitArgs {
    Bvh4Node* nodes; Bvh4Node* node_t1; int n_nodes;
    const SphereGeom* spheres; const QuadRec* quads; const Instance* instances; int n_spheres, n_quads;
    void* sphere_box[2]; void* quad_box[2];      // 24 B per record; [1] only with motion
    unsigned* extent_key;
    std::vector<int> level_first;                // breadth-first levels of the tree (BuildStats::level_first)
    float cam_center[3];
    bool motion;
};
// Recomputes every box of the tree from the device records.  motion: nodes <- boxes at time 0, node_t1 <- (time 1 - time 0).
// Returns after the kernels are queued on `st` (one stream synchronisation inside, for the scene extent the padding needs).
cudaError_t refit_run(const RefitArgs& a, cudaStream_t st, float* pad_out, float* extent_out);

}  // namespace mort
