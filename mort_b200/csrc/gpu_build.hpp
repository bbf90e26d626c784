// gpu_build.hpp — the GPU tree builder's entry point (gpu_build.cu), same contract as build_bvh4 (bvh_build.cpp).
#pragma once
#include <string>
#include <vector>

#include "flatten.hpp"

namespace mort {

// flags: bit 0 = plain per-thread atomics everywhere (no warp / block aggregation; A/B and debugging)
bool gpu_build_bvh4(const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& nodes, std::vector<int>& order_out, BuildStats& stats,
                    const BuildOptions& opt, void* cuda_stream, int flags, std::string* err);

}  // namespace mort
