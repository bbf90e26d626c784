// flatten.hpp — host side of mort_commit(): turns the reference-shaped Scene (handles, lists, instance
// wrappers, media, the reference's own BVH objects) into the flat HBM layout of device_types.h and builds
// the binned-SAH 4-wide BVH over every visible leaf primitive.
#pragma once
#include <string>
#include <vector>

#include "device_types.h"
#include "scene.hpp"

namespace mort {

// One visible sphere / quad as world::hit (world.cuh:104-171) reaches it.  `order` is its position in the
// reference's visit sequence: on bit-equal t the reference keeps the LATER candidate (objects.cuh:73-76,197),
// so "largest order wins" reproduces its tie rule under any traversal order.
struct LeafRef { int type, idx, inst, order, top_type, top_idx; };

struct BuildStats {
    int n_leaves = 0, n_nodes = 0, n_bvh2_nodes = 0, max_depth = 0, n_leaf_slots = 0;
    double sah_cost = 0, build_ms = 0;
    float pad = 0, scene_extent = 0;
    std::vector<int> level_first;              // 4-wide nodes are laid out breadth-first: level d = [level_first[d], level_first[d+1])
    int built_on_gpu = 0, gpu_levels = 0, gpu_small_subtrees = 0;
    size_t gpu_workspace_bytes = 0; double gpu_kernel_ms = 0;   // GPU builder: workspace, CUDA-event time of its kernels
};
struct BuildOptions {
    int max_leaf = MORT_MAX_LEAF; float k_trav = 1.0f;
    int threads = 0;            // host builder: 0 = all cores
    int gpu_small = 0;          // GPU builder: subtrees of at most this many primitives are built by one thread each (0 = 64)
};

struct FlatScene {
    std::vector<Bvh4Node> nodes;
    std::vector<SphereGeom> spheres; std::vector<PrimInfo> sphere_info;
    std::vector<QuadRec> quads;
    std::vector<uint8_t> sphere_cls, quad_cls;   // SHADE_* per record
    std::vector<Instance> instances;
    std::vector<Material> materials;
    std::vector<Texture> textures;
    std::vector<NoiseTables> noises;
    std::vector<Medium> media;
    std::vector<BoundaryPrim> boundary;
    std::vector<LightPrim> lights;
    std::vector<LeafRef> leaves;               // visit order (kept for tests / stats)
    std::vector<uint8_t> sphere_pinned;        // per sphere slot: copied into a light or medium-boundary record (mort_update_sphere then needs a new commit)
    int n_moving = 0;                          // leaves that are moving spheres
    int light_kind = LIGHT_NONE;
    int post_media_order = 0, two_pass = 0, empty = 1, linear = 0;
    CameraParams cam;
    BuildStats stats;
};

// bvh_build.cpp: binned SAH over boxes -> BVH2 -> collapse to 4-wide, breadth-first layout.
struct BuildPrim { float lo[3], hi[3]; int type; int ref; };   // ref = index into FlatScene::leaves
// A tree builder: boxes in, 4-wide nodes (leaf words = positions in order_out) + the leaf order out.  The host builder
// (build_bvh4) is the default; capi.cu passes the GPU builder (gpu_build.cu) for large scenes.  Both give the same tree.
typedef bool (*BvhBuildFn)(void* user, const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& nodes, std::vector<int>& order_out,
                           BuildStats& stats, const BuildOptions& opt, std::string* err);
bool flatten_scene(const Scene& s, FlatScene& out, std::string* err, const BuildOptions& opt = BuildOptions(), BvhBuildFn build = nullptr, void* build_user = nullptr);
void camera_params(const Camera& c, CameraParams& out);

struct Bvh4Leaf { int first, count, type; };                    // range in `order_out`
void build_bvh4(const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& nodes, std::vector<int>& order_out,
                BuildStats& stats, const BuildOptions& opt = BuildOptions());

}  // namespace mort
