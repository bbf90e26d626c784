// mort_main.cpp — the `mort <scene 1-10>` entry point (mort.cu:633-689), headless: instead of opening a GLUT
// window and re-rendering forever (gpu_anim.h, out of scope on a headless B200 box) it renders `--frames`
// frames through the C ABI, prints the reference's "Avg. time per frame" line (mort.cu:116-119) and writes
// the image as PPM (top-down; the frame itself is bottom-up like the reference's, camera.cuh:70-78).
// `--accumulate` turns the frame loop into a progressive render (every frame adds a new sample set to the same
// image; `--checkpoint FILE [--resume]` saves / continues it), `--load FILE.mscn` renders a dumped scene.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mort_b200.h"

static int usage() { printf("Usage: mort <number_between_1_and_11> [--width W] [--aspect A] [--spp S] [--depth D] [--seed X] [--frames F]\n"
                            "            [--mode auto|mega|wave|pool] [--stage N] [--bps blocks/SM] [--tpb threads] [--pool paths/block] [--field G [--fieldcam 0|1]]\n"
                            "            [--builder auto|host|gpu] [--motion-bounds] [--gpu-small N] [--gpu-flags N]   tree build (GPU from 16384 leaves up by default)\n"
                            "            [--assets DIR] [--out image.ppm|image.png] [--hdr image.pfm] [--device K] [--load scene.mscn] [--dump scene.mscn]\n"
                            "            [--scene-file scene.txt] [--dump-text scene.txt]\n"
                            "            [--accumulate [--checkpoint FILE [--resume]] [--preview K]]   --preview: rewrite --out every K frames while the image converges\n"
                            "            [--gpus N [--split sample|tile]]   N GPUs of this box, one NCCL reduce of the partial frames per frame\n"); return -1; }

int main(int argc, char** argv) {
    if (argc < 2) return usage();                       // mort.cu:638-641
    int scene = atoi(argv[1]);
    int width = 0, spp = 0, depth = 0, frames = 1, device = 0, stage = 0, mode = MORT_MODE_AUTO, bps = 0, tpb = 0, field = 0, fieldcam = 0, pool = 0, refill = 0, gpus = 1, split = MORT_SPLIT_SAMPLE, xflags = 0;
    float aspect = 0; unsigned seed = 69420; std::string assets = "mort_b200/assets", out, hdr, load, dump, ckpt, text_in, text_out;
    bool accumulate = false, resume = false; int preview = 0;
    mort_build_opts bo; memset(&bo, 0, sizeof(bo));
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto nx = [&]() -> const char* { if (i + 1 >= argc) { usage(); exit(-1); } return argv[++i]; };
        if (a == "--width") width = atoi(nx()); else if (a == "--aspect") aspect = (float)atof(nx());
        else if (a == "--spp") spp = atoi(nx()); else if (a == "--depth") depth = atoi(nx());
        else if (a == "--seed") seed = (unsigned)strtoul(nx(), 0, 10); else if (a == "--frames") frames = atoi(nx());
        else if (a == "--assets") assets = nx(); else if (a == "--out") out = nx(); else if (a == "--hdr") hdr = nx(); else if (a == "--device") device = atoi(nx());
        else if (a == "--stage") stage = atoi(nx());
        else if (a == "--load") load = nx(); else if (a == "--dump") dump = nx();
        else if (a == "--scene-file") text_in = nx(); else if (a == "--dump-text") text_out = nx();
        else if (a == "--accumulate") accumulate = true; else if (a == "--checkpoint") ckpt = nx(); else if (a == "--resume") resume = true; else if (a == "--preview") preview = atoi(nx());
        else if (a == "--bps") bps = atoi(nx()); else if (a == "--tpb") tpb = atoi(nx());
        else if (a == "--field") field = atoi(nx()); else if (a == "--fieldcam") fieldcam = atoi(nx());
        else if (a == "--gpus") gpus = atoi(nx()); else if (a == "--split") { std::string m = nx(); split = m == "tile" ? MORT_SPLIT_TILE : MORT_SPLIT_SAMPLE; }
        else if (a == "--builder") { std::string m = nx(); bo.builder = m == "host" ? MORT_BUILD_HOST : m == "gpu" ? MORT_BUILD_GPU : MORT_BUILD_AUTO; }
        else if (a == "--motion-bounds") bo.motion_bounds = 1; else if (a == "--gpu-small") bo.gpu_small = atoi(nx()); else if (a == "--gpu-flags") bo.gpu_flags = atoi(nx());
        else if (a == "--pool") pool = atoi(nx()); else if (a == "--xflags") xflags = atoi(nx()); else if (a == "--refill") refill = atoi(nx());
        else if (a == "--mode") { std::string m = nx(); mode = m == "wave" ? MORT_MODE_WAVEFRONT : m == "mega" ? MORT_MODE_MEGAKERNEL : m == "pool" ? MORT_MODE_POOL : MORT_MODE_AUTO; }
        else return usage();
    }
    if (gpus > 1) {
        // one process, N GPUs: the same scene is built and committed on every rank's context; each frame is one
        // mort_group_render = shares rendered concurrently + one NCCL reduce + tone map on rank 0
        if (accumulate || !dump.empty() || !text_out.empty()) { fprintf(stderr, "mort: --gpus does not combine with --accumulate / --dump / --dump-text\n"); return -1; }
        mort_group* g = nullptr;
        if (mort_group_create(gpus, nullptr, &g) != MORT_OK) { fprintf(stderr, "mort: cannot drive %d CUDA devices (NCCL and %d GPUs are required)\n", gpus, gpus); return 2; }
        for (int r = 0; r < gpus; r++) {
            mort_ctx* c = mort_group_ctx(g, r);
            int rc = !text_in.empty() ? mort_load_scene_text(c, text_in.c_str(), assets.c_str()) : !load.empty() ? mort_load_scene(c, load.c_str(), assets.c_str())
                     : field > 0 ? mort_build_sphere_field(c, field, 69420, fieldcam) : mort_build_scene(c, scene, assets.c_str());
            if (rc == MORT_OK) rc = mort_override_camera(c, width, aspect, spp, depth);
            if (rc == MORT_OK) rc = mort_set_build_opts(c, &bo);
            if (rc == MORT_OK) rc = mort_commit(c);
            if (rc != MORT_OK) { fprintf(stderr, "mort: rank %d: %s\n", r, mort_last_error(c)); mort_group_destroy(g); return 3; }
        }
        mort_stats st; mort_get_stats(mort_group_ctx(g, 0), &st);
        std::vector<uint8_t> img((size_t)st.width * st.height * 4);
        mort_render_opts o; mort_default_render_opts(&o); o.seed = seed; o.mode = mode; o.blocks_per_sm = bps; o.threads_per_block = tpb; o.pool_paths = pool; o.pool_refill = refill; o.pool_flags = xflags;
        double total = 0, coll = 0, kmin = 0; mort_group_stats gs; memset(&gs, 0, sizeof(gs));
        for (int f = 0; f < frames; f++) {
            o.frame = (uint32_t)f;
            if (mort_group_render(g, &o, split, img.data(), nullptr) != MORT_OK) { fprintf(stderr, "mort: render: %s\n", mort_group_last_error(g)); mort_group_destroy(g); return 3; }
            mort_group_get_stats(g, &gs);
            total += gs.kernel_ms_max + gs.collective_ms; coll += gs.collective_ms; kmin += gs.kernel_ms_min;
            printf("Avg. time per frame: %3.1f ms\n", total / (f + 1));
        }
        const double ms = gs.kernel_ms_max + gs.collective_ms;
        printf("{\"scene\":%d,\"n_gpus\":%d,\"split\":\"%s\",\"width\":%d,\"height\":%d,\"spp_eff\":%d,\"depth\":%d,\"ms\":%.3f,\"kernel_ms_max\":%.3f,\"kernel_ms_min\":%.3f,"
               "\"collective_ms\":%.3f,\"collective_mb\":%.1f,\"msamples_per_s\":%.3f,\"mrays_per_s\":%.3f}\n",
               scene, gpus, split == MORT_SPLIT_TILE ? "tile" : "sample", st.width, st.height, st.sqrt_spp * st.sqrt_spp, st.bounce_limit, ms, gs.kernel_ms_max, gs.kernel_ms_min,
               gs.collective_ms, gs.collective_bytes / 1048576.0, (double)gs.samples / (ms * 1e3), (double)gs.segments / (ms * 1e3));
        if (!out.empty() && mort_write_image(out.c_str(), img.data(), st.width, st.height) != MORT_OK) { fprintf(stderr, "mort: cannot write %s (.ppm or .png)\n", out.c_str()); mort_group_destroy(g); return 4; }
        mort_group_destroy(g);
        return 0;
    }
    mort_ctx* ctx = nullptr;
    if (mort_create(device, &ctx) != MORT_OK) { fprintf(stderr, "mort: no usable CUDA device %d (this renderer has no CPU path)\n", device); return 2; }
    auto die = [&](const char* what) { fprintf(stderr, "mort: %s: %s\n", what, mort_last_error(ctx)); mort_destroy(ctx); return 3; };
    if (!text_in.empty()) { if (mort_load_scene_text(ctx, text_in.c_str(), assets.c_str()) != MORT_OK) return die("scene file"); }
    else if (!load.empty()) { if (mort_load_scene(ctx, load.c_str(), assets.c_str()) != MORT_OK) return die("load"); }
    else if (field > 0) { if (mort_build_sphere_field(ctx, field, 69420, fieldcam) != MORT_OK) return die("sphere field"); }
    else if (mort_build_scene(ctx, scene, assets.c_str()) != MORT_OK) return die("scene");
    if (mort_override_camera(ctx, width, aspect, spp, depth) != MORT_OK) return die("camera");
    if (!dump.empty() && mort_dump_scene(ctx, dump.c_str()) != MORT_OK) return die("dump");
    if (!text_out.empty() && mort_dump_scene_text(ctx, text_out.c_str()) != MORT_OK) return die("dump-text");
    if (mort_set_build_opts(ctx, &bo) != MORT_OK) return die("build options");
    if (mort_commit(ctx) != MORT_OK) return die("commit");
    mort_build_info bi; mort_get_build_info(ctx, &bi);
    mort_stats st; mort_get_stats(ctx, &st);
    std::vector<uint8_t> img((size_t)st.width * st.height * 4);
    std::vector<float> acc(hdr.empty() ? 0 : (size_t)st.width * st.height * 4);
    mort_render_opts o; mort_default_render_opts(&o); o.seed = seed; o.mode = mode; o.stage_nodes = stage; o.blocks_per_sm = bps; o.threads_per_block = tpb; o.pool_paths = pool; o.pool_refill = refill; o.pool_flags = xflags;
    double total = 0;
    uint32_t frames_total = 1;                         // frames averaged into the written image
    if (accumulate) {
        // one call per frame so the reference's running "Avg. time per frame" line keeps its meaning; the checkpoint
        // (if any) is read before the first frame and rewritten after every frame
        for (int f = 0; f < frames; f++) {
            const bool last = f + 1 == frames;
            // the headless stand-in for the reference's live window (gpu_anim.h): the running mean is written out every `preview` frames
            const bool show = last || (preview > 0 && !out.empty() && (f + 1) % preview == 0);
            if (mort_render_progressive(ctx, &o, 1, ckpt.empty() ? nullptr : ckpt.c_str(), (resume || f > 0) ? 1 : 0, show ? img.data() : nullptr,
                                        (last && !acc.empty()) ? acc.data() : nullptr, &frames_total) != MORT_OK) return die("render");
            if (show && !last && mort_write_image(out.c_str(), img.data(), st.width, st.height) != MORT_OK) { fprintf(stderr, "mort: cannot write %s (.ppm or .png)\n", out.c_str()); mort_destroy(ctx); return 4; }
            mort_get_stats(ctx, &st);
            total += st.last_render_ms;
            printf("Avg. time per frame: %3.1f ms\n", total / (f + 1));
        }
        printf("{\"frames_accumulated\":%u,\"spp_total\":%lld}\n", frames_total, (long long)frames_total * st.sqrt_spp * st.sqrt_spp);
    } else
    for (int f = 0; f < frames; f++) {
        o.frame = (uint32_t)f;
        if (mort_render(ctx, &o, img.data(), acc.empty() ? nullptr : acc.data()) != MORT_OK) return die("render");
        mort_get_stats(ctx, &st);
        total += st.last_render_ms;
        printf("Avg. time per frame: %3.1f ms\n", total / (f + 1));
    }
    double samples = (double)st.width * st.height * st.sqrt_spp * st.sqrt_spp;
    printf("{\"scene\":%d,\"width\":%d,\"height\":%d,\"spp_eff\":%d,\"depth\":%d,\"ms\":%.3f,\"msamples_per_s\":%.3f,\"mrays_per_s\":%.3f,\"nodes\":%d,\"leaves\":%d,\"regs\":%d,\"bps\":%d,\"build_ms\":%.2f,\"flatten_ms\":%.2f,\"upload_ms\":%.2f,\"built_on_gpu\":%d,\"gpu_stream_ms\":%.2f,\"gpu_levels\":%d,\"motion_nodes\":%d,\"sah\":%.4f}\n",
           scene, st.width, st.height, st.sqrt_spp * st.sqrt_spp, st.bounce_limit, st.last_render_ms, samples / (st.last_render_ms * 1e3),
           (double)st.last_segments / (st.last_render_ms * 1e3), st.n_nodes, st.n_leaves, st.regs_per_thread, st.blocks_per_sm, st.build_ms,
           bi.flatten_ms, st.upload_ms, bi.built_on_gpu, bi.gpu_stream_ms, bi.gpu_levels, bi.motion_nodes, st.sah_cost);
    if (!out.empty() && mort_write_image(out.c_str(), img.data(), st.width, st.height) != MORT_OK) { fprintf(stderr, "mort: cannot write %s (.ppm or .png)\n", out.c_str()); mort_destroy(ctx); return 4; }
    // linear radiance as PFM (rows bottom-up — the frame's own order, camera.cuh:70-78); NaN-poisoned pixels stay NaN
    if (!hdr.empty() && mort_write_pfm(hdr.c_str(), acc.data(), st.width, st.height, (float)(1.0 / ((double)st.sqrt_spp * st.sqrt_spp * frames_total))) != MORT_OK) {
        fprintf(stderr, "mort: cannot write %s\n", hdr.c_str()); mort_destroy(ctx); return 4;
    }
    mort_destroy(ctx);
    return 0;
}
