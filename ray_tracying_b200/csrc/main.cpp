// main.cpp -- command-line front end with the reference binary's switches
// (reference Code/raytracer.cpp:356-487):
//     Raytracer -input scene.json [-output out.ppm] [-bvh] [-s N] [-light_sample N]
// The reference resolves -input under ../../ASCII/ and -output under ../../Output/ relative to
// its build directory; we do the same unless the argument contains a '/', in which case it is
// taken as a path. Extra switches (not in the reference): -depth D, -seed S, -textures DIR,
// -ids FILE (raw int32 primary hit-ID buffer), -stats, -gpus N (shard the frame over N GPUs of this
// box by screen tiles, rt_render_multi; 0 = all visible GPUs), -tile W H, -tile_block B.
// Everything is done through the C ABI in include/rt_render.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rt_render.h"

int main(int argc, char* argv[]) {
    rt_render_params params;
    rt_render_params_default(&params);
    std::string scene_name, output_name = "output.ppm", texture_dir, ids_file;
    bool print_stats = false;
    int gpus = 1;

    for (int i = 1; i < argc; ++i) {
        if (std::strcmp(argv[i], "-bvh") == 0) params.use_bvh = 1;
        else if (std::strcmp(argv[i], "-s") == 0 && i + 1 < argc) params.samples_sqrt = std::atoi(argv[++i]);
        else if (std::strcmp(argv[i], "-light_sample") == 0 && i + 1 < argc) params.light_samples = std::atoi(argv[++i]);
        else if (std::strcmp(argv[i], "-input") == 0 && i + 1 < argc) scene_name = argv[++i];
        else if (std::strcmp(argv[i], "-output") == 0 && i + 1 < argc) output_name = argv[++i];
        else if (std::strcmp(argv[i], "-depth") == 0 && i + 1 < argc) params.max_depth = std::atoi(argv[++i]);
        else if (std::strcmp(argv[i], "-seed") == 0 && i + 1 < argc) params.seed = std::strtoull(argv[++i], nullptr, 10);
        else if (std::strcmp(argv[i], "-textures") == 0 && i + 1 < argc) texture_dir = argv[++i];
        else if (std::strcmp(argv[i], "-ids") == 0 && i + 1 < argc) ids_file = argv[++i];
        else if (std::strcmp(argv[i], "-stats") == 0) print_stats = true;
        else if (std::strcmp(argv[i], "-gpus") == 0 && i + 1 < argc) gpus = std::atoi(argv[++i]);
        else if (std::strcmp(argv[i], "-tile") == 0 && i + 2 < argc) { params.tile_w = std::atoi(argv[++i]); params.tile_h = std::atoi(argv[++i]); }
        else if (std::strcmp(argv[i], "-tile_block") == 0 && i + 1 < argc) params.reserved[5] = std::atoi(argv[++i]);
    }
    if (scene_name.empty()) {
        std::fprintf(stderr, "Error: Please specify scene file name\n");
        std::printf("Correct usage: ./Raytracer -name {scene_file_name.json}\n");
        return 1;
    }
    const std::string scene_file = scene_name.find('/') == std::string::npos ? "../../ASCII/" + scene_name : scene_name;
    const std::string output_file = output_name.find('/') == std::string::npos ? "../../Output/" + output_name : output_name;

    rt_scene* scene = nullptr;
    if (rt_scene_load_json(scene_file.c_str(), texture_dir.empty() ? nullptr : texture_dir.c_str(), &scene) != RT_OK) {
        std::fprintf(stderr, "An error occurred: %s\n", rt_last_error());
        return 1;
    }
    int32_t width = 0, height = 0, n_shapes = 0;
    rt_scene_resolution(scene, &width, &height);
    rt_scene_counts(scene, &n_shapes, nullptr, nullptr, nullptr, nullptr);
    if (width == 0 || height == 0) {
        std::fprintf(stderr, "Error: Camera resolution is 0. Check scene.json.\n");
        rt_scene_destroy(scene);
        return 1;
    }
    if (n_shapes == 0) std::fprintf(stderr, "Warning: No shapes loaded to render.\n");
    std::printf("BVH built. Mode: %s\n", params.use_bvh ? "ON" : "OFF");
    std::printf("Rendering %dx%d with %dx%d samples and %d light sampling points ...\n", width, height,
                params.samples_sqrt, params.samples_sqrt, params.light_samples);

    std::vector<uint8_t> rgb((size_t)width * height * 3);
    std::vector<int32_t> ids;
    if (!ids_file.empty()) ids.resize((size_t)width * height);
    rt_render_stats stats;
    if (gpus <= 0) gpus = rt_device_count();
    const int rc_render = gpus > 1 ? rt_render_multi(scene, &params, gpus, nullptr, rgb.data(), ids.empty() ? nullptr : ids.data(), nullptr, &stats)
                                   : rt_render(scene, &params, rgb.data(), ids.empty() ? nullptr : ids.data(), nullptr, &stats);
    if (rc_render != RT_OK) {
        std::fprintf(stderr, "An error occurred: %s\n", rt_last_error());
        rt_scene_destroy(scene);
        return 1;
    }
    std::printf("Rendering complete.\n");
    if (print_stats)
        std::printf("rays %llu (primary %llu, shadow %llu, secondary %llu), %d GPU%s, kernel %.3f ms, %.1f Mrays/s, end to end %.3f ms\n",
                    (unsigned long long)stats.rays, (unsigned long long)stats.primary_rays,
                    (unsigned long long)stats.shadow_rays, (unsigned long long)stats.secondary_rays, gpus, gpus > 1 ? "s" : "",
                    stats.kernel_ms, stats.kernel_ms > 0 ? (double)stats.rays / stats.kernel_ms * 1e-3 : 0.0, stats.total_ms);
    int rc = 0;
    if (rt_write_ppm(output_file.c_str(), width, height, rgb.data()) != RT_OK) {
        std::fprintf(stderr, "%s\n", rt_last_error());
        rc = 1;
    } else {
        std::printf("Image written to %s\n", output_file.c_str());
    }
    if (!ids_file.empty()) {
        FILE* f = std::fopen(ids_file.c_str(), "wb");
        if (f) { std::fwrite(ids.data(), sizeof(int32_t), ids.size(), f); std::fclose(f); }
        else rc = 1;
    }
    rt_scene_destroy(scene);
    return rc;
}
