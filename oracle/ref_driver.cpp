// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A thin driver around the UNMODIFIED reference sources under /root/reference/Code.
// It is linked against the reference's own object files (shapes.cpp, acceleration.cpp,
// camera.cpp, image.cpp, json_loader.cpp and raytracer.cpp compiled with
// -Dmain=reference_main), see oracle/Makefile; nothing from the reference is copied here.
// Outputs go to oracle/_ref/ only.
//
// What it does (all through the reference's public functions):
//   * loads a scene.json with Camera(), load_lights_from_json(), load_shapes_from_json()
//     (reference json_loader.cpp:103,164; camera.cpp:239) and builds BVH (acceleration.cpp:7);
//   * "ids" mode  : per pixel, centre ray via Camera::pixelToRay_thin_lens (camera.cpp:97)
//     then BVH::get_intersection (acceleration.cpp:142) -> primitive index (load order) and t;
//   * "render" mode: the frame loop of raytracer.cpp:433-476 (stratified samples, Trace(),
//     gamma 1.1, clamp, *255.999) for a band of rows, with a caller-given mt19937 seed
//     instead of std::random_device so runs are repeatable;
//   * "bvh" mode  : dumps the reference tree in pre-order (boxes + leaf primitive indices);
//   * --depth D   : Trace() is entered at depth (MAX_RECURSION_DEPTH - D); the reference only
//     uses `depth` for its `depth > MAX_RECURSION_DEPTH` cut-off (raytracer.cpp:290), so this
//     is exactly a reference whose limit is D, without touching its source.
//
// The reference's BVH object is not thread-safe (temp_hit_vec member) and its RNG is one serial
// stream, so parallelism for timing is by PROCESS over row bands (--rows), see bench.py.

#include "raytracer.hpp"
#include "camera.hpp"
#include "image.hpp"
#include "json_loader.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Options {
    std::string scene, mode = "render", out_ppm, out_raw, out_ids, out_bvh;
    bool use_bvh = true;
    int s = 1, light_samples = 1, depth = MAX_RECURSION_DEPTH;
    unsigned seed = 1;
    int row0 = 0, row1 = -1;
    int col0 = 0, col1 = -1;   // render / ids modes: only columns [col0,col1) of the rows (outputs keep the full width)
    float fixed_time = -1.0f;  // ids mode: ray.time; <0 -> 0
    int repeat = 1;            // render mode: render the band this many times, report each time
};

void usage() {
    std::fprintf(stderr,
        "ref_driver --scene F [--mode render|ids|bvh] [--bvh 0|1] [--s N] [--light-samples N]\n"
        "           [--depth D] [--seed S] [--rows Y0 Y1] [--cols X0 X1] [--time T] [--repeat N]\n"
        "           [--out-ppm F] [--out-raw F] [--out-ids F] [--out-bvh F]\n");
}

bool parse(int argc, char** argv, Options& o) {
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&](const char* what) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", what); std::exit(2); }
            return argv[++i];
        };
        if (a == "--scene") o.scene = next("--scene");
        else if (a == "--mode") o.mode = next("--mode");
        else if (a == "--bvh") o.use_bvh = std::atoi(next("--bvh")) != 0;
        else if (a == "--s") o.s = std::atoi(next("--s"));
        else if (a == "--light-samples") o.light_samples = std::atoi(next("--light-samples"));
        else if (a == "--depth") o.depth = std::atoi(next("--depth"));
        else if (a == "--seed") o.seed = (unsigned)std::strtoul(next("--seed"), nullptr, 10);
        else if (a == "--rows") { o.row0 = std::atoi(next("--rows")); o.row1 = std::atoi(next("--rows")); }
        else if (a == "--cols") { o.col0 = std::atoi(next("--cols")); o.col1 = std::atoi(next("--cols")); }
        else if (a == "--time") o.fixed_time = (float)std::atof(next("--time"));
        else if (a == "--repeat") o.repeat = std::max(1, std::atoi(next("--repeat")));
        else if (a == "--out-ppm") o.out_ppm = next("--out-ppm");
        else if (a == "--out-raw") o.out_raw = next("--out-raw");
        else if (a == "--out-ids") o.out_ids = next("--out-ids");
        else if (a == "--out-bvh") o.out_bvh = next("--out-bvh");
        else { usage(); return false; }
    }
    if (o.scene.empty()) { usage(); return false; }
    if (o.depth < 0 || o.depth > MAX_RECURSION_DEPTH) {
        std::fprintf(stderr, "--depth must be in [0,%d]\n", MAX_RECURSION_DEPTH);
        return false;
    }
    return true;
}

void dump_node(const node& n, const std::unordered_map<const Shapes*, int>& index, FILE* f) {
    const bool leaf = !n.left && !n.right;
    std::fprintf(f, "%c %a %a %a %a %a %a", leaf ? 'L' : 'I',
                 n.bounding_box.min_point[0], n.bounding_box.min_point[1], n.bounding_box.min_point[2],
                 n.bounding_box.max_point[0], n.bounding_box.max_point[1], n.bounding_box.max_point[2]);
    if (leaf) {
        std::fprintf(f, " %zu", n.objects.size());
        for (const Shapes* s : n.objects) std::fprintf(f, " %d", index.at(s));
    }
    std::fprintf(f, "\n");
    if (n.left) dump_node(*n.left, index, f);
    if (n.right) dump_node(*n.right, index, f);
}

}  // namespace

int main(int argc, char** argv) {
    Options opt;
    if (!parse(argc, argv, opt)) return 2;

    try {
        Camera camera(opt.scene);
        auto [width, height] = camera.getResolution();
        if (width <= 0 || height <= 0) { std::fprintf(stderr, "bad resolution\n"); return 1; }
        std::vector<Light> lights = load_lights_from_json(opt.scene);
        std::vector<std::unique_ptr<Shapes>> shapes = load_shapes_from_json(opt.scene);

        std::vector<Shapes*> ptrs;
        std::unordered_map<const Shapes*, int> index;
        for (size_t i = 0; i < shapes.size(); ++i) {
            ptrs.push_back(shapes[i].get());
            index[shapes[i].get()] = (int)i;
        }
        auto t_b0 = std::chrono::steady_clock::now();
        BVH bvh(ptrs);
        auto t_b1 = std::chrono::steady_clock::now();
        const double build_s = std::chrono::duration<double>(t_b1 - t_b0).count();

        const int row0 = std::max(0, opt.row0);
        const int row1 = (opt.row1 < 0 || opt.row1 > height) ? height : opt.row1;
        const int col0 = std::max(0, opt.col0);
        const int col1 = (opt.col1 < 0 || opt.col1 > width) ? width : opt.col1;

        if (opt.mode == "bvh") {
            FILE* f = opt.out_bvh.empty() ? stdout : std::fopen(opt.out_bvh.c_str(), "w");
            if (!f) return 1;
            std::fprintf(f, "shapes %zu\n", shapes.size());
            if (!shapes.empty()) dump_node(bvh.root, index, f);
            if (f != stdout) std::fclose(f);
            return 0;
        }

        std::mt19937 gen(opt.seed);
        std::uniform_real_distribution<double> dist(0.0, 1.0);

        if (opt.mode == "ids") {
            std::vector<int> ids((size_t)width * (row1 - row0));
            std::vector<float> ts((size_t)width * (row1 - row0));
            auto t0 = std::chrono::steady_clock::now();
            for (int y = row0; y < row1; ++y)
                for (int x = col0; x < col1; ++x) {
                    auto [origin, direction] = camera.pixelToRay_thin_lens({x + 0.5f, y + 0.5f}, gen, dist);
                    Ray ray = {origin, direction};
                    ray.time = opt.fixed_time < 0 ? 0.0f : opt.fixed_time;
                    Hit hit = bvh.get_intersection(ray, opt.use_bvh);
                    size_t k = (size_t)(y - row0) * width + x;
                    ids[k] = hit.shape ? index.at(hit.shape) : -1;
                    ts[k] = hit.t;
                }
            auto t1 = std::chrono::steady_clock::now();
            if (!opt.out_ids.empty()) {
                FILE* f = std::fopen(opt.out_ids.c_str(), "wb");
                if (!f) return 1;
                int hdr[4] = {width, row1 - row0, row0, 0};
                std::fwrite(hdr, sizeof(int), 4, f);
                std::fwrite(ids.data(), sizeof(int), ids.size(), f);
                std::fwrite(ts.data(), sizeof(float), ts.size(), f);
                std::fclose(f);
            }
            std::printf("{\"mode\":\"ids\",\"width\":%d,\"rows\":%d,\"seconds\":%.6f,\"build_seconds\":%.6f}\n",
                        width, row1 - row0, std::chrono::duration<double>(t1 - t0).count(), build_s);
            return 0;
        }

        if (opt.mode != "render") { usage(); return 2; }

        // The reference frame loop (raytracer.cpp:433-476) for rows [row0,row1).
        const int start_depth = MAX_RECURSION_DEPTH - opt.depth;
        const int rows = row1 - row0;
        std::vector<unsigned char> rgb((size_t)width * rows * 3);
        std::vector<float> raw((size_t)width * rows * 3);
        std::vector<double> times;
        for (int rep = 0; rep < opt.repeat; ++rep) {
        gen.seed(opt.seed);
        auto t0 = std::chrono::steady_clock::now();
        for (int y = row0; y < row1; ++y) {
            for (int x = col0; x < col1; ++x) {
                Color c = {0.0f, 0.0f, 0.0f};
                if (opt.s <= 1) {
                    auto [origin, direction] = camera.pixelToRay_thin_lens({x + 0.5f, y + 0.5f}, gen, dist);
                    Ray ray = {origin, direction};
                    ray.time = (float)dist(gen);
                    c = Trace(ray, bvh, lights, start_depth, opt.use_bvh, gen, dist, opt.light_samples);
                } else {
                    for (int j = 0; j < opt.s; ++j)
                        for (int i = 0; i < opt.s; ++i) {
                            double ox = dist(gen), oy = dist(gen);
                            double sx = (i + ox) / opt.s, sy = (j + oy) / opt.s;
                            auto [origin, direction] = camera.pixelToRay_thin_lens({x + sx, y + sy}, gen, dist);
                            Ray ray = {origin, direction};
                            ray.time = (float)dist(gen);
                            c = c + Trace(ray, bvh, lights, start_depth, opt.use_bvh, gen, dist, opt.light_samples);
                        }
                    c = c / (float)(opt.s * opt.s);
                }
                const float gamma = 1.1f;
                float ch[3] = {std::pow(c.r, 1.0f / gamma), std::pow(c.g, 1.0f / gamma), std::pow(c.b, 1.0f / gamma)};
                size_t k = ((size_t)(y - row0) * width + x) * 3;
                raw[k] = c.r; raw[k + 1] = c.g; raw[k + 2] = c.b;
                for (int q = 0; q < 3; ++q) {
                    int v = static_cast<int>(std::max(0.0f, std::min(1.0f, ch[q])) * 255.999);
                    rgb[k + q] = (unsigned char)std::max(0, std::min(v, 255));
                }
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        times.push_back(std::chrono::duration<double>(t1 - t0).count());
        }  // repeat
        const double secs = times.back();

        if (!opt.out_ppm.empty()) {
            Image img(width, rows);
            for (int y = 0; y < rows; ++y)
                for (int x = 0; x < width; ++x) {
                    size_t k = ((size_t)y * width + x) * 3;
                    img.setPixel(x, y, rgb[k], rgb[k + 1], rgb[k + 2]);
                }
            img.write(opt.out_ppm);  // reference image.cpp:53 (P3 writer)
        }
        if (!opt.out_raw.empty()) {
            FILE* f = std::fopen(opt.out_raw.c_str(), "wb");
            if (!f) return 1;
            int hdr[4] = {width, rows, row0, 0};
            std::fwrite(hdr, sizeof(int), 4, f);
            std::fwrite(rgb.data(), 1, rgb.size(), f);
            std::fwrite(raw.data(), sizeof(float), raw.size(), f);
            std::fclose(f);
        }
        std::string all;
        for (size_t i = 0; i < times.size(); ++i) { char b[32]; std::snprintf(b, sizeof(b), "%s%.6f", i ? "," : "", times[i]); all += b; }
        std::printf("{\"mode\":\"render\",\"width\":%d,\"rows\":%d,\"cols\":%d,\"spp\":%d,\"seconds\":%.6f,\"all_seconds\":[%s],\"build_seconds\":%.6f}\n",
                    width, rows, col1 - col0, opt.s <= 1 ? 1 : opt.s * opt.s, secs, all.c_str(), build_s);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
