// scene.cpp -- scene.json loader, constructor-level scene creation, transforms, bounding boxes,
// P3 PPM I/O. Host side of the drop-in boundary.
//
// Mirrors (same float operations, same order; compile with -ffp-contract=off):
//   Camera::readCameraSpec        reference Code/camera.cpp:14-58
//   parse_material                reference Code/json_loader.cpp:30-97
//   load_lights_from_json         reference Code/json_loader.cpp:103-158
//   load_shapes_from_json         reference Code/json_loader.cpp:164-338
//   Shapes::buildTransformationMatrices  reference Code/shapes.cpp:92-139
//   *::get_bounding_box           reference Code/shapes.cpp:264-287, 335-343, 425-433, 496-503
//   Image::read / Image::write    reference Code/image.cpp:53-133
#include "scene.hpp"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <unordered_map>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "json_min.hpp"

namespace rtb {
namespace {

using jsonmin::Value;

// ---- 4x4 helpers (shapes.cpp:92-149) ---------------------------------------------------------
void mat_mul(const float A[4][4], const float B[4][4], float R[4][4]) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 4; ++k) acc += A[i][k] * B[k][j];
            R[i][j] = acc;
        }
}

void mat_set(float M[4][4], float a00, float a01, float a02, float a03, float a10, float a11, float a12, float a13,
             float a20, float a21, float a22, float a23) {
    float v[4][4] = {{a00, a01, a02, a03}, {a10, a11, a12, a13}, {a20, a21, a22, a23}, {0, 0, 0, 1}};
    std::memcpy(M, v, sizeof(v));
}

// The reference calls the unqualified C functions cos()/sin() on floats (shapes.cpp:101-103);
// with libstdc++ that resolves to the double overloads, and the result is narrowed to float.
inline float cos_ref(float r) { return (float)::cos((double)r); }
inline float sin_ref(float r) { return (float)::sin((double)r); }

void build_transforms(const float t[3], const float r[3], const float s[3], HostPrim& p) {
    float scale_m[4][4], rot_m[4][4], trans_m[4][4];
    mat_set(scale_m, s[0], 0, 0, 0, 0, s[1], 0, 0, 0, 0, s[2], 0);
    const float cx = cos_ref(r[0]), sx = sin_ref(r[0]);
    const float cy = cos_ref(r[1]), sy = sin_ref(r[1]);
    const float cz = cos_ref(r[2]), sz = sin_ref(r[2]);
    mat_set(rot_m, cy * cz, sx * sy * cz - cx * sz, cx * sy * cz + sx * sz, 0,
                   cy * sz, sx * sy * sz + cx * cz, cx * sy * sz - sx * cz, 0,
                   -sy,     sx * cy,                cx * cy,                0);
    mat_set(trans_m, 1, 0, 0, t[0], 0, 1, 0, t[1], 0, 0, 1, t[2]);
    float rs[4][4];
    mat_mul(rot_m, scale_m, rs);
    mat_mul(trans_m, rs, p.o2w);

    float inv_s[4][4], inv_r[4][4], inv_t[4][4];
    mat_set(inv_s, 1.0f / s[0], 0, 0, 0, 0, 1.0f / s[1], 0, 0, 0, 0, 1.0f / s[2], 0);
    mat_set(inv_r, rot_m[0][0], rot_m[1][0], rot_m[2][0], 0,
                   rot_m[0][1], rot_m[1][1], rot_m[2][1], 0,
                   rot_m[0][2], rot_m[1][2], rot_m[2][2], 0);
    mat_set(inv_t, 1, 0, 0, -t[0], 0, 1, 0, -t[1], 0, 0, 1, -t[2]);
    float sr[4][4];
    mat_mul(inv_s, inv_r, sr);
    mat_mul(sr, inv_t, p.w2o);
}

// transformPoint (shapes.cpp:151-158): the bottom row is (0,0,0,1) by construction, w == 1.
void xform_point(const float m[4][4], const float p[3], float out[3]) {
    const float w = m[3][0] * p[0] + m[3][1] * p[1] + m[3][2] * p[2] + m[3][3];
    for (int i = 0; i < 3; ++i) out[i] = m[i][0] * p[0] + m[i][1] * p[1] + m[i][2] * p[2] + m[i][3];
    if (std::fabs((double)(w - 1.0f)) > (double)1e-6f && w != 0) { out[0] /= w; out[1] /= w; out[2] /= w; }
}

void box_reset(Box& b) {
    for (int i = 0; i < 3; ++i) { b.lo[i] = FLT_MAX; b.hi[i] = -FLT_MAX; }
}
void box_add(Box& b, const float p[3]) {
    for (int i = 0; i < 3; ++i) { b.lo[i] = std::min(b.lo[i], p[i]); b.hi[i] = std::max(b.hi[i], p[i]); }
}

void compute_box(HostPrim& p) {
    box_reset(p.box);
    if (p.type == RT_SPHERE) {
        static const float c[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                      {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
        for (int k = 0; k < 8; ++k) {
            float w[3];
            xform_point(p.o2w, c[k], w);
            box_add(p.box, w);
            float moved[3] = {w[0] + p.velocity[0], w[1] + p.velocity[1], w[2] + p.velocity[2]};
            box_add(p.box, moved);
        }
    } else if (p.type == RT_CUBE) {
        static const float c[8][3] = {{-0.5f, -0.5f, -0.5f}, {0.5f, -0.5f, -0.5f}, {0.5f, 0.5f, -0.5f}, {-0.5f, 0.5f, -0.5f},
                                      {-0.5f, -0.5f, 0.5f},  {0.5f, -0.5f, 0.5f},  {0.5f, 0.5f, 0.5f},  {-0.5f, 0.5f, 0.5f}};
        for (int k = 0; k < 8; ++k) { float w[3]; xform_point(p.o2w, c[k], w); box_add(p.box, w); }
    } else if (p.type == RT_RECTANGLE) {
        static const float c[4][3] = {{-0.5f, -0.5f, 0}, {0.5f, -0.5f, 0}, {0.5f, 0.5f, 0}, {-0.5f, 0.5f, 0}};
        for (int k = 0; k < 4; ++k) { float w[3]; xform_point(p.o2w, c[k], w); box_add(p.box, w); }
    } else {  // RT_PLANE, shapes.cpp:496-503
        const float padding = 1e-4f;
        for (int k = 0; k < 4; ++k) box_add(p.box, p.corners[k]);
        for (int i = 0; i < 3; ++i) { p.box.lo[i] -= padding; p.box.hi[i] += padding; }
    }
}

// Plane normal (shapes.cpp:446-451): cross(c1-c0, c2-c0), normalised; invalid if |n| < 1e-6.
void compute_plane_normal(HostPrim& p) {
    float e1[3], e2[3];
    for (int i = 0; i < 3; ++i) { e1[i] = p.corners[1][i] - p.corners[0][i]; e2[i] = p.corners[2][i] - p.corners[0][i]; }
    float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const float len = (float)::sqrt((double)(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]));
    p.normal_valid = !(len < 1e-6f);
    for (int i = 0; i < 3; ++i) p.normal[i] = n[i] / len;
}

void identity(float M[4][4]) { mat_set(M, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0); }

HostPrim make_prim(int type, int material, const float t[3], const float r[3], const float s[3], const float vel[3],
                   const float* corners12) {
    HostPrim p;
    p.type = type;
    p.material = material;
    identity(p.w2o);
    identity(p.o2w);
    std::memset(p.corners, 0, sizeof(p.corners));
    if (type == RT_PLANE) {
        std::memcpy(p.corners, corners12, 12 * sizeof(float));
        compute_plane_normal(p);
    } else {
        build_transforms(t, r, s, p);
        if (type == RT_SPHERE) std::memcpy(p.velocity, vel, 3 * sizeof(float));
    }
    compute_box(p);
    return p;
}

// ---- materials ------------------------------------------------------------------------------
rt_material_desc default_material() {  // material.hpp:52-70
    rt_material_desc m{};
    m.diffuse_color[0] = m.diffuse_color[1] = m.diffuse_color[2] = 0.8f;
    m.specular_color[0] = m.specular_color[1] = m.specular_color[2] = 1.0f;
    m.k_ambient = 0.1f; m.k_diffuse = 0.9f; m.k_specular = 0.3f;
    m.shininess = 20.0f; m.roughness = 0.0f;
    m.reflectivity = 0.0f; m.transparency = 0.0f; m.refractive_index = 1.0f;
    m.texture = -1;
    return m;
}

struct MaterialTable {
    std::vector<rt_material_desc>& out;
    std::unordered_map<uint64_t, std::vector<int>> index;  // FNV-1a of the record -> ids with that hash
    explicit MaterialTable(std::vector<rt_material_desc>& o) : out(o) {}
    int intern(const rt_material_desc& m) {
        const unsigned char* b = reinterpret_cast<const unsigned char*>(&m);
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(m); ++i) { h ^= b[i]; h *= 1099511628211ull; }
        std::vector<int>& ids = index[h];
        for (int id : ids) if (std::memcmp(&out[(size_t)id], &m, sizeof(m)) == 0) return id;
        const int id = (int)out.size();
        out.push_back(m);
        ids.push_back(id);
        return id;
    }
};

struct TextureTable {
    HostScene& scene;
    std::string dir;
    std::unordered_map<std::string, int> index;  // resolved path -> texture id or -1
    int load(const std::string& texture_file) {
        // json_loader.cpp:78-80: drop the 3-character extension, append "ppm".
        if (texture_file.size() < 3) return -1;
        std::string name = texture_file.substr(0, texture_file.size() - 3) + "ppm";
        std::string path = dir + "/" + name;
        auto it = index.find(path);
        if (it != index.end()) return it->second;
        Texture t;
        int id = -1;
        if (read_ppm_p3(path, t) && t.width != 0) {
            id = (int)scene.textures.size();
            scene.textures.push_back(std::move(t));
        } else {
            std::cerr << "Warning: Failed to load texture file: " << path << std::endl;
        }
        index.emplace(path, id);
        return id;
    }
};

// parse_material (json_loader.cpp:30-97). The texture is returned by NAME: elements are converted on
// several threads, and the texture table (file I/O, ids in first-use order) is filled afterwards, serially,
// in element order. `warn` collects what the reference would have printed.
rt_material_desc parse_material(const Value& mj, std::string& texture_file, std::string& warn) {
    rt_material_desc m = default_material();
    texture_file.clear();
    try {
        if (const Value* v = mj.find("diffuse_color")) v->as_float3(m.diffuse_color);
        if (const Value* v = mj.find("specular_color")) v->as_float3(m.specular_color);
        m.k_ambient = mj.value_float("k_ambient", 0.1f);
        m.k_diffuse = mj.value_float("k_diffuse", 0.6f);
        m.k_specular = mj.value_float("k_specular", 0.6f);
        float roughness = mj.value_float("roughness", 0.001f);
        roughness = std::max(0.001f, roughness);
        const float r = std::max(0.001f, std::min(1.0f, roughness));
        m.shininess = 5.0f / (r * r);
        m.roughness = mj.value_float("roughness", 0.0f);
        m.reflectivity = mj.value_float("reflectivity", 0.0f);
        m.transparency = mj.value_float("transparency", 0.0f);
        m.refractive_index = mj.value_float("refractive_index", 1.0f);
        if (const Value* tf = mj.find("texture_file")) {
            if (tf->is_string() && !tf->str().empty()) texture_file = tf->str();
        }
    } catch (const std::exception& e) {
        warn += std::string("Warning: Error parsing material data: ") + e.what() + "\n";
        texture_file.clear();
        return default_material();
    }
    return m;
}

void load_camera(const Value& root, rt_camera_desc& c) {
    // camera.cpp:26-48. Missing blocks are an error here (the reference prints and keeps zeros,
    // which main() then rejects as "Camera resolution is 0", raytracer.cpp:403).
    if (!root.contains("cameras") || !root.contains("render"))
        throw std::runtime_error("JSON file is missing required keys (cameras, render)");
    const Value& cams = root.at("cameras");
    if (!cams.is_array() || cams.array().empty()) throw std::runtime_error("'cameras' must be a non-empty array");
    const Value& cj = cams.array()[0];
    c.focal_length = cj.at("focal_length").as_float();
    c.aperture = cj.value_float("aperture", 0.0f);
    c.focus_dist = cj.value_float("focus_dist", 10.0f);
    cj.at("location").as_float3(c.location);
    cj.at("gaze_vector").as_float3(c.gaze);
    cj.at("up_vector").as_float3(c.up);
    c.sensor_width = cj.at("sensor_width").as_int();
    c.sensor_height = cj.at("sensor_height").as_int();
    c.res_x = root.at("render").at("resolution_x").as_int();
    c.res_y = root.at("render").at("resolution_y").as_int();
}

void load_lights(const Value& root, std::vector<rt_light_desc>& lights) {
    const Value* lj = root.find("lights");
    if (!lj) { std::cerr << "Warning: No valid lights were loaded." << std::endl; return; }
    if (!lj->is_array()) { std::cerr << "Warning: 'lights' key found but is not an array. No lights loaded." << std::endl; return; }
    for (const Value& l : lj->array()) {
        if (!l.is_object()) { std::cerr << "Warning: Skipping non-object entry in 'lights' array." << std::endl; continue; }
        try {
            if (!l.contains("location") || !l.contains("color") || !l.contains("intensity")) {
                std::cerr << "Warning: Skipping invalid light definition." << std::endl;
                continue;
            }
            rt_light_desc d{};
            l.at("location").as_float3(d.location);
            l.at("color").as_float3(d.color);
            d.intensity = l.at("intensity").as_float();
            d.radius = l.value_float("radius", 0.0f);
            if (d.intensity <= 0) { std::cerr << "Warning: Skipping light with non-positive intensity." << std::endl; continue; }
            lights.push_back(d);
        } catch (const std::exception& e) {
            std::cerr << "Warning: Error parsing light entry: " << e.what() << std::endl;
        }
    }
    if (lights.empty()) std::cerr << "Warning: No valid lights were loaded." << std::endl;
}

// One element of "spheres" / "cubes" / "rectangles" / "planes" after conversion (json_loader.cpp:180-332).
// The primitive is written straight into its slot of HostScene::prims; what the serial pass still needs
// (material record, texture name, warnings) travels here.
struct ShapeOut {
    bool ok = false;
    HostPrim& prim;            // material = -1 until the serial pass interns `mat`
    rt_material_desc mat;
    std::string texture_file;  // resolved in the serial pass
    std::string warn;          // what the reference would have printed for this element
    explicit ShapeOut(HostPrim& slot) : prim(slot) {}
};

// A parsed "material" block: what parse_material returned for its text.
struct MaterialEntry {
    rt_material_desc mat;
    std::string texture_file, warn;
};

// `material` = the element's parsed material block, or nullptr when the element has none (default material).
void convert_shape(int type, const Value& j, const MaterialEntry* material, ShapeOut& o) {
    static const float zero3[3] = {0, 0, 0};
    o.ok = false;
    if (!j.is_object()) return;
    auto material_of = [&](const Value&) {
        o.mat = default_material();
        o.texture_file.clear();
        if (material) { o.mat = material->mat; o.texture_file = material->texture_file; o.warn += material->warn; }
    };
    try {
        if (type == RT_SPHERE) {  // json_loader.cpp:180-234
            float t[3], r[3] = {0, 0, 0}, sc[3] = {1, 1, 1}, vel[3] = {0, 0, 0};
            j.at("location").as_float3(t);
            if (const Value* v = j.find("rotation")) v->as_float3(r);
            const Value* sv = j.find("scale");
            if (sv && sv->is_array()) sv->as_float3(sc);
            else if (const Value* rv = j.find("radius")) { float rad = rv->as_float(); sc[0] = sc[1] = sc[2] = rad; }
            material_of(j);
            if (const Value* v = j.find("velocity")) v->as_float3(vel);
            vel[0] = vel[0] / 5; vel[1] = vel[1] / 5; vel[2] = vel[2] / 5;
            o.prim = make_prim(RT_SPHERE, -1, t, r, sc, vel, nullptr);
        } else if (type == RT_CUBE) {  // json_loader.cpp:237-278
            if (!j.contains("translation") || !j.contains("rotation")) {
                o.warn += "Warning: Skipping invalid cube definition.\n";
                return;
            }
            float t[3], r[3], sc[3] = {1, 1, 1};
            j.at("translation").as_float3(t);
            j.at("rotation").as_float3(r);
            if (const Value* sv = j.find("scale")) {
                if (sv->is_array()) sv->as_float3(sc);
                else if (sv->is_number()) { float v = sv->as_float(); sc[0] = sc[1] = sc[2] = v; }
            }
            material_of(j);
            o.prim = make_prim(RT_CUBE, -1, t, r, sc, zero3, nullptr);
        } else if (type == RT_RECTANGLE) {  // json_loader.cpp:282-301
            float t[3], r[3], sc[3];
            j.at("translation").as_float3(t);
            j.at("rotation").as_float3(r);
            j.at("scale").as_float3(sc);
            material_of(j);
            o.prim = make_prim(RT_RECTANGLE, -1, t, r, sc, zero3, nullptr);
        } else {  // planes, json_loader.cpp:304-332
            const Value* cj = j.find("corners");
            if (!cj || !cj->is_array() || cj->array().size() != 4) {
                o.warn += "Warning: Skipping invalid plane definition.\n";
                return;
            }
            float corners[12];
            for (int k = 0; k < 4; ++k) cj->array()[k].as_float3(corners + 3 * k);
            material_of(j);
            o.prim = make_prim(RT_PLANE, -1, zero3, zero3, zero3, zero3, corners);
        }
        o.ok = true;
    } catch (const std::exception& e) {
        static const char* what[4] = {"Warning: Error parsing sphere: ", "Warning: Error parsing cube entry: ",
                                      "Warning: Error parsing rectangle: ", "Warning: Error parsing plane entry: "};
        o.warn += std::string(what[type]) + e.what() + "\n";
    }
}

unsigned host_threads() {
    static const unsigned n = [] {
        const char* e = std::getenv("RT_B200_HOST_THREADS");
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        return e ? (unsigned)std::max(1, std::atoi(e)) : std::min(hw, 64u);
    }();
    return n;
}

// The four shape arrays, in the loader's push order (spheres, cubes, rectangles, planes). Their elements
// were left as text ranges by the document parse (jsonmin::DeferredArray): here every element is parsed and
// converted on its own, chunks of elements on different threads; interning of materials and loading of
// textures -- whose ids depend on first-use order -- and the warnings follow serially in element order, so
// the scene is the one a serial loader builds.
void load_shapes(const std::vector<jsonmin::DeferredArray>& arrays, HostScene& s, MaterialTable& mats, TextureTable& textures) {
    static const char* keys[4] = {"spheres", "cubes", "rectangles", "planes"};
    static const int types[4] = {RT_SPHERE, RT_CUBE, RT_RECTANGLE, RT_PLANE};
    size_t total = 0;
    for (const jsonmin::DeferredArray& a : arrays) total += a.elements.size();
    s.prims.reserve(total);
    for (int c = 0; c < 4; ++c) {
        const jsonmin::DeferredArray* arr = nullptr;
        for (const jsonmin::DeferredArray& a : arrays) if (a.key == keys[c]) arr = &a;
        if (!arr || arr->elements.empty()) continue;
        const size_t n = arr->elements.size();
        const size_t base = s.prims.size();
        s.prims.resize(base + n);
        // per element: ok flag, material record; texture names and warnings are rare and kept per thread
        std::vector<uint8_t> ok(n, 0);
        std::vector<rt_material_desc> mat(n);
        struct Note { size_t i; std::string texture_file, warn; };
        const unsigned threads = (unsigned)std::max<size_t>(1, std::min<size_t>(host_threads(), n / 2048 + 1));
        std::vector<std::vector<Note>> notes(threads);
        std::vector<std::string> fatal(threads);
        auto work = [&](unsigned t) {
            const size_t lo = n * t / threads, hi = n * (t + 1) / threads;
            // material blocks memoised by their text (per thread): a scene repeats a handful of them millions of times
            std::unordered_map<std::string, MaterialEntry> seen;
            for (size_t i = lo; i < hi; ++i) {
                Value j;
                const MaterialEntry* material = nullptr;
                try {
                    std::pair<const char*, const char*> mr;
                    j = jsonmin::parse_range_skipping(arr->elements[i].first, arr->elements[i].second, "material", mr);
                    if (mr.first) {
                        std::string text(mr.first, mr.second);
                        auto it = seen.find(text);
                        if (it == seen.end()) {
                            MaterialEntry e;
                            const Value mj = jsonmin::parse_range(mr.first, mr.second);
                            e.mat = parse_material(mj, e.texture_file, e.warn);
                            if (seen.size() > 4096) seen.clear();  // scenes with a material per shape: no point in remembering them
                            it = seen.emplace(std::move(text), std::move(e)).first;
                        }
                        material = &it->second;
                    }
                } catch (const std::exception& e) {  // malformed JSON: the whole document is rejected, like the reference's parse
                    if (fatal[t].empty()) fatal[t] = e.what();
                    return;
                }
                ShapeOut o(s.prims[base + i]);
                convert_shape(types[c], j, material, o);
                ok[i] = o.ok ? 1 : 0;
                mat[i] = o.mat;
                if (!o.texture_file.empty() || !o.warn.empty()) notes[t].push_back({i, std::move(o.texture_file), std::move(o.warn)});
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; ++t) pool.emplace_back(work, t);
        work(0);
        for (std::thread& th : pool) th.join();
        for (const std::string& f : fatal) if (!f.empty()) throw std::runtime_error(f);
        // serial, in element order: warnings, textures and materials (their ids depend on first use), compaction
        unsigned nt = 0;
        size_t ni = 0, kept = base;
        for (size_t i = 0; i < n; ++i) {
            while (nt < threads && ni >= notes[nt].size()) { ++nt; ni = 0; }
            const Note* note = (nt < threads && notes[nt][ni].i == i) ? &notes[nt][ni++] : nullptr;
            if (note && !note->warn.empty()) std::cerr << note->warn << std::flush;
            if (!ok[i]) continue;
            if (note && !note->texture_file.empty()) mat[i].texture = textures.load(note->texture_file);
            if (kept != base + i) s.prims[kept] = s.prims[base + i];
            s.prims[kept].material = mats.intern(mat[i]);
            ++kept;
        }
        s.prims.resize(kept);
    }
    if (s.prims.empty()) std::cerr << "Warning: No valid shapes were loaded." << std::endl;
}

void normalize3(const float v[3], float out[3]) {  // camera.cpp:60-68
    const float mag = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (mag == 0.0f) { out[0] = out[1] = out[2] = 0.0f; return; }
    out[0] = v[0] / mag; out[1] = v[1] / mag; out[2] = v[2] / mag;
}
void cross3(const float a[3], const float b[3], float out[3]) {  // camera.cpp:70-87
    const float x = a[1] * b[2] - a[2] * b[1];
    const float y = a[2] * b[0] - a[0] * b[2];
    const float z = a[0] * b[1] - a[1] * b[0];
    out[0] = x; out[1] = y; out[2] = z;
}

}  // namespace

void finalize_scene(HostScene& s) {
    // Camera basis (camera.cpp:109-115); the reference recomputes it for every ray with the same
    // operands, so computing it once gives the same bits.
    float tmp[3];
    normalize3(s.cam.gaze, s.zdir);
    cross3(s.cam.up, s.zdir, tmp);
    normalize3(tmp, s.xdir);
    cross3(s.zdir, s.xdir, tmp);
    normalize3(tmp, s.ydir);
    build_bvh(s);
    flatten_scene(s);
}

void load_scene_json(const std::string& path, const std::string& texture_dir, HostScene& s) {
    const bool timing = std::getenv("RT_B200_DEBUG") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = now();
    // the file is mapped, not copied: the shape elements are parsed straight out of the page cache
    struct Mapping {
        const char* data = nullptr;
        size_t size = 0;
        int fd = -1;
        ~Mapping() {
            if (data && size) ::munmap(const_cast<char*>(data), size);
            if (fd >= 0) ::close(fd);
        }
    } text;
    {
        text.fd = ::open(path.c_str(), O_RDONLY);
        if (text.fd < 0) throw std::runtime_error("Could not open JSON file: " + path);
        struct stat st;
        if (::fstat(text.fd, &st) != 0 || !S_ISREG(st.st_mode)) throw std::runtime_error("Could not open JSON file: " + path);
        text.size = (size_t)st.st_size;
        if (text.size > 0) {
            void* m = ::mmap(nullptr, text.size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, text.fd, 0);
            if (m == MAP_FAILED) throw std::runtime_error("Could not map JSON file: " + path);
            text.data = static_cast<const char*>(m);
        }
    }
    const auto t1 = now();
    // the document without the elements of the four shape arrays (those stay text ranges into `text`)
    std::vector<jsonmin::DeferredArray> shape_arrays;
    jsonmin::Parser parser(text.data, text.data + text.size);
    parser.defer({"spheres", "cubes", "rectangles", "planes"}, &shape_arrays);
    parser.set_threads(host_threads());
    Value root = parser.parse_document();
    const auto t2 = now();
    if (!root.is_object()) throw std::runtime_error("scene root must be a JSON object");
    load_camera(root, s.cam);
    load_lights(root, s.lights);
    MaterialTable mats(s.materials);
    TextureTable textures{s, texture_dir.empty() ? std::string("../../Textures") : texture_dir, {}};
    load_shapes(shape_arrays, s, mats, textures);
    const auto t3 = now();
    finalize_scene(s);
    if (timing)
        std::fprintf(stderr, "[rt_b200] load %s: read %.3f s, parse %.3f s, shapes %.3f s, bvh %.3f s, flatten %.3f s\n", path.c_str(),
                     secs(t0, t1), secs(t1, t2), secs(t2, t3), s.build_seconds, secs(t3, now()) - s.build_seconds);
}

void create_scene_from_desc(const rt_scene_desc& d, HostScene& s) {
    s.cam = d.camera;
    s.lights.assign(d.lights, d.lights + d.n_lights);
    s.materials.assign(d.materials, d.materials + d.n_materials);
    for (int i = 0; i < d.n_textures; ++i) {
        Texture t;
        t.width = d.textures[i].width;
        t.height = d.textures[i].height;
        t.rgb.assign(d.textures[i].rgb, d.textures[i].rgb + (size_t)t.width * t.height * 3);
        s.textures.push_back(std::move(t));
    }
    if (s.materials.empty()) s.materials.push_back(default_material());
    for (const rt_material_desc& m : s.materials)
        if (m.texture >= d.n_textures) throw std::runtime_error("material references a texture that does not exist");
    s.prims.reserve(d.n_shapes);
    for (int i = 0; i < d.n_shapes; ++i) {
        const rt_shape_desc& sh = d.shapes[i];
        if (sh.type < RT_SPHERE || sh.type > RT_PLANE) throw std::runtime_error("unknown shape type");
        if (sh.material < 0 || sh.material >= (int)s.materials.size()) throw std::runtime_error("shape material out of range");
        s.prims.push_back(make_prim(sh.type, sh.material, sh.translation, sh.rotation, sh.scale, sh.velocity, sh.corners));
    }
    finalize_scene(s);
}

// ---- P3 PPM (image.cpp:53-133) ----------------------------------------------------------------
bool read_ppm_p3(const std::string& path, Texture& out) {
    std::ifstream file(path);
    if (!file.is_open()) { std::cerr << "Error: Could not open file " << path << " for reading\n"; return false; }
    std::string magic, line;
    file >> magic;
    if (magic != "P3") { std::cerr << "Error: Only P3 PPM format is supported\n"; return false; }
    file >> std::ws;
    while (file.peek() == '#') { std::getline(file, line); file >> std::ws; }
    int w = 0, h = 0, maxc = 0;
    file >> w >> h >> maxc;
    if (w <= 0 || h <= 0) return false;
    if (maxc != 255) std::cerr << "Warning: Max color value is " << maxc << ", expected 255\n";
    out.width = w;
    out.height = h;
    out.rgb.resize((size_t)w * h * 3);
    for (size_t i = 0; i < out.rgb.size(); ++i) {
        int v = 0;
        file >> v;
        out.rgb[i] = (uint8_t)std::max(0, std::min(v, 255));
    }
    return true;
}

bool write_ppm_p3(const std::string& path, int width, int height, const uint8_t* rgb) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) { std::cerr << "Error: Could not open file " << path << " for writing\n"; return false; }
    // Same byte stream as Image::write: "P3\nW H\n255\n", pixels as "r g b" joined by two spaces.
    std::string buf;
    buf.reserve((size_t)width * 14 + 16);
    std::fprintf(f, "P3\n%d %d\n255\n", width, height);
    char tmp[16];
    for (int y = 0; y < height; ++y) {
        buf.clear();
        for (int x = 0; x < width; ++x) {
            const uint8_t* p = rgb + ((size_t)y * width + x) * 3;
            int n = std::snprintf(tmp, sizeof(tmp), "%d %d %d", (int)p[0], (int)p[1], (int)p[2]);
            buf.append(tmp, (size_t)n);
            if (x < width - 1) buf.append("  ");
        }
        buf.push_back('\n');
        std::fwrite(buf.data(), 1, buf.size(), f);
    }
    std::fclose(f);
    return true;
}

}  // namespace rtb
