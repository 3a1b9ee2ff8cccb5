// rt_api.cpp -- the extern "C" layer declared in include/rt_render.h (host part).
// Device entry points (rt_scene_upload, rt_render_device, rt_render, ...) live in render.cu.
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include "api_internal.hpp"
#include "json_min.hpp"

namespace rtb {
thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace rtb

using rtb::HostScene;

struct rt_scene {
    HostScene host;
};

HostScene* rtb::host_of(rt_scene* s) { return &s->host; }
const HostScene* rtb::host_of(const rt_scene* s) { return &s->host; }

extern "C" {

const char* rt_last_error(void) { return rtb::g_last_error.c_str(); }
int rt_version(void) { return 2; }

void rt_render_params_default(rt_render_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->use_bvh = 0;       // raytracer.cpp:361
    p->samples_sqrt = 4;  // raytracer.cpp:362
    p->light_samples = 1; // raytracer.cpp:363
    p->max_depth = 10;    // raytracer.hpp:11
    p->seed = 1;
    p->fixed_time = -1.0f;
    p->rank = 0;
    p->world = 1;
    p->tile_w = 32;
    p->tile_h = 32;
}

int rt_scene_load_json(const char* scene_path, const char* texture_dir, rt_scene** out) {
    if (!scene_path || !out) { rtb::set_error("rt_scene_load_json: null argument"); return RT_ERR_INVALID; }
    *out = nullptr;
    rt_scene* s = new rt_scene();
    try {
        rtb::load_scene_json(scene_path, texture_dir ? texture_dir : "", s->host);
    } catch (const std::exception& e) {
        rtb::set_error(std::string("rt_scene_load_json: ") + e.what());
        delete s;
        return RT_ERR_IO;
    }
    *out = s;
    return RT_OK;
}

int rt_scene_create(const rt_scene_desc* desc, rt_scene** out) {
    if (!desc || !out) { rtb::set_error("rt_scene_create: null argument"); return RT_ERR_INVALID; }
    *out = nullptr;
    if (desc->n_shapes < 0 || desc->n_lights < 0 || desc->n_materials < 0 || desc->n_textures < 0 ||
        (desc->n_shapes > 0 && !desc->shapes) || (desc->n_lights > 0 && !desc->lights) ||
        (desc->n_materials > 0 && !desc->materials) || (desc->n_textures > 0 && !desc->textures)) {
        rtb::set_error("rt_scene_create: inconsistent counts / null arrays");
        return RT_ERR_INVALID;
    }
    rt_scene* s = new rt_scene();
    try {
        rtb::create_scene_from_desc(*desc, s->host);
    } catch (const std::exception& e) {
        rtb::set_error(std::string("rt_scene_create: ") + e.what());
        delete s;
        return RT_ERR_INVALID;
    }
    *out = s;
    return RT_OK;
}

void rt_scene_destroy(rt_scene* scene) {
    if (!scene) return;
    rtb::device_release(scene->host);
    delete scene;
}

int rt_scene_resolution(const rt_scene* scene, int32_t* width, int32_t* height) {
    if (!scene) { rtb::set_error("null scene"); return RT_ERR_INVALID; }
    if (width) *width = scene->host.cam.res_x;
    if (height) *height = scene->host.cam.res_y;
    return RT_OK;
}

int rt_scene_counts(const rt_scene* scene, int32_t* n_shapes, int32_t* n_lights, int32_t* n_materials,
                    int32_t* n_nodes, int32_t* n_leaves) {
    if (!scene) { rtb::set_error("null scene"); return RT_ERR_INVALID; }
    const HostScene& h = scene->host;
    if (n_shapes) *n_shapes = (int32_t)h.prims.size();
    if (n_lights) *n_lights = (int32_t)h.lights.size();
    if (n_materials) *n_materials = (int32_t)h.materials.size();
    if (n_nodes) *n_nodes = (int32_t)h.tree.size();
    if (n_leaves) *n_leaves = h.n_leaves;
    return RT_OK;
}

int rt_scene_shape_order(const rt_scene* scene, int32_t* out, int32_t n) {
    if (!scene || !out) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    const HostScene& h = scene->host;
    if (n != (int32_t)h.order.size()) { rtb::set_error("rt_scene_shape_order: n must equal the shape count"); return RT_ERR_INVALID; }
    for (int32_t i = 0; i < n; ++i) out[i] = h.order[i];
    return RT_OK;
}

int rt_scene_dump_bvh(const rt_scene* scene, rt_bvh_node_dump* out, int32_t max_nodes) {
    if (!scene || (!out && max_nodes > 0)) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    const HostScene& h = scene->host;
    int32_t n = 0;
    for (const rtb::TreeNode& t : h.tree) {
        if (n >= max_nodes) break;
        rt_bvh_node_dump& d = out[n++];
        std::memset(&d, 0, sizeof(d));
        d.is_leaf = t.left < 0 ? 1 : 0;
        for (int i = 0; i < 3; ++i) { d.box_min[i] = t.box.lo[i]; d.box_max[i] = t.box.hi[i]; }
        if (d.is_leaf) {
            d.count = t.count;
            for (int k = 0; k < t.count && k < 4; ++k) d.prims[k] = h.order[t.first + k];
        }
    }
    return n;
}

int rt_json_number(const char* text, double* value, int32_t* is_integer) {
    if (!text || !value) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    try {
        const jsonmin::Value v = jsonmin::parse(std::string(text));
        if (!v.is_number()) { rtb::set_error("not a JSON number"); return RT_ERR_INVALID; }
        *value = v.kind == jsonmin::Value::Int ? (double)v.i : v.d;
        if (is_integer) *is_integer = v.kind == jsonmin::Value::Int ? 1 : 0;
        return RT_OK;
    } catch (const std::exception& e) {
        rtb::set_error(e.what());
        return RT_ERR_INVALID;
    }
}

int rt_scene_dump_wide(const rt_scene* scene, float* out, int32_t max_nodes, int32_t* depth) {
    if (!scene || (!out && max_nodes > 0)) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    const HostScene& h = scene->host;
    const int32_t n = (int32_t)std::min<size_t>(h.dwide.size(), (size_t)std::max(0, max_nodes));
    if (n > 0) std::memcpy(out, h.dwide.data(), (size_t)n * sizeof(rtb::DWide));
    if (depth) *depth = h.wide_depth;
    return (int32_t)h.dwide.size();
}

int rt_write_ppm(const char* path, int32_t width, int32_t height, const uint8_t* rgb8) {
    if (!path || !rgb8 || width <= 0 || height <= 0) { rtb::set_error("rt_write_ppm: bad argument"); return RT_ERR_INVALID; }
    if (!rtb::write_ppm_p3(path, width, height, rgb8)) { rtb::set_error(std::string("rt_write_ppm: cannot write ") + path); return RT_ERR_IO; }
    return RT_OK;
}

int rt_read_ppm(const char* path, int32_t* width, int32_t* height, uint8_t** rgb8) {
    if (!path || !width || !height || !rgb8) { rtb::set_error("rt_read_ppm: null argument"); return RT_ERR_INVALID; }
    rtb::Texture t;
    if (!rtb::read_ppm_p3(path, t)) { rtb::set_error(std::string("rt_read_ppm: cannot read ") + path); return RT_ERR_IO; }
    *width = t.width;
    *height = t.height;
    *rgb8 = (uint8_t*)std::malloc(t.rgb.size());
    if (!*rgb8) { rtb::set_error("rt_read_ppm: out of memory"); return RT_ERR_IO; }
    std::memcpy(*rgb8, t.rgb.data(), t.rgb.size());
    return RT_OK;
}

void rt_free(void* p) { std::free(p); }

}  // extern "C"
