// scene.hpp -- host-side scene model and the device (HBM) record layouts.
//
// Host code mirrors the reference's loader and BVH builder so that the device structures hold
// exactly the values the reference computes (same float operations in the same order; this
// translation unit is compiled with -ffp-contract=off, the CUDA side with -fmad=false).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rt_render.h"

namespace rtb {

// ---------------------------------------------------------------------------------------------
// Device record layouts (plain floats so that host code needs no CUDA headers).
// ---------------------------------------------------------------------------------------------
struct alignas(16) F4 { float x, y, z, w; };

// One primitive = 128 bytes = one L2 line, 8 x float4:
//   q0 = (velocity.xyz, bits: type | material << 2)
//   transformed shapes (sphere, cube, rectangle):
//     q1..q3 = rows 0..2 of world_to_object, q4..q6 = rows 0..2 of object_to_world
//   plane:
//     q1 = (c0.xyz, c3.x) q2 = (c1.xyz, c3.y) q3 = (c2.xyz, c3.z) q4 = (unit normal, valid ? 1 : 0)
//   q7 = (bits: load-order index, 0, 0, 0)
// The first 64 bytes are what a miss needs; the second 64 only matter for a hit.
struct alignas(128) DPrim { F4 q[8]; };

// One internal node = 64 bytes: both children's boxes and references.
//   a = (L.min.x, L.min.y, L.min.z, L.max.x)  b = (L.max.y, L.max.z, R.min.x, R.min.y)
//   c = (R.min.z, R.max.x, R.max.y, R.max.z)  d = bits(left_ref, right_ref, 0, 0)
// ref >= 0: index of an internal node; ref < 0: leaf, ~ref = index into the leaf records.
struct alignas(64) DNode { F4 a, b, c, d; };

// One leaf of the reference tree (<= 4 primitives) = 128 bytes:
//   l0 = (box.lo.xyz, bits first primitive)
//   l1 = (box.hi.xyz, bits meta): meta bits 0-2 = count, bits 4+4T..7+4T = mask of the leaf's
//        primitives that have type T (so a warp can run one type's test routine at a time)
//   l2..l7 = 24 floats: for primitive k, floats [6k, 6k+6) = its own box (lo.xyz, hi.xyz),
//            INFLATED outward, used only to skip primitives the ray clearly misses.
// The leaf box is the reference's (exact): whether it passes AABB::intersect decides whether the
// leaf's primitives are tested at all.
struct alignas(128) DLeaf { F4 l[8]; };

// Material = 64 bytes:
//   m0 = (diffuse.rgb, k_ambient) m1 = (specular.rgb, k_diffuse)
//   m2 = (k_specular, shininess, roughness, reflectivity)
//   m3 = (transparency, refractive_index, bits texture index, 0)
struct alignas(16) DMaterial { F4 m[4]; };

// Light = 32 bytes: l0 = (location, intensity) l1 = (color, radius)
struct alignas(16) DLight { F4 l[2]; };

struct DTexture { int32_t width, height; uint32_t offset, pad; };  // offset into the u8 pool

inline int32_t leaf_ref(int leaf_index) { return ~leaf_index; }

// ---------------------------------------------------------------------------------------------
// Host model
// ---------------------------------------------------------------------------------------------
struct Box {
    float lo[3], hi[3];
};

struct HostPrim {
    int type = 0;
    int material = 0;
    float velocity[3] = {0, 0, 0};
    float w2o[4][4];
    float o2w[4][4];
    float corners[4][3];
    float normal[3] = {0, 0, 0};  // planes
    bool normal_valid = false;
    Box box;
};

struct Texture {
    int width = 0, height = 0;
    std::vector<uint8_t> rgb;
};

struct TreeNode {
    Box box;
    int left = -1, right = -1;  // indices into HostScene::tree, -1 for leaves
    int first = 0, count = 0;   // range over the sorted order
};

struct DeviceScene;  // defined in render.cu

struct HostScene {
    rt_camera_desc cam{};
    float xdir[3], ydir[3], zdir[3];  // camera basis, camera.cpp:109-115
    std::vector<rt_light_desc> lights;
    std::vector<rt_material_desc> materials;
    std::vector<Texture> textures;
    std::vector<HostPrim> prims;  // load order

    // BVH (reference construction)
    std::vector<int> order;       // sorted position -> load-order index (the reference's shape_list)
    std::vector<TreeNode> tree;   // pre-order, tree[0] = root (empty when there are no shapes)
    int n_leaves = 0;

    // flattened for the device
    std::vector<DPrim> dprims;    // sorted order
    std::vector<DNode> dnodes;
    std::vector<DLeaf> dleaves;
    int32_t root_ref = 0;
    Box root_box{};
    std::vector<DMaterial> dmaterials;
    std::vector<DLight> dlights;
    std::vector<DTexture> dtextures;
    std::vector<uint8_t> texels;

    DeviceScene* dev = nullptr;

    double build_seconds = 0.0;
};

// scene.cpp
void finalize_scene(HostScene& s);  // camera basis, matrices already set; builds BVH + flattens
void load_scene_json(const std::string& path, const std::string& texture_dir, HostScene& s);
void create_scene_from_desc(const rt_scene_desc& d, HostScene& s);
bool read_ppm_p3(const std::string& path, Texture& out);
bool write_ppm_p3(const std::string& path, int width, int height, const uint8_t* rgb);

// bvh.cpp
void build_bvh(HostScene& s);
void flatten_scene(HostScene& s);

}  // namespace rtb
