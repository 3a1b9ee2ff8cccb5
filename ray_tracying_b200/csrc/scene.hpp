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

// One WIDE node = 128 bytes = one L2 line: up to 4 children, structure-of-arrays so that four
// 256-bit loads fetch it (rt_device.cuh: ldg256):
//   f[ 0.. 3] = lo.x of children 0..3    f[ 4.. 7] = hi.x
//   f[ 8..11] = lo.y                     f[12..15] = hi.y
//   f[16..19] = lo.z                     f[20..23] = hi.z
//   f[24] = bits: node index of child 0 -- the INNER children of a node occupy slots 0..ni-1 and are
//           consecutive nodes (first + slot); primitive children follow in slots ni..n-1
//   f[25] = bits: meta = valid mask (bits 0-3) | primitive mask (bits 4-7: the child is a primitive) |
//           primitive types, 2 bits per child (bits 16-23)
//   f[26] = Q: quadratic cull coefficient (largest over the spheres BELOW this node, else 0)
//   f[27..30] = bits: sorted position of the primitive in slot 0..3 (primitive children only)
//   f[31] unused
// The tree is a 4-wide BVH over the primitives' CULLING boxes (bvh.cpp: SAH build, cull_pad); a box
// of this tree is only ever used to skip work. The reference's own condition for testing a shape --
// the exact box of its reference leaf passes AABB::intersect -- is evaluated per primitive on
// HostScene::dleafbox (the "gate"). The reference tree itself stays on the host (sort order for
// tie-breaks, leaf boxes, tests) and is uploaded only for the literal validation mode.
struct alignas(128) DWide { float f[32]; };
constexpr uint32_t WIDE_PRIM_SHIFT = 4;  // meta >> 4 & 15 = primitive mask

// Material = 64 bytes:
//   m0 = (diffuse.rgb, k_ambient) m1 = (specular.rgb, k_diffuse)
//   m2 = (k_specular, shininess, roughness, reflectivity)
//   m3 = (transparency, refractive_index, bits texture index, 0)
struct alignas(16) DMaterial { F4 m[4]; };

// Light = 32 bytes: l0 = (location, intensity) l1 = (color, radius)
struct alignas(16) DLight { F4 l[2]; };

struct DTexture { int32_t width, height; uint32_t offset, pad; };  // offset into the u8 pool

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

struct DeviceSet;  // defined in render.cu: the scene's copies on the CUDA devices

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
    std::vector<DWide> dwide;     // dwide[0] = root (empty when there are no shapes)
    std::vector<F4> dleafbox;     // per sorted primitive: (lo.xyz, 0), (hi.xyz, 0) of the EXACT box of its reference leaf
    int wide_depth = 0;           // levels of the wide tree
    int stack_need = 1;           // traversal stack entries a ray can need (max over root->leaf paths of the pushed siblings) + 1
    std::vector<DMaterial> dmaterials;
    std::vector<DLight> dlights;
    std::vector<DTexture> dtextures;
    std::vector<uint8_t> texels;

    DeviceSet* dev = nullptr;

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
