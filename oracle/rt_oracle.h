/* rt_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product (ray_tracying_b200/) never does.
 *
 * The input structs are declared here independently of include/rt_render.h; they describe the
 * same constructor-level quantities (the reference's Camera / Light / Material / Shapes
 * constructor arguments), so a test can hand the same numpy arrays to both sides.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_camera {
    float location[3], gaze[3], up[3];
    float focal_length;
    int32_t sensor_width, sensor_height;
    float aperture, focus_dist;
    int32_t res_x, res_y;
} orc_camera;

typedef struct orc_light { float location[3], color[3], intensity, radius; } orc_light;

typedef struct orc_material {
    float diffuse[3], specular[3];
    float k_ambient, k_diffuse, k_specular, shininess, roughness, reflectivity, transparency, refractive_index;
    int32_t texture;
} orc_material;

typedef struct orc_shape {
    int32_t type; /* 0 sphere, 1 cube, 2 rectangle, 3 plane */
    int32_t material;
    float translation[3], rotation[3], scale[3], velocity[3];
    float corners[12];
} orc_shape;

typedef struct orc_texture { int32_t width, height; const uint8_t* rgb; } orc_texture;

typedef struct orc_scene_desc {
    orc_camera camera;
    int32_t n_lights; const orc_light* lights;
    int32_t n_materials; const orc_material* materials;
    int32_t n_shapes; const orc_shape* shapes;
    int32_t n_textures; const orc_texture* textures;
} orc_scene_desc;

typedef struct orc_params {
    int32_t use_bvh, samples_sqrt, light_samples, max_depth;
    uint64_t seed;
    float fixed_time;   /* >= 0: fixed shutter time; < 0: random per sample */
    int32_t row0, row1; /* rows [row0,row1) are rendered; row1 <= 0 means all */
    int32_t threads;    /* worker threads over the window's pixels (results do not depend on it) */
    int32_t col0, col1; /* columns [col0,col1) of those rows are rendered; col1 <= 0 means all */
} orc_params;

typedef struct orc_scene orc_scene;

int orc_scene_create(const orc_scene_desc* desc, orc_scene** out);
void orc_scene_destroy(orc_scene* s);
/* shape_list after BVH construction (load-order indices), n = n_shapes */
int orc_scene_shape_order(const orc_scene* s, int32_t* out, int32_t n);
/* pre-order node dump: kind (1 = leaf), 6 box floats, count, up to 4 load-order indices */
typedef struct orc_node_dump { int32_t is_leaf; float lo[3], hi[3]; int32_t count; int32_t prims[4]; } orc_node_dump;
int orc_scene_dump_bvh(const orc_scene* s, orc_node_dump* out, int32_t max_nodes);

/* Renders rows [row0,row1) x columns [col0,col1). Buffers are FULL-frame sized (res_y*res_x); only the
 * rendered pixels are written. rays[3] (optional) receives primary / shadow / secondary ray counts. Any output may be NULL. */
int orc_render(const orc_scene* s, const orc_params* p, uint8_t* rgb8, int32_t* hit_ids, float* hit_t, float* linear,
               uint64_t* rays);

/* Philox4x32-10 block, for the known-answer test. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
