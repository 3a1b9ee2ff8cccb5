/* rt_render.h -- C ABI of the B200-native renderer (librt_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path: the frame loop in
 * Code/raytracer.cpp:433-476 (compute_pixel_color -> Trace -> shade -> BVH::get_intersection ->
 * Shapes::intersect) and the data that loop consumes. Plain pointers and sizes only; no C++ or
 * torch types. Every entry point names the reference interface it replaces.
 *
 * All functions return RT_OK (0) or a negative rt_status; rt_last_error() gives the message for
 * the calling thread. There is no CPU fallback: without a CUDA device every render call fails
 * with RT_ERR_CUDA.
 */
#ifndef RT_RENDER_H
#define RT_RENDER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID = -1, /* bad argument */
    RT_ERR_IO = -2,      /* file could not be opened / parsed */
    RT_ERR_CUDA = -3,    /* CUDA runtime error or no device */
    RT_ERR_SCENE = -4    /* scene is unusable (e.g. zero resolution, raytracer.cpp:403) */
} rt_status;

/* Primitive kinds = the reference's Shapes subclasses (Code/shapes.hpp:90-139). */
typedef enum rt_shape_type {
    RT_SPHERE = 0,    /* Sphere: unit sphere under T*R*S, with linear velocity (shapes.cpp:193) */
    RT_CUBE = 1,      /* Cube: unit cube [-0.5,0.5]^3 under T*R*S (shapes.cpp:349) */
    RT_RECTANGLE = 2, /* Rectangle: unit square on z=0 under T*R*S (shapes.cpp:293) */
    RT_PLANE = 3      /* Plane: world-space quad given by 4 corners (shapes.cpp:438) */
} rt_shape_type;

/* Camera members as Camera::readCameraSpec leaves them (camera.cpp:14-58). sensor_* are ints
 * because the reference reads them with get<int>() (camera.cpp:39-40). */
typedef struct rt_camera_desc {
    float location[3];
    float gaze[3];
    float up[3];
    float focal_length;
    int32_t sensor_width, sensor_height;
    float aperture;   /* 0 = pinhole (camera.cpp:138) */
    float focus_dist; /* default 10 */
    int32_t res_x, res_y;
} rt_camera_desc;

/* Light (light.hpp:5-13). */
typedef struct rt_light_desc {
    float location[3];
    float color[3];
    float intensity;
    float radius; /* > 0 -> soft shadows with light_samples jittered targets (raytracer.cpp:207) */
} rt_light_desc;

/* Material as parse_material returns it (json_loader.cpp:30-97; material.hpp:47-76). */
typedef struct rt_material_desc {
    float diffuse_color[3];
    float specular_color[3];
    float k_ambient, k_diffuse, k_specular;
    float shininess; /* 5 / clamp(roughness,0.001,1)^2 (json_loader.cpp:56-61) */
    float roughness; /* glossy-reflection fuzz radius (raytracer.cpp:312) */
    float reflectivity, transparency, refractive_index;
    int32_t texture; /* index into textures, -1 = none */
} rt_material_desc;

/* Constructor arguments of Sphere/Cube/Rectangle/Plane (shapes.hpp:92-139). For spheres,
 * `velocity` is the constructor argument, i.e. the JSON value already divided by 5
 * (json_loader.cpp:221-223). `corners` is used by RT_PLANE only. */
typedef struct rt_shape_desc {
    int32_t type;     /* rt_shape_type */
    int32_t material; /* index into materials */
    float translation[3];
    float rotation[3];
    float scale[3];
    float velocity[3];
    float corners[12];
} rt_shape_desc;

/* 8-bit RGB texture, row-major from the top-left, as Image::read leaves it (image.cpp:86-133). */
typedef struct rt_texture_desc {
    int32_t width, height;
    const uint8_t* rgb;
} rt_texture_desc;

typedef struct rt_scene_desc {
    rt_camera_desc camera;
    int32_t n_lights;
    const rt_light_desc* lights;
    int32_t n_materials;
    const rt_material_desc* materials;
    int32_t n_shapes;
    const rt_shape_desc* shapes; /* in the loader's push order: spheres, cubes, rectangles, planes */
    int32_t n_textures;
    const rt_texture_desc* textures;
} rt_scene_desc;

/* Command-line switches of the reference binary (raytracer.cpp:360-389) plus what a multi-GPU
 * caller needs. Zero-initialise, then rt_render_params_default(). */
typedef struct rt_render_params {
    int32_t use_bvh;       /* -bvh: 1 = tree (acceleration.cpp:67-117), 0 = linear scan (:123-138) */
    int32_t samples_sqrt;  /* -s N : N x N stratified samples per pixel; <=1 = one centre ray */
    int32_t light_samples; /* -light_sample N */
    int32_t max_depth;     /* recursion limit, reference constant MAX_RECURSION_DEPTH = 10 */
    uint64_t seed;         /* Philox key; the reference seeds mt19937 from random_device */
    float fixed_time;      /* >= 0: every ray gets this shutter time; < 0: uniform [0,1) per sample */
    int32_t rank, world;   /* this process renders screen tiles t with t % world == rank */
    int32_t tile_w, tile_h; /* screen tile size used for that interleave (multiples of 8 and 4) */
    int32_t collect_stats; /* 1: also count node visits / primitive tests (slower) */
    int32_t reserved[7];   /* [0] bit 0: literal reference traversal (visit everything, exact tests only; validation)
                              [1] bit 0: record CUDA events around the trace/shadow/shade/light launches
                                         (rt_scene_last_kernel_times)
                              [2] bit 0: serialise the launches of a frame on the caller's stream (by default the
                                         shadow/light kernels of a level overlap the next level on a second stream)
                              [3], [4] : render window, x0 | x1 << 16 and y0 | y1 << 16 (both 0 = the whole frame):
                                         only pixels with x0 <= x < x1, y0 <= y < y1 are rendered -- the frame loop
                                         of raytracer.cpp:433-476 restricted to a region (band-sampled validation
                                         against the CPU reference at full benchmark sizes)
                              [5]      : tile block B (0 or 1 = none): tiles are dealt to the ranks in B x B groups
                                         instead of singly, so that a rank touches a compact part of the scene per group
                              [6]      : 0 */
} rt_render_params;

typedef struct rt_render_stats {
    uint64_t rays;          /* get_intersection calls: primary + shadow + reflection + refraction */
    uint64_t primary_rays;
    uint64_t shadow_rays;
    uint64_t secondary_rays;
    uint64_t node_visits;   /* only with collect_stats */
    uint64_t prim_tests;    /* only with collect_stats */
    float kernel_ms;        /* CUDA-event time of the render kernels of this call */
    float total_ms;         /* CUDA-event time of the whole call on the stream (incl. copies) */
    int32_t launches;       /* kernels launched by this call */
    int32_t pixels;         /* pixels this rank rendered */
} rt_render_stats;

typedef struct rt_scene rt_scene; /* opaque: host scene + flattened BVH + device copies */

const char* rt_last_error(void);
int rt_version(void);

/* Number of CUDA devices visible (0 on a CPU-only box; never an error). */
int rt_device_count(void);

void rt_render_params_default(rt_render_params* p);

/* Replaces Camera::Camera + load_lights_from_json + load_shapes_from_json (camera.cpp:239,
 * json_loader.cpp:103,164) and BVH::BVH (acceleration.cpp:7): parses the scene.json schema,
 * resolves "texture_file": "x.jpg" to <texture_dir>/x.ppm like json_loader.cpp:78-80 (texture_dir
 * NULL = "../../Textures"), builds the reference's tree and flattens it. Host only: no GPU needed. */
int rt_scene_load_json(const char* scene_path, const char* texture_dir, rt_scene** out);

/* Same, from constructor-level arrays (synthetic scenes without a JSON round trip). */
int rt_scene_create(const rt_scene_desc* desc, rt_scene** out);

void rt_scene_destroy(rt_scene* scene);

/* Introspection used by the host logic tests (all host side). */
int rt_scene_resolution(const rt_scene* scene, int32_t* width, int32_t* height);
int rt_scene_counts(const rt_scene* scene, int32_t* n_shapes, int32_t* n_lights, int32_t* n_materials,
                    int32_t* n_nodes, int32_t* n_leaves);
/* BVH leaf order: out[i] = load-order index of the primitive at sorted position i, i.e. the
 * reference's shape_list after BVH construction (acceleration.cpp:20-64). n = n_shapes. */
int rt_scene_shape_order(const rt_scene* scene, int32_t* out, int32_t n);
/* Pre-order dump of the tree, comparable with the reference's node structure: for each node
 * kind (0 internal, 1 leaf), box[6], and for leaves count + up to 4 load-order indices.
 * Returns the number of nodes written (<= max_nodes) or a negative status. */
typedef struct rt_bvh_node_dump {
    int32_t is_leaf;
    float box_min[3], box_max[3];
    int32_t count;
    int32_t prims[4];
} rt_bvh_node_dump;
int rt_scene_dump_bvh(const rt_scene* scene, rt_bvh_node_dump* out, int32_t max_nodes);

/* How the scene loader reads one JSON number (nlohmann semantics: int64 for plain integers, otherwise
 * the correctly rounded IEEE double of strtod) -- exposed so that the tests can hold the loader's
 * fast path to strtod bit for bit. */
int rt_json_number(const char* text, double* value, int32_t* is_integer);

/* The flattened 4-wide device tree (host copy), 32 floats per node, layout in
 * ray_tracying_b200/csrc/scene.hpp (DWide): child boxes as structure of arrays, first child, meta
 * word (valid / gate masks, leaf flag, primitive types), sphere cull coefficient. Writes at most
 * max_nodes nodes, returns the number of nodes of the tree; *depth = its number of levels. */
int rt_scene_dump_wide(const rt_scene* scene, float* out, int32_t max_nodes, int32_t* depth);

/* Copies primitives, nodes, materials, lights and textures to the current CUDA device (the
 * "scene resident in HBM" state). Idempotent; rt_render* call it on demand. `bytes` (optional)
 * receives the number of bytes copied host->device. The copy is asynchronous on the default stream;
 * every later frame, on whatever stream, waits for it through an event, and a re-upload after
 * rt_scene_evict waits for the last frame that still reads the old copy. There is ONE frame in flight
 * per scene and device: frames enqueued on different streams are ordered one behind the other. */
int rt_scene_upload(rt_scene* scene, uint64_t* bytes);
/* Marks the device copy stale so that the next upload / render copies the scene host->device
 * again (end-to-end timing). Device and page-locked host allocations are kept until
 * rt_scene_destroy: an upload is one asynchronous copy of one arena. */
int rt_scene_evict(rt_scene* scene);

/* CUDA-event timings of the most recent rt_render_device / rt_render call on this scene: the
 * render kernels alone, and everything the call put on the stream. The events are recorded on
 * the launching stream by every call; this getter waits for them, so it can be used after an
 * asynchronous rt_render_device(stats = NULL) without perturbing the timed region. Returns
 * RT_ERR_SCENE if that (asynchronous) frame overflowed a ray queue and therefore dropped rays.
 * The overflow of an asynchronous frame is STICKY: the next rt_render_device / rt_render /
 * rt_scene_last_timing call on the scene that finds the flag halves the batch size and returns
 * RT_ERR_SCENE once; the caller renders the frame again. */
int rt_scene_last_timing(rt_scene* scene, float* kernel_ms, float* total_ms);

/* Per-kernel-class CUDA-event times of the most recent frame rendered with reserved[1] bit 0 set:
 * ms4 / launches4 = {trace_kernel, shadow_kernel, shade_kernel, light_kernel} (sum of the launch
 * durations, number of TIMED launches -- at most 512 per frame); frame_launches4 = all launches of
 * the frame per class. Events are recorded on the launching stream; waits for them. */
int rt_scene_last_kernel_times(rt_scene* scene, float* ms4, int32_t* launches4, int32_t* frame_launches4);

/* Number of pixels this rank renders for (rank, world, tile). */
int rt_shard_pixels(const rt_scene* scene, const rt_render_params* p, int64_t* n_pixels);

/* Replaces the frame loop raytracer.cpp:433-476 for this rank's tiles. Device-resident outputs,
 * launched on `stream` (a cudaStream_t passed as void*; NULL = default stream), asynchronous
 * unless stats != NULL (then it synchronises the stream to read counters and timings).
 *   rgb8    : width*height*3 bytes, gamma 1.1 + clamp + *255.999 (raytracer.cpp:446-457); may be NULL
 *   hit_ids : width*height int32, load-order index of the primitive hit by the first sample's
 *             primary ray, -1 = background; may be NULL
 *   linear  : width*height*3 float, averaged linear colour before gamma; may be NULL
 * Full-frame buffers; a rank with world > 1 writes only the pixels of its own tiles.
 * Ray queues have a fixed capacity; a scene whose ray trees branch heavily (reflective AND
 * transparent materials) can overflow them, which a synchronous call (stats != NULL) detects and
 * answers by re-rendering with smaller batches, remembered for later calls on this scene; an
 * asynchronous call reports it through the next call (see rt_scene_last_timing). Scenes without a
 * material that is both reflective and transparent cannot overflow (a level never has more rays than
 * the batch). */
int rt_render_device(rt_scene* scene, const rt_render_params* p, uint8_t* rgb8, int32_t* hit_ids, float* linear,
                     void* stream, rt_render_stats* stats);

/* Same, with HOST output buffers: (upload scene if needed) -> render -> copy back. This is the
 * end-to-end call; total_ms covers all of it. */
int rt_render(rt_scene* scene, const rt_render_params* p, uint8_t* rgb8, int32_t* hit_ids, float* linear,
              rt_render_stats* stats);

/* The same end-to-end call over SEVERAL GPUs of this process (the north-star's "image sharded across the
 * GPUs of one box by interleaved screen tiles, scene and BVH replicated, tiles copied to host at frame
 * end"): replaces the serial frame loop raytracer.cpp:433-476 for the whole frame. One persistent host
 * thread per device: scene upload (all devices copy from one page-locked staging buffer, in parallel),
 * render of the device's tiles (p->tile_w x p->tile_h, dealt round-robin, see reserved[5]) into a packed
 * buffer, ONE device->host copy per output into page-locked memory, scatter into the caller's HOST frame
 * buffers. No inter-GPU traffic, no NCCL. devices = NULL means ordinals 0..n_devices-1. p->rank / p->world
 * must be 0 / 1. stats: ray counts summed over the devices, kernel_ms = slowest device, total_ms = host
 * wall clock of the whole call. The result is bit-identical to rt_render on one device. */
int rt_render_multi(rt_scene* scene, const rt_render_params* p, int32_t n_devices, const int32_t* devices,
                    uint8_t* rgb8, int32_t* hit_ids, float* linear, rt_render_stats* stats);

/* Kernels launched for this scene so far (all devices): rt_render* calls add what they enqueue. */
int rt_scene_launch_count(rt_scene* scene, uint64_t* launches);

/* Measured ceiling of the traversal loop on the current device (the denominator of bench.py's
 * roofline): the loop's own node step -- four conservative slab tests, the sorting network, push / pop
 * on the shared-memory stack -- run by fully converged warps over n_nodes SYNTHETIC nodes (L1-resident)
 * whose four children all contain the scene, so that every step does its full work (four passes, three
 * pushes, one descent); `steps` node visits per thread, best of `repeats` launches. any_hit selects the
 * occlusion-query flavour (shadow kernels) or the closest-hit one. Returns box tests per second: what
 * the traversal kernels would reach with no divergence, no cache misses, no fetch / primitive phases. */
int rt_traversal_peak(rt_scene* scene, int32_t any_hit, int32_t n_nodes, int32_t steps, int32_t repeats,
                      double* box_tests_per_s, float* ms);

/* Self-tests of the exactness machinery on the device (tests/test_gpu_properties.py).
 * rt_selftest_boxes: n random (ray, box) pairs -- generic, near-axis-parallel directions, huge
 *   coordinates, grazing faces / edges / corners to a few ulps -- through the conservative slab test
 *   and the reference's exact AABB::intersect (shapes.cpp:55-72). out8 = {tests, exact passes,
 *   conservative passes, surely passes, VIOLATIONS exact && !conservative, VIOLATIONS surely &&
 *   !exact, rays skipped because a direction component is <= 1e-6, 0}.
 * rt_selftest_cull: for every primitive of the scene, rays_per_primitive rays aimed at and around it
 *   from near and far: the exact intersection routine against the conservative test of the
 *   primitive's culling box. out8 = {tests, exact hits, culling passes, VIOLATIONS hit && !pass, ...}. */
int rt_selftest_boxes(uint64_t seed, int64_t n, uint64_t* out8);
int rt_selftest_cull(rt_scene* scene, uint64_t seed, int32_t rays_per_primitive, uint64_t* out8);

/* Replaces Image::write / Image::read (image.cpp:53-84, 86-133): ASCII P3 PPM. */
int rt_write_ppm(const char* path, int32_t width, int32_t height, const uint8_t* rgb8);
int rt_read_ppm(const char* path, int32_t* width, int32_t* height, uint8_t** rgb8 /* free with rt_free */);
void rt_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* RT_RENDER_H */
