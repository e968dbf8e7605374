/*
 * ezrt.h -- C ABI of ezrt_b200: the B200-native drop-in for EzRT's path-tracing hot path.
 *
 * The reference (AKGWSB/EzRT) has no plugin/FFI interface; its de-facto boundary is "what
 * main() hands to pass1 and what pass1 leaves in lastFrame" (SURVEY.md 8b).  Every entry
 * point below cites the reference code it replaces.  Path aliases: P2/ P3/ P4/ P5/ = the
 * "source code" directory of tutorial part 2..5; fsh = shaders/fshader.fsh.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; all functions return 0 on success or
 *     a negative ezrt_status; ezrt_last_error() gives a thread-local message.  The
 *     reference prints and exit(-1)s (P5/main.cpp:178-182, :283-286); this ABI never exits.
 *   - the caller owns every input array and every output buffer; a scene handle owns its
 *     device allocations.  One handle = one CUDA device; calls on one handle are
 *     serialised by the caller; different handles may be driven from different threads.
 *   - "tris" is an array of Triangle_encoded (P5/main.cpp:60-69): 36 packed floats
 *     (p1 p2 p3 n1 n2 n3 emissive baseColor param1..param4), stride 144 B.
 *   - "nodes" is an array of BVHNode_encoded (P5/main.cpp:71-76): 12 packed floats
 *     (left,right,0)(n,index,0) AA BB, ints stored as floats; element 0 is the dummy
 *     testNode, the root is element 1 (P5/main.cpp:830-838, P5/fsh:263).
 *   - framebuffers are linear-radiance fp32, row 0 = bottom row (GL convention), the
 *     content of pass1's colour attachment 0 / lastFrame (P5/fsh:942-947) before any
 *     tone mapping.
 */
#ifndef EZRT_H
#define EZRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EZRT_TRIANGLE_FLOATS 36 /* Triangle_encoded, P5/main.cpp:60-69 */
#define EZRT_BVHNODE_FLOATS 12  /* BVHNode_encoded,  P5/main.cpp:71-76 */
#define EZRT_MATERIAL_FLOATS 18 /* emissive, baseColor, 12 scalars: P5/main.cpp:27-42 */

typedef enum ezrt_status {
    EZRT_OK = 0,
    EZRT_ERR_INVALID = -1,   /* bad argument */
    EZRT_ERR_IO = -2,        /* file could not be opened / parsed */
    EZRT_ERR_CUDA = -3,      /* CUDA runtime error or no device */
    EZRT_ERR_BAD_TREE = -4,  /* node array is not a tree the traversal can walk */
    EZRT_ERR_NOMEM = -5
} ezrt_status;

/* Integrator modes = the four per-pixel loops the reference ships. */
typedef enum ezrt_mode {
    EZRT_MODE_DIFFUSE_P3 = 0,      /* P3/fsh:376-446  diffuse, uniform hemisphere, wang-hash   */
    EZRT_MODE_DISNEY_ANISO_P4 = 1, /* P4/fsh:478-550  anisotropic Disney, uniform hemisphere   */
    EZRT_MODE_DISNEY_SOBOL_P5 = 2, /* P5/fsh:762-807  Disney + Sobol/CP rotation               */
    EZRT_MODE_DISNEY_IS_MIS_P5 = 3 /* P5/fsh:810-890  BRDF + HDR importance sampling, MIS      */
} ezrt_mode;

/* BVH traversal policy.  All three return bit-identical hits (tests assert it).
 * ACCEL (default): the device builds its own acceleration tree over the same triangles (the
 *   reference's sweep SAH without its INF = 114514 cost cut-off, P5/main.cpp:20,:493), finds the
 *   global closest hit there, keeps it when the shader's hitBVH provably reaches that triangle's
 *   leaf in the REFERENCE tree and no other triangle ties, and otherwise re-traces the ray with the
 *   exact reference-order traversal (DESIGN.md "accel").
 * REFERENCE walks every box the ray overlaps, exactly as P5/fsh:254-306 (no best-distance pruning).
 * PRUNED walks the reference tree in the shader's order but skips a sub-tree whose box entry lies
 *   beyond the current best hit by a conservative margin (DESIGN.md "pruning"). */
typedef enum ezrt_traverse {
    EZRT_TRAVERSE_ACCEL = 0,
    EZRT_TRAVERSE_REFERENCE = 1,
    EZRT_TRAVERSE_PRUNED = 2
} ezrt_traverse;

typedef enum ezrt_pipeline {
    EZRT_PIPELINE_WAVEFRONT = 0, /* generate / extend / shade / shadow / blend kernels */
    EZRT_PIPELINE_MEGAKERNEL = 1 /* one thread per pixel, literal loop (cross-check)   */
} ezrt_pipeline;

/* Everything display() passes to pass1 per frame (P5/main.cpp:709-745, :919-923) plus
 * the constants that are literals in the shaders. */
typedef struct ezrt_render_params {
    int32_t width, height;   /* uniform width,height  (P5/main.cpp:922-923)                  */
    int32_t spp;             /* number of consecutive display() calls to perform             */
    uint32_t first_frame;    /* frameCounter of the first call (P5/main.cpp:719); 0 = fresh  */
    int32_t max_bounce;      /* literal in main(): P5/fsh:935 (2), P4/fsh:540 (4), P3 (2)    */
    int32_t mode;            /* ezrt_mode                                                    */
    float eye[3];            /* uniform eye           (P5/main.cpp:710-711, :717)            */
    float camera_rotate[16]; /* uniform cameraRotate, column-major inverse(lookAt) (:712-718) */
    float env_color[3];      /* colour of a miss when the scene has no HDR map               */
    int32_t traverse;        /* ezrt_traverse                                                */
    int32_t pipeline;        /* ezrt_pipeline                                                */
    int32_t out_channels;    /* 3 (RGB) or 4 (RGBA, alpha = 1 as in P5/fsh:947)              */
    /* image partition for multi-GPU rendering: the image is cut into 16x16 tiles, tile
     * (tx,ty) belongs to part (tx+ty) % part_count; a part renders only its tiles and
     * writes them compactly (tile-major) unless part_count == 1. */
    int32_t part_rank, part_count;
    int32_t frames_per_batch; /* wavefront: display() calls traced concurrently (0 = auto)   */
    int32_t profile;          /* 1: bracket every kernel with CUDA events (ezrt_get_kernel_times);
                                 2: count the records the accel traversal fetches (ezrt_counters.node_visits / tri_tests;
                                    a slower instantiation of the same kernels -- never inside a timed region) */
    int32_t reserved[3];      /* [0]: flags, EZRT_PARAM_*; [1], [2]: 0 */
} ezrt_render_params;

/* ezrt_render_params.reserved[0]: keep counting -- ezrt_get_counters / ezrt_get_kernel_times then report the sums over all
   renders since the last one issued without this flag (a benchmark loop reads them once, without a sync per render) */
#define EZRT_PARAM_ACCUMULATE 1

typedef struct ezrt_counters {
    uint64_t rays;          /* hitBVH invocations: primary + bounce + shadow     */
    uint64_t primary_rays, bounce_rays, shadow_rays;
    uint64_t samples;       /* pixel-samples = fragment shader invocations                   */
    uint64_t kernel_launches;
    double device_ms;       /* CUDA-event time of the last render on its stream              */
    uint64_t deferred_rays; /* accel policy: rays re-traced by the exact reference-order pass */
    uint64_t node_visits;   /* profile = 2: acceleration-tree node records fetched ...       */
    uint64_t tri_tests;     /*              ... and triangle records (64 B) fetched by the accel kernels */
    uint64_t node_visits_96;/*              of node_visits: 96-byte records (16-bit planes / W8); the rest are 128-byte records */
    uint64_t node_bytes, tri_bytes; /*      bytes of those records                           */
} ezrt_counters;

typedef struct ezrt_scene ezrt_scene; /* device-resident scene (replaces the two TBOs + 2 textures) */

const char* ezrt_last_error(void);
int ezrt_version(void);

/* ----------------------------------------------------------------------------------------
 * Device side: the hot path.
 * -------------------------------------------------------------------------------------- */

/* Replaces the texture-buffer uploads P5/main.cpp:878-906: copies the reference-layout
 * arrays to the GPU `device` and repacks them for the kernels.  hdr / hdr_cache may be
 * NULL (then hdr_w = hdr_h = 0).  hdr rows are stored top-to-bottom as HDRLoader returns
 * them (P5/lib/hdrloader.cpp:76-91); hdr_cache is calculateHdrCache()'s output
 * (P5/main.cpp:592-689).  hdr_filter_linear: 1 = GL_LINEAR (P5/main.cpp:196-199),
 * 0 = GL_NEAREST (P3/main.cpp:195-196). */
int ezrt_scene_create(int device, const float* tris, int n_triangles, const float* nodes, int n_nodes,
                      const float* hdr, const float* hdr_cache, int hdr_w, int hdr_h,
                      int hdr_filter_linear, ezrt_scene** out_scene);
int ezrt_scene_destroy(ezrt_scene* scene);

/* render(width,height,spp) -> framebuffer: equals `spp` consecutive display() calls
 * (P5/main.cpp:697-748) each drawing pass1 (P5/fsh:894-949) and copying to lastFrame.
 * `framebuffer` is a HOST buffer, in/out: when first_frame > 0 it must hold lastFrame.
 * Size: n_pixels(part) * out_channels floats, see ezrt_partition_pixels(). */
int ezrt_render(ezrt_scene* scene, const ezrt_render_params* params, float* framebuffer);

/* Same, but `d_framebuffer` is DEVICE memory on the scene's GPU and the work is enqueued on
 * `cuda_stream` (a cudaStream_t, NULL = default stream).  Steady state: no host synchronisation.  The FIRST
 * call for a given (image size, partition, batch size) allocates the scene's scratch buffers (cudaMalloc synchronises
 * the device) and uploads the tile list (one stream synchronisation).  The scratch belongs to the scene: at most ONE
 * render of a scene may be in flight at a time -- enqueue renders of the same scene on one stream, or order them with
 * events; different scenes (one per GPU) are independent. */
int ezrt_render_device(ezrt_scene* scene, const ezrt_render_params* params, float* d_framebuffer,
                       void* cuda_stream);

/* Counters of the most recent render on this scene (synchronises the stream). */
int ezrt_get_counters(ezrt_scene* scene, ezrt_counters* out);

/* Per-kernel-class device time of the most recent render with params.profile = 1 (CUDA events on
 * the render's stream; synchronises).  Classes: 0 extend (hitBVH, closest hit), 1 shade,
 * 2 shadow (hitBVH, any hit), 3 other (generate, blend, tally).  ms[4], launches[4]. */
int ezrt_get_kernel_times(ezrt_scene* scene, double* ms, uint64_t* launches);

/* Number of pixels part `rank` of `count` owns for a width x height image. */
int64_t ezrt_partition_pixels(int width, int height, int rank, int count);
/* Device kernel: scatter the compact tile-major buffer of part `rank` into a full
 * width x height framebuffer (both device pointers, same channel count).  The device tile lists are cached per
 * (device, image size, rank, count) -- the gather runs once per render -- and live until ezrt_partition_cache_clear(). */
int ezrt_partition_scatter(const float* d_compact, float* d_full, int width, int height, int channels,
                           int rank, int count, void* cuda_stream);
/* Frees the tile lists ezrt_partition_scatter cached on the calling thread's current device (all devices: device < 0). */
int ezrt_partition_cache_clear(int device);
/* Host version of the same scatter (used by the CPU/gloo path and by tests). */
int ezrt_partition_scatter_host(const float* compact, float* full, int width, int height, int channels,
                                int rank, int count);

/* Single-function entry points of the hot path (device), for parity tests: trace `n` rays
 * (origins/dirs: n x 3 floats, host) through hitBVH (P5/fsh:254-306; C++ twin
 * P2/main.cpp:466-485) and return per ray: hit flag, distance, triangle index, isInside,
 * hit point, shading normal. any_hit=1 stops at the first hit (shadow rays, P5/fsh:826-829). */
int ezrt_trace_rays(ezrt_scene* scene, int n, const float* origins, const float* dirs, int traverse,
                    int any_hit, int p3_normal_fudge, int32_t* out_hit, float* out_distance,
                    int32_t* out_triangle, int32_t* out_inside, float* out_point, float* out_normal);

/* Evaluate the BRDF functions on the device for n (V,N,L,material) tuples: which = 0
 * BRDF_Evaluate (P5/fsh:500-549), 1 BRDF_Evaluate aniso (P4/fsh:412-473), 2 BRDF_Pdf
 * (P5/fsh:715-752; result in out[3*i]), 3 SampleBRDF (P5/fsh:633-664; xi = n x 3). */
int ezrt_eval_brdf(int device, int which, int n, const float* V, const float* N, const float* L,
                   const float* xi, const float* materials, float* out);

/* Evaluate a ezrt_math.h function on the device for n inputs (parity of the arithmetic
 * definition): which = 0 sin, 1 cos, 2 log, 3 exp, 4 pow(a,b), 5 atan2(a,b), 6 asin. */
int ezrt_eval_math(int device, int which, int n, const float* a, const float* b, float* out);

/* The binary SAH tree ezrt_scene_create derives its acceleration tree from (NOT the reference's tree: the same
 * exhaustive sweep as buildBVHwithSAH, P5/main.cpp:458-589, without the INF = 114514 cost sentinel, leaves of <= leaf_n
 * triangles), built where = 0 on the GPU (csrc/accel_build.cu, what ezrt_scene_create uses) or where = 1 on the host
 * (host_scene.cpp); both produce the same array.  links_out: 4 ints per node (left, right, n, index; node 0 = root),
 * boxes_out: 6 floats per node (AA, BB), order_out: n_triangles triangle indices (the tree's triangle order); any may be
 * NULL.  nodes_cap = capacity of links_out / boxes_out in nodes.  Returns the node count or a negative status;
 * *ms = wall-clock of the build (device: upload of the triangles and read-back of the tree included). */
int ezrt_accel_build(int device, const float* tris, int n_triangles, int leaf_n, int where, int32_t* links_out,
                     float* boxes_out, int nodes_cap, uint32_t* order_out, double* ms);

/* ----------------------------------------------------------------------------------------
 * Post pass (SURVEY.md 8f "next" row 3): what the user sees.
 * -------------------------------------------------------------------------------------- */

/* pass3 (P5/shaders/pass3.fsh:14-25): toneMapping(c, limit) = c * 1.0 / (1.0 + lum/limit) with
 * lum = 0.3 r + 0.6 g + 0.1 b (limit = 1.5 in the shader), then pow(c, 1/2.2).  Device kernel:
 * d_in has `channels` (3 or 4) floats per pixel, d_out 3 floats per pixel. */
int ezrt_post_tonemap(const float* d_in, int channels, float* d_out, int64_t n_pixels, float limit, void* cuda_stream);
/* Host: write a linear framebuffer (row 0 = bottom, as ezrt_render returns it) as an 8-bit RGB PNG, top row
 * first; tonemap = 1 applies pass3 first; quantisation is P1's imshow(): (unsigned char)clamp(v*255, 0, 255)
 * (P1/main.cpp:173-194; P1 wrote its PNG with the vendored svpng.inc, this is an independent writer). */
int ezrt_write_png(const char* path, const float* framebuffer, int width, int height, int channels, int tonemap);

/* ----------------------------------------------------------------------------------------
 * Host side: the scene pipeline that feeds the path (stays on the CPU, north_star).
 * -------------------------------------------------------------------------------------- */

typedef struct ezrt_trilist ezrt_trilist; /* std::vector<Triangle> of P5/main.cpp:801 */

ezrt_trilist* ezrt_trilist_create(void);
void ezrt_trilist_destroy(ezrt_trilist* list);
int ezrt_trilist_size(const ezrt_trilist* list);

/* getTransformMatrix (P5/main.cpp:255-271): translate * rotate(x,y,z degrees) * scale,
 * column-major out[16]. */
void ezrt_transform_matrix(const float rotate_deg[3], const float translate[3], const float scale[3],
                           float out[16]);
/* readObj (P5/main.cpp:274-392) incl. its unit-box normalisation quirk (:317-318),
 * transform, smooth-normal generation and per-mesh material (18 floats: emissive,
 * baseColor, subsurface, metallic, specular, specularTint, roughness, anisotropic, sheen,
 * sheenTint, clearcoat, clearcoatGloss, IOR, transmission). */
int ezrt_trilist_read_obj(ezrt_trilist* list, const char* path, const float material[EZRT_MATERIAL_FLOATS],
                          const float trans[16], int smooth_normal);
/* smooth_normal is a flag word: bit 0 = smoothNormal; EZRT_OBJ_HARDENED additionally accepts negative
 * (relative) vertex indices and fan-triangulates polygons -- the reference cuts a polygon to its first three
 * vertices (P5/main.cpp:321-336), which stays the default behaviour. */
#define EZRT_OBJ_HARDENED 2
/* Same parser on an in-memory OBJ text (synthetic meshes). */
int ezrt_trilist_read_obj_text(ezrt_trilist* list, const char* text, size_t len,
                               const float material[EZRT_MATERIAL_FLOATS], const float trans[16],
                               int smooth_normal);
/* Append already-built triangles in Triangle_encoded layout. */
int ezrt_trilist_append_encoded(ezrt_trilist* list, const float* tris, int n);

typedef enum ezrt_bvh_builder {
    EZRT_BVH_SAH_FAST = 0,    /* same tree as SAH_LITERAL, keys pre-computed, subtrees in parallel */
    EZRT_BVH_SAH_LITERAL = 1, /* buildBVHwithSAH exactly as written (P5/main.cpp:458-589)   */
    EZRT_BVH_MEDIAN = 2,      /* buildBVH (P5/main.cpp:395-455)                              */
    EZRT_BVH_SAH_NO_SENTINEL = 3 /* NOT the reference tree: same sweep SAH without the INF = 114514
                                    cost cut-off (P5/main.cpp:20,:493); what the device builds
                                    internally as its acceleration tree (DESIGN.md "accel")      */
} ezrt_bvh_builder;

/* Sorts the list's triangles in place and builds the node array: nodes{testNode};
 * buildBVHwithSAH(triangles, nodes, 0, N-1, leaf_n) (P5/main.cpp:830-838).  Returns the
 * node count (incl. dummy node 0) or a negative status. */
int ezrt_trilist_build_bvh(ezrt_trilist* list, int leaf_n, int builder);
int ezrt_trilist_node_count(const ezrt_trilist* list);
/* 1 if this host's std::sort orders equal keys as libstdc++ does, i.e. the SAH_LITERAL / SAH_FAST / MEDIAN builders
 * reproduce the triangle order of the reference built with libstdc++ (its buildBVH* sort with order-only comparators,
 * P5/main.cpp:403-413, :560-568) and the goldens of this repository; 0 = the trees are valid but differ from the
 * reference's wherever centroid coordinates tie (ezrt_trilist_build_bvh then warns once on stderr). */
int ezrt_host_sort_is_reference(void);
/* Encode as P5/main.cpp:843-871 into caller buffers (36 floats/triangle, 12 floats/node). */
int ezrt_trilist_encode_triangles(const ezrt_trilist* list, float* tris_out);
int ezrt_trilist_encode_nodes(const ezrt_trilist* list, float* nodes_out);

/* Scene description file in place of the hard-coded scene blocks of main()
 * (P3/main.cpp:688-701, P4/main.cpp:687-729, P5/main.cpp:795-823): directives `set <field> ...`, `reset`,
 * `mesh <obj> smooth|flat [hardened] rotate.. translate.. scale..`, `camera <rotatAngle> <upAngle> <r>`,
 * `hdr <path>`; appends the meshes to `list`; camera[3] and hdr_path (nullable) receive the other settings. */
int ezrt_scene_file_load(const char* path, ezrt_trilist* list, float camera[3], char* hdr_path, size_t hdr_path_cap);

/* HDRLoader::load (P5/lib/hdrloader.cpp:29-97): Radiance .hdr -> float RGB rows.  Call with
 * cols = NULL to query width/height.  Returns 0, or 1 when the pixel data ended early or was malformed (a run marker
 * with nothing to repeat, a run length beyond 32 bits): like the reference the rows decoded so far are kept and the
 * rest is zero, but unlike it nothing outside the scanline buffer is read; < 0 on errors. */
int ezrt_hdr_load(const char* path, int* width, int* height, float* cols);
/* calculateHdrCache (P5/main.cpp:592-689): (sample_x, sample_y, pdf) lookup texture. */
int ezrt_hdr_cache(const float* hdr, int width, int height, float* cache_out);

/* calculateHdrCache on the GPU `device`: same bits as ezrt_hdr_cache -- every fp32
 * sum keeps the reference's order -- host arrays in and out; device_ms (nullable) = kernel time. */
int ezrt_hdr_cache_device(int device, const float* hdr, int width, int height, float* cache_out, double* device_ms);

/* Camera of display() (P5/main.cpp:710-713): orbit angles in degrees + radius ->
 * eye, cameraRotate = inverse(lookAt(eye, 0, +y)), column-major. */
void ezrt_camera_orbit(float rotate_angle_deg, float up_angle_deg, float r, float eye[3],
                       float camera_rotate[16]);

#ifdef __cplusplus
}
#endif
#endif /* EZRT_H */
