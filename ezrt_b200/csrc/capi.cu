// capi.cu -- device half of the C ABI (include/ezrt.h): scene upload/repack, render
// orchestration (wavefront + megakernel), image partition helpers and the single-function
// test entry points.  Host-side scene building lives in host_scene.cpp.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <array>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "device_scene.h"
#include "ezrt.h"
#include "ezrt_internal.h"
#include "ezrt_math.h"
#include "kernels.h"
#include "w8_node.h"

#define CU_CHECK(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return ezrt_set_error(EZRT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                  __FILE__, __LINE__);                                                  \
    } while (0)

namespace {

struct DeviceBuffer {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return EZRT_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) return ezrt_set_error(EZRT_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
        bytes = need;
        return EZRT_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

// tiles of part `rank` of `count` (ezrt_internal.h): row-major tile order, (tx+ty)%count == rank
std::vector<TileDev> partition_tiles(int width, int height, int rank, int count) {
    std::vector<TileDev> tiles;
    int tx_n = (width + EZRT_TILE - 1) / EZRT_TILE, ty_n = (height + EZRT_TILE - 1) / EZRT_TILE;
    int offset = 0;
    for (int ty = 0; ty < ty_n; ty++)
        for (int tx = 0; tx < tx_n; tx++) {
            if ((tx + ty) % count != rank) continue;
            TileDev t;
            t.x0 = tx * EZRT_TILE;
            t.y0 = ty * EZRT_TILE;
            t.w = std::min(EZRT_TILE, width - t.x0);
            t.h = std::min(EZRT_TILE, height - t.y0);
            t.pixel_offset = offset;
            offset += t.w * t.h;
            tiles.push_back(t);
        }
    return tiles;
}

}  // namespace

struct ezrt_scene {
    int device = 0;
    int n_sms = 148;
    SceneDev dev{};
    DeviceBuffer nodes, tri_geo, tri_shade, materials, hdr, hdr_cache;
    DeviceBuffer acc_hot, acc_tri_ref, tri_leaf, leaf_box, defer_buf, acc_tri_leaf, ref_to_acc, acc_wide, acc_wide_q16;   // acc_hot = W8 nodes | geometry | shading records of the accel order
    int acc_depth = 0;
    int n_materials = 0;
    int tree_depth = 0;
    // render state (lazily sized)
    DeviceBuffer tiles_buf, queue_buf[2], shadow_buf, lo_buf, le_buf, counters_buf, totals_buf, fb_buf, sort_buf;
    void* hot_base = nullptr;   // accel nodes | geometry | shading records (L2 persisting window)
    size_t hot_bytes = 0;
    size_t l2_persist_bytes = 0;
    size_t max_window_bytes = 0;  // persisting L2 set-aside granted by the device (0 = feature off)
    bool regular_tree = true;  // false: only the literal REFERENCE traversal is valid for the caller's tree
    bool have_accel = true;    // false: no acceleration tree (irregular caller tree, or a degenerate one too deep for the stacks): ACCEL runs as PRUNED
    int camera_pixel_major = 1;  // camera pass: a warp traces the samples of one pixel (env EZRT_CAMERA_ORDER=frame: an 8x4 block of one frame)
    int sort_rays = 0;  // env EZRT_SORT_RAYS=1 enables the bounce-ray sort (measured: no gain with per-lane refill)
    int tiles_key[4] = {-1, -1, -1, -1};
    std::vector<TileDev> tiles;
    size_t n_pixels = 0;       // pixels of the owned tiles
    size_t fmax = 0, fmax_key[2] = {0, 0};   // frames per batch that fit the free memory, for (slots per frame, bytes per slot)
    cudaStream_t own_stream = nullptr, copy_stream = nullptr;
    // the accel policy's deferred lane: exact traversal + shading of the few deferred rays beside the main k_shade
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    DeviceBuffer side_hit_buf;
    int deferred_lane = 1;   // env EZRT_DEFERRED_LANE=0: the exact pass in line, as in round 1
    cudaEvent_t fb_event = nullptr, fb_wait = nullptr;   // ezrt_render: the H2D of lastFrame runs beside the tracing kernels
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    bool have_timing = false;
    unsigned long long launches = 0;
    // params.profile: one event pair per launch, summed per kernel class by ezrt_get_kernel_times
    std::vector<cudaEvent_t> ev_pool;
    struct Span { int cls; int e0, e1; };
    std::vector<Span> spans;
    size_t ev_used = 0;
    bool profiling = false;
    int span_begin(int cls, cudaStream_t st) {
        if (!profiling) return -1;
        while (ev_pool.size() < ev_used + 2) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return -1;
            ev_pool.push_back(e);
        }
        Span sp{cls, (int)ev_used, (int)ev_used + 1};
        ev_used += 2;
        cudaEventRecord(ev_pool[sp.e0], st);
        spans.push_back(sp);
        return (int)spans.size() - 1;
    }
    void span_end(int id, cudaStream_t st) {
        if (id >= 0) cudaEventRecord(ev_pool[spans[id].e1], st);
    }
};

namespace {

int carve_queue(DeviceBuffer& buf, size_t capacity, PathQueue& q) {
    size_t per = sizeof(float4) * 4 + sizeof(float2);
    int rc = buf.ensure(per * capacity + 256);
    if (rc) return rc;
    char* p = (char*)buf.p;
    q.ray_o = (float4*)p; p += sizeof(float4) * capacity;
    q.ray_d = (float4*)p; p += sizeof(float4) * capacity;
    q.hist = (float4*)p;  p += sizeof(float4) * capacity;
    q.fr = (float4*)p;    p += sizeof(float4) * capacity;
    q.hit = (float2*)p;
    return EZRT_OK;
}
int carve_shadow(DeviceBuffer& buf, size_t capacity, ShadowQueue& q) {
    int rc = buf.ensure((size_t)EZRT_SHADOW_SLOT_BYTES * capacity + 256);
    if (rc) return rc;
    char* p = (char*)buf.p;
    q.ray_o = (float4*)p; p += sizeof(float4) * capacity;
    q.ray_d = (float4*)p; p += sizeof(float4) * capacity;
    q.nrm = (float4*)p;   p += sizeof(float4) * capacity;
    q.view = (float4*)p;  p += sizeof(float4) * capacity;
    q.hist = (float4*)p;  p += sizeof(float4) * capacity;
    q.lit = (unsigned char*)p;
    return EZRT_OK;
}

int validate_params(const ezrt_scene* scene, const ezrt_render_params* p) {
    if (!scene || !p) return ezrt_set_error(EZRT_ERR_INVALID, "render: null argument");
    if (p->width <= 0 || p->height <= 0 || p->spp < 0) return ezrt_set_error(EZRT_ERR_INVALID, "render: bad image size/spp");
    if (p->mode < 0 || p->mode > 3) return ezrt_set_error(EZRT_ERR_INVALID, "render: unknown mode %d", p->mode);
    if (p->max_bounce < 0 || p->max_bounce > 64) return ezrt_set_error(EZRT_ERR_INVALID, "render: max_bounce out of range");
    if (p->out_channels != 3 && p->out_channels != 4) return ezrt_set_error(EZRT_ERR_INVALID, "render: out_channels must be 3 or 4");
    if (p->part_count < 1 || p->part_rank < 0 || p->part_rank >= p->part_count)
        return ezrt_set_error(EZRT_ERR_INVALID, "render: bad partition %d/%d", p->part_rank, p->part_count);
    if (p->mode == EZRT_MODE_DISNEY_IS_MIS_P5 && (!scene->dev.hdr || !scene->dev.hdr_cache))
        return ezrt_set_error(EZRT_ERR_INVALID, "render: IS/MIS mode needs an HDR map and its cache");
    return EZRT_OK;
}

int prepare_tiles(ezrt_scene* s, const ezrt_render_params* p, cudaStream_t st) {
    int key[4] = {p->width, p->height, p->part_rank, p->part_count};
    if (memcmp(key, s->tiles_key, sizeof(key)) == 0) return EZRT_OK;
    s->tiles = partition_tiles(p->width, p->height, p->part_rank, p->part_count);
    size_t bytes = sizeof(TileDev) * std::max<size_t>(1, s->tiles.size());
    int rc = s->tiles_buf.ensure(bytes);
    if (rc) return rc;
    if (!s->tiles.empty()) CU_CHECK(cudaMemcpyAsync(s->tiles_buf.p, s->tiles.data(), sizeof(TileDev) * s->tiles.size(), cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaStreamSynchronize(st));  // s->tiles is pageable host memory
    memcpy(s->tiles_key, key, sizeof(key));
    s->n_pixels = 0;
    for (const TileDev& t : s->tiles) s->n_pixels += (size_t)t.w * t.h;
    return EZRT_OK;
}

struct ScatterEntry { TileDev* d_tiles; int n; };
std::map<std::array<int, 5>, ScatterEntry>& scatter_cache() {
    static std::map<std::array<int, 5>, ScatterEntry> cache;
    return cache;
}
std::mutex& scatter_mutex() {
    static std::mutex mu;
    return mu;
}

RenderDev make_render_dev(const ezrt_scene* s, const ezrt_render_params* p) {
    RenderDev rd;
    rd.width = p->width; rd.height = p->height;
    rd.mode = p->mode; rd.max_bounce = p->max_bounce; rd.traverse = p->traverse;
    memcpy(rd.eye, p->eye, sizeof(rd.eye));
    memcpy(rd.cam, p->camera_rotate, sizeof(rd.cam));
    memcpy(rd.env, p->env_color, sizeof(rd.env));
    rd.first_frame = p->first_frame;
    rd.out_channels = p->out_channels;
    rd.compact_out = (p->part_count > 1) ? 1 : 0;
    rd.n_tiles = (int)s->tiles.size();
    rd.accel_space = (s->regular_tree && s->have_accel && p->traverse == EZRT_TRAVERSE_ACCEL && p->pipeline == EZRT_PIPELINE_WAVEFRONT) ? 1 : 0;
    return rd;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------
// scene upload + repack (replaces the TBO / texture uploads P5/main.cpp:878-906)
// ------------------------------------------------------------------------------------------
int ezrt_scene_create(int device, const float* tris, int n_triangles, const float* nodes, int n_nodes, const float* hdr,
                      const float* hdr_cache, int hdr_w, int hdr_h, int hdr_filter_linear, ezrt_scene** out_scene) {
    if (!tris || !nodes || !out_scene || n_triangles <= 0 || n_nodes < 2)
        return ezrt_set_error(EZRT_ERR_INVALID, "scene_create: need triangles and at least the dummy + root node");
    if ((hdr || hdr_cache) && (hdr_w <= 0 || hdr_h <= 0)) return ezrt_set_error(EZRT_ERR_INVALID, "scene_create: bad HDR size");
    if (n_triangles >= (1 << 24)) return ezrt_set_error(EZRT_ERR_INVALID, "scene_create: more than 2^24 triangles (ints-as-floats limit)");
    int n_dev = 0;
    CU_CHECK(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return ezrt_set_error(EZRT_ERR_CUDA, "scene_create: no CUDA device %d (have %d)", device, n_dev);
    CU_CHECK(cudaSetDevice(device));
    // env EZRT_VERBOSE=1: wall-clock of the stages on stderr (all host work except the uploads)
    const bool verbose = getenv("EZRT_VERBOSE") && atoi(getenv("EZRT_VERBOSE")) != 0;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!verbose) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[ezrt_scene_create] %-44s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };

    // ---- decode + validate the tree (getBVHNode, P5/fsh:138-155) ----
    // Host work on the caller's tree only: runs on its own thread while this one feeds the GPU (triangle upload, records,
    // acceleration-tree build).  Errors are carried back as (code, message): ezrt_set_error is per thread.
    struct HNode { int left, right, n, index; };
    std::vector<HNode> hn(n_nodes);
    std::vector<int> inner_id(n_nodes, -1);
    int n_inner = 0, max_depth = 0, n_top = 0, root_ref = 0;
    bool regular_tree = true;
    std::vector<float4> gnodes, leaf_box;
    std::vector<int> tri_leaf(n_triangles, 0);
    std::string ref_msg;
    auto ref_fail = [&](int code, const char* fmt, int a = 0, int b = 0, int c = 0) {
        char buf[256];
        snprintf(buf, sizeof(buf), fmt, a, b, c);
        ref_msg = buf;
        return code;
    };
    auto ref_stage = [&]() -> int {
    for (int i = 0; i < n_nodes; i++) {
        const float* s = nodes + (size_t)i * EZRT_BVHNODE_FLOATS;
        hn[i].left = (int)s[0]; hn[i].right = (int)s[1]; hn[i].n = (int)s[3]; hn[i].index = (int)s[4];
    }
    std::vector<char> seen(n_nodes, 0);
    {
        std::vector<std::pair<int, int>> stk;
        stk.push_back({1, 1});
        while (!stk.empty()) {
            auto [i, depth] = stk.back();
            stk.pop_back();
            if (i < 1 || i >= n_nodes) return ref_fail(EZRT_ERR_BAD_TREE, "scene_create: child index %d out of range", i);
            if (seen[i]) return ref_fail(EZRT_ERR_BAD_TREE, "scene_create: node %d reachable twice", i);
            seen[i] = 1;
            max_depth = std::max(max_depth, depth);
            const HNode& nd = hn[i];
            if (nd.n > 0) {
                if (nd.n > EZRT_LEAF_MAX_N) return ref_fail(EZRT_ERR_BAD_TREE, "scene_create: leaf %d holds %d > %d triangles", i, nd.n, EZRT_LEAF_MAX_N);
                if (nd.index < 0 || nd.index + nd.n > n_triangles) return ref_fail(EZRT_ERR_BAD_TREE, "scene_create: leaf %d range out of bounds", i);
            } else {
                // the shader would read the dummy node 0 for a missing child (P5/fsh:278-302); not supported
                if (nd.left <= 0 || nd.right <= 0) return ref_fail(EZRT_ERR_BAD_TREE, "scene_create: inner node %d lacks a child", i);
                stk.push_back({nd.right, depth + 1});
                stk.push_back({nd.left, depth + 1});
            }
        }
    }
    if (max_depth + 1 > EZRT_MAX_STACK) return ref_fail(EZRT_ERR_BAD_TREE, "scene_create: tree depth %d exceeds %d", max_depth, EZRT_MAX_STACK - 1);
    // The ACCEL and PRUNED policies rely on what buildBVH* guarantees: every triangle in exactly one leaf,
    // leaf boxes bounding their triangles, child boxes inside their parent's.  A caller-supplied tree that
    // breaks any of these is only ever walked literally (REFERENCE policy), whatever the params ask for.
    {
        std::vector<unsigned char> cover(n_triangles, 0);
        for (int i = 1; i < n_nodes && regular_tree; i++) {
            if (!seen[i]) continue;
            const float* B = nodes + (size_t)i * EZRT_BVHNODE_FLOATS;
            if (hn[i].n > 0) {
                for (int k = 0; k < hn[i].n && regular_tree; k++) {
                    const int t = hn[i].index + k;
                    if (cover[t]++) regular_tree = false;
                    const float* v = tris + (size_t)t * EZRT_TRIANGLE_FLOATS;
                    for (int c = 0; c < 9; c++)
                        if (!(v[c] >= B[6 + c % 3] && v[c] <= B[9 + c % 3])) regular_tree = false;
                }
            } else {
                for (int c : {hn[i].left, hn[i].right}) {
                    const float* C = nodes + (size_t)c * EZRT_BVHNODE_FLOATS;
                    for (int a = 0; a < 3; a++)
                        if (!(C[6 + a] >= B[6 + a] && C[9 + a] <= B[9 + a])) regular_tree = false;
                }
            }
        }
        for (int t = 0; t < n_triangles && regular_tree; t++)
            if (cover[t] != 1) regular_tree = false;
    }
    // Record numbering: the top of the tree level by level (these records are staged in shared memory
    // by the traversal kernels: 61% of all inner-node visits hit depth <= 9 on the 1M-triangle scene),
    // whole levels while they fit EZRT_TOP_NODES_MAX; everything below in the builder's pre-order.
    {
        std::vector<int> level, next;
        if (hn[1].n <= 0) level.push_back(1);
        while (!level.empty() && n_top + (int)level.size() <= EZRT_TOP_NODES_MAX) {
            next.clear();
            for (int i : level) {
                inner_id[i] = n_top++;
                if (hn[hn[i].left].n <= 0) next.push_back(hn[i].left);
                if (hn[hn[i].right].n <= 0) next.push_back(hn[i].right);
            }
            level.swap(next);
        }
    }
    n_inner = n_top;
    for (int i = 1; i < n_nodes; i++)
        if (seen[i] && hn[i].n <= 0 && inner_id[i] < 0) inner_id[i] = n_inner++;

    auto child_ref = [&](int c) -> int {
        if (hn[c].n > 0) return (int)(EZRT_LEAF_FLAG | ((uint32_t)hn[c].index << 7) | (uint32_t)hn[c].n);
        return inner_id[c];
    };
    auto pack_node = [](float4* g, const float* LA, const float* LB, const float* RA, const float* RB, int rl, int rr) {
        float fl, fr;
        memcpy(&fl, &rl, 4);
        memcpy(&fr, &rr, 4);
        g[0] = make_float4(LA[0], LA[1], LB[0], LB[1]);  // left : AA.x AA.y | BB.x BB.y
        g[1] = make_float4(RA[0], RA[1], RB[0], RB[1]);  // right: AA.x AA.y | BB.x BB.y
        g[2] = make_float4(LA[2], LB[2], RA[2], RB[2]);  // left AA.z BB.z  | right AA.z BB.z
        g[3] = make_float4(fl, fr, 0.0f, 0.0f);          // child references
    };
    gnodes.assign((size_t)std::max(1, n_inner) * 4, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    for (int i = 1; i < n_nodes; i++) {
        if (inner_id[i] < 0) continue;
        const float* L = nodes + (size_t)hn[i].left * EZRT_BVHNODE_FLOATS;
        const float* R = nodes + (size_t)hn[i].right * EZRT_BVHNODE_FLOATS;
        pack_node(&gnodes[(size_t)inner_id[i] * 4], L + 6, L + 9, R + 6, R + 9, child_ref(hn[i].left), child_ref(hn[i].right));
    }
    // reference leaf of every triangle + leaf boxes (accel policy: "does the shader reach this leaf?")
    for (int i = 1; i < n_nodes; i++) {
        if (!seen[i] || hn[i].n <= 0) continue;
        const float* B = nodes + (size_t)i * EZRT_BVHNODE_FLOATS;
        int slot = (int)(leaf_box.size() / 2);
        leaf_box.push_back(make_float4(B[6], B[7], B[8], 0.0f));
        leaf_box.push_back(make_float4(B[9], B[10], B[11], 0.0f));
        for (int k = 0; k < hn[i].n; k++) tri_leaf[hn[i].index + k] = slot;
    }

    root_ref = child_ref(1);
    return EZRT_OK;
    };   // ref_stage
    // ---- the scene object, the worker for the caller's tree, and the clean-up of every early return ----
    ezrt_scene* sc = new (std::nothrow) ezrt_scene();
    if (!sc) return ezrt_set_error(EZRT_ERR_NOMEM, "scene_create: out of host memory");
    sc->device = device;
    int ref_rc = EZRT_OK;
    double ref_ms = 0.0;
    struct Cleanup {   // declared after everything the worker touches: joined first, then those locals go
        ezrt_scene* sc = nullptr;
        std::thread worker;
        DeviceBuffer raw;
        ~Cleanup() {
            if (worker.joinable()) worker.join();
            raw.release();
            if (sc) ezrt_scene_destroy(sc);
        }
    } guard;
    guard.sc = sc;
    auto ref_timed = [&]() {
        const auto t0 = std::chrono::steady_clock::now();
        ref_rc = ref_stage();
        ref_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    try {
        guard.worker = std::thread(ref_timed);
    } catch (...) {
        ref_timed();   // no thread to be had: in line
    }
    // ---- triangles: upload the caller's array once; geometry records, shading records, material runs and the scene
    // bounds are made from it on the device (scene_prep.cu) ----
    DeviceBuffer& raw = guard.raw;
    {
        int r = raw.ensure((size_t)n_triangles * EZRT_TRIANGLE_FLOATS * sizeof(float));
        if (r) return r;
        if (cudaMemcpy(raw.p, tris, (size_t)n_triangles * EZRT_TRIANGLE_FLOATS * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
            return ezrt_set_error(EZRT_ERR_CUDA, "scene_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    lap("triangles: upload");
    EzrtPrepInfo prep;
    {
        int r = sc->tri_geo.ensure((size_t)n_triangles * 4 * sizeof(float4));
        if (!r) r = sc->tri_shade.ensure((size_t)n_triangles * 3 * sizeof(float4));
        if (!r) r = ezrt_prep_records((const float*)raw.p, n_triangles, sc->tri_geo.p, sc->tri_shade.p, prep);
        if (r) return r;
    }
    std::vector<float4> mats;
    for (int m = 0; m < prep.n_materials; m++) {
        const float* v = &prep.materials[(size_t)m * EZRT_MATERIAL_FLOATS];
        mats.push_back(make_float4(v[0], v[1], v[2], v[3]));
        mats.push_back(make_float4(v[4], v[5], v[6], v[7]));
        mats.push_back(make_float4(v[8], v[9], v[10], v[11]));
        mats.push_back(make_float4(v[12], v[13], v[14], v[15]));
        mats.push_back(make_float4(v[16], v[17], 0.0f, 0.0f));
    }
    const float max_abs = prep.max_abs;
    float bmin[3], bmax[3];
    for (int k = 0; k < 3; k++) { bmin[k] = prep.bmin[k]; bmax[k] = prep.bmax[k]; }
    lap("triangles: geometry / shading records, materials (device)");
    // ---- acceleration tree: sentinel-free SAH over the same triangles (DESIGN.md "accel"), collapsed to 8-wide
    // quantised nodes (accel_w8.cpp, w8_node.h).  A tree too deep for the traversal stacks (degenerate input) is
    // dropped: the scene then renders with the PRUNED policy on the caller's tree, as for an irregular tree.
    const float prune_delta = max_abs * 1.52587890625e-05f;  // 2^-16 * scene extent
    EzrtW4Tree w4;
    EzrtRawArray<float>& acc_wide = w4.nodes;
    EzrtRawArray<uint32_t>& acc_wide_q = w4.q16;
    std::vector<uint32_t> w8_words;
    int acc_wide_root = 0;
    std::vector<uint32_t> acc_order;
    int acc_depth = 0, w8_depth = 0, w8_near_bit[3] = {0, 1, 2};
    bool have_accel = true;
    {
        std::vector<EzrtAccelNode> an;
        std::vector<uint32_t> order_bin;
        // on the GPU (accel_build.cu; the same tree node for node); env EZRT_BUILD=host: the host builder (host_scene.cpp)
        const char* be = getenv("EZRT_BUILD");
        if (be && !strcmp(be, "host")) {
            ezrt_build_accel(tris, n_triangles, W8_MAX_LEAF_TRIS, an, order_bin);
            lap("acceleration tree: binary SAH build (host)");
        } else {
            const int brc = ezrt_build_accel_device((const float*)raw.p, n_triangles, W8_MAX_LEAF_TRIS, an, order_bin, nullptr);
            if (brc < 0) return brc;
            lap("acceleration tree: binary SAH build (device)");
        }
        // boxes inflated by 2*delta: a hit hitTriangle accepts lies within delta of its triangle's box, so
        // the inflated boxes of the whole ancestor chain are entered no later than the hit distance
        const float pad = 2.0f * prune_delta;
        // Which form of the tree the accel kernels walk (env EZRT_ACCEL, read at scene creation):
        //   4 (default): 4-wide nodes with exact fp32 boxes, children sorted by entry distance (k_extend_accel): measured
        //                faster on B200 (3272 vs 3005 Mrays/s on C3, profiles/sweep_w8_r2.txt)
        //   8          : 8-wide nodes with 8-bit quantised boxes in octant order (k_extend_w8): 45 % fewer L1 wavefronts per
        //                ray, 36 % more instructions
        int accel_form = 4;
        if (const char* we = getenv("EZRT_ACCEL")) accel_form = (atoi(we) == 8) ? 8 : 4;
        if (an[0].n > 0) accel_form = 8;   // a single-leaf tree: the W8 builder handles it
        EzrtW8Tree w8;
        ezrt_w8_axis_bits(bmin, bmax, w8_near_bit);
        if (accel_form == 8) {
            const int wrc = ezrt_build_w8(an, order_bin, pad, max_abs, w8_near_bit, w8);
            if (wrc != 0 || w8.depth > EZRT_W8_SMEM_STACK + W8_LOCAL_STACK) {
                have_accel = false;
            } else {
                acc_order = w8.tri_order;
                w8_words.swap(w8.nodes);
                w8_depth = w8.depth;
            }
        } else {
            // 4-wide collapse: SAH-optimal choice of each node's children (EzrtCollapse; env EZRT_W4_COLLAPSE=greedy: round 1's
            // "replace the largest inner child" rule); leaves = sub-trees of <= 4 consecutive triangles; packed by accel_w8.cpp
            const char* ce = getenv("EZRT_W4_COLLAPSE");
            const bool greedy = ce && !strcmp(ce, "greedy");
            const char* qe = getenv("EZRT_ACCEL_Q16");
            const bool want_q16 = !(qe && atoi(qe) == 0);
            if (ezrt_build_w4(an, pad, max_abs, greedy, want_q16, ezrt_host_threads(), w4) != 0) {
                have_accel = false;
            } else {
                acc_order = order_bin;
                acc_wide_root = w4.root;
                acc_depth = w4.depth;
                if (3 * w4.depth + 2 > EZRT_ACCEL_STACK) { w4.nodes.clear(); have_accel = false; }  // too deep for the traversal stack
                if (!have_accel) w4.q16.clear();
            }
        }
        if (!have_accel) {
            acc_order.resize(n_triangles);
            for (int i = 0; i < n_triangles; i++) acc_order[i] = (uint32_t)i;
        }
        raw.release();
    }
    lap("acceleration tree: collapse, pack");
    // ---- the caller's tree is needed from here on ----
    if (guard.worker.joinable()) guard.worker.join();
    if (verbose) fprintf(stderr, "[ezrt_scene_create] %-44s %8.1f ms (worker thread, overlapped)\n", "reference tree: decode, validate, repack", ref_ms);
    if (ref_rc) return ezrt_set_error(ref_rc, "%s", ref_msg.c_str());
    lap("wait for the reference-tree worker");
    cudaDeviceProp prop;   // cudaGetDeviceProperties takes milliseconds: once per device and process
    {
        static std::mutex mu;
        static std::map<int, cudaDeviceProp> cache;
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(device);
        if (it == cache.end()) {
            memset(&prop, 0, sizeof(prop));
            if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) cudaGetLastError();
            else cache[device] = prop;
        } else {
            prop = it->second;
        }
    }
    if (prop.multiProcessorCount > 0) {
        sc->n_sms = prop.multiProcessorCount;
        sc->max_window_bytes = (size_t)prop.accessPolicyMaxWindowSize;
    }
    int rc = EZRT_OK;
    auto upload = [&](DeviceBuffer& b, const void* src, size_t bytes) -> int {
        int r = b.ensure(std::max<size_t>(bytes, 16));
        if (r) return r;
        if (bytes && cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
            return ezrt_set_error(EZRT_ERR_CUDA, "scene_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return EZRT_OK;
    };
    EzrtLap sub("ezrt_scene_create uploads");
    if (!rc) rc = upload(sc->nodes, gnodes.data(), gnodes.size() * sizeof(float4));
    if (!rc) rc = upload(sc->materials, mats.data(), mats.size() * sizeof(float4));
    sub("reference-tree nodes, materials");
    // acceleration tree nodes | triangle geometry | shading records in ONE allocation: the window the
    // L2 persisting-access policy is set on while a render runs (the data every ray touches at random)
    const size_t acc_nodes_bytes = ((std::max<size_t>(w8_words.size(), 4) * sizeof(uint32_t) + 255) / 256) * 256;
    const size_t acc_geo_bytes = (((size_t)n_triangles * 4 * sizeof(float4) + 255) / 256) * 256;
    const size_t acc_shade_bytes = (((size_t)n_triangles * 3 * sizeof(float4) + 255) / 256) * 256;
    if (!rc) rc = sc->acc_hot.ensure(acc_nodes_bytes + acc_geo_bytes + acc_shade_bytes);
    if (!rc) {
        char* base = (char*)sc->acc_hot.p;
        cudaError_t e = w8_words.empty() ? cudaSuccess : cudaMemcpy(base, w8_words.data(), w8_words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = ezrt_set_error(EZRT_ERR_CUDA, "scene_create: upload failed: %s", cudaGetErrorString(e));
        sc->hot_base = base;
        sc->hot_bytes = acc_nodes_bytes + acc_geo_bytes + acc_shade_bytes;
    }
    sub("hot allocation");
    if (!rc) rc = upload(sc->acc_tri_ref, acc_order.data(), acc_order.size() * sizeof(uint32_t));
    if (!rc && !acc_wide.empty()) rc = upload(sc->acc_wide, acc_wide.data(), acc_wide.size() * sizeof(float));
    if (!rc && !acc_wide_q.empty()) rc = upload(sc->acc_wide_q16, acc_wide_q.data(), acc_wide_q.size() * sizeof(uint32_t));
    if (!rc) rc = upload(sc->tri_leaf, tri_leaf.data(), tri_leaf.size() * sizeof(int));
    if (!rc) rc = upload(sc->leaf_box, leaf_box.data(), leaf_box.size() * sizeof(float4));
    sub("order, wide nodes, leaf map, leaf boxes");
    // geometry, shading records and the reference-leaf map in the acceleration tree's order, and the inverse permutation:
    // gathered on the device
    if (!rc) rc = sc->acc_tri_leaf.ensure((size_t)n_triangles * sizeof(int));
    if (!rc) rc = sc->ref_to_acc.ensure((size_t)n_triangles * sizeof(uint32_t));
    if (!rc)
        rc = ezrt_prep_gather(sc->tri_geo.p, sc->tri_shade.p, (const int*)sc->tri_leaf.p, (const uint32_t*)sc->acc_tri_ref.p, n_triangles,
                              (char*)sc->acc_hot.p + acc_nodes_bytes, (char*)sc->acc_hot.p + acc_nodes_bytes + acc_geo_bytes,
                              (int*)sc->acc_tri_leaf.p, (uint32_t*)sc->ref_to_acc.p);
    sub("gather");
    if (!rc && hdr) rc = upload(sc->hdr, hdr, sizeof(float) * 3 * (size_t)hdr_w * hdr_h);
    if (!rc && hdr_cache) rc = upload(sc->hdr_cache, hdr_cache, sizeof(float) * 3 * (size_t)hdr_w * hdr_h);
    sub("environment map, cache");
    if (!rc && cudaStreamCreateWithFlags(&sc->own_stream, cudaStreamNonBlocking) != cudaSuccess) rc = ezrt_set_error(EZRT_ERR_CUDA, "scene_create: stream");
    if (!rc && (cudaEventCreate(&sc->ev_start) != cudaSuccess || cudaEventCreate(&sc->ev_stop) != cudaSuccess)) rc = ezrt_set_error(EZRT_ERR_CUDA, "scene_create: events");
    sub("stream, events");
    if (rc) return rc;
    lap("uploads, records in the tree's order (device)");
    guard.sc = nullptr;   // from here on the scene is the caller's
    sc->n_materials = prep.n_materials;
    sc->regular_tree = regular_tree;
    sc->have_accel = have_accel;
    sc->tree_depth = max_depth;
    SceneDev& d = sc->dev;
    d.nodes = (const float4*)sc->nodes.p;
    d.tri_geo = (const float4*)sc->tri_geo.p;
    d.tri_shade = (const float4*)sc->tri_shade.p;
    d.materials = (const float4*)sc->materials.p;
    d.hdr = hdr ? (const float*)sc->hdr.p : nullptr;
    d.hdr_cache = hdr_cache ? (const float*)sc->hdr_cache.p : nullptr;
    d.hdr_w = hdr_w; d.hdr_h = hdr_h; d.hdr_linear = hdr_filter_linear ? 1 : 0;
    d.root_ref = root_ref;
    d.w8_nodes = w8_words.empty() ? nullptr : (const uint4*)sc->acc_hot.p;
    for (int k = 0; k < 3; k++) d.w8_near_bit[k] = w8_near_bit[k];
    d.w8_stack_entries = std::max(1, std::min(w8_depth, EZRT_W8_SMEM_STACK));
    d.w8_origin_limit = W8_ORIGIN_LIMIT_REL * max_abs;
    d.w8_decode_bits = W8_DECODE_BITS;
    d.w8_tri_weight = 2;
    if (const char* e = getenv("EZRT_TRI_W")) d.w8_tri_weight = std::max(1, std::min(64, atoi(e)));
    d.acc_tri_geo = (const float4*)((const char*)sc->acc_hot.p + acc_nodes_bytes);
    d.acc_tri_ref = (const uint32_t*)sc->acc_tri_ref.p;
    d.acc_wide_nodes = acc_wide.empty() ? nullptr : (const float4*)sc->acc_wide.p;
    d.acc_wide_root_ref = acc_wide_root;
    d.acc_wide_q16 = acc_wide_q.empty() ? nullptr : (const uint4*)sc->acc_wide_q16.p;
    d.q16_decode_bits = 0x4B000000u;
    d.tri_l1_bypass = ((size_t)n_triangles * 64 > ((size_t)4 << 20)) ? 1 : 0;  // > 4 MB of triangle records: stream them past L1
    if (const char* e = getenv("EZRT_TRI_L1_BYPASS")) d.tri_l1_bypass = atoi(e) != 0;
    d.acc_tri_shade = (const float4*)((const char*)sc->acc_hot.p + acc_nodes_bytes + acc_geo_bytes);
    d.acc_tri_leaf = (const int*)sc->acc_tri_leaf.p;
    d.ref_to_acc = (const uint32_t*)sc->ref_to_acc.p;
    d.tri_leaf = (const int*)sc->tri_leaf.p;
    d.leaf_box = (const float4*)sc->leaf_box.p;
    sc->acc_depth = acc_depth;
    d.n_triangles = n_triangles;
    d.n_inner = n_inner;
    d.top_nodes = n_top;
    if (const char* e = getenv("EZRT_TOP_NODES")) {
        d.top_nodes = std::max(0, std::min(n_top, atoi(e)));
    }
    d.prune_delta = prune_delta;  // 2^-16 * scene extent (DESIGN.md "pruning")
    for (int k = 0; k < 3; k++) {
        d.bmin[k] = bmin[k];
        float ext = bmax[k] - bmin[k];
        d.cell_scale[k] = (ext > 0.0f) ? 32.0f / ext : 0.0f;
    }
    if (const char* e = getenv("EZRT_SORT_RAYS")) sc->sort_rays = atoi(e);
    if (const char* e = getenv("EZRT_DEFERRED_LANE")) sc->deferred_lane = atoi(e) != 0;
    if (const char* e = getenv("EZRT_CAMERA_ORDER")) sc->camera_pixel_major = strcmp(e, "frame") != 0;
    {   // optional L2 persistence for the randomly-accessed tree data (env EZRT_L2_PERSIST=1)
        // measured on C3: -8 % (the set-aside shrinks the L2 left for the streaming queue traffic) -> off by default
        const char* e = getenv("EZRT_L2_PERSIST");
        if (e && atoi(e) != 0 && prop.persistingL2CacheMaxSize > 0) {
            size_t want = std::min<size_t>(sc->hot_bytes, (size_t)prop.persistingL2CacheMaxSize);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) sc->l2_persist_bytes = want;
            cudaGetLastError();
        }
    }
    d.refill_thresh = 24;
    d.refill_thresh_camera = 0;
    if (const char* e = getenv("EZRT_REFILL_CAM")) d.refill_thresh_camera = std::max(0, std::min(32, atoi(e)));
    d.inner_thresh = 16;
    d.leaf_thresh = 16;   // round 2 (cheaper leaf passes, S-1M): 16 measured +3 % over 12 (profiles/sweep_thresh_r2.txt); chunks: sweep_camera_r2.txt
    d.work_chunk = 32;
    d.work_chunk_camera = 64;
    if (const char* e = getenv("EZRT_CHUNK_CAM")) d.work_chunk_camera = std::max(32, std::min(65536, atoi(e)));
    if (const char* e = getenv("EZRT_CHUNK")) d.work_chunk = std::max(32, std::min(65536, atoi(e)));
    if (const char* e = getenv("EZRT_LEAF_T")) d.leaf_thresh = std::max(1, std::min(33, atoi(e)));
    if (const char* e = getenv("EZRT_REFILL_T")) d.refill_thresh = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("EZRT_INNER_T")) d.inner_thresh = std::max(1, std::min(32, atoi(e)));
    *out_scene = sc;
    return EZRT_OK;
}

int ezrt_scene_destroy(ezrt_scene* s) {
    if (!s) return EZRT_OK;
    cudaSetDevice(s->device);
    s->nodes.release(); s->tri_geo.release(); s->tri_shade.release(); s->materials.release();
    s->hdr.release(); s->hdr_cache.release(); s->tiles_buf.release();
    s->acc_hot.release(); s->acc_tri_ref.release(); s->tri_leaf.release(); s->leaf_box.release(); s->defer_buf.release();
    s->acc_tri_leaf.release(); s->ref_to_acc.release(); s->acc_wide.release(); s->acc_wide_q16.release();
    s->queue_buf[0].release(); s->queue_buf[1].release(); s->shadow_buf.release();
    s->lo_buf.release(); s->le_buf.release(); s->counters_buf.release(); s->totals_buf.release(); s->fb_buf.release(); s->sort_buf.release();
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->side_stream) cudaStreamDestroy(s->side_stream);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    s->side_hit_buf.release();
    if (s->fb_event) cudaEventDestroy(s->fb_event);
    if (s->ev_start) cudaEventDestroy(s->ev_start);
    if (s->ev_stop) cudaEventDestroy(s->ev_stop);
    for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
    delete s;
    return EZRT_OK;
}

// ------------------------------------------------------------------------------------------
// render
// ------------------------------------------------------------------------------------------
int ezrt_render_device(ezrt_scene* s, const ezrt_render_params* p, float* d_fb, void* cuda_stream) {
    int rc = validate_params(s, p);
    if (rc) return rc;
    if (!d_fb) return ezrt_set_error(EZRT_ERR_INVALID, "render: null framebuffer");
    CU_CHECK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    rc = prepare_tiles(s, p, st);
    if (rc) return rc;
    rc = s->totals_buf.ensure(sizeof(unsigned long long) * 8);
    if (rc) return rc;
    unsigned long long* totals = (unsigned long long*)s->totals_buf.p;
    // EZRT_PARAM_ACCUMULATE: counters, kernel spans and the device-time bracket continue from the previous render
    // (a benchmark reads them once after K renders instead of synchronising after every one)
    const bool accumulate = (p->reserved[0] & EZRT_PARAM_ACCUMULATE) != 0 && s->have_timing;
    if (!accumulate) {
        CU_CHECK(cudaEventRecord(s->ev_start, st));
        CU_CHECK(cudaMemsetAsync(totals, 0, sizeof(unsigned long long) * 8, st));
        s->launches = 0;
        s->spans.clear();
        s->ev_used = 0;
    }
    s->have_timing = false;   // set again once ev_stop is recorded (an early error return must not leave a dangling bracket)
    s->profiling = (p->profile == 1);
    RenderDev rd = make_render_dev(s, p);
    const TileDev* d_tiles = (const TileDev*)s->tiles_buf.p;
    const bool prune = s->regular_tree && (p->traverse != EZRT_TRAVERSE_REFERENCE);
    const bool accel = s->regular_tree && s->have_accel && (p->traverse == EZRT_TRAVERSE_ACCEL);
    if (rd.n_tiles == 0 || p->spp == 0) {
        CU_CHECK(cudaEventRecord(s->ev_stop, st));
        s->have_timing = true;
        return EZRT_OK;
    }

    if (p->pipeline == EZRT_PIPELINE_MEGAKERNEL) {
        if (s->fb_wait) {
            CU_CHECK(cudaStreamWaitEvent(st, s->fb_wait, 0));
            s->fb_wait = nullptr;
        }
        int sp = s->span_begin(0, st);
        launch_megakernel(s->dev, rd, d_tiles, prune, p->spp, d_fb, totals, st);
        s->span_end(sp, st);
        s->launches++;
        CU_CHECK(cudaGetLastError());
        CU_CHECK(cudaEventRecord(s->ev_stop, st));
        s->have_timing = true;
        return EZRT_OK;
    }

    const size_t per_frame = (size_t)rd.n_tiles * EZRT_TILE_PIXELS;
    const bool is_mode = (p->mode == EZRT_MODE_DISNEY_IS_MIS_P5);
    int F = p->frames_per_batch;
    if (F <= 0) F = (int)std::max<size_t>(1, ((size_t)32 << 20) / per_frame);  // ~32 M sample slots per batch (~7.5 GB of state):
                                                                              // long queues amortise the persistent kernels' ramp-up and tail
    F = std::min(F, p->spp);
    {   // bound the batch by the memory that is actually there (scratch already held by this scene counts as available);
        // asked once per (slots per frame, integrator): cudaMemGetInfo is a driver round trip, the render path is launch-only
        const size_t per_slot = 2 * (sizeof(float4) * 4 + sizeof(float2)) + 2 * sizeof(float4) + sizeof(uint32_t) +
                                (is_mode ? (size_t)EZRT_SHADOW_SLOT_BYTES : 0) + (s->sort_rays ? 2 * sizeof(uint32_t) : 0);
        if (s->fmax_key[0] != per_frame || s->fmax_key[1] != per_slot) {
            size_t free_b = 0, total_b = 0;
            s->fmax = (size_t)1 << 30;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                const size_t held = s->queue_buf[0].bytes + s->queue_buf[1].bytes + s->shadow_buf.bytes + s->lo_buf.bytes + s->le_buf.bytes +
                                    s->defer_buf.bytes + s->sort_buf.bytes;
                const size_t avail = (size_t)((double)(free_b + held) * 0.9);
                s->fmax = avail / per_slot / per_frame;
                if (s->fmax < 1) return ezrt_set_error(EZRT_ERR_NOMEM, "render: %zu MB free, one frame of wavefront state needs %zu MB", free_b >> 20, (per_slot * per_frame) >> 20);
            }
            s->fmax_key[0] = per_frame;
            s->fmax_key[1] = per_slot;
        }
        F = (int)std::min<size_t>((size_t)F, s->fmax);
    }
    const size_t capacity = per_frame * (size_t)F;
    if (capacity >= ((size_t)1 << 31)) return ezrt_set_error(EZRT_ERR_INVALID, "render: batch too large");
    PathQueue q[2];
    ShadowQueue sq{};
    if ((rc = carve_queue(s->queue_buf[0], capacity, q[0]))) return rc;
    if ((rc = carve_queue(s->queue_buf[1], capacity, q[1]))) return rc;
    if ((rc = carve_shadow(s->shadow_buf, is_mode ? capacity : 1, sq))) return rc;
    if ((rc = s->lo_buf.ensure(sizeof(float4) * capacity))) return rc;
    if ((rc = s->le_buf.ensure(sizeof(float4) * capacity))) return rc;
    const int n_stages = p->max_bounce + 2;
    // counters: [0,n) queue sizes, [n,2n) shadow sizes, [2n,3n) extend work, [3n,4n) shadow work,
    // [4n,6n) deferred-ray counts of the accel passes (extend, shadow), [6n,8n) work counters of their exact passes
    const int n_counters = 8 * n_stages;
    if ((rc = s->counters_buf.ensure(sizeof(uint32_t) * n_counters))) return rc;
    uint32_t* cnt = (uint32_t*)s->counters_buf.p;
    uint32_t *q_count = cnt, *s_count = cnt + n_stages, *w_ext = cnt + 2 * n_stages, *w_sh = cnt + 3 * n_stages;
    uint32_t *d_ext = cnt + 4 * n_stages, *d_sh = cnt + 5 * n_stages, *dw_ext = cnt + 6 * n_stages, *dw_sh = cnt + 7 * n_stages;
    if ((rc = s->defer_buf.ensure(sizeof(uint32_t) * (capacity + 64)))) return rc;
    uint32_t* defer_list = (uint32_t*)s->defer_buf.p;
    float4* Lo = (float4*)s->lo_buf.p;
    float4* Le = (float4*)s->le_buf.p;
    uint32_t *sort_keys = nullptr, *sort_perm = nullptr, *sort_bins = nullptr;
    if (s->sort_rays) {   // the optional bounce-ray sort (exact policies only)
        if ((rc = s->sort_buf.ensure(sizeof(uint32_t) * (2 * capacity + EZRT_SORT_BINS + 64)))) return rc;
        sort_keys = (uint32_t*)s->sort_buf.p;
        sort_perm = sort_keys + capacity;
        sort_bins = sort_perm + capacity;
    }

    unsigned long long* const count_ptr = (p->profile == 2) ? totals + 5 : nullptr;  // node visits, triangle tests of the W8 kernels
    // the deferred lane needs a second stream (highest priority: its small blocks go first when an SM has room), two events and
    // the hit records of up to EZRT_SIDE_CAP deferred rays
    const bool lane = accel && s->deferred_lane;
    float2* side_hit = nullptr;
    if (lane) {
        if (!s->side_stream) {
            int lo_pri = 0, hi_pri = 0;
            cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
            if (cudaStreamCreateWithPriority(&s->side_stream, cudaStreamNonBlocking, hi_pri) != cudaSuccess) return ezrt_set_error(EZRT_ERR_CUDA, "render: side stream");
            if (cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) != cudaSuccess)
                return ezrt_set_error(EZRT_ERR_CUDA, "render: events");
        }
        if ((rc = s->side_hit_buf.ensure(sizeof(float2) * (size_t)EZRT_SIDE_CAP))) return rc;
        side_hit = (float2*)s->side_hit_buf.p;
    }
    const bool l2_window = accel && s->l2_persist_bytes > 0 && s->hot_bytes > 0;
    if (l2_window) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = s->hot_base;
        attr.accessPolicyWindow.num_bytes = std::min<size_t>(s->hot_bytes, (size_t)s->max_window_bytes);
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)s->l2_persist_bytes / (double)attr.accessPolicyWindow.num_bytes);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
    }
    for (int done = 0; done < p->spp; done += F) {
        const int nf = std::min(F, p->spp - done);
        const uint32_t n_slots = (uint32_t)(per_frame * (size_t)nf);
        const uint32_t batch_first = p->first_frame + (uint32_t)done;
        CU_CHECK(cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * n_counters, st));
        const bool fused_camera = accel;   // accel policy: camera rays are generated inside the first extend kernel
        int sp = -1;
        if (!fused_camera) {
            sp = s->span_begin(3, st);
            launch_generate(rd, d_tiles, n_slots, batch_first, q[0], &q_count[0], s->n_sms, st);
            s->span_end(sp, st);
            s->launches++;
        }
        for (int b = 0; b <= p->max_bounce; b++) {
            PathQueue& qin = q[b & 1];
            PathQueue& qout = q[(b + 1) & 1];
            const uint32_t* perm = nullptr;
            if (b > 0 && s->sort_rays) {  // optional experiment
                sp = s->span_begin(3, st);
                launch_ray_sort(s->dev, qin, &q_count[b], sort_keys, sort_bins, sort_perm, n_slots, s->n_sms, st);
                s->span_end(sp, st);
                s->launches += 3;
                perm = sort_perm;
            }
            sp = s->span_begin(0, st);
            const int exact_gate = lane ? 2 : 0;   // lane: only an overflowing list of deferred rays is traced in line
            if (accel && fused_camera && b == 0) {
                launch_extend_camera(s->dev, rd, d_tiles, batch_first, n_slots, s->camera_pixel_major ? (uint32_t)nf : 0u, qin, &w_ext[b], defer_list, &d_ext[b], &dw_ext[b],
                                     s->n_sms, count_ptr, st, exact_gate);
                s->launches++;
            } else if (accel) {
                launch_extend_accel(s->dev, qin, &q_count[b], &w_ext[b], defer_list, &d_ext[b], &dw_ext[b], n_slots, s->n_sms, count_ptr, perm, st, exact_gate);
                s->launches++;
            } else {
                launch_extend(s->dev, prune, false, qin, &q_count[b], &w_ext[b], perm, 0, n_slots, s->n_sms, st);
            }
            s->span_end(sp, st);
            const uint32_t n_fused = (fused_camera && b == 0) ? n_slots : 0u;
            if (lane) {   // fork: the deferred rays of this bounce are traced exactly and shaded on the side stream ...
                CU_CHECK(cudaEventRecord(s->ev_fork, st));
                CU_CHECK(cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0));
                launch_deferred_lane(s->dev, rd, d_tiles, b, batch_first, qin, defer_list, &d_ext[b], &dw_ext[b], side_hit, qout, &q_count[b + 1], sq, &s_count[b],
                                     Lo, Le, n_fused, (uint32_t)nf, s->n_sms, s->side_stream);
                CU_CHECK(cudaEventRecord(s->ev_join, s->side_stream));
                s->launches += 2;
            }
            sp = s->span_begin(1, st);
            launch_shade(s->dev, rd, d_tiles, b, batch_first, qin, &q_count[b], qout, &q_count[b + 1], sq, &s_count[b], Lo, Le,
                         n_slots, n_fused, (uint32_t)nf, s->n_sms, st);
            if (lane) CU_CHECK(cudaStreamWaitEvent(st, s->ev_join, 0));   // ... while this k_shade shades all the others; join
            s->span_end(sp, st);
            s->launches += 2;
            if (is_mode && b < p->max_bounce) {
                sp = s->span_begin(2, st);
                if (accel) {
                    launch_shadow_accel(s->dev, sq, &s_count[b], &w_sh[b], Lo, defer_list, &d_sh[b], &dw_sh[b], n_slots, s->n_sms, count_ptr, st);
                    s->launches++;
                } else {
                    launch_shadow(s->dev, prune, sq, &s_count[b], &w_sh[b], Lo, nullptr, n_slots, s->n_sms, st);
                }
                s->span_end(sp, st);
                sp = s->span_begin(1, st);   // shading work: counted with k_shade
                launch_nee(s->dev, rd, sq, &s_count[b], Lo, n_slots, s->n_sms, st);
                s->span_end(sp, st);
                s->launches += 2;
            }
        }
        sp = s->span_begin(3, st);
        if (s->fb_wait) {   // ezrt_render: lastFrame arrives on the copy stream
            CU_CHECK(cudaStreamWaitEvent(st, s->fb_wait, 0));
            s->fb_wait = nullptr;
        }
        launch_blend(rd, d_tiles, nf, batch_first, Lo, Le, d_fb, st);
        launch_tally(q_count, s_count, d_ext, d_sh, p->max_bounce + 1, totals, fused_camera ? (uint32_t)(s->n_pixels * (size_t)nf) : 0u, st);
        s->span_end(sp, st);
        s->launches += 2;
    }
    if (l2_window) {  // leave the caller's stream as it was
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
    }
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaEventRecord(s->ev_stop, st));
    s->have_timing = true;
    return EZRT_OK;
}

int ezrt_render(ezrt_scene* s, const ezrt_render_params* p, float* framebuffer) {
    int rc = validate_params(s, p);
    if (rc) return rc;
    if (!framebuffer) return ezrt_set_error(EZRT_ERR_INVALID, "render: null framebuffer");
    CU_CHECK(cudaSetDevice(s->device));
    int64_t npix = ezrt_partition_pixels(p->width, p->height, p->part_rank, p->part_count);
    size_t bytes = sizeof(float) * (size_t)npix * p->out_channels;
    rc = s->fb_buf.ensure(std::max<size_t>(bytes, 16));
    if (rc) return rc;
    cudaStream_t st = s->own_stream;
    // lastFrame is only needed by the first k_blend: its upload runs on a second stream, under the tracing kernels
    s->fb_wait = nullptr;
    static const bool overlap_upload = []() { const char* e = getenv("EZRT_RENDER_OVERLAP"); return !(e && atoi(e) == 0); }();
    if (p->first_frame > 0 && !overlap_upload) {
        CU_CHECK(cudaMemcpyAsync(s->fb_buf.p, framebuffer, bytes, cudaMemcpyHostToDevice, st));
    } else if (p->first_frame > 0) {
        if (!s->copy_stream && cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return ezrt_set_error(EZRT_ERR_CUDA, "render: stream");
        if (!s->fb_event && cudaEventCreateWithFlags(&s->fb_event, cudaEventDisableTiming) != cudaSuccess) return ezrt_set_error(EZRT_ERR_CUDA, "render: event");
        CU_CHECK(cudaMemcpyAsync(s->fb_buf.p, framebuffer, bytes, cudaMemcpyHostToDevice, s->copy_stream));
        CU_CHECK(cudaEventRecord(s->fb_event, s->copy_stream));
        s->fb_wait = s->fb_event;
    }
    rc = ezrt_render_device(s, p, (float*)s->fb_buf.p, st);
    if (s->fb_wait) {  // the render returned before its first blend (error, empty part): do not leave the copy behind
        cudaStreamWaitEvent(st, s->fb_wait, 0);
        s->fb_wait = nullptr;
    }
    if (rc) return rc;
    CU_CHECK(cudaMemcpyAsync(framebuffer, s->fb_buf.p, bytes, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    return EZRT_OK;
}

int ezrt_get_counters(ezrt_scene* s, ezrt_counters* out) {
    if (!s || !out) return ezrt_set_error(EZRT_ERR_INVALID, "get_counters: null argument");
    memset(out, 0, sizeof(*out));
    if (!s->have_timing) return EZRT_OK;
    CU_CHECK(cudaSetDevice(s->device));
    CU_CHECK(cudaEventSynchronize(s->ev_stop));
    unsigned long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CU_CHECK(cudaMemcpy(t, s->totals_buf.p, sizeof(t), cudaMemcpyDeviceToHost));
    out->deferred_rays = t[4];
    out->node_visits = t[5] + t[7];          // t[5]: visits of 128-byte exact nodes, t[7]: of 96-byte nodes (Q16 form, W8)
    out->tri_tests = t[6];
    out->node_bytes = t[5] * 128ull + t[7] * 96ull;
    out->tri_bytes = t[6] * 64ull;
    out->node_visits_96 = t[7];
    float ms = 0.0f;
    CU_CHECK(cudaEventElapsedTime(&ms, s->ev_start, s->ev_stop));
    out->primary_rays = t[0]; out->bounce_rays = t[1]; out->shadow_rays = t[2];
    out->rays = t[0] + t[1] + t[2];
    out->samples = t[3];
    out->kernel_launches = s->launches;
    out->device_ms = ms;
    return EZRT_OK;
}

int ezrt_get_kernel_times(ezrt_scene* s, double* ms, uint64_t* launches) {
    if (!s || !ms || !launches) return ezrt_set_error(EZRT_ERR_INVALID, "get_kernel_times: null argument");
    for (int k = 0; k < 4; k++) { ms[k] = 0.0; launches[k] = 0; }
    if (!s->have_timing || s->spans.empty()) return EZRT_OK;
    CU_CHECK(cudaSetDevice(s->device));
    CU_CHECK(cudaEventSynchronize(s->ev_stop));
    for (const auto& sp : s->spans) {
        float t = 0.0f;
        CU_CHECK(cudaEventElapsedTime(&t, s->ev_pool[sp.e0], s->ev_pool[sp.e1]));
        ms[sp.cls] += t;
        launches[sp.cls] += 1;
    }
    return EZRT_OK;
}

// ------------------------------------------------------------------------------------------
// image partition
// ------------------------------------------------------------------------------------------
int64_t ezrt_partition_pixels(int width, int height, int rank, int count) {
    if (width <= 0 || height <= 0 || count < 1 || rank < 0 || rank >= count) return EZRT_ERR_INVALID;
    int64_t n = 0;
    for (const TileDev& t : partition_tiles(width, height, rank, count)) n += (int64_t)t.w * t.h;
    return n;
}

int ezrt_partition_scatter(const float* d_compact, float* d_full, int width, int height, int channels, int rank, int count,
                           void* cuda_stream) {
    if (!d_compact || !d_full || channels < 1) return ezrt_set_error(EZRT_ERR_INVALID, "partition_scatter: bad argument");
    if (width <= 0 || height <= 0 || count < 1 || rank < 0 || rank >= count) return ezrt_set_error(EZRT_ERR_INVALID, "partition_scatter: bad partition");
    // device tile lists are cached per (device, image, part): the gather runs once per render
    auto& cache = scatter_cache();
    std::mutex& mu = scatter_mutex();
    int device = 0;
    CU_CHECK(cudaGetDevice(&device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    ScatterEntry ent;
    {
        std::lock_guard<std::mutex> lock(mu);
        std::array<int, 5> key = {device, width, height, rank, count};
        auto it = cache.find(key);
        if (it == cache.end()) {
            std::vector<TileDev> tiles = partition_tiles(width, height, rank, count);
            ScatterEntry e{nullptr, (int)tiles.size()};
            if (!tiles.empty()) {
                CU_CHECK(cudaMalloc(&e.d_tiles, sizeof(TileDev) * tiles.size()));
                CU_CHECK(cudaMemcpy(e.d_tiles, tiles.data(), sizeof(TileDev) * tiles.size(), cudaMemcpyHostToDevice));
            }
            it = cache.emplace(key, e).first;
        }
        ent = it->second;
    }
    if (ent.n == 0) return EZRT_OK;
    launch_partition_scatter(d_compact, d_full, ent.d_tiles, ent.n, width, channels, st);
    CU_CHECK(cudaGetLastError());
    return EZRT_OK;
}

int ezrt_partition_cache_clear(int device) {
    std::lock_guard<std::mutex> lock(scatter_mutex());
    auto& cache = scatter_cache();
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto it = cache.begin(); it != cache.end();) {
        if (device < 0 || it->first[0] == device) {
            if (it->second.d_tiles) {
                cudaSetDevice(it->first[0]);
                cudaFree(it->second.d_tiles);
            }
            it = cache.erase(it);
        } else {
            ++it;
        }
    }
    cudaSetDevice(prev);
    cudaGetLastError();
    return EZRT_OK;
}

int ezrt_partition_scatter_host(const float* compact, float* full, int width, int height, int channels, int rank, int count) {
    if (!compact || !full || channels < 1) return ezrt_set_error(EZRT_ERR_INVALID, "partition_scatter_host: bad argument");
    for (const TileDev& t : partition_tiles(width, height, rank, count))
        for (int iy = 0; iy < t.h; iy++)
            for (int ix = 0; ix < t.w; ix++) {
                size_t src = ((size_t)t.pixel_offset + (size_t)iy * t.w + ix) * channels;
                size_t dst = ((size_t)(t.y0 + iy) * width + (t.x0 + ix)) * channels;
                for (int c = 0; c < channels; c++) full[dst + c] = compact[src + c];
            }
    return EZRT_OK;
}

// ------------------------------------------------------------------------------------------
// post pass
// ------------------------------------------------------------------------------------------
int ezrt_post_tonemap(const float* d_in, int channels, float* d_out, int64_t n_pixels, float limit, void* cuda_stream) {
    if (!d_in || !d_out || (channels != 3 && channels != 4) || n_pixels < 0) return ezrt_set_error(EZRT_ERR_INVALID, "post_tonemap: bad argument");
    launch_tonemap(d_in, channels, d_out, (long long)n_pixels, limit, (cudaStream_t)cuda_stream);
    CU_CHECK(cudaGetLastError());
    return EZRT_OK;
}

// ------------------------------------------------------------------------------------------
// single-function entry points (parity tests)
// ------------------------------------------------------------------------------------------
int ezrt_trace_rays(ezrt_scene* s, int n, const float* origins, const float* dirs, int traverse, int any_hit, int p3_normal_fudge,
                    int32_t* out_hit, float* out_distance, int32_t* out_triangle, int32_t* out_inside, float* out_point,
                    float* out_normal) {
    if (!s || n < 0 || !origins || !dirs || !out_hit || !out_distance || !out_triangle || !out_inside || !out_point || !out_normal)
        return ezrt_set_error(EZRT_ERR_INVALID, "trace_rays: null argument");
    if (n == 0) return EZRT_OK;
    CU_CHECK(cudaSetDevice(s->device));
    // the rays go through the production path: a ray queue traced by the persistent extend kernels
    std::vector<float4> ho(n), hd(n);
    for (int i = 0; i < n; i++) {
        ho[i] = make_float4(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2], 0.0f);
        hd[i] = make_float4(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], 0.0f);
    }
    DeviceBuffer buf;
    const size_t N = (size_t)n;
    int rc = buf.ensure(sizeof(float4) * 2 * N + sizeof(float2) * N + sizeof(float) * 7 * N + sizeof(int) * 4 * N + 1024);
    if (rc) return rc;
    char* p = (char*)buf.p;
    PathQueue q{};
    q.ray_o = (float4*)p; p += sizeof(float4) * N;
    q.ray_d = (float4*)p; p += sizeof(float4) * N;
    q.hit = (float2*)p; p += sizeof(float2) * N;
    float* d_point = (float*)p; p += sizeof(float) * 3 * N;
    float* d_normal = (float*)p; p += sizeof(float) * 3 * N;
    float* d_dist = (float*)p; p += sizeof(float) * N;
    int* d_hit = (int*)p; p += sizeof(int) * N;
    int* d_tri = (int*)p; p += sizeof(int) * N;
    int* d_inside = (int*)p; p += sizeof(int) * N;
    uint32_t* d_defer = (uint32_t*)p; p += sizeof(uint32_t) * N;
    uint32_t* d_cnt = (uint32_t*)p;  // [0] n, [1] work, [2] deferred, [3] deferred work
    if (!s->regular_tree) traverse = EZRT_TRAVERSE_REFERENCE;
    else if (traverse == EZRT_TRAVERSE_ACCEL && (!s->have_accel || any_hit)) traverse = EZRT_TRAVERSE_PRUNED;  // any-hit probes take the exact kernel
    cudaStream_t st = s->own_stream;
    const uint32_t counters[4] = {(uint32_t)n, 0u, 0u, 0u};
    cudaError_t e = cudaMemcpyAsync(q.ray_o, ho.data(), sizeof(float4) * N, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q.ray_d, hd.data(), sizeof(float4) * N, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cnt, counters, sizeof(counters), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        if (traverse == EZRT_TRAVERSE_ACCEL)
            launch_extend_accel(s->dev, q, d_cnt, d_cnt + 1, d_defer, d_cnt + 2, d_cnt + 3, (uint32_t)n, s->n_sms, nullptr, nullptr, st);
        else
            launch_extend(s->dev, traverse != EZRT_TRAVERSE_REFERENCE, any_hit != 0, q, d_cnt, d_cnt + 1, nullptr, 0, (uint32_t)n, s->n_sms, st);
        launch_trace_finish(s->dev, n, q, p3_normal_fudge, traverse == EZRT_TRAVERSE_ACCEL, d_hit, d_dist, d_tri, d_inside, d_point, d_normal, st);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_hit, d_hit, sizeof(int) * N, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_distance, d_dist, sizeof(float) * N, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_triangle, d_tri, sizeof(int) * N, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_inside, d_inside, sizeof(int) * N, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_point, d_point, sizeof(float) * 3 * N, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_normal, d_normal, sizeof(float) * 3 * N, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    buf.release();
    if (e != cudaSuccess) return ezrt_set_error(EZRT_ERR_CUDA, "trace_rays: %s", cudaGetErrorString(e));
    return EZRT_OK;
}

int ezrt_eval_brdf(int device, int which, int n, const float* V, const float* N, const float* L, const float* xi,
                   const float* materials, float* out) {
    if (n < 0 || !V || !N || !materials || !out || which < 0 || which > 3) return ezrt_set_error(EZRT_ERR_INVALID, "eval_brdf: bad argument");
    if (which != 3 && !L) return ezrt_set_error(EZRT_ERR_INVALID, "eval_brdf: L required");
    if (which == 3 && !xi) return ezrt_set_error(EZRT_ERR_INVALID, "eval_brdf: xi required");
    if (n == 0) return EZRT_OK;
    CU_CHECK(cudaSetDevice(device));
    DeviceBuffer buf;
    size_t f3 = sizeof(float) * 3 * (size_t)n;
    int rc = buf.ensure(f3 * 5 + sizeof(float) * 18 * (size_t)n + 256);
    if (rc) return rc;
    float* dV = (float*)buf.p;
    float* dN = dV + 3 * (size_t)n;
    float* dL = dN + 3 * (size_t)n;
    float* dXi = dL + 3 * (size_t)n;
    float* dOut = dXi + 3 * (size_t)n;
    float* dM = dOut + 3 * (size_t)n;
    cudaError_t e = cudaMemcpy(dV, V, f3, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dN, N, f3, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && L) e = cudaMemcpy(dL, L, f3, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && xi) e = cudaMemcpy(dXi, xi, f3, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dM, materials, sizeof(float) * 18 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        launch_eval_brdf(which, n, dV, dN, L ? dL : nullptr, xi ? dXi : nullptr, dM, dOut, 0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dOut, f3, cudaMemcpyDeviceToHost);
    buf.release();
    if (e != cudaSuccess) return ezrt_set_error(EZRT_ERR_CUDA, "eval_brdf: %s", cudaGetErrorString(e));
    return EZRT_OK;
}

int ezrt_accel_build(int device, const float* tris, int n_triangles, int leaf_n, int where, int32_t* links_out, float* boxes_out,
                     int nodes_cap, uint32_t* order_out, double* ms) {
    if (!tris || n_triangles <= 0 || leaf_n < 1) return ezrt_set_error(EZRT_ERR_INVALID, "accel_build: bad argument");
    std::vector<EzrtAccelNode> an;
    std::vector<uint32_t> order;
    const auto t0 = std::chrono::steady_clock::now();
    int n_nodes = 0;
    if (where == 1) {
        n_nodes = ezrt_build_accel(tris, n_triangles, leaf_n, an, order);
    } else {
        int n_dev = 0;
        CU_CHECK(cudaGetDeviceCount(&n_dev));
        if (device < 0 || device >= n_dev) return ezrt_set_error(EZRT_ERR_CUDA, "accel_build: no CUDA device %d (have %d)", device, n_dev);
        CU_CHECK(cudaSetDevice(device));
        DeviceBuffer raw;
        int rc = raw.ensure((size_t)n_triangles * EZRT_TRIANGLE_FLOATS * sizeof(float));
        if (rc) return rc;
        if (cudaMemcpy(raw.p, tris, (size_t)n_triangles * EZRT_TRIANGLE_FLOATS * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            raw.release();
            return ezrt_set_error(EZRT_ERR_CUDA, "accel_build: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        n_nodes = ezrt_build_accel_device((const float*)raw.p, n_triangles, leaf_n, an, order, nullptr);
        raw.release();
    }
    if (n_nodes < 0) return n_nodes;
    if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if ((links_out || boxes_out) && nodes_cap < n_nodes) return ezrt_set_error(EZRT_ERR_INVALID, "accel_build: %d nodes, room for %d", n_nodes, nodes_cap);
    for (int i = 0; i < n_nodes; i++) {
        if (links_out) { links_out[4 * i] = an[i].left; links_out[4 * i + 1] = an[i].right; links_out[4 * i + 2] = an[i].n; links_out[4 * i + 3] = an[i].index; }
        if (boxes_out)
            for (int k = 0; k < 3; k++) { boxes_out[6 * i + k] = an[i].AA[k]; boxes_out[6 * i + 3 + k] = an[i].BB[k]; }
    }
    if (order_out) memcpy(order_out, order.data(), sizeof(uint32_t) * (size_t)n_triangles);
    return n_nodes;
}

int ezrt_eval_math(int device, int which, int n, const float* a, const float* b, float* out) {
    if (n < 0 || !a || !out || which < 0 || which > 6) return ezrt_set_error(EZRT_ERR_INVALID, "eval_math: bad argument");
    if (n == 0) return EZRT_OK;
    CU_CHECK(cudaSetDevice(device));
    DeviceBuffer buf;
    size_t fN = sizeof(float) * (size_t)n;
    int rc = buf.ensure(fN * 3 + 256);
    if (rc) return rc;
    float* dA = (float*)buf.p;
    float* dB = dA + n;
    float* dO = dB + n;
    cudaError_t e = cudaMemcpy(dA, a, fN, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && b) e = cudaMemcpy(dB, b, fN, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        launch_eval_math(which, n, dA, b ? dB : nullptr, dO, 0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dO, fN, cudaMemcpyDeviceToHost);
    buf.release();
    if (e != cudaSuccess) return ezrt_set_error(EZRT_ERR_CUDA, "eval_math: %s", cudaGetErrorString(e));
    return EZRT_OK;
}

}  // extern "C"
