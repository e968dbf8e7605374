// ezrt_main.cpp -- the reference's main() + display() loop as a C++ host over the C ABI (include/ezrt.h).
//
// What P5/main.cpp does with OpenGL, this does with libezrt_b200.so:
//   main()     :759-871  readObj x N, buildBVHwithSAH, encode            -> ezrt_trilist_* (or a scene file)
//              :873-906  texture buffers, HDR map, calculateHdrCache     -> ezrt_hdr_load / ezrt_hdr_cache / ezrt_scene_create
//   display()  :697-748  camera, frameCounter++, pass1 / pass2 / pass3   -> ezrt_camera_orbit, ezrt_render (spp frames), ezrt_write_png
// No GL, no window: the accumulated frame is tone-mapped (pass3) and written as a PNG.
//
//   g++ -O2 -std=c++17 -Iinclude examples/ezrt_main.cpp -Lezrt_b200 -lezrt_b200 -Wl,-rpath,'$ORIGIN/../ezrt_b200' -o examples/ezrt_main
//   examples/ezrt_main --p5 "<reference>/part 5 .../source code" out.png [--size 512 512] [--spp 64] [--mode 3] [--bounces 2]
//   examples/ezrt_main --scene scene.txt out.png ...
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "ezrt.h"

static void die(const char* what) {
    fprintf(stderr, "ezrt_main: %s: %s\n", what, ezrt_last_error());
    exit(1);
}
#define CHECK(call, what) do { if ((call) < 0) die(what); } while (0)

int main(int argc, char** argv) {
    std::string p5_dir, scene_file, out_png;
    int width = 512, height = 512, spp = 16, mode = EZRT_MODE_DISNEY_IS_MIS_P5, bounces = 2, traverse = EZRT_TRAVERSE_ACCEL, device = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--p5" && i + 1 < argc) p5_dir = argv[++i];
        else if (a == "--scene" && i + 1 < argc) scene_file = argv[++i];
        else if (a == "--size" && i + 2 < argc) { width = atoi(argv[++i]); height = atoi(argv[++i]); }
        else if (a == "--spp" && i + 1 < argc) spp = atoi(argv[++i]);
        else if (a == "--mode" && i + 1 < argc) mode = atoi(argv[++i]);
        else if (a == "--bounces" && i + 1 < argc) bounces = atoi(argv[++i]);
        else if (a == "--traverse" && i + 1 < argc) traverse = atoi(argv[++i]);
        else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else out_png = a;
    }
    if (out_png.empty() || (p5_dir.empty() == scene_file.empty())) {
        fprintf(stderr, "usage: ezrt_main (--p5 <P5 source dir> | --scene <scene file>) out.png [--size W H] [--spp N] [--mode 0..3] [--bounces B] [--traverse 0..2]\n");
        return 2;
    }

    // ---- scene (P5/main.cpp:768-800)
    ezrt_trilist* list = ezrt_trilist_create();
    float camera[3] = {90.0f, 10.0f, 2.0f};  // rotatAngle, upAngle, r (:764-766)
    std::string hdr_path;
    if (!p5_dir.empty()) {
        const float zero[3] = {0, 0, 0};
        float trans[16];
        // Material m; m.roughness = 0.5; m.specular = 1.0; m.metallic = 1.0; m.clearcoat = 1.0; m.clearcoatGloss = 0.0; m.baseColor = (1, 0.73, 0.25)
        float m[EZRT_MATERIAL_FLOATS] = {0, 0, 0, 1.0f, 0.73f, 0.25f, 0.0f, 1.0f, 1.0f, 0.0f, 0.5f, 0.0f, 0.0f, 0.5f, 1.0f, 0.0f, 1.0f, 0.0f};
        const float t0[3] = {0, -0.5f, 0}, s0[3] = {0.75f, 0.75f, 0.75f};
        ezrt_transform_matrix(zero, t0, s0, trans);
        CHECK(ezrt_trilist_read_obj(list, (p5_dir + "/models/teapot.obj").c_str(), m, trans, 1), "readObj teapot");
        m[10] = 0.01f; m[7] = 0.1f; m[8] = 1.0f; m[3] = m[4] = m[5] = 1.0f;  // roughness, metallic, specular, baseColor (:778-781)
        const float len = 13000.0f, s1[3] = {len, 0.01f, len};
        ezrt_transform_matrix(zero, t0, s1, trans);
        CHECK(ezrt_trilist_read_obj(list, (p5_dir + "/models/quad.obj").c_str(), m, trans, 0), "readObj quad");
        hdr_path = p5_dir + "/HDR/chinese_garden_2k.hdr";
    } else {
        char hp[4096];
        CHECK(ezrt_scene_file_load(scene_file.c_str(), list, camera, hp, sizeof(hp)), "scene file");
        hdr_path = hp;
    }
    const int n_tris = ezrt_trilist_size(list);
    auto t_build = std::chrono::steady_clock::now();
    const int n_nodes = ezrt_trilist_build_bvh(list, 8, EZRT_BVH_SAH_FAST);  // buildBVHwithSAH(triangles, nodes, 0, n-1, 8) (:799)
    if (n_nodes < 0) die("buildBVHwithSAH");
    const double build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_build).count();
    std::vector<float> tris((size_t)n_tris * EZRT_TRIANGLE_FLOATS), nodes((size_t)n_nodes * EZRT_BVHNODE_FLOATS);
    CHECK(ezrt_trilist_encode_triangles(list, tris.data()), "encode triangles");
    CHECK(ezrt_trilist_encode_nodes(list, nodes.data()), "encode nodes");
    ezrt_trilist_destroy(list);
    printf("scene: %d triangles, %d BVH nodes (built in %.2f s)\n", n_tris, n_nodes, build_s);

    // ---- environment (:891-905)
    std::vector<float> hdr, cache;
    int hw = 0, hh = 0;
    if (!hdr_path.empty()) {
        CHECK(ezrt_hdr_load(hdr_path.c_str(), &hw, &hh, nullptr), "HDRLoader::load");
        hdr.resize((size_t)hw * hh * 3);
        cache.resize(hdr.size());
        CHECK(ezrt_hdr_load(hdr_path.c_str(), &hw, &hh, hdr.data()), "HDRLoader::load");
        CHECK(ezrt_hdr_cache_device(device, hdr.data(), hw, hh, cache.data(), nullptr), "calculateHdrCache");
        printf("environment: %d x %d\n", hw, hh);
    } else if (mode == EZRT_MODE_DISNEY_IS_MIS_P5) {
        mode = EZRT_MODE_DISNEY_SOBOL_P5;  // no map to importance-sample
    }

    ezrt_scene* scene = nullptr;
    CHECK(ezrt_scene_create(device, tris.data(), n_tris, nodes.data(), n_nodes, hdr.empty() ? nullptr : hdr.data(),
                            cache.empty() ? nullptr : cache.data(), hw, hh, /*linear filter*/ 1, &scene), "scene_create");

    // ---- display() x spp (:697-748)
    ezrt_render_params p;
    memset(&p, 0, sizeof(p));
    p.width = width; p.height = height; p.spp = spp; p.first_frame = 0; p.max_bounce = bounces; p.mode = mode;
    p.traverse = traverse; p.pipeline = EZRT_PIPELINE_WAVEFRONT; p.out_channels = 3; p.part_rank = 0; p.part_count = 1;
    ezrt_camera_orbit(camera[0], camera[1], camera[2], p.eye, p.camera_rotate);
    std::vector<float> fb((size_t)width * height * 3);
    CHECK(ezrt_render(scene, &p, fb.data()), "render");
    ezrt_counters c;
    CHECK(ezrt_get_counters(scene, &c), "counters");
    printf("%d frames of %d x %d: %llu rays in %.1f ms on the device = %.0f Mrays/s (%llu kernel launches)\n", spp, width, height,
           (unsigned long long)c.rays, c.device_ms, c.rays / (c.device_ms * 1e3), (unsigned long long)c.kernel_launches);
    CHECK(ezrt_write_png(out_png.c_str(), fb.data(), width, height, 3, /*pass3 tone map*/ 1), "write_png");
    printf("wrote %s\n", out_png.c_str());
    ezrt_scene_destroy(scene);
    return 0;
}
