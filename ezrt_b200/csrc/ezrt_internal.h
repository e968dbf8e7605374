// ezrt_internal.h -- shared between the translation units of libezrt_b200.so (not installed).
#ifndef EZRT_INTERNAL_H
#define EZRT_INTERNAL_H

#include "ezrt.h"

#ifdef __cplusplus
extern "C" {
#endif

// records a thread-local message for ezrt_last_error() and returns `code`
int ezrt_set_error(int code, const char* fmt, ...)
#if defined(__GNUC__)
    __attribute__((format(printf, 2, 3)))
#endif
    ;

#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
#include <stdint.h>

#include <vector>
struct EzrtAccelNode {
    int left, right, n, index;  // children (0 = none; the root is node 0 and never a child), leaf range
    float AA[3], BB[3];
};
// host_scene.cpp: sentinel-free SAH tree over the triangles of a Triangle_encoded array
int ezrt_build_accel(const float* tris, int n_tris, int leaf_n, std::vector<EzrtAccelNode>& nodes, std::vector<uint32_t>& order);
#endif

// Image partition shared by host and device code (ezrt_render_params.part_rank/part_count):
// 16x16 tiles, tile (tx,ty) -> part (tx+ty) % count; a part stores its tiles in row-major
// tile order, pixels row-major inside a tile (edge tiles are clipped).
#define EZRT_PART_TILE 16

#endif
