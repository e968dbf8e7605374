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

// accel_build.cu: the same tree, node for node, built on the GPU from the Triangle_encoded array in device memory (the
// current device).  Returns the node count or a negative status; *levels = depth of the tree.
int ezrt_build_accel_device(const float* d_tris, int n_tris, int leaf_n, std::vector<EzrtAccelNode>& nodes, std::vector<uint32_t>& order,
                            int* levels);

// accel_w8.cpp: SAH-optimal collapse of the binary tree to `width`-wide nodes (dynamic programming; shared by the 4-wide
// exact-box form and the 8-wide quantised form)
struct EzrtCollapse {
    const std::vector<EzrtAccelNode>* an = nullptr;
    int width = 0;
    std::vector<int> first, count;   // per binary node: first triangle / triangle count of its sub-tree (in the binary tree's order)
    std::vector<char> as_leaf;       // per binary node: as a child of a wide node it is ONE leaf (all its triangles)
    std::vector<float> C;            // C[n * 8 + i]: cost of representing n's sub-tree by <= i roots
    int build(const std::vector<EzrtAccelNode>& an, int width, int max_leaf, double cost_node, double cost_tri);
    int collect(int n0, int k0, int* roots) const;
    int children(int b, int* ch) const;   // b: an inner binary node that is not as_leaf; returns the child count (<= width)
};

// accel_w8.cpp: the same tree collapsed to 8-wide nodes with 8-bit quantised child boxes (w8_node.h)
struct EzrtW8Tree {
    std::vector<uint32_t> nodes;       // W8_NODE_WORDS words per node; node 0 = root; breadth-first numbering
    std::vector<uint32_t> tri_order;   // triangle i of the tree = tri_order[i] of the caller's triangle array
    std::vector<int> leaf_first;       // per binary node: first triangle (new order) of that leaf, -1 for inner nodes
    int depth;                         // levels of 8-wide nodes
    int n_nodes;
    long long n_children;              // occupied slots over all nodes (fill statistics)
};
// `order_in`: triangle order of the binary tree (ezrt_build_accel); `pad`: box inflation (2 * prune_delta);
// axis_bit[a]: significance (0..2) of axis a in the slot index (largest scene extent -> bit 2).  Returns 0 or < 0.
void ezrt_w8_axis_bits(const float bmin[3], const float bmax[3], int axis_bit[3]);
int ezrt_build_w8(const std::vector<EzrtAccelNode>& an, const std::vector<uint32_t>& order_in, float pad, float max_abs_coord,
                  const int axis_bit[3], EzrtW8Tree& out);
#endif

// Image partition shared by host and device code (ezrt_render_params.part_rank/part_count):
// 16x16 tiles, tile (tx,ty) -> part (tx+ty) % count; a part stores its tiles in row-major
// tile order, pixels row-major inside a tile (edge tiles are clipped).
#define EZRT_PART_TILE 16

#endif
