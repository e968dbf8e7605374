// accel_w8.cpp -- host builder of the 8-wide quantised acceleration tree (layout and rationale: w8_node.h).
// Input is the binary sentinel-free SAH tree of ezrt_build_accel (host_scene.cpp); this file collapses it to
// 8-wide nodes, puts the children into octant-ordered slots, renumbers the triangles so that the leaf children
// of a node are consecutive, and quantises the child boxes conservatively.  The reference has no counterpart
// (its hitBVH walks the 48-byte binary nodes, P5/fsh:254-306); results stay the reference's through the
// deferral rule of the accel policy.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "ezrt_internal.h"
#include "w8_node.h"

namespace {

inline float box_area(const EzrtAccelNode& c) {
    const float x = c.BB[0] - c.AA[0], y = c.BB[1] - c.AA[1], z = c.BB[2] - c.AA[2];
    return x * y + x * z + y * z;
}

}  // namespace

// Significance of the axes in the slot index: the axis of largest scene extent decides first (bit 2).
void ezrt_w8_axis_bits(const float bmin[3], const float bmax[3], int axis_bit[3]) {
    int idx[3] = {0, 1, 2};
    float ext[3] = {bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2]};
    std::stable_sort(idx, idx + 3, [&](int a, int b) { return ext[a] < ext[b]; });  // ascending extent
    for (int k = 0; k < 3; k++) axis_bit[idx[k]] = k;
}

// ------------------------------------------------------------------------------------------
// Optimal collapse of the binary tree to `width`-wide nodes by dynamic programming over the binary tree (surface-area
// heuristic, after Ylitie et al. 2017 section 3.1; children have larger indices than their parent):
//   C(n,1) = min(cost of n as ONE leaf (<= max_leaf triangles, consecutive), area(n) * cost_node + best split of n into <= width roots)
//   C(n,i) = min(C(n,i-1), min_k C(left,k) + C(right,i-k))          i = 2..width-1: n's sub-tree represented by <= i roots
// Visiting a wide node costs cost_node, testing one triangle cost_tri (ratio of the kernels' instruction counts).
// ------------------------------------------------------------------------------------------
int EzrtCollapse::build(const std::vector<EzrtAccelNode>& an_, int width_, int max_leaf, double cost_node, double cost_tri) {
    an = &an_;
    width = width_;
    const int NB = (int)an_.size();
    first.assign(NB, 0);
    count.assign(NB, 0);
    as_leaf.assign(NB, 0);
    C.assign((size_t)NB * 8, 0.0f);
    if (width < 2 || width > 8) return -1;
    for (int i = NB - 1; i >= 0; i--) {
        const EzrtAccelNode& nd = an_[i];
        const double area = (double)box_area(nd);
        float* c = &C[(size_t)i * 8];
        if (nd.n > 0) {
            if (nd.n > max_leaf) return -3;
            first[i] = nd.index;
            count[i] = nd.n;
            as_leaf[i] = 1;
            for (int k = 1; k <= 7; k++) c[k] = (float)(area * nd.n * cost_tri);
            continue;
        }
        const float *cl = &C[(size_t)nd.left * 8], *cr = &C[(size_t)nd.right * 8];
        first[i] = std::min(first[nd.left], first[nd.right]);
        count[i] = count[nd.left] + count[nd.right];
        float dist[9];
        for (int j = 2; j <= width; j++) {
            float best = 3.0e38f;
            for (int k = 1; k < j; k++)
                if (k <= width - 1 && j - k <= width - 1) best = std::min(best, cl[k] + cr[j - k]);
            dist[j] = best;
        }
        const float c_internal = (float)(area * cost_node) + dist[width];
        const float c_leaf = (count[i] <= max_leaf && first[nd.left] + count[nd.left] == first[nd.right]) ? (float)(area * count[i] * cost_tri) : 3.0e38f;
        as_leaf[i] = c_leaf <= c_internal;
        c[1] = std::min(c_leaf, c_internal);
        for (int j = 2; j <= width - 1; j++) c[j] = std::min(c[j - 1], dist[j]);
        for (int j = width; j <= 7; j++) c[j] = c[width - 1];
    }
    return 0;
}

// roots of the best representation of sub-tree n0 by <= k0 roots (left to right)
int EzrtCollapse::collect(int n0, int k0, int* roots) const {
    const std::vector<EzrtAccelNode>& a = *an;
    struct Pick { int n, k; };
    int cnt = 0;
    Pick stack[32];
    int sp = 0;
    stack[sp++] = {n0, k0};
    while (sp > 0) {
        Pick p = stack[--sp];
        const float* c = &C[(size_t)p.n * 8];
        while (p.k > 1 && c[p.k] == c[p.k - 1]) p.k--;
        if (p.k == 1 || a[p.n].n > 0) { roots[cnt++] = p.n; continue; }
        const EzrtAccelNode& nd = a[p.n];
        const float *cl = &C[(size_t)nd.left * 8], *cr = &C[(size_t)nd.right * 8];
        int bk = 1;
        float best = 3.0e38f;
        for (int k = 1; k < p.k; k++)
            if (cl[k] + cr[p.k - k] < best) { best = cl[k] + cr[p.k - k]; bk = k; }
        stack[sp++] = {nd.right, p.k - bk};   // right first on the stack: the left roots come out first
        stack[sp++] = {nd.left, bk};
    }
    return cnt;
}

// children (binary nodes; as_leaf[] says which become leaves) of the wide node made from inner binary node b
int EzrtCollapse::children(int b, int* ch) const {
    const std::vector<EzrtAccelNode>& a = *an;
    const EzrtAccelNode& nd = a[b];
    const float *cl = &C[(size_t)nd.left * 8], *cr = &C[(size_t)nd.right * 8];
    int bk = 1;
    float best = 3.0e38f;
    for (int k = 1; k <= width - 1; k++)
        if (cl[k] + cr[width - k] < best) { best = cl[k] + cr[width - k]; bk = k; }
    int cnt = collect(nd.left, bk, ch);
    cnt += collect(nd.right, width - bk, ch + cnt);
    return cnt;
}

#define W8_COST_NODE 1.0
#define W8_COST_TRI 0.4

int ezrt_build_w8(const std::vector<EzrtAccelNode>& an, const std::vector<uint32_t>& order_in, float pad, float max_abs_coord,
                  const int axis_bit[3], EzrtW8Tree& out) {
    out.nodes.clear();
    out.tri_order.clear();
    out.leaf_first.assign(an.size(), -1);
    out.depth = 0;
    out.n_nodes = 0;
    out.n_children = 0;
    if (an.empty()) return -1;
    const int NB = (int)an.size();
    const double min_step = (double)max_abs_coord * (double)W8_MIN_STEP_REL;
    EzrtCollapse col;
    const int crc = col.build(an, 8, W8_MAX_LEAF_TRIS, W8_COST_NODE, W8_COST_TRI);
    if (crc) return crc;
    const std::vector<int>& first = col.first;
    const std::vector<int>& count = col.count;
    const std::vector<char>& as_leaf = col.as_leaf;

    struct Item { int bnode, level; };
    std::vector<Item> queue;   // wide node i is built from queue[i] (breadth-first: a node's inner children are consecutive)
    queue.push_back({0, 1});
    out.tri_order.reserve(order_in.size());
    std::vector<int> new_pos(order_in.size(), -1);   // position in order_in -> position in the new order
    for (size_t wi = 0; wi < queue.size(); wi++) {
        const int b = queue[wi].bnode, level = queue[wi].level;
        out.depth = std::max(out.depth, level);
        // ---- the node's children: the best split of b's sub-tree into <= 8 roots
        int ch[8];
        int cnt = 0;
        if (an[b].n > 0 || (wi == 0 && as_leaf[b])) {  // the whole tree is one leaf slot
            ch[cnt++] = b;
        } else {
            cnt = col.children(b, ch);
        }
        // ---- octant-ordered slots: greedy assignment of (child, slot) pairs by dot(child centre - node centre, slot corner)
        double nlo[3] = {3.0e38, 3.0e38, 3.0e38}, nhi[3] = {-3.0e38, -3.0e38, -3.0e38};  // union of the padded child boxes (exact in double)
        for (int k = 0; k < cnt; k++)
            for (int a = 0; a < 3; a++) {
                nlo[a] = std::min(nlo[a], (double)an[ch[k]].AA[a] - (double)pad);
                nhi[a] = std::max(nhi[a], (double)an[ch[k]].BB[a] + (double)pad);
            }
        int slot_child[8];
        for (int s = 0; s < 8; s++) slot_child[s] = -1;
        {
            double cost[8][8];
            for (int k = 0; k < cnt; k++)
                for (int s = 0; s < 8; s++) {
                    double c = 0.0;
                    for (int a = 0; a < 3; a++) {
                        const double off = 0.5 * ((double)an[ch[k]].AA[a] + (double)an[ch[k]].BB[a]) - 0.5 * (nlo[a] + nhi[a]);
                        c += ((s >> axis_bit[a]) & 1) ? off : -off;
                    }
                    cost[k][s] = c;
                }
            bool child_done[8] = {false, false, false, false, false, false, false, false};
            for (int round = 0; round < cnt; round++) {
                int bk = -1, bs = -1;
                double bc = -1.0e300;
                for (int k = 0; k < cnt; k++) {
                    if (child_done[k]) continue;
                    for (int s = 0; s < 8; s++)
                        if (slot_child[s] < 0 && cost[k][s] > bc) { bc = cost[k][s]; bk = k; bs = s; }
                }
                child_done[bk] = true;
                slot_child[bs] = ch[bk];
            }
        }
        // ---- record
        uint32_t w[W8_NODE_WORDS];
        memset(w, 0, sizeof(w));
        uint8_t qlo[3][8], qhi[3][8], meta[8];
        memset(meta, 0, sizeof(meta));
        float origin[3], scale[3];
        for (int a = 0; a < 3; a++) {
            // 252 steps span the extent; the stored planes lie at least W8_SLACK_STEPS outside the exact ones (decode rounding, w8_node.h)
            double step = (nhi[a] - nlo[a]) / 252.0;
            int e;
            frexp(std::max(step, 1.0e-300), &e);            // 2^e > step
            double sc = ldexp(1.0, e);
            while (sc < min_step) sc *= 2.0;
            scale[a] = (float)sc;                            // a power of two: exact
            float org = (float)(nlo[a] - sc);
            while ((double)org > nlo[a] - sc) org = nextafterf(org, -3.0e38f);  // never above: low planes must not move up
            origin[a] = org;
            for (int s = 0; s < 8; s++) {
                const int c = slot_child[s];
                if (c < 0) { qlo[a][s] = 255; qhi[a][s] = 0; continue; }  // inverted: never hit
                const double lo = (double)an[c].AA[a] - (double)pad, hi = (double)an[c].BB[a] + (double)pad;
                double ql = floor((lo - (double)org) / sc - W8_SLACK_STEPS);
                double qh = ceil((hi - (double)org) / sc + W8_SLACK_STEPS);
                if (ql < 0.0 || qh > 255.0 || ql > qh) return -2;  // cannot happen with the 252-step rule; refuse rather than clamp
                qlo[a][s] = (uint8_t)ql;
                qhi[a][s] = (uint8_t)qh;
            }
        }
        uint32_t imask = 0;
        const uint32_t tri_base = (uint32_t)out.tri_order.size();
        const uint32_t child_base = (uint32_t)queue.size();
        for (int s = 0; s < 8; s++) {
            const int c = slot_child[s];
            if (c < 0) continue;
            out.n_children++;
            if (as_leaf[c]) {  // one leaf slot: all triangles of c's sub-tree (consecutive in order_in)
                const uint32_t off = (uint32_t)out.tri_order.size() - tri_base;
                if (count[c] > W8_MAX_LEAF_TRIS || off + (uint32_t)count[c] > W8_MAX_NODE_TRIS) return -3;
                meta[s] = (uint8_t)(((uint32_t)count[c] << 5) | off);
                for (int k = 0; k < count[c]; k++) {
                    new_pos[(size_t)first[c] + k] = (int)out.tri_order.size();
                    out.tri_order.push_back(order_in[(size_t)first[c] + k]);
                }
            } else {
                imask |= 1u << s;
                queue.push_back({c, level + 1});
            }
        }
        memcpy(&w[W8_W_ORIGIN], origin, 12);
        memcpy(&w[W8_W_SCALE], scale, 12);
        w[W8_W_CHILD_BASE] = child_base;
        w[W8_W_TRI_BASE] = tri_base;
        for (int a = 0; a < 3; a++) {
            memcpy(&w[W8_W_QLO + 2 * a], qlo[a], 8);
            memcpy(&w[W8_W_QHI + 2 * a], qhi[a], 8);
        }
        memcpy(&w[W8_W_META], meta, 8);
        w[W8_W_IMASK] = imask;
        out.nodes.insert(out.nodes.end(), w, w + W8_NODE_WORDS);
    }
    out.n_nodes = (int)queue.size();
    if (out.tri_order.size() != order_in.size()) return -4;
    for (int i = 0; i < NB; i++)
        if (an[i].n > 0) out.leaf_first[i] = new_pos[an[i].index];
    return 0;
}
