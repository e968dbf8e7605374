// w4_check.cpp -- CPU check of the 4-wide acceleration tree builder (ezrt_b200/csrc/accel_w8.cpp ezrt_build_w4; test tool, not product):
// (1) the arrays do not depend on the number of threads (1 against 2, 5, 16), for both collapse rules;
// (2) structure: every triangle of the binary tree's order lies in exactly one leaf, every child box contains the boxes of the
//     triangles below it inflated by `pad`, the Q16 planes contain the exact boxes, references stay inside the array.
//   build/w4_check tris.f32 n_tris
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "ezrt.h"
#include "ezrt_internal.h"
#include "w8_node.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: w4_check tris.f32 n_tris\n"); return 2; }
    const int n = atoi(argv[2]);
    std::vector<float> tris((size_t)n * 36);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(tris.data(), 4, tris.size(), f) != tris.size()) { fprintf(stderr, "cannot read triangles\n"); return 1; }
    fclose(f);
    float max_abs = 0;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 9; k++) max_abs = std::max(max_abs, fabsf(tris[(size_t)i * 36 + k]));
    std::vector<EzrtAccelNode> an;
    std::vector<uint32_t> order;
    ezrt_build_accel(tris.data(), n, W8_MAX_LEAF_TRIS, an, order);
    const float pad = 2.0f * max_abs * 1.52587890625e-05f;
    int bad = 0;
    for (int greedy = 0; greedy < 2; greedy++) {
        EzrtW4Tree ref;
        if (ezrt_build_w4(an, pad, max_abs, greedy != 0, true, 1, ref) != 0) { printf("build failed\n"); return 1; }
        for (int threads : {2, 5, 16}) {
            setenv("EZRT_W4_FORCE_THREADS", "1", 1);   // small inputs too
            EzrtW4Tree t;
            ezrt_build_w4(an, pad, max_abs, greedy != 0, true, threads, t);
            const bool same = t.nodes.size() == ref.nodes.size() && !memcmp(t.nodes.data(), ref.nodes.data(), ref.nodes.size() * 4) &&
                              t.q16.size() == ref.q16.size() && !memcmp(t.q16.data(), ref.q16.data(), ref.q16.size() * 4) && t.depth == ref.depth;
            if (!same) { bad++; printf("greedy %d threads %d: arrays differ\n", greedy, threads); }
        }
        // structure
        const int NW = (int)(ref.nodes.size() / 32);
        std::vector<int> covered(n, 0);
        long long viol = 0;
        struct It { int node; float lo[3], hi[3]; };
        std::vector<It> stk;
        stk.push_back({ref.root, {-3e38f, -3e38f, -3e38f}, {3e38f, 3e38f, 3e38f}});
        long long visited = 0;
        while (!stk.empty()) {
            const It it = stk.back();
            stk.pop_back();
            visited++;
            const float* r = &ref.nodes[(size_t)it.node * 32];
            const uint32_t* q = ref.q16.empty() ? nullptr : &ref.q16[(size_t)it.node * 24];
            int refs[4];
            memcpy(refs, r + 24, 16);
            for (int k = 0; k < 4; k++) {
                const float lo[3] = {r[4 * k], r[4 * k + 1], r[16 + 2 * k]}, hi[3] = {r[4 * k + 2], r[4 * k + 3], r[16 + 2 * k + 1]};
                if ((uint32_t)refs[k] == 0x80000000u) continue;   // absent
                if (q) {   // the quantised planes contain the exact box
                    if (q[18 + k] != (uint32_t)refs[k]) viol++;
                    for (int a = 0; a < 3; a++) {
                        float org, sc;
                        memcpy(&org, &q[a], 4);
                        memcpy(&sc, &q[3 + a], 4);
                        const uint32_t w = q[6 + 3 * k + a];
                        const double ql = (double)org + (double)sc * (double)(w & 0xffffu), qh = (double)org + (double)sc * (double)(w >> 16);
                        if (!(ql <= (double)lo[a] && qh >= (double)hi[a])) viol++;
                    }
                }
                if (refs[k] < 0) {
                    const int first = (int)(((uint32_t)refs[k] & 0x7fffffffu) >> 7), cnt = refs[k] & 127;
                    if (cnt < 1 || cnt > W8_MAX_LEAF_TRIS || first < 0 || first + cnt > n) { viol++; continue; }
                    for (int t = first; t < first + cnt; t++) {
                        covered[t]++;
                        const float* v = &tris[(size_t)order[t] * 36];
                        for (int c = 0; c < 9; c++)
                            if (!(v[c] - pad >= lo[c % 3] - 1e-30f && v[c] + pad <= hi[c % 3] + 1e-30f)) viol++;
                    }
                } else {
                    if (refs[k] >= NW) { viol++; continue; }
                    It c;
                    c.node = refs[k];
                    for (int a = 0; a < 3; a++) { c.lo[a] = lo[a]; c.hi[a] = hi[a]; }
                    stk.push_back(c);
                }
            }
        }
        long long uncovered = 0;
        for (int t = 0; t < n; t++) uncovered += covered[t] != 1;
        printf("greedy %d: %d wide nodes (%lld reached), depth %d, q16 %s, violations %lld, triangles not covered exactly once %lld\n", greedy, NW, visited,
               ref.depth, ref.q16.empty() ? "no" : "yes", viol, uncovered);
        if (viol || uncovered || visited != NW) bad++;
    }
    printf(bad ? "FAILED\n" : "OK\n");
    return bad ? 1 : 0;
}
