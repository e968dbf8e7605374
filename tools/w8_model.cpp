// w8_model.cpp -- CPU model of the W8 acceleration-tree traversal (development + test tool; not product, not oracle).
// Builds the tree with the PRODUCT builders (ezrt_build_accel + ezrt_build_w8), then walks it with exactly the decode
// arithmetic and visit rule of the device kernel (w8_node.h, device_functions.cuh: octant-ordered hit masks, group
// stack, triangle masks) and checks every ray's closest-hit distance against brute force over all triangles
// (small scenes) or against the exact-box traversal of the binary tree (large scenes).  Prints work counts per ray.
//   g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -mfma -Iinclude -Iezrt_b200/csrc tools/w8_model.cpp \
//       ezrt_b200/csrc/host_scene.cpp ezrt_b200/csrc/accel_w8.cpp ezrt_b200/csrc/errors.cpp -o build/w8_model
//   build/w8_model tris.f32 n_tris rays.f32 [brute]        rays: 7-float records (o, d, kind) as oracle_set_ray_dump writes
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "ezrt.h"
#include "ezrt_internal.h"
#include "ezrt_math.h"
#include "w8_node.h"

struct TriRec { ez_vec3 p1, p2, p3, N; float d0; };

// tri_test_t<TIES> of device_functions.cuh (hitTriangle P5/fsh:160-217 on the repacked record)
static int tri_test(const TriRec& r, ez_vec3 o, ez_vec3 d, float best, float& tout) {
    float nd = ez_dot(r.N, d);
    if (ez_abs(nd) < 0.00001f) return 0;
    float t = EZ_DIV(r.d0 - ez_dot(o, r.N), nd);
    if (t < 0.0005f) return 0;
    if (!(t <= best)) return 0;
    ez_vec3 P = ez_add(o, ez_scale(d, t));
    float s1 = ez_dot(ez_cross(ez_sub(r.p2, r.p1), ez_sub(P, r.p1)), r.N);
    float s2 = ez_dot(ez_cross(ez_sub(r.p3, r.p2), ez_sub(P, r.p2)), r.N);
    float s3 = ez_dot(ez_cross(ez_sub(r.p1, r.p3), ez_sub(P, r.p3)), r.N);
    bool r1 = (s1 > 0.0f && s2 > 0.0f && s3 > 0.0f), r2 = (s1 < 0.0f && s2 < 0.0f && s3 < 0.0f);
    if (!(r1 || r2)) return 0;
    tout = t;
    return (t == best) ? 2 : 1;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: w8_model tris.f32 n_tris rays.f32 [brute]\n"); return 2; }
    const int n = atoi(argv[2]);
    const bool brute = argc > 4 && !strcmp(argv[4], "brute");
    std::vector<float> tris((size_t)n * 36);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(tris.data(), 4, tris.size(), f) != tris.size()) { fprintf(stderr, "cannot read triangles\n"); return 1; }
    fclose(f);
    std::vector<float> rays;
    {
        FILE* rf = fopen(argv[3], "rb");
        float r[7];
        while (rf && fread(r, 4, 7, rf) == 7) rays.insert(rays.end(), r, r + 7);
        if (rf) fclose(rf);
    }
    const int NR = (int)(rays.size() / 7);
    float maxc = 0, bmin[3] = {3e38f, 3e38f, 3e38f}, bmax[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 9; k++) {
            const float v = tris[(size_t)i * 36 + k];
            maxc = fmaxf(maxc, fabsf(v));
            bmin[k % 3] = fminf(bmin[k % 3], v);
            bmax[k % 3] = fmaxf(bmax[k % 3], v);
        }
    const float delta = maxc * 1.52587890625e-05f, pad = 2.0f * delta;
    int axis_bit[3];
    ezrt_w8_axis_bits(bmin, bmax, axis_bit);

    std::vector<EzrtAccelNode> an;
    std::vector<uint32_t> order;
    ezrt_build_accel(tris.data(), n, 4, an, order);
    EzrtW8Tree w8;
    const int rc = ezrt_build_w8(an, order, pad, maxc, axis_bit, w8);
    if (rc) { fprintf(stderr, "ezrt_build_w8 failed: %d\n", rc); return 1; }
    printf("binary nodes %zu, 8-wide nodes %d (%.1f MB), depth %d, mean fill %.2f, axis bits x%d y%d z%d\n", an.size(), w8.n_nodes,
           w8.n_nodes * 96.0 / 1e6, w8.depth, (double)w8.n_children / w8.n_nodes, axis_bit[0], axis_bit[1], axis_bit[2]);
    {   // the 4-wide collapse the default kernel uses (same dynamic programme, width 4): every triangle in exactly one leaf of <= 4
        EzrtCollapse c4;
        if (c4.build(an, 4, W8_MAX_LEAF_TRIS, 1.0, 0.3) != 0) { fprintf(stderr, "4-wide collapse failed\n"); return 1; }
        std::vector<char> seen(n, 0);
        long nodes4 = 0, kids = 0, bad = 0;
        std::vector<int> todo;
        if (an[0].n <= 0 && !c4.as_leaf[0]) todo.push_back(0);
        while (!todo.empty()) {
            const int b = todo.back();
            todo.pop_back();
            int ch[8];
            const int cnt = c4.children(b, ch);
            nodes4++;
            kids += cnt;
            if (cnt < 2 || cnt > 4) bad++;
            for (int k = 0; k < cnt; k++) {
                if (c4.as_leaf[ch[k]]) {
                    if (c4.count[ch[k]] < 1 || c4.count[ch[k]] > W8_MAX_LEAF_TRIS) bad++;
                    for (int t = 0; t < c4.count[ch[k]]; t++) { if (seen[c4.first[ch[k]] + t]++) bad++; }
                } else {
                    todo.push_back(ch[k]);
                }
            }
        }
        if (nodes4 > 0) for (int i = 0; i < n; i++) if (seen[i] != 1) bad++;
        printf("4-wide collapse: %ld nodes, mean fill %.2f, violations %ld\n", nodes4, nodes4 ? (double)kids / nodes4 : 0.0, bad);
        if (bad) return 4;
    }
    std::vector<TriRec> rec(n);
    for (int i = 0; i < n; i++) {
        const float* s = &tris[(size_t)w8.tri_order[i] * 36];
        TriRec& r = rec[i];
        r.p1 = ez_v3(s[0], s[1], s[2]); r.p2 = ez_v3(s[3], s[4], s[5]); r.p3 = ez_v3(s[6], s[7], s[8]);
        r.N = ez_normalize(ez_cross(ez_sub(r.p2, r.p1), ez_sub(r.p3, r.p1)));
        r.d0 = ez_dot(r.N, r.p1);
    }

    // experiment switches (environment): W8M_SORT=1 visit hit children by entry distance instead of octant order;
    // W8M_GMIN=1 keep the smallest entry distance of a pushed group and drop the group at pop when it is beyond the best hit;
    // W8M_EXACT=1 exact child boxes instead of the quantised ones (how much the 8-bit planes cost)
    const bool x_sort = getenv("W8M_SORT") && atoi(getenv("W8M_SORT")), x_gmin = getenv("W8M_GMIN") && atoi(getenv("W8M_GMIN"));
    double nv[3] = {0, 0, 0}, nt[3] = {0, 0, 0}, npush[3] = {0, 0, 0}, cntk[3] = {0, 0, 0};
    long mismatch = 0, skipped = 0, ties = 0;
    int max_sp = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : mismatch, skipped, ties) reduction(max : max_sp)
    for (int r = 0; r < NR; r++) {
        const float* R = &rays[(size_t)r * 7];
        const ez_vec3 o = ez_v3(R[0], R[1], R[2]), d = ez_v3(R[3], R[4], R[5]);
        const int kind = std::min(2, std::max(0, (int)R[6]));
        const float inv[3] = {EZ_DIV(1.0f, d.x), EZ_DIV(1.0f, d.y), EZ_DIV(1.0f, d.z)}, oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
        const float ax = fabsf(inv[0]), ay = fabsf(inv[1]), az = fabsf(inv[2]);
        const float olim = W8_ORIGIN_LIMIT_REL * maxc;
        if (!(ax <= W8_INV_LIMIT && ay <= W8_INV_LIMIT && az <= W8_INV_LIMIT && ax >= W8_INV_MIN && ay >= W8_INV_MIN && az >= W8_INV_MIN) || !(fabsf(oo[0]) <= olim && fabsf(oo[1]) <= olim && fabsf(oo[2]) <= olim)) {
            skipped++;  // the kernel hands these to the exact traversal
            continue;
        }
        const float slack = delta * fmaxf(ax, fmaxf(ay, az));
        uint32_t near_mask = 0;
        for (int a = 0; a < 3; a++) if (dd[a] >= 0.0f) near_mask |= 1u << axis_bit[a];
        float best = EZ_INF;
        bool tie = false;
        struct Group { uint32_t base, bits; float tmin; uint32_t order; float ts[8]; } st[64];
        float g_ts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const bool x_tsel = getenv("W8M_TSEL") && atoi(getenv("W8M_TSEL"));  // per-child entry distance kept with the group (ideal pop pruning)
        const bool x_exact = getenv("W8M_EXACT") && atoi(getenv("W8M_EXACT"));
        float g_tmin = 0.0f;
        uint32_t g_order = 0;   // W8M_SORT: slots in visit order, 4 bits each, first = lowest nibble
        int sp = 0;
        uint32_t g_base = 0, g_bits = 0;   // bits: imask (low 8) | hits in priority positions (bits 8..15)
        int node = 0;
        double my_nv = 0, my_nt = 0, my_push = 0;
        while (true) {
            uint32_t t_base = 0, t_mask = 0;
            if (node >= 0) {
                const uint32_t* w = &w8.nodes[(size_t)node * W8_NODE_WORDS];
                my_nv += 1;
                const float limit = best + (best * 0.000244140625f + slack);
                float A[3], B[3];
                for (int a = 0; a < 3; a++) {
                    float org, sc;
                    memcpy(&org, &w[W8_W_ORIGIN + a], 4);
                    memcpy(&sc, &w[W8_W_SCALE + a], 4);
                    B[a] = sc * inv[a];
                    A[a] = fmaf(-W8_DECODE_BIAS, B[a], (org - oo[a]) * inv[a]);
                }
                const uint8_t* qlo = (const uint8_t*)&w[W8_W_QLO];
                const uint8_t* qhi = (const uint8_t*)&w[W8_W_QHI];
                const uint8_t* meta = (const uint8_t*)&w[W8_W_META];
                const uint32_t imask = w[W8_W_IMASK] & 255u;
                uint32_t hits8 = 0;
                float tmin_s[8];
                for (int s = 0; s < 8; s++) {
                    float tn[3], tf[3];
                    for (int a = 0; a < 3; a++) {
                        const uint8_t lo = qlo[8 * a + s], hi = qhi[8 * a + s];
                        const uint8_t nr = dd[a] >= 0.0f ? lo : hi, fr = dd[a] >= 0.0f ? hi : lo;
                        tn[a] = fmaf(W8_DECODE_BIAS + (float)nr + (x_exact ? (dd[a] >= 0.0f ? 1.5f : -1.5f) : 0.0f), B[a], A[a]);
                        tf[a] = fmaf(W8_DECODE_BIAS + (float)fr + (x_exact ? (dd[a] >= 0.0f ? -1.5f : 1.5f) : 0.0f), B[a], A[a]);
                    }
                    const float tmin = fmaxf(fmaxf(tn[0], tn[1]), fmaxf(tn[2], 0.0f));
                    const float tmax = fminf(fminf(tf[0], tf[1]), fminf(tf[2], limit));
                    if (tmin <= tmax) hits8 |= 1u << s;
                    tmin_s[s] = tmin;
                }
                uint32_t inner = hits8 & imask, leaf = hits8 & ~imask, perm = 0;
                for (int s = 0; s < 8; s++) if (inner >> s & 1) perm |= 1u << (s ^ near_mask);
                for (int s = 0; s < 8; s++)
                    if (leaf >> s & 1) t_mask |= ((1u << (meta[s] >> 5)) - 1u) << (meta[s] & 31u);
                t_base = w[W8_W_TRI_BASE];
                if (g_bits >> 8) { st[sp].base = g_base; st[sp].bits = g_bits; st[sp].tmin = g_tmin; st[sp].order = g_order; memcpy(st[sp].ts, g_ts, sizeof(g_ts)); sp++; my_push += 1; max_sp = std::max(max_sp, sp); }
                g_base = w[W8_W_CHILD_BASE];
                g_bits = imask | (perm << 8);
                g_tmin = 3.0e38f;
                for (int s = 0; s < 8; s++) if (inner >> s & 1) g_tmin = fminf(g_tmin, tmin_s[s]);
                memcpy(g_ts, tmin_s, sizeof(g_ts));
                if (x_sort) {  // hit inner slots by ascending entry distance; bits 8.. = slot mask (unpermuted)
                    int idx[8], m = 0;
                    for (int s = 0; s < 8; s++) if (inner >> s & 1) idx[m++] = s;
                    std::sort(idx, idx + m, [&](int a, int b) { return tmin_s[a] < tmin_s[b]; });
                    g_order = 0;
                    for (int k = m - 1; k >= 0; k--) g_order = (g_order << 4) | (uint32_t)idx[k];
                    g_bits = imask | (inner << 8);
                }
            }
            while (t_mask) {  // the node's triangles, lowest offset first
                const int k = __builtin_ctz(t_mask);
                t_mask &= t_mask - 1;
                my_nt += 1;
                float t;
                const int h = tri_test(rec[t_base + k], o, d, best, t);
                if (h == 2) tie = true;
                else if (h == 1) { best = t; tie = false; }
            }
            bool done = false;
            while ((g_bits >> 8) == 0) {
                if (sp == 0) { done = true; break; }
                --sp;
                g_base = st[sp].base;
                g_bits = st[sp].bits;
                g_tmin = st[sp].tmin;
                g_order = st[sp].order;
                memcpy(g_ts, st[sp].ts, sizeof(g_ts));
                if (x_gmin && g_tmin > best + (best * 0.000244140625f + slack)) g_bits &= 255u;  // the whole group lies beyond the best hit
            }
            if (done) break;
            int slot;
            if (x_sort) {
                slot = (int)(g_order & 15u);
                g_order >>= 4;
                g_bits ^= 1u << (8 + slot);
            } else {
                const int p = 31 - __builtin_clz(g_bits >> 8);      // highest priority position
                g_bits ^= 1u << (8 + p);
                slot = p ^ (int)near_mask;
            }
            if (x_tsel && g_ts[slot] > best + (best * 0.000244140625f + slack)) { node = -1; continue; }
            if (x_gmin) {  // entry distance of what stays behind in the group (model: exact minimum over the remaining hit slots)
                // NOTE: needs the t values of the node the group came from; the model keeps them only for the node just visited,
                // so the minimum is taken when the group is created (below) and is a lower bound afterwards
            }
            node = (int)(g_base + __builtin_popcount(g_bits & 255u & ((1u << slot) - 1u)));
        }
        if (tie) ties++;
        // ---- check
        float want = EZ_INF;
        if (brute) {
            for (int i = 0; i < n; i++) { float t; if (tri_test(rec[i], o, d, want, t) != 0) want = t; }
        } else {  // exact (padded) boxes of the binary tree, pruned
            int stk[128], sp2 = 0, cur = 0;
            while (true) {
                const EzrtAccelNode& nd = an[cur];
                bool descend = false;
                if (nd.n > 0) {
                    const int first = w8.leaf_first[cur];
                    for (int k = 0; k < nd.n; k++) { float t; if (tri_test(rec[first + k], o, d, want, t) != 0) want = t; }
                } else {
                    const float limit = want + (want * 0.000244140625f + slack);
                    int c[2] = {nd.left, nd.right};
                    bool hit[2];
                    for (int j = 0; j < 2; j++) {
                        float t0 = -3e38f, t1 = 3e38f;
                        for (int a = 0; a < 3; a++) {
                            const float ta = ((an[c[j]].AA[a] - pad) - oo[a]) * inv[a], tb = ((an[c[j]].BB[a] + pad) - oo[a]) * inv[a];
                            t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                        }
                        hit[j] = t1 >= t0 && t1 > 0.0f && !(t0 > limit);
                    }
                    if (hit[0] && hit[1]) { stk[sp2++] = c[1]; cur = c[0]; descend = true; }
                    else if (hit[0]) { cur = c[0]; descend = true; }
                    else if (hit[1]) { cur = c[1]; descend = true; }
                }
                if (descend) continue;
                if (sp2 == 0) break;
                cur = stk[--sp2];
            }
        }
        if (memcmp(&want, &best, 4) != 0) mismatch++;
#pragma omp critical
        { nv[kind] += my_nv; nt[kind] += my_nt; npush[kind] += my_push; cntk[kind] += 1; }
    }
    const char* names[3] = {"camera", "bounce", "shadow"};
    for (int k = 0; k < 3; k++)
        if (cntk[k] > 0)
            printf("%s rays %.0f: %.2f node visits, %.2f triangle tests, %.2f pushes per ray\n", names[k], cntk[k], nv[k] / cntk[k], nt[k] / cntk[k], npush[k] / cntk[k]);
    printf("max stack depth %d, rays left to the exact kernel %ld, rays with a tie %ld\n", max_sp, skipped, ties);
    printf("closest-hit distances differing from %s: %ld of %d rays\n", brute ? "brute force" : "the exact-box traversal", mismatch, NR);
    return mismatch == 0 ? 0 : 3;
}
