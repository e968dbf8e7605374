// accel_stats.cpp -- CPU model of the acceleration-tree traversal: work counts per ray for candidate layouts.
// Development tool (not product, not oracle): decides which layout is worth GPU time.
//   g++ -O2 -std=c++17 -fopenmp -Iinclude -Iezrt_b200/csrc tools/accel_stats.cpp ezrt_b200/csrc/host_scene.cpp ezrt_b200/csrc/errors.cpp -o /tmp/accel_stats
//   /tmp/accel_stats tris.f32 n_tris
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "ezrt_internal.h"

struct V { float x, y, z; };
static V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V norm(V a) { float l = sqrtf(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }

struct WNode { int n; float lo[8][3], hi[8][3]; int ref[8]; };  // ref >= 0 inner wide node, < 0: leaf ~(first<<4|cnt)... encoded below
struct Leaf { int first, cnt; };

static uint32_t rng_state = 12345;
static float rnd() { rng_state = rng_state * 1664525u + 1013904223u; return (rng_state >> 8) * (1.0f / 16777216.0f); }

int main(int argc, char** argv) {
    int n = atoi(argv[2]);
    std::vector<float> tris((size_t)n * 36);
    FILE* f = fopen(argv[1], "rb");
    if (fread(tris.data(), 4, tris.size(), f) != tris.size()) return 1;
    fclose(f);
    // rays: argv[3] = file of (o, d, kind) 7-float records dumped by the oracle (oracle_set_ray_dump), argv[4] = kind filter;
    // otherwise synthetic bounce rays from random surface points, cosine-ish hemisphere
    int NR = 200000;
    std::vector<V> ro, rd;
    if (argc > 3) {
        FILE* rf = fopen(argv[3], "rb");
        const int want = argc > 4 ? atoi(argv[4]) : -1;
        float r[7];
        while (fread(r, 4, 7, rf) == 7)
            if (want < 0 || (int)r[6] == want) { ro.push_back({r[0], r[1], r[2]}); rd.push_back({r[3], r[4], r[5]}); }
        fclose(rf);
        NR = (int)ro.size();
        printf("replaying %d recorded rays (kind %d)\n", NR, want);
    } else {
        ro.resize(NR); rd.resize(NR);
    }
    for (int i = 0; argc <= 3 && i < NR; i++) {
        int t = (int)(rnd() * n) % n;
        const float* p = &tris[(size_t)t * 36];
        V a{p[0], p[1], p[2]}, b{p[3], p[4], p[5]}, c{p[6], p[7], p[8]};
        float u = rnd(), v = rnd();
        if (u + v > 1) { u = 1 - u; v = 1 - v; }
        V P{a.x + u * (b.x - a.x) + v * (c.x - a.x), a.y + u * (b.y - a.y) + v * (c.y - a.y), a.z + u * (b.z - a.z) + v * (c.z - a.z)};
        V N = norm(cross(sub(b, a), sub(c, a)));
        V d;
        do { d = {2 * rnd() - 1, 2 * rnd() - 1, 2 * rnd() - 1}; } while (dot(d, d) > 1 || dot(d, d) < 1e-4f);
        d = norm(d);
        if (dot(d, N) < 0) d = {-d.x, -d.y, -d.z};
        d = norm(V{d.x + N.x, d.y + N.y, d.z + N.z});
        ro[i] = {P.x + 1e-3f * N.x, P.y + 1e-3f * N.y, P.z + 1e-3f * N.z};
        rd[i] = d;
    }
    for (int leaf_n : {8, 4}) {
        std::vector<EzrtAccelNode> an;
        std::vector<uint32_t> order;
        ezrt_build_accel(tris.data(), n, leaf_n, an, order);
        std::vector<float> geo((size_t)n * 9);
        for (int i = 0; i < n; i++) memcpy(&geo[(size_t)i * 9], &tris[(size_t)order[i] * 36], 36);
        for (int width : {4, 8}) for (int order_mode : {0, 1}) {
            // collapse
            std::vector<WNode> wn;
            std::vector<Leaf> leaves;
            auto area = [&](int c) { float x = an[c].BB[0] - an[c].AA[0], y = an[c].BB[1] - an[c].AA[1], z = an[c].BB[2] - an[c].AA[2]; return x * y + x * z + y * z; };
            std::function<int(int)> build = [&](int b) -> int {
                int id = (int)wn.size();
                wn.push_back(WNode());
                int ch[8] = {an[b].left, an[b].right};
                int cnt = 2;
                while (cnt < width) {
                    int best = -1; float ba = -1;
                    for (int k = 0; k < cnt; k++) if (an[ch[k]].n <= 0 && area(ch[k]) > ba) { ba = area(ch[k]); best = k; }
                    if (best < 0) break;
                    int c = ch[best];
                    ch[best] = an[c].left; ch[cnt++] = an[c].right;
                }
                WNode w; w.n = cnt;
                for (int k = 0; k < cnt; k++) {
                    const EzrtAccelNode& c = an[ch[k]];
                    for (int a = 0; a < 3; a++) { w.lo[k][a] = c.AA[a]; w.hi[k][a] = c.BB[a]; }
                    if (c.n > 0) { w.ref[k] = -1 - (int)leaves.size(); leaves.push_back({c.index, c.n}); }
                    else w.ref[k] = build(ch[k]);
                }
                wn[id] = w;
                return id;
            };
            build(0);
            // traverse
            double nv = 0, lv = 0, tt = 0, hits = 0, maxsp = 0, pushes = 0, multi = 0;
#pragma omp parallel for reduction(+ : nv, lv, tt, hits, pushes, multi) reduction(max : maxsp)
            for (int r = 0; r < NR; r++) {
                V o = ro[r], d = rd[r];
                float inv[3] = {1 / d.x, 1 / d.y, 1 / d.z}, oo[3] = {o.x, o.y, o.z};
                float best = 1e30f;
                struct E { int ref; float t; } st[128];
                int sp = 0;
                int cur = 0;
                while (true) {
                    if (cur >= 0) {
                        const WNode& w = wn[cur];
                        nv += 1;
                        E hit[8]; int nh = 0;
                        for (int k = 0; k < w.n; k++) {
                            float t0 = 0, t1 = best;
                            for (int a = 0; a < 3; a++) {
                                float ta = (w.lo[k][a] - oo[a]) * inv[a], tb = (w.hi[k][a] - oo[a]) * inv[a];
                                t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                            }
                            if (t1 >= t0) hit[nh++] = {w.ref[k], t0};
                        }
                        if (order_mode == 0) std::sort(hit, hit + nh, [](const E& a, const E& b) { return a.t < b.t; });
                        else if (order_mode == 1) {  // nearest first, the others in slot order
                            int m = 0;
                            for (int k = 1; k < nh; k++) if (hit[k].t < hit[m].t) m = k;
                            if (nh) std::swap(hit[0], hit[m]);
                        } else if (nh > 2) {  // nearest first, farthest last (pushed first), middle unsorted
                            int m = 0;
                            for (int k = 1; k < nh; k++) if (hit[k].t < hit[m].t) m = k;
                            std::swap(hit[0], hit[m]);
                            int f = 1;
                            for (int k = 2; k < nh; k++) if (hit[k].t > hit[f].t) f = k;
                            std::swap(hit[nh - 1], hit[f]);
                        } else if (nh == 2 && hit[1].t < hit[0].t) std::swap(hit[0], hit[1]);
                        for (int k = nh - 1; k >= 1; k--) { st[sp++] = hit[k]; pushes += 1; }
                        if (nh >= 3) multi += 1;
                        if ((double)sp > maxsp) maxsp = sp;
                        if (nh) { cur = hit[0].ref; continue; }
                    } else {
                        const Leaf& L = leaves[-1 - cur];
                        lv += 1; tt += L.cnt;
                        for (int k = 0; k < L.cnt; k++) {
                            const float* g = &geo[(size_t)(L.first + k) * 9];
                            V a{g[0], g[1], g[2]}, e1 = sub(V{g[3], g[4], g[5]}, a), e2 = sub(V{g[6], g[7], g[8]}, a);
                            V pv = cross(d, e2); float det = dot(e1, pv);
                            if (fabsf(det) < 1e-12f) continue;
                            float id = 1 / det; V tv = sub(o, a); float u = dot(tv, pv) * id;
                            if (u < 0 || u > 1) continue;
                            V qv = cross(tv, e1); float v = dot(d, qv) * id;
                            if (v < 0 || u + v > 1) continue;
                            float t = dot(e2, qv) * id;
                            if (t > 1e-4f && t < best) best = t;
                        }
                    }
                    bool got = false;
                    while (sp > 0) { E e = st[--sp]; if (e.t > best) continue; cur = e.ref; got = true; break; }
                    if (!got) break;
                }
                if (best < 1e30f) hits += 1;
            }
            printf("leaf<=%d width %d order %d: wide nodes %zu leaves %zu | per ray: node visits %.1f leaf visits %.2f tri tests %.1f pushes %.1f visits with >=3 children hit %.1f%% hit %.2f maxsp %.0f | node bytes/ray %.0f (128B/64B/80B)\n",
                   leaf_n, width, order_mode, wn.size(), leaves.size(), nv / NR, lv / NR, tt / NR, pushes / NR, 100.0 * multi / nv, hits / NR, maxsp,
                   nv / NR * (width == 2 ? 64 : width == 4 ? 128 : 80));
        }
    }
    return 0;
}
