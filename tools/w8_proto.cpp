// w8_proto.cpp -- CPU prototype of the 8-wide, 8-bit-quantised acceleration node planned for round 2.
// Development tool (not product, not oracle).  It answers, without a GPU:
//   (1) are the quantised child boxes conservative under the DEVICE decode arithmetic
//         f = 2^23 + q  (one PRMT),   t = fma(f, B, A'),   B = scale * inv_d,   A' = fma(-2^23, B, (origin - o) * inv_d)
//       i.e. does the traversal still find the global closest hit?  (checked against an exhaustive test of every ray against
//       the triangles its true-box traversal reaches: same best distance bit for bit)
//   (2) how many extra node visits does the quantisation cost (vs. exact boxes, same tree)?
//   g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -mfma -Iinclude -Iezrt_b200/csrc tools/w8_proto.cpp ezrt_b200/csrc/host_scene.cpp ezrt_b200/csrc/errors.cpp -o /tmp/w8_proto
//   /tmp/w8_proto tris.f32 n_tris rays.f32   (rays: 7-float records from oracle_set_ray_dump)
#include <math.h>
#include <stdint.h>
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

struct W8 {
    int n;
    float lo[8][3], hi[8][3];       // exact (padded) child boxes
    float origin[3];
    int exp[3];                     // scale = 2^exp
    uint8_t qlo[8][3], qhi[8][3];   // stored bytes: plane = origin + q * scale, one step of padding included
    int ref[8];                     // >= 0 inner node, < 0: ~leaf
};
struct Leaf { int first, cnt; };

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int n = atoi(argv[2]);
    std::vector<float> tris((size_t)n * 36);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(tris.data(), 4, tris.size(), f) != tris.size()) return 1;
    fclose(f);
    std::vector<V> ro, rd;
    {
        FILE* rf = fopen(argv[3], "rb");
        float r[7];
        while (rf && fread(r, 4, 7, rf) == 7) { ro.push_back({r[0], r[1], r[2]}); rd.push_back({r[3], r[4], r[5]}); }
        if (rf) fclose(rf);
    }
    const int NR = (int)ro.size();
    float maxc = 0;
    for (int i = 0; i < n; i++) for (int k = 0; k < 9; k++) maxc = fmaxf(maxc, fabsf(tris[(size_t)i * 36 + k]));
    const float delta = maxc * 1.52587890625e-05f, pad = 2.0f * delta;  // as capi.cu: boxes inflated by 2 delta

    std::vector<EzrtAccelNode> an;
    std::vector<uint32_t> order;
    ezrt_build_accel(tris.data(), n, 4, an, order);
    std::vector<float> geo((size_t)n * 9);
    for (int i = 0; i < n; i++) memcpy(&geo[(size_t)i * 9], &tris[(size_t)order[i] * 36], 36);

    std::vector<W8> wn;
    std::vector<Leaf> leaves;
    auto area = [&](int c) { float x = an[c].BB[0] - an[c].AA[0], y = an[c].BB[1] - an[c].AA[1], z = an[c].BB[2] - an[c].AA[2]; return x * y + x * z + y * z; };
    long clamp_fail = 0;
    std::function<int(int)> build = [&](int b) -> int {
        const int id = (int)wn.size();
        wn.push_back(W8());
        int ch[8] = {an[b].left, an[b].right};
        int cnt = 2;
        while (cnt < 8) {
            int best = -1; float ba = -1;
            for (int k = 0; k < cnt; k++) if (an[ch[k]].n <= 0 && area(ch[k]) > ba) { ba = area(ch[k]); best = k; }
            if (best < 0) break;
            const int c = ch[best];
            ch[best] = an[c].left; ch[cnt++] = an[c].right;
        }
        W8 w; w.n = cnt;
        float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
        for (int k = 0; k < cnt; k++)
            for (int a = 0; a < 3; a++) {
                w.lo[k][a] = an[ch[k]].AA[a] - pad; w.hi[k][a] = an[ch[k]].BB[a] + pad;
                mn[a] = fminf(mn[a], w.lo[k][a]); mx[a] = fmaxf(mx[a], w.hi[k][a]);
            }
        for (int a = 0; a < 3; a++) {
            // 252 usable steps for the extent, stored values 1..254 hold the rounded planes, 0 / 255 the padding step
            int e; frexpf((mx[a] - mn[a]) / 252.0f, &e);   // 2^e > extent / 252
            w.exp[a] = e;
            const double scale = ldexp(1.0, e);
            w.origin[a] = (float)((double)mn[a] - scale);   // one step below the lowest plane
            // make sure the float origin is not above the true value (rounding of the subtraction)
            while ((double)w.origin[a] > (double)mn[a] - scale) w.origin[a] = nextafterf(w.origin[a], -3e38f);
            for (int k = 0; k < cnt; k++) {
                const double ql = floor(((double)w.lo[k][a] - (double)w.origin[a]) / scale) - 1.0;  // one step of padding (decode rounding)
                const double qh = ceil(((double)w.hi[k][a] - (double)w.origin[a]) / scale) + 1.0;
                if (ql < 0 || qh > 255) clamp_fail++;
                w.qlo[k][a] = (uint8_t)std::max(0.0, ql);
                w.qhi[k][a] = (uint8_t)std::min(255.0, qh);
            }
        }
        for (int k = 0; k < cnt; k++) {
            const EzrtAccelNode& c = an[ch[k]];
            if (c.n > 0) { w.ref[k] = -1 - (int)leaves.size(); leaves.push_back({c.index, c.n}); }
            else w.ref[k] = build(ch[k]);
        }
        wn[id] = w;
        return id;
    };
    build(0);
    printf("8-wide nodes %zu, leaves %zu, quantisation range failures %ld\n", wn.size(), leaves.size(), clamp_fail);

    for (int quant = 0; quant < 2; quant++) {
        double nv = 0, lv = 0, tt = 0, nhit = 0;
        long mismatch = 0;
        std::vector<float> best_out(NR);
#pragma omp parallel for reduction(+ : nv, lv, tt, nhit)
        for (int r = 0; r < NR; r++) {
            const V o = ro[r], d = rd[r];
            const float inv[3] = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z}, oo[3] = {o.x, o.y, o.z};
            if (!(fabsf(inv[0]) < 3e38f && fabsf(inv[1]) < 3e38f && fabsf(inv[2]) < 3e38f)) { best_out[r] = -1; continue; }
            const float slack = delta * fmaxf(fabsf(inv[0]), fmaxf(fabsf(inv[1]), fabsf(inv[2])));
            float best = 114514.0f;
            struct E { int ref; float t; } st[160];
            int sp = 0, cur = 0;
            while (true) {
                if (cur >= 0) {
                    const W8& w = wn[cur];
                    nv += 1;
                    const float limit = best + (best * 0.000244140625f + slack);
                    float A[3], B[3];
                    for (int a = 0; a < 3; a++) {
                        const float scale = ldexpf(1.0f, w.exp[a]);
                        B[a] = scale * inv[a];
                        A[a] = fmaf(-8388608.0f, B[a], (w.origin[a] - oo[a]) * inv[a]);
                    }
                    E hit[8]; int nh = 0;
                    for (int k = 0; k < w.n; k++) {
                        float t0 = -3e38f, t1 = 3e38f;
                        for (int a = 0; a < 3; a++) {
                            float ta, tb;
                            if (quant) {
                                ta = fmaf(8388608.0f + (float)w.qlo[k][a], B[a], A[a]);
                                tb = fmaf(8388608.0f + (float)w.qhi[k][a], B[a], A[a]);
                            } else {
                                ta = (w.lo[k][a] - oo[a]) * inv[a];
                                tb = (w.hi[k][a] - oo[a]) * inv[a];
                            }
                            t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                        }
                        if (t1 >= t0 && t1 > 0.0f && !(t0 > limit)) hit[nh++] = {w.ref[k], t0};
                    }
                    std::sort(hit, hit + nh, [](const E& a, const E& b) { return a.t < b.t; });
                    for (int k = nh - 1; k >= 1; k--) st[sp++] = hit[k];
                    if (nh) { cur = hit[0].ref; continue; }
                } else {
                    const Leaf& L = leaves[-1 - cur];
                    lv += 1; tt += L.cnt;
                    for (int k = 0; k < L.cnt; k++) {  // the reference's plane + edge test (P5/fsh:160-217), plain fp32 here
                        const float* g = &geo[(size_t)(L.first + k) * 9];
                        const V p1{g[0], g[1], g[2]}, p2{g[3], g[4], g[5]}, p3{g[6], g[7], g[8]};
                        V N = cross(sub(p2, p1), sub(p3, p1));
                        const float l = sqrtf(dot(N, N));
                        N = {N.x / l, N.y / l, N.z / l};
                        const float nd = dot(N, d);
                        if (!(fabsf(nd) >= 0.00001f)) continue;
                        const float t = (dot(N, p1) - dot(o, N)) / nd;
                        if (t < 0.0005f || !(t < best)) continue;
                        const V P{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t};
                        const float s1 = dot(cross(sub(p2, p1), sub(P, p1)), N), s2 = dot(cross(sub(p3, p2), sub(P, p2)), N),
                                    s3 = dot(cross(sub(p1, p3), sub(P, p3)), N);
                        if ((s1 > 0 && s2 > 0 && s3 > 0) || (s1 < 0 && s2 < 0 && s3 < 0)) best = t;
                    }
                }
                bool got = false;
                while (sp > 0) {
                    const E e = st[--sp];
                    if (e.t > best + (best * 0.000244140625f + slack)) continue;
                    cur = e.ref; got = true; break;
                }
                if (!got) break;
            }
            best_out[r] = best;
            if (best < 114514.0f) nhit += 1;
        }
        static std::vector<float> exact;
        if (!quant) exact = best_out;
        else for (int r = 0; r < NR; r++) if (memcmp(&exact[r], &best_out[r], 4) != 0) mismatch++;
        printf("%s boxes: per ray %.2f node visits, %.2f leaf visits, %.2f triangle tests, hit fraction %.3f%s\n", quant ? "quantised" : "exact    ",
               nv / NR, lv / NR, tt / NR, nhit / NR, quant ? "" : "");
        if (quant) printf("closest-hit distances differing from the exact-box traversal: %ld of %d rays\n", mismatch, NR);
    }
    return 0;
}
