/*
 * ezrt_math.h -- the NORMATIVE fp32 arithmetic of the EzRT path-tracing hot path.
 *
 * The reference (AKGWSB/EzRT) runs this path as GLSL (P5/shaders/fshader.fsh) on top of
 * driver-implemented built-ins (normalize, cross, mix, sin, cos, atan, asin, log, pow ...)
 * and glm on the host; neither is bit-specified, and the reference ships no test that pins
 * them (how parity is pinned nevertheless: DESIGN.md section 2).  This header therefore DEFINES every such
 * operation as a fixed sequence of IEEE-754 binary32  + - * / sqrt fma  operations, so the
 * same source gives bit-identical results under g++ (host, -ffp-contract=off -mfma) and
 * nvcc (device, -fmad=false, default -prec-div/-prec-sqrt/-ftz=false).
 *
 * Rules:
 *   - every fused multiply-add is spelled EZ_FMA(); nothing else may be contracted;
 *   - min/max are the GLSL/glm ternaries: min(x,y) = (y<x)?y:x, max(x,y) = (x<y)?y:x
 *     (P5/fsh:226-230; glm/detail/func_common.inl);
 *   - dot/cross use FMA chains (what a GPU shader compiler emits for the GLSL built-ins);
 *   - transcendental functions are Cephes-style single-precision kernels (published
 *     algorithm, S. Moshier, "Cephes Mathematical Library", sinf/cosf/logf/expf/atanf/
 *     asinf) restated with explicit operation order.
 *
 * Used by: the CUDA kernels (ezrt_b200/csrc), the host scene pipeline, and the CPU oracle
 * (oracle/), which must share these primitive definitions to be comparable bit-for-bit.
 */
#ifndef EZRT_MATH_H
#define EZRT_MATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define EZ_HD __host__ __device__ __forceinline__
#else
#define EZ_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define EZ_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define EZ_SQRT(a) __fsqrt_rn(a)
#define EZ_DIV(a, b) __fdiv_rn((a), (b))
#else
#define EZ_FMA(a, b, c) __builtin_fmaf((a), (b), (c))
#define EZ_SQRT(a) __builtin_sqrtf(a)
#define EZ_DIV(a, b) ((a) / (b))
#endif

/* the shader's constants: P5/fsh:27-28 (PI is one ulp below float(pi)) */
#define EZ_PI 3.1415926f
#define EZ_INF 114514.0f

/* ------------------------------------------------------------------ bit casts */
EZ_HD uint32_t ez_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
EZ_HD float ez_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

/* ------------------------------------------------------------------ scalar helpers */
EZ_HD float ez_min(float x, float y) { return (y < x) ? y : x; }
EZ_HD float ez_max(float x, float y) { return (x < y) ? y : x; }
EZ_HD float ez_abs(float x) { return ez_u2f(ez_f2u(x) & 0x7fffffffu); }
EZ_HD float ez_clamp(float x, float lo, float hi) { return ez_min(ez_max(x, lo), hi); }
/* GLSL mix(x,y,a) = x*(1-a) + y*a */
EZ_HD float ez_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
EZ_HD float ez_sqr(float x) { return x * x; }
/* floor for |x| < 2^31 */
EZ_HD float ez_floor(float x) {
    float t = (float)(int)x;
    return (t > x) ? (t - 1.0f) : t;
}
/* uint -> float, round to nearest even (GLSL float(uint)) */
EZ_HD float ez_u32_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint2float_rn(u);
#else
    return (float)u;
#endif
}

/* ------------------------------------------------------------------ vec3 */
struct ez_vec3 {
    float x, y, z;
};
typedef struct ez_vec3 ez_vec3;

EZ_HD ez_vec3 ez_v3(float x, float y, float z) { ez_vec3 v; v.x = x; v.y = y; v.z = z; return v; }
EZ_HD ez_vec3 ez_add(ez_vec3 a, ez_vec3 b) { return ez_v3(a.x + b.x, a.y + b.y, a.z + b.z); }
EZ_HD ez_vec3 ez_sub(ez_vec3 a, ez_vec3 b) { return ez_v3(a.x - b.x, a.y - b.y, a.z - b.z); }
EZ_HD ez_vec3 ez_mul(ez_vec3 a, ez_vec3 b) { return ez_v3(a.x * b.x, a.y * b.y, a.z * b.z); }
EZ_HD ez_vec3 ez_scale(ez_vec3 a, float s) { return ez_v3(a.x * s, a.y * s, a.z * s); }
EZ_HD ez_vec3 ez_divs(ez_vec3 a, float s) { return ez_v3(EZ_DIV(a.x, s), EZ_DIV(a.y, s), EZ_DIV(a.z, s)); }
EZ_HD ez_vec3 ez_neg(ez_vec3 a) { return ez_v3(-a.x, -a.y, -a.z); }
EZ_HD ez_vec3 ez_vmin(ez_vec3 a, ez_vec3 b) { return ez_v3(ez_min(a.x, b.x), ez_min(a.y, b.y), ez_min(a.z, b.z)); }
EZ_HD ez_vec3 ez_vmax(ez_vec3 a, ez_vec3 b) { return ez_v3(ez_max(a.x, b.x), ez_max(a.y, b.y), ez_max(a.z, b.z)); }
EZ_HD ez_vec3 ez_vmix(ez_vec3 a, ez_vec3 b, float t) {
    return ez_v3(ez_mix(a.x, b.x, t), ez_mix(a.y, b.y, t), ez_mix(a.z, b.z, t));
}
/* dot = fma(z,z, fma(y,y, x*x)) */
EZ_HD float ez_dot(ez_vec3 a, ez_vec3 b) { return EZ_FMA(a.z, b.z, EZ_FMA(a.y, b.y, a.x * b.x)); }
/* cross: each component a*b - c*d = fma(a, b, -(c*d)) */
EZ_HD ez_vec3 ez_cross(ez_vec3 a, ez_vec3 b) {
    return ez_v3(EZ_FMA(a.y, b.z, -(a.z * b.y)),
                 EZ_FMA(a.z, b.x, -(a.x * b.z)),
                 EZ_FMA(a.x, b.y, -(a.y * b.x)));
}
/* normalize(v) = v * (1/sqrt(dot(v,v)))   (glm: v * inversesqrt(dot(v,v))) */
EZ_HD ez_vec3 ez_normalize(ez_vec3 v) {
    float inv = EZ_DIV(1.0f, EZ_SQRT(ez_dot(v, v)));
    return ez_v3(v.x * inv, v.y * inv, v.z * inv);
}
/* reflect(I,N) = I - 2*dot(N,I)*N */
EZ_HD ez_vec3 ez_reflect(ez_vec3 I, ez_vec3 N) {
    float k = 2.0f * ez_dot(N, I);
    return ez_v3(I.x - k * N.x, I.y - k * N.y, I.z - k * N.z);
}

/* ------------------------------------------------------------------ frexp / ldexp */
/* x = m * 2^e, m in [0.5,1); x must be finite and > 0 */
EZ_HD float ez_frexp_pos(float x, int* e) {
    uint32_t u = ez_f2u(x);
    int bias = 0;
    if ((u & 0x7f800000u) == 0u) { /* denormal: scale by 2^24 (exact) */
        x = x * 16777216.0f;
        u = ez_f2u(x);
        bias = -24;
    }
    *e = (int)((u >> 23) & 0xffu) - 126 + bias;
    return ez_u2f((u & 0x807fffffu) | 0x3f000000u);
}
/* z * 2^n, z finite; two-step so that the intermediate scale factors stay normal */
EZ_HD float ez_ldexp(float z, int n) {
    if (n > 254) n = 254;
    if (n < -252) n = -252;
    int n1 = n / 2;
    int n2 = n - n1;
    float s1 = ez_u2f((uint32_t)(n1 + 127) << 23);
    float s2 = ez_u2f((uint32_t)(n2 + 127) << 23);
    return (z * s1) * s2;
}

/* ------------------------------------------------------------------ sin / cos */
#define EZ_FOPI 1.27323954473516f
#define EZ_DP1 0.78515625f
#define EZ_DP2 2.4187564849853515625e-4f
#define EZ_DP3 3.77489497744594108e-8f

EZ_HD float ez_sin_poly(float x, float z) {
    float y = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * x;
    return y + x;
}
EZ_HD float ez_cos_poly(float z) {
    float y = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z;
    y = y - 0.5f * z;
    return y + 1.0f;
}
/* valid for |x| < 8192 (the path only produces |x| <= 4*pi); larger arguments return 0 / 1 */
EZ_HD float ez_sin(float xx) {
    float sign = 1.0f;
    float x = xx;
    if (x < 0.0f) { sign = -1.0f; x = -x; }
    if (!(x < 8192.0f)) return 0.0f;
    int j = (int)(EZ_FOPI * x);
    float y = (float)j;
    if (j & 1) { j += 1; y += 1.0f; }
    j &= 7;
    if (j > 3) { sign = -sign; j -= 4; }
    x = ((x - y * EZ_DP1) - y * EZ_DP2) - y * EZ_DP3;
    float z = x * x;
    float r = (j == 1 || j == 2) ? ez_cos_poly(z) : ez_sin_poly(x, z);
    return (sign < 0.0f) ? -r : r;
}
EZ_HD float ez_cos(float xx) {
    float sign = 1.0f;
    float x = xx;
    if (x < 0.0f) x = -x;
    if (!(x < 8192.0f)) return 1.0f;
    int j = (int)(EZ_FOPI * x);
    float y = (float)j;
    if (j & 1) { j += 1; y += 1.0f; }
    j &= 7;
    if (j > 3) { j -= 4; sign = -sign; }
    if (j > 1) sign = -sign;
    x = ((x - y * EZ_DP1) - y * EZ_DP2) - y * EZ_DP3;
    float z = x * x;
    float r = (j == 1 || j == 2) ? ez_sin_poly(x, z) : ez_cos_poly(z);
    return (sign < 0.0f) ? -r : r;
}

/* ------------------------------------------------------------------ log / exp / pow */
/* natural log; x <= 0 or NaN returns -EZ_HUGE (x==0) or NaN */
EZ_HD float ez_log(float x) {
    if (!(x > 0.0f)) {
        if (x == 0.0f) return -3.402823466e38f;
        return ez_u2f(0x7fc00000u);
    }
    if (ez_f2u(x) >= 0x7f800000u) return x; /* +inf */
    int e;
    x = ez_frexp_pos(x, &e);
    if (x < 0.707106781186547524f) { e -= 1; x = x + x - 1.0f; }
    else { x = x - 1.0f; }
    float z = x * x;
    float y = ((((((((7.0376836292e-2f * x - 1.1514610310e-1f) * x + 1.1676998740e-1f) * x
                    - 1.2420140846e-1f) * x + 1.4249322787e-1f) * x - 1.6668057665e-1f) * x
                 + 2.0000714765e-1f) * x - 2.4999993993e-1f) * x + 3.3333331174e-1f) * x * z;
    float fe = (float)e;
    if (e != 0) y = y + (-2.12194440e-4f) * fe;
    y = y + (-0.5f) * z;
    z = x + y;
    if (e != 0) z = z + 0.693359375f * fe;
    return z;
}
EZ_HD float ez_exp(float x) {
    if (x != x) return x;
    if (x > 88.72283905206835f) return ez_u2f(0x7f800000u);
    if (x < -103.278929903431851103f) return 0.0f;
    float z = ez_floor(1.44269504088896341f * x + 0.5f);
    x = x - z * 0.693359375f;
    x = x - z * (-2.12194440e-4f);
    int n = (int)z;
    z = x * x;
    z = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x
           + 4.1665795894e-2f) * x + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    return ez_ldexp(z, n);
}
/* GLSL pow(x,y), x > 0:  exp(y*log(x)); pow(0,y>0) = 0; undefined (NaN) for x < 0 */
EZ_HD float ez_pow(float x, float y) {
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : ((y == 0.0f) ? 1.0f : ez_u2f(0x7f800000u));
    if (x < 0.0f) return ez_u2f(0x7fc00000u);
    return ez_exp(y * ez_log(x));
}

/* ------------------------------------------------------------------ atan / asin */
#define EZ_TRUE_PI 3.14159265358979323846f
#define EZ_PIO2 1.5707963267948966192f
#define EZ_PIO4 0.7853981633974483096f

EZ_HD float ez_atan(float xx) {
    float sign = 1.0f;
    float x = xx;
    if (x < 0.0f) { sign = -1.0f; x = -x; }
    float y;
    if (x > 2.414213562373095f) { y = EZ_PIO2; x = -EZ_DIV(1.0f, x); }
    else if (x > 0.4142135623730950f) { y = EZ_PIO4; x = EZ_DIV(x - 1.0f, x + 1.0f); }
    else { y = 0.0f; }
    float z = x * x;
    y = y + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
              - 3.33329491539e-1f) * z * x + x);
    return (sign < 0.0f) ? -y : y;
}
/* GLSL atan(y,x) */
EZ_HD float ez_atan2(float y, float x) {
    if (x == 0.0f) {
        if (y > 0.0f) return EZ_PIO2;
        if (y < 0.0f) return -EZ_PIO2;
        return 0.0f;
    }
    if (y == 0.0f) return (x < 0.0f) ? EZ_TRUE_PI : 0.0f;
    float w = 0.0f;
    if (x < 0.0f) w = (y < 0.0f) ? -EZ_TRUE_PI : EZ_TRUE_PI;
    float z = ez_atan(EZ_DIV(y, x));
    return w + z;
}
/* GLSL asin(x); |x|>1 (undefined in GLSL) is treated as |x|=1 */
EZ_HD float ez_asin(float xx) {
    float sign = 1.0f;
    float a = xx;
    if (a < 0.0f) { sign = -1.0f; a = -a; }
    if (a != a) return a;
    if (a > 1.0f) a = 1.0f;
    if (a < 1.0e-4f) return xx;
    float x, z;
    int flag;
    if (a > 0.5f) { z = 0.5f * (1.0f - a); x = EZ_SQRT(z); flag = 1; }
    else { x = a; z = x * x; flag = 0; }
    z = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z
          + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * x + x;
    if (flag) { z = z + z; z = EZ_PIO2 - z; }
    return (sign < 0.0f) ? -z : z;
}

/* ------------------------------------------------------------------ post pass */
/* pass3.fsh:14-25: toneMapping(c, limit) = c * 1.0 / (1.0 + lum / limit), then pow(c, vec3(1.0/2.2)) */
EZ_HD ez_vec3 ez_tonemap_pass3(ez_vec3 c, float limit) {
    float luminance = 0.3f * c.x + 0.6f * c.y + 0.1f * c.z;
    float den = 1.0f + EZ_DIV(luminance, limit);
    ez_vec3 t = ez_v3(EZ_DIV(c.x * 1.0f, den), EZ_DIV(c.y * 1.0f, den), EZ_DIV(c.z * 1.0f, den));
    const float g = EZ_DIV(1.0f, 2.2f);
    return ez_v3(ez_pow(t.x, g), ez_pow(t.y, g), ez_pow(t.z, g));
}

#endif /* EZRT_MATH_H */
