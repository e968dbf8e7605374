// glm/glm.hpp -- STAND-IN for the glm headers the reference's main.cpp includes (glm is not vendored
// by the reference and not installed here).
//
// *** TEST INFRASTRUCTURE (oracle/), NOT PRODUCT.  Lets oracle/ref_host_shim.cpp compile the reference's
// *** own main.cpp (readObj, buildBVH, buildBVHwithSAH, calculateHdrCache, the encode loops of main())
// *** from where it lies, so the product's host pipeline can be compared with the reference's code.
//
// Only what main.cpp uses.  Operation orders follow glm 0.9.9 (type_mat4x4.inl, matrix_transform.inl,
// func_geometric.inl); sin/cos/sqrt are the normative ones of include/ezrt_math.h, exactly as in the
// product's own restatement (ezrt_b200/csrc/host_scene.cpp) -- so the model matrix is NOT independently
// pinned by this stand-in; everything main.cpp does with the transformed vertices is.
#ifndef EZRT_STUB_GLM_HPP
#define EZRT_STUB_GLM_HPP

#include "ezrt_math.h"

namespace glm {

struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }

struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

struct mat4 {
    vec4 c[4];  // columns
    mat4() { c[0] = vec4(1, 0, 0, 0); c[1] = vec4(0, 1, 0, 0); c[2] = vec4(0, 0, 1, 0); c[3] = vec4(0, 0, 0, 1); }
    mat4(const vec4& a, const vec4& b, const vec4& d, const vec4& e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
// type_mat4x4.inl operator*(mat4, vec4): (m0*v.x + m1*v.y) + (m2*v.z + m3*v.w)
inline vec4 operator*(const mat4& m, const vec4& v) { return (m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w); }
// type_mat4x4.inl operator*(mat4, mat4): Result[i] = A0*B[i][0] + A1*B[i][1] + A2*B[i][2] + A3*B[i][3]
inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int i = 0; i < 4; i++) r[i] = ((a[0] * b[i][0] + a[1] * b[i][1]) + a[2] * b[i][2]) + a[3] * b[i][3];
    return r;
}

inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
// templates, as in glm: an unqualified sin(float) in main.cpp then resolves to <cmath>'s overload (display() only)
template <class T> inline T sin(T x) { return (T)ez_sin((float)x); }
template <class T> inline T cos(T x) { return (T)ez_cos((float)x); }
inline float min(float a, float b) { return (b < a) ? b : a; }  // func_common.inl
inline float max(float a, float b) { return (a < b) ? b : a; }
inline float dot(const vec3& a, const vec3& b) { return ez_dot(ez_v3(a.x, a.y, a.z), ez_v3(b.x, b.y, b.z)); }
inline vec3 cross(const vec3& a, const vec3& b) {
    ez_vec3 r = ez_cross(ez_v3(a.x, a.y, a.z), ez_v3(b.x, b.y, b.z));
    return vec3(r.x, r.y, r.z);
}
inline vec3 normalize(const vec3& a) {
    ez_vec3 r = ez_normalize(ez_v3(a.x, a.y, a.z));
    return vec3(r.x, r.y, r.z);
}
inline float length(const vec3& a) { return EZ_SQRT(dot(a, a)); }

// matrix_transform.inl
inline mat4 scale(const mat4& m, const vec3& v) { return mat4(m[0] * v.x, m[1] * v.y, m[2] * v.z, m[3]); }
inline mat4 translate(const mat4& m, const vec3& v) {
    mat4 r = m;
    r[3] = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3];
    return r;
}
inline mat4 rotate(const mat4& m, float angle, const vec3& v) {
    const float c = ez_cos(angle), s = ez_sin(angle);
    vec3 axis = normalize(v);
    vec3 temp = axis * (1.0f - c);
    float R[3][3];
    R[0][0] = c + temp.x * axis.x;
    R[0][1] = temp.x * axis.y + s * axis.z;
    R[0][2] = temp.x * axis.z - s * axis.y;
    R[1][0] = temp.y * axis.x - s * axis.z;
    R[1][1] = c + temp.y * axis.y;
    R[1][2] = temp.y * axis.z + s * axis.x;
    R[2][0] = temp.z * axis.x + s * axis.y;
    R[2][1] = temp.z * axis.y - s * axis.x;
    R[2][2] = c + temp.z * axis.z;
    mat4 r;
    for (int i = 0; i < 3; i++) r[i] = (m[0] * R[i][0] + m[1] * R[i][1]) + m[2] * R[i][2];
    r[3] = m[3];
    return r;
}
// display() only (never executed by the shim; the product's camera is tested against the oracle elsewhere)
inline mat4 lookAt(const vec3& eye, const vec3& center, const vec3& up) {
    vec3 f = normalize(center - eye), s = normalize(cross(f, up)), u = cross(s, f);
    mat4 r;
    r[0][0] = s.x; r[1][0] = s.y; r[2][0] = s.z;
    r[0][1] = u.x; r[1][1] = u.y; r[2][1] = u.z;
    r[0][2] = -f.x; r[1][2] = -f.y; r[2][2] = -f.z;
    r[3][0] = -dot(s, eye); r[3][1] = -dot(u, eye); r[3][2] = dot(f, eye);
    return r;
}
inline mat4 inverse(const mat4& m) { return m; }  // declared for display(); not executed

inline const float* value_ptr(const vec3& v) { return &v.x; }
inline const float* value_ptr(const mat4& m) { return &m.c[0].x; }

}  // namespace glm

#endif
