// Minimal stand-in for the un-vendored g-truc/glm headers.
//
// TEST INFRASTRUCTURE ONLY.  The reference's dptr kernels include <glm/glm.hpp>
// but the glm submodule (src/submodules/dptr/.gitmodules:1-3 -> third_party/glm)
// is not fetched in /root/reference and no version is pinned.  This file is our
// own from-scratch implementation of the handful of glm operations those
// kernels use (column-major mat3, vec3/vec4, transpose, dot, max) so that the
// UNMODIFIED reference .cu files can be compiled into oracle/_ref/ and run on
// the GPU box as the real-reference arm of the parity tests.  It is never
// included by the product (splatter_a_video_b200/csrc).
//
// Arithmetic order follows glm's documented definitions: mat3*mat3 is
// result[c][r] = a[0][r]*b[c][0] + a[1][r]*b[c][1] + a[2][r]*b[c][2].
#pragma once
#include <cuda_runtime.h>

#define GLM_SHIM_FN __host__ __device__ __forceinline__

namespace glm {

struct vec3 {
    float x, y, z;
    GLM_SHIM_FN vec3() : x(0.f), y(0.f), z(0.f) {}
    GLM_SHIM_FN explicit vec3(float s) : x(s), y(s), z(s) {}
    GLM_SHIM_FN vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    GLM_SHIM_FN float &operator[](int i) { return (&x)[i]; }
    GLM_SHIM_FN const float &operator[](int i) const { return (&x)[i]; }
    GLM_SHIM_FN vec3 &operator+=(const vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLM_SHIM_FN vec3 &operator-=(const vec3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    GLM_SHIM_FN vec3 &operator+=(float s) { x += s; y += s; z += s; return *this; }
    GLM_SHIM_FN vec3 &operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
    float x, y, z, w;
    GLM_SHIM_FN vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    GLM_SHIM_FN vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    GLM_SHIM_FN float &operator[](int i) { return (&x)[i]; }
    GLM_SHIM_FN const float &operator[](int i) const { return (&x)[i]; }
};

GLM_SHIM_FN vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_SHIM_FN vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_SHIM_FN vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
GLM_SHIM_FN vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLM_SHIM_FN vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLM_SHIM_FN vec3 operator*(const vec3 &a, const vec3 &b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_SHIM_FN vec3 operator+(const vec3 &a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }

GLM_SHIM_FN float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GLM_SHIM_FN vec3 max(const vec3 &a, float s) {
    return vec3(a.x < s ? s : a.x, a.y < s ? s : a.y, a.z < s ? s : a.z);
}

// Column-major 3x3: m[c] is column c, m[c][r] is (row r, column c).
struct mat3 {
    vec3 c[3];
    GLM_SHIM_FN mat3() { c[0] = vec3(1.f, 0.f, 0.f); c[1] = vec3(0.f, 1.f, 0.f); c[2] = vec3(0.f, 0.f, 1.f); }
    GLM_SHIM_FN explicit mat3(float d) { c[0] = vec3(d, 0.f, 0.f); c[1] = vec3(0.f, d, 0.f); c[2] = vec3(0.f, 0.f, d); }
    GLM_SHIM_FN mat3(float x0, float y0, float z0,
                     float x1, float y1, float z1,
                     float x2, float y2, float z2) {
        c[0] = vec3(x0, y0, z0); c[1] = vec3(x1, y1, z1); c[2] = vec3(x2, y2, z2);
    }
    GLM_SHIM_FN vec3 &operator[](int i) { return c[i]; }
    GLM_SHIM_FN const vec3 &operator[](int i) const { return c[i]; }
};

GLM_SHIM_FN mat3 operator*(const mat3 &a, const mat3 &b) {
    mat3 r(0.f);
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 3; ++row)
            r[col][row] = a[0][row] * b[col][0] + a[1][row] * b[col][1] + a[2][row] * b[col][2];
    return r;
}
GLM_SHIM_FN mat3 operator*(float s, const mat3 &a) {
    mat3 r(0.f);
    for (int col = 0; col < 3; ++col) r[col] = s * a[col];
    return r;
}
GLM_SHIM_FN mat3 operator*(const mat3 &a, float s) { return s * a; }
GLM_SHIM_FN vec3 operator*(const mat3 &a, const vec3 &v) {
    return vec3(a[0][0] * v.x + a[1][0] * v.y + a[2][0] * v.z,
                a[0][1] * v.x + a[1][1] * v.y + a[2][1] * v.z,
                a[0][2] * v.x + a[1][2] * v.y + a[2][2] * v.z);
}
GLM_SHIM_FN mat3 transpose(const mat3 &a) {
    return mat3(a[0][0], a[1][0], a[2][0],
                a[0][1], a[1][1], a[2][1],
                a[0][2], a[1][2], a[2][2]);
}

} // namespace glm
