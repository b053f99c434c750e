// Minimal stand-in for the handful of glm symbols the reference rasterizer uses.
//
// The reference does not vendor or pin glm (its .gitignore:212 excludes third_party/glm; README.md:60,88
// asks for `apt install libglm-dev` or a HEAD clone) and glm is not installed in this image.  This header
// is OUR code (not copied from glm): column-major mat3 like glm (m[i] is column i, m[i][j] is row j of
// column i), products summed k = 0,1,2 left to right.  It exists only so that oracle/Makefile can compile
// the reference's own .cu files, unmodified, into oracle/_ref/ (test / baseline infrastructure).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GLMS_HD __host__ __device__ inline
#else
#define GLMS_HD inline
#endif

namespace glm {

struct vec3 {
    float x, y, z;
    GLMS_HD vec3() : x(0), y(0), z(0) {}
    GLMS_HD vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    GLMS_HD explicit vec3(float a) : x(a), y(a), z(a) {}
    GLMS_HD float& operator[](int i) { return (&x)[i]; }
    GLMS_HD const float& operator[](int i) const { return (&x)[i]; }
    GLMS_HD vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLMS_HD vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
    GLMS_HD vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
    float x, y, z, w;
    GLMS_HD vec4() : x(0), y(0), z(0), w(0) {}
    GLMS_HD vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};

GLMS_HD vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLMS_HD vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLMS_HD vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLMS_HD vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLMS_HD vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
GLMS_HD float dot(const vec3& a, const vec3& b) { float t0 = a.x * b.x, t1 = a.y * b.y, t2 = a.z * b.z; return t0 + t1 + t2; }
GLMS_HD float length(const vec3& a) { return sqrtf(dot(a, a)); }
GLMS_HD vec3 max(const vec3& a, float s) { return vec3(fmaxf(a.x, s), fmaxf(a.y, s), fmaxf(a.z, s)); }

struct mat3 {
    vec3 c[3];  // columns
    GLMS_HD mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    GLMS_HD explicit mat3(float d) { c[0] = vec3(d, 0, 0); c[1] = vec3(0, d, 0); c[2] = vec3(0, 0, d); }
    GLMS_HD mat3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2) {
        c[0] = vec3(x0, y0, z0); c[1] = vec3(x1, y1, z1); c[2] = vec3(x2, y2, z2);
    }
    GLMS_HD vec3& operator[](int i) { return c[i]; }
    GLMS_HD const vec3& operator[](int i) const { return c[i]; }
};

GLMS_HD mat3 transpose(const mat3& m) {
    return mat3(m[0][0], m[1][0], m[2][0], m[0][1], m[1][1], m[2][1], m[0][2], m[1][2], m[2][2]);
}
// (a*b)[col][row] = sum_k a[k][row] * b[col][k]
GLMS_HD mat3 operator*(const mat3& a, const mat3& b) {
    mat3 r(0.0f);
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++)
            r[col][row] = a[0][row] * b[col][0] + a[1][row] * b[col][1] + a[2][row] * b[col][2];
    return r;
}
GLMS_HD vec3 operator*(const mat3& a, const vec3& v) {
    return vec3(a[0][0] * v.x + a[1][0] * v.y + a[2][0] * v.z,
                a[0][1] * v.x + a[1][1] * v.y + a[2][1] * v.z,
                a[0][2] * v.x + a[1][2] * v.y + a[2][2] * v.z);
}
GLMS_HD mat3 operator*(float s, const mat3& a) {
    mat3 r(0.0f);
    for (int i = 0; i < 3; i++) r[i] = s * a[i];
    return r;
}

}  // namespace glm
