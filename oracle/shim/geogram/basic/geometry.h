/* oracle/shim -- TEST INFRASTRUCTURE ONLY; see geogram/basic/common.h in this directory. */
#pragma once
#include <geogram/basic/common.h>
namespace GEO {
struct vec3 {
    double x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    double& operator[](index_t i) { return (&x)[i]; }
    const double& operator[](index_t i) const { return (&x)[i]; }
    double* data() { return &x; }
    const double* data() const { return &x; }
};
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
template <class T2> inline vec3 operator*(const vec3& a, T2 s) { return vec3(a.x * double(s), a.y * double(s), a.z * double(s)); }
template <class T2> inline vec3 operator*(T2 s, const vec3& a) { return vec3(double(s) * a.x, double(s) * a.y, double(s) * a.z); }
template <class T2> inline vec3 operator/(const vec3& a, T2 s) { return vec3(a.x / double(s), a.y / double(s), a.z / double(s)); }
inline double dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline double length2(const vec3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double length(const vec3& a) { return ::sqrt(length2(a)); }
inline double distance2(const vec3& a, const vec3& b) { return length2(b - a); }
inline double distance(const vec3& a, const vec3& b) { return length(b - a); }
inline vec3 normalize(const vec3& v) {
    double s = length(v);
    if (s > 1e-30) s = 1.0 / s;
    return s * v;
}
struct Box {
    double xyz_min[3];
    double xyz_max[3];
    bool contains(const vec3& b) const {
        for (coord_index_t c = 0; c < 3; ++c)
            if (b[c] < xyz_min[c] || b[c] > xyz_max[c]) return false;
        return true;
    }
};
inline bool bboxes_overlap(const Box& B1, const Box& B2) {
    for (coord_index_t c = 0; c < 3; ++c) {
        if (B1.xyz_max[c] < B2.xyz_min[c]) return false;
        if (B1.xyz_min[c] > B2.xyz_max[c]) return false;
    }
    return true;
}
inline void bbox_union(Box& target, const Box& B1, const Box& B2) {
    for (coord_index_t c = 0; c < 3; ++c) {
        target.xyz_min[c] = std::min(B1.xyz_min[c], B2.xyz_min[c]);
        target.xyz_max[c] = std::max(B1.xyz_max[c], B2.xyz_max[c]);
    }
}
namespace Geom {
inline double distance2(const vec3& a, const vec3& b) { return GEO::distance2(a, b); }
inline double tetra_signed_volume(const vec3& p, const vec3& q, const vec3& r, const vec3& s) {
    return dot(q - p, cross(r - p, s - p)) / 6.0;
}
}
}
