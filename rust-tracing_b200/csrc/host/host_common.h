// Small f64 vector type for the host-side scene code, with the reference's operator
// semantics (vec3.rs): Vec3 / FP multiplies by the reciprocal (vec3.rs:244-249) and
// normalize multiplies by length().recip() (vec3.rs:119-131).
#pragma once
#include <cmath>
#include <string>

namespace rt_host {

extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    explicit V3(const double* p) : x(p[0]), y(p[1]), z(p[2]) {}
    static V3 splat(double v) { return V3(v, v, v); }
    V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
    V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
    V3 operator-() const { return V3(-x, -y, -z); }
    V3 operator*(const V3& o) const { return V3(x * o.x, y * o.y, z * o.z); }
    V3 operator*(double s) const { return V3(x * s, y * s, z * s); }
    V3 operator/(double s) const { return *this * (1.0 / s); }
    double dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
    double length_squared() const { return dot(*this); }
    double length() const { return std::sqrt(dot(*this)); }
    V3 normalize() const { return *this * (1.0 / length()); }
    V3 cross(const V3& o) const { return V3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
    void store(double* p) const { p[0] = x; p[1] = y; p[2] = z; }
};

}  // namespace rt_host
