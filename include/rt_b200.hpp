// rt_b200.hpp — C++ mirror of the crate's scene-construction surface, header-only over the
// C ABI in rt_b200.h. Names and argument meaning follow the reference one for one so that a
// scene written against the crate (main.rs:56-639) transliterates line by line:
//
//   Rust                                               here
//   Arc::new(SolidColor::new(r,g,b))                   s.SolidColor(r,g,b)
//   Arc::new(Lambertian::new(tex))                     s.Lambertian(tex)
//   Arc::new(Sphere::new(c, r, mat))                   s.Sphere(c, r, mat)
//   Sphere::new(..).with_target(t)                     s.Sphere(c, r, mat, t)
//   Quad::cube(&a, &b, mat)                            s.cube(a, b, mat)
//   world.add(obj)                                     world.add(obj)
//   Arc::new(BVHNode::new(&mut list))                  s.BVHNode(list)
//   Camera::new(CameraSettings{..})                    rt::Camera(settings)
//
// Handles (Texture/Material/Hittable) are ids into the scene being built; they play the role
// of the reference's Arc<dyn Trait>. Errors throw rt::Error on this side of the ABI only.
#ifndef RT_B200_HPP
#define RT_B200_HPP

#include "rt_b200.h"

#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace rt {

using FP = double;  // common.rs:1

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

inline int check(int rc) {
    if (rc < 0) throw Error(rc, rt_last_error());
    return rc;
}

struct Vec3 {  // vec3.rs:9-16
    FP x, y, z;
    constexpr Vec3() : x(0), y(0), z(0) {}
    constexpr Vec3(FP x_, FP y_, FP z_) : x(x_), y(y_), z(z_) {}
    static constexpr Vec3 splat(FP v) { return Vec3(v, v, v); }
    static constexpr Vec3 ZERO() { return splat(0.0); }
    static constexpr Vec3 ONE() { return splat(1.0); }
    static constexpr Vec3 RIGHT() { return Vec3(1, 0, 0); }
    static constexpr Vec3 UP() { return Vec3(0, 1, 0); }
    static constexpr Vec3 FORWARD() { return Vec3(0, 0, 1); }
    static constexpr Vec3 LEFT() { return Vec3(-1, 0, 0); }
    static constexpr Vec3 DOWN() { return Vec3(0, -1, 0); }
    static constexpr Vec3 BACKWARD() { return Vec3(0, 0, -1); }
    constexpr Vec3 operator+(const Vec3& o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
    constexpr Vec3 operator-(const Vec3& o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
    constexpr Vec3 operator-() const { return Vec3(-x, -y, -z); }
    constexpr Vec3 operator*(const Vec3& o) const { return Vec3(x * o.x, y * o.y, z * o.z); }
    constexpr Vec3 operator*(FP s) const { return Vec3(x * s, y * s, z * s); }
    constexpr Vec3 operator/(FP s) const { return *this * (1.0 / s); }  // vec3.rs:244-249
    FP dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
    FP length_squared() const { return dot(*this); }
    FP length() const { return std::sqrt(dot(*this)); }
};
using Point3 = Vec3;
using Color = Vec3;
inline constexpr Vec3 operator*(FP s, const Vec3& v) { return v * s; }

struct Texture { int id = -1; };
struct Material { int id = -1; };
struct Hittable { int id = -1; };

struct HittableList {  // hittable.rs:50-59
    std::vector<Hittable> objects;
    void add(Hittable h) { objects.push_back(h); }
};

class Scene {
  public:
    explicit Scene(uint64_t bvh_seed = 2) { check(rt_builder_create(bvh_seed, &b_)); }
    ~Scene() { rt_builder_destroy(b_); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    rt_builder* raw() { return b_; }
    rt_builder* release() { rt_builder* b = b_; b_ = nullptr; return b; }

    // texture.rs
    Texture SolidColor(FP r, FP g, FP bl) { return {check(rt_tex_solid(b_, r, g, bl))}; }
    Texture SolidColor(const Color& c) { return SolidColor(c.x, c.y, c.z); }  // From<Color>, texture.rs:27
    Texture CheckerTexture(FP scale, Texture even, Texture odd) { return {check(rt_tex_checker(b_, scale, even.id, odd.id))}; }
    Texture CheckerTexture(FP scale, const Color& even, const Color& odd) {  // new_from_colors, texture.rs:51
        return CheckerTexture(scale, SolidColor(even), SolidColor(odd));
    }
    Texture ImageTexture(int width, int height, const uint8_t* rgb8) { return {check(rt_tex_image(b_, width, height, rgb8))}; }
    Texture NoiseTexture(FP scale, uint64_t perlin_seed) { return {check(rt_tex_noise(b_, scale, perlin_seed))}; }

    // material.rs
    Material Lambertian(Texture albedo) { return {check(rt_mat_lambertian(b_, albedo.id))}; }
    Material Metal(const Color& albedo, FP fuzz) {
        const double a[3] = {albedo.x, albedo.y, albedo.z};
        return {check(rt_mat_metal(b_, a, fuzz))};
    }
    Material Dielectric(FP ir) { return {check(rt_mat_dielectric(b_, ir))}; }
    Material DiffuseLight(Texture emit) { return {check(rt_mat_diffuse_light(b_, emit.id))}; }
    Material Isotropic(Texture albedo) { return {check(rt_mat_isotropic(b_, albedo.id))}; }

    // sphere.rs, quad.rs, hittable.rs, constant_medium.rs, bvh.rs
    Hittable Sphere(const Point3& c, FP radius, Material m) {
        const double cc[3] = {c.x, c.y, c.z};
        return {check(rt_hit_sphere(b_, cc, radius, m.id))};
    }
    Hittable Sphere(const Point3& c, FP radius, Material m, const Point3& target) {  // .with_target
        const double cc[3] = {c.x, c.y, c.z}, tt[3] = {target.x, target.y, target.z};
        return {check(rt_hit_moving_sphere(b_, cc, tt, radius, m.id))};
    }
    Hittable Quad(const Point3& q, const Vec3& u, const Vec3& v, Material m) {
        const double qq[3] = {q.x, q.y, q.z}, uu[3] = {u.x, u.y, u.z}, vv[3] = {v.x, v.y, v.z};
        return {check(rt_hit_quad(b_, qq, uu, vv, m.id))};
    }
    Hittable cube(const Point3& a, const Point3& bb, Material m) {  // Quad::cube
        const double aa[3] = {a.x, a.y, a.z}, bbb[3] = {bb.x, bb.y, bb.z};
        return {check(rt_hit_cube(b_, aa, bbb, m.id))};
    }
    Hittable List(const HittableList& l) {
        std::vector<int> ids;
        for (auto h : l.objects) ids.push_back(h.id);
        return {check(rt_hit_list(b_, ids.data(), (int)ids.size()))};
    }
    Hittable Translate(Hittable obj, const Vec3& offset) {
        const double o[3] = {offset.x, offset.y, offset.z};
        return {check(rt_hit_translate(b_, obj.id, o))};
    }
    Hittable RotateY(Hittable obj, FP angle) { return {check(rt_hit_rotate_y(b_, obj.id, angle))}; }
    Hittable ConstantMedium(Hittable boundary, FP density, Texture albedo) {
        return {check(rt_hit_constant_medium(b_, boundary.id, density, albedo.id))};
    }
    Hittable ConstantMedium(Hittable boundary, FP density, const Color& albedo) {  // new_from_color
        return ConstantMedium(boundary, density, SolidColor(albedo));
    }
    Hittable BVHNode(const HittableList& l) {  // BVHNode::new(&mut list)
        std::vector<int> ids;
        for (auto h : l.objects) ids.push_back(h.id);
        return {check(rt_hit_bvh(b_, ids.data(), (int)ids.size()))};
    }

    rt_scene_desc finish(Hittable world) {
        rt_scene_desc d;
        check(rt_builder_finish(b_, world.id, &d));
        return d;
    }

  private:
    rt_builder* b_ = nullptr;
};

struct CameraSettings : rt_camera_settings {  // camera.rs:8-37
    CameraSettings() { rt_camera_settings_default(this); }
};

struct Camera : rt_camera_desc {  // camera.rs:38-110
    explicit Camera(const rt_camera_settings& s) { check(rt_camera_new(&s, this)); }
};

}  // namespace rt

#endif
