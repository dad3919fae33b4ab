// The nine CLI scenes of the reference (main.rs:56-639, dispatch main.rs:645-656) written
// against the C++ mirror API. Geometry, materials and camera literals are the reference's;
// its unseeded layout RNG is replaced by a seeded one with the same distributions.
#include "../../../include/rt_b200.hpp"
#include "host_common.h"
#include "host_rng.h"

#include <cstring>

using namespace rt;

namespace {

struct Ctx {
    Scene& s;
    rt_host::HostRng rng;
    const rt_scene_request& req;
    Color random_color() {  // Color::random(), vec3.rs:42-44
        const double x = rng.random(), y = rng.random(), z = rng.random();
        return Color(x, y, z);
    }
    Color random_color_range(double lo, double hi) {  // Vec3::random_range, vec3.rs:46-52
        const double x = rng.range(lo, hi), y = rng.range(lo, hi), z = rng.range(lo, hi);
        return Color(x, y, z);
    }
    Texture earth_texture() {  // ImageTexture::new("assets/earth-large.jpg"), main.rs:178,590
        if (!req.earth_rgb8 || req.earth_width <= 0 || req.earth_height <= 0)
            throw Error(RT_ERR_INVALID_ARGUMENT, "this scene needs the decoded earth image (earth_rgb8/earth_width/earth_height)");
        return s.ImageTexture(req.earth_width, req.earth_height, req.earth_rgb8);
    }
};

void set3(double* d, const Vec3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

// main.rs:56-138
HittableList random_balls(Ctx& c, CameraSettings& cam) {
    Scene& s = c.s;
    HittableList world;
    Material ground_material = s.Lambertian(s.SolidColor(Color::splat(0.5)));
    world.add(s.Sphere(Point3::DOWN() * 1000.0, 1000.0, ground_material));

    for (int a = -11; a < 11; ++a) {
        for (int b = -11; b < 11; ++b) {
            const FP choose_mat = c.rng.random();
            const FP cx = (FP)a + 0.9 * c.rng.random();
            const FP cz = (FP)b + 0.9 * c.rng.random();
            const Point3 center(cx, 0.2, cz);
            if ((center - Point3(4.0, 0.2, 0.0)).length() > 0.9) {
                if (choose_mat < 0.8) {
                    const Color c1 = c.random_color();
                    const Color c2 = c.random_color();
                    Texture color = s.SolidColor(c1 * c2);
                    const Point3 target = center + Vec3::UP() * c.rng.random() * 0.5;
                    world.add(s.Sphere(center, 0.2, s.Lambertian(color), target));
                } else if (choose_mat < 0.95) {
                    const Color albedo = c.random_color_range(0.5, 1.0);
                    const FP fuzz = c.rng.range(0.0, 0.5);
                    world.add(s.Sphere(center, 0.2, s.Metal(albedo, fuzz)));
                } else {
                    world.add(s.Sphere(center, 0.2, s.Dielectric(1.5)));
                }
            }
        }
    }
    world.add(s.Sphere(Point3(0.0, 1.0, 0.0), 1.0, s.Dielectric(1.5)));
    world.add(s.Sphere(Point3(-4.0, 1.0, 0.0), 1.0, s.Lambertian(s.SolidColor(0.4, 0.2, 0.1))));
    world.add(s.Sphere(Point3(4.0, 1.0, 0.0), 1.0, s.Metal(Color(0.7, 0.6, 0.5), 0.0)));

    cam.aspect_ratio = 16.0 / 9.0;
    cam.image_width = 600;
    cam.samples_per_pixel = 128;
    cam.max_depth = 8;
    set3(cam.background, Color(0.7, 0.8, 1.0));
    cam.vfov = 20.0;
    set3(cam.look_from, Point3(13.0, 2.0, 3.0));
    set3(cam.look_at, Point3::ZERO());
    cam.defocus_angle = 0.6;
    cam.focus_dist = 10.0;
    return world;
}

void book1_camera(CameraSettings& cam) {  // shared literal of main.rs:158-170,222-234
    cam.aspect_ratio = 16.0 / 9.0;
    cam.image_width = 1200;
    cam.samples_per_pixel = 128;
    cam.max_depth = 8;
    set3(cam.background, Color(0.7, 0.8, 1.0));
    cam.vfov = 20.0;
    set3(cam.look_from, Point3(13.0, 2.0, 3.0));
    set3(cam.look_at, Point3::ZERO());
}

// main.rs:140-173
HittableList two_spheres(Ctx& c, CameraSettings& cam) {
    Scene& s = c.s;
    HittableList world;
    Material checker = s.Lambertian(s.CheckerTexture(0.32, Color(0.2, 0.3, 0.1), Color::splat(0.9)));
    world.add(s.Sphere(Point3(0.0, -10.0, 0.0), 10.0, checker));
    world.add(s.Sphere(Point3(0.0, 10.0, 0.0), 10.0, checker));
    book1_camera(cam);
    return world;
}

// main.rs:175-203
HittableList earth(Ctx& c, CameraSettings& cam) {
    Scene& s = c.s;
    HittableList world;
    Material earth_texture = s.Lambertian(c.earth_texture());
    world.add(s.Sphere(Point3(0.0, 0.0, 0.0), 2.0, earth_texture));
    book1_camera(cam);
    set3(cam.look_from, Point3(12.0, 0.0, 0.0));
    return world;
}

// main.rs:205-237
HittableList two_perlin_spheres(Ctx& c, CameraSettings& cam) {
    Scene& s = c.s;
    HittableList world;
    Material perlin_texture = s.Lambertian(s.NoiseTexture(4.0, c.req.perlin_seed));
    world.add(s.Sphere(Point3(0.0, -1000.0, 0.0), 1000.0, perlin_texture));
    world.add(s.Sphere(Point3(0.0, 2.0, 0.0), 2.0, perlin_texture));
    book1_camera(cam);
    return world;
}

// main.rs:239-294
HittableList quads(Ctx& c, CameraSettings& cam) {
    Scene& s = c.s;
    HittableList world;
    Material left_red = s.Lambertian(s.SolidColor(1.0, 0.2, 0.2));
    Material back_green = s.Lambertian(s.SolidColor(0.2, 1.0, 0.2));
    Material right_blue = s.Lambertian(s.SolidColor(0.2, 0.2, 1.0));
    Material upper_orange = s.Lambertian(s.SolidColor(1.0, 0.5, 0.0));
    Material lower_teal = s.Lambertian(s.SolidColor(0.2, 0.8, 0.8));
    world.add(s.Quad(Point3(-3.0, -2.0, 5.0), Vec3::BACKWARD() * 4.0, Vec3::UP() * 4.0, left_red));
    world.add(s.Quad(Point3(-2.0, -2.0, 0.0), Vec3::RIGHT() * 4.0, Vec3::UP() * 4.0, back_green));
    world.add(s.Quad(Point3(3.0, -2.0, 1.0), Vec3::FORWARD() * 4.0, Vec3::UP() * 4.0, right_blue));
    world.add(s.Quad(Point3(-2.0, 3.0, 1.0), Vec3::RIGHT() * 4.0, Vec3::FORWARD() * 4.0, upper_orange));
    world.add(s.Quad(Point3(-2.0, -3.0, 5.0), Vec3::RIGHT() * 4.0, Vec3::BACKWARD() * 4.0, lower_teal));
    cam.aspect_ratio = 1.0;
    cam.image_width = 1200;
    cam.samples_per_pixel = 128;
    cam.max_depth = 8;
    set3(cam.background, Color(0.7, 0.8, 1.0));
    cam.vfov = 80.0;
    set3(cam.look_from, Point3::FORWARD() * 9.0);
    set3(cam.look_at, Point3::ZERO());
    return world;
}

// main.rs:296-342
HittableList simple_light(Ctx& c, CameraSettings& cam) {
    Scene& s = c.s;
    HittableList world;
    Texture perlin_texture = s.NoiseTexture(4.0, c.req.perlin_seed);
    world.add(s.Sphere(Point3(0.0, -1000.0, 0.0), 1000.0, s.Lambertian(perlin_texture)));
    world.add(s.Sphere(Point3(0.0, 2.0, 0.0), 2.0, s.Lambertian(perlin_texture)));
    Material diffuse_light = s.DiffuseLight(s.SolidColor(4.0, 4.0, 4.0));
    world.add(s.Quad(Point3(3.0, 1.0, -2.0), Vec3::RIGHT() * 2.0, Vec3::UP() * 2.0, diffuse_light));
    world.add(s.Sphere(Point3(0.0, 7.0, 0.0), 2.0, diffuse_light));
    cam.aspect_ratio = 16.0 / 9.0;
    cam.image_width = 600;
    cam.samples_per_pixel = 1024;
    cam.max_depth = 8;
    set3(cam.background, Color::ZERO());
    cam.vfov = 20.0;
    set3(cam.look_from, Point3(26.0, 3.0, 6.0));
    set3(cam.look_at, Point3::UP() * 2.0);
    return world;
}

void cornell_camera(CameraSettings& cam) {  // main.rs:406-418,491-503
    cam.aspect_ratio = 1.0;
    cam.image_width = 600;
    cam.samples_per_pixel = 4096;
    cam.max_depth = 8;
    set3(cam.background, Color::ZERO());
    cam.vfov = 40.0;
    set3(cam.look_from, Point3(278.0, 278.0, -800.0));
    set3(cam.look_at, Point3(278.0, 278.0, 0.0));
}

// main.rs:344-421
HittableList cornell_box(Ctx& c, CameraSettings& cam) {
    Scene& s = c.s;
    HittableList world;
    Material red = s.Lambertian(s.SolidColor(0.65, 0.05, 0.05));
    Material white = s.Lambertian(s.SolidColor(0.73, 0.73, 0.73));
    Material green = s.Lambertian(s.SolidColor(0.12, 0.45, 0.15));
    Material light = s.DiffuseLight(s.SolidColor(15.0, 15.0, 15.0));
    world.add(s.Quad(Point3(555.0, 0.0, 555.0), Vec3::UP() * 555.0, Vec3::BACKWARD() * 555.0, green));
    world.add(s.Quad(Point3::ZERO(), Vec3::UP() * 555.0, Vec3::FORWARD() * 555.0, red));
    world.add(s.Quad(Point3(343.0, 554.0, 332.0), Vec3::LEFT() * 130.0, Vec3::BACKWARD() * 105.0, light));
    world.add(s.Quad(Point3::FORWARD() * 555.0, Vec3::RIGHT() * 555.0, Vec3::BACKWARD() * 555.0, white));
    world.add(s.Quad(Point3::ONE() * 555.0, Vec3::LEFT() * 555.0, Vec3::BACKWARD() * 555.0, white));
    world.add(s.Quad(Point3(555.0, 0.0, 555.0), Vec3::LEFT() * 555.0, Vec3::UP() * 555.0, white));

    Hittable box1 = s.cube(Point3::ZERO(), Point3(165.0, 330.0, 165.0), white);
    box1 = s.RotateY(box1, 15.0);
    box1 = s.Translate(box1, Vec3(265.0, 0.0, 295.0));
    world.add(box1);

    Hittable box2 = s.cube(Point3::ZERO(), Point3::splat(165.0), white);
    box2 = s.RotateY(box2, -18.0);
    box2 = s.Translate(box2, Vec3(130.0, 0.0, 65.0));
    world.add(box2);

    cornell_camera(cam);
    return world;
}

// main.rs:423-506
HittableList cornell_smoke(Ctx& c, CameraSettings& cam) {
    Scene& s = c.s;
    HittableList world;
    Material red = s.Lambertian(s.SolidColor(0.65, 0.05, 0.05));
    Material white = s.Lambertian(s.SolidColor(0.73, 0.73, 0.73));
    Material green = s.Lambertian(s.SolidColor(0.12, 0.45, 0.15));
    Material light = s.DiffuseLight(s.SolidColor(7.0, 7.0, 7.0));
    world.add(s.Quad(Point3(555.0, 0.0, 555.0), Vec3::UP() * 555.0, Vec3::BACKWARD() * 555.0, green));
    world.add(s.Quad(Point3::ZERO(), Vec3::UP() * 555.0, Vec3::FORWARD() * 555.0, red));
    world.add(s.Quad(Point3(113.0, 554.0, 127.0), Vec3::RIGHT() * 330.0, Vec3::FORWARD() * 305.0, light));
    world.add(s.Quad(Point3::FORWARD() * 555.0, Vec3::RIGHT() * 555.0, Vec3::BACKWARD() * 555.0, white));
    world.add(s.Quad(Point3::ONE() * 555.0, Vec3::LEFT() * 555.0, Vec3::BACKWARD() * 555.0, white));
    world.add(s.Quad(Point3(555.0, 0.0, 555.0), Vec3::LEFT() * 555.0, Vec3::UP() * 555.0, white));

    Hittable box1 = s.cube(Point3::ZERO(), Point3(165.0, 330.0, 165.0), white);
    box1 = s.RotateY(box1, 15.0);
    box1 = s.Translate(box1, Vec3(265.0, 0.0, 295.0));
    world.add(s.ConstantMedium(box1, 0.01, Color::ZERO()));

    Hittable box2 = s.cube(Point3::ZERO(), Point3::splat(165.0), white);
    box2 = s.RotateY(box2, -18.0);
    box2 = s.Translate(box2, Vec3(130.0, 0.0, 65.0));
    world.add(s.ConstantMedium(box2, 0.01, Color::ONE()));

    cornell_camera(cam);
    return world;
}

// main.rs:508-639
HittableList final_scene(Ctx& c, CameraSettings& cam) {
    Scene& s = c.s;
    HittableList world;
    Material ground = s.Lambertian(s.SolidColor(0.48, 0.83, 0.53));
    HittableList boxes1;
    for (int i = 0; i < 20; ++i) {
        for (int j = 0; j < 20; ++j) {
            const FP side = 100.0;
            const FP x0 = -1000.0 + (FP)i * side;
            const FP x1 = x0 + side;
            const FP z0 = -1000.0 + (FP)j * side;
            const FP z1 = z0 + side;
            const FP y0 = 0.0;
            const FP y1 = c.rng.range(1.0, 101.0);
            boxes1.add(s.cube(Point3(x0, y0, z0), Point3(x1, y1, z1), ground));
        }
    }
    // Green ground boxes
    world.add(s.BVHNode(boxes1));
    // Light source
    world.add(s.Quad(Point3(123.0, 554.0, 147.0), Vec3::RIGHT() * 300.0, Vec3::FORWARD() * 265.0,
                     s.DiffuseLight(s.SolidColor(7.0, 7.0, 7.0))));
    // Motion-blurred sphere
    const Point3 center1(400.0, 400.0, 200.0);
    const Point3 center2 = center1 + Vec3::RIGHT() * 30.0;
    Material sphere_material = s.Lambertian(s.SolidColor(0.7, 0.3, 0.1));
    world.add(s.Sphere(center1, 50.0, sphere_material, center2));
    // Glass sphere
    world.add(s.Sphere(Point3(260.0, 150.0, 45.0), 50.0, s.Dielectric(1.5)));
    // Fuzzy metal sphere
    world.add(s.Sphere(Point3(0.0, 150.0, 145.0), 50.0, s.Metal(Color(0.8, 0.8, 0.9), 1.0)));
    // Subsurface-scattering sphere: the glass boundary is in the world AND bounds a medium
    Hittable boundary = s.Sphere(Point3(360.0, 150.0, 145.0), 70.0, s.Dielectric(1.5));
    world.add(boundary);
    world.add(s.ConstantMedium(boundary, 0.2, Color(0.2, 0.4, 0.9)));
    // Global scene fog
    boundary = s.Sphere(Point3::ZERO(), 5000.0, s.Dielectric(1.5));
    world.add(s.ConstantMedium(boundary, 0.0001, Color::ONE()));
    // Earth sphere
    Material earth_material = s.Lambertian(c.earth_texture());
    world.add(s.Sphere(Point3(400.0, 200.0, 400.0), 100.0, earth_material));
    // Noise sphere
    Texture perlin_texture = s.NoiseTexture(0.1, c.req.perlin_seed);
    world.add(s.Sphere(Point3(220.0, 280.0, 300.0), 80.0, s.Lambertian(perlin_texture)));
    // Box of spheres
    HittableList boxes2;
    Material white = s.Lambertian(s.SolidColor(0.73, 0.73, 0.73));
    for (int k = 0; k < 1000; ++k) {
        const Color p = c.random_color_range(0.0, 165.0);
        boxes2.add(s.Sphere(p, 10.0, white));
    }
    world.add(s.Translate(s.RotateY(s.BVHNode(boxes2), 15.0), Vec3(-100.0, 270.0, 395.0)));

    cam.aspect_ratio = 1.0;
    cam.image_width = 800;
    cam.samples_per_pixel = 8192;
    cam.max_depth = 40;
    set3(cam.background, Color::ZERO());
    cam.vfov = 40.0;
    set3(cam.look_from, Point3(478.0, 278.0, -600.0));
    set3(cam.look_at, Point3(278.0, 278.0, 0.0));
    return world;
}

}  // namespace

extern "C" int rt_scene_builtin(const rt_scene_request* req, rt_builder** builder_out, rt_scene_desc* scene_out,
                                rt_camera_settings* settings_out) {
    if (!req || !builder_out || !scene_out || !settings_out)
        return rt_host::fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_builtin: null argument");
    try {
        Scene s(req->bvh_seed);
        Ctx c{s, rt_host::HostRng(req->scene_seed), *req};
        CameraSettings cam;
        HittableList world;
        switch (req->scene) {  // main.rs:645-656
            case 1: world = two_spheres(c, cam); break;
            case 2: world = earth(c, cam); break;
            case 3: world = two_perlin_spheres(c, cam); break;
            case 4: world = quads(c, cam); break;
            case 5: world = simple_light(c, cam); break;
            case 6: world = cornell_box(c, cam); break;
            case 7: world = cornell_smoke(c, cam); break;
            case 8: world = final_scene(c, cam); break;
            case 0:
            default: world = random_balls(c, cam); break;
        }
        Hittable bvh = s.BVHNode(world);  // main.rs:659
        if (req->image_width > 0) cam.image_width = req->image_width;
        if (req->samples_per_pixel > 0) cam.samples_per_pixel = req->samples_per_pixel;
        if (req->max_depth > 0) cam.max_depth = req->max_depth;
        *scene_out = s.finish(bvh);
        std::memcpy(settings_out, static_cast<rt_camera_settings*>(&cam), sizeof(rt_camera_settings));
        *builder_out = s.release();
        return RT_OK;
    } catch (const Error& e) {
        return rt_host::fail(e.status, e.what());
    } catch (const std::exception& e) {
        return rt_host::fail(RT_ERR_INTERNAL, e.what());
    }
}
