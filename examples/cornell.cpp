// examples/cornell.cpp — the Cornell box of main.rs:344-421 written against the C++ mirror (include/rt_b200.hpp),
// built and rendered through the C ABI. Build:
//   g++ -std=c++17 -Iinclude examples/cornell.cpp -Lrust-tracing_b200/csrc -lrt_b200 -Wl,-rpath,$PWD/rust-tracing_b200/csrc -o cornell
// Without a GPU it stops after printing the device layout (rt_context_create reports RT_ERR_NO_DEVICE: no CPU fallback).
#include "rt_b200.hpp"

#include <cstdio>
#include <vector>

using namespace rt;

int main(int argc, char** argv) {
    const int spp = argc > 1 ? std::atoi(argv[1]) : 64;
    try {
        Scene s(/*bvh_seed=*/2);
        HittableList world;
        Material red = s.Lambertian(s.SolidColor(0.65, 0.05, 0.05));
        Material white = s.Lambertian(s.SolidColor(0.73, 0.73, 0.73));
        Material green = s.Lambertian(s.SolidColor(0.12, 0.45, 0.15));
        Material light = s.DiffuseLight(s.SolidColor(15.0, 15.0, 15.0));
        world.add(s.Quad(Point3(555, 0, 555), Vec3::UP() * 555.0, Vec3::BACKWARD() * 555.0, green));
        world.add(s.Quad(Point3::ZERO(), Vec3::UP() * 555.0, Vec3::FORWARD() * 555.0, red));
        world.add(s.Quad(Point3(343, 554, 332), Vec3::LEFT() * 130.0, Vec3::BACKWARD() * 105.0, light));
        world.add(s.Quad(Point3::FORWARD() * 555.0, Vec3::RIGHT() * 555.0, Vec3::BACKWARD() * 555.0, white));
        world.add(s.Quad(Point3::ONE() * 555.0, Vec3::LEFT() * 555.0, Vec3::BACKWARD() * 555.0, white));
        world.add(s.Quad(Point3(555, 0, 555), Vec3::LEFT() * 555.0, Vec3::UP() * 555.0, white));
        world.add(s.Translate(s.RotateY(s.cube(Point3::ZERO(), Point3(165, 330, 165), white), 15.0), Vec3(265, 0, 295)));
        world.add(s.Translate(s.RotateY(s.cube(Point3::ZERO(), Point3::splat(165.0), white), -18.0), Vec3(130, 0, 65)));
        rt_scene_desc desc = s.finish(s.BVHNode(world));   // main.rs:659

        CameraSettings cs;
        cs.aspect_ratio = 1.0;
        cs.image_width = 200;
        cs.samples_per_pixel = spp;
        cs.max_depth = 8;
        cs.vfov = 40.0;
        cs.look_from[0] = 278; cs.look_from[1] = 278; cs.look_from[2] = -800;
        cs.look_at[0] = 278; cs.look_at[1] = 278; cs.look_at[2] = 0;
        Camera cam(cs);

        rt_layout_info info;
        check(rt_scene_layout(&desc, 0u, &info));
        std::printf("scene: %d hittables, %d bvh nodes -> %d stream words (%d inner, %d quad, %d box, %d instance ops)\n",
                    desc.n_hittables, desc.n_bvh_nodes, info.n_words, info.n_inner, info.n_quad, info.n_box, info.n_xform);

        rt_context* ctx = nullptr;
        if (rt_context_create(0, &ctx) < 0) {
            std::printf("no GPU: %s\n", rt_last_error());
            return 0;
        }
        rt_scene* dev = nullptr;
        check(rt_scene_upload(ctx, &desc, &dev));
        std::vector<float> sums((size_t)cam.image_width * cam.image_height * 4);
        check(rt_render(ctx, dev, &cam, 0, spp, /*seed=*/0, sums.data()));   // renderer.rs:26-49
        double mean = 0.0;
        for (size_t k = 0; k < sums.size(); k += 4) mean += (sums[k] + sums[k + 1] + sums[k + 2]) / 3.0;
        std::printf("rendered %lldx%lld at %d spp, mean radiance %.4f\n", (long long)cam.image_width, (long long)cam.image_height, spp,
                    mean / (sums.size() / 4) / spp);
        rt_scene_destroy(dev);
        rt_context_destroy(ctx);
    } catch (const Error& e) {
        std::fprintf(stderr, "error %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
