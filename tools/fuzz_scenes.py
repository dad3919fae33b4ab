"""Random scenes for device-vs-oracle parity sweeps (tools/fuzz_parity.py, tests/test_gpu_hits.py): random mixes of
spheres (static / moving), quads with arbitrary edge vectors (most stick out of the diagonal box Quad::new gives them,
quad.rs:41-43), cubes, Translate / RotateY instances (also nested in each other's subtrees), BVHs nested in BVHs, lists
and media."""
import numpy as np

import rust_tracing_b200 as rt


def random_scene(seed):
    rng = np.random.default_rng(seed)
    s = rt.Scene(bvh_seed=int(rng.integers(1, 1 << 30)))
    mats = [s.Lambertian(s.SolidColor(*rng.random(3))), s.Metal(rng.random(3), float(rng.random())), s.Dielectric(1.5),
            s.Lambertian(s.CheckerTexture(0.5, (0.1, 0.1, 0.1), (0.9, 0.9, 0.9))), s.DiffuseLight(s.SolidColor(4, 4, 4))]
    pick = lambda: mats[int(rng.integers(0, len(mats)))]

    def prim(depth=0):
        if depth < 2 and rng.random() < 0.15:     # an instance of a primitive: nests when the group around it is instanced too
            inner = prim(depth + 1)
            if rng.random() < 0.5:
                inner = s.RotateY(inner, float(rng.uniform(-90, 90)))
            return s.Translate(inner, rng.uniform(-2, 2, 3))
        k = int(rng.integers(0, 4))
        c = rng.uniform(-8, 8, 3)
        if k == 0:
            return s.Sphere(c, float(rng.uniform(0.3, 2.0)), pick())
        if k == 1:
            return s.Sphere(c, float(rng.uniform(0.3, 1.5)), pick(), target=c + rng.uniform(-1, 1, 3))
        if k == 2:
            u, v = rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 3)
            return s.Quad(c, u, v, pick())
        return s.cube(c, c + rng.uniform(0.5, 3.0, 3), pick())

    def group(n):
        l = rt.HittableList()
        for _ in range(n):
            l.add(prim())
        return s.BVHNode(l) if rng.random() < 0.7 else s.List(l)

    world = rt.HittableList()
    for _ in range(int(rng.integers(3, 8))):
        world.add(prim())
    for _ in range(int(rng.integers(1, 4))):
        g = group(int(rng.integers(2, 12)))
        r = rng.random()
        if r < 0.3:
            g = s.Translate(g, rng.uniform(-5, 5, 3))
        elif r < 0.6:
            g = s.Translate(s.RotateY(g, float(rng.uniform(-90, 90))), rng.uniform(-5, 5, 3))
        elif r < 0.7:
            g = s.RotateY(s.Translate(g, rng.uniform(-5, 5, 3)), float(rng.uniform(-90, 90)))
        world.add(g)
    for _ in range(int(rng.integers(0, 3))):
        c = rng.uniform(-6, 6, 3)
        r = rng.random()
        if r < 0.4:
            b = s.Sphere(c, float(rng.uniform(1, 3)), mats[2])
        elif r < 0.8:
            b = s.Translate(s.RotateY(s.cube((0, 0, 0), rng.uniform(1, 4, 3), mats[0]), float(rng.uniform(-60, 60))), c)
        else:
            b = group(3)            # a generic (program) boundary
        world.add(s.ConstantMedium(b, float(rng.uniform(0.1, 1.5)), rng.random(3)))
    nested = rt.HittableList()
    nested.add(group(6))
    nested.add(prim())
    world.add(s.BVHNode(nested))
    s.finish(s.BVHNode(world) if rng.random() < 0.8 else s.List(world))
    return s
