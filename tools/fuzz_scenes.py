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


def rich_scene(seed):
    """A wider mix than random_scene (which tests pin by seed, so it stays as it is): noise / image / nested checker textures
    (sometimes more NoiseTextures than fit in shared memory), emissive textures, fuzz up to 1, a 1000-unit ground sphere and a
    camera-enclosing fog sphere (f64 sphere code), media inside instances and inside instanced groups (they stay in the op
    stream), media with textured albedo or a moving boundary, and now and then a BVH big enough that the op stream no longer
    fits in shared memory (the kernel then reads it from global memory). Some combinations are ones the device layout
    rejects (RT_ERR_UNSUPPORTED at upload): callers skip those."""
    rng = np.random.default_rng(seed)
    s = rt.Scene(bvh_seed=int(rng.integers(1, 1 << 30)))
    n_noise = int(rng.choice([0, 1, 2, 6]))
    noises = [s.NoiseTexture(float(rng.uniform(0.2, 4.0)), perlin_seed=int(rng.integers(1, 1000))) for _ in range(n_noise)]
    img = (rng.random((32, 64, 3)) * 255).astype(np.uint8)
    image = s.ImageTexture(img)
    solid = lambda: s.SolidColor(*rng.random(3))

    def texture(depth=0):
        r = rng.random()
        if r < 0.35 or (depth >= 2):
            return solid()
        if r < 0.55 and noises:
            return noises[int(rng.integers(0, len(noises)))]
        if r < 0.7:
            return image
        return s.CheckerTexture(float(rng.uniform(0.2, 2.0)), texture(depth + 1), texture(depth + 1))

    mats = [s.Lambertian(texture()) for _ in range(4)]
    mats += [s.Metal(rng.random(3), float(rng.choice([0.0, 0.3, 1.0]))), s.Dielectric(float(rng.choice([1.5, 1.0 / 1.5, 2.4]))),
             s.DiffuseLight(texture()), s.DiffuseLight(s.SolidColor(5, 4, 3))]
    pick = lambda: mats[int(rng.integers(0, len(mats)))]

    def medium(boundary):
        albedo = texture() if rng.random() < 0.4 else rng.random(3)
        return s.ConstantMedium(boundary, float(rng.uniform(0.05, 1.5)), albedo)

    def prim(depth=0):
        r = rng.random()
        if depth < 2 and r < 0.15:
            inner = prim(depth + 1)
            if rng.random() < 0.5:
                inner = s.RotateY(inner, float(rng.uniform(-180, 180)))
            return s.Translate(inner, rng.uniform(-2, 2, 3))
        if r < 0.22:      # a medium as a primitive of the group: inside an instance it cannot be hoisted
            c = rng.uniform(-8, 8, 3)
            k = rng.random()
            if k < 0.4:
                b = s.Sphere(c, float(rng.uniform(0.8, 2.5)), mats[5])
            elif k < 0.6:
                b = s.Sphere(c, float(rng.uniform(0.8, 2.0)), mats[5], target=c + rng.uniform(-1, 1, 3))
            else:
                b = s.Translate(s.RotateY(s.cube((0, 0, 0), rng.uniform(1, 4, 3), mats[0]), float(rng.uniform(-60, 60))), c)
            return medium(b)
        k = int(rng.integers(0, 4))
        c = rng.uniform(-8, 8, 3)
        if k == 0:
            return s.Sphere(c, float(rng.uniform(0.3, 2.0)), pick())
        if k == 1:
            return s.Sphere(c, float(rng.uniform(0.3, 1.5)), pick(), target=c + rng.uniform(-1, 1, 3))
        if k == 2:
            return s.Quad(c, rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 3), pick())
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
            g = s.Translate(s.RotateY(g, float(rng.uniform(-180, 180))), rng.uniform(-5, 5, 3))
        elif r < 0.7:
            g = s.RotateY(s.Translate(g, rng.uniform(-5, 5, 3)), float(rng.uniform(-180, 180)))
        world.add(g)
    if rng.random() < 0.5:
        world.add(s.Sphere((0, -1012, 0), 1000.0, pick()))                       # ground: radius > 200 -> f64 sphere code
    if rng.random() < 0.3:
        world.add(s.ConstantMedium(s.Sphere((0, 0, 0), 300.0, mats[5]), float(rng.uniform(0.002, 0.02)), (1, 1, 1)))   # fog around the camera
    if rng.random() < 0.15:                                                    # a stream too large for shared memory
        big = rt.HittableList()
        for _ in range(2600):
            c = rng.uniform(-9, 9, 3)
            big.add(s.Sphere(c, float(rng.uniform(0.05, 0.25)), pick()))
        world.add(s.Translate(s.BVHNode(big), (0, 14, 0)) if rng.random() < 0.5 else s.BVHNode(big))
    for _ in range(int(rng.integers(0, 3))):
        c = rng.uniform(-6, 6, 3)
        r = rng.random()
        if r < 0.4:
            b = s.Sphere(c, float(rng.uniform(1, 3)), mats[5])
        elif r < 0.8:
            b = s.Translate(s.RotateY(s.cube((0, 0, 0), rng.uniform(1, 4, 3), mats[0]), float(rng.uniform(-60, 60))), c)
        else:
            b = group(3)
        world.add(medium(b))
    s.finish(s.BVHNode(world) if rng.random() < 0.8 else s.List(world))
    return s
