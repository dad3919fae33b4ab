"""Host-side mirror of the crate's surface over the C ABI (include/rt_b200.h).

Names and argument meaning follow the reference (sphere.rs, quad.rs, hittable.rs,
constant_medium.rs, bvh.rs, material.rs, texture.rs, camera.rs, renderer.rs) so scenes and
tests read like the crate's own code:

    s = Scene()
    ground = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
    world = HittableList()
    world.add(s.Sphere((0, -1000, 0), 1000.0, ground))
    bvh = s.BVHNode(world)
    cam = Camera(CameraSettings(image_width=400, look_from=(13, 2, 3), look_at=(0, 0, 0), vfov=20))
    sums = render(cam, s.finish(bvh))          # per-pixel SUM over spp, like renderer.rs:26-49

All compute goes through the CUDA library; nothing here falls back to the CPU.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A


def _d3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


class Handle(int):
    """An id into the scene under construction (plays the role of Arc<dyn Trait>)."""


class HittableList:
    """hittable.rs:50-59."""

    def __init__(self):
        self.objects = []

    def add(self, obj):
        self.objects.append(int(obj))


class Scene:
    """Owns an rt_builder; its methods are the crate's constructors."""

    def __init__(self, bvh_seed=2, _raw=None):
        self._lib = A.lib()
        if _raw is not None:
            self._b = _raw
        else:
            b = C.c_void_p()
            A.check(self._lib.rt_builder_create(int(bvh_seed), C.byref(b)))
            self._b = b
        self._keep = []  # host arrays the description points into
        self.desc = None

    def close(self):
        if self._b is not None:
            self._lib.rt_builder_destroy(self._b)
            self._b = None
            self.desc = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- texture.rs
    def SolidColor(self, r, g=None, b=None):
        if g is None:
            r, g, b = r
        return Handle(A.check(self._lib.rt_tex_solid(self._b, r, g, b)))

    def CheckerTexture(self, scale, even, odd):
        if not isinstance(even, Handle):
            even = self.SolidColor(even)   # new_from_colors, texture.rs:51
        if not isinstance(odd, Handle):
            odd = self.SolidColor(odd)
        return Handle(A.check(self._lib.rt_tex_checker(self._b, scale, even, odd)))

    def ImageTexture(self, rgb8):
        """rgb8: uint8 array (H, W, 3) — the decoded image (texture.rs:76-80 decodes a file)."""
        arr = np.ascontiguousarray(rgb8, dtype=np.uint8)
        if arr.ndim != 3 or arr.shape[2] != 3:
            raise ValueError("ImageTexture expects an (H, W, 3) uint8 array")
        h, w = arr.shape[:2]
        return Handle(A.check(self._lib.rt_tex_image(self._b, w, h, arr.ctypes.data)))

    def NoiseTexture(self, scale, perlin_seed=3):
        return Handle(A.check(self._lib.rt_tex_noise(self._b, scale, int(perlin_seed))))

    def NoiseTextureFromTables(self, scale, ranvec, perm_x, perm_y, perm_z):
        """rt_tex_noise_tables: NoiseTexture whose Perlin tables the host drew itself (perlin.rs:16-25)."""
        rv = np.ascontiguousarray(ranvec, dtype=np.float64).reshape(256, 3)
        px, py, pz = (np.ascontiguousarray(p, dtype=np.int32).reshape(256) for p in (perm_x, perm_y, perm_z))
        return Handle(A.check(self._lib.rt_tex_noise_tables(self._b, scale, rv.ctypes.data, px.ctypes.data, py.ctypes.data, pz.ctypes.data)))

    # ---- material.rs
    def Lambertian(self, albedo):
        return Handle(A.check(self._lib.rt_mat_lambertian(self._b, albedo)))

    def Metal(self, albedo, fuzz):
        return Handle(A.check(self._lib.rt_mat_metal(self._b, _d3(albedo), fuzz)))

    def Dielectric(self, ir):
        return Handle(A.check(self._lib.rt_mat_dielectric(self._b, ir)))

    def DiffuseLight(self, emit):
        return Handle(A.check(self._lib.rt_mat_diffuse_light(self._b, emit)))

    def Isotropic(self, albedo):
        return Handle(A.check(self._lib.rt_mat_isotropic(self._b, albedo)))

    # ---- sphere.rs / quad.rs / hittable.rs / constant_medium.rs / bvh.rs
    def Sphere(self, center, radius, material, target=None):
        if target is None:
            return Handle(A.check(self._lib.rt_hit_sphere(self._b, _d3(center), radius, material)))
        return Handle(A.check(self._lib.rt_hit_moving_sphere(self._b, _d3(center), _d3(target), radius, material)))

    def Quad(self, q, u, v, material):
        return Handle(A.check(self._lib.rt_hit_quad(self._b, _d3(q), _d3(u), _d3(v), material)))

    def cube(self, a, b, material):
        return Handle(A.check(self._lib.rt_hit_cube(self._b, _d3(a), _d3(b), material)))

    def List(self, hlist):
        ids = (C.c_int * max(1, len(hlist.objects)))(*hlist.objects)
        return Handle(A.check(self._lib.rt_hit_list(self._b, ids, len(hlist.objects))))

    def Translate(self, obj, offset):
        return Handle(A.check(self._lib.rt_hit_translate(self._b, obj, _d3(offset))))

    def RotateY(self, obj, angle):
        return Handle(A.check(self._lib.rt_hit_rotate_y(self._b, obj, angle)))

    def ConstantMedium(self, boundary, density, albedo):
        if not isinstance(albedo, Handle):
            albedo = self.SolidColor(albedo)   # new_from_color, constant_medium.rs:28
        return Handle(A.check(self._lib.rt_hit_constant_medium(self._b, boundary, density, albedo)))

    def BVHNode(self, hlist):
        ids = (C.c_int * max(1, len(hlist.objects)))(*hlist.objects)
        return Handle(A.check(self._lib.rt_hit_bvh(self._b, ids, len(hlist.objects))))

    def BVHNodeOnDevice(self, ctx, hlist):
        """rt_hit_bvh_device: BVHNode::new with the sorting on the GPU; the same node array as BVHNode(hlist)."""
        ids = (C.c_int * max(1, len(hlist.objects)))(*hlist.objects)
        return Handle(A.check(self._lib.rt_hit_bvh_device(self._b, ctx._h, ids, len(hlist.objects))))

    def BVHFromNodes(self, nodes):
        """rt_hit_bvh_nodes: a BVH the host built itself. `nodes`: pre-order sequence of (bbox[6], left, right, object)
        with left / right indexing into the sequence (-1 for leaves) - what a Rust host reads off its own BVHNode."""
        arr = (A.BvhNodeDesc * len(nodes))()
        for k, (bbox, left, right, obj) in enumerate(nodes):
            arr[k].bbox[:] = [float(x) for x in bbox]
            arr[k].left, arr[k].right, arr[k].object, arr[k].axis = int(left), int(right), int(obj), -1
        return Handle(A.check(self._lib.rt_hit_bvh_nodes(self._b, arr, len(nodes))))

    def finish(self, world):
        d = A.SceneDesc()
        A.check(self._lib.rt_builder_finish(self._b, world, C.byref(d)))
        self.desc = d
        return self


class CameraSettings(A.CameraSettingsC):
    """camera.rs:8-37 — keyword arguments override CameraSettings::default()."""

    def __init__(self, **kw):
        super().__init__()
        A.lib().rt_camera_settings_default(C.byref(self))
        for k, v in kw.items():
            if k in ("look_from", "look_at", "vup", "background"):
                setattr(self, k, _d3(v))
            else:
                setattr(self, k, v)


class Camera(A.CameraDesc):
    """camera.rs:38-110."""

    def __init__(self, settings):
        super().__init__()
        A.check(A.lib().rt_camera_new(C.byref(settings), C.byref(self)))

    @property
    def shape(self):
        return int(self.image_height), int(self.image_width)


SCENE_NAMES = ["random_balls", "two_spheres", "earth", "two_perlin_spheres", "quads", "simple_light",
               "cornell_box", "cornell_smoke", "final_scene"]   # main.rs:47


def synthetic_earth(width=6400, height=3200, seed=11):
    """Deterministic stand-in for assets/earth-large.jpg (same 6400x3200 RGB8 shape): blue 'ocean'
    with low-frequency green/brown 'continents' and white caps. Used when the JPEG is not available
    (the GPU box has no copy of the reference's assets)."""
    rng = np.random.default_rng(seed)
    gh, gw = 33, 65
    coarse = rng.random((gh, gw)).astype(np.float32)
    coarse[:, -1] = coarse[:, 0]
    ys = np.linspace(0, gh - 1, height, dtype=np.float32)
    xs = np.linspace(0, gw - 1, width, dtype=np.float32)
    y0 = np.minimum(ys.astype(np.int32), gh - 2)
    x0 = np.minimum(xs.astype(np.int32), gw - 2)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy)
    fx = fx * fx * (3 - 2 * fx)
    a = coarse[y0][:, x0]
    b = coarse[y0][:, x0 + 1]
    c = coarse[y0 + 1][:, x0]
    d = coarse[y0 + 1][:, x0 + 1]
    field = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
    fine = ((np.arange(height)[:, None] * 7 + np.arange(width)[None, :] * 13) % 32).astype(np.float32) / 32.0
    land = field > 0.55
    lat = np.abs(np.linspace(-1, 1, height, dtype=np.float32))[:, None]
    img = np.empty((height, width, 3), dtype=np.uint8)
    img[..., 0] = np.where(land, 60 + 90 * field + 20 * fine, 10 + 20 * fine)
    img[..., 1] = np.where(land, 90 + 80 * field + 20 * fine, 40 + 60 * field)
    img[..., 2] = np.where(land, 40 + 30 * fine, 110 + 100 * field)
    cap = (lat > 0.9)
    img[np.broadcast_to(cap, (height, width))] = 240
    return img


_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def jpeg_entropy_decode(data):
    """rt_jpeg_entropy_decode (host only): -> (JpegInfo, int16 coefficients)."""
    lib = A.lib()
    buf = np.frombuffer(data, dtype=np.uint8)
    info = A.JpegInfo()
    A.check(lib.rt_jpeg_entropy_decode(buf.ctypes.data, len(buf), C.byref(info), None, 0))
    coef = np.empty(info.coef_count, dtype=np.int16)
    A.check(lib.rt_jpeg_entropy_decode(buf.ctypes.data, len(buf), C.byref(info), coef.ctypes.data, coef.size))
    return info, coef


def load_earth(path=None, ctx=None):
    """The decoded earth image (texture.rs:76-80 decodes `assets/earth-large.jpg`, main.rs:179,591) as RGB8: decoded on
    the GPU by rt_jpeg_decode when a Context is given (same bytes), by PIL otherwise. Looked for, in order: `path`, $RT_B200_EARTH, ./assets/earth-large.jpg, <repo>/assets/earth-large.jpg (the
    byte copy tools/make_reference_fixtures.py / build() make where the reference checkout exists; it travels to the
    GPU box) and the reference checkout itself. Only if none exists: the synthetic stand-in, and the source string
    says so. Returns (array, source)."""
    candidates = [path, os.environ.get("RT_B200_EARTH"), os.path.join("assets", "earth-large.jpg"),
                  os.path.join(_REPO_ROOT, "assets", "earth-large.jpg"),
                  os.path.join(os.environ.get("RT_REFERENCE", "/root/reference"), "assets", "earth-large.jpg")]
    for p in candidates:
        if p and os.path.exists(p):
            if ctx is not None and p.lower().endswith((".jpg", ".jpeg")):
                with open(p, "rb") as f:
                    return ctx.jpeg_decode(f.read()), p
            from PIL import Image
            Image.MAX_IMAGE_PIXELS = None
            return np.asarray(Image.open(p).convert("RGB"), dtype=np.uint8), p
    return synthetic_earth(), "synthetic"


def builtin_scene(scene, image_width=0, samples_per_pixel=0, max_depth=0, scene_seed=1, bvh_seed=2,
                  perlin_seed=3, earth=None):
    """The CLI scenes of main.rs:56-639 by index or name. Returns (Scene, CameraSettings)."""
    if isinstance(scene, str):
        scene = SCENE_NAMES.index(scene)
    lib = A.lib()
    req = A.SceneRequest()
    req.scene = scene
    req.image_width = image_width
    req.samples_per_pixel = samples_per_pixel
    req.max_depth = max_depth
    req.scene_seed, req.bvh_seed, req.perlin_seed = scene_seed, bvh_seed, perlin_seed
    keep = None
    if scene in (2, 8):
        if earth is None:
            earth, _ = load_earth()
        keep = np.ascontiguousarray(earth, dtype=np.uint8)
        req.earth_height, req.earth_width = keep.shape[:2]
        req.earth_rgb8 = keep.ctypes.data_as(C.POINTER(C.c_uint8))
    raw = C.c_void_p()
    desc = A.SceneDesc()
    settings = CameraSettings()
    A.check(lib.rt_scene_builtin(C.byref(req), C.byref(raw), C.byref(desc), C.byref(settings)))
    s = Scene(_raw=raw)
    s.desc = desc
    return s, settings


def layout_flags(prune=True, box_primitives=True, hoist_media=True, ops_in_smem=True, generic_kernel=False):
    """RT_LAYOUT_* switches of the flattening (tests and A/B runs; the defaults are the product's layout)."""
    return ((0 if prune else A.RT_LAYOUT_NO_PRUNE) | (0 if box_primitives else A.RT_LAYOUT_NO_BOX_PRIMITIVES) |
            (0 if hoist_media else A.RT_LAYOUT_NO_HOIST) | (0 if ops_in_smem else A.RT_LAYOUT_OPS_IN_GLOBAL) |
            (A.RT_LAYOUT_GENERIC_KERNEL if generic_kernel else 0))


def scene_layout(scene, flags=0):
    """Host-side dry run of the flattening rt_scene_upload performs (no GPU): dict of op counts and sizes."""
    info = A.LayoutInfo()
    A.check(A.lib().rt_scene_layout(C.byref(scene.desc), flags, C.byref(info)))
    return {name: getattr(info, name) for name, _ in A.LayoutInfo._fields_}


def scene_ops(scene, flags=0):
    """rt_scene_ops_export: the flattened traversal stream (host dry run). Returns a dict with `words` (float32 (N, 4)),
    `n_world_words`, `media_ops` (word indices of the hoisted media) and `first_link` (link of op 0)."""
    lib = A.lib()
    n, nw, nm, fl = C.c_int64(), C.c_int32(), C.c_int32(), C.c_uint32()
    media = (C.c_int32 * 8)()
    A.check(lib.rt_scene_ops_export(C.byref(scene.desc), flags, None, 0, C.byref(n), None, None, None, None))
    words = np.zeros((n.value, 4), dtype=np.float32)
    A.check(lib.rt_scene_ops_export(C.byref(scene.desc), flags, words.ctypes.data_as(C.POINTER(C.c_float)), n.value, C.byref(n),
                                    C.byref(nw), media, C.byref(nm), C.byref(fl)))
    return {"words": words, "n_world_words": nw.value, "media_ops": [int(media[k]) for k in range(nm.value)],
            "first_link": fl.value}


class DeviceScene:
    def __init__(self, ctx, scene, flags=0):
        self.ctx = ctx
        self._lib = A.lib()
        h = C.c_void_p()
        A.check(self._lib.rt_scene_upload_ex(ctx._h, C.byref(scene.desc), flags, C.byref(h)))
        self._h = h

    def close(self):
        if self._h is not None:
            self._lib.rt_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One per GPU (rt_context)."""

    def __init__(self, device_id=0):
        self._lib = A.lib()
        h = C.c_void_p()
        A.check(self._lib.rt_context_create(device_id, C.byref(h)))
        self._h = h
        self.device_id = device_id

    def close(self):
        if self._h is not None:
            self._lib.rt_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm, khz, mem = C.c_int(), C.c_int(), C.c_size_t()
        A.check(self._lib.rt_device_info(self._h, C.byref(sm), C.byref(khz), C.byref(mem)))
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "total_mem": mem.value}

    def upload(self, scene, flags=0):
        return DeviceScene(self, scene, flags)

    def render(self, dscene, cam, sample_begin=0, sample_count=None, seed=0):
        """rt_render: host float32 (H, W, 4) sums (x,y,z radiance, w sample count)."""
        if sample_count is None:
            sample_count = cam.samples_per_pixel
        h, w = cam.shape
        out = np.empty((h, w, 4), dtype=np.float32)
        A.check(self._lib.rt_render(self._h, dscene._h, C.byref(cam), sample_begin, sample_count, seed,
                                    out.ctypes.data))
        return out

    def render_rgb8(self, dscene, cam, sample_begin=0, sample_count=None, seed=0):
        """rt_render_rgb8: render + device-side color_to_rgb(sum/spp); host uint8 (H, W, 3)."""
        if sample_count is None:
            sample_count = cam.samples_per_pixel
        h, w = cam.shape
        out = np.empty((h, w, 3), dtype=np.uint8)
        A.check(self._lib.rt_render_rgb8(self._h, dscene._h, C.byref(cam), sample_begin, sample_count, seed, out.ctypes.data))
        return out

    def render_accumulate(self, dscene, cam, sample_begin, sample_count, seed, d_sum_ptr, stream=0):
        """rt_render_accumulate into a device float4 buffer (e.g. a torch tensor's data_ptr())."""
        A.check(self._lib.rt_render_accumulate(self._h, dscene._h, C.byref(cam), sample_begin, sample_count,
                                               seed, C.c_void_p(d_sum_ptr), C.c_void_p(stream)))

    def finalize_rgb8(self, d_sum_ptr, n_pixels, spp, stream=0):
        """rt_finalize_rgb8 on `stream` (the stream the framebuffer was rendered on)."""
        out = np.empty((n_pixels, 3), dtype=np.uint8)
        A.check(self._lib.rt_finalize_rgb8(self._h, C.c_void_p(d_sum_ptr), n_pixels, float(spp), out.ctypes.data,
                                           C.c_void_p(stream)))
        return out

    def stats(self):
        st = A.RenderStats()
        A.check(self._lib.rt_render_get_stats(self._h, C.byref(st)))
        return {"paths": st.paths, "segments": st.segments, "kernel_launches": st.kernel_launches,
                "last_kernel_ms": st.last_kernel_ms}

    def count_ops(self, dscene, cam, sample_begin=0, sample_count=1, seed=0):
        """Op counts of the instrumented kernel over a sample range (dict name -> count)."""
        buf = np.zeros(64, dtype=np.uint64)
        names = C.c_char_p()
        n = A.check(self._lib.rt_render_count_ops(self._h, dscene._h, C.byref(cam), sample_begin, sample_count, seed,
                                                  buf.ctypes.data, len(buf), C.byref(names)))
        return dict(zip(names.value.decode().split(","), (int(x) for x in buf[:n])))

    def hit_batch(self, dscene, rays, t_min=0.001, t_max=float("inf"), seed=7):
        rays = np.ascontiguousarray(rays, dtype=A.ray_dtype())
        out = np.zeros(len(rays), dtype=A.hit_dtype())
        A.check(self._lib.rt_hit_batch(self._h, dscene._h, rays.ctypes.data, len(rays), t_min, t_max, seed,
                                       out.ctypes.data))
        return out

    def texture_batch(self, dscene, tex, uvp):
        uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
        out = np.empty((len(uvp), 3), dtype=np.float64)
        A.check(self._lib.rt_texture_batch(self._h, dscene._h, int(tex), uvp.ctypes.data, len(uvp), out.ctypes.data))
        return out

    def get_ray_batch(self, cam, pixel_index, sample_index, seed=0):
        pix = np.ascontiguousarray(pixel_index, dtype=np.int64)
        smp = np.ascontiguousarray(sample_index, dtype=np.int64)
        out = np.zeros(len(pix), dtype=A.ray_dtype())
        A.check(self._lib.rt_get_ray_batch(self._h, C.byref(cam), pix.ctypes.data, smp.ctypes.data, len(pix), seed,
                                           out.ctypes.data))
        return out

    def scatter_batch(self, dscene, rays, hits, pixel, sample, segment=0, seed=0):
        """rt_scatter_batch: Material::emitted / scatter on (ray, hit record) pairs; structured array (A.scatter_dtype())."""
        rays = np.ascontiguousarray(rays, dtype=A.ray_dtype())
        hits = np.ascontiguousarray(hits, dtype=A.hit_dtype())
        pixel = np.ascontiguousarray(pixel, dtype=np.uint32)
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        assert len(rays) == len(hits) == len(pixel) == len(sample)
        out = np.zeros(len(rays), dtype=A.scatter_dtype())
        A.check(self._lib.rt_scatter_batch(self._h, dscene._h, rays.ctypes.data, hits.ctypes.data, len(rays), seed,
                                           pixel.ctypes.data, sample.ctypes.data, int(segment), out.ctypes.data))
        return out

    def jpeg_decode(self, data, backend="own"):
        """rt_jpeg_decode (host Huffman + this library's kernels, bytes equal to libjpeg-turbo's) or, backend="nvjpeg",
        rt_jpeg_decode_nvjpeg: JPEG bytes -> uint8 (H, W, 3)."""
        fn = self._lib.rt_jpeg_decode if backend == "own" else self._lib.rt_jpeg_decode_nvjpeg
        buf = np.frombuffer(data, dtype=np.uint8)
        w, h = C.c_int(), C.c_int()
        A.check(fn(self._h, buf.ctypes.data, len(buf), C.byref(w), C.byref(h), None, 0))
        out = np.empty((h.value, w.value, 3), dtype=np.uint8)
        A.check(fn(self._h, buf.ctypes.data, len(buf), C.byref(w), C.byref(h), out.ctypes.data, out.nbytes))
        return out

    def bvh_export(self, dscene, bvh_hittable, capacity=1 << 20):
        buf = np.empty(capacity, dtype=np.int32)
        n = C.c_int32()
        A.check(self._lib.rt_bvh_export(dscene._h, int(bvh_hittable), buf.ctypes.data, capacity, C.byref(n)))
        return buf[: n.value].copy()

    def measure_fp32_peak(self):
        v = C.c_double()
        A.check(self._lib.rt_measure_fp32_peak(self._h, C.byref(v)))
        return v.value


def render_multi(contexts, dscenes, cam, sample_begin=0, sample_count=None, seed=0, weights=None):
    """rt_render_multi: one Context + DeviceScene per GPU of this process; returns (host float32 (H, W, 4) sums, shares)."""
    n = len(contexts)
    assert n == len(dscenes) and n >= 1
    if sample_count is None:
        sample_count = cam.samples_per_pixel
    h, w = cam.shape
    out = np.empty((h, w, 4), dtype=np.float32)
    cs = (C.c_void_p * n)(*[c._h for c in contexts])
    ss = (C.c_void_p * n)(*[d._h for d in dscenes])
    ws = (C.c_double * n)(*[float(x) for x in weights]) if weights is not None else None
    shares = (C.c_int64 * n)()
    A.check(A.lib().rt_render_multi(cs, ss, n, C.byref(cam), sample_begin, sample_count, seed, ws, out.ctypes.data, shares))
    return out, [int(x) for x in shares]


_default_ctx = {}


def default_context(device_id=0):
    if device_id not in _default_ctx:
        _default_ctx[device_id] = Context(device_id)
    return _default_ctx[device_id]


def color_to_rgb8(sums, spp):
    """Host reference of the post-process for small arrays (color.rs:12-19 on sum/spp); the device
    path is Context.finalize_rgb8."""
    c = np.asarray(sums, dtype=np.float64)[..., :3] * (1.0 / spp)
    with np.errstate(invalid="ignore"):
        g = np.power(c, 1.0 / 2.2)
    g = np.clip(g, 0.0, 0.999)
    g = np.where(np.isnan(g), 0.0, g)
    return (256.0 * g).astype(np.uint8)


def render(camera, world, output_file_name=None, seed=0, device_id=0):
    """Drop-in for `pub fn render(camera, world, output_file_name)` (renderer.rs:12): renders
    camera.samples_per_pixel samples per pixel on the GPU. With output_file_name it writes `<name>.png` like
    renderer.rs:53-74 (the bytes come from the device, rt_render_rgb8) and returns the (H, W, 3) uint8 image;
    without, it returns the (H, W, 3) float32 SUM image of renderer.rs:26-49."""
    ctx = default_context(device_id)
    ds = ctx.upload(world)
    try:
        if output_file_name:
            from PIL import Image
            rgb = ctx.render_rgb8(ds, camera, 0, camera.samples_per_pixel, seed)
            Image.fromarray(rgb).save(f"{output_file_name}.png")
            return rgb
        return ctx.render(ds, camera, 0, camera.samples_per_pixel, seed)[..., :3]
    finally:
        ds.close()
