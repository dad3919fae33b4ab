"""ctypes view of include/rt_b200.h — structures, loader and prototypes.

The shared library (csrc/librt_b200.so: host scene code + sm_100a kernels) is the product;
this module only describes its C ABI to Python. There is no CPU fallback: if the library is
missing, import fails loudly (run `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(HERE, "csrc", "librt_b200.so")   # RT_B200_LIB: development variants only

RT_OK = 0
RT_ERR_INVALID_ARGUMENT = -1
RT_ERR_OUT_OF_RANGE = -2
RT_ERR_UNSUPPORTED = -3
RT_ERR_CUDA = -4
RT_ERR_NO_DEVICE = -5
RT_ERR_OUT_OF_MEMORY = -6
RT_ERR_INTERNAL = -7

RT_TEX_SOLID, RT_TEX_CHECKER, RT_TEX_IMAGE, RT_TEX_NOISE = range(4)
RT_MAT_LAMBERTIAN, RT_MAT_METAL, RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE_LIGHT, RT_MAT_ISOTROPIC = range(5)
(RT_HIT_SPHERE, RT_HIT_QUAD, RT_HIT_LIST, RT_HIT_TRANSLATE, RT_HIT_ROTATE_Y,
 RT_HIT_CONSTANT_MEDIUM, RT_HIT_BVH) = range(7)
RT_FLAG_MOVING = 1
RT_FLAG_CUBE_LIST = 2
RT_LAYOUT_NO_PRUNE, RT_LAYOUT_NO_BOX_PRIMITIVES, RT_LAYOUT_NO_HOIST, RT_LAYOUT_OPS_IN_GLOBAL, RT_LAYOUT_GENERIC_KERNEL = 1, 2, 4, 8, 16

d3 = C.c_double * 3
d6 = C.c_double * 6


class TextureDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("_pad", C.c_int32),
                ("color", d3), ("scale", C.c_double)]


class MaterialDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tex", C.c_int32), ("albedo", d3), ("param", C.c_double)]


class HittableDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mat", C.c_int32), ("child", C.c_int32), ("count", C.c_int32),
                ("flags", C.c_uint32), ("_pad", C.c_int32),
                ("v0", d3), ("v1", d3), ("v2", d3), ("v3", d3), ("n", d3),
                ("s0", C.c_double), ("s1", C.c_double), ("bbox", d6)]


class BvhNodeDesc(C.Structure):
    _fields_ = [("bbox", d6), ("left", C.c_int32), ("right", C.c_int32), ("object", C.c_int32),
                ("axis", C.c_int32)]


class PerlinDesc(C.Structure):
    _fields_ = [("ranvec", (C.c_double * 3) * 256), ("perm_x", C.c_int32 * 256),
                ("perm_y", C.c_int32 * 256), ("perm_z", C.c_int32 * 256)]


class JpegInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("components", C.c_int32),
                ("h_samp", C.c_int32 * 3), ("v_samp", C.c_int32 * 3), ("blocks_w", C.c_int32 * 3), ("blocks_h", C.c_int32 * 3),
                ("adobe_rgb", C.c_int32), ("quant", (C.c_uint16 * 64) * 3), ("coef_offset", C.c_int64 * 3),
                ("coef_count", C.c_int64)]


class ImageDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb8", C.POINTER(C.c_uint8))]


class SceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("world", C.c_int32),
                ("n_textures", C.c_int32), ("n_materials", C.c_int32), ("n_hittables", C.c_int32),
                ("n_list_items", C.c_int32), ("n_bvh_nodes", C.c_int32), ("n_perlins", C.c_int32),
                ("n_images", C.c_int32), ("_pad", C.c_int32),
                ("textures", C.POINTER(TextureDesc)), ("materials", C.POINTER(MaterialDesc)),
                ("hittables", C.POINTER(HittableDesc)), ("list_items", C.POINTER(C.c_int32)),
                ("bvh_nodes", C.POINTER(BvhNodeDesc)), ("perlins", C.POINTER(PerlinDesc)),
                ("images", C.POINTER(ImageDesc))]


class CameraSettingsC(C.Structure):
    _fields_ = [("aspect_ratio", C.c_double), ("image_width", C.c_int64),
                ("samples_per_pixel", C.c_int32), ("max_depth", C.c_int32), ("vfov", C.c_double),
                ("look_from", d3), ("look_at", d3), ("vup", d3), ("defocus_angle", C.c_double),
                ("focus_dist", C.c_double), ("background", d3)]


class CameraDesc(C.Structure):
    _fields_ = [("image_width", C.c_int64), ("image_height", C.c_int64),
                ("samples_per_pixel", C.c_int32), ("max_depth", C.c_int32),
                ("background", d3), ("center", d3), ("pixel00_loc", d3), ("pixel_delta_u", d3),
                ("pixel_delta_v", d3), ("defocus_angle", C.c_double), ("defocus_disk_u", d3),
                ("defocus_disk_v", d3)]


class RayDesc(C.Structure):
    _fields_ = [("origin", d3), ("direction", d3), ("time", C.c_double)]


class HitDesc(C.Structure):
    _fields_ = [("t", C.c_double), ("p", d3), ("normal", d3), ("u", C.c_double), ("v", C.c_double),
                ("hit", C.c_int32), ("front_face", C.c_int32), ("prim_id", C.c_int32),
                ("mat_id", C.c_int32)]


class SceneRequest(C.Structure):
    _fields_ = [("scene", C.c_int32), ("image_width", C.c_int32), ("samples_per_pixel", C.c_int32),
                ("max_depth", C.c_int32), ("scene_seed", C.c_uint64), ("bvh_seed", C.c_uint64),
                ("perlin_seed", C.c_uint64), ("earth_width", C.c_int32), ("earth_height", C.c_int32),
                ("earth_rgb8", C.POINTER(C.c_uint8))]


class LayoutInfo(C.Structure):
    _fields_ = [("n_words", C.c_int32), ("n_inner", C.c_int32), ("n_sphere", C.c_int32), ("n_quad", C.c_int32),
                ("n_box", C.c_int32), ("n_xform", C.c_int32), ("n_medium_in_stream", C.c_int32),
                ("n_medium_hoisted", C.c_int32), ("n_precise_spheres", C.c_int32), ("n_bvh", C.c_int32),
                ("device_bytes", C.c_int64)]


class RenderStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("last_kernel_ms", C.c_float)]


# numpy dtypes with the same layout (for bulk ray / hit buffers)
def ray_dtype():
    import numpy as np
    return np.dtype([("origin", "f8", 3), ("direction", "f8", 3), ("time", "f8")])


def hit_dtype():
    import numpy as np
    return np.dtype([("t", "f8"), ("p", "f8", 3), ("normal", "f8", 3), ("u", "f8"), ("v", "f8"),
                     ("hit", "i4"), ("front_face", "i4"), ("prim_id", "i4"), ("mat_id", "i4")])


def scatter_dtype():
    import numpy as np
    return np.dtype([("ray_out", ray_dtype()), ("attenuation", "f8", 3), ("emitted", "f8", 3), ("scattered", "i4"), ("_pad", "i4")])


P = C.POINTER
vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/rt_b200.h declares
PROTOTYPES = {
    "rt_last_error": (C.c_char_p, []),
    "rt_abi_version": (C.c_int, []),
    "rt_builder_create": (C.c_int, [C.c_uint64, P(vp)]),
    "rt_builder_destroy": (None, [vp]),
    "rt_tex_solid": (C.c_int, [vp, C.c_double, C.c_double, C.c_double]),
    "rt_tex_checker": (C.c_int, [vp, C.c_double, C.c_int, C.c_int]),
    "rt_tex_image": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "rt_tex_noise": (C.c_int, [vp, C.c_double, C.c_uint64]),
    "rt_tex_noise_tables": (C.c_int, [vp, C.c_double, vp, vp, vp, vp]),
    "rt_mat_lambertian": (C.c_int, [vp, C.c_int]),
    "rt_mat_metal": (C.c_int, [vp, P(C.c_double), C.c_double]),
    "rt_mat_dielectric": (C.c_int, [vp, C.c_double]),
    "rt_mat_diffuse_light": (C.c_int, [vp, C.c_int]),
    "rt_mat_isotropic": (C.c_int, [vp, C.c_int]),
    "rt_hit_sphere": (C.c_int, [vp, P(C.c_double), C.c_double, C.c_int]),
    "rt_hit_moving_sphere": (C.c_int, [vp, P(C.c_double), P(C.c_double), C.c_double, C.c_int]),
    "rt_hit_quad": (C.c_int, [vp, P(C.c_double), P(C.c_double), P(C.c_double), C.c_int]),
    "rt_hit_cube": (C.c_int, [vp, P(C.c_double), P(C.c_double), C.c_int]),
    "rt_hit_list": (C.c_int, [vp, P(C.c_int), C.c_int]),
    "rt_hit_translate": (C.c_int, [vp, C.c_int, P(C.c_double)]),
    "rt_hit_rotate_y": (C.c_int, [vp, C.c_int, C.c_double]),
    "rt_hit_constant_medium": (C.c_int, [vp, C.c_int, C.c_double, C.c_int]),
    "rt_hit_bvh": (C.c_int, [vp, P(C.c_int), C.c_int]),
    "rt_hit_bvh_nodes": (C.c_int, [vp, vp, C.c_int]),
    "rt_tex_checker_inv": (C.c_int, [vp, C.c_double, C.c_int, C.c_int]),
    "rt_hit_rotate_y_sincos": (C.c_int, [vp, C.c_int, C.c_double, C.c_double]),
    "rt_hit_constant_medium_nid": (C.c_int, [vp, C.c_int, C.c_double, C.c_int]),
    "rt_builder_finish": (C.c_int, [vp, C.c_int, P(SceneDesc)]),
    "rt_camera_new": (C.c_int, [P(CameraSettingsC), P(CameraDesc)]),
    "rt_camera_settings_default": (None, [P(CameraSettingsC)]),
    "rt_scene_builtin": (C.c_int, [P(SceneRequest), P(vp), P(SceneDesc), P(CameraSettingsC)]),
    "rt_context_create": (C.c_int, [C.c_int, P(vp)]),
    "rt_context_destroy": (None, [vp]),
    "rt_device_info": (C.c_int, [vp, P(C.c_int), P(C.c_int), P(C.c_size_t)]),
    "rt_scene_upload": (C.c_int, [vp, P(SceneDesc), P(vp)]),
    "rt_scene_destroy": (None, [vp]),
    "rt_hit_bvh_device": (C.c_int, [vp, vp, P(C.c_int), C.c_int]),
    "rt_bvh_axis_draws": (C.c_int, [C.c_int]),
    "rt_bvh_build_device": (C.c_int, [vp, vp, C.c_int, vp, vp, vp]),
    "rt_jpeg_decode": (C.c_int, [vp, vp, C.c_size_t, P(C.c_int), P(C.c_int), vp, C.c_size_t]),
    "rt_jpeg_decode_nvjpeg": (C.c_int, [vp, vp, C.c_size_t, P(C.c_int), P(C.c_int), vp, C.c_size_t]),
    "rt_jpeg_entropy_decode": (C.c_int, [vp, C.c_size_t, P(JpegInfo), vp, C.c_size_t]),
    "rt_scene_upload_ex": (C.c_int, [vp, P(SceneDesc), C.c_uint32, P(vp)]),
    "rt_scene_layout": (C.c_int, [P(SceneDesc), C.c_uint32, P(LayoutInfo)]),
    "rt_scene_ops_export": (C.c_int, [P(SceneDesc), C.c_uint32, P(C.c_float), C.c_int64, P(C.c_int64), P(C.c_int32), P(C.c_int32),
                                      P(C.c_int32), P(C.c_uint32)]),
    "rt_render_accumulate": (C.c_int, [vp, vp, P(CameraDesc), C.c_int64, C.c_int64, C.c_uint64, vp, vp]),
    "rt_render": (C.c_int, [vp, vp, P(CameraDesc), C.c_int64, C.c_int64, C.c_uint64, vp]),
    "rt_render_multi": (C.c_int, [P(vp), P(vp), C.c_int, P(CameraDesc), C.c_int64, C.c_int64, C.c_uint64, P(C.c_double), vp, P(C.c_int64)]),
    "rt_render_rgb8": (C.c_int, [vp, vp, P(CameraDesc), C.c_int64, C.c_int64, C.c_uint64, vp]),
    "rt_finalize_rgb8": (C.c_int, [vp, vp, C.c_int64, C.c_double, vp, vp]),
    "rt_render_get_stats": (C.c_int, [vp, P(RenderStats)]),
    "rt_render_count_ops": (C.c_int, [vp, vp, P(CameraDesc), C.c_int64, C.c_int64, C.c_uint64, vp, C.c_int, P(C.c_char_p)]),
    "rt_hit_batch": (C.c_int, [vp, vp, vp, C.c_int64, C.c_double, C.c_double, C.c_uint64, vp]),
    "rt_texture_batch": (C.c_int, [vp, vp, C.c_int, vp, C.c_int64, vp]),
    "rt_get_ray_batch": (C.c_int, [vp, P(CameraDesc), vp, vp, C.c_int64, C.c_uint64, vp]),
    "rt_scatter_batch": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_uint64, vp, vp, C.c_uint32, vp]),
    "rt_bvh_export": (C.c_int, [vp, C.c_int, vp, C.c_int32, P(C.c_int32)]),
    "rt_measure_fp32_peak": (C.c_int, [vp, P(C.c_double)]),
}

_lib = None


class RtError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"rt_b200 error {status}: {message}")
        self.status = status


def lib():
    """Load csrc/librt_b200.so (once) and attach prototypes. Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc < 0:
        raise RtError(rc, lib().rt_last_error().decode("utf-8", "replace"))
    return rc
