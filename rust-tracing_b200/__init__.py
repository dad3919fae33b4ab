"""B200-native path-tracing hot path of Husenap/rust-tracing (renderer.rs:26-49,139-155).

The product is csrc/librt_b200.so (host scene code + hand-written sm_100a kernels) behind the C ABI
of include/rt_b200.h; this package is the thin host-side mirror of the crate's surface.
Import name: `rust_tracing_b200` (the directory is `rust-tracing_b200/`; see ../rust_tracing_b200/).
"""
from . import _abi
from .api import (Camera, CameraSettings, Context, DeviceScene, Handle, HittableList, Scene, SCENE_NAMES,
                  builtin_scene, color_to_rgb8, default_context, jpeg_entropy_decode, layout_flags, load_earth, render, render_multi, scene_layout, scene_ops,
                  synthetic_earth)

__all__ = ["Camera", "CameraSettings", "Context", "DeviceScene", "Handle", "HittableList", "Scene", "SCENE_NAMES",
           "builtin_scene", "color_to_rgb8", "default_context", "jpeg_entropy_decode", "layout_flags", "load_earth", "render", "render_multi", "scene_layout", "scene_ops", "synthetic_earth", "_abi"]
