"""bindings/rust/b200.rs cannot be compiled in this image (no cargo / rustc), so it is checked the way a compiler's FFI lint
would: every #[repr(C)] struct and every extern "C" prototype against include/rt_b200.h (tools/check_rust_abi.py), the
struct sizes against the ctypes view that tests/test_abi.py ties to the built library, and the coverage of the crate's
closed sets (16 types need a `flatten`, Camera a `to_desc`)."""
import ctypes as C
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RS = os.path.join(ROOT, "bindings", "rust", "b200.rs")


def checker():
    spec = importlib.util.spec_from_file_location("check_rust_abi", os.path.join(ROOT, "tools", "check_rust_abi.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_rust_declarations_match_the_header(rt, capsys):
    rc, sizes, structs, funcs = checker().main()
    assert rc == 0, capsys.readouterr().out
    A = rt._abi
    want = {"rt_texture_desc": A.TextureDesc, "rt_material_desc": A.MaterialDesc, "rt_hittable_desc": A.HittableDesc,
            "rt_bvh_node_desc": A.BvhNodeDesc, "rt_perlin_desc": A.PerlinDesc, "rt_scene_desc": A.SceneDesc,
            "rt_camera_desc": A.CameraDesc, "rt_render_stats": A.RenderStats}
    for name, ct in want.items():
        assert sizes[name] == C.sizeof(ct), name
    assert len(structs["rt_scene_desc"]) == 17            # the struct round 1 shipped empty
    for must in ("rt_builder_finish", "rt_scene_upload", "rt_render", "rt_render_multi", "rt_render_accumulate", "rt_finalize_rgb8",
                 "rt_hit_bvh_nodes", "rt_tex_noise_tables", "rt_hit_rotate_y_sincos", "rt_hit_constant_medium_nid", "rt_tex_checker_inv"):
        assert must in funcs


def test_every_closed_set_member_flattens():
    src = open(RS).read()
    for ty in ("SolidColor", "CheckerTexture", "ImageTexture", "NoiseTexture", "Lambertian", "Metal", "Dielectric", "DiffuseLight",
               "Isotropic", "Sphere", "Quad", "HittableList", "Translate", "RotateY", "ConstantMedium", "BVHNode"):
        m = re.search(r"impl %s \{.*?\n\}" % ty, src, flags=re.S)
        assert m and "fn flatten(&self, b: &mut SceneBuilder) -> c_int" in m.group(0), ty
    assert re.search(r"impl Camera \{.*?pub fn to_desc\(&self\) -> rt_camera_desc", src, flags=re.S)
    assert "placeholder" not in src and "same pattern" not in src


def test_the_checker_catches_mistakes(tmp_path, monkeypatch):
    """A swapped field, a dropped parameter and a wrong pointer type must each be reported."""
    m = checker()
    src = open(RS).read()
    broken = [src.replace("    pub n_textures: i32,\n    pub n_materials: i32,", "    pub n_materials: i32,\n    pub n_textures: i32,"),
              src.replace("seed: u64, host_sum_rgba: *mut c_float) -> c_int;", "host_sum_rgba: *mut c_float) -> c_int;", 1),
              src.replace("pub rgb8: *const u8,", "pub rgb8: u64,"),
              src.replace("pub fn rt_mat_dielectric(b: *mut rt_builder, ir: c_double)", "pub fn rt_mat_dielectric(b: *mut rt_builder, ir: c_float)")]
    for k, text in enumerate(broken):
        assert text != src, k
        fake_root = tmp_path / f"r{k}"
        (fake_root / "bindings" / "rust").mkdir(parents=True)
        (fake_root / "include").mkdir()
        (fake_root / "bindings" / "rust" / "b200.rs").write_text(text)
        (fake_root / "include" / "rt_b200.h").write_text(open(os.path.join(ROOT, "include", "rt_b200.h")).read())
        monkeypatch.setattr(m, "ROOT", str(fake_root))
        assert m.main()[0] == 1, k
