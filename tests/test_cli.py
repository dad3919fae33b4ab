"""CLI surface of main.rs:40-54 (no GPU needed for parsing)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cli():
    spec = importlib.util.spec_from_file_location("rt_cli", os.path.join(ROOT, "rust-tracing_b200", "__main__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_defaults_match_reference():
    a = load_cli().parse_args([])
    assert (a.live, a.scene, a.output) == (False, 0, "output")     # main.rs:44,48,52


def test_flags():
    a = load_cli().parse_args(["-l", "-s", "8", "-o", "final", "--width", "400", "--spp", "64", "--depth", "40"])
    assert a.live and a.scene == 8 and a.output == "final" and (a.width, a.spp, a.depth) == (400, 64, 40)
    b = load_cli().parse_args(["--live", "--scene", "3", "--output", "x"])
    assert b.live and b.scene == 3 and b.output == "x"
