import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rt():
    """The product package. Builds the CUDA library in-tree if it is missing (nvcc cross-compiles)."""
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "rust-tracing_b200", "csrc", "librt_b200.so")):
        g.build()
    import rust_tracing_b200
    return rust_tracing_b200


@pytest.fixture(scope="session")
def ob():
    """The CPU oracle (test infrastructure only)."""
    from oracle import binding
    binding.build()
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def ctx(rt):
    """Device context; fails loudly (no CPU fallback) if there is no GPU."""
    return rt.Context(0)


@pytest.fixture(scope="session")
def earth(rt):
    arr, _ = rt.load_earth()
    return arr


def small_scene(rt, idx, earth=None, width=None, **kw):
    w = width or {0: 200, 1: 200, 2: 200, 3: 200, 4: 160, 5: 200, 6: 150, 7: 150, 8: 160}[idx]
    s, cs = rt.builtin_scene(idx, image_width=w, earth=earth if idx in (2, 8) else None, **kw)
    return s, rt.Camera(cs)
