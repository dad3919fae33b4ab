"""Import shim: `rust-tracing_b200/` (the package directory the layout asks for) is not a valid
Python identifier, so `import rust_tracing_b200` resolves here and re-exports that package."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rust-tracing_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
