"""u2pl_b200 -- B200-native (sm_100a) implementation of U2PL's per-step semi-supervised
training hot path, behind the reference's own call signatures.

    u2pl_b200.ops        torch-facing wrappers over the C ABI (include/u2pl_b200.h)
    u2pl_b200.u2pl       drop-in mirror of the reference package `u2pl` (models / utils)
    u2pl_b200.install()  registers that mirror as the top-level package `u2pl` so train_semi.py imports it
"""
import os
import sys

__version__ = "0.1.0"
_HERE = os.path.dirname(os.path.abspath(__file__))


class _MirrorFinder:
    """Meta-path finder that resolves ONLY the top-level name `u2pl` to the mirror package u2pl_b200/u2pl.  (Putting the
    package directory itself on sys.path would also expose build.py, ops.py, step.py ... as top-level modules and shadow
    a user's own `build` / `ops`.)"""

    @staticmethod
    def find_spec(name, path=None, target=None):
        if name != "u2pl":
            return None
        import importlib.util
        root = os.path.join(_HERE, "u2pl")
        return importlib.util.spec_from_file_location("u2pl", os.path.join(root, "__init__.py"), submodule_search_locations=[root])


def install():
    """Make `import u2pl` resolve to the drop-in mirror shipped in this package."""
    if not any(isinstance(f, type) and f.__name__ == "_MirrorFinder" for f in sys.meta_path):
        sys.meta_path.insert(0, _MirrorFinder)
    for name in [m for m in sys.modules if m == "u2pl" or m.startswith("u2pl.")]:
        mod = sys.modules[name]
        if not (getattr(mod, "__file__", None) or "").startswith(_HERE):
            del sys.modules[name]
