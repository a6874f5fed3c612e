"""u2pl_b200 -- B200-native (sm_100a) implementation of U2PL's per-step semi-supervised
training hot path, behind the reference's own call signatures.

    u2pl_b200.ops        torch-facing wrappers over the C ABI (include/u2pl_b200.h)
    u2pl_b200.u2pl       drop-in mirror of the reference package `u2pl` (models / utils)
    u2pl_b200.install()  puts that mirror on sys.path as `u2pl` so train_semi.py imports it
"""
import os
import sys

__version__ = "0.1.0"
_HERE = os.path.dirname(os.path.abspath(__file__))


def install():
    """Make `import u2pl` resolve to the drop-in mirror shipped in this package."""
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    for name in [m for m in sys.modules if m == "u2pl" or m.startswith("u2pl.")]:
        mod = sys.modules[name]
        if not getattr(mod, "__file__", "").startswith(_HERE):
            del sys.modules[name]
