"""Test infrastructure (not product): import the UNMODIFIED reference
(scikit-gstat 1.0.23) from /root/reference in the build container.

The reference imports matplotlib and imageio at module import time
(skgstat/Variogram.py:17 -> plotting/variogram_plot.py:2, data/_loader.py:8-11);
neither is installed here, so empty stub modules are seeded first.  Nothing
the hot path computes touches them.  /root/reference does not exist on the GPU
box; there the copy installed under oracle/_ref (see INSTALLED_ROOT) is used - by
bench.py's reference arm / cpu_baseline leg only.  oracle/make_golden.py (fixture
generation) and the tests that skip when the reference is absent use it too.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
# `pip install --no-deps --target oracle/_ref /root/reference` (done by __graft_entry__.build() in
# the build container; git-ignored, but it travels to the GPU box with the repo snapshot)
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """Where the unmodified reference can be imported from: the source tree if present (build
    container), else the copy installed under oracle/_ref (GPU box), else None."""
    for root in (REFERENCE_ROOT, INSTALLED_ROOT):
        if os.path.isdir(os.path.join(root, "skgstat")):
            return root
    return None


def available() -> bool:
    return reference_root() is not None


def import_reference():
    if "skgstat" in sys.modules:
        return sys.modules["skgstat"]
    root = reference_root()
    if root is None:
        raise ImportError("reference not present at %s or %s" % (REFERENCE_ROOT, INSTALLED_ROOT))
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "imageio", "imageio.v3"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.collections"].LineCollection = object
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object

    def _imread(path):
        from PIL import Image
        return np.asarray(Image.open(path))

    sys.modules["imageio"].imread = _imread
    sys.modules["imageio.v3"].imread = _imread
    sys.modules["imageio"].v3 = sys.modules["imageio.v3"]
    if root not in sys.path:
        sys.path.insert(0, root)
    import skgstat  # noqa: E402
    return skgstat
