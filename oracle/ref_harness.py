"""Test infrastructure (not product): import the UNMODIFIED reference
(scikit-gstat 1.0.23) from /root/reference in the build container.

The reference imports matplotlib and imageio at module import time
(skgstat/Variogram.py:17 -> plotting/variogram_plot.py:2, data/_loader.py:8-11);
neither is installed here, so empty stub modules are seeded first.  Nothing
the hot path computes touches them.  /root/reference does not exist on the GPU
box: this module is only used by oracle/make_golden.py (fixture generation)
and by tests that skip when the reference is absent.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "skgstat"))


def import_reference():
    if "skgstat" in sys.modules:
        return sys.modules["skgstat"]
    if not available():
        raise ImportError("reference not present at %s" % REFERENCE_ROOT)
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
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import skgstat  # noqa: E402
    return skgstat
