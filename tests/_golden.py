"""Loader of the committed golden vectors (tests/golden/*.npz; made by tests/golden/make_golden.py from the C oracle:
PARITY UNPINNED, see that script's header)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    spec = json.loads(str(z["spec"]))
    bc = spec["bc"]
    spec["bc"] = (bc[0],) if len(bc) == 1 else (bc[0], tuple(bc[1]) if isinstance(bc[1], list) else bc[1])
    arrays = {k: (np.asfortranarray(z[k]) if z[k].ndim == 2 else z[k]) for k in z.files if k != "spec"}
    for k in ("ms", "qs", "mm"):
        if k in arrays:
            spec[k] = arrays[k]
    return spec, arrays
