import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    """Golden case as a dict; cases that share the common synthetic rollout get it merged in."""
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    if "actions" not in d and name.split("_")[0] in ("mppi", "mppiq", "cem", "dmd", "rs"):
        d.update(dict(np.load(os.path.join(GOLDEN, "common.npz"))))
    return {k: (v.item() if v.ndim == 0 else v) for k, v in d.items()}
