"""Loading the committed reference fixtures (tests/golden/*.npz, see make_golden.py)."""
import ast
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    out["meta"] = ast.literal_eval(str(z["meta"]))
    return out


def names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


MODULE_CASES = ["sot512_cut", "sot512_nocut", "sot512_logf_cut", "sot2048_cut", "sot2048_nocut",
                "sot2048_logf_unsorted", "sot512_p1_nosquare", "sot512_p3"]


def oracle_kwargs(ctor):
    """Reference ctor kwargs -> oracle/sot_oracle.py keyword arguments."""
    return dict(p=ctor.get("p", 1), square=bool(ctor.get("square_dist", False)),
                cut_scale=bool(ctor.get("dont_normalize", False)),
                limit=bool(ctor.get("limit_quantile_range", False)),
                require_sort=bool(ctor.get("require_sort", True)))
