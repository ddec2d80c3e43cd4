"""Loader for the committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py
from the reference's own matcher classes)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# (pgmatch_*.npz are the stage-7 vectors of tests/golden/make_golden_pgmatch.py: tests/test_pgmatch_oracle.py, tests/test_gpu_pgmatch.py;
# sample_*.npz are the full-size fixtures of tests/golden/make_fullsize.py: another format, used by tests/test_gpu_fullsize.py and bench.py --verify)
NAMES = sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))) if not n.startswith(("sample_", "pgmatch_")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["params"] = dict(seed=int(g["seed"]), min_chars_per_mismatch=int(g["min_chars_per_mismatch"]),
                       mode=bytes(g["mode"]).decode(), pre_seed=int(g["pre_seed"]), pre_mode=bytes(g["pre_mode"]).decode(),
                       rev_compl=bool(int(g["rev_compl"])))
    g["read_len"] = int(g["read_len"])
    return g
