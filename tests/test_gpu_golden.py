"""GPU: the CUDA path, through the C ABI, against the committed golden vectors (outputs of the
reference's own classes, tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

from golden_util import NAMES, load
from pgrc_b200 import matcher

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_reference_golden(name):
    g = load(name)
    p = g["params"]
    got = matcher.map_reads_into_pg(g["text"], g["lq_packed"], g["n_packed"] if len(g["n_packed"]) else None, g["read_len"],
                                    rev_compl_pg=p["rev_compl"], pre_reads_exact_matching_chars=p["pre_seed"],
                                    reads_exact_matching_chars=p["seed"], min_chars_per_mismatch=p["min_chars_per_mismatch"],
                                    pre_matching_mode=p["pre_mode"], matching_mode=p["mode"])
    assert np.array_equal(got.pos, g["pos"])
    assert np.array_equal(got.rc, g["rc"])
    assert np.array_equal(got.mm, g["mm"])
    assert got.matched == int(g["matched"])
    exact_only = g["read_len"] == min(p["seed"], g["read_len"]) and p["pre_seed"] == 0
    if not exact_only:
        assert np.array_equal(got.per_mm, g["per_mm"])
    assert got.stats["candidates"] > 0
