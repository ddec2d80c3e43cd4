"""CPU: the C-ABI library loads and exports every symbol include/pgrc_gpu_matcher.h declares; without a
GPU the product fails loudly (no CPU fallback); the product never touches oracle/."""
import ctypes
import os
import re

import pytest

from pgrc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pgrc_gpu_matcher.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgm_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.pgm_abi_version() == 1


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.PgmStats) == 8 * (1 + 256 + 6)
    assert ctypes.sizeof(_lib.PgmAccumulators) == 8 * 6
    assert ctypes.sizeof(_lib.PgmTimings) == 16 * len(_lib.KERNEL_NAMES)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.pgm_create(0, ctypes.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.pgm_last_error(None)
    from pgrc_b200 import matcher
    with pytest.raises(matcher.PgmError):
        matcher.GpuReadsMatcher(0)


def test_shard_plan_is_callable_without_gpu():
    from pgrc_b200 import matcher
    for pg_len in (0, 1, 1000, 12345, 10**6 + 7, 9 * 10**9):
        for world in (1, 2, 3, 8):
            prev_end = 0
            for rank in range(world):
                sb, sl, ob, oe = matcher.shard_plan(pg_len, rank, world)
                assert ob == prev_end and oe >= ob
                prev_end = oe
                assert sb % 32 == 0 and sb <= ob and sb + sl >= oe and sb + sl <= pg_len
                assert ob == 0 or sb == 0 or ob - sb >= _lib.PGM_SHARD_HALO
                assert oe == pg_len or sb + sl == pg_len or sb + sl - oe >= _lib.PGM_SHARD_HALO
            assert prev_end == pg_len


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pgrc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import oracle|from oracle)", src, flags=re.M), f
                assert "pgrc_oracle" not in src and "libpgrc_ref" not in src, f


def test_route_buffer_layout_and_calls_without_gpu():
    assert ctypes.sizeof(_lib.PgmRouteBuffer) == 8 + 8 + 4 + 4 + 8 * _lib.PGM_ROUTE_MAX_WORLD
    lib = _lib.load()
    assert lib.pgm_route_config(None, 0, 1, None, 0) == -1          # null context: invalid argument, no crash
    from pgrc_b200 import matcher
    assert matcher.read_ranges(10, 3) == [0, 3, 6, 10] and matcher.read_ranges(0, 2) == [0, 0, 0]
