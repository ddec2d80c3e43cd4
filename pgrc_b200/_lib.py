"""ctypes binding of the C ABI declared in include/pgrc_gpu_matcher.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There
is no fallback: if it is missing, importing the product fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgrc_gpu.so")

PGM_OK = 0
PGM_DISABLED_PREFIX_MODE = 0xFFFF
PGM_SHARD_HALO = 512

STATUS_NAMES = {0: "PGM_OK", -1: "PGM_ERR_INVALID_ARG", -2: "PGM_ERR_NO_DEVICE", -3: "PGM_ERR_CUDA",
                -4: "PGM_ERR_OOM", -5: "PGM_ERR_BAD_SYMBOL", -6: "PGM_ERR_UNSUPPORTED", -7: "PGM_ERR_STATE"}

# every symbol include/pgrc_gpu_matcher.h declares
EXPORTS = ["pgm_abi_version", "pgm_create", "pgm_destroy", "pgm_last_error", "pgm_set_stream", "pgm_synchronize", "pgm_upload",
           "pgm_set_text", "pgm_set_text_shard", "pgm_shard_plan", "pgm_set_reads", "pgm_match_begin", "pgm_match_begin_interleaved",
           "pgm_scan_pass", "pgm_get_accumulators", "pgm_put_accumulators", "pgm_resolve_pass", "pgm_get_results", "pgm_get_mismatches", "pgm_copmem_begin", "pgm_copmem_pass", "pgm_map_reads",
           "pgm_kernel_launches", "pgm_set_tuning", "pgm_set_profiling", "pgm_get_timings",
           "pgm_route_config", "pgm_route_rounds", "pgm_route_slot", "pgm_route_export", "pgm_route_pull", "pgm_route_scan_launch", "pgm_route_probe_launch", "pgm_route_fetch", "pgm_route_begin", "pgm_route_recv", "pgm_route_build", "pgm_route_scan", "pgm_route_probe",
           "pgm_route_verify",
           "pgm_group_create", "pgm_group_destroy", "pgm_group_last_error", "pgm_group_size", "pgm_group_set_text", "pgm_group_set_reads",
           "pgm_group_upload", "pgm_group_match_begin", "pgm_group_pass", "pgm_group_copmem_begin", "pgm_group_copmem_pass",
           "pgm_group_get_results", "pgm_group_get_mismatches",
           "pgm_mem_index", "pgm_mem_match", "pgm_mem_get_matches",
           "pgm_group_mem_index", "pgm_group_mem_match", "pgm_group_mem_get_matches", "pgm_mem_match_share", "pgm_mem_get_share"]

KERNEL_NAMES = ["pack_text", "rc_text", "unpack_reads", "init_state", "build_table", "scan", "resolve", "finalize", "accumulators",
                "scan_filter", "scan_probe", "scan_verify", "mismatches", "copmem_index", "copmem_query",
                "route_build", "route_scan", "route_probe", "route_verify", "mem_pack", "mem_query", "mem_emit",
                "copmem_stage1", "copmem_stage2"]
PGM_ROUTE_MAX_WORLD = 16
PGM_ROUTE_PATTERNS, PGM_ROUTE_WINDOWS, PGM_ROUTE_CANDIDATES = 0, 1, 2


class PgmStats(ctypes.Structure):
    _fields_ = [("matched", ctypes.c_uint64), ("per_mm", ctypes.c_uint64 * 256),
                ("patterns_inserted", ctypes.c_uint64), ("table_slots", ctypes.c_uint64),
                ("candidates", ctypes.c_uint64), ("verified", ctypes.c_uint64),
                ("accepted", ctypes.c_uint64), ("filter_positives", ctypes.c_uint64)]


class PgmTimings(ctypes.Structure):
    _fields_ = [("ms", ctypes.c_double * len(KERNEL_NAMES)), ("launches", ctypes.c_uint64 * len(KERNEL_NAMES))]


class PgmAccumulators(ctypes.Structure):
    _fields_ = [("best_key", ctypes.c_void_p), ("first_other_order", ctypes.c_void_p),
                ("same_pos_mask", ctypes.c_void_p), ("same_pos_mm", ctypes.c_void_p),
                ("touched", ctypes.c_void_p), ("n_reads", ctypes.c_uint64)]


class PgmRouteBuffer(ctypes.Structure):
    _fields_ = [("base", ctypes.c_void_p), ("stride_bytes", ctypes.c_uint64), ("entry_bytes", ctypes.c_uint32),
                ("world", ctypes.c_uint32), ("count", ctypes.c_uint64 * PGM_ROUTE_MAX_WORLD)]


class PgmRoutePeer(ctypes.Structure):
    _fields_ = [("pid", ctypes.c_uint64), ("ptr", ctypes.c_void_p), ("device", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("ipc_handle", ctypes.c_ubyte * 64)]


class PgmError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). pgrc_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, u64, u32, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    lib.pgm_abi_version.restype = ci
    lib.pgm_create.restype = ci; lib.pgm_create.argtypes = [ci, ctypes.POINTER(vp)]
    lib.pgm_destroy.restype = None; lib.pgm_destroy.argtypes = [vp]
    lib.pgm_last_error.restype = ctypes.c_char_p; lib.pgm_last_error.argtypes = [vp]
    lib.pgm_set_stream.restype = ci; lib.pgm_set_stream.argtypes = [vp, vp]
    lib.pgm_synchronize.restype = ci; lib.pgm_synchronize.argtypes = [vp]
    lib.pgm_upload.restype = ci; lib.pgm_upload.argtypes = [vp]
    lib.pgm_set_text.restype = ci; lib.pgm_set_text.argtypes = [vp, vp, u64]
    lib.pgm_set_text_shard.restype = ci; lib.pgm_set_text_shard.argtypes = [vp, vp, u64, u64, u64, u64, u64]
    lib.pgm_shard_plan.restype = ci
    lib.pgm_shard_plan.argtypes = [u64, ci, ci] + [ctypes.POINTER(u64)] * 4
    lib.pgm_set_reads.restype = ci; lib.pgm_set_reads.argtypes = [vp, vp, u32, vp, u32, u32]
    lib.pgm_match_begin.restype = ci; lib.pgm_match_begin.argtypes = [vp, u32, u32, u32, u32, ci]
    lib.pgm_match_begin_interleaved.restype = ci; lib.pgm_match_begin_interleaved.argtypes = [vp, u32, u32, u32, u32, ci]
    lib.pgm_scan_pass.restype = ci; lib.pgm_scan_pass.argtypes = [vp, ci]
    lib.pgm_get_accumulators.restype = ci; lib.pgm_get_accumulators.argtypes = [vp, ctypes.POINTER(PgmAccumulators)]
    lib.pgm_put_accumulators.restype = ci; lib.pgm_put_accumulators.argtypes = [vp]
    lib.pgm_resolve_pass.restype = ci; lib.pgm_resolve_pass.argtypes = [vp, ci]
    lib.pgm_get_results.restype = ci; lib.pgm_get_results.argtypes = [vp, vp, vp, vp, ctypes.POINTER(PgmStats)]
    lib.pgm_get_mismatches.restype = ci
    lib.pgm_get_mismatches.argtypes = [vp, vp, vp, vp, u64, ctypes.POINTER(u64)]
    lib.pgm_copmem_begin.restype = ci; lib.pgm_copmem_begin.argtypes = [vp, u32, u32, u32, ci]
    lib.pgm_copmem_pass.restype = ci; lib.pgm_copmem_pass.argtypes = [vp, ci]
    lib.pgm_map_reads.restype = ci
    lib.pgm_map_reads.argtypes = [vp, u32, u32, u32, u32, ctypes.c_char, ctypes.c_char, ci, vp, vp, vp,
                                  ctypes.POINTER(PgmStats)]
    lib.pgm_kernel_launches.restype = u64; lib.pgm_kernel_launches.argtypes = [vp]
    lib.pgm_set_tuning.restype = ci; lib.pgm_set_tuning.argtypes = [vp, ci, ci, ci, ci]
    lib.pgm_set_profiling.restype = ci; lib.pgm_set_profiling.argtypes = [vp, ci]
    lib.pgm_get_timings.restype = ci; lib.pgm_get_timings.argtypes = [vp, ctypes.POINTER(PgmTimings)]
    rb = ctypes.POINTER(PgmRouteBuffer)
    lib.pgm_route_config.restype = ci; lib.pgm_route_config.argtypes = [vp, ci, ci, ctypes.POINTER(u64), u64]
    lib.pgm_route_rounds.restype = ci; lib.pgm_route_rounds.argtypes = [vp, ctypes.POINTER(u32)]
    lib.pgm_route_slot.restype = ci; lib.pgm_route_slot.argtypes = [vp, ci]
    lib.pgm_route_scan_launch.restype = ci; lib.pgm_route_scan_launch.argtypes = [vp, ci, u32]
    lib.pgm_route_probe_launch.restype = ci; lib.pgm_route_probe_launch.argtypes = [vp, ci, u32, ctypes.POINTER(u64)]
    lib.pgm_route_fetch.restype = ci; lib.pgm_route_fetch.argtypes = [vp, ci, rb]
    lib.pgm_route_export.restype = ci; lib.pgm_route_export.argtypes = [vp, ci, ctypes.POINTER(PgmRoutePeer)]
    lib.pgm_route_pull.restype = ci; lib.pgm_route_pull.argtypes = [vp, ci, ctypes.POINTER(PgmRoutePeer), ctypes.POINTER(PgmRouteBuffer)]
    lib.pgm_route_begin.restype = ci; lib.pgm_route_begin.argtypes = [vp, u32, u32, u32, u32, ci, rb]
    lib.pgm_route_recv.restype = ci; lib.pgm_route_recv.argtypes = [vp, ci, u64, ctypes.POINTER(vp)]
    lib.pgm_route_build.restype = ci; lib.pgm_route_build.argtypes = [vp, u64]
    lib.pgm_route_scan.restype = ci; lib.pgm_route_scan.argtypes = [vp, ci, u32, rb]
    lib.pgm_route_probe.restype = ci; lib.pgm_route_probe.argtypes = [vp, ci, u32, ctypes.POINTER(u64), rb]
    lib.pgm_route_verify.restype = ci; lib.pgm_route_verify.argtypes = [vp, ci, u64]
    lib.pgm_group_create.restype = ci; lib.pgm_group_create.argtypes = [ci, ctypes.POINTER(ci), ctypes.POINTER(vp)]
    lib.pgm_group_destroy.restype = None; lib.pgm_group_destroy.argtypes = [vp]
    lib.pgm_group_last_error.restype = ctypes.c_char_p; lib.pgm_group_last_error.argtypes = [vp]
    lib.pgm_group_size.restype = ci; lib.pgm_group_size.argtypes = [vp]
    lib.pgm_group_set_text.restype = ci; lib.pgm_group_set_text.argtypes = [vp, vp, u64]
    lib.pgm_group_set_reads.restype = ci; lib.pgm_group_set_reads.argtypes = [vp, vp, u32, vp, u32, u32]
    lib.pgm_group_upload.restype = ci; lib.pgm_group_upload.argtypes = [vp]
    lib.pgm_group_match_begin.restype = ci; lib.pgm_group_match_begin.argtypes = [vp, u32, u32, u32, u32, ci, ci]
    lib.pgm_group_pass.restype = ci; lib.pgm_group_pass.argtypes = [vp, ci]
    lib.pgm_group_copmem_begin.restype = ci; lib.pgm_group_copmem_begin.argtypes = [vp, u32, u32, u32, ci]
    lib.pgm_group_copmem_pass.restype = ci; lib.pgm_group_copmem_pass.argtypes = [vp, ci]
    lib.pgm_group_get_results.restype = ci; lib.pgm_group_get_results.argtypes = [vp, vp, vp, vp, ctypes.POINTER(PgmStats)]
    lib.pgm_group_get_mismatches.restype = ci; lib.pgm_group_get_mismatches.argtypes = [vp, vp, vp, vp, u64, ctypes.POINTER(u64)]
    lib.pgm_mem_index.restype = ci; lib.pgm_mem_index.argtypes = [vp, u32, u32, ctypes.POINTER(u32)]
    lib.pgm_mem_match.restype = ci; lib.pgm_mem_match.argtypes = [vp, vp, u64, ci, ci, u32, ctypes.POINTER(u64)]
    lib.pgm_mem_get_matches.restype = ci; lib.pgm_mem_get_matches.argtypes = [vp, vp, u64]
    lib.pgm_mem_match_share.restype = ci; lib.pgm_mem_match_share.argtypes = [vp, vp, u64, ci, ci, u32, ci, ci, ctypes.POINTER(u64)]
    lib.pgm_mem_get_share.restype = ci; lib.pgm_mem_get_share.argtypes = [vp, vp, vp, u64]
    lib.pgm_group_mem_index.restype = ci; lib.pgm_group_mem_index.argtypes = [vp, u32, u32, ctypes.POINTER(u32)]
    lib.pgm_group_mem_match.restype = ci; lib.pgm_group_mem_match.argtypes = [vp, vp, u64, ci, ci, u32, ctypes.POINTER(u64)]
    lib.pgm_group_mem_get_matches.restype = ci; lib.pgm_group_mem_get_matches.argtypes = [vp, vp, u64]
    _lib = lib
    return lib
