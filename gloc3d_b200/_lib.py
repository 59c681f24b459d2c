"""ctypes binding of libgloc3d.so (the C ABI declared in include/gloc3d.h).

The library is the product: there is no Python or CPU fallback.  If the shared
object is missing, importing this module raises; if no sm_100 GPU is usable, every
compute entry point returns GLOC_ERR_CUDA and the wrappers raise GlocError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgloc3d.so")

GLOC_OK = 0
GLOC_ERR_INVALID = 1
GLOC_ERR_CUDA = 2
GLOC_ERR_NOT_BUILT = 3
GLOC_ERR_RANGE = 4
GLOC_ERR_NOMEM = 5

KNN_AUTO, KNN_EXACT_SCAN, KNN_SHORTLIST = 0, 1, 2
LOC_VERIFY_ALL, LOC_FIRST_MATCH = 0, 1


class GlocError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgloc3d error {code}: {msg}")
        self.code = code


class KnnStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("queries", "kernel_launches", "shortlist_queries",
                                          "fallback_queries", "shortlist_rows", "last_mode")]


class CsmResult(C.Structure):
    _fields_ = [("found", C.c_int), ("score", C.c_float), ("scan_index", C.c_int),
                ("x_offset", C.c_int), ("y_offset", C.c_int), ("reserved", C.c_int),
                ("pose_x", C.c_double), ("pose_y", C.c_double), ("pose_yaw", C.c_double)]

    def as_tuple(self):
        return (self.found, self.score, self.scan_index, self.x_offset, self.y_offset,
                self.pose_x, self.pose_y, self.pose_yaw)


class Profile(C.Structure):
    _fields_ = [("dominant_ms", C.c_double), ("dominant_launches", C.c_uint64)]


class CsmStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("matches", "kernel_launches", "coarse_candidates",
                                          "refined_nodes")]


class GridInfo(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("resolution", C.c_double),
                ("max_x", C.c_double), ("max_y", C.c_double)]


class LocParams(C.Structure):
    _fields_ = [("k", C.c_int), ("n_lin", C.c_int), ("n_ang", C.c_int), ("ang_step", C.c_double),
                ("depth", C.c_int), ("min_score", C.c_float), ("policy", C.c_int)]


class LocResult(C.Structure):
    _fields_ = [("located", C.c_int), ("candidate", C.c_int), ("best_candidate", C.c_int),
                ("n_verified", C.c_int), ("db_index", C.c_uint64), ("match", CsmResult)]


class LocStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("queries", "pairs_verified", "waves", "kernel_launches", "pairs_migrated")]


class LocProfile(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("retrieval_ms", C.c_double), ("calls", C.c_uint64)]


class BevInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("min_ix", C.c_int), ("min_iy", C.c_int),
                ("ox", C.c_double), ("oy", C.c_double), ("resolution", C.c_double),
                ("n_occupied", C.c_uint64), ("n_points_in_range", C.c_uint64)]


# name -> (restype, argtypes); mirrors include/gloc3d.h one to one
_vp, _sz, _i, _d, _f = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_float
_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
SIGNATURES = {
    "gloc_version": (_i, []),
    "gloc_last_error": (C.c_char_p, []),
    "gloc_device_count": (_i, []),
    "gloc_knn_create": (_i, [C.POINTER(_vp), _sz, _i]),
    "gloc_knn_destroy": (None, [_vp]),
    "gloc_knn_set_db": (_i, [_vp, _vp, _sz]),
    "gloc_knn_set_db_device": (_i, [_vp, _vp, _sz]),
    "gloc_knn_append": (_i, [_vp, _vp, _sz]),
    "gloc_knn_size": (_sz, [_vp]),
    "gloc_knn_dim": (_sz, [_vp]),
    "gloc_knn_set_search_limit": (_i, [_vp, _sz]),
    "gloc_knn_set_index_offset": (_i, [_vp, C.c_uint64]),
    "gloc_knn_set_mode": (_i, [_vp, _i]),
    "gloc_knn_get_stats": (_i, [_vp, C.POINTER(KnnStats)]),
    "gloc_knn_set_profiling": (_i, [_vp, _i]),
    "gloc_knn_get_profile": (_i, [_vp, C.POINTER(Profile)]),
    "gloc_knn_query": (_i, [_vp, _vp, _sz, _sz, _vp, _vp]),
    "gloc_knn_query_device": (_i, [_vp, _vp, _sz, _sz, _vp, _vp, _vp]),
    "gloc_knn_merge_topk_device": (_i, [_vp, _vp, _sz, _sz, _sz, _vp, _vp, _i, _vp]),
    "gloc_csm_create": (_i, [C.POINTER(_vp), _i]),
    "gloc_csm_destroy": (None, [_vp]),
    "gloc_csm_add_grid_cells": (_i, [_vp, _vp, _i, _i, _d, _d, _d, _ip]),
    "gloc_csm_add_grid_u8": (_i, [_vp, _vp, _i, _i, _d, _d, _d, _ip]),
    "gloc_csm_num_grids": (_i, [_vp]),
    "gloc_csm_store_bytes": (_i, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gloc_comm_unique_id": (_i, [_vp, _sz]),
    "gloc_comm_create": (_i, [C.POINTER(_vp), _vp, _i, _i, _i]),
    "gloc_comm_create_local": (_i, [C.POINTER(_vp), _i, _ip]),
    "gloc_comm_destroy": (None, [_vp]),
    "gloc_comm_rank": (_i, [_vp]),
    "gloc_comm_size": (_i, [_vp]),
    "gloc_comm_nccl_version": (_i, []),
    "gloc_knn_query_sharded": (_i, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _i]),
    "gloc_knn_query_sharded_device": (_i, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _i, _vp]),
    "gloc_loc_localize_sharded": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, C.POINTER(LocParams), _vp, _vp, _vp, _vp, _i]),
    "gloc_loc_create": (_i, [C.POINTER(_vp), _vp, _vp]),
    "gloc_loc_destroy": (None, [_vp]),
    "gloc_loc_set_row_grids": (_i, [_vp, _vp, _sz]),
    "gloc_loc_localize": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, C.POINTER(LocParams), _vp, _vp, _vp, _vp]),
    "gloc_loc_localize_device": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, C.POINTER(LocParams), _vp, _vp, _vp, _vp]),
    "gloc_loc_get_stats": (_i, [_vp, C.POINTER(LocStats)]),
    "gloc_loc_share_grids": (_i, [_vp, _vp]),
    "gloc_loc_unshare_grids": (_i, [_vp, _vp]),
    "gloc_loc_assign_pairs": (_i, [_vp, _sz, _i, _vp]),
    "gloc_loc_set_profiling": (_i, [_vp, _i]),
    "gloc_loc_get_profile": (_i, [_vp, C.POINTER(LocProfile)]),
    "gloc_csm_get_precomputation_grid": (_i, [_vp, _i, _i, _vp]),
    "gloc_csm_match_batch": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _d, _i, _f,
                                  C.POINTER(CsmResult)]),
    "gloc_csm_discretize": (_i, [_vp, _vp, _i, _d, _d, _d, _i, _d, _d, _d, _d, _vp]),
    "gloc_csm_search_params": (_i, [_d, _d, _vp, _i, _d, _ip, _ip, _dp]),
    "gloc_csm_grid_to_points": (_i, [_vp, _i, _i, _d, _d, _d, _vp, _i, _ip]),
    "gloc_csm_get_stats": (_i, [_vp, C.POINTER(CsmStats)]),
    "gloc_csm_set_profiling": (_i, [_vp, _i]),
    "gloc_csm_get_profile": (_i, [_vp, C.POINTER(Profile)]),
    "gloc_bev_create": (_i, [C.POINTER(_vp), _i, _f, _f]),
    "gloc_bev_destroy": (None, [_vp]),
    "gloc_bev_project": (_i, [_vp, _vp, _sz, _i, C.POINTER(BevInfo)]),
    "gloc_bev_get_image": (_i, [_vp, _vp, _sz]),
    "gloc_bev_get_cnn_input": (_i, [_vp, _i, _i, _vp]),
    "gloc_bev_get_cnn_input_roi": (_i, [_vp, _i, _i, _vp, _vp]),
    "gloc_enc_forward_padded_device": (_i, [_vp, _vp, _vp, _i, _vp]),
    "gloc_enc_forward_padded": (_i, [_vp, _vp, _vp, _i, _vp]),
    "gloc_desc_extract_padded": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp]),
    "gloc_bev_get_occupied_points": (_i, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "gloc_bev_kernel_launches": (C.c_uint64, [_vp]),
    "gloc_csm_add_grid_from_bev": (_i, [_vp, _vp, _ip]),
    "gloc_csm_add_grid_from_bev_aligned": (_i, [_vp, _vp, _ip]),
    "gloc_csm_get_grid_info": (_i, [_vp, _i, C.POINTER(GridInfo)]),
    "gloc_grid_file_write": (_i, [C.c_char_p, C.POINTER(GridInfo), C.POINTER(_vp), _sz]),
    "gloc_grid_file_write_tagged": (_i, [C.c_char_p, C.POINTER(GridInfo), C.POINTER(_vp), _sz, C.c_uint32]),
    "gloc_grid_file_tag": (C.c_uint32, [_vp]),
    "gloc_csm_save_grids_tagged": (_i, [_vp, C.c_char_p, C.c_uint32]),
    "gloc_grid_file_open": (_i, [C.c_char_p, C.POINTER(_vp), C.POINTER(_sz)]),
    "gloc_grid_file_next": (_i, [_vp, C.POINTER(GridInfo), _vp, _sz]),
    "gloc_grid_file_close": (None, [_vp]),
    "gloc_csm_save_grids": (_i, [_vp, C.c_char_p]),
    "gloc_csm_load_grids": (_i, [_vp, C.c_char_p, _ip, _ip]),
    "gloc_vlad_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "gloc_vlad_destroy": (None, [_vp]),
    "gloc_vlad_forward_device": (_i, [_vp, _vp, _i, _i, _vp]),
    "gloc_vlad_forward": (_i, [_vp, _vp, _i, _i, _vp]),
    "gloc_vlad_kernel_launches": (C.c_uint64, [_vp]),
    "gloc_enc_create": (_i, [C.POINTER(_vp), _i, _i, _i, C.POINTER(_vp), C.POINTER(_vp)]),
    "gloc_enc_destroy": (None, [_vp]),
    "gloc_enc_feature_shape": (_i, [_vp, _ip, _ip]),
    "gloc_enc_forward_device": (_i, [_vp, _vp, _i, _vp]),
    "gloc_enc_forward": (_i, [_vp, _vp, _i, _vp]),
    "gloc_enc_kernel_launches": (C.c_uint64, [_vp]),
    "gloc_knn_pair_workers": (_i, [_i]),
    "gloc_desc_extract": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "gloc_bench_smem_gather": (_i, [_i, C.POINTER(C.c_double)]),
}

_lib = None


def lib():
    """The loaded library.  Raises ImportError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C gloc3d_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the binding diverge
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != GLOC_OK:
        raise GlocError(rc, lib().gloc_last_error().decode("utf-8", "replace"))
