"""ctypes binding of libpgs_b200.so (the C ABI declared in include/pgs_b200.h).

There is no CPU fallback: every op in this package goes through this library, and the library
only holds sm_100a device code.  If the shared object is missing (and cannot be built) the
import of the op fails loudly.
"""
import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_size_t, c_void_p, c_char_p, c_float, c_double

import torch

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# name -> (restype, argtypes); kept in one place so tests can check it against the header
SIGNATURES = {
    "pgs_version": (c_int, []),
    "pgs_last_error": (c_char_p, []),
    "pgs_launch_count": (c_int64, []),
    "pgs_cmap_capacity": (c_int64, [c_int64]),
    "pgs_cmap_build_scratch_bytes": (c_size_t, [c_int64]),
    "pgs_cmap_build": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pgs_kmap_build": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                               c_void_p, c_void_p]),
    "pgs_kmap_pairs_scratch_bytes": (c_size_t, [c_int64, c_int32]),
    "pgs_kmap_pairs": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    "pgs_conv_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32,
                             c_int32, c_void_p, c_void_p]),
    "pgs_bq_grid_scratch_bytes": (c_size_t, [c_int64]),
    "pgs_bq_grid_build": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pgs_bq_pack_queries": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "pgs_bq_query": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_float,
                             c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgs_bq_export": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "pgs_nn1_query": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_float,
                              c_int32, c_void_p, c_void_p, c_void_p]),
    "pgs_prop_gt_iou": (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "pgs_prop_nms_scratch_bytes": (c_size_t, [c_int64]),
    "pgs_prop_cross_nms": (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "pgs_rg_init": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "pgs_rg_propagate": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "pgs_hdb_scratch_bytes": (c_size_t, [c_int64, c_int32]),
    "pgs_hdb_search_stats": (c_int, [c_void_p, c_int32]),
    "pgs_hdb_morton_rank": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "pgs_hdb_mst": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_size_t, c_void_p]),
    "pgs_hdb_labels_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_double, c_void_p, c_void_p]),
    "pgs_ms_iterate": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_float, c_int32, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "pgs_ms_assign": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "pgs_conv_tc_supported": (c_int, [c_int32, c_int32]),
    "pgs_conv_tc_scratch_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "pgs_conv_fwd_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32,
                                c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pgs_conv_mma_supported": (c_int, [c_int32, c_int32]),
    "pgs_conv_mma_scratch_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "pgs_conv_fwd_mma": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32,
                                 c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pgs_conv_mma_split_supported": (c_int, [c_int32, c_int32]),
    "pgs_kmap_row_masks": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "pgs_kmap_permute": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "pgs_conv_fwd_mma_split": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32,
                                       c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pgs_bn_forward": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_float,
                               c_float, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgs_bn_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                                c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgs_conv_dw_mma_supported": (c_int, [c_int32, c_int32]),
    "pgs_conv_bwd_weight_mma": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                        c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "pgs_conv_prep_weights_batch": (c_int, [c_void_p, c_int32, c_int64, c_void_p]),
    "pgs_bn_forward_ex": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_float,
                                  c_float, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgs_bn_backward_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgs_add2": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pgs_cat2": (c_int, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int64, c_int32, c_void_p]),
    "pgs_unet_record_bytes": (None, [c_void_p]),
    "pgs_unet_forward": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "pgs_unet_backward_scratch_elems": (c_int64, [c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "pgs_unet_backward": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "pgs_conv_bwd_weight": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                    c_int32, c_int32, c_int32, c_void_p, c_void_p]),
}


c_void_p = ctypes.c_void_p


class PgsError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load():
    """Load (building first if sources are newer and nvcc exists) and return the ctypes handle."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if os.path.exists(_build.NVCC):
        try:
            _build.build()
        except Exception as e:  # pragma: no cover - build errors must be visible
            raise PgsError("building libpgs_b200.so failed: %s" % e)
    if not os.path.exists(path):
        raise PgsError(
            "libpgs_b200.so not found at %s and nvcc is unavailable: the B200 CUDA extension is required "
            "(there is no CPU fallback). Run `python -m panopticsegforlargescalepointcloud_b200.build`." % path
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        raise PgsError(load().pgs_last_error().decode())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """Current CUDA stream of the current device as a void* (torch.cuda.current_stream() costs ~17 us per call in
    Python; the raw getter ~0.3 us -- it is called for every kernel launch)."""
    if _raw_stream is not None:
        return c_void_p(_raw_stream(torch.cuda.current_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise PgsError("expected a CUDA tensor: the B200 path has no CPU implementation")
    if not t.is_contiguous():
        raise PgsError("expected a contiguous tensor")
    return c_void_p(t.data_ptr())


def launch_count():
    return int(load().pgs_launch_count())


class nvtx_range:
    """NVTX range around a host-side phase (coordinate / kernel maps, region growing, HDBSCAN, ...): visible in Nsight
    Systems and `ncu --nvtx`; a no-op costing two cheap calls when no tool is attached."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return False
