"""ctypes binding of libciclope_b200.so (include/ciclope_b200.h).

There is NO CPU fallback: if the CUDA library is missing or no CUDA device is present the
public functions raise.  torch is used for device memory, streams and torch.distributed only.
"""
import ctypes
import os

import torch  # noqa: F401  (loads libcudart before our library; provides device memory)

_HERE = os.path.dirname(os.path.abspath(__file__))
# CICLOPE_B200_LIB: alternative build of the same library (kernel tuning experiments, profiles/tools)
LIB_PATH = os.environ.get("CICLOPE_B200_LIB") or os.path.join(_HERE, "libciclope_b200.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_vp = ctypes.c_void_p
c_sz = ctypes.c_size_t


class CfeSetDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("s0", c_i64), ("s1", c_i64), ("r0", c_i64), ("r1", c_i64), ("c0", c_i64), ("c1", c_i64),
                ("vs0", c_i64), ("vs1", c_i64), ("vr0", c_i64), ("vr1", c_i64), ("vc0", c_i64), ("vc1", c_i64),
                ("n_cand", c_i64), ("tile_base", c_i64)]


class CfeItemSeg(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("item0", c_i64), ("lit_off", c_i64), ("lit_len", c_i64), ("id0", c_i64), ("rank0", c_i64)]


_SIGNATURES = {
    "cfe_version": (c_int, []),
    "cfe_last_error": (ctypes.c_char_p, []),
    "cfe_launch_count": (ctypes.c_uint64, []),
    "cfe_pack_bits": (c_int, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp]),
    "cfe_pack_bits_le": (c_int, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp]),
    "cfe_max_u8": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "cfe_value_span_u8": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "cfe_scan_workspace_bytes": (c_sz, [c_i64]),
    "cfe_scan_elements": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_scan_run_starts": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_scan_nodes": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_scan_nodes_index": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_emit_nodes_bits": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "cfe_fill_list": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "cfe_scan_u32": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_ccl_max_runs": (c_sz, [c_i64, c_i64]),
    "cfe_ccl_largest": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_ccl_largest_bits": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_ccl_write_mask_bits": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_ccl_label": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "cfe_ccl_select": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_ccl_write_mask": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_ccl_write_select": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp,
                                     c_vp]),
    "cfe_ccl_largest_select": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_int,
                                       c_vp, c_vp]),
    "cfe_ccl_plane_roots": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_ccl_interface_edges": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "cfe_ccl_interface_forest": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp,
                                         c_vp, c_i64, c_vp]),
    "cfe_ccl_boundary_vertices": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp,
                                          c_vp, c_vp, c_vp]),
    "cfe_ccl_boundary_solve": (c_int, [c_vp, c_int, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_uf_solve": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "cfe_sets_count": (c_int, [c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_sets_fill": (c_int, [c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp,
                              c_vp]),
    "cfe_set_counts": (c_int, [c_vp, c_int, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_sets_onepass_tile": (c_i64, []),
    "cfe_sets_onepass": (c_int, [c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp,
                                 c_vp, c_vp]),
    "cfe_refnode_sums": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_fill_cells": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "cfe_gather_gv": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "cfe_fill_points": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "cfe_format_coords": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "cfe_format_coords_wide": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "cfe_node_lengths_wide": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "cfe_emit_nodes_wide": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_emit_nodes": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "cfe_emit_elements_workspace_bytes": (c_sz, [c_i64]),
    "cfe_element_lengths": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_i64, c_i64,
                                    c_i64, c_int, c_vp, c_vp, c_vp]),
    "cfe_fill_elements": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "cfe_fill_elements_gv": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp,
                                     c_vp]),
    "cfe_emit_elements": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_i64, c_i64,
                                  c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_emit_items": (c_int, [c_vp, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_put_literal": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_gen_node_lengths": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_gen_emit_nodes": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cfe_gen_elem_lengths": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "cfe_gen_emit_elements": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "cfe_gv_histogram": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    "cfe_peer_exchange": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_int, c_i64, c_i64, ctypes.c_uint32,
                                  ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, c_vp, c_vp, c_vp]),
    "cfe_box_andnot_u8": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "cfe_masked_set_u8": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "cfe_masked_keep_u8": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "cfe_disk_dilate_bits": (c_int, [c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_vp, c_vp]),
    "cfe_line_or_bits": (c_int, [c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "cfe_unpack_bits": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp]),
    "cfe_border_flags": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "cfe_gauss1d": (c_int, [c_vp, c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_vp]),
    "cfe_spline_prefilter": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp]),
    "cfe_spline_zoom": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                c_vp, c_int, c_vp, c_vp]),
    "cfe_threshold_f64": (c_int, [c_vp, c_i64, ctypes.c_double, c_vp, c_vp]),
    "cfe_host_pyset_order": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "cfe_proj_bits": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "cfe_scale_crop_f32": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "cfe_threshold_u8": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    "cfe_gv_partition_workspace_bytes": (ctypes.c_size_t, [c_i64]),
    "cfe_gv_partition": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
}

EXPORTED_SYMBOLS = sorted(_SIGNATURES)

_lib = None


class CfeError(RuntimeError):
    pass


def load():
    """Load the CUDA library (no compute).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CfeError("%s not found: build it with `python -m ciclope_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise CfeError("ciclope_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return load()


# When set to a list, every entry-point call is bracketed by CUDA events on the current stream and
# (name, start_event, stop_event) is appended: bench.py's per-kernel timing.  None = off.
PROFILE = None


def call(name, *args):
    """Call an int-returning entry point; raise CfeError with cfe_last_error() on failure."""
    lib = load()
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        PROFILE.append((name, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise CfeError("%s failed (%d): %s" % (name, rc, lib.cfe_last_error().decode()))


def ptr(t):
    """Device (or pinned host) pointer of a torch tensor, or NULL for None."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
