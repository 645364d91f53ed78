"""ctypes binding of libcns_sm100.so (declared in include/cns_b200.h).

There is deliberately no fallback: if the shared library is missing or a CUDA device
is not present the product path raises.  PyTorch is used only to own device memory
and streams; every tensor crosses the boundary as a raw pointer.
"""
import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcns_sm100.so")
if os.environ.get("CNS_LIB"):           # another build of the same library (A-B runs, scripts/build_variant.sh)
    LIB_PATH = os.path.abspath(os.environ["CNS_LIB"])

CNS_OK = 0
PREC_FP32, PREC_F16, PREC_BF16 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "f16": PREC_F16, "bf16": PREC_BF16}


class CnsError(RuntimeError):
    pass


class cns_batch(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int64), ("n_clusters", C.c_int64), ("n_graphs", C.c_int64),
        ("n_e01", C.c_int64), ("n_e11_cur", C.c_int64), ("n_e11_tar", C.c_int64),
        ("x_cur", C.c_void_p), ("x_tar", C.c_void_p), ("pos_cur", C.c_void_p), ("pos_tar", C.c_void_p),
        ("l0_to_l1_j", C.c_void_p), ("l0_to_l1_i", C.c_void_p),
        ("l1_edge_cur", C.c_void_p), ("l1_edge_tar", C.c_void_p),
        ("l1_cur_stride", C.c_int64), ("l1_tar_stride", C.c_int64),
        ("cluster_mask", C.c_void_p), ("cluster_centers", C.c_void_p),
        ("node_graph", C.c_void_p), ("num_clusters", C.c_void_p), ("tPo_norm", C.c_void_p),
        ("flags", C.c_int64), ("max_clusters_per_graph", C.c_int64), ("target_cache", C.c_void_p),
        ("clu_rowptr", C.c_void_p),
    ]


BATCH_L1_CUR_COMPLETE, BATCH_L1_TAR_COMPLETE = 1, 2


_lib = None


def _declare(lib):
    vp, i64, i32, sz = C.c_void_p, C.c_int64, C.c_int, C.c_size_t
    sig = {
        "cns_version": (C.c_char_p, []),
        "cns_last_error_string": (C.c_char_p, []),
        "cns_device_count": (i32, []),
        "cns_create": (i32, [C.POINTER(vp), i32]),
        "cns_destroy": (i32, [vp]),
        "cns_load_weights": (i32, [vp, C.POINTER(vp), C.POINTER(i64), i32, vp]),
        "cns_load_weights_device": (i32, [vp, C.POINTER(vp), C.POINTER(i64), i32, vp]),
        "cns_param_name": (C.c_char_p, [i32]),
        "cns_param_numel": (i64, [i32]),
        "cns_set_precision": (i32, [vp, i32]),
        "cns_set_tc_tile": (i32, [vp, i32]),
        "cns_set_option": (i32, [vp, C.c_char_p, C.c_longlong]),
        "cns_get_counter": (C.c_longlong, [vp, C.c_char_p]),
        "cns_profile_begin": (i32, [vp]),
        "cns_profile_end": (i32, [vp, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
        "cns_forward_workspace_bytes": (sz, [C.POINTER(cns_batch)]),
        "cns_forward_save_bytes": (sz, [C.POINTER(cns_batch)]),
        "cns_graphvs_forward": (i32, [vp, C.POINTER(cns_batch), vp, vp, vp, vp, vp, vp, sz, vp, sz, vp]),
        "cns_forward_status": (i32, [vp, vp]),
        "cns_objectives": (i32, [vp, vp, vp, i64, C.c_float, vp, vp, vp, vp]),
        "cns_postprocess": (i32, [vp, vp, vp, i64, vp, vp]),
        "cns_clip_adamw_workspace_bytes": (sz, []),
        "cns_clip_adamw": (i32, [vp, vp, vp, vp, vp, i64] + [C.c_float] * 6 + [i64, C.c_float, vp, vp, sz, vp]),
        "cns_backward_workspace_bytes": (sz, [C.POINTER(cns_batch)]),
        "cns_param_offset": (i64, [vp, i32]),
        "cns_graphvs_backward": (i32, [vp, C.POINTER(cns_batch), vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, sz, vp]),
        "cns_graphvs_forward_host": (i32, [vp, C.POINTER(cns_batch), vp, vp, vp, vp, vp]),
        "cns_servo_step_host": (i32, [vp, C.POINTER(cns_batch), i32, vp, vp, vp]),
        "cns_csr_by_dst": (i32, [vp, i64, i64, vp, vp, vp, sz, vp]),
        "cns_edgeconv_max": (i32, [vp, vp, vp, i64, i32, vp, vp]),
        "cns_graphnorm_relu": (i32, [vp, vp, i64, vp, vp, vp, vp, vp]),
        "cns_linear": (i32, [vp, i32, vp, i32, vp, vp, vp, i32, i64, i32, vp, vp]),
        "cns_target_cache_create": (i32, [vp, C.POINTER(cns_batch), vp, C.POINTER(vp)]),
        "cns_target_cache_destroy": (i32, [vp]),
        "cns_frame_mask": (i32, [vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, sz, vp]),
        "cns_graph_mask_workspace_bytes": (sz, [i64, i64, i64]),
        "cns_graph_mask": (i32, [vp, vp, vp, vp, vp, i64, i64, i64, vp, vp, vp, i64, vp, vp, vp, sz, vp]),
        "cns_linear_split_workspace_bytes": (sz, []),
        "cns_linear_split": (i32, [vp, i64, vp, vp, vp, vp, i32, vp, vp, sz, vp]),
        "cns_weight_grad_workspace_bytes": (sz, []),
        "cns_weight_grad": (i32, [vp, i64, vp, i64, i64, vp, vp, i32, i32, vp, sz, vp]),
        "cns_align_matches": (i32, [vp, i64, vp, i64, i64, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp, vp, sz, vp]),
        "cns_pixel_stop_workspace_bytes": (sz, [i64]),
        "cns_pixel_stop_error": (i32, [vp, vp, vp, vp, vp, i64, i64, C.c_float, C.c_float, vp, vp, vp, vp, vp, sz, vp]),
        "cns_label_nearest": (i32, [vp, vp, i64, vp, vp, vp, vp]),
        "cns_ap_similarity": (i32, [vp, vp, vp, vp, i64, vp, vp]),
        "cns_segment_median_workspace_bytes": (sz, [i64]),
        "cns_segment_median": (i32, [vp, vp, i64, i64, vp, vp, sz, vp]),
        "cns_ap_prepare": (i32, [vp, vp, vp, vp, vp, vp, i64, vp]),
        "cns_ap_labels": (i32, [vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, sz, vp]),
        "cns_graph_layout": (i32, [vp, vp, vp, vp, i64, vp, vp, vp, vp]),
        "cns_ap_workspace_bytes": (sz, [i64, i64, i64, i32]),
        "cns_affinity_propagation": (i32, [vp, vp, vp, vp, i64, i64, i64, C.c_double, i32, i32, vp, vp, vp, vp, sz, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)   # AttributeError here == header / library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    return sig


EXPORTED_SYMBOLS = None


def load():
    """dlopen the in-tree library (no compute, works without a GPU)."""
    global _lib, EXPORTED_SYMBOLS
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise CnsError("%s not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or ./build.sh (there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        EXPORTED_SYMBOLS = sorted(_declare(lib).keys())
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != CNS_OK:
        msg = load().cns_last_error_string().decode()
        raise CnsError("%s failed with status %d: %s" % (what or "libcns_sm100", rc, msg))


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def param_names():
    lib = load()
    return [lib.cns_param_name(k).decode() for k in range(43)]


class Handle(object):
    """Owns a cns_handle (packed device weights)."""

    def __init__(self, device: torch.device):
        lib = load()
        if not torch.cuda.is_available():
            raise CnsError("cns_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device)
        h = C.c_void_p()
        check(lib.cns_create(C.byref(h), self.device.index or 0), "cns_create")
        self._h = h
        self.precision = "fp32"
        tile = os.environ.get("CNS_TC_TILE")
        if tile:
            check(lib.cns_set_tc_tile(h, int(tile)), "cns_set_tc_tile")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                load().cns_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def raw(self):
        return self._h

    def set_precision(self, name: str):
        check(load().cns_set_precision(self._h, PRECISIONS[name]), "cns_set_precision")
        self.precision = name

    def set_option(self, name: str, value: int):
        check(load().cns_set_option(self._h, name.encode(), int(value)), "cns_set_option(%s)" % name)

    def counter(self, name: str) -> int:
        return int(load().cns_get_counter(self._h, name.encode()))

    def load_weights(self, tensors, on_device: bool):
        lib = load()
        assert len(tensors) == 43
        keep = [t.detach().contiguous().float() for t in tensors]
        ptrs = (C.c_void_p * 43)(*[t.data_ptr() for t in keep])
        numels = (C.c_int64 * 43)(*[t.numel() for t in keep])
        if on_device:
            check(lib.cns_load_weights_device(self._h, ptrs, numels, 43, stream_ptr(self.device)),
                  "cns_load_weights_device")
            self._keepalive = keep   # the D2D copies are asynchronous
        else:
            check(lib.cns_load_weights(self._h, ptrs, numels, 43, stream_ptr(self.device)), "cns_load_weights")
