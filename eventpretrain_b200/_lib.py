"""ctypes binding of the C ABI declared in include/eventpretrain_b200.h.

There is NO CPU fallback: if the CUDA library has not been built, or no CUDA device is present,
every operator raises.  (The CPU oracle under oracle/ is test infrastructure and is never
imported from this package.)
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeventpretrain_b200.so")

EP_U8, EP_I8, EP_U16, EP_I16, EP_I32, EP_I64, EP_F32, EP_F64, EP_U32 = range(1, 10)
EP_NORM_COUNT, EP_NORM_MEM, EP_NORM_MEM_GUARD = 1, 2, 3
EP_ORDER_CPQ, EP_ORDER_PQC = 0, 1
EP_BIN_FORCE_GLOBAL, EP_BIN_FORCE_TILED, EP_BIN_FORCE_PLANE = 1, 4, 8
EP_EINVAL, EP_EWORKSPACE, EP_EUNSUPPORTED, EP_EALIGN = -1, -2, -3, -4
EP_RESIZE_NEAREST, EP_RESIZE_BILINEAR, EP_RESIZE_BICUBIC = 0, 1, 2

c_void_p, c_int, c_int64, c_size_t, c_float, c_double = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                                         ctypes.c_size_t, ctypes.c_float, ctypes.c_double)


P = ctypes.POINTER


class EventsSoa(ctypes.Structure):
    _fields_ = [("x", c_void_p), ("y", c_void_p), ("t", c_void_p), ("p", c_void_p),
                ("xy_dtype", c_int), ("t_dtype", c_int), ("p_dtype", c_int), ("batch", c_int),
                ("t_div", c_double), ("offsets", c_void_p), ("offsets_host", c_void_p), ("t_base", c_void_p)]


class EventsAos(ctypes.Structure):
    _fields_ = [("events", c_void_p), ("dtype", c_int), ("n", c_int64)]


class ProfileStats(ctypes.Structure):
    _fields_ = [("ms", c_double * 3), ("launches", c_int * 3)]


class ViewParams(ctypes.Structure):
    _fields_ = [("crop_x", c_int), ("crop_y", c_int), ("crop_w", c_int), ("crop_h", c_int), ("hflip", c_int),
                ("time_flip", c_int), ("negate", c_int), ("reserved", c_int)]


class BinParams(ctypes.Structure):
    _fields_ = [("height", c_int), ("width", c_int), ("num_bins", c_int), ("count_channels", c_int),
                ("scale_x", c_double), ("scale_y", c_double), ("time_f32", c_int), ("flags", c_int)]


# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    "ep_abi_version": (c_int, []),
    "ep_status_string": (ctypes.c_char_p, [c_int]),
    "ep_launch_count": (ctypes.c_ulonglong, []),
    "ep_profile_enable": (c_int, [c_int]),
    "ep_profile_read": (c_int, [P(ProfileStats)]),
    "ep_bin_events_workspace_bytes": (c_size_t, [P(BinParams), c_int, P(c_size_t)]),
    "ep_bin_events_workspace_bytes_for": (c_size_t, [P(EventsSoa), P(BinParams)]),
    "ep_reshape_axis_multiplier_host": (c_int, [c_double, c_void_p]),
    "ep_bin_events": (c_int, [c_void_p, P(EventsSoa), P(BinParams), c_void_p, c_void_p, c_void_p, c_void_p,
                              c_size_t, c_void_p]),
    "ep_bin_events_stats": (c_int, [c_void_p, P(EventsSoa), P(BinParams), c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p, c_void_p]),
    "ep_plane_statistics_workspace_bytes": (c_size_t, [c_int]),
    "ep_plane_statistics": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t]),
    "ep_bin_events_aos": (c_int, [c_void_p, P(EventsAos), P(BinParams), c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    "ep_normalise_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ep_normalise": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t]),
    "ep_mem_hotpixel_workspace_bytes": (c_size_t, [c_int]),
    "ep_mem_hotpixel": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_size_t]),
    "ep_evrep_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int64]),
    "ep_evrep_workspace_bytes_for": (c_size_t, [c_void_p, c_int, c_int]),
    "ep_evrep": (c_int, [c_void_p, P(EventsSoa), c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ep_time_surface_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ep_time_surface": (c_int, [c_void_p, P(EventsSoa), c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ep_view_augment": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ep_diffmap_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_void_p]),
    "ep_patchify_normpix": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "ep_target_patch_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                     c_void_p]),
    "ep_target_patch_loss_masked": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                            c_float, c_void_p]),
    "ep_swin_group_windows_host": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ep_collate_aos_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_int]),
    "ep_pack_transport_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int]),
    "ep_collate_transport4_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "ep_mask_from_noise": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ep_patch_density": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "ep_gather_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ep_patchify_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p]),
    "ep_block_mask_expand": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ep_swin_apply_mask": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "ep_swin_scatter_dense": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ep_gather_tokens_nchw": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ep_scatter_add_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ep_sum_over_batch": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "ep_unshuffle_tokens_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ep_scatter_add_tokens_nchw": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ep_swin_group_tables_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ep_gather_sum_nchw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ep_unshuffle_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                    c_void_p]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Load the C-ABI library; fail loudly (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m eventpretrain_b200.build` "
            "(nvcc, sm_100a).  eventpretrain_b200 has no CPU / PyTorch fallback by design.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here == header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().ep_status_string(status).decode()
        raise RuntimeError(f"{what} failed: {msg} (status {status})")
