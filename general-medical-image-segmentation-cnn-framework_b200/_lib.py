"""ctypes binding of libb200seg.so (the C ABI declared in include/b200seg.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200seg.so")

_P, _I, _L, _F, _D, _Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_size_t
_CODES = {"p": _P, "i": _I, "l": _L, "f": _F, "d": _D, "z": _Z, "Q": ctypes.c_uint64}


class ConvGeom(ctypes.Structure):
    """b200seg_conv_geom"""
    _fields_ = [(n, ctypes.c_int32) for n in
                ("n", "d", "h", "w", "cin", "od", "oh", "ow", "cout", "k", "stride", "pad", "dil")]


# name -> argument type codes ('g' = const b200seg_conv_geom*), in header order
SIGNATURES = {
    "b200seg_conv3d_uses_tensor_cores": "g",
    "b200seg_ncdhw_f32_to_ndhwc_bf16": "ppiilp",
    "b200seg_ndhwc_bf16_to_ncdhw_f32": "ppiilp",
    "b200seg_ncdhw_f32_to_ndhwc_bf16_pitched": "pplii" + "l" + "p",
    "b200seg_ndhwc_bf16_to_ncdhw_f32_pitched": "plpii" + "l" + "p",
    "b200seg_pack_conv_weight": "ppiiiiiip",
    "b200seg_pack_conv_weight_padded": "ppiiiiiip",
    "b200seg_pack_weights_batched": "pppiiip",
    "b200seg_unpack_wgrads_batched": "pppiiip",
    "b200seg_pad_channels": "plipilp",
    "b200seg_unpack_conv_wgrad": "ppiiiiiip",
    "b200seg_conv3d_fprop": "gplppplppzp",
    "b200seg_conv3d_fprop_act_supported": "g",
    "b200seg_conv3d_fprop_act": "gplp" + "pp" + "if" + "pl" + "p",
    "b200seg_conv3d_dgrad": "gplpplppzp",
    "b200seg_conv3d_wgrad": "gplplppzp",
    "b200seg_pack_convt_weight": "ppiiip",
    "b200seg_convt_k2s2_fwd": "plpp" + "pl" + "iiiiii" + "p",
    "b200seg_convt_k2s2_dgrad": "plp" + "pl" + "iiiiii" + "p",
    "b200seg_convt_k2s2_wgrad": "plpl" + "p" + "iiiiii" + "p",
    "b200seg_channel_stats": "pllii" + "pp",
    "b200seg_norm_finalize": "pdii" + "pppp" + "ffi" + "pp",
    "b200seg_norm_eval_coef": "ppppp" + "fi" + "pp",
    "b200seg_norm_act_fwd": "plp" + "lii" + "if" + "p" + "pl" + "pl" + "p",
    "b200seg_norm_act_fwd_stats": "plp" + "d" + "pppp" + "ffi" + "p" + "li" + "if" + "p" + "pl" + "pl" + "p",
    "b200seg_norm_act_bwd_reduce": "plpl" + "p" + "lii" + "if" + "p" + "pl" + "ppp" + "p",
    "b200seg_norm_act_bwd_apply": "plpl" + "pp" + "d" + "lii" + "if" + "p" + "pl" + "pl" + "pl" + "p",
    "b200seg_norm_act_bwd_apply_acc": "plpl" + "pp" + "d" + "lii" + "if" + "p" + "pl" + "pl" + "pl" + "pl" + "p",
    "b200seg_maxpool2_fwd": "plplp" + "iiiii" + "p",
    "b200seg_maxpool2_bwd": "plp" + "pl" + "pl" + "iiiii" + "p",
    "b200seg_maxpool2_idx_to_torch": "pp" + "iiiii" + "p",
    "b200seg_upsample2_fwd": "plpl" + "iiiii" + "p",
    "b200seg_upsample2_bwd": "plpl" + "iiiii" + "p",
    "b200seg_add": "plplpl" + "li" + "p",
    "b200seg_dropout": "plpl" + "ll" + "if" + "pQ" + "i" + "p",
    "b200seg_space_to_depth": "plp" + "iiiiiiiiii" + "p",
    "b200seg_depth_to_space": "ppl" + "iiiiiiiiii" + "p",
    "b200seg_convt1_k2s2_fwd": "pppp" + "liii" + "p",
    "b200seg_convt1_k2s2_bwd": "ppppp" + "liii" + "p",
    "b200seg_reverse_gate_fwd": "plp" + "pl" + "li" + "p",
    "b200seg_reverse_gate_bwd": "plplp" + "plp" + "li" + "p",
    "b200seg_channel_blend_fwd": "plp" + "plp" + "pl" + "lii" + "p",
    "b200seg_channel_blend_bwd_reduce": "plplpl" + "pp" + "lii" + "p",
    "b200seg_channel_blend_bwd_apply": "pl" + "ppp" + "plpl" + "lii" + "p",
    "b200seg_f32_sigmoid_fwd": "ppl" + "p",
    "b200seg_f32_sigmoid_bwd": "pppl" + "p",
    "b200seg_dropout2": "plpl" + "ll" + "iif" + "pQQ" + "i" + "p",
    "b200seg_classmap_up2_add": "ppp" + "liii" + "p",
    "b200seg_classmap_down2_sum": "pp" + "liii" + "p",
    "b200seg_head_conv1x1_fwd": "plppp" + "ilii" + "p",
    "b200seg_head_conv1x1_bwd": "ppl" + "p" + "pl" + "pp" + "ilii" + "p",
    "b200seg_argmax_labels": "pp" + "ili" + "p",
    "b200seg_loss_reduce": "pp" + "ilii" + "ppp",
    "b200seg_loss_grad": "pp" + "ili" + "p" + "ffff" + "pp" + "pp",
    "b200seg_dice_sums": "ppp" + "iil" + "if" + "pp",
    "b200seg_dice_grad": "ppp" + "iil" + "if" + "ppp" + "pp",
    "b200seg_seg_counts": "ppl" + "pp",
    "b200seg_mask_edge_points": "piii" + "fff" + "pQpp",
    "b200seg_min_distances": "plplpp",
    "b200seg_pad3d_fwd": "plpl" + "iiiii" + "ii" + "p",
    "b200seg_pad3d_bwd": "plpl" + "iiiii" + "ii" + "p",
    "b200seg_volume_stats": "plpp",
    "b200seg_znorm_finalize": "plpp",
    "b200seg_crop_patch": "pi" + "iiii" + "iii" + "iii" + "ppp",
    "b200seg_window_accumulate_crop": "ppp" + "l" + "iiiiiii" + "p" + "iii" + "p",
    "b200seg_window_keys_to_labels": "pp" + "l" + "p",
    "b200seg_window_accumulate_average": "pp" + "iiiii" + "pp" + "iii" + "p",
    "b200seg_window_finalize": "pp" + "il" + "pp",
    "b200seg_p2p_alloc": "pp",
    "b200seg_p2p_alloc_bytes": "zpp",
    "b200seg_p2p_grad_allreduce": "pppp" + "iii" + "pp",
    "b200seg_p2p_open": "pp",
    "b200seg_p2p_close": "pi",
    "b200seg_p2p_allreduce": "pi" + "p" + "ii" + "p" + "ii" + "d" + "pppp" + "ff" + "i" + "p" + "p",
    "b200seg_adam_step": "pppp" + "l" + "fffff" + "i" + "f" + "p",
    "b200seg_adam_step_dev": "pppp" + "l" + "pp" + "p",
    "b200seg_adam_step_fused": "ppppp" + "pp" + "iii" + "pp" + "i" + "p",
}
STRING_FUNCS = ("b200seg_version", "b200seg_last_error")
SIZE_FUNCS = {"b200seg_conv3d_workspace_bytes": "g", "b200seg_conv3d_ws_bytes": "gi", "b200seg_p2p_mailbox_bytes": ""}
INT64_FUNCS = {"b200seg_umma_launch_count": ""}


class B200SegError(RuntimeError):
    pass


_lib = None


def load():
    """Load the extension; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200SegError("libb200seg.so is missing at %s: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)

    def types(codes):
        return [ctypes.POINTER(ConvGeom) if c == "g" else _CODES[c] for c in codes]

    for name, codes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = types(codes)
        fn.restype = ctypes.c_int
    for name in STRING_FUNCS:
        getattr(lib, name).restype = ctypes.c_char_p
        getattr(lib, name).argtypes = []
    for name, codes in SIZE_FUNCS.items():
        getattr(lib, name).restype = ctypes.c_size_t
        getattr(lib, name).argtypes = types(codes)
    for name, codes in INT64_FUNCS.items():
        getattr(lib, name).restype = ctypes.c_int64
        getattr(lib, name).argtypes = types(codes)
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an ABI function; raise B200SegError with the library's message on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise B200SegError("%s failed (%d): %s" % (name, rc, lib.b200seg_last_error().decode()))


def exported_symbols():
    return list(SIGNATURES) + list(STRING_FUNCS) + list(SIZE_FUNCS) + list(INT64_FUNCS)
