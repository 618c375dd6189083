"""ctypes binding of libbsed.so (the C ABI declared in include/bsed.h).

There is no fallback: if the shared library is missing, or no B200 is usable, calls raise.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbsed.so")

BSED_F_TRAIN = 1
BSED_F_SAVE = 2
MAX_CNN_LAYERS = 8

# every symbol include/bsed.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = [
    "bsed_version", "bsed_last_error", "bsed_create", "bsed_destroy", "bsed_frontend_n_frames",
    "bsed_melspec", "bsed_amp_to_db_workspace_bytes", "bsed_amp_to_db", "bsed_median_decode",
    "bsed_plan_create", "bsed_plan_destroy", "bsed_plan_param_count", "bsed_plan_bn_buffer_count",
    "bsed_plan_out_frames", "bsed_plan_param_offsets", "bsed_plan_workspace_bytes", "bsed_crnn_forward",
    "bsed_crnn_backward", "bsed_predictor_param_count", "bsed_predictor_param_offsets",
    "bsed_predictor_ldl", "bsed_predictor_workspace_bytes", "bsed_predictor_forward", "bsed_predictor_backward",
    "bsed_plan_debug_tensor", "bsed_mt_loss", "bsed_opt_ema_step", "bsed_ema_buffers", "bsed_gemm_nn",
    "bsed_gemm_tn", "bsed_conv3x3", "bsed_launch_count", "bsed_profile_begin", "bsed_profile_end",
    "bsed_conv3x3_tc", "bsed_gemm_nt_tc", "bsed_conv3x3_wgrad", "bsed_conv3x3_wgrad_workspace_bytes",
    "bsed_plan_set_precision", "bsed_plan_get_precision",
    "bsed_disc_param_count", "bsed_disc_bn_buffer_count", "bsed_disc_workspace_bytes", "bsed_disc_forward",
    "bsed_disc_backward", "bsed_disc_bce", "bsed_disc_set_precision", "bsed_loss_terms", "bsed_roll_clips",
    "bsed_ipc_export", "bsed_ipc_open", "bsed_ipc_close", "bsed_dp_opt_ema_step",
    "bsed_im2col_nhwc", "bsed_add_relu", "bsed_maxpool_nhwc", "bsed_avgpool_nhwc", "bsed_sigmoid_rows",
    "bsed_bn_rows_workspace_bytes", "bsed_bn_rows_train", "bsed_bn_rows_backward", "bsed_col2im_nhwc",
    "bsed_maxpool_nhwc_backward", "bsed_avgpool_nhwc_backward", "bsed_sigmoid_rows_backward", "bsed_gemm_tn_tc",
    "bsed_logmel", "bsed_conv3x3_tc3", "bsed_gemm_nt_tc3", "bsed_scale_f32", "bsed_step_state_advance",
    "bsed_set_step_state",
]
PRECISIONS = {"fp32": 0, "tf32": 1, "tf32x3": 2}


class CrnnCfg(C.Structure):
    _fields_ = [("n_frames", C.c_int), ("n_mels", C.c_int), ("n_cnn", C.c_int),
                ("filters", C.c_int * MAX_CNN_LAYERS), ("pool_t", C.c_int * MAX_CNN_LAYERS),
                ("pool_f", C.c_int * MAX_CNN_LAYERS), ("rnn_hidden", C.c_int), ("rnn_layers", C.c_int),
                ("n_class", C.c_int), ("dropout", C.c_float), ("bn_eps", C.c_float),
                ("bn_momentum", C.c_float), ("fpn", C.c_int)]


class Group(C.Structure):
    _fields_ = [("params", C.c_void_p), ("bn_buffers", C.c_void_p), ("num_batches_tracked", C.c_void_p),
                ("first_clip", C.c_int), ("n_clips", C.c_int)]


class OptCfg(C.Structure):
    _fields_ = [("kind", C.c_int), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("weight_decay", C.c_float), ("momentum", C.c_float),
                ("grad_scale", C.c_float), ("ema_alpha", C.c_float), ("step", C.c_int64),
                ("ema_step", C.c_int64)]


class LossTerm(C.Structure):
    _fields_ = [("kind", C.c_int), ("pred_first", C.c_int), ("n_clips", C.c_int), ("ref", C.c_void_p),
                ("roll", C.c_void_p), ("ref_is_strong", C.c_int), ("weight", C.c_float), ("grad_weight", C.c_float),
                ("slot", C.c_int)]


class StepState(C.Structure):
    """bsed_step_state (include/bsed.h): the device-resident per-iteration scalars of a captured training step."""
    _fields_ = [("global_step", C.c_int64), ("opt_step", C.c_int64), ("dp_epoch", C.c_int64), ("keys", C.c_uint32 * 16),
                ("lr", C.c_float), ("cons_w", C.c_float), ("step_size", C.c_float), ("bc2_sqrt", C.c_float),
                ("ema_a", C.c_float), ("ema_b", C.c_float), ("first_step", C.c_int32), ("pad", C.c_int32)]


class StepCfg(C.Structure):
    _fields_ = [("dropout_seed", C.c_uint64), ("key_mul", C.c_int64), ("key_add", C.c_int64), ("beta1", C.c_float),
                ("beta2", C.c_float), ("ema_alpha", C.c_float), ("max_consistency_cost", C.c_float),
                ("rampup_length", C.c_int64)]


LOSS_BCE_STRONG, LOSS_BCE_WEAK, LOSS_MSE_STRONG, LOSS_MSE_WEAK = 0, 1, 2, 3

_lib = None
_lock = threading.Lock()


def load():
    """dlopen libbsed.so and declare the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, f32, u32, u64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32, C.c_uint64, C.c_size_t
        P = C.POINTER

        def proto(name, res, *args):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = list(args)

        proto("bsed_version", i32)
        proto("bsed_last_error", C.c_char_p)
        proto("bsed_create", i32, i32, P(vp))
        proto("bsed_destroy", i32, vp)
        proto("bsed_frontend_n_frames", i32, i32)
        proto("bsed_melspec", i32, vp, vp, i32, i32, vp, vp)
        proto("bsed_amp_to_db_workspace_bytes", sz, i32)
        proto("bsed_amp_to_db", i32, vp, vp, vp, f32, i32, i32, i32, vp, vp, vp, vp, sz, vp)
        proto("bsed_logmel", i32, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp)
        proto("bsed_median_decode", i32, vp, vp, i32, i32, i32, f32, i32, vp, i32, vp, vp)
        proto("bsed_plan_create", i32, vp, P(CrnnCfg), i32, P(vp))
        proto("bsed_plan_destroy", i32, vp)
        proto("bsed_plan_set_precision", i32, vp, i32)
        proto("bsed_plan_get_precision", i32, vp)
        proto("bsed_plan_param_count", i64, vp)
        proto("bsed_plan_bn_buffer_count", i64, vp)
        proto("bsed_plan_out_frames", i32, vp)
        proto("bsed_plan_param_offsets", i32, vp, P(i64), i32)
        proto("bsed_plan_workspace_bytes", sz, vp)
        proto("bsed_crnn_forward", i32, vp, P(Group), i32, vp, i32, i32, u64, u64, vp, vp, sz, vp)
        proto("bsed_crnn_backward", i32, vp, u32, vp, vp, i32, vp, sz, vp)
        proto("bsed_predictor_param_count", i64, vp)
        proto("bsed_predictor_param_offsets", i32, vp, P(i64), i32)
        proto("bsed_predictor_ldl", i32)
        proto("bsed_predictor_workspace_bytes", sz, vp, i32)
        proto("bsed_predictor_forward", i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, sz, vp)
        proto("bsed_predictor_backward", i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, vp, sz, vp)
        proto("bsed_plan_debug_tensor", i32, vp, vp, C.c_char_p, P(vp), P(i64))
        proto("bsed_mt_loss", i32, vp, vp, vp, i32, i32, i32, i32, i32, vp, i32, i32, vp, vp, f32, vp, vp, vp, vp)
        proto("bsed_loss_terms", i32, vp, vp, vp, i32, i32, i32, P(LossTerm), i32, vp, i32, vp, vp, vp)
        proto("bsed_roll_clips", i32, vp, vp, vp, vp, vp, i32, i32, i32, vp)
        proto("bsed_ipc_export", i32, vp, vp, C.c_char_p, P(u64))
        proto("bsed_ipc_open", i32, vp, C.c_char_p, u64, P(vp))
        proto("bsed_ipc_close", i32, vp, vp, u64)
        proto("bsed_dp_opt_ema_step", i32, vp, i32, i32, P(vp), P(vp), P(vp), P(vp), i64, vp, vp, i64, P(OptCfg), vp)
        proto("bsed_im2col_nhwc", i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp)
        proto("bsed_add_relu", i32, vp, vp, vp, i64, vp)
        proto("bsed_maxpool_nhwc", i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp)
        proto("bsed_avgpool_nhwc", i32, vp, vp, vp, i32, i32, i32, vp)
        proto("bsed_sigmoid_rows", i32, vp, vp, i32, vp, i32, i32, vp)
        proto("bsed_bn_rows_workspace_bytes", sz, i32)
        proto("bsed_bn_rows_train", i32, vp, vp, i64, i32, vp, vp, f32, f32, vp, vp, vp, vp, i32, vp, vp, vp, sz, vp)
        proto("bsed_bn_rows_backward", i32, vp, vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, sz, vp)
        proto("bsed_col2im_nhwc", i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp)
        proto("bsed_maxpool_nhwc_backward", i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp)
        proto("bsed_avgpool_nhwc_backward", i32, vp, vp, vp, i32, i32, i32, vp)
        proto("bsed_sigmoid_rows_backward", i32, vp, vp, vp, vp, i32, i32, i32, vp)
        proto("bsed_opt_ema_step", i32, vp, vp, vp, vp, vp, vp, i64, P(OptCfg), vp)
        proto("bsed_ema_buffers", i32, vp, vp, vp, i64, vp, vp, i32, f32, i64, vp)
        proto("bsed_gemm_nn", i32, vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, vp)
        proto("bsed_gemm_tn", i32, vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, vp)
        proto("bsed_launch_count", u64)
        proto("bsed_profile_begin", i32, i32)
        proto("bsed_profile_end", i32, P(C.c_double), P(C.c_double), P(C.c_double), P(i32))
        proto("bsed_conv3x3_wgrad_workspace_bytes", sz, vp)
        proto("bsed_conv3x3_wgrad", i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, sz, vp)
        proto("bsed_conv3x3_tc", i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp)
        proto("bsed_gemm_tn_tc", i32, vp, vp, i32, vp, i32, vp, i32, i32, i32, i64, vp, sz, vp)
        proto("bsed_gemm_nt_tc", i32, vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, vp)
        proto("bsed_conv3x3", i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp)
        proto("bsed_conv3x3_tc3", i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp)
        proto("bsed_gemm_nt_tc3", i32, vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, vp, vp)
        proto("bsed_disc_set_precision", i32, vp, i32)
        proto("bsed_scale_f32", i32, vp, vp, vp, i64, f32, vp)
        proto("bsed_step_state_advance", i32, vp, vp, P(StepCfg), vp)
        proto("bsed_set_step_state", i32, vp, vp)
        proto("bsed_disc_param_count", i64)
        proto("bsed_disc_bn_buffer_count", i64)
        proto("bsed_disc_workspace_bytes", sz, i32)
        proto("bsed_disc_forward", i32, vp, vp, vp, vp, vp, i32, i32, vp, vp, sz, vp)
        proto("bsed_disc_backward", i32, vp, vp, vp, vp, i32, vp, i32, vp, vp, sz, vp)
        proto("bsed_disc_bce", i32, vp, vp, vp, i32, vp, vp, vp)
        _lib = lib
    return _lib


class BsedError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = load().bsed_last_error()
        raise BsedError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


_handles = {}


def handle(device_index):
    """One library context per (process, device)."""
    lib = load()
    if device_index not in _handles:
        h = C.c_void_p()
        check(lib.bsed_create(int(device_index), C.byref(h)), "bsed_create")
        _handles[device_index] = h
    return _handles[device_index]


def ptr(t):
    """Device pointer of a contiguous torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_contiguous(), "libbsed takes contiguous tensors"
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
