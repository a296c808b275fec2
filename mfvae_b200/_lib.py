"""ctypes binding of libmfvae_b200.so (the C ABI declared in include/mfvae.h).

There is no CPU or eager fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmfvae_b200.so")
MAX_HIDDEN = 8

PREC_FP32, PREC_BF16 = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2
FUSE_AUTO, FUSE_NONE, FUSE_ENCODER, FUSE_LOSS, FUSE_NOFOLD_IDX, FUSE_NOFOLD_ACT = 0, 1, 2, 4, 8, 16
LOSS_DEFAULT, LOSS_HUBER, LOSS_MSE, LOSS_JOINT_MSE = 0, 1, 2, 3
(T_IDX_EMB, T_ENC_W, T_ENC_B, T_ACT_TABLE, T_SDEC_W, T_SDEC_B, T_RDEC_W, T_RDEC_B, T_RLIN_W, T_RLIN_B, T_ACTENC_W,
 T_ACTENC_B) = range(12)


class MfvaeConfig(C.Structure):
    _fields_ = [("n_agents", C.c_int32), ("idx_features", C.c_int32), ("latent", C.c_int32),
                ("act_features", C.c_int32), ("n_enc_hidden", C.c_int32), ("enc_hidden", C.c_int32 * MAX_HIDDEN),
                ("n_dec_hidden", C.c_int32), ("dec_hidden", C.c_int32 * MAX_HIDDEN),
                ("obs_dim", C.POINTER(C.c_int32)), ("n_act", C.POINTER(C.c_int32)),
                ("kl_weight", C.c_float), ("r_weight", C.c_float), ("huber", C.c_int32),
                ("precision", C.c_int32), ("engine", C.c_int32), ("optimize_encoders", C.c_int32),
                ("fusion", C.c_int32), ("continuous_act", C.c_int32), ("act_hidden", C.c_int32)]


class MfvaeTensorInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("agent", C.c_int32), ("layer", C.c_int32), ("rows", C.c_int32),
                ("cols", C.c_int32), ("ld", C.c_int32), ("offset", C.c_int64)]


class MfvaeArenas(C.Structure):
    _fields_ = [("d_param", C.c_void_p), ("d_grad", C.c_void_p), ("d_m", C.c_void_p), ("d_v", C.c_void_p),
                ("d_shadow_bf16", C.c_void_p)]


class MfvaeBatch(C.Structure):
    _fields_ = [("d_obs", C.c_void_p), ("d_act", C.c_void_p), ("d_next", C.c_void_p), ("d_rew", C.c_void_p),
                ("d_idx", C.c_void_p), ("d_eps", C.c_void_p), ("batch", C.c_int32), ("sample0", C.c_int64),
                ("batch_global", C.c_int64), ("seed", C.c_uint64), ("step", C.c_uint64), ("obs_bf16", C.c_int32), ("next_bf16", C.c_int32)]


class MfvaeGemmTiming(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("groups", C.c_int32), ("kind", C.c_int32),
                ("ms", C.c_float)]


class MfvaeOutputs(C.Structure):
    _fields_ = [("d_recon_s", C.c_void_p), ("recon_s_ld", C.c_int32), ("d_recon_r", C.c_void_p),
                ("recon_r_ld", C.c_int32), ("d_latent", C.c_void_p), ("d_losses", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/mfvae.h declares
_vp, _i32, _i64, _u64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
SIGNATURES = {
    "mfvae_last_error": (C.c_char_p, []),
    "mfvae_version": (C.c_int, []),
    "mfvae_create": (C.c_int, [C.POINTER(MfvaeConfig), C.c_int, C.POINTER(_vp)]),
    "mfvae_destroy": (C.c_int, [_vp]),
    "mfvae_arena_elems": (_i64, [_vp]),
    "mfvae_optimized_elems": (_i64, [_vp]),
    "mfvae_tensor_count": (_i32, [_vp]),
    "mfvae_tensor_table": (C.c_int, [_vp, C.POINTER(MfvaeTensorInfo), _i32]),
    "mfvae_bind_arenas": (C.c_int, [_vp, C.POINTER(MfvaeArenas)]),
    "mfvae_refresh_shadow": (C.c_int, [_vp, _vp]),
    "mfvae_refresh_shadow_range": (C.c_int, [_vp, _i64, _i64, _vp]),
    "mfvae_workspace_bytes": (_i64, [_vp, _i32]),
    "mfvae_bind_workspace": (C.c_int, [_vp, _vp, _i64, _i32]),
    "mfvae_forward": (C.c_int, [_vp, C.POINTER(MfvaeBatch), C.POINTER(MfvaeOutputs), _vp]),
    "mfvae_loss": (C.c_int, [_vp, C.POINTER(MfvaeBatch), _i32, _vp]),
    "mfvae_set_loss_weights": (C.c_int, [_vp, _f, _f]),
    "mfvae_set_loss_weights3": (C.c_int, [_vp, _f, _f, _f]),
    "mfvae_backward": (C.c_int, [_vp, C.POINTER(MfvaeBatch), _vp]),
    "mfvae_backward_ext": (C.c_int, [_vp, C.POINTER(MfvaeBatch), _vp, _i64, _vp, _i64, _vp, _vp]),
    "mfvae_adam_step": (C.c_int, [_vp, _f, _f, _f, _f, _i64, _vp]),
    "mfvae_adam_step_overlapped": (C.c_int, [_vp, _f, _f, _f, _f, _i64, _vp]),
    "mfvae_adam_range": (C.c_int, [_vp, _i64, _i64, _f, _f, _f, _f, _i64, _vp]),
    "mfvae_wait_decoder_reads": (C.c_int, [_vp, _vp]),
    "mfvae_bucket_read_wait": (C.c_int, [_vp, _i32, _vp]),
    "mfvae_set_sm_reserve": (C.c_int, [_vp, _i32]),
    "mfvae_fwd_bwd": (C.c_int, [_vp, C.POINTER(MfvaeBatch), C.POINTER(MfvaeOutputs), _vp]),
    "mfvae_train_step": (C.c_int, [_vp, C.POINTER(MfvaeBatch), _f, _f, _f, _f, _i64, _i32, C.POINTER(MfvaeOutputs), _vp]),
    "mfvae_launch_count": (C.c_uint64, []),
    "mfvae_profile_enable": (C.c_int, [_vp, _i32]),
    "mfvae_profile_read": (_i32, [_vp, C.POINTER(MfvaeGemmTiming), _i32]),
    "mfvae_bucket_count": (_i32, [_vp]),
    "mfvae_bucket": (C.c_int, [_vp, _i32, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_vp)]),
    "mfvae_bucket_wait": (C.c_int, [_vp, _i32, _vp]),
    "mfvae_loss_wait": (C.c_int, [_vp, _vp]),
    "mfvae_comm_bind": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _i32]),
    "mfvae_opt_join": (C.c_int, [_vp, _vp]),
    "mfvae_comm_window_bytes": (_i64, [_vp, _i32]),
    "mfvae_allreduce_grads": (C.c_int, [_vp, _i64, _i64, _i32, _f, _f, _f, _f, _i64, _vp]),
    "mfvae_allreduce_losses": (C.c_int, [_vp, _vp]),
    "mfvae_reparam_kl": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _u64, _u64, _i64, _i64, _vp, _vp, _vp]),
    "mfvae_recon_loss": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _i32, _i64, _i32, _i32, _f, _i64, _vp, _vp, _vp]),
    "mfvae_adam_flat": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i64, _vp]),
    "mfvae_philox_normal": (C.c_int, [_vp, _i64, _i32, _u64, _u64, _i64, _vp]),
    "mfvae_gemm": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64,
                             _vp, _i64, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _i64, _i32, _vp]),
    "mfvae_ring_create": (C.c_int, [_i32, _i32, _i32, _i64, _vp, C.POINTER(_vp)]),
    "mfvae_ring_destroy": (C.c_int, [_vp]),
    "mfvae_ring_row_floats": (_i64, [_i32, _i32, _i32]),
    "mfvae_ring_size": (_i64, [_vp]),
    "mfvae_ring_add": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "mfvae_ring_sample": (C.c_int, [_vp, _i64, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m mfvae_b200.build` "
                               "(there is no CPU / eager fallback for this path)")
        import torch  # noqa: F401  the library links the CUDA runtime dynamically; torch has libcudart.so.12 loaded already
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError = header / library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("mfvae_b200: " + lib().mfvae_last_error().decode(errors="replace"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
