"""
lib.py — ctypes binding of libt4k.so (include/t4k.h).  The CUDA library IS the product: if it
is missing or fails to load this module raises — there is no Python/CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libt4k.so")

# enums of include/t4k.h (same numbering as the reference's math_op / t4_layer / t4_loss)
(ABS, NEG, EXP, LN, LOG, TANH, RELU, SIGM, SQRT, RCP, SAT, IDEN, FILL, GFILL, SCALE, POW,
 ADD, SUB, MUL, DIV, MOD, MAX, MIN) = range(23)
(L_NONE, L_CONV, L_LINEAR, L_FLATTEN, L_RELU, L_TANH, L_SIGMOID, L_SELU, L_LEAKYRL, L_ELU,
 L_DROPOUT, L_SOFTMAX, L_LOGSMAX, L_AVGPOOL, L_MAXPOOL, L_MINPOOL, L_BATCHNM, L_USAMPLE,
 L_DCONV) = range(19)
LOSS_MSE, LOSS_BCE, LOSS_CE, LOSS_NLL = range(4)
UNIFORM, NORMAL = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TC, GEMM_TCF, GEMM_TC_BF16X3, GEMM_MMA, GEMM_TL = 0, 1, 2, 3, 4, 5, 6
EINVAL, ENOSUP, ENOMEM = -1, -2, -3
COMM_HANDLE_BYTES = 64

_p, _i, _f, _l, _u64 = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint64

# name -> (restype, argtypes); every symbol include/t4k.h declares
PROTOTYPES = {
    "t4k_version": (_i, []),
    "t4k_strerror": (C.c_char_p, [_i]),
    "t4k_device_count": (_i, []),
    "t4k_sm_count": (_i, []),
    "t4k_sync": (_i, [_p]),
    "t4k_launch_count": (C.c_long, []),
    "t4k_set_pdl": (_i, [_i]),
    "t4k_set_workspace_bank": (_i, [_i]),
    "t4k_map": (_i, [_i, _p, _f, _l, _p]),
    "t4k_ts_op": (_i, [_i, _p, _f, _p, _l, _p]),
    "t4k_tt_op": (_i, [_i, _p, _p, _p, _l, _i, _i, _p]),
    "t4k_copy": (_i, [_p, _p, _l, _p]),
    "t4k_transpose": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "t4k_identity": (_i, [_p, _i, _i, _i, _i, _p]),
    "t4k_sum": (_i, [_p, _l, _p, _p]),
    "t4k_nvar": (_i, [_p, _f, _l, _p, _p]),
    "t4k_minmax": (_i, [_p, _l, _i, _p, _p]),
    "t4k_avg_std": (_i, [_p, _l, _p, _p]),
    "t4k_dot": (_i, [_p, _p, _p, _f, _f, _i, _i, _i, _i, _p]),
    "t4k_loss": (_i, [_i, _p, _p, _l, _i, _p, _p]),
    "t4k_nan_inf": (_i, [_p, _l, _p, _p]),
    "t4k_gemm": (_i, [_p, _p, _p, _f, _f, _i, _i, _i, _i, _i, _i, _i, _l, _l, _l, _p]),
    "t4k_gemm_ex": (_i, [_i, _p, _p, _p, _f, _f, _i, _i, _i, _i, _i, _i, _i, _l, _l, _l, _p]),
    "t4k_set_gemm_tl": (_i, [_i, _i]),
    "t4k_gemm_tl_trace": (_i, [_p]),
    "t4k_bias": (_i, [_p, _p, _i, _i, _p]),
    "t4k_linear_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "t4k_activate_fwd": (_i, [_i, _p, _p, _p, _f, _l, _p]),
    "t4k_softmax_fwd": (_i, [_p, _p, _i, _i, _p]),
    "t4k_logsoftmax_fwd": (_i, [_p, _p, _i, _i, _p]),
    "t4k_conv2d_fwd": (_i, [_p, _p, _p, _p] + [_i] * 10 + [_p]),
    "t4k_set_conv_engine": (_i, [_i]),
    "t4k_pool_fwd": (_i, [_i, _p, _p] + [_i] * 7 + [_p]),
    "t4k_batchnorm_fwd": (_i, [_p] * 6 + [_i] * 3 + [_p]),
    "t4k_dbias": (_i, [_p, _p, _i, _i, _p]),
    "t4k_linear_bwd": (_i, [_p] * 6 + [_i] * 4 + [_p]),
    "t4k_linear_bwd_ex": (_i, [_p] * 6 + [_i] * 5 + [_p]),
    "t4k_dconv2d_fwd": (_i, [_p] * 4 + [_i] * 10 + [_p]),
    "t4k_dconv2d_bwd": (_i, [_p] * 6 + [_i] * 11 + [_p]),
    "t4k_set_carveout": (_i, [_i]),
    "t4k_dropout_fwd": (_i, [_p, _p, _p, _f, _l, _l, _l, _p]),
    "t4k_rand_sharded": (_i, [_p, _l, _l, _l, _i, _f, _f, _p]),
    "t4k_batchnorm_fwd_dp": (_i, [_p] * 7 + [_i] * 4 + [_p]),
    "t4k_batchnorm_bwd_dp": (_i, [_p] * 8 + [_i] * 5 + [_p]),
    "t4k_head_train_scratch_floats": (_l, [_i] * 5),
    "t4k_linear_act_head_train": (_i, [_i] + [_p] * 6 + [_f] + [_p] * 8 + [_i] * 4 + [_p]),
    "t4k_head_grad_finish": (_i, [_p, _i, _i, _i, _p, _p, _p, _p]),
    "t4k_linear_bwd_pair": (_i, [_p] * 5 + [_i] * 3 + [_p]),
    "t4k_linear_bwd_from_head": (_i, [_p] * 8 + [_i] * 4 + [_p]),
    "t4k_linear_dx_from_head": (_i, [_p] * 6 + [_i] * 4 + [_p]),
    "t4k_linear_bwd_act": (_i, [_p] * 8 + [_i] * 5 + [_p]),
    "t4k_linear_act_fwd": (_i, [_i] + [_p] * 6 + [_f] + [_i] * 3 + [_p]),
    "t4k_mlp_head_fwd": (_i, [_p] * 5 + [_i] * 3 + [_p]),
    "t4k_mlp_head_fwd_dup": (_i, [_p] * 6 + [_i] * 3 + [_p]),
    "t4k_linear_act_head_fwd": (_i, [_i] + [_p] * 6 + [_f] + [_p] * 5 + [_i] * 4 + [_p]),
    "t4k_mlp_head_bwd": (_i, [_p] * 10 + [_i] * 4 + [_p]),
    "t4k_activate_bwd": (_i, [_p, _p, _p, _l, _p]),
    "t4k_conv2d_bwd": (_i, [_p] * 6 + [_i] * 11 + [_p]),
    "t4k_pool_bwd": (_i, [_i, _p, _p] + [_i] * 7 + [_p]),
    "t4k_batchnorm_bwd": (_i, [_p] * 7 + [_i] * 4 + [_p]),
    "t4k_conv_pool_relu_fwd": (_i, [_p] * 9 + [_i] * 10 + [_p]),
    "t4k_conv_pool_relu_fwd_feed": (_i, [_p, _p, _i, _f, _f, _p, _p, _i] + [_p] * 9 + [_i] * 10 + [_p]),
    "t4k_conv_pool_relu_bwd": (_i, [_p] * 10 + [_i] * 11 + [_p]),
    "t4k_sgd": (_i, [_p, _p, _p, _i, _f, _f, _l, _p]),
    "t4k_adam": (_i, [_p, _p, _p, _p, _f, _f, _f, _l, _p]),
    "t4k_adamw": (_i, [_p, _p, _p, _p, _f, _f, _f, _f, _l, _p]),
    "t4k_optim_multi_range": (_i, [_i, _p, _p, _p, _p, _p, _i, _l, _l, _f, _f, _f, _f, _p]),
    "t4k_conv_pool_relu_bwd_opt": (_i, [_p] * 10 + [_i] * 11 + [_p, _p]),
    "t4k_optim_multi": (_i, [_i, _p, _p, _p, _p, _p, _i, _l, _f, _f, _f, _f, _p]),
    "t4k_comm_create": (_i, [_i, _i, _l, C.POINTER(_p), _p]),
    "t4k_comm_connect": (_i, [_p, _p]),
    "t4k_comm_connect_local": (_i, [_p, C.POINTER(_p)]),
    "t4k_comm_destroy": (_i, [_p]),
    "t4k_comm_status": (_i, [_p]),
    "t4k_comm_poll": (_i, [_p]),
    "t4k_comm_capacity": (_l, [_p]),
    "t4k_shard_info": (_i, [_l, _i, _i, C.POINTER(_l), C.POINTER(_l)]),
    "t4k_allreduce_sum": (_i, [_p, _p, _l, _p]),
    "t4k_optim_multi_dp": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _l, _f, _f, _f, _f, _p, _i, _l, _p]),
    "t4k_optim_multi_dp_range": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _l, _l, _l, _f, _f, _f, _f, _p, _i, _l, _p]),
    "t4k_conv_pool_relu_bwd_mid_event": (_i, [_p]),
    "t4k_comm_chunk_floats": (_l, [_p]),
    "t4k_comm_scalar_mirror": (_i, [_p, _p]),
    "t4k_dp_push_owner": (_l, [_p, _p, _l, _l, _p]),
    "t4k_optim_multi_dp_rs": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _l, _l, _f, _f, _f, _f, _i, _p]),
    "t4k_dp_push_dma": (_l, [_p, _p, _l, _l, C.c_uint32, _p]),
    "t4k_dp_push": (_l, [_p, _p, _l, _l, _p]),
    "t4k_rand_seed": (_i, [_u64]),
    "t4k_rand": (_i, [_p, _l, _i, _f, _f, _p]),
    "t4k_rand_at": (_i, [_p, _l, _i, _f, _f, _u64, _u64, _p]),
    "t4k_rand_tick": (_i, [_p]),
    "t4k_dataset_load": (_i, [_p, _p, _l, _f, _f, _p, _p, _i, _p, _i, _p]),
    "t4k_onehot": (_i, [_p, _p, _i, _i, _p]),
    "t4k_hit": (_i, [_p, _p, _i, _i, _p, _p]),
}


class T4KError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen libt4k.so and attach prototypes.  Raises if the CUDA extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise T4KError("libt4k.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C tensorforth_b200/csrc` — there is no CPU fallback" % SO_PATH)
    L = C.CDLL(SO_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(L, name)            # AttributeError if the ABI and the header drift apart
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        raise T4KError("%s failed: rc=%d (%s)" % (what or "t4k call", rc, load().t4k_strerror(rc).decode()))
    return rc
