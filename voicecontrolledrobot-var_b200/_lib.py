"""ctypes binding of libvar_b200.so (C ABI declared in include/var_b200.h).

The library is the product: if it is missing, or a call fails, this module raises --
there is no CPU or PyTorch fallback anywhere in the package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvar_b200.so")

VAR_ERRORS = {-1: "VAR_ERR_ARG", -2: "VAR_ERR_CUDA", -3: "VAR_ERR_UNSUPPORTED", -4: "VAR_ERR_WORKSPACE"}


class VarB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C voicecontrolledrobot-var_b200/csrc`). The VAR hot path has no CPU fallback.")
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

_p, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes); mirrors include/var_b200.h one to one
PROTOTYPES = {
    "var_version": (_i, []),
    "var_last_error": (C.c_char_p, []),
    "var_mfcc_plan_create": (_i, [_i, _i, _i, _i, _i, C.POINTER(_p)]),
    "var_mfcc_plan_destroy": (_i, [_p]),
    "var_mfcc_num_frames": (_i, [_p, _i]),
    "var_mfcc_fwd": (_i, [_p, _p, _p, _p, _i, _i, _p, _p]),
    "var_sampler_seed": (_i, [_p, C.c_uint64, _p]),
    "var_sampler_epoch": (_i, [_p, _i, _p, _p]),
    "var_sampler_batch": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "var_sampler_batch_tasks": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p,
                                     _p, _p]),
    "var_sampler_set_state": (_i, [_p, _p, _i, _p]),
    "var_host_gather_rows": (_i, [_p, _i64, _p, _i, _p, _i]),
    "var_host_gather_clips": (_i64, [_p, _p, _p, _i, _p, _p, _i]),
    "var_net_create": (_i, [_i, _i, _i, C.POINTER(_p)]),
    "var_net_destroy": (_i, [_p]),
    "var_net_param_floats": (_i64, [_p]),
    "var_net_num_tensors": (_i, [_p]),
    "var_net_tensor_info": (_i, [_p, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i64),
                                 C.POINTER(_i64)]),
    "var_net_bind": (_i, [_p, _p, _p, _p]),
    "var_net_load_tensor": (_i, [_p, _i, _p, _p]),
    "var_net_store_tensor": (_i, [_p, _i, _i, _p, _p]),
    "var_net_refresh_mma": (_i, [_p, _p]),
    "var_net_workspace_bytes": (_i64, [_p, _i, _i, _i]),
    "var_net_set_overlap": (_i, [_p, _i]),
    "var_net_grad_bucket": (_i, [_p, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "var_net_set_bucket_event": (_i, [_p, _i, _p]),
    "var_net_raw_dims": (_i, [_p, C.POINTER(_i), C.POINTER(_i)]),
    "var_net_forward": (_i, [_p, _p, _i, _i, _p, _i, _p, _i64, _i, _p, _p, _p, _p, _p]),
    "var_net_backward": (_i, [_p, _p, _p, _p, _i64, _p]),
    "var_net_triplet_step": (_i, [_p, _p, _i, _p, _i, _f, _f, _p, _i64, _p, _p, _p]),
    "var_net_reward": (_i, [_p, _p, _i, _p, _p, _p, _i, _p, _i64, _p, _p, _p, _p, _p]),
    "var_reward_normalize": (_i, [_p, _p, _i, _p, _p, C.c_double, C.c_double, C.c_double, _i, _p, _p, _p]),
    "var_adam_step": (_i, [_p, _p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i64, _f, _p]),
    "var_pack_weight": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "var_unpack_weight": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "var_conv2d_fwd": (_i, [_p, _i, _p, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _p]),
    "var_conv2d_dgrad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "var_conv2d_wgrad": (_i, [_p, _i, _p, _f, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "var_cvt_f16": (_i, [_p, _p, _i64, _p]),
    "var_grad_to_f16_scaled": (_i, [_p, _p, _i64, _p, _p, _p]),
    "var_conv2d_fwd_h16": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _p]),
    "var_conv2d_dgrad_h16": (_i, [_p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "var_conv2d_wgrad_h16": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "var_linear_h16": (_i, [_p, _i64, _p, _p, _p, _i64, _p, _i, _i64, _p, _i, _i, _i, _i, _p]),
    "var_linear_wgrad_h16": (_i, [_p, _i64, _p, _i64, _p, _i, _p, _i, _i, _i, _p]),
    "var_maxpool2x2_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "var_maxpool2x2_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "var_launch_count": (C.c_longlong, []),
    "var_launch_count_add": (C.c_int, [C.c_longlong]),
    "var_prof_begin": (_i, []),
    "var_prof_end": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong), _i]),
    "var_prof_num_tags": (_i, []),
    "var_h16_flags": (_i, []),
    "var_triplet_fwd_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p,
                                 _p, _p, _p, _p, _p]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


PROF_TAGS = ["gemm_fwd", "gemm_dgrad", "gemm_scalar", "gru_step", "wgrad", "colsum", "mfcc", "tail", "pool", "adam",
             "gru_cell_bwd", "sampler", "misc", "gemm_fwd16", "gemm_dgrad16", "wgrad16", "gru_bwd"]


def prof_end():
    """-> {tag: (ms, flops, launches)} of the launches since var_prof_begin()."""
    n = lib.var_prof_num_tags()
    ms, fl, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_longlong * n)()
    check(lib.var_prof_end(ms, fl, cnt, n), "var_prof_end")
    return {PROF_TAGS[i]: (ms[i], fl[i], cnt[i]) for i in range(n) if cnt[i]}


def last_error():
    msg = lib.var_last_error()
    return msg.decode() if msg else ""


def check(rc, what):
    if rc is not None and rc < 0:
        raise VarB200Error(f"{what} failed: {VAR_ERRORS.get(rc, rc)} {last_error()}")
    return rc


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
