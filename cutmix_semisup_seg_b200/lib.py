"""ctypes binding of libb200seg.so (C ABI: include/b200seg.h).

The product path has no CPU fallback: if the shared library is missing or a call fails this module
raises.  PyTorch is only used for device memory (data_ptr) and the current CUDA stream.
"""
import ctypes
import os

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# B200SEG_LIB: another build of the same library (A/B timing of two builds on one box, tools/gpu_session.sh libab)
LIB_PATH = os.environ.get('B200SEG_LIB') or os.path.join(_PKG_DIR, 'libb200seg.so')

c_int = ctypes.c_int
c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_f32 = ctypes.c_float
c_vp = ctypes.c_void_p
c_sz = ctypes.c_size_t


class B2Error(RuntimeError):
    pass


class ConvParams(ctypes.Structure):
    """struct b2_conv_params (include/b200seg.h)."""
    _fields_ = [
        ('a', c_vp), ('a_lo', c_vp), ('b', c_vp), ('b_lo', c_vp), ('d', c_vp),
        ('n', c_i32), ('ih', c_i32), ('iw', c_i32), ('k', c_i32), ('lda', c_i32),
        ('nb', c_i32), ('tb', c_i32), ('ldb', c_i32),
        ('oh', c_i32), ('ow', c_i32),
        ('fh', c_i32), ('fw', c_i32), ('ldd', c_i32), ('ostride', c_i32), ('ooh', c_i32), ('oow', c_i32),
        ('istride', c_i32),
        ('n_taps', c_i32),
        ('taps', c_vp),
        ('scale', c_vp), ('shift', c_vp),
        ('addend', c_vp), ('ld_add', c_i32),
        ('gate', c_vp), ('ld_gate', c_i32),
        ('scale2', c_vp),
        ('relu', c_i32), ('accumulate', c_i32), ('n_split', c_i32),
        ('max_ctas', c_i32),
        ('stats', c_vp), ('ld_stats', c_i32),
        ('stats_sub', c_vp), ('ld_stats_sub', c_i32),
    ]


class WgradParams(ctypes.Structure):
    """struct b2_wgrad_params (include/b200seg.h)."""
    _fields_ = [
        ('dy', c_vp), ('dy_lo', c_vp), ('x', c_vp), ('x_lo', c_vp), ('dw', c_vp),
        ('n', c_i32), ('oh', c_i32), ('ow', c_i32), ('m', c_i32), ('ldy', c_i32),
        ('ih', c_i32), ('iw', c_i32), ('c', c_i32), ('ldx', c_i32),
        ('istride', c_i32),
        ('n_taps', c_i32), ('taps', c_vp),
        ('tw', c_i32),
        ('accumulate', c_i32), ('n_split', c_i32),
        ('workspace', c_vp), ('workspace_bytes', c_sz),
        ('max_ctas', c_i32),
        ('row_scale', c_vp),
        ('kchunk', c_i32),
    ]


_SIGS = {
    'b2_version': (c_int, []),
    'b2_num_sms': (c_int, []),
    'b2_ema_step': (c_int, [c_vp, c_i64, c_f32, c_f32, c_vp]),
    'b2_ema_step_flat': (c_int, [c_vp, c_vp, c_i64, c_f32, c_f32, c_vp]),
    'b2_opt_ema_step': (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, ctypes.c_double, ctypes.c_double, c_f32, c_f32, c_f32, c_int,
                                c_int, c_f32, c_f32, c_vp]),
    'b2_argmax_confusion': (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_i64, c_vp, c_vp, c_vp]),
    'b2_box_mask_rasterize': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp]),
    'b2_mix': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    'b2_mix_per_sample': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    'b2_ict_conf_mean': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_f32, c_vp]),
    'b2_ict_consistency_fwd_bwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_f32,
                                           c_int, c_vp]),
    'b2_affine_grid_sample': (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b2_aug_consistency_fwd_bwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32,
                                           c_int, c_vp]),
    'b2_col2im': (c_int, [c_vp, c_vp] + [c_int] * 14 + [c_vp]),
    'b2_sample_reduce_blocks': (c_i64, [c_i64]),
    'b2_sample_l2norm': (c_int, [c_vp, c_int, c_i64, c_vp, c_vp, c_vp]),
    'b2_vat_adaptive_radius': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp]),
    'b2_add_scaled_per_sample': (c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_int, c_i64, c_vp]),
    'b2_normalize_to_tensor': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    'b2_u8_to_tensor': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    'b2_crop_flip_normalize': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'b2_crop_flip_u8': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    'b2_geom_u8': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    'b2_colour_jitter': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    'b2_upsample2x_add': (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b2_upsample2x_bwd': (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b2_mul_mask': (c_int, [c_vp, c_int, c_vp, c_f32, c_vp, c_int, c_i64, c_int, c_vp]),
    'b2_avgpool2x2': (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b2_avgpool2x2_bwd': (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b2_scale_channels': (c_int, [c_vp, c_int, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_vp]),
    'b2_consistency_num_partials': (c_i64, [c_int, c_i64]),
    'b2_consistency_fwd_bwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int,
                                       c_f32, c_int, c_vp]),
    'b2_consistency_finalize': (c_int, [c_vp, c_i64, c_i64, c_f32, c_int, c_f32, c_f32, c_vp, c_vp]),
    'b2_ce_num_partials': (c_i64, [c_int, c_i64]),
    'b2_ce_fwd_bwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_i64, c_vp]),
    'b2_ce_finalize': (c_int, [c_vp, c_i64, c_vp, c_vp]),
    'b2_scale_inplace': (c_int, [c_vp, c_i64, c_vp, c_f32, c_vp]),
    'b2_conv_gemm': (c_int, [ctypes.POINTER(ConvParams), c_vp]),
    'b2_conv_stats_rows': (c_i64, [ctypes.POINTER(ConvParams)]),
    'b2_conv_gemm_plan': (c_int, [ctypes.POINTER(ConvParams), c_int, c_vp]),
    'b2_conv_wgrad_workspace': (c_sz, [ctypes.POINTER(WgradParams)]),
    'b2_conv_wgrad': (c_int, [ctypes.POINTER(WgradParams), c_vp]),
    'b2_conv_wgrad_plan_check': (c_int, [ctypes.POINTER(WgradParams), c_vp]),
    'b2_conv_wgrad_plan_balance': (c_int, [ctypes.POINTER(WgradParams), c_vp]),
    'b2_split_tf32': (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    'b2_transpose_w': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'b2_bn_fold_multi': (c_int, [c_vp, c_int, c_int, c_vp]),
    'b2_transpose_w_multi': (c_int, [c_vp, c_int, c_i64, c_vp]),
    'b2_relu_gate': (c_int, [c_vp, c_int, c_vp, c_int, c_i64, c_int, c_vp]),
    'b2_slice_copy': (c_int, [c_vp, c_int, c_vp, c_int, c_i64, c_int, c_int, c_vp]),
    'b2_nchw_to_nhwc': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b2_nhwc_to_nchw': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b2_im2col': (c_int, [c_vp, c_vp] + [c_int] * 13 + [c_vp]),
    'b2_maxpool3x3s2': (c_int, [c_vp, c_vp, c_vp] + [c_int] * 6 + [c_vp]),
    'b2_maxpool3x3s2_bwd': (c_int, [c_vp, c_vp, c_vp] + [c_int] * 6 + [c_vp]),
    'b2_bilinear_fwd': (c_int, [c_vp, c_vp] + [c_int] * 10 + [c_vp]),
    'b2_bilinear_bwd': (c_int, [c_vp, c_vp] + [c_int] * 10 + [c_vp, c_f32, c_int, c_vp]),
    'b2_bilinear_bwd_nchw_workspace_floats': (c_i64, [c_int, c_int, c_int, c_int]),
    'b2_bilinear_bwd_nchw': (c_int, [c_vp, c_vp, c_vp] + [c_int] * 8 + [c_vp, c_f32, c_int, c_vp]),
    'b2_gap_fwd': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    'b2_gap_bwd': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b2_bcast_fwd': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    'b2_bcast_bwd': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    'b2_bn_workspace_doubles': (c_i64, [c_i64, c_int]),
    'b2_bn_stats': (c_int, [c_vp, c_i64, c_int, c_int, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'b2_bn_apply': (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_f32, c_vp, c_int,
                            c_vp, c_int, c_vp]),
    'b2_bn_bwd': (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_int, c_vp,
                          c_f32, c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    'b2_bn_fold': (c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_int, c_vp]),
    'b2_bn_eval_param_grad': (c_int, [c_vp, c_int, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_int,
                                      c_vp, c_vp, c_int, c_vp, c_vp]),
    'b2_bn_eval_param_grad_from_stats': (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    'b2_bn_stats_workspace_doubles': (c_i64, [c_int]),
    'b2_bn_eval_param_grad_wdot_from_stats': (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_f32,
                                                      c_vp, c_vp, c_int, c_vp, c_vp]),
    'b2_bn_eval_param_grad_wdot': (c_int, [c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_f32,
                                           c_vp, c_vp, c_int, c_vp, c_vp]),
    'b2_dropout_mask': (c_int, [c_vp, c_i64, c_f32, c_u64, c_u64, c_vp, c_vp]),
    'b2_add_inplace': (c_int, [c_vp, c_vp, c_i64, c_vp]),
    'b2_fill': (c_int, [c_vp, c_f32, c_i64, c_vp]),
    'b2_colsum': (c_int, [c_vp, c_int, c_i64, c_int, c_vp, c_int, c_vp, c_vp]),
}

# Functions that return a size/count rather than an error code.
_NON_STATUS = {'b2_version', 'b2_num_sms', 'b2_consistency_num_partials', 'b2_sample_reduce_blocks', 'b2_ce_num_partials',
               'b2_conv_wgrad_workspace', 'b2_bn_workspace_doubles', 'b2_conv_stats_rows',
               'b2_bn_stats_workspace_doubles', 'b2_bilinear_bwd_nchw_workspace_floats'}

_lib = None


def exported_symbols():
    """Names declared in include/b200seg.h that the shared library must export."""
    return sorted(list(_SIGS.keys()) + ['b2_last_error'])


def load():
    """Load libb200seg.so; raise loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2Error('libb200seg.so not found at {} — run `python -c "import __graft_entry__ as g; g.build()"` '
                      '(there is no CPU fallback for the B200 hot path)'.format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    lib.b2_last_error.restype = ctypes.c_char_p
    lib.b2_last_error.argtypes = []
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    lib.b2_debug_set.restype = None
    lib.b2_debug_set.argtypes = [c_int, c_int]
    lib.b2_debug_trace.restype = None
    lib.b2_debug_trace.argtypes = [c_vp]
    _lib = lib
    return lib


def call(name, *args):
    """Call a status-returning entry point; raise B2Error with b2_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _NON_STATUS:
        return rc
    if rc != 0:
        raise B2Error('{} failed (code {}): {}'.format(name, rc, lib.b2_last_error().decode()))
    return 0


def stream_ptr():
    """cudaStream_t of torch's current stream, as an integer for the C ABI."""
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B2Error('B200 hot path received a non-CUDA tensor; there is no CPU fallback')


def require_dense(dtype, *tensors):
    """The C ABI takes raw pointers: a tensor of another dtype or a non-contiguous view would be silently reinterpreted.
    Raises B2Error naming the offending argument position instead."""
    for i, t in enumerate(tensors):
        if t is None:
            continue
        if t.dtype != dtype:
            raise B2Error('B200 hot path: argument {} has dtype {}, expected {} (cast it; the kernels read raw pointers)'
                          .format(i, t.dtype, dtype))
        if not t.is_contiguous():
            raise B2Error('B200 hot path: argument {} is not contiguous (shape {}, strides {})'.format(i, tuple(t.shape), t.stride()))
