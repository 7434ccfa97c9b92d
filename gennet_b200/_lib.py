"""ctypes binding of libgennet_b200.so (the C ABI declared in include/gennet_b200.h).

The product has NO CPU path: importing this module without the built library, or calling
into it without a B200, raises.  torch is used only to own device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# GENNET_B200_LIB: developer override (kernel experiments built to another file); the product is the in-tree library
LIB_PATH = os.environ.get('GENNET_B200_LIB') or os.path.join(_HERE, 'libgennet_b200.so')

c_f = ctypes.c_float
c_d = ctypes.c_double
c_i = ctypes.c_int
c_ll = ctypes.c_longlong
c_u64 = ctypes.c_uint64
c_p = ctypes.c_void_p

# name -> argtypes (return type is always int unless listed in _RESTYPES)
_SIGS = {
    'gn_version': [],
    'gn_device_ok': [],
    'gn_fft_plan_create': [c_i, ctypes.POINTER(c_p)],
    'gn_fft_plan_destroy': [c_p],
    'gn_whiten_td_f32': [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_p],
    'gn_irfft_f32': [c_p, c_p, c_p, c_p, c_i, c_f, c_i, c_i, c_p],
    'gn_synth_f32': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_f, c_u64, c_u64, c_p],
    'gn_bbh_assemble_f32': [c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_mean_std_f32': [c_p, c_ll, c_p, c_p],
    'gn_add_scaled_f32': [c_p, c_p, c_f, c_ll, c_p],
    'gn_burst_waveforms_f32': [c_p, c_p, c_i, c_i, c_f, c_f, c_f, c_f, c_p],
    'gn_conv1d_fwd_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_conv1d_dgrad_f32': [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_wgrad_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv_w_to_bf16': [c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_cast_f32_to_bf16': [c_p, c_p, c_ll, c_p],
    'gn_cast_bf16_to_f32': [c_p, c_p, c_ll, c_p],
    'gn_conv1d_fwd_bf16': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_conv1d_dgrad_bf16': [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_conv1d_wgrad_bf16': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_split_f32_bf16': [c_p, c_p, c_ll, c_i, c_p],
    'gn_conv_w_split_bf16': [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_fwd_bf16x3': [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_p],
    'gn_conv1d_dgrad_bf16x3': [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_p],
    'gn_conv1d_wgrad_bf16x3': [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_split_pad_f32_bf16': [c_p, c_p, c_ll, c_i, c_i, c_i, c_p],
    'gn_dense_w_split_bf16': [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    'gn_dense_fwd_bf16x3': [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_p],
    'gn_dense_dgrad_bf16x3': [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_p],
    'gn_dense_wgrad_bf16x3': [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_amax_f32': [c_p, c_ll, c_p, c_p],
    'gn_split_f32_f16x2': [c_p, c_p, c_p, c_i, c_ll, c_p],
    'gn_split_colsum_f32_f16x2': [c_p, c_p, c_p, c_i, c_ll, c_i, c_p, c_p],
    'gn_conv_w_split_f16x2': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_conv1d_fwd_f16x2': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_conv1d_fwd_stats_f16x2': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_conv1d_dgrad_f16x2': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_conv1d_wgrad_f16x2': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_split_pad_f32_f16x2': [c_p, c_p, c_p, c_ll, c_i, c_i, c_p],
    'gn_dense_w_split_f16x2': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_dense_fwd_f16x2': [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_dense_dgrad_f16x2': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_dense_wgrad_f16x2': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_smallcin_fwd_bf16': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_conv1d_smallcin_wgrad_bf16': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_smallcin_dgrad_bf16': [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_cout1_fwd_bf16': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_cout1_dgrad_bf16': [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_cout1_wgrad_bf16': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_smallcin_fwd_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_conv1d_smallcin_wgrad_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_smallcin_dgrad_f32': [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_cout1_fwd_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_cout1_dgrad_f32': [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv1d_cout1_wgrad_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_dense_small_fwd_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_dense_small_dgrad_f32': [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_dense_small_wgrad_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_upsample1d_fwd_bf16': [c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    'gn_upsample1d_bwd_bf16': [c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    'gn_dense_small_fwd_bf16': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_dense_small_dgrad_bf16': [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_dense_small_wgrad_bf16': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_act_bwd_bf16': [c_p, c_p, c_p, c_ll, c_i, c_f, c_p],
    'gn_conv2d_w2_pack_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_conv2d_w2_unpack_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    'gn_dense_fwd_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_p],
    'gn_dense_dgrad_f32': [c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_dense_wgrad_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_bn_stats_f32': [c_p, c_ll, c_i, c_p, c_p, c_p],
    'gn_bn_finalize_f32': [c_p, c_p, c_d, c_i, c_f, c_f, c_p, c_p, c_p, c_i, c_p, c_d, c_p],
    'gn_bn_apply_f32': [c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_i, c_f, c_i, c_p],
    'gn_bn_bwd_sums_f32': [c_p, c_p, c_p, c_ll, c_i, c_p, c_p],
    'gn_bn_bwd_apply_f32': [c_p, c_p, c_p, c_p, c_p, c_d, c_p, c_p, c_p, c_ll, c_i, c_p],
    'gn_bn_stats_bf16': [c_p, c_ll, c_i, c_p, c_p],
    'gn_chain_fwd_bf16': [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_i, c_f, c_i, c_f, c_p, c_u64, c_u64, c_ll, c_i, c_p],
    'gn_chain_bwd_sums_bf16': [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_i, c_f, c_p, c_u64, c_u64, c_ll, c_i, c_p, c_p],
    'gn_chain_bwd_bf16': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_d, c_i, c_f, c_i, c_f, c_p, c_u64, c_u64, c_p, c_p, c_ll, c_i, c_p],
    'gn_bn_sums_f32': [c_p, c_ll, c_i, c_p, c_p],
    'gn_chain_fwd_f32': [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_i, c_f, c_i, c_f, c_p, c_u64, c_u64, c_ll, c_i, c_p],
    'gn_chain_bwd_sums_f32': [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_i, c_f, c_p, c_u64, c_u64, c_ll, c_i, c_p, c_p],
    'gn_chain_bwd_f32': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_d, c_i, c_f, c_i, c_f, c_p, c_u64, c_u64, c_p, c_p, c_ll, c_i, c_p],
    'gn_chain_fwd_amax_f32': [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_i, c_f, c_i, c_f, c_p, c_u64, c_u64, c_ll, c_i, c_p, c_p],
    'gn_chain_fwd_planes_f32': [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_i, c_f, c_i, c_f, c_p, c_u64, c_u64, c_ll, c_i, c_p, c_p, c_f, c_p],
    'gn_chain_bwd_amax_f32': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_d, c_i, c_f, c_i, c_f, c_p, c_u64, c_u64, c_p, c_p, c_ll, c_i, c_p, c_p],
    'gn_act_fwd_f32': [c_p, c_p, c_ll, c_i, c_f, c_p],
    'gn_act_bwd_f32': [c_p, c_p, c_p, c_ll, c_i, c_f, c_p],
    'gn_noise_fwd_f32': [c_p, c_p, c_p, c_ll, c_i, c_f, c_p],
    'gn_noise_bwd_f32': [c_p, c_p, c_p, c_ll, c_i, c_f, c_p],
    'gn_noise_draw_f32': [c_p, c_ll, c_i, c_f, c_u64, c_u64, c_p],
    'gn_uniform_f32': [c_p, c_ll, c_f, c_f, c_u64, c_u64, c_p],
    'gn_normal_f32': [c_p, c_ll, c_f, c_f, c_u64, c_u64, c_p],
    'gn_upsample1d_fwd_f32': [c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    'gn_upsample1d_bwd_f32': [c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    'gn_maxpool1d_fwd_f32': [c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    'gn_maxpool1d_bwd_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    'gn_gap_fwd_f32': [c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_gap_bwd_f32': [c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_transpose_f32': [c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_reg_terms_f32': [c_p, c_p, c_ll, c_f, c_f, c_p, c_p],
    'gn_axpy_f32': [c_p, c_p, c_f, c_ll, c_p],
    'gn_gather_rows_f32': [c_p, c_p, c_p, c_i, c_ll, c_p],
    'gn_kde2d_pdf_f32': [c_p, c_i, c_p, c_i, c_d, c_d, c_d, c_d, c_p, c_p],
    'gn_overlap_sums_f32': [c_p, c_p, c_ll, c_p, c_p],
    'gn_percentiles_f32': [c_p, c_i, c_i, c_p, c_i, c_p, c_p],
    'gn_maxnorm_roll_f32': [c_p, c_p, c_p, c_i, c_i, c_p],
    'gn_flip_transpose_f32': [c_p, c_p, c_i, c_i, c_i, c_p],
    'gn_stack_residual_fwd_f32': [c_p, c_p, c_p, c_i, c_i, c_p],
    'gn_stack_residual_bwd_f32': [c_p, c_p, c_i, c_i, c_p],
    'gn_stack_pair_f32': [c_p, c_p, c_p, c_ll, c_p],
    'gn_residual_moments_fwd_f32': [c_p, c_p, c_p, c_i, c_i, c_p],
    'gn_residual_moments_bwd_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_d, c_p],
    'gn_loss_fwd_bwd_f32': [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_i, c_i, c_p],
    'gn_adam_step_f32': [c_p, c_p, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_f, c_p],
    'gn_sgd_step_f32': [c_p, c_p, c_ll, c_f, c_f, c_p],
}

# constants mirrored from the header
ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_LEAKY, ACT_RELU_MAX, ACT_ELU = range(7)
NOISE_DROPOUT, NOISE_GDROPOUT, NOISE_GNOISE = range(3)
LOSS_BCE, LOSS_MSE, LOSS_CHISQ = range(3)

_lib = None


class GennetError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GennetError('%s is missing: build it with `python gennet_b200/build.py` '
                          '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.gn_last_error.restype = ctypes.c_char_p
    lib.gn_last_error.argtypes = []
    for name, args in _SIGS.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.argtypes = args
        fn.restype = c_i
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGS) + ['gn_last_error']


def require_device():
    lib = load()
    if not torch.cuda.is_available() or not lib.gn_device_ok():
        raise GennetError('gennet_b200 needs a B200 (sm_100) GPU: %s' %
                          (lib.gn_last_error().decode() or 'torch.cuda.is_available() is False'))


# launch accounting: every C-ABI call below launches at least one kernel of this library
COUNTS = {'calls': 0}
PROFILE = None      # set to a list to record (name, tag, start_event, end_event, args) per call


def call(name, *args, tag=None):
    """Invoke an entry point; raise GennetError with gn_last_error() on failure."""
    COUNTS['calls'] += 1
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(load(), name)(*args)
        e1.record()
        PROFILE.append((name, tag, e0, e1, args))
    else:
        rc = getattr(load(), name)(*args)
    if rc != 0:
        raise GennetError('%s failed (%d): %s' % (name, rc, load().gn_last_error().decode()))


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GennetError('expected a CUDA tensor (the library has no CPU path)')
    if dtype is not None and t.dtype != dtype:
        raise GennetError('expected dtype %s, got %s' % (dtype, t.dtype))
    if not t.is_contiguous():
        raise GennetError('expected a contiguous tensor')
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
