"""ctypes binding of libpromonet_b200.so (the C ABI in include/promonet_b200.h)

There is no fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p, POINTER

from promonet_b200 import build as _build

MATH_FP32_SIMT = 0
MATH_BF16X3_TC = 1


class Error(RuntimeError):
    """A libpromonet_b200 call returned a negative status"""


# name -> (restype, argtypes); must list every symbol of include/promonet_b200.h
SIGNATURES = {
    'pmn_last_error': (c_char_p, []),
    'pmn_version': (c_int, []),
    'pmn_launch_count': (c_int64, []),
    'pmn_profile_enable': (None, [c_int]),
    'pmn_profile_reset': (None, []),
    'pmn_profile_read': (c_int, [c_char_p, POINTER(ctypes.c_double), POINTER(c_int64)]),
    'pmn_generator_create': (c_int, [POINTER(c_void_p)]),
    'pmn_generator_destroy': (None, [c_void_p]),
    'pmn_generator_set_tensor': (
        c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int, c_void_p]),
    'pmn_generator_finalize': (c_int, [c_void_p, c_int, c_void_p]),
    'pmn_generator_set_pair_mask': (c_int, [c_void_p, ctypes.c_uint]),
    'pmn_generator_set_f8': (c_int, [c_void_p, c_int]),
    'pmn_generator_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int]),
    'pmn_generator_forward': (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    'pmn_generator_staging_bytes': (c_size_t, [c_int, c_int, c_int]),
    'pmn_generator_forward_host': (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t,
         c_void_p, c_size_t, c_void_p]),
    'pmn_generator_features': (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int, c_int, c_void_p]),
    'pmn_fargan_create': (c_int, [POINTER(c_void_p)]),
    'pmn_fargan_destroy': (None, [c_void_p]),
    'pmn_fargan_set_tensor': (
        c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int, c_void_p]),
    'pmn_fargan_finalize': (c_int, [c_void_p, c_void_p]),
    'pmn_fargan_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int]),
    'pmn_fargan_forward': (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    'pmn_spectral_workspace_bytes': (c_size_t, [c_int, c_int]),
    'pmn_spectral_features': (
        c_int,
        [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int,
         c_void_p, c_size_t, c_void_p]),
    'pmn_linear_to_mel': (c_int, [c_void_p, c_void_p, c_float, c_int, c_int, c_void_p]),
    'pmn_viterbi_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'pmn_viterbi_decode': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
         c_void_p, c_size_t, c_void_p]),
    'pmn_pitch_create': (c_int, [POINTER(c_void_p)]),
    'pmn_pitch_destroy': (None, [c_void_p]),
    'pmn_pitch_set_tensor': (
        c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int, c_void_p]),
    'pmn_pitch_finalize': (c_int, [c_void_p, c_int, c_void_p]),
    'pmn_pitch_frames': (c_int, [c_int, c_int, ctypes.c_double]),
    'pmn_pitch_workspace_bytes': (
        c_size_t, [c_int, c_int, c_int, ctypes.c_double, c_int]),
    'pmn_pitch_forward': (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_double, c_float, c_float,
         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
         c_void_p, c_size_t, c_void_p]),
    'pmn_weight_norm_fold': (
        c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'pmn_pack_conv1d_weight': (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pmn_conv1d': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
         c_float, c_int, c_void_p]),
    'pmn_debug_tc_counters': (None, [c_void_p]),
    'pmn_conv1d_tc_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'pmn_conv1d_tc': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
         c_void_p, c_size_t, c_void_p]),
    'pmn_conv1d_tc_f8': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
         c_void_p, c_size_t, c_void_p]),
    'pmn_conv1d_tc_general': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
         c_float, c_float, c_void_p, c_size_t, c_void_p]),
    'pmn_conv_pair_tc_workspace_bytes': (c_size_t, [c_int, c_int]),
    'pmn_debug_pair_tc': (None, [c_void_p, c_int]),
    'pmn_conv_pair_tc': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float,
         c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    'pmn_conv_transpose1d': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
         c_int, c_int, c_float, c_void_p]),
    'pmn_conv_transpose1d_tc_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'pmn_conv_transpose1d_tc': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
         c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    # ---- training-step operators ----
    'pmn_conv_gemm': (
        c_int,
        [c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p,
         c_int, c_float, c_void_p, c_float, c_void_p, c_float, c_int, c_void_p, c_void_p]),
    'pmn_conv_gemm_tc': (
        c_int,
        [c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p,
         c_int, c_float, c_void_p, c_float, c_void_p, c_float, c_int, c_void_p, c_void_p]),
    'pmn_conv_tc_channel_pad': (c_int, [c_int]),
    'pmn_conv_tc_packed_floats': (c_size_t, [c_int, c_int, c_int]),
    'pmn_debug_train_tc_counters': (None, [c_void_p]),
    'pmn_pack_weight_taps': (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'pmn_conv_wgrad': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_float,
         c_void_p, c_void_p, c_void_p]),
    'pmn_conv_wgrad_tc': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_float,
         c_void_p, c_void_p, c_void_p]),
    'pmn_extract_grouped': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'pmn_prepare_weights': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'pmn_transpose_weight': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pmn_weight_norm_backward_table': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'pmn_weight_norm_backward': (
        c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'pmn_reflect_pad': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'pmn_reflect_pad_backward': (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pmn_axpby': (c_int, [c_float, c_void_p, c_float, c_void_p, c_int64, c_void_p]),
    'pmn_mse_to_target': (
        c_int, [c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    'pmn_l1_mean': (
        c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    'pmn_adamw': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
         c_float, c_int, c_float, c_void_p, c_void_p]),
    'pmn_adamw_peer': (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_float, c_float,
         c_float, c_float, c_float, c_int, c_void_p, c_void_p]),
    'pmn_row_sum': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pmn_features': (
        c_int,
        [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
         c_void_p, c_int, c_int, c_void_p]),
    'pmn_pitch_bins': (
        c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p]),
    'pmn_embedding_backward': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pmn_global_features': (
        c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pmn_stft_magnitude': (
        c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    'pmn_stft_magnitude_backward': (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_int, c_void_p]),
    'pmn_mel_loss': (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    'pmn_channel_sum': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'pmn_copy_columns': (
        c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int64, c_int, c_int, c_void_p]),
    'pmn_grid_sample': (
        c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pmn_dft_basis': (c_int, [c_void_p, c_int, c_void_p]),
    'pmn_spectral_convergence': (
        c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    'pmn_frame_overlap_add': (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    # ---- multi-resolution discriminator front end ----
    'pmn_dft_basis_rect': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'pmn_complex_magnitude': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pmn_complex_magnitude_backward': (
        c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    # ---- in-training validation ----
    'pmn_metrics_update': (
        c_int,
        [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p]),
    'pmn_edit_contour': (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_float,
         c_float, c_void_p]),
}
METRICS_SLOTS = 12    # PMN_METRICS_SLOTS


class WeightDesc(ctypes.Structure):
    """pmn_weight_desc"""
    _fields_ = [
        ('v', c_void_p), ('g', c_void_p), ('w', c_void_p), ('packed', c_void_p),
        ('packed_t', c_void_p), ('wt', c_void_p), ('dense', c_void_p),
        ('dim0', c_int), ('dim1', c_int), ('taps', c_int), ('groups', c_int)]


class WeightNormDesc(ctypes.Structure):
    """pmn_weight_norm_desc"""
    _fields_ = [
        ('v', c_void_p), ('g', c_void_p), ('gw', c_void_p), ('gv', c_void_p), ('gg', c_void_p),
        ('dim0', c_int), ('inner', c_int)]


class ConvGeometry(ctypes.Structure):
    """pmn_conv_geometry"""
    _fields_ = [(name, c_int) for name in (
        'batch', 'c_in', 'c_out', 'h_in', 'w_in', 'h_out', 'w_out',
        'kh', 'kw', 'sh', 'sw', 'dh', 'dw', 'ph', 'pw',
        'channel_stride', 'position_stride', 'batch_stride')]


_library = None


def library():
    """Load (building first if the in-tree .so is stale or absent)"""
    global _library
    if _library is None:
        path = _build.build()
        lib = ctypes.CDLL(str(path))
        for name, (restype, argtypes) in SIGNATURES.items():
            function = getattr(lib, name)
            function.restype = restype
            function.argtypes = argtypes
        _library = lib
    return _library


def check(status):
    if status != 0:
        message = library().pmn_last_error()
        raise Error(
            f'libpromonet_b200 status {status}: '
            f'{message.decode() if message else "unknown"}')


def ptr(tensor):
    """Device (or host) address of a contiguous tensor; None -> NULL"""
    if tensor is None:
        return None
    if not tensor.is_contiguous():
        raise ValueError('tensor passed to libpromonet_b200 must be contiguous')
    return tensor.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(library().pmn_launch_count())


def profile(enabled):
    """Turn per-kernel CUDA-event timing on or off (resets the totals)"""
    library().pmn_profile_reset()
    library().pmn_profile_enable(int(bool(enabled)))


def profile_read(kernel):
    """(total device milliseconds, launches) of `kernel` since profile(True)"""
    total, launches = ctypes.c_double(), c_int64()
    check(library().pmn_profile_read(
        kernel.encode(), ctypes.byref(total), ctypes.byref(launches)))
    return total.value, launches.value
