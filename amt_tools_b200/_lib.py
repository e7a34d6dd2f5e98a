"""
ctypes binding of libamtfeat.so (include/amtfeat.h).  This is the whole Python <-> native boundary:
plain pointers and sizes; torch is used only to own device memory and to name the CUDA stream.

There is no CPU fallback: if the shared library is missing the import fails loudly, and
`amtfeat_process` on a host-only plan returns AMTFEAT_ERR_NO_DEVICE.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('AMTFEAT_LIB') or os.path.join(_HERE, 'libamtfeat.so')   # env override: kernel A/B builds

MAX_HARMONICS = 16
(WAVEFORM, STFT, MEL, VQT, HVQT, POWER) = range(6)
OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_WORKSPACE = range(5)
RES_KAISER_BEST, RES_KAISER_FAST = range(2)


class Config(C.Structure):
    _fields_ = [
        ('kind', C.c_int32), ('hop_length', C.c_int32), ('sample_rate', C.c_double),
        ('decibels', C.c_int32), ('center', C.c_int32), ('win_length', C.c_int32), ('n_fft', C.c_int32),
        ('n_mels', C.c_int32), ('htk', C.c_int32), ('n_bins', C.c_int32), ('bins_per_octave', C.c_int32),
        ('fmin', C.c_double), ('gamma', C.c_double), ('n_harmonics', C.c_int32), ('n_decim_taps', C.c_int32),
        ('harmonics', C.c_double * MAX_HARMONICS), ('decim_taps', C.POINTER(C.c_double)),
    ]


class AmtfeatError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'libamtfeat.so is missing (%s). Build it with `python amt_tools_b200/build.py` '
            '(needs nvcc); amt_tools_b200 has no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    i64p = C.POINTER(C.c_int64)
    sig = {
        'amtfeat_version': (C.c_int, []),
        'amtfeat_last_error': (C.c_char_p, []),
        'amtfeat_plan_create': (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(P)]),
        'amtfeat_plan_destroy': (None, [P]),
        'amtfeat_expected_frames': (C.c_int64, [P, C.c_int64]),
        'amtfeat_output_frames': (C.c_int64, [P, C.c_int64]),
        'amtfeat_sample_range': (C.c_int, [P, C.c_int64, i64p, i64p]),
        'amtfeat_num_samples_required': (C.c_int64, [P]),
        'amtfeat_times': (C.c_int, [P, C.c_int64, C.c_int, C.POINTER(C.c_double), C.c_int64]),
        'amtfeat_early_ds_count': (C.c_int, [P, C.c_int]),
        'amtfeat_num_channels': (C.c_int, [P]),
        'amtfeat_feature_size': (C.c_int, [P]),
        'amtfeat_out_shape': (C.c_int, [P, C.c_int64, C.c_int64 * 3, C.POINTER(C.c_int)]),
        'amtfeat_plan_describe': (C.c_int, [P, C.c_char_p, C.c_size_t]),
        'amtfeat_clip_describe': (C.c_int, [P, C.c_int64, C.c_char_p, C.c_size_t]),
        'amtfeat_workspace_bytes': (C.c_size_t, [P, C.c_int, i64p]),
        'amtfeat_launch_count': (C.c_int, [P, C.c_int, i64p]),
        'amtfeat_process': (C.c_int, [P, P, i64p, i64p, i64p, C.c_int, P, P, C.c_size_t, P]),
        'amtfeat_process_raw': (C.c_int, [P, P, i64p, i64p, i64p, C.c_int, P, P, C.c_size_t, P]),
        'amtfeat_range_reference': (C.c_int, [P, P, C.c_int64, C.c_int64, C.c_int64, P, P]),
        'amtfeat_range_finish': (C.c_int, [P, P, C.c_int64, C.c_int64, C.c_int64, P, P, C.c_int64, C.c_int64, P]),
        'amtfeat_process_host': (C.c_int, [P, P, i64p, i64p, i64p, C.c_int, P, C.c_int64, C.c_int64, P, P, P,
                                           C.c_size_t, P]),
        'amtfeat_pipeline_create': (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_size_t, C.POINTER(P)]),
        'amtfeat_pipeline_submit': (C.c_int, [P, P, P, i64p, i64p, i64p, C.c_int, P, C.c_int64, C.c_int64, i64p]),
        'amtfeat_pipeline_wait': (C.c_int, [P, C.c_int64]),
        'amtfeat_pipeline_destroy': (None, [P]),
        'amtfeat_framify_hops': (C.c_int64, [C.c_int64, C.c_int, C.c_int, C.c_int]),
        'amtfeat_framify': (C.c_int, [P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, P, P]),
        'amtfeat_resampler_create': (C.c_int, [C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(P)]),
        'amtfeat_resampler_destroy': (None, [P]),
        'amtfeat_resampler_out_len': (C.c_int64, [P, C.c_int64]),
        'amtfeat_resampler_table': (C.c_int64, [P, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        'amtfeat_ingest_workspace_bytes': (C.c_size_t, [C.c_int]),
        'amtfeat_resample': (C.c_int, [P, P, i64p, i64p, C.c_int, P, i64p, P, C.c_size_t, P]),
        'amtfeat_pcm16_to_float': (C.c_int, [P, C.c_int64, C.c_float, P, P]),
        'amtfeat_to_mono': (C.c_int, [P, C.c_int64, C.c_int, P, P]),
        'amtfeat_rms_norm': (C.c_int, [P, i64p, i64p, C.c_int, P, C.c_size_t, P]),
        'amtfeat_profile_enable': (C.c_int, [P, C.c_int]),
        'amtfeat_profile_read': (C.c_int, [P, C.c_char_p, C.c_size_t]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib, sorted(sig)


lib, EXPORTS = _load()


def check(status):
    if status != OK:
        msg = (lib.amtfeat_last_error() or b'').decode()
        if status == ERR_INVALID:
            raise ValueError(msg)
        raise AmtfeatError('libamtfeat status %d: %s' % (status, msg))


def i64_array(values):
    return (C.c_int64 * len(values))(*[int(v) for v in values])
