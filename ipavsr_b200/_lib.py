"""ctypes binding of libipavsr_b200.so (the C-ABI declared in include/ipavsr_b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc (in-tree); if that is not
possible, or a call returns a non-zero status, a RuntimeError is raised.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libipavsr_b200.so')
HEADER = os.path.join(HERE, '..', 'include', 'ipavsr_b200.h')

_lib = None

P = C.c_void_p
I = C.c_int
F = C.c_float
U64 = C.c_uint64
I64 = C.c_int64

_SIGS = {
    'ipavsr_last_error': (C.c_char_p, []),
    'ipavsr_version': (I, []),
    'ipavsr_source_hash': (C.c_char_p, []),
    'ipavsr_launch_count': (U64, []),
    'ipavsr_launch_count_add': (None, [U64]),
    'ipavsr_device_info': (I, [P, P, P, P]),
    'ipavsr_gemm': (I, [I, I, I, I, I, I, P, I, P, I, P, I, P, I, I, P, U64, P]),
    'ipavsr_gemm_workspace_bytes': (U64, [I, I, I, I, I, I]),
    'ipavsr_gemm_tc_supported': (I, [I, I, I, I, I, P, I, P, I, P, I]),
    'ipavsr_tf32_split_rna': (I, [P, P, P, U64, P]),
    'ipavsr_gemm_tf32x3_presplit': (I, [I, I, I, I, I, P, P, I, P, P, I, P, I, P, I, I, P, P, P]),
    'ipavsr_amax': (I, [P, I, I64, I, P, P]),
    'ipavsr_f16_split': (I, [P, I, I64, I, P, P, I, P, P, I, P]),
    'ipavsr_f16_split_segments': (I, [P, P, P, U64, P, I, P, P, P]),
    'ipavsr_gemm_f16x3': (I, [I, I, I, I, I, P, P, I, P, P, P, I, P, P, I, P, I, I, P, P, P, I, P]),
    'ipavsr_gemm_f16_supported': (I, [I, I, I, P, I, P, I]),
    'ipavsr_dense_bwd_prep': (I, [P, I, P, I, P, I, P, I, I, I, I, P, P]),
    'ipavsr_dense_bwd_prep_f16': (I, [P, I, P, I, P, I, I, I, I, P, P, P, I, P, P]),
    'ipavsr_dense_bwd_prep_f16_supported': (I, [P, I, P, I, I, P, P, I]),
    'ipavsr_colsum': (I, [P, I, P, I, I, I, P]),
    'ipavsr_delta_fwd': (I, [P, I, P, I, I, I, I, I, I, P]),
    'ipavsr_delta_bwd': (I, [P, I, P, I, I, I, I, I, I, P]),
    'ipavsr_lstm_fwd': (I, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, P, U64, P]),
    'ipavsr_lstm_bwd': (I, [P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, I, I, P, U64, P]),
    'ipavsr_lstm_workspace_bytes': (U64, [I, I, I]),
    'ipavsr_lstm_fwd_f16': (I, [P, P, P, P, I, P, P, P, P, P, P, P, P, I, I, I, I, I, P]),
    'ipavsr_lstm_fwd_f16_supported': (I, [I, I, I, I]),
    'ipavsr_lstm_bwd_f16': (I, [P, P, P, P, P, I, P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, I, P, P, P, P, P, U64, P]),
    'ipavsr_lstm_bwd_f16_supported': (I, [I, I, I, I, F]),
    'ipavsr_fuse_sum': (I, [P, P, I, P, P, I, I, I, P]),
    'ipavsr_adasum_bwd_coeff': (I, [P, I, P, P, I, P, I, I, I, P]),
    'ipavsr_copy2d': (I, [P, I, P, I, I, I, P, I, P]),
    'ipavsr_slice_last': (I, [P, I, P, I, I, I, I, I, I, P]),
    'ipavsr_batch_gather': (I, [P, I, P, P, P, P, P, I, P, P, I, I, I, P]),
    'ipavsr_vote_eval': (I, [P, I, P, P, I, I, I, P, P, P, P]),
    'ipavsr_dropout': (I, [P, I, P, P, I, I, I, F, P]),
    'ipavsr_dropout_mask': (I, [P, U64, F, U64, U64, P]),
    'ipavsr_bn_stats': (I, [P, I, P, I, I, P]),
    'ipavsr_bn_fwd': (I, [P, I, P, I, P, P, P, P, P, P, P, I, I, I64, F, F, I, I, P]),
    'ipavsr_bn_bwd_stats': (I, [P, I, P, I, P, P, P, I, I, P]),
    'ipavsr_bn_bwd': (I, [P, I, P, I, P, P, P, P, P, I, P, P, I, I, I64, I, P]),
    'ipavsr_lstm_steps_supported': (I, [I, I, I, I]),
    'ipavsr_lstm_steps_workspace_bytes': (U64, [I, I, I]),
    'ipavsr_lstm_fwd_f16_steps': (I, [P, P, P, P, P, I, P, P, P, P, P, P, P, P, I, I, I, I, I, P, P, U64, P]),
    'ipavsr_lstm_bwd_f16_steps': (I, [P, P, P, P, P, I, P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, I, P, P, P, P, P, U64,
                                      P]),
    'ipavsr_softmax': (I, [P, I, P, I, I, I, P]),
    'ipavsr_temporal_softmax_loss': (I, [P, I, P, P, P, P, I, I, I, F, P, P]),
    'ipavsr_categorical_crossentropy': (I, [P, I, P, P, P, I, I, I, F, P, P]),
    'ipavsr_squared_error': (I, [P, I, P, I, P, P, I, I64, I, F, P]),
    'ipavsr_l2_penalty': (I, [P, P, U64, P, P, P, F, P]),
    'ipavsr_optim_step': (I, [I, P, P, P, P, U64, F, P, P, F, F, F, F, F, P]),
    'ipavsr_norm_samplewise': (I, [P, I, P, I, I64, I, P]),
    'ipavsr_norm_featurewise_stats': (I, [P, I, P, P, P, I64, I, P]),
    'ipavsr_norm_featurewise_apply': (I, [P, I, P, P, P, I, I64, I, P]),
    'ipavsr_seq_mean_sub': (I, [P, I, P, I, P, I, I, P]),
    'ipavsr_diff_image': (I, [P, I, P, I, P, I, I, P]),
    'ipavsr_deltas_fir': (I, [P, I, P, I, P, I, I, I, I, P]),
    'ipavsr_deltas_fir_f32': (I, [P, I, P, I, P, I, I, I, I, P]),
    'ipavsr_zigzag_indices': (I, [I, I, P]),
    'ipavsr_dct_basis': (I, [P, I, P, I, I, P]),
    'ipavsr_dct_project': (I, [P, I, P, I, P, I, I64, I, I, P]),
    'ipavsr_gather_cols': (I, [P, I, P, P, I, I64, I, P]),
    'ipavsr_col_abs_sum': (I, [P, I, P, I64, I, P]),
    'ipavsr_reorder': (I, [P, I, P, I, I64, I, I, I, P]),
    'ipavsr_align_fill': (I, [P, I, P, I, P, P, P, I, I, I64, P]),
    'ipavsr_gather_rows': (I, [P, I64, P, I64, I, P, P, I64, P]),
    'ipavsr_colsum_masked': (I, [P, I, P, I, P, I, I, I, P]),
    'ipavsr_upload_ragged': (I, [P, I64, I64, P, I64, P, P, I, P]),
    'ipavsr_debug_gemm_timestamps': (I, [P]),
    'ipavsr_debug_gemm_persistent_launches': (U64, []),
    'ipavsr_debug_lstm_timestamps': (I, [P]),
    'ipavsr_fill': (I, [P, U64, F, P]),
    'ipavsr_tf32_split': (I, [P, P, P, U64, P]),
}


def header_symbols():
    """Every function name declared in include/ipavsr_b200.h."""
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ipavsr_[a-z0-9_]+)\s*\(', text)))


def load():
    """Load (building first if necessary) the shared library and attach argument types."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build
    want = build.source_hash()
    if not os.path.exists(LIB_PATH) or build.built_hash() != want:
        # missing or stale (sources edited since the last build): rebuild in-tree when nvcc is there, otherwise refuse —
        # calling a library built from other sources through these signatures would corrupt arguments silently
        if not os.path.exists(build.NVCC):
            raise RuntimeError('libipavsr_b200.so is %s and nvcc (%s) is not available to build it'
                               % ('missing' if not os.path.exists(LIB_PATH) else 'stale', build.NVCC))
        build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    got = lib.ipavsr_source_hash().decode()
    if got != want:
        raise RuntimeError('libipavsr_b200.so was built from other sources (hash %s, tree %s): run python -m ipavsr_b200.build'
                           % (got, want))
    _lib = lib
    return lib


class IpavsrError(RuntimeError):
    pass


def check(status, what=''):
    if status != 0:
        msg = load().ipavsr_last_error()
        raise IpavsrError('%s failed with status %d: %s' % (what or 'ipavsr call', status,
                                                            msg.decode() if msg else '?'))


_fn_cache = {}


def call(name, *args):
    """Invoke a status-returning entry point and raise on failure."""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    rc = fn(*args)
    if rc != 0:
        check(rc, name)
