"""ctypes binding of libb200gp.so (include/b200gp.h).  There is no CPU fallback: a missing or
unloadable library is an ImportError-grade failure raised the moment anything needs the device."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libb200gp.so')
CSRC = os.path.join(_HERE, 'csrc')

_lib = None

c_int_p = ctypes.POINTER(ctypes.c_int32)
c_double_p = ctypes.POINTER(ctypes.c_double)


class B200GPError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into ms_project_b200/libb200gp.so (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(['make', '-C', CSRC, '-j8'], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise B200GPError('building libb200gp.so failed')
    return LIB_PATH


def _declare(lib):
    vp = ctypes.c_void_p
    lib.b200gp_version.restype = ctypes.c_char_p
    lib.b200gp_version.argtypes = []
    lib.b200gp_create.restype = ctypes.c_int
    lib.b200gp_create.argtypes = [ctypes.c_int, vp, ctypes.POINTER(vp)]
    lib.b200gp_destroy.restype = ctypes.c_int
    lib.b200gp_destroy.argtypes = [vp]
    lib.b200gp_last_error.restype = ctypes.c_char_p
    lib.b200gp_last_error.argtypes = [vp]
    lib.b200gp_set_quirks.restype = ctypes.c_int
    lib.b200gp_set_quirks.argtypes = [vp, ctypes.c_int]
    lib.b200gp_set_lookup.restype = ctypes.c_int
    lib.b200gp_set_lookup.argtypes = [vp, vp, ctypes.c_int]
    lib.b200gp_set_profiling.restype = ctypes.c_int
    lib.b200gp_set_profiling.argtypes = [vp, ctypes.c_int]
    lib.b200gp_set_panel_tiles.restype = ctypes.c_int
    lib.b200gp_set_panel_tiles.argtypes = [vp, ctypes.c_int]
    lib.b200gp_set_tile64.restype = ctypes.c_int
    lib.b200gp_set_tile64.argtypes = [vp, ctypes.c_int]
    lib.b200gp_set_potrf_la.restype = ctypes.c_int
    lib.b200gp_set_potrf_la.argtypes = [vp, ctypes.c_int]
    lib.b200gp_set_tma.restype = ctypes.c_int
    lib.b200gp_set_tma.argtypes = [vp, ctypes.c_int]
    lib.b200gp_set_fused_grad.restype = ctypes.c_int
    lib.b200gp_set_fused_grad.argtypes = [vp, ctypes.c_int]
    lib.b200gp_launch_count.restype = ctypes.c_longlong
    lib.b200gp_launch_count.argtypes = [vp]
    lib.b200gp_last_phase_ms.restype = ctypes.c_int
    lib.b200gp_last_phase_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    lib.b200gp_workspace_bytes.restype = ctypes.c_size_t
    lib.b200gp_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.b200gp_bind.restype = ctypes.c_int
    lib.b200gp_bind.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp]
    lib.b200gp_build_K.restype = ctypes.c_int
    lib.b200gp_build_K.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp]
    lib.b200gp_lml_grad.restype = ctypes.c_int
    lib.b200gp_lml_grad.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    lib.b200gp_posterior.restype = ctypes.c_int
    lib.b200gp_posterior.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    lib.b200gp_potrf_work_bytes.restype = ctypes.c_size_t
    lib.b200gp_potrf_work_bytes.argtypes = [ctypes.c_int]
    lib.b200gp_potrf.restype = ctypes.c_int
    lib.b200gp_potrf.argtypes = [vp, vp, ctypes.c_int, vp, c_double_p, c_int_p]
    lib.b200gp_hellinger_precompute.restype = ctypes.c_int
    lib.b200gp_hellinger_precompute.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int,
                                                ctypes.c_int, vp, vp, vp]
    lib.b200gp_hellinger_pairs_work_bytes.restype = ctypes.c_size_t
    lib.b200gp_hellinger_pairs_work_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.b200gp_hellinger_pairs.restype = ctypes.c_int
    lib.b200gp_hellinger_pairs.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int,
                                           ctypes.c_double, vp, vp, vp]
    lib.b200gp_gram_logdet.restype = ctypes.c_int
    lib.b200gp_gram_logdet.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp]
    lib.b200gp_hellinger_pairs_large_work_bytes.restype = ctypes.c_size_t
    lib.b200gp_hellinger_pairs_large_work_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.b200gp_hellinger_pairs_large.restype = ctypes.c_int
    lib.b200gp_hellinger_pairs_large.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int,
                                                 vp, vp, ctypes.c_int, vp, vp, vp]
    lib.b200gp_vector_pairs.restype = ctypes.c_int
    lib.b200gp_vector_pairs.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int, vp]
    lib.b200gp_centered_alignments.restype = ctypes.c_int
    lib.b200gp_centered_alignments.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
    lib.b200gp_alignment_min_slots.restype = ctypes.c_int
    lib.b200gp_alignment_min_slots.argtypes = [ctypes.c_int, ctypes.c_int]

EXPORTS = ['b200gp_version', 'b200gp_create', 'b200gp_destroy', 'b200gp_last_error', 'b200gp_set_quirks', 'b200gp_set_lookup',
           'b200gp_set_profiling', 'b200gp_set_panel_tiles', 'b200gp_set_fused_grad', 'b200gp_set_tile64', 'b200gp_set_potrf_la', 'b200gp_set_tma', 'b200gp_launch_count', 'b200gp_last_phase_ms', 'b200gp_workspace_bytes',
           'b200gp_bind', 'b200gp_build_K', 'b200gp_lml_grad', 'b200gp_posterior', 'b200gp_potrf_work_bytes', 'b200gp_potrf',
           'b200gp_hellinger_precompute', 'b200gp_hellinger_pairs_work_bytes', 'b200gp_hellinger_pairs', 'b200gp_gram_logdet', 'b200gp_hellinger_pairs_large_work_bytes',
           'b200gp_hellinger_pairs_large', 'b200gp_vector_pairs', 'b200gp_centered_alignments', 'b200gp_alignment_min_slots']


def lib():
    """The loaded library; raises B200GPError (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200GPError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                              '(there is no CPU fallback)' % LIB_PATH)
        try:
            l = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise B200GPError('cannot load %s: %s' % (LIB_PATH, e))
        _declare(l)
        _lib = l
    return _lib


def check(ctx, rc, what):
    if rc != 0:
        msg = lib().b200gp_last_error(ctx)
        raise B200GPError('%s failed (rc=%d): %s' % (what, rc, msg.decode() if msg else ''))
