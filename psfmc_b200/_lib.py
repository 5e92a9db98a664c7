"""
ctypes binding of the C ABI declared in include/psfmc_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a)
as ``psfmc_b200/libpsfmc_b200.so``. There is no CPU fallback: if the library is
missing or fails to load, importing this module's :func:`load` raises.
"""
import ctypes
import os

PEER_HANDLE_BYTES = 64
ABI_VERSION = 2

SKY, POINT, SERSIC = 0, 1, 2
FLAG_ANGLE_DEGREES, FLAG_BILINEAR = 1, 2
P_ADU, P_X, P_Y, P_MAG, P_REFF, P_REFF_B, P_INDEX, P_ANGLE = 0, 0, 1, 2, 3, 4, 5, 6
NSLOTS = 7
MAX_COMPONENTS = 32
PREC_FP64, PREC_FP32, PREC_FP64_RAWF32 = 0, 1, 2
DESC_NO_FP64_RESCUE = 1
DESC_LOW_LATENCY = 2
PRECISIONS = {'fp64': PREC_FP64, 'fp32': PREC_FP32, 'fp64_rawf32': PREC_FP64_RAWF32}
IMAGE_BITS = {'raw_model': 1, 'convolved_model': 2, 'residual': 4,
              'composite_ivm': 8, 'point_source_subtracted': 16}


class Slot(ctypes.Structure):
    _fields_ = [('theta_index', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('value', ctypes.c_double)]


class Component(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('flags', ctypes.c_int32),
                ('slot', Slot * NSLOTS)]


class Desc(ctypes.Structure):
    _fields_ = [
        ('abi_version', ctypes.c_int32),
        ('height', ctypes.c_int32), ('width', ctypes.c_int32),
        ('obs_data', ctypes.POINTER(ctypes.c_double)),
        ('obs_var', ctypes.POINTER(ctypes.c_double)),
        ('bad_px', ctypes.POINTER(ctypes.c_uint8)),
        ('n_psf', ctypes.c_int32),
        ('psf_height', ctypes.c_int32), ('psf_width', ctypes.c_int32),
        ('psf', ctypes.POINTER(ctypes.c_double)),
        ('psf_var', ctypes.POINTER(ctypes.c_double)),
        ('mag_zeropoint', ctypes.c_double),
        ('n_components', ctypes.c_int32),
        ('components', ctypes.POINTER(Component)),
        ('psf_index', Slot),
        ('precision', ctypes.c_int32),
        ('n_devices', ctypes.c_int32),
        ('devices', ctypes.POINTER(ctypes.c_int32)),
        ('max_batch', ctypes.c_int32),
        ('flags', ctypes.c_int32),
    ]


class Info(ctypes.Structure):
    _fields_ = [
        ('height', ctypes.c_int32), ('width', ctypes.c_int32),
        ('n_components', ctypes.c_int32), ('n_sersic', ctypes.c_int32),
        ('n_point', ctypes.c_int32), ('n_psf', ctypes.c_int32),
        ('precision', ctypes.c_int32), ('n_devices', ctypes.c_int32),
        ('path', ctypes.c_int32), ('kernels_per_call', ctypes.c_int32),
        ('flops_per_eval', ctypes.c_double),
        ('fft_flops_per_eval', ctypes.c_double),
        ('hbm_bytes_per_eval', ctypes.c_double),
        ('launches_total', ctypes.c_int64),
        ('kappa_table', ctypes.c_int32), ('rescued_total', ctypes.c_int32),
        ('graph_replays', ctypes.c_int32), ('rescued_on_device', ctypes.c_int32),
    ]


PRIOR_OTHER, PRIOR_UNIFORM, PRIOR_NORMAL, PRIOR_WEIBULL_MIN = 0, 1, 2, 3


class PriorColumn(ctypes.Structure):
    _fields_ = [('family', ctypes.c_int32), ('theta_index', ctypes.c_int32),
                ('valid', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('loc', ctypes.c_double), ('scale', ctypes.c_double),
                ('log_scale', ctypes.c_double), ('log_norm', ctypes.c_double),
                ('shape', ctypes.c_double), ('log_shape', ctypes.c_double)]


class PriorTerm(ctypes.Structure):
    _fields_ = [('component', ctypes.c_int32), ('first_column', ctypes.c_int32),
                ('n_columns', ctypes.c_int32), ('reserved', ctypes.c_int32)]


class PriorRule(ctypes.Structure):
    _fields_ = [('component', ctypes.c_int32), ('a_index', ctypes.c_int32),
                ('b_index', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('a_value', ctypes.c_double), ('b_value', ctypes.c_double)]


# int (*other_columns)(void *user, const double *theta, int64 n, int64 ld, double *logp,
#                      int64 ld_logp)
OTHER_COLUMNS_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_double), ctypes.c_int64,
                                    ctypes.c_int64, ctypes.POINTER(ctypes.c_double),
                                    ctypes.c_int64)


class PriorPlan(ctypes.Structure):
    _fields_ = [('columns', ctypes.POINTER(PriorColumn)),
                ('terms', ctypes.POINTER(PriorTerm)),
                ('rules', ctypes.POINTER(PriorRule)),
                ('n_columns', ctypes.c_int32), ('n_terms', ctypes.c_int32),
                ('n_rules', ctypes.c_int32), ('n_components', ctypes.c_int32),
                ('other_columns', OTHER_COLUMNS_FN), ('user', ctypes.c_void_p)]


class Ensemble(ctypes.Structure):
    _fields_ = [('n_walkers', ctypes.c_int64), ('n_dim', ctypes.c_int64),
                ('a', ctypes.c_double),
                ('pos', ctypes.POINTER(ctypes.c_double)),
                ('lnprob', ctypes.POINTER(ctypes.c_double)),
                ('mt_key', ctypes.POINTER(ctypes.c_uint32)),
                ('mt_pos', ctypes.POINTER(ctypes.c_int32)),
                ('chain', ctypes.POINTER(ctypes.c_double)),
                ('lnprob_chain', ctypes.POINTER(ctypes.c_double)),
                ('chain_len', ctypes.c_int64), ('chain_start', ctypes.c_int64),
                ('thin', ctypes.c_int64),
                ('n_accepted', ctypes.POINTER(ctypes.c_double)),
                ('flags', ctypes.c_int64)]


ENS_SHARDED, ENS_DEVICE = 1, 2


# every symbol include/psfmc_b200.h declares
EXPORTED_SYMBOLS = (
    'psfmc_engine_create', 'psfmc_engine_destroy', 'psfmc_lnlike_batch',
    'psfmc_lnlike_batch_begin', 'psfmc_lnlike_batch_end',
    'psfmc_lnlike_batch_device', 'psfmc_render_batch', 'psfmc_accumulate_batch',
    'psfmc_engine_info', 'psfmc_prior_columns', 'psfmc_prior_sum',
    'psfmc_engine_profile', 'psfmc_engine_profile_read',
    'psfmc_fp32_peak_probe', 'psfmc_last_error', 'psfmc_abi_version',
    'psfmc_peer_create', 'psfmc_peer_connect', 'psfmc_lnlike_batch_exchange',
    'psfmc_peer_gathered', 'psfmc_lnpost_batch', 'psfmc_lnpost_batch_sharded',
    'psfmc_ensemble_run', 'psfmc_rng_fill',
)

_DEFAULT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                             'libpsfmc_b200.so')
_cache = {}


class EngineError(RuntimeError):
    def __init__(self, code, message):
        super(EngineError, self).__init__(
            'psfmc_b200 error {}: {}'.format(code, message))
        self.code = code


def library_path():
    return os.environ.get('PSFMC_B200_LIB', _DEFAULT_PATH)


def load(path=None):
    """Load (once per path) and prototype the shared library."""
    path = path or library_path()
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise ImportError(
            'psfmc_b200: CUDA library not found at {} -- build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a). '
            'There is no CPU fallback.'.format(path))
    lib = ctypes.CDLL(path)
    dbl_p = ctypes.POINTER(ctypes.c_double)
    lib.psfmc_abi_version.restype = ctypes.c_int
    lib.psfmc_abi_version.argtypes = []
    lib.psfmc_last_error.restype = ctypes.c_char_p
    lib.psfmc_last_error.argtypes = []
    lib.psfmc_engine_create.restype = ctypes.c_int
    lib.psfmc_engine_create.argtypes = [ctypes.POINTER(Desc),
                                        ctypes.POINTER(ctypes.c_void_p)]
    lib.psfmc_engine_destroy.restype = None
    lib.psfmc_engine_destroy.argtypes = [ctypes.c_void_p]
    lib.psfmc_lnlike_batch.restype = ctypes.c_int
    lib.psfmc_lnlike_batch.argtypes = [ctypes.c_void_p, dbl_p, ctypes.c_int64,
                                       ctypes.c_int64, dbl_p]
    lib.psfmc_lnlike_batch_begin.restype = ctypes.c_int
    lib.psfmc_lnlike_batch_begin.argtypes = [ctypes.c_void_p, dbl_p, ctypes.c_int64,
                                             ctypes.c_int64, dbl_p]
    lib.psfmc_lnlike_batch_end.restype = ctypes.c_int
    lib.psfmc_lnlike_batch_end.argtypes = [ctypes.c_void_p]
    lib.psfmc_lnlike_batch_device.restype = ctypes.c_int
    lib.psfmc_lnlike_batch_device.argtypes = [
        ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
        ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    lib.psfmc_render_batch.restype = ctypes.c_int
    lib.psfmc_render_batch.argtypes = [ctypes.c_void_p, dbl_p, ctypes.c_int64,
                                       ctypes.c_int64, ctypes.c_uint32, dbl_p]
    lib.psfmc_accumulate_batch.restype = ctypes.c_int
    lib.psfmc_accumulate_batch.argtypes = [ctypes.c_void_p, dbl_p, ctypes.c_int64,
                                           ctypes.c_int64, ctypes.c_uint32, dbl_p]
    lib.psfmc_engine_info.restype = ctypes.c_int
    lib.psfmc_engine_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(Info)]
    lib.psfmc_prior_columns.restype = ctypes.c_int
    lib.psfmc_prior_columns.argtypes = [
        ctypes.POINTER(PriorColumn), ctypes.c_int32, dbl_p, ctypes.c_int64,
        ctypes.c_int64, dbl_p, ctypes.c_int64]
    lib.psfmc_prior_sum.restype = ctypes.c_int
    lib.psfmc_prior_sum.argtypes = [
        dbl_p, ctypes.c_int64, ctypes.c_int64, dbl_p, ctypes.c_int64,
        ctypes.POINTER(PriorTerm), ctypes.c_int32, ctypes.POINTER(PriorRule),
        ctypes.c_int32, ctypes.c_int32, dbl_p]
    lib.psfmc_engine_profile.restype = ctypes.c_int
    lib.psfmc_engine_profile.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.psfmc_engine_profile_read.restype = ctypes.c_int
    lib.psfmc_engine_profile_read.argtypes = [ctypes.c_void_p, dbl_p,
                                              ctypes.POINTER(ctypes.c_int64)]
    lib.psfmc_peer_create.restype = ctypes.c_int
    lib.psfmc_peer_create.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    lib.psfmc_peer_connect.restype = ctypes.c_int
    lib.psfmc_peer_connect.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_void_p]
    lib.psfmc_lnlike_batch_exchange.restype = ctypes.c_int
    lib.psfmc_lnlike_batch_exchange.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
        ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    lib.psfmc_peer_gathered.restype = ctypes.c_int
    lib.psfmc_peer_gathered.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.psfmc_lnpost_batch.restype = ctypes.c_int
    lib.psfmc_lnpost_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(PriorPlan), dbl_p,
                                       ctypes.c_int64, ctypes.c_int64, dbl_p]
    lib.psfmc_lnpost_batch_sharded.restype = ctypes.c_int
    lib.psfmc_lnpost_batch_sharded.argtypes = [ctypes.c_void_p, ctypes.POINTER(PriorPlan),
                                               dbl_p, ctypes.c_int64, ctypes.c_int64, dbl_p]
    lib.psfmc_ensemble_run.restype = ctypes.c_int
    lib.psfmc_ensemble_run.argtypes = [ctypes.c_void_p, ctypes.POINTER(PriorPlan),
                                       ctypes.POINTER(Ensemble), ctypes.c_int64]
    lib.psfmc_rng_fill.restype = ctypes.c_int
    lib.psfmc_rng_fill.argtypes = [ctypes.POINTER(ctypes.c_uint32),
                                   ctypes.POINTER(ctypes.c_int32), ctypes.c_int32,
                                   ctypes.c_int64, ctypes.c_int64, dbl_p]
    lib.psfmc_fp32_peak_probe.restype = ctypes.c_int
    lib.psfmc_fp32_peak_probe.argtypes = [ctypes.c_int32, dbl_p, dbl_p]
    if lib.psfmc_abi_version() != ABI_VERSION:
        raise ImportError('psfmc_b200: ABI version mismatch between {} and the '
                          'Python binding'.format(path))
    _cache[path] = lib
    return lib


def check(lib, code):
    if code != 0:
        raise EngineError(code, lib.psfmc_last_error().decode('utf-8', 'replace'))
