"""ctypes binding of libwoltka_b200.so (the C-ABI in include/woltka_b200.h).

There is no CPU fallback: if the shared library is missing the import of any
compute entry point raises, and if no sm_100 device is present `wk_create`
fails with the library's own message.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libwoltka_b200.so')

UNITS = 720720
MAX_ENTRIES = 8

KIND_NONE, KIND_FREE, KIND_RANK, KIND_NONE_ID = 0, 1, 2, 3
F_UNIQ, F_ABOVE, F_MAJOR, F_UNASSIGNED, F_SIZES = 1, 2, 4, 8, 16

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/woltka_b200.h one to one
SIGNATURES = {
    'wk_last_error': (C.c_char_p, []),
    'wk_abi_version': (C.c_int, []),
    'wk_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'wk_create': (C.c_int, [C.c_int, C.POINTER(_vp)]),
    'wk_destroy': (C.c_int, [_vp]),
    'wk_set_stream': (C.c_int, [_vp, _vp]),
    'wk_sync': (C.c_int, [_vp]),
    'wk_launch_count': (C.c_int64, [_vp]),
    'wk_set_tuning': (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    'wk_set_option': (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    'wk_last_kernel': (C.c_char_p, [_vp]),
    'wk_release_cached_memory': (C.c_int, [C.c_int]),
    'wk_host_alloc': (C.c_int, [C.POINTER(_vp), C.c_int64]),
    'wk_host_free': (C.c_int, [_vp]),
    'wk_set_tree': (C.c_int, [_vp, _vp, C.c_int32, C.c_int32]),
    'wk_set_plan': (C.c_int, [_vp, _vp, C.c_int32, C.c_uint32, C.c_double,
                              C.c_int32, C.c_int64]),
    'wk_resize_counts': (C.c_int, [_vp, C.c_int32, C.c_int64]),
    'wk_set_subjects': (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    'wk_classify_chunk': (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp,
                                    C.c_int64, C.c_int32]),
    'wk_classify_packed': (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int64, _vp,
                                     _vp, C.c_int64, C.c_int32]),
    'wk_classify_packed_bits': (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int64, _vp,
                                     _vp, C.c_int64, C.c_int32]),
    'wk_classify_device': (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp,
                                     C.c_int64, C.c_int32]),
    'wk_ordinal_set_genes': (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int32,
                                       C.c_int64]),
    'wk_ordinal_chunk': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64,
                                   C.c_double, _vp, _vp, C.c_int64,
                                   C.c_int32]),
    'wk_ordinal_device': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64,
                                    C.c_double, _vp, _vp, C.c_int64,
                                    C.c_int32]),
    'wk_ordinal_fetch_pairs': (C.c_int, [_vp, _i64p, _vp, _vp, C.c_int64]),
    'wk_fetch_counts': (C.c_int, [_vp, _vp]),
    'wk_fetch_overflow': (C.c_int, [_vp, _i64p, _vp, _vp, _vp, C.c_int64]),
    'wk_fetch_strata': (C.c_int, [_vp, _i64p, _vp, _vp, _vp, _vp, _vp,
                                  C.c_int64]),
    'wk_reset_counts': (C.c_int, [_vp]),
    'wk_strata_export_device': (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), _i64p]),
    'wk_strata_import_device': (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    'wk_strata_reserve': (C.c_int, [_vp, C.c_int64]),
    'wk_reset_strata': (C.c_int, [_vp]),
    'wk_overflow_export_device': (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), _i64p]),
    'wk_overflow_import_device': (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int]),
    'wk_set_assign_output': (C.c_int, [_vp, C.c_int]),
    'wk_fetch_assignments': (C.c_int, [_vp, _vp, C.c_int64]),
    'wk_counts_device': (C.c_int, [_vp, C.POINTER(_vp), _i64p]),
    'wk_parse_sam': (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _i64p, _i64p,
                               _i32p, _i32p]),
    'wk_parse_text': (C.c_int, [_vp, _vp, C.c_int64, C.c_int, C.c_int, _i64p,
                                _i64p, _i32p, _i32p]),
    'wk_parse_fetch_names': (C.c_int, [_vp, C.c_int, C.c_int32, C.c_int32, _vp,
                                       C.c_int64, _i64p, _vp]),
    'wk_parse_fetch_columns': (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    'wk_classify_parsed': (C.c_int, [_vp, _vp, C.c_int32, C.c_int32]),
    'wk_parse_block': (C.c_int, [_vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int,
                                 _i64p, _i64p, _i64p, _i32p, _i32p]),
    'wk_parse_fetch_coords': (C.c_int, [_vp, _vp, _vp, _vp]),
    'wk_parse_options': (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, C.c_int32,
                                   C.c_int]),
    'wk_ordinal_parsed': (C.c_int, [_vp, _vp, C.c_int32, C.c_double, _vp,
                                    C.c_int32, C.c_int32]),
    'wk_cover_add': (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64]),
    'wk_cover_merge': (C.c_int, [_vp, _i64p]),
    'wk_cover_fetch': (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64]),
}

_lib = None


class WoltkaB200Error(RuntimeError):
    """Raised for any non-zero status of the C-ABI."""

    def __init__(self, code, msg):
        super().__init__(f'[wk status {code}] {msg}')
        self.code = code


def load():
    """Load the shared library (once) and declare every signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` (nvcc, '
            'sm_100a). woltka_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = ABI mismatch, be loud
        fn.restype = res
        fn.argtypes = args
    if lib.wk_abi_version() != 1:
        raise ImportError('libwoltka_b200.so has an unexpected ABI version')
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise WoltkaB200Error(status, load().wk_last_error().decode())
