"""ctypes binding of libscoary_b200.so (include/scoary_b200.h).

There is no fallback: if the shared library is missing this module raises, and
if no B200 is present sb_create() fails and Engine() raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCOARY_B200_LIB") or os.path.join(_HERE, "libscoary_b200.so")   # override: kernel experiments

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f64p = ctypes.POINTER(ctypes.c_double)


class SbStats(ctypes.Structure):
    _fields_ = [
        ("kernel_launches", ctypes.c_int64),
        ("h2d_bytes", ctypes.c_int64),
        ("d2h_bytes", ctypes.c_int64),
        ("tests_contingency", ctypes.c_int64),
        ("tests_walks", ctypes.c_int64),
        ("ms_pack", ctypes.c_double),
        ("ms_fisher", ctypes.c_double),
        ("ms_shuffle", ctypes.c_double),
        ("ms_walk", ctypes.c_double),
        ("ms_permute", ctypes.c_double),
        ("ms_reduce", ctypes.c_double),
        ("launches_permute", ctypes.c_int64),
        ("sm_count", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("calls_transposed", ctypes.c_int64),
        ("ms_epilogue", ctypes.c_double),
    ]


# name -> (restype, argtypes); every symbol include/scoary_b200.h declares
SIGNATURES = {
    "sb_version": (ctypes.c_int, []),
    "sb_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "sb_destroy": (None, [ctypes.c_void_p]),
    "sb_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "sb_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "sb_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "sb_set_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "sb_set_permute_mode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "sb_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(SbStats)]),
    "sb_stats_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "sb_set_genes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]),
    "sb_set_genes_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]),
    "sb_set_trait": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "sb_set_tree": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "sb_contingency_fisher": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sb_contingency_fisher_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sb_contingency_fisher_multi": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p]),
    "sb_contingency_fisher_multi_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                                          ctypes.c_void_p, ctypes.c_void_p]),
    "sb_pairwise": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "sb_pairwise_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "sb_permute": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                  ctypes.c_uint64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p]),
    "sb_permute_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                         ctypes.c_uint64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "sb_permute_range": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                        ctypes.c_int32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]),
    "sb_permute_range_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
                                               ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64, ctypes.c_void_p,
                                               ctypes.c_void_p]),
    "sb_debug_shuffled_labels": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64, ctypes.c_void_p]),
    "sb_upgma": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "sb_debug_pipe_rates": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "sb_adjust_pvalues": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "sb_adjust_pvalues_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.POINTER(ctypes.c_int64)]),
    "sb_binom_two_sided": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "sb_csv_row_starts": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p]),
    "sb_csv_pack_rows": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                          ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "sb_csv_scan_rows": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char, ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.c_void_p]),
    "sb_csv_free": (None, [ctypes.c_void_p]),
    "sb_csv_gather_fields": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "sb_vcf_line_starts": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "sb_vcf_count_rows": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p,
                                           ctypes.c_int64, ctypes.c_void_p]),
    "sb_vcf_pack_rows": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32,
                                          ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "sb_debug_compile_tree": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "sb_debug_compile_tree2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                              ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "sb_int32_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]),
}

_lib = None


def load():
    """Load libscoary_b200.so and attach signatures.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "scoary_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C scoary_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
