"""ctypes binding of libcarc_b200.so (include/carc_b200.h).

There is no CPU fallback: if the shared library has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C carcassonne_b200/csrc``) importing this module fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CARC_B200_LIB: another build of the same library (A/B measurements of a kernel change); default: the in-tree build
LIB_PATH = os.environ.get("CARC_B200_LIB") or os.path.join(_HERE, "libcarc_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "carcassonne_b200: %s is missing -- build the CUDA library first (make -C carcassonne_b200/csrc); "
        "there is no CPU fallback." % LIB_PATH
    )

lib = C.CDLL(LIB_PATH)

c_i64 = C.c_int64
c_int = C.c_int
c_vp = C.c_void_p
c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)

#: every symbol include/carc_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "carc_version": (c_int, []),
    "carc_last_error": (C.c_char_p, []),
    "carc_dmma_peak": (c_int, [c_int, c_dp, c_vp]),
    "carc_dmma_rate": (c_int, [c_int, c_int, c_int, c_dp, c_vp]),
    "carc_fp64_mix_rate": (c_int, [c_int, c_int, c_int, c_int, c_dp, c_dp, c_vp]),
    "carc_malloc": (c_int, [C.POINTER(c_vp), C.c_size_t]),
    "carc_free": (c_int, [c_vp]),
    "carc_malloc_host": (c_int, [C.POINTER(c_vp), C.c_size_t]),
    "carc_free_host": (c_int, [c_vp]),
    "carc_memcpy_h2d": (c_int, [c_vp, c_vp, C.c_size_t, c_vp]),
    "carc_memcpy_d2h": (c_int, [c_vp, c_vp, C.c_size_t, c_vp]),
    "carc_stream_synchronize": (c_int, [c_vp]),
    "carc_permute": (c_int, [c_vp, c_vp, c_int, c_i64p, c_i32p, c_int, c_int, c_vp]),
    "carc_axpby": (c_int, [c_i64, c_dp, c_vp, c_dp, c_vp, c_int, c_vp]),
    "carc_mode_product": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "carc_mul": (c_int, [c_i64, c_vp, c_vp, c_vp]),
    "carc_dotc": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp]),
    "carc_dotu": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp]),
    "carc_sumsq": (c_int, [c_i64, c_vp, c_vp, c_vp]),
    "carc_count_nonfinite": (c_int, [c_i64, c_vp, c_vp, c_vp]),
    "carc_operator_num_groups": (c_int, [c_vp]),
    "carc_operator_executed_flops": (C.c_double, [c_vp]),
    "carc_comm_create": (c_int, [C.POINTER(c_vp), c_int, c_int, c_i64]),
    "carc_comm_local_handles": (c_int, [c_vp, c_vp]),
    "carc_comm_connect": (c_int, [c_vp, c_vp]),
    "carc_comm_allreduce": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "carc_comm_status": (c_int, [c_vp, C.POINTER(c_int)]),
    "carc_comm_destroy": (c_int, [c_vp]),
    "carc_operator_set_comm": (c_int, [c_vp, c_vp]),
    "carc_operator_create_dense": (c_int, [C.POINTER(c_vp), c_vp, c_i64]),
    "carc_operator_dimension": (c_i64, [c_vp]),
    "carc_lu_inverse_blocks_elems": (c_i64, [c_int]),
    "carc_lu_invert_diagonal_blocks": (c_int, [c_vp, c_int, c_vp, c_vp]),
    "carc_lu_solve_blocks": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "carc_relax": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, C.c_double, c_int, C.c_double, c_int, c_int, c_dp, c_vp]),
    "carc_gmres": (c_int, [c_vp, c_vp, c_vp, C.c_double, c_int, c_int, C.POINTER(c_int), c_dp, c_vp]),
    "carc_cg": (c_int, [c_vp, c_vp, c_vp, C.c_double, c_int, C.POINTER(c_int), c_dp, c_vp]),
    "carc_lu_factor": (c_int, [c_vp, c_int, c_vp, C.POINTER(c_int), c_vp]),
    "carc_cholesky_factor_as_lu": (c_int, [c_vp, c_int, c_vp, C.c_double, C.POINTER(c_int), c_vp]),
    "carc_lu_solve": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp]),
    "carc_zgemm_hermitian": (c_int, [c_int, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "carc_index_table": (c_int, [c_int, c_i64p, c_i64p, c_vp, c_vp]),
    "carc_zgemm_tab": (c_int, [c_int, c_int, c_i64, c_i64, c_i64, c_dp, c_vp, c_i64, c_vp, c_i64, c_dp, c_vp,
                               c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "carc_zgemm": (c_int, [c_int, c_int, c_i64, c_i64, c_i64, c_dp, c_vp, c_i64, c_vp, c_i64, c_dp, c_vp,
                           c_i64p, c_i64p, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "carc_operator_create": (c_int, [C.POINTER(c_vp), c_int, c_int, c_int, c_int, c_int]),
    "carc_operator_add_term": (c_int, [c_vp, c_vp, c_vp, c_i64, c_dp]),
    "carc_operator_finalize": (c_int, [c_vp]),
    "carc_operator_apply": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "carc_operator_set_path": (c_int, [c_vp, c_int]),
    "carc_operator_path": (c_int, [c_vp]),
    "carc_stage3f_profile_read": (c_int, [c_vp]),
    "carc_stage3_path": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_i64, c_int]),
    "carc_stage3_describe_stars": (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "carc_stage3f_describe": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_i64, c_vp, c_int]),
    "carc_absorb_side_into_corner": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp]),
    "carc_double_layer_center": (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "carc_absorb_center_into_side": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "carc_form_stage1": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "carc_form_stage2": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "carc_stage3_form_matrix": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_int, c_vp, c_int, c_vp]),
    "carc_normalize_axis": (c_int, [c_vp, c_vp, c_int, c_int, c_int, C.c_double, c_vp, c_vp, c_vp, c_vp]),
    "carc_product_compressor": (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_int, C.c_double, c_vp, c_vp, c_vp,
                                        c_vp]),
    "carc_operator_num_terms": (c_int, [c_vp]),
    "carc_operator_cost_of_multiply": (c_i64, [c_vp]),
    "carc_operator_destroy": (c_int, [c_vp]),
    "carc_qr": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "carc_svd_small": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "carc_normalizer_matrices": (c_int, [c_vp, c_vp, c_vp, c_int, C.c_double, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "carc_stage3_matvec_host": (c_int, [c_int, C.POINTER(c_vp), C.POINTER(c_vp), c_i64p, C.POINTER(c_dp),
                                        c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

# error codes (include/carc_b200.h)
OK, ERR_CUDA, ERR_DIMENSION_MISMATCH, ERR_RANK, ERR_VALUE, ERR_RELAX_FAILED, ERR_INVARIANT, ERR_NO_CONVERGENCE, \
    ERR_UNSUPPORTED, ERR_EXCHANGE = range(10)
OP_N, OP_T, OP_C, OP_J = range(4)


class CarcError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libcarc_b200 error {}: {}".format(code, message))
        self.code = code


def check(rc):
    if rc != 0:
        msg = lib.carc_last_error().decode("utf-8", "replace")
        if rc == ERR_VALUE:
            raise ValueError(msg)
        raise CarcError(rc, msg)


def cplx2(z):
    z = complex(z)
    return (C.c_double * 2)(z.real, z.imag)
