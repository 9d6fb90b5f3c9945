"""ctypes binding of `libscdef_b200.so` (C ABI in `include/scdef_b200.h`).

There is no CPU fallback: if the shared library is missing the import of the training path
fails with a clear error (build it with `python -c "import __graft_entry__ as g; g.build()"`
or `scdef_b200/csrc/build.sh`).
"""
from __future__ import annotations

import ctypes as C
import os

MAX_LAYERS = 8
ABI_VERSION = 8

PASS_LOCAL, PASS_GLOBAL = 1, 2
NOISE_PHILOX, NOISE_EXPLICIT = 0, 1
LIK_AUTO, LIK_SIMT, LIK_TC = 0, 1, 2
FLAG_REUSE_W0_SAMPLES, FLAG_SKIP_LOSS, FLAG_LIK_FIRST = 1, 2, 4

_f32p = C.c_void_p  # device pointers travel as integers


class ModelDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_layers", C.c_int32),
        ("layer_sizes", C.c_int32 * MAX_LAYERS),
        ("n_cells", C.c_int64),
        ("n_genes", C.c_int32),
        ("n_batches", C.c_int32),
        ("use_brd", C.c_int32),
        ("marginalize_alpha", C.c_int32),
        ("supervised_layer", C.c_int32),
        ("alpha_floor", C.c_float),
        ("top_alpha", C.c_float),
        ("brd", C.c_float),
        ("brd_mean", C.c_float),
        ("shrinkage_shape", C.c_float),
        ("shrinkage_rate", C.c_float),
        ("cell_scale_shape", C.c_float),
        ("gene_scale_shape", C.c_float),
        ("X", _f32p),
        ("batch_id", C.c_void_p),
        ("batch_lib_ratio", _f32p),
        ("cell_alpha_factor", _f32p),
        ("gene_ratio", _f32p),
        ("x_lgamma_rowsum", _f32p),
        ("row_slot", C.c_void_p),
        ("w_prior_shape", _f32p * MAX_LAYERS),
        ("w_prior_rate", _f32p * MAX_LAYERS),
        ("w_prior_shape_scalar", C.c_float * MAX_LAYERS),
        ("w_prior_rate_scalar", C.c_float * MAX_LAYERS),
        ("fixed_top_z", _f32p),
    ]


class Blocks(C.Structure):
    _fields_ = [
        ("cell_scale", _f32p),
        ("z", _f32p),
        ("gene_scale", _f32p),
        ("brd", _f32p),
        ("W", _f32p * MAX_LAYERS),
        ("ard", _f32p),
        ("alpha", _f32p),
    ]


class Noise(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("coupling", C.c_int32),
        ("seed", C.c_uint64),
        ("step", C.c_uint64),
        ("eps_cell_scale", _f32p),
        ("eps_z", _f32p),
        ("eps_gene_scale", _f32p),
        ("eps_brd", _f32p),
        ("eps_W", _f32p * MAX_LAYERS),
        ("eps_ard", _f32p),
        ("eps_alpha", _f32p),
        ("local_row_offset", C.c_int64),
        ("local_rows_total", C.c_int64),
    ]


class Call(C.Structure):
    _fields_ = [
        ("pass_", C.c_int32),
        ("num_samples", C.c_int32),
        ("batch", C.c_int32),
        ("lik_impl", C.c_int32),
        ("annealing", C.c_float),
        ("stop_gradients", C.c_float * MAX_LAYERS),
        ("stop_cell_budgets", C.c_float),
        ("stop_gene_budgets", C.c_float),
        ("alpha", C.c_float),
        ("scale", C.c_float),
        ("global_weight", C.c_float),
        ("indices", C.c_void_p),
        ("X_batch", _f32p),
        ("flags", C.c_int32),
        ("reserved", C.c_int32),
    ]


EXPORTS = [
    "scdef_abi_version", "scdef_ctx_create", "scdef_ctx_destroy", "scdef_last_error",
    "scdef_workspace_bytes", "scdef_elbo_grad", "scdef_adam_local", "scdef_adam_global",
    "scdef_posterior", "scdef_row_stats", "scdef_noise_fill", "scdef_profile_enable", "scdef_profile_read",
    "scdef_adam_local_lazy", "scdef_adam_bias_table", "scdef_assignment_counts",
    "scdef_csr_row_stats", "scdef_csr_gather_rows", "scdef_assign_confident", "scdef_init_gamma_block",
    "scdef_u16_row_stats", "scdef_u16_gather_rows", "scdef_count_stats", "scdef_signature_scores",
    "scdef_set_w0_grad_event",
]
LAZY_CATCHUP, LAZY_UPDATE, LAZY_CATCHUP_ALL = 0, 1, 2
PROF_KINDS = ["lik", "adam_local_z", "hier", "sample", "wprior", "finish", "adam_global", "priors", "adam_local_c", "lik_main", "lik_main_global"]

_LIB = None


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libscdef_b200.so")


class ScdefLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library and declare the prototypes (once)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ScdefLibraryError(
            f"{path} not found: the CUDA extension is not built and there is no CPU fallback. "
            "Run scdef_b200/csrc/build.sh (needs nvcc)."
        )
    lib = C.CDLL(path)
    lib.scdef_abi_version.restype = C.c_int
    lib.scdef_ctx_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]
    lib.scdef_ctx_create.restype = C.c_int
    lib.scdef_ctx_destroy.argtypes = [C.c_void_p]
    lib.scdef_ctx_destroy.restype = None
    lib.scdef_last_error.argtypes = [C.c_void_p]
    lib.scdef_last_error.restype = C.c_char_p
    lib.scdef_workspace_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.scdef_workspace_bytes.restype = C.c_size_t
    lib.scdef_elbo_grad.argtypes = [
        C.c_void_p, C.POINTER(Call), C.POINTER(Blocks), C.POINTER(Noise), C.c_void_p,
        C.POINTER(Blocks), C.c_void_p, C.c_size_t, C.c_void_p,
    ]
    lib.scdef_elbo_grad.restype = C.c_int
    lib.scdef_adam_local.argtypes = [
        C.c_void_p, C.POINTER(Blocks), C.POINTER(Blocks), C.POINTER(Blocks), C.POINTER(Blocks),
        C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
        C.POINTER(C.c_int32), C.c_void_p,
    ]
    lib.scdef_adam_local.restype = C.c_int
    lib.scdef_adam_local_lazy.argtypes = [
        C.c_void_p, C.c_int, C.POINTER(Blocks), C.POINTER(Blocks), C.POINTER(Blocks), C.POINTER(Blocks),
        C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
        C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
    ]
    lib.scdef_adam_local_lazy.restype = C.c_int
    lib.scdef_adam_bias_table.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_int]
    lib.scdef_adam_bias_table.restype = C.c_int
    lib.scdef_adam_global.argtypes = [
        C.c_void_p, C.POINTER(Blocks), C.POINTER(Blocks), C.POINTER(Blocks), C.POINTER(Blocks),
        C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p,
    ]
    lib.scdef_adam_global.restype = C.c_int
    lib.scdef_posterior.argtypes = [C.c_void_p, C.POINTER(Blocks), C.POINTER(Blocks), C.POINTER(Blocks), C.c_void_p]
    lib.scdef_posterior.restype = C.c_int
    lib.scdef_assignment_counts.argtypes = [C.c_void_p, C.POINTER(Blocks), C.c_void_p, C.c_void_p]
    lib.scdef_assignment_counts.restype = C.c_int
    lib.scdef_csr_row_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scdef_csr_row_stats.restype = C.c_int
    lib.scdef_csr_gather_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.scdef_csr_gather_rows.restype = C.c_int
    lib.scdef_u16_row_stats.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scdef_u16_row_stats.restype = C.c_int
    lib.scdef_u16_gather_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.scdef_u16_gather_rows.restype = C.c_int
    lib.scdef_count_stats.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scdef_count_stats.restype = C.c_int
    lib.scdef_assign_confident.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_float, C.c_uint64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
    lib.scdef_assign_confident.restype = C.c_int
    lib.scdef_signature_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scdef_signature_scores.restype = C.c_int
    lib.scdef_set_w0_grad_event.argtypes = [C.c_void_p, C.c_void_p]
    lib.scdef_set_w0_grad_event.restype = C.c_int
    lib.scdef_init_gamma_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                           C.c_float, C.c_float, C.c_uint64, C.c_int32, C.c_int64, C.c_void_p]
    lib.scdef_init_gamma_block.restype = C.c_int
    lib.scdef_row_stats.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scdef_row_stats.restype = C.c_int
    lib.scdef_noise_fill.argtypes = [C.c_void_p, C.POINTER(Noise), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.scdef_noise_fill.restype = C.c_int
    lib.scdef_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.scdef_profile_enable.restype = C.c_int
    lib.scdef_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.scdef_profile_read.restype = C.c_int
    got = lib.scdef_abi_version()
    if got != ABI_VERSION:
        raise ScdefLibraryError(f"ABI mismatch: library {got}, binding {ABI_VERSION}; rebuild the extension")
    _LIB = lib
    return lib
