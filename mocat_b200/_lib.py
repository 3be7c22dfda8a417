"""ctypes binding of libmocat_b200.so (C-ABI: include/mocat_b200.h).

The shared object is built in-tree by `mocat_b200/csrc/build.sh` (or `__graft_entry__.build()`).
There is NO CPU fallback: if the library is missing or no sm_100 device is present every compute
entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmocat_b200.so")

MB_MAX_SMALL_DIM = 8
MB_HIST_MAX = 16384
MB_MAX_WORLD = 8
MB_PAIRDIST_ACC_BYTES = 1024 + 2048 * 4
MB_ABC_WS_BYTES = 1024 + 2048 * 4

LIK_RASTRIGIN, LIK_GAUSSIAN, LIK_NONE = 0, 1, 2
MOVE_MALA, MOVE_RW = 0, 1
RESAMPLE_SYSTEMATIC, RESAMPLE_MULTINOMIAL = 0, 1
SSM_LINEAR_GAUSSIAN, SSM_LORENZ96 = 0, 1
PROPOSAL_BOOTSTRAP, PROPOSAL_OPTIMAL, PROPOSAL_ENKF = 0, 1, 2

c_f, c_d, c_i32, c_i64, c_u32, c_u64, c_vp = C.c_float, C.c_double, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_void_p


class Control(C.Structure):
    _fields_ = [(k, c_d) for k in ("wmax", "s1", "s2", "lse", "lse2", "log_ess", "ess", "log_z", "beta",
                                   "alpha_mean", "aux0", "aux1")] + \
               [("nan_count", c_i64), ("alpha_fx", c_i64), ("seed", c_u64)] + \
               [(k, c_i32) for k in ("iter", "resample", "done", "search_iters", "resampled", "pad0")]


class Hist(C.Structure):
    _fields_ = [(k, c_d) for k in ("beta", "ess", "log_z", "alpha_mean", "lse")] + \
               [("resampled", c_i32), ("search_iters", c_i32)]


CONTROL_DTYPE = np.dtype([(k, "<f8") for k in ("wmax", "s1", "s2", "lse", "lse2", "log_ess", "ess", "log_z", "beta",
                                              "alpha_mean", "aux0", "aux1")] +
                         [("nan_count", "<i8"), ("alpha_fx", "<i8"), ("seed", "<u8")] +
                         [(k, "<i4") for k in ("iter", "resample", "done", "search_iters", "resampled", "pad0")])
HIST_DTYPE = np.dtype([(k, "<f8") for k in ("beta", "ess", "log_z", "alpha_mean", "lse")] +
                      [("resampled", "<i4"), ("search_iters", "<i4")])
assert CONTROL_DTYPE.itemsize == C.sizeof(Control) == 144
assert HIST_DTYPE.itemsize == C.sizeof(Hist) == 48


class Target(C.Structure):
    _fields_ = [("kind", c_i32), ("dim", c_i32), ("prior_mean", c_f), ("prior_std", c_f), ("prior_pscale", c_f),
                ("a", c_f), ("mean", c_f * MB_MAX_SMALL_DIM), ("prec_sqrt", c_f * (MB_MAX_SMALL_DIM ** 2))]


class Move(C.Structure):
    _fields_ = [("kind", c_i32), ("mcmc_steps", c_i32), ("leapfrog_steps", c_i32), ("stepsize", c_f)]


class Temper(C.Structure):
    _fields_ = [("max_temperature", c_d), ("ess_retain", c_d), ("ess_resample", c_d), ("tol", c_d),
                ("max_search_iter", c_i32), ("max_iter", c_i32), ("schedule", c_vp), ("schedule_len", c_i32),
                ("pad", c_i32)]


_M2 = c_f * (MB_MAX_SMALL_DIM ** 2)


class SSM(C.Structure):
    _fields_ = [("kind", c_i32), ("dim", c_i32), ("dim_obs", c_i32), ("substeps", c_i32),
                ("m0", c_f * MB_MAX_SMALL_DIM), ("L0", _M2), ("F", _M2), ("LQ", _M2), ("H", _M2), ("Rps", _M2),
                ("lik_const", c_f), ("forcing", c_f), ("dt", c_f), ("q_std", c_f), ("r_std", c_f),
                ("init_mean", c_f), ("init_std", c_f), ("proposal", c_i32)]


class Shard(C.Structure):
    _fields_ = [("rank", c_i32), ("world", c_i32), ("n_local", c_i64), ("n_total", c_i64),
                ("x_peers", c_vp * MB_MAX_WORLD), ("cdf_peers", c_vp * MB_MAX_WORLD), ("totals", c_vp),
                ("anc_peers", c_vp * MB_MAX_WORLD), ("lw_peers", c_vp * MB_MAX_WORLD), ("ws_peers", c_vp * MB_MAX_WORLD)]


class GK(C.Structure):
    _fields_ = [("m", c_i32), ("c", c_f), ("prior_min", c_f), ("prior_max", c_f), ("buffer", c_f), ("data", c_f * 16)]


class Teki(C.Structure):
    """mb_teki: device-resident record of a tempered-EKI ensemble (include/mocat_b200.h)"""
    _fields_ = [("temperature", c_d), ("prev_temperature", c_d), ("alph", c_d), ("ess", c_d),
                ("cov_x", c_d * 16), ("cov_xy", c_d * 64), ("cov_y", c_d * 256), ("cov_y_given_x", c_d * 256),
                ("chol", c_d * 256), ("prec", c_d * 256), ("gain", c_d * 64), ("stds", c_d * 4), ("prior_stds", c_d * 4),
                ("value_nan", c_i64), ("perturb_nan", c_i64), ("iter", c_i32), ("done", c_i32), ("search_iters", c_i32),
                ("pad", c_i32)]


class TekiPrm(C.Structure):
    _fields_ = [("max_temperature", c_d), ("nugget", c_d), ("term_std", c_d), ("ess_threshold", c_d), ("tol", c_d),
                ("max_search_iter", c_i32), ("max_iter", c_i32), ("mode", c_i32), ("schedule_len", c_i32),
                ("schedule", c_vp)]


def fill_matrix(dst, mat):
    """row-major (r, c) -> padded MB_MAX_SMALL_DIM x MB_MAX_SMALL_DIM ctypes array"""
    mat = np.atleast_2d(np.asarray(mat, dtype=np.float64))
    assert mat.shape[0] <= MB_MAX_SMALL_DIM and mat.shape[1] <= MB_MAX_SMALL_DIM
    for r in range(mat.shape[0]):
        for c in range(mat.shape[1]):
            dst[r * MB_MAX_SMALL_DIM + c] = float(mat[r, c])


# name -> (restype, argtypes)   -- must list EVERY function declared in include/mocat_b200.h
SIGNATURES = {
    "mb_last_error": (C.c_char_p, []),
    "mb_abi_version": (C.c_int, []),
    "mb_create": (c_vp, [C.c_int]),
    "mb_destroy": (None, [c_vp]),
    "mb_sm_count": (C.c_int, [c_vp]),
    "mb_lse_ess": (C.c_int, [c_vp, c_vp, c_vp, c_d, c_i64, c_vp, c_vp]),
    "mb_temper_adapt": (C.c_int, [c_vp, c_vp, c_vp, c_i64, C.POINTER(Temper), C.c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "mb_cumsum_lw": (C.c_int, [c_vp, c_vp, c_i64, c_vp, C.c_int, c_vp, c_vp]),
    "mb_cumsum_f32": (C.c_int, [c_vp, c_vp, c_i64, c_d, c_vp, c_vp]),
    "mb_ancestors": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp, c_u64, c_u32, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "mb_gather_state": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "mb_smc_init": (C.c_int, [c_vp, C.POINTER(Target), c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp, c_vp, c_vp, c_u64,
                              c_i64, c_vp, c_vp]),
    "mb_smc_move": (C.c_int, [c_vp, C.POINTER(Target), C.POINTER(Move), c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp,
                              c_vp, c_vp, c_u64, c_i64, c_vp, c_vp, c_vp]),
    "mb_pf_init": (C.c_int, [c_vp, C.POINTER(SSM), c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_u64, c_i64, c_d, c_vp,
                             c_vp, c_vp, c_vp]),
    "mb_pf_step": (C.c_int, [c_vp, C.POINTER(SSM), c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_u64, c_u32,
                             c_i64, c_d, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mb_kalman_filter": (C.c_int, [c_vp, C.POINTER(SSM), c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "mb_backward_sample": (C.c_int, [c_vp, C.POINTER(SSM), c_f, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_u64, c_u32, c_vp,
                                     c_vp, c_vp]),
    "mb_rm_adapt": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_d, c_d, c_d, c_vp, c_vp, c_vp]),
    "mb_stitch_sample": (C.c_int, [c_vp, C.POINTER(SSM), c_f, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_u64, c_u32, c_vp, c_vp]),
    "mb_transition_potential": (C.c_int, [c_vp, C.POINTER(SSM), c_f, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "mb_ancestors_sharded": (C.c_int, [c_vp, c_vp, C.c_int, c_u64, c_u32, c_vp, c_i64, c_vp, c_vp]),
    "mb_strata_count": (C.c_int, [c_i64]),
    "mb_strata_hist": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_u64, c_u32, c_vp, c_vp, C.c_int, c_vp]),
    "mb_strata_reduce": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp, C.c_int, c_vp, c_vp]),
    "mb_ancestors_sorted": (C.c_int, [c_vp, c_vp, c_i64, c_vp, C.c_int, c_vp, c_vp, C.c_int, c_u64, c_u32, c_i64, c_i64,
                                      c_vp, c_i64, c_vp, c_vp]),
    "mb_alloc": (C.c_int, [c_vp, C.c_size_t, C.POINTER(c_vp)]),
    "mb_free": (C.c_int, [c_vp, c_vp]),
    "mb_ipc_get_handle": (C.c_int, [c_vp, c_vp, c_vp]),
    "mb_ipc_open": (C.c_int, [c_vp, c_vp, C.POINTER(c_vp)]),
    "mb_ipc_close": (C.c_int, [c_vp, c_vp]),
    "mb_comm_create": (C.c_int, [c_vp, C.c_int, C.c_int, C.POINTER(c_vp), c_vp]),
    "mb_comm_connect": (C.c_int, [c_vp, c_vp]),
    "mb_comm_destroy": (None, [c_vp]),
    "mb_comm_allgather": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "mb_pf_l96_init": (C.c_int, [c_vp, C.POINTER(SSM), c_vp, c_i64, c_i64, c_vp, c_vp, c_u64, c_i64, c_d, c_vp, c_vp, c_vp,
                                 c_vp]),
    "mb_pf_l96_step": (C.c_int, [c_vp, C.POINTER(SSM), c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_u64, c_u32, c_i64,
                                 c_d, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mb_weighted_moments_rows": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mb_weighted_moment_sums_rows": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mb_gather_rows": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, C.c_int, c_vp]),
    "mb_ksd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, c_f, C.c_int, c_vp, c_vp]),
    "mb_rows_mean_cov": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_f, c_vp, c_vp, c_vp, c_vp]),
    "mb_enkf_analysis": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_u64, C.c_uint32, c_i64, c_vp, c_vp, c_vp, c_vp,
                                   c_vp, c_vp]),
    "mb_rs_workspace_bytes": (C.c_size_t, [c_i64]),
    "mb_rs_tile_sums": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, C.c_int, c_vp]),
    "mb_rs_ancestors": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, C.c_int, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "mb_rs_heavy": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, C.c_int, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "mb_weighted_moments": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mb_quantile": (C.c_int, [c_vp, c_vp, c_i64, c_d, c_vp, c_vp]),
    "mb_colstats": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, c_vp, c_vp]),
    "mb_target_potential_grad": (C.c_int, [c_vp, C.POINTER(Target), c_d, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "mb_workspace_generation": (C.c_uint64, [c_vp]),
    "mb_cond_begin": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "mb_cond_end": (C.c_int, [c_vp, c_vp]),
    "mb_prior_sample": (C.c_int, [c_vp, c_f, c_f, C.c_int, c_i64, c_u64, c_i64, c_vp, c_vp]),
    "mb_logistic_potential_grad": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, c_f, c_f, c_d, c_vp, C.c_int, c_vp, c_vp,
                                             C.c_int, c_vp]),
    "mb_teki_workspace_doubles": (C.c_int, [C.c_int]),
    "mb_teki_init": (C.c_int, [c_vp, C.POINTER(GK), c_vp, c_vp, c_i64, C.c_int, c_u64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "mb_teki_update": (C.c_int, [c_vp, C.POINTER(GK), C.POINTER(TekiPrm), c_vp, c_vp, c_i64, c_u64, c_i64, c_vp, c_vp,
                                 c_vp, c_vp, c_vp, c_vp]),
    "mb_abc_init": (C.c_int, [c_vp, C.POINTER(GK), c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_u64,
                              c_i64, c_vp, c_vp]),
    "mb_abc_move": (C.c_int, [c_vp, C.POINTER(GK), C.c_int, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp,
                              c_vp, c_vp, c_vp, c_vp, c_u64, c_i64, c_vp, c_vp]),
    "mb_abc_adapt": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_d, c_d, c_d,
                               C.c_int, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "mb_abc_move_sharded": (C.c_int, [c_vp, C.POINTER(GK), C.c_int, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                      c_u64, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mb_abc_adapt_stage": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_d, c_d,
                                     c_d, C.c_int, c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "mb_svgd_phi": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, C.c_int, c_vp]),
    "mb_pairdist_bandwidth": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, c_vp]),
    "mb_svgd_phi_rows": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, C.c_int, C.c_int, c_vp]),
    "mb_pairdist_partial": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "mb_pairdist_finish": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "mb_gaussian_kernel": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_f, c_vp, c_vp]),
    "mb_adagrad": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f, c_f, c_vp]),
}


class MocatB200Error(RuntimeError):
    pass


class Library:
    """Loaded shared object + one context per device."""

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise MocatB200Error(
                f"{path} not found: build it with mocat_b200/csrc/build.sh (nvcc, sm_100a). "
                "mocat_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(self.dll, name):
                continue                      # optional groups are checked by tests/test_abi.py
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
        if self.dll.mb_abi_version() != 1:
            raise MocatB200Error("ABI version mismatch")
        self._ctx = {}

    def last_error(self):
        return self.dll.mb_last_error().decode()

    def ctx(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise MocatB200Error("no CUDA device: mocat_b200 runs on B200 (sm_100a) only and has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        if device not in self._ctx:
            h = self.dll.mb_create(int(device))
            if not h:
                raise MocatB200Error(self.last_error())
            self._ctx[device] = h
        return self._ctx[device]

    def call(self, name, *args):
        rc = getattr(self.dll, name)(*args)
        if rc != 0:
            raise MocatB200Error(f"{name} failed ({rc}): {self.last_error()}")


_LIB = None


def get():
    global _LIB
    if _LIB is None:
        _LIB = Library()
    return _LIB


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
