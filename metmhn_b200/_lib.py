"""ctypes binding of libmetmhn_b200.so (include/metmhn_b200.h).

There is deliberately no fallback: if the shared library is missing, or the machine has no CUDA
device, every compute call raises.  The oracle under /oracle is test infrastructure and is never
imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMH_LIB") or os.path.join(_HERE, "libmetmhn_b200.so")   # MMH_LIB: A/B builds of the same library

MMH_OK, MMH_EINVAL, MMH_ECUDA, MMH_ENOMEM, MMH_ETOOLARGE, MMH_ENCCL = 0, -1, -2, -3, -4, -5
NCCL_ID_BYTES = 128
MAX_MUT = 28

EXPORTS = ("mmh_create", "mmh_value_grad", "mmh_value", "mmh_eval_weighted", "mmh_eval_device", "mmh_sync",
           "mmh_set_profile", "mmh_per_patient",
           "mmh_stats", "mmh_destroy", "mmh_last_error", "mmh_measure_fp64_tflops",
           "mmh_nccl_unique_id", "mmh_comm_init", "mmh_comm_destroy",
           "mmh_multi_create", "mmh_multi_value_grad", "mmh_multi_value", "mmh_multi_destroy", "mmh_simulate", "mmh_learn", "mmh_per_patient_grads",
           "mmh_row_cost")


class Stats(C.Structure):
    _fields_ = [("n_dat", C.c_int64), ("n_em", C.c_int64), ("n_spaces", C.c_int64), ("n_chunks", C.c_int64),
                ("n_launches", C.c_int64), ("states_value_grad", C.c_double), ("alg_bytes", C.c_double),
                ("alg_flops", C.c_double), ("exec_fma", C.c_double), ("last_ms", C.c_double),
                ("scratch_bytes", C.c_double), ("k_hist", (C.c_int64 * 64) * 4), ("class_ms", C.c_double * 8)]


class MetMHNError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"metmhn_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MetMHNError(MMH_ECUDA, f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; "
                                     "g.build()'` or `make -C metmhn_b200/csrc`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    L.mmh_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64]
    L.mmh_value_grad.argtypes = [C.c_void_p, dp, C.c_double, dp, dp]
    L.mmh_value.argtypes = [C.c_void_p, dp, C.c_double, dp]
    L.mmh_eval_weighted.argtypes = [C.c_void_p, dp, C.c_double, C.c_double, C.c_int, dp, C.c_void_p]
    L.mmh_eval_device.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]
    L.mmh_sync.argtypes = [C.c_void_p]
    L.mmh_set_profile.argtypes = [C.c_void_p, C.c_int]
    L.mmh_per_patient.argtypes = [C.c_void_p, dp, dp]
    L.mmh_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.mmh_destroy.argtypes = [C.c_void_p]
    L.mmh_destroy.restype = None
    L.mmh_last_error.restype = C.c_char_p
    L.mmh_measure_fp64_tflops.argtypes = [C.c_int, dp]
    L.mmh_nccl_unique_id.argtypes = [C.c_char_p]
    L.mmh_comm_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    L.mmh_comm_destroy.argtypes = [C.c_void_p]
    L.mmh_multi_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                   C.POINTER(C.c_int), C.c_int, C.c_int64]
    L.mmh_multi_value_grad.argtypes = [C.c_void_p, dp, C.c_double, dp, dp]
    L.mmh_multi_value.argtypes = [C.c_void_p, dp, C.c_double, dp]
    L.mmh_multi_destroy.argtypes = [C.c_void_p]
    L.mmh_multi_destroy.restype = None
    L.mmh_simulate.argtypes = [C.c_int, dp, C.c_int64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.mmh_per_patient_grads.argtypes = [C.c_void_p, dp, C.c_int64, C.c_int64, dp, dp]
    L.mmh_learn.argtypes = [C.c_void_p, dp, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_double, dp, dp,
                            C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    for name in EXPORTS:
        if name not in ("mmh_destroy", "mmh_last_error", "mmh_multi_destroy", "mmh_row_cost"):
            getattr(L, name).restype = C.c_int
    L.mmh_row_cost.restype = C.c_double
    L.mmh_row_cost.argtypes = [C.c_void_p, C.c_int]
    _lib = L
    return L


def check(rc):
    if rc != MMH_OK:
        raise MetMHNError(rc, lib().mmh_last_error().decode("utf-8", "replace"))


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Handle:
    """Device-resident, preprocessed dataset (one per GPU / shard)."""

    def __init__(self, dat, device=0, chunk_bytes=0):
        dat = np.ascontiguousarray(np.asarray(dat), dtype=np.int8)
        if dat.ndim != 2 or dat.shape[1] < 5 or (dat.shape[1] - 3) % 2:
            raise MetMHNError(MMH_EINVAL, "dat must be (n_dat, 2n+3) int8")
        self.n_mut = (dat.shape[1] - 3) // 2
        self.n_tot = self.n_mut + 1
        self.npar = self.n_tot * (self.n_tot + 2)
        self.n_dat = dat.shape[0]
        self.device = device
        self._h = C.c_void_p()
        check(lib().mmh_create(C.byref(self._h), self.n_mut, dat.ctypes.data_as(C.c_void_p), dat.shape[0],
                               dat.shape[1], device, int(chunk_bytes)))

    def _params(self, params):
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())
        if p.shape[0] != self.npar:
            raise MetMHNError(MMH_EINVAL, f"params must have {(self.npar)} entries")
        return p

    def value_grad(self, params, perc_met):
        p = self._params(params)
        score = C.c_double()
        grad = np.empty(self.npar)
        check(lib().mmh_value_grad(self._h, _dptr(p), float(perc_met), C.byref(score), _dptr(grad)))
        return score.value, grad

    def value(self, params, perc_met):
        p = self._params(params)
        score = C.c_double()
        check(lib().mmh_value(self._h, _dptr(p), float(perc_met), C.byref(score)))
        return score.value

    def eval_weighted(self, params, w_type0, w_other, want_grad=True, out_dev_ptr=None, to_host=True):
        p = self._params(params)
        out = np.empty(self.npar + 1) if to_host else None
        check(lib().mmh_eval_weighted(self._h, _dptr(p), float(w_type0), float(w_other), int(bool(want_grad)),
                                      _dptr(out) if to_host else None,
                                      C.c_void_p(out_dev_ptr) if out_dev_ptr else None))
        if not to_host:
            return None
        return (out[0], out[1:]) if want_grad else (out[0], None)

    def eval_device(self, d_params_ptr, w_type0, w_other, d_out_ptr, want_grad=True):
        """Queue one evaluation with device-resident parameters / result (raw device pointers)."""
        check(lib().mmh_eval_device(self._h, C.c_void_p(d_params_ptr), float(w_type0), float(w_other),
                                    int(bool(want_grad)), C.c_void_p(d_out_ptr)))

    def sync(self):
        check(lib().mmh_sync(self._h))

    def set_profile(self, on):
        check(lib().mmh_set_profile(self._h, int(bool(on))))

    def per_patient(self, params):
        p = self._params(params)
        out = np.zeros(max(self.n_dat, 1))
        check(lib().mmh_per_patient(self._h, _dptr(p), _dptr(out)))
        return out[: self.n_dat]

    def per_patient_grads(self, params, first_row=0, n_rows=None):
        """(logp[n_rows], grads[n_rows, npar]) of single rows (test hook, mmh_per_patient_grads)."""
        p = self._params(params)
        n_rows = self.n_dat - first_row if n_rows is None else int(n_rows)
        lp = np.zeros(max(n_rows, 1))
        g = np.zeros((max(n_rows, 1), self.npar))
        check(lib().mmh_per_patient_grads(self._h, _dptr(p), int(first_row), n_rows, _dptr(lp), _dptr(g)))
        return lp[:n_rows], g[:n_rows]

    def stats(self):
        s = Stats()
        check(lib().mmh_stats(self._h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in Stats._fields_ if k not in ("k_hist", "class_ms")}
        d["class_ms"] = dict(zip(("setup", "solve_fwd", "solve_adj", "stats", "finish", "other", "pfin"), list(s.class_ms)[:7]))
        d["k_hist"] = {t: {k: int(s.k_hist[t][k]) for k in range(64) if s.k_hist[t][k]} for t in range(4)}
        return d

    def learn(self, x0, perc_met, w_penal, eps=1e-5, max_iter=100000, ftol=1e-4):
        """The whole L-BFGS fit inside the library (mmh_learn): returns (x, f, iterations, evaluations)."""
        p = self._params(x0)
        x = np.empty(self.npar)
        f = C.c_double()
        it, ev = C.c_int64(), C.c_int64()
        check(lib().mmh_learn(self._h, _dptr(p), float(perc_met), float(w_penal), float(eps), int(max_iter), float(ftol),
                              _dptr(x), C.byref(f), C.byref(it), C.byref(ev)))
        return x, f.value, it.value, ev.value

    def comm_init(self, unique_id: bytes, nranks: int, rank: int):
        """Attach an NCCL communicator (collective over all ranks): from now on every evaluation on this handle
        ends with an in-library all-reduce of the result on the handle's stream."""
        if len(unique_id) != NCCL_ID_BYTES:
            raise MetMHNError(MMH_EINVAL, "the NCCL unique id has 128 bytes")
        check(lib().mmh_comm_init(self._h, unique_id, int(nranks), int(rank)))

    def comm_destroy(self):
        if self._h:
            check(lib().mmh_comm_destroy(self._h))

    def close(self):
        if self._h:
            lib().mmh_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    """128 opaque bytes from ncclGetUniqueId; rank 0 creates them and ships them to the other ranks."""
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    check(lib().mmh_nccl_unique_id(buf))
    return buf.raw


class MultiHandle:
    """One process, several GPUs (mmh_multi_*): rows partitioned by the cost model, one in-library all-reduce."""

    def __init__(self, dat, devices, chunk_bytes=0):
        dat = np.ascontiguousarray(np.asarray(dat), dtype=np.int8)
        if dat.ndim != 2 or dat.shape[1] < 5 or (dat.shape[1] - 3) % 2:
            raise MetMHNError(MMH_EINVAL, "dat must be (n_dat, 2n+3) int8")
        self.n_mut = (dat.shape[1] - 3) // 2
        self.n_tot = self.n_mut + 1
        self.npar = self.n_tot * (self.n_tot + 2)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        self._m = C.c_void_p()
        check(lib().mmh_multi_create(C.byref(self._m), self.n_mut, dat.ctypes.data_as(C.c_void_p), dat.shape[0],
                                     dat.shape[1], devs, len(devices), int(chunk_bytes)))

    def value_grad(self, params, perc_met):
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())
        if p.shape[0] != self.npar:
            raise MetMHNError(MMH_EINVAL, f"params must have {self.npar} entries")
        score = C.c_double()
        grad = np.empty(self.npar)
        check(lib().mmh_multi_value_grad(self._m, _dptr(p), float(perc_met), C.byref(score), _dptr(grad)))
        return score.value, grad

    def value(self, params, perc_met):
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())
        score = C.c_double()
        check(lib().mmh_multi_value(self._m, _dptr(p), float(perc_met), C.byref(score)))
        return score.value

    def close(self):
        if self._m:
            lib().mmh_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def simulate(log_theta, log_d_p, log_d_m, n_sim: int, seed: int = 0, device: int = 0):
    """GPU Gillespie sampler (`metmhn/simulations.py:110-147` `simulate_dat`): returns (geno int8 (n_sim, 2n+1), order int8
    (n_sim,)) with order 1 = PT diagnosed first, 2 = MT first, 0 = never seeded."""
    th = np.asarray(log_theta, dtype=np.float64)
    n_tot = th.shape[0]
    p = np.ascontiguousarray(np.concatenate([th.ravel(), np.asarray(log_d_p, dtype=np.float64).ravel(),
                                             np.asarray(log_d_m, dtype=np.float64).ravel()]))
    if th.shape != (n_tot, n_tot) or p.shape[0] != n_tot * (n_tot + 2):
        raise MetMHNError(MMH_EINVAL, "log_theta must be (n+1, n+1), log_d_p / log_d_m (n+1,)")
    geno = np.zeros((int(n_sim), 2 * (n_tot - 1) + 1), dtype=np.int8)
    order = np.zeros(int(n_sim), dtype=np.int8)
    check(lib().mmh_simulate(n_tot - 1, _dptr(p), int(n_sim), int(seed) & 0xFFFFFFFFFFFFFFFF, int(device),
                             geno.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.c_void_p)))
    return geno, order


def measure_fp64_tflops(device=0):
    v = C.c_double()
    check(lib().mmh_measure_fp64_tflops(device, C.byref(v)))
    return v.value
