"""ctypes binding of libdicegp.so (the C ABI in include/dicegp.h).

The product path has no CPU fallback: if the shared library is missing or no CUDA device
is present, every entry point raises -- loudly -- instead of computing on the host.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DICEGP_LIB") or os.path.join(_HERE, "libdicegp.so")

ABI_VERSION = 2

CHAIN_RUNNING, CHAIN_CONVERGED, CHAIN_EXHAUSTED, CHAIN_NOMEM = 0, 1, 2, 3

# every symbol include/dicegp.h declares (tests check that the library exports all of them)
SYMBOLS = (
    "dicegp_abi_version", "dicegp_create", "dicegp_destroy", "dicegp_last_error",
    "dicegp_get_limits", "dicegp_upload_genes", "dicegp_set_chains",
    "dicegp_main_chains_from_prerun", "dicegp_run_host_stream", "dicegp_run_device_stream",
    "dicegp_chain_status", "dicegp_download_trace", "dicegp_posterior_summary",
    "dicegp_last_run_stats", "dicegp_set_trace_budget", "dicegp_set_prerun_chains",
    "dicegp_timer_start", "dicegp_timer_stop", "dicegp_counters", "dicegp_eval_loglik",
    "dicegp_dump_stream", "dicegp_work_counters",
)


class DiceGPError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("libdicegp error %d: %s" % (code, text))
        self.code = code


class ChainDesc(C.Structure):
    """struct dicegp_chain_desc"""
    _fields_ = [
        ("gene", C.c_int32), ("t0", C.c_int32), ("Tc", C.c_int32),
        ("M", C.c_int32), ("initial", C.c_int32), ("gap", C.c_int32),
        ("kinv_off", C.c_int64), ("prop_factor_off", C.c_int64),
        ("jump_scale_off", C.c_int64), ("jump_const_off", C.c_int64),
        ("y0_off", C.c_int64), ("prior_const", C.c_double),
    ]


class Limits(C.Structure):
    _fields_ = [("max_isoforms", C.c_int32), ("max_times", C.c_int32),
                ("abi_version", C.c_int32), ("sm_count", C.c_int32),
                ("max_latent", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def load():
    """dlopen libdicegp.so and declare prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libdicegp.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C diceseq_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pi32, pi64, pd = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
    lib.dicegp_abi_version.restype = C.c_int
    lib.dicegp_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.dicegp_destroy.argtypes = [vp]
    lib.dicegp_last_error.argtypes = [vp]
    lib.dicegp_last_error.restype = C.c_char_p
    lib.dicegp_get_limits.argtypes = [vp, C.POINTER(Limits)]
    lib.dicegp_upload_genes.argtypes = [vp, i32, i32, pi32, pi32, pd, i64, pd, i64]
    lib.dicegp_set_chains.argtypes = [vp, i32, C.POINTER(ChainDesc), pd, i64]
    lib.dicegp_main_chains_from_prerun.argtypes = [vp, i32, i32, i32, dbl, pd, pd, dbl, pd]
    lib.dicegp_run_host_stream.argtypes = [vp, i32, pi32, i32, pd, pi64, i64, pd]
    lib.dicegp_run_device_stream.argtypes = [vp, C.c_uint64, pi64]
    lib.dicegp_chain_status.argtypes = [vp, i32, pi32, pi32, pi32, pi32, pi32, pi32]
    lib.dicegp_download_trace.argtypes = [vp, i32, i32, i32, pd, pd, pd, pd]
    lib.dicegp_posterior_summary.argtypes = [vp, i32, pi32, pd, pd, pi64, pd]
    lib.dicegp_last_run_stats.argtypes = [vp, pd, pi64, pi64, pi64]
    lib.dicegp_set_trace_budget.argtypes = [vp, i64]
    lib.dicegp_set_prerun_chains.argtypes = [vp, dbl, dbl, i32, i32, i32]
    lib.dicegp_timer_start.argtypes = [vp]
    lib.dicegp_timer_stop.argtypes = [vp, pd]
    lib.dicegp_counters.argtypes = [vp, pi64, pi64, pi64]
    lib.dicegp_eval_loglik.argtypes = [vp, i32, i32, i32, pd, pd]
    lib.dicegp_dump_stream.argtypes = [vp, i32, C.c_uint64, i64, i32, i32, pd, pd]
    lib.dicegp_work_counters.argtypes = [vp, pi64, i32, i32]
    for name in SYMBOLS:
        if name != "dicegp_last_error":
            getattr(lib, name).restype = C.c_int
    if lib.dicegp_abi_version() != ABI_VERSION:
        raise ImportError("libdicegp.so ABI %d != binding ABI %d" % (lib.dicegp_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype)) if arr is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class ChainShapes:
    """(C, Tc, M) of every installed chain, as arrays (80k pre-run chains per batch: no
    per-chain Python objects on the sampling path); indexes like a list of tuples."""

    def __init__(self, C_, Tc, M):
        self.C = np.asarray(C_, dtype=np.int64)
        self.Tc = np.broadcast_to(np.asarray(Tc, dtype=np.int64), self.C.shape)
        self.M = np.broadcast_to(np.asarray(M, dtype=np.int64), self.C.shape)

    def __len__(self):
        return self.C.size

    def __getitem__(self, i):
        return int(self.C[i]), int(self.Tc[i]), int(self.M[i])


def shape_limits():
    """Compiled shape limits of the library (no device needed): dict of max_isoforms, max_times, max_latent."""
    lim = Limits()
    rc = load().dicegp_get_limits(None, C.byref(lim))
    if rc != 0:
        raise DiceGPError(rc, "dicegp_get_limits")
    return {"max_isoforms": lim.max_isoforms, "max_times": lim.max_times, "max_latent": lim.max_latent}


class Context:
    """One handle per GPU (dicegp_create / dicegp_destroy)."""

    def __init__(self, device=0):
        self.lib = load()
        self._h = C.c_void_p()
        rc = self.lib.dicegp_create(int(device), C.byref(self._h))
        if rc != 0:
            raise DiceGPError(rc, self.lib.dicegp_last_error(None).decode())
        self.device = device
        self.gene_C = None
        self.T = 0
        self.chain_shapes = ChainShapes([], 1, 1)      # (C, Tc, M) per chain

    def close(self):
        if self._h:
            self.lib.dicegp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise DiceGPError(rc, self.lib.dicegp_last_error(self._h).decode())

    def limits(self):
        lim = Limits()
        self._check(self.lib.dicegp_get_limits(self._h, C.byref(lim)))
        return lim

    # -- genes ---------------------------------------------------------------------------
    def upload_genes(self, packed):
        """packed: diceseq_b200.packer.PackedGenes"""
        n_iso, n_reads = _i32(packed.n_iso), _i32(packed.n_reads)
        W, lens = _f64(packed.W), _f64(packed.len_iso)
        self._check(self.lib.dicegp_upload_genes(
            self._h, packed.G, packed.T, _p(n_iso, C.c_int32), _p(n_reads, C.c_int32),
            _p(W, C.c_double), W.size, _p(lens, C.c_double), lens.size))
        self.gene_C = n_iso.copy()
        self.T = packed.T
        self.chain_shapes = ChainShapes([], 1, 1)

    # -- chains --------------------------------------------------------------------------
    def set_chains(self, descs, pool):
        arr = (ChainDesc * len(descs))(*descs)
        pool = _f64(pool)
        self._check(self.lib.dicegp_set_chains(self._h, len(descs), arr, _p(pool, C.c_double), pool.size))
        self.chain_shapes = ChainShapes([self.gene_C[d.gene] for d in descs], [d.Tc for d in descs],
                                        [d.M for d in descs])

    def set_prerun_chains(self, theta1, var0=0.1, M=100, initial=100, gap=50):
        self._check(self.lib.dicegp_set_prerun_chains(self._h, float(theta1), float(var0), M, initial, gap))
        self.chain_shapes = ChainShapes(np.repeat(self.gene_C, self.T), 1, M)

    def timer_start(self):
        self._check(self.lib.dicegp_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.lib.dicegp_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def counters(self):
        a, b, d = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.dicegp_counters(self._h, C.byref(a), C.byref(b), C.byref(d)))
        return {"kernel_launches": a.value, "h2d_bytes": b.value, "d2h_bytes": d.value}

    def work_counters(self, reset=False):
        """What the chain kernels counted since the last reset (dicegp_work_counters)."""
        v = np.zeros(8, dtype=np.int64)
        self._check(self.lib.dicegp_work_counters(self._h, _p(v, C.c_int64), 8, 1 if reset else 0))
        return dict(zip(("w_bytes", "rounds", "candidate_evals", "steps", "fma"), (int(x) for x in v[:5])))

    def main_chains_from_prerun(self, M, initial, gap, theta1, kinv, prop_factor, prior_const):
        kinv = _f64(kinv)
        pf = _f64(prop_factor) if prop_factor is not None else None
        G = len(self.gene_C)
        var = np.zeros((G, self.limits().max_isoforms))
        self._check(self.lib.dicegp_main_chains_from_prerun(
            self._h, M, initial, gap, float(theta1), _p(kinv, C.c_double),
            _p(pf, C.c_double), float(prior_const), _p(var, C.c_double)))
        self.chain_shapes = ChainShapes(self.gene_C, self.T, M)
        return var

    # -- running -------------------------------------------------------------------------
    def run_host_stream(self, chain_ids, n_steps, delta, delta_off, u):
        ids, doff = _i32(chain_ids), _i64(delta_off)
        delta, u = _f64(delta), _f64(u)
        assert u.size == ids.size * n_steps
        self._check(self.lib.dicegp_run_host_stream(
            self._h, ids.size, _p(ids, C.c_int32), int(n_steps), _p(delta, C.c_double),
            _p(doff, C.c_int64), delta.size, _p(u, C.c_double)))

    def dump_stream(self, chain, seed, stream_id, n_steps, step0=0):
        """(delta (n_steps, C-1, Tc), u (n_steps,)) of the device stream of an installed chain
        (dicegp_dump_stream): what run_device_stream(seed, stream_id) feeds that chain with."""
        Cn, Tc, _M = self.chain_shapes[chain]
        delta = np.empty((n_steps, Cn - 1, Tc))
        u = np.empty(n_steps)
        self._check(self.lib.dicegp_dump_stream(self._h, int(chain), C.c_uint64(int(seed)), int(stream_id), int(step0),
                                                int(n_steps), _p(delta, C.c_double), _p(u, C.c_double)))
        return delta, u

    def run_device_stream(self, seed, stream_id=None):
        sid = _i64(stream_id) if stream_id is not None else None
        self._check(self.lib.dicegp_run_device_stream(self._h, C.c_uint64(int(seed)), _p(sid, C.c_int64)))

    # -- results -------------------------------------------------------------------------
    def eval_loglik(self, gene, y, t0=0):
        """lik[t] = sum_n log(W_t[n,:] . fsi[:,t]) of the state y (C x Tc) for the time points
        [t0, t0+Tc) of an uploaded gene (dicegp_eval_loglik; model_GP.py:335-345)."""
        y = _f64(y)
        Tc = y.shape[1]
        if y.shape[0] != int(self.gene_C[gene]):
            raise ValueError("y must have one row per isoform of the gene")
        lik = np.empty(Tc)
        self._check(self.lib.dicegp_eval_loglik(self._h, int(gene), int(t0), int(Tc), _p(y, C.c_double),
                                                _p(lik, C.c_double)))
        return lik

    def chain_status(self, chain_ids=None):
        if chain_ids is None:
            chain_ids = np.arange(len(self.chain_shapes))
        ids = _i32(chain_ids)
        out = [np.zeros(ids.size, dtype=np.int32) for _ in range(5)]
        self._check(self.lib.dicegp_chain_status(self._h, ids.size, _p(ids, C.c_int32),
                                                 *[_p(o, C.c_int32) for o in out]))
        return dict(zip(("m", "n_rows", "cnt", "state", "flags"), out))

    def download_trace(self, chain, n_rows, first_row=0):
        Cn, Tc, _M = self.chain_shapes[chain]
        psi = np.empty((n_rows, Cn, Tc))
        y = np.empty((n_rows, Cn, Tc))
        pro = np.empty(n_rows)
        lik = np.empty(n_rows)
        self._check(self.lib.dicegp_download_trace(
            self._h, int(chain), int(first_row), int(n_rows), _p(psi, C.c_double),
            _p(y, C.c_double), _p(pro, C.c_double), _p(lik, C.c_double)))
        return psi, y, pro, lik

    def set_trace_budget(self, nbytes):
        self._check(self.lib.dicegp_set_trace_budget(self._h, int(nbytes)))

    def posterior_summary(self, chain_ids=None, flat=False):
        """Device-side burn-in summaries: (psi_mean list, psi_ci list, lik_marg array)."""
        if chain_ids is None:
            chain_ids = np.arange(len(self.chain_shapes))
        ids = _i32(chain_ids)
        sizes = self.chain_shapes.C[ids] * self.chain_shapes.Tc[ids]
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        mean = np.empty(int(off[-1]))
        ci = np.empty(2 * int(off[-1]))
        lik = np.empty(ids.size)
        self._check(self.lib.dicegp_posterior_summary(
            self._h, ids.size, _p(ids, C.c_int32), _p(mean, C.c_double), _p(ci, C.c_double),
            _p(off[:-1].copy(), C.c_int64), _p(lik, C.c_double)))
        if flat:
            return mean, ci, lik, off
        means, cis = [], []
        for k, i in enumerate(ids):
            Cn, Tc, _M = self.chain_shapes[i]
            means.append(mean[off[k]:off[k + 1]].reshape(Cn, Tc))
            cis.append(ci[2 * off[k]:2 * off[k + 1]].reshape(Cn, Tc, 2))
        return means, cis, lik

    def last_run_stats(self):
        ms = C.c_double()
        launches, evals, steps = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.dicegp_last_run_stats(self._h, C.byref(ms), C.byref(launches),
                                                   C.byref(evals), C.byref(steps)))
        return {"kernel_ms": ms.value, "launches": launches.value, "read_evals": evals.value,
                "steps": steps.value}
