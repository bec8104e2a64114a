"""ctypes binding of libma_b200.so (include/ma_b200.h) — the thin Python side of the drop-in boundary.

The library is the product; this file only marshals numpy buffers.  It fails loudly when the CUDA extension
has not been built or no CUDA device is present: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libma_b200.so")

KSW_RIGHT = 0x02
KSW_EXTZ_ONLY = 0x40
KSW_REV_CIGAR = 0x80


class Params(ctypes.Structure):
    _fields_ = [
        ("match", ctypes.c_int32), ("mismatch", ctypes.c_int32), ("gap", ctypes.c_int32), ("extend", ctypes.c_int32),
        ("gap2", ctypes.c_int32), ("extend2", ctypes.c_int32), ("sv_penalty", ctypes.c_int32),
        ("seeding_technique", ctypes.c_int32), ("min_seed_length", ctypes.c_int32),
        ("min_ambiguity", ctypes.c_int32), ("max_ambiguity", ctypes.c_int32),
        ("seed_drop_min_size", ctypes.c_int32), ("seed_drop_factor", ctypes.c_double),
        ("max_num_soc", ctypes.c_int32), ("min_num_soc", ctypes.c_int32), ("soc_width", ctypes.c_int32),
        ("rectangular_soc", ctypes.c_int32),
        ("soc_score_drop", ctypes.c_double), ("harm_score_min", ctypes.c_int32),
        ("harm_score_min_rel", ctypes.c_double), ("score_diff_tolerance", ctypes.c_double),
        ("max_score_lookahead", ctypes.c_int32), ("switch_qlen", ctypes.c_int32),
        ("max_delta_dist", ctypes.c_double), ("min_delta_dist", ctypes.c_int32),
        ("optimistic_gap_estimation", ctypes.c_int32), ("gap_cost_cutting", ctypes.c_int32),
        ("max_gap_area", ctypes.c_int32), ("genome_size_disable", ctypes.c_int64),
        ("disable_heuristics", ctypes.c_int32),
        ("padding", ctypes.c_int32), ("bandwidth_ext", ctypes.c_int32), ("min_bandwidth_gap", ctypes.c_int32),
        ("zdrop", ctypes.c_int32), ("srand_base", ctypes.c_uint32),
    ]


KSW_TASK_DTYPE = np.dtype([("qoff", "<i8"), ("toff", "<i8"), ("qlen", "<i4"), ("tlen", "<i4"), ("w", "<i4"),
                           ("zdrop", "<i4"), ("flag", "<i4"), ("tag", "<i4")])
KSW_RESULT_DTYPE = np.dtype([("max", "<i4"), ("zdropped", "<i4"), ("max_q", "<i4"), ("max_t", "<i4"),
                             ("mqe", "<i4"), ("mqe_t", "<i4"), ("mte", "<i4"), ("mte_q", "<i4"), ("score", "<i4"),
                             ("n_cigar", "<i4"), ("reach_end", "<i4"), ("status", "<i4"), ("cigar_off", "<i8"),
                             ("cells", "<i8")])

_lib = None


class MaB200Error(RuntimeError):
    pass


def load_library():
    """Loads libma_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MaB200Error("libma_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C ma_b200/csrc`); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.ma_b200_params_preset.argtypes = [ctypes.c_char_p, ctypes.POINTER(Params)]
    lib.ma_b200_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.ma_b200_destroy.argtypes = [vp]
    lib.ma_b200_destroy.restype = None
    lib.ma_b200_last_error.argtypes = [vp]
    lib.ma_b200_last_error.restype = ctypes.c_char_p
    lib.ma_b200_launch_count.argtypes = [vp]
    lib.ma_b200_launch_count.restype = i64
    lib.ma_b200_set_params.argtypes = [vp, ctypes.POINTER(Params)]
    lib.ma_b200_ksw_upload.argtypes = [vp, i64, vp, vp, i64]
    lib.ma_b200_ksw_run.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.ma_b200_ksw_download.argtypes = [vp, vp, vp, i64, ctypes.POINTER(i64)]
    lib.ma_b200_ksw_batch.argtypes = [vp, i64, vp, vp, i64, vp, vp, i64, ctypes.POINTER(i64)]
    _lib = lib
    return lib


def preset(name: str) -> Params:
    p = Params()
    if load_library().ma_b200_params_preset(name.encode(), ctypes.byref(p)) != 0:
        raise MaB200Error("unknown preset %r" % name)
    return p


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class Context:
    """One CUDA device + stream + device buffers (ma_b200_ctx)."""

    def __init__(self, device: int = 0, preset_name: str = "default"):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.ma_b200_create(device, ctypes.byref(h))
        if rc != 0:
            raise MaB200Error("ma_b200_create(device=%d) failed with %d: no CUDA device / driver "
                              "(there is no CPU fallback)" % (device, rc))
        self.h = h
        self.params = preset(preset_name)
        self.set_params(self.params)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ma_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise MaB200Error("ma_b200 error %d: %s" % (rc, self.lib.ma_b200_last_error(self.h).decode()))

    def set_params(self, p: Params):
        self.params = p
        self._check(self.lib.ma_b200_set_params(self.h, ctypes.byref(p)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.ma_b200_launch_count(self.h))

    # ---- banded DP ------------------------------------------------------------------------------------------
    def ksw_upload(self, tasks: np.ndarray, seq: np.ndarray):
        tasks = np.ascontiguousarray(tasks, dtype=KSW_TASK_DTYPE)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        self._ksw_n = len(tasks)
        self._ksw_keep = (tasks, seq)
        self._check(self.lib.ma_b200_ksw_upload(self.h, len(tasks), _ptr(tasks), _ptr(seq), seq.size))

    def ksw_run(self) -> float:
        ms = ctypes.c_float(0)
        self._check(self.lib.ma_b200_ksw_run(self.h, ctypes.byref(ms)))
        return ms.value

    def ksw_download(self):
        res = np.zeros(self._ksw_n, dtype=KSW_RESULT_DTYPE)
        words = ctypes.c_int64(0)
        # first call with an empty slab to learn the size
        cig = np.zeros(1, dtype=np.uint32)
        rc = self.lib.ma_b200_ksw_download(self.h, _ptr(res), _ptr(cig), 0, ctypes.byref(words))
        if rc not in (0, -3):
            self._check(rc)
        cig = np.zeros(max(1, words.value), dtype=np.uint32)
        self._check(self.lib.ma_b200_ksw_download(self.h, _ptr(res), _ptr(cig), cig.size, ctypes.byref(words)))
        return res, cig[:words.value]

    def ksw_batch(self, tasks: np.ndarray, seq: np.ndarray, cigar_cap_words: int | None = None):
        """kswcpp_dispatch for a batch, host buffers in and out (ma_b200_ksw_batch)."""
        tasks = np.ascontiguousarray(tasks, dtype=KSW_TASK_DTYPE)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        res = np.zeros(len(tasks), dtype=KSW_RESULT_DTYPE)
        if cigar_cap_words is None:
            cigar_cap_words = int((tasks["qlen"].astype(np.int64) + tasks["tlen"] + 2).sum()) + 1
        cig = np.zeros(cigar_cap_words, dtype=np.uint32)
        words = ctypes.c_int64(0)
        self._check(self.lib.ma_b200_ksw_batch(self.h, len(tasks), _ptr(tasks), _ptr(seq), seq.size, _ptr(res),
                                               _ptr(cig), cig.size, ctypes.byref(words)))
        return res, cig[:words.value]


def pack_ksw_tasks(pairs):
    """pairs: iterable of (w, zdrop, flag, query uint8[], target uint8[]) -> (tasks, seq slab)."""
    pairs = list(pairs)
    tasks = np.zeros(len(pairs), dtype=KSW_TASK_DTYPE)
    chunks, off = [], 0
    for i, (w, zdrop, flag, q, t) in enumerate(pairs):
        tasks[i] = (off, off + len(q), len(q), len(t), w, zdrop, flag, i)
        chunks.append(np.asarray(q, dtype=np.uint8))
        chunks.append(np.asarray(t, dtype=np.uint8))
        off += len(q) + len(t)
    seq = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    return tasks, np.concatenate([seq, np.zeros(16, dtype=np.uint8)])
