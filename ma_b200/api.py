"""ctypes binding of libma_b200.so (include/ma_b200.h) — the thin Python side of the drop-in boundary.

The library is the product; this file only marshals numpy buffers.  It fails loudly when the CUDA extension
has not been built or no CUDA device is present: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MA_B200_LIB: another build of the same library (A/B of compile-time knobs, scripts/); never a different backend
LIB_PATH = os.environ.get("MA_B200_LIB") or os.path.join(_HERE, "libma_b200.so")

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
        ("report_n", ctypes.c_int32), ("min_alignment_score", ctypes.c_int32),
        ("max_supplementary_per_prim", ctypes.c_int32), ("use_paired_reads", ctypes.c_int32),
        ("max_overlap_supplementary", ctypes.c_double), ("paired_mean", ctypes.c_double),
        ("paired_std", ctypes.c_double), ("paired_bonus", ctypes.c_double),
    ]


KSW_TASK_DTYPE = np.dtype([("qoff", "<i8"), ("toff", "<i8"), ("qlen", "<i4"), ("tlen", "<i4"), ("w", "<i4"),
                           ("zdrop", "<i4"), ("flag", "<i4"), ("tag", "<i4")])
KSW_RESULT_DTYPE = np.dtype([("max", "<i4"), ("zdropped", "<i4"), ("max_q", "<i4"), ("max_t", "<i4"),
                             ("mqe", "<i4"), ("mqe_t", "<i4"), ("mte", "<i4"), ("mte_q", "<i4"), ("score", "<i4"),
                             ("n_cigar", "<i4"), ("reach_end", "<i4"), ("status", "<i4"), ("cigar_off", "<i8"),
                             ("cells", "<i8")])

SEED_DTYPE = np.dtype([("q", "<i4"), ("len", "<i4"), ("r", "<i8"), ("amb", "<u4"), ("fw", "<i4"), ("delta", "<i8")])
SEGMENT_DTYPE = np.dtype([("start", "<i4"), ("size", "<i4"), ("sa_start", "<i8"), ("sa_rev", "<i8"),
                          ("sa_size", "<i8")])
SET_DTYPE = np.dtype([("read", "<i4"), ("ordinal", "<i4"), ("soc_index", "<u4"), ("n", "<i4"), ("seed_off", "<i8"),
                      ("task_off", "<i4"), ("n_tasks", "<i4"), ("win_begin", "<u8"), ("win_end", "<u8"),
                      ("valid", "<i4"), ("pad", "<i4")])
ALN_DTYPE = np.dtype([("begin_ref", "<i8"), ("end_ref", "<i8"), ("score", "<i8"), ("begin_q", "<i4"),
                      ("end_q", "<i4"), ("length", "<i4"), ("n_runs", "<i4"), ("soc_index", "<u4"), ("read", "<i4"),
                      ("run_off", "<i8"), ("rank", "<i4"), ("flags", "<i4"), ("mapq", "<f8"), ("rank_mq", "<i4"),
                      ("pair_rank", "<i4")])
INFO_DTYPE = np.dtype([("seed_off", "<i8"), ("n_seeds", "<i4"), ("set_off", "<i4"), ("n_sets", "<i4"),
                       ("status", "<i4")])
READ_ELISTS, READ_ESEGMENTS, READ_ESETS, READ_EBAND = 1, 2, 4, 8
STAGE_SEEDS, STAGE_SETS, STAGE_ALIGN, STAGE_MAPQ = 1, 2, 3, 4
ALN_SECONDARY, ALN_SUPPLEMENTARY, ALN_FIRST_MATE = 1, 2, 4


class AlignStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("n_reads", "n_seeds", "n_sets", "n_set_seeds", "n_tasks", "n_runs",
                                              "n_cigar_words", "n_ext", "n_invpsi", "n_dropped", "dp_cells",
                                              "n_lookup")] + \
               [(n, ctypes.c_float) for n in ("ms_seed", "ms_locate", "ms_socharm", "ms_plan", "ms_dp",
                                              "ms_assemble", "ms_total")] + [("launches", ctypes.c_int32),
                                                                              ("n_failed", ctypes.c_int32),
                                                                              ("n_reported", ctypes.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


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
    lib.ma_b200_ksw_set_extension_only.argtypes = [vp, ctypes.c_int32]
    lib.ma_b200_set_batch_split.argtypes = [vp, i64]
    lib.ma_b200_create_sibling.argtypes = [vp, ctypes.POINTER(vp)]
    lib.ma_b200_set_reported_only.argtypes = [vp, ctypes.c_int32]
    lib.ma_b200_ksw_upload.argtypes = [vp, i64, vp, vp, i64]
    lib.ma_b200_ksw_run.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.ma_b200_ksw_download.argtypes = [vp, vp, vp, i64, ctypes.POINTER(i64)]
    lib.ma_b200_ksw_batch.argtypes = [vp, i64, vp, vp, i64, vp, vp, i64, ctypes.POINTER(i64)]
    lib.ma_b200_index_upload.argtypes = [vp, vp, i64, vp, i64, i64, vp, i64, ctypes.c_int32, vp, i64, i64, vp, vp,
                                         ctypes.c_int32]
    lib.ma_b200_index_build.argtypes = [vp, vp, i64, vp, vp, ctypes.c_int32]
    lib.ma_b200_index_sizes.argtypes = [vp] + [ctypes.POINTER(i64)] * 4 + [vp]
    lib.ma_b200_index_download.argtypes = [vp, vp, vp, vp]
    lib.ma_b200_align_upload.argtypes = [vp, i64, vp, vp]
    lib.ma_b200_align_run.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(AlignStats)]
    lib.ma_b200_align_download_info.argtypes = [vp, vp]
    lib.ma_b200_align_download_segments.argtypes = [vp, vp, vp]
    lib.ma_b200_align_download_seeds.argtypes = [vp, vp, i64]
    lib.ma_b200_align_download_sets.argtypes = [vp, vp, i64, vp, i64]
    lib.ma_b200_align_download.argtypes = [vp, vp, vp, i64, vp, i64]
    lib.ma_b200_align_batch.argtypes = [vp, i64, vp, vp, vp, vp, i64, vp, i64, ctypes.POINTER(AlignStats)]
    lib.ma_b200_gather_probe.argtypes = [vp, i64, ctypes.POINTER(ctypes.c_double)]
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

    def sibling(self):
        """A second context on the same device sharing this one's index (ma_b200_create_sibling): for two batches in
        flight from two host threads. This context must outlive it."""
        other = Context.__new__(Context)
        other.lib = self.lib
        h = ctypes.c_void_p()
        self._check(self.lib.ma_b200_create_sibling(self.h, ctypes.byref(h)))
        other.h = h
        other.params = self.params
        other._stats, other._n_reads = {}, 0
        other._reported_only = getattr(self, "_reported_only", False)
        return other

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

    # ---- index ----------------------------------------------------------------------------------------------
    def index_upload(self, ix):
        """ix: ma_b200.index.Index (the reference's own arrays)."""
        bwt = np.ascontiguousarray(ix.bwt, dtype=np.uint32)
        sa = np.ascontiguousarray(ix.sa, dtype=np.int64)
        pac = np.ascontiguousarray(ix.pac, dtype=np.uint8)
        L2 = np.ascontiguousarray(ix.L2, dtype=np.int64)
        cs = np.ascontiguousarray(ix.contig_start, dtype=np.int64)
        cl = np.ascontiguousarray(ix.contig_len, dtype=np.int64)
        self._check(self.lib.ma_b200_index_upload(self.h, _ptr(bwt), bwt.size, _ptr(L2), ix.primary, ix.ref_len,
                                                  _ptr(sa), sa.size, ix.sa_intv, _ptr(pac), pac.size, ix.fwd_len,
                                                  _ptr(cs), _ptr(cl), len(cs)))
        self._contigs = (cs.copy(), cl.copy())

    def index_build(self, fwd_codes: np.ndarray, contig_start, contig_len):
        """Builds the FM-index on the GPU from the forward strand; it stays resident in HBM."""
        fwd = np.ascontiguousarray(fwd_codes, dtype=np.uint8)
        cs = np.ascontiguousarray(contig_start, dtype=np.int64)
        cl = np.ascontiguousarray(contig_len, dtype=np.int64)
        self._check(self.lib.ma_b200_index_build(self.h, _ptr(fwd), fwd.size, _ptr(cs), _ptr(cl), len(cs)))
        self._contigs = (cs.copy(), cl.copy())

    def index_download(self, names=None):
        from .index import Index
        nw, ns, npac, prim = (ctypes.c_int64(0) for _ in range(4))
        L2 = np.zeros(5, dtype=np.int64)
        self._check(self.lib.ma_b200_index_sizes(self.h, ctypes.byref(nw), ctypes.byref(ns), ctypes.byref(npac),
                                                 ctypes.byref(prim), _ptr(L2)))
        ix = Index()
        ix.bwt = np.zeros(nw.value, dtype=np.uint32)
        ix.sa = np.zeros(ns.value, dtype=np.int64)
        ix.pac = np.zeros(npac.value, dtype=np.uint8)
        self._check(self.lib.ma_b200_index_download(self.h, _ptr(ix.bwt), _ptr(ix.sa), _ptr(ix.pac)))
        ix.L2, ix.primary, ix.ref_len, ix.fwd_len = L2, prim.value, int(L2[4]), int(L2[4]) // 2
        ix.contig_start, ix.contig_len = self._contigs
        ix.contig_names = list(names) if names else ["chr%d" % (i + 1) for i in range(len(ix.contig_start))]
        return ix

    # ---- alignment path -------------------------------------------------------------------------------------
    def align_upload(self, reads: np.ndarray, offsets: np.ndarray):
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self._n_reads = len(offsets) - 1
        self._check(self.lib.ma_b200_align_upload(self.h, self._n_reads, _ptr(reads), _ptr(offsets)))

    def align_run(self, stage: int = STAGE_ALIGN, keep_segments: int = 0) -> dict:
        st = AlignStats()
        self._keep_segments = keep_segments
        self._check(self.lib.ma_b200_align_run(self.h, stage, keep_segments, ctypes.byref(st)))
        self._stats = st.as_dict()
        return self._stats

    def download_info(self):
        info = np.zeros(self._n_reads, dtype=INFO_DTYPE)
        self._check(self.lib.ma_b200_align_download_info(self.h, _ptr(info)))
        return info

    def download_segments(self):
        segs = np.zeros((self._n_reads, self._keep_segments), dtype=SEGMENT_DTYPE)
        n = np.zeros(self._n_reads, dtype=np.int32)
        self._check(self.lib.ma_b200_align_download_segments(self.h, _ptr(segs), _ptr(n)))
        return segs, n

    def download_seeds(self):
        seeds = np.zeros(max(1, self._stats["n_seeds"]), dtype=SEED_DTYPE)
        self._check(self.lib.ma_b200_align_download_seeds(self.h, _ptr(seeds), seeds.size))
        return seeds[:self._stats["n_seeds"]]

    def download_sets(self):
        sets = np.zeros(max(1, self._stats["n_sets"]), dtype=SET_DTYPE)
        seeds = np.zeros(max(1, self._stats["n_set_seeds"]), dtype=SEED_DTYPE)
        self._check(self.lib.ma_b200_align_download_sets(self.h, _ptr(sets), sets.size, _ptr(seeds), seeds.size))
        return sets[:self._stats["n_sets"]], seeds

    def download_alignments(self):
        info = np.zeros(self._n_reads, dtype=INFO_DTYPE)
        alns = np.zeros(max(1, self._stats["n_sets"]), dtype=ALN_DTYPE)
        runs = np.zeros(max(1, self._stats["n_runs"]), dtype=np.uint32)
        self._check(self.lib.ma_b200_align_download(self.h, _ptr(info), _ptr(alns), alns.size, _ptr(runs),
                                                    runs.size))
        return info, alns[:self._stats["n_sets"]], runs

    def align_batch(self, reads: np.ndarray, offsets: np.ndarray, cap_alns: int | None = None,
                    cap_runs: int | None = None):
        """The drop-in call: host buffers in, alignment records out (ma_b200_align_batch)."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        cap_alns = cap_alns or 4 * n + 1024
        cap_runs = cap_runs or 64 * n + 4096
        info = np.zeros(n, dtype=INFO_DTYPE)
        alns = np.zeros(cap_alns, dtype=ALN_DTYPE)
        runs = np.zeros(cap_runs, dtype=np.uint32)
        st = AlignStats()
        rc = self.lib.ma_b200_align_batch(self.h, n, _ptr(reads), _ptr(offsets), _ptr(info), _ptr(alns),
                                          cap_alns, _ptr(runs), cap_runs, ctypes.byref(st))
        if rc == -3 and (st.n_sets > cap_alns or st.n_runs > cap_runs) and not getattr(self, "_reported_only", False):
            # MA_B200_ENOMEM: the record arrays were too small (long reads carry hundreds of runs per alignment); the
            # results are still on the device and the stats hold the exact counts: fetch them into arrays of that size
            alns = np.zeros(max(1, st.n_sets), dtype=ALN_DTYPE)
            runs = np.zeros(max(1, st.n_runs), dtype=np.uint32)
            rc = self.lib.ma_b200_align_download(self.h, _ptr(info), _ptr(alns), alns.size, _ptr(runs), runs.size)
        self._check(rc)
        self._stats = st.as_dict()
        n_out = st.n_reported if getattr(self, "_reported_only", False) else st.n_sets
        return info, alns[:n_out], runs[:st.n_runs], self._stats

    def set_reported_only(self, on: bool):
        """Downloads after STAGE_MAPQ deliver only the records MappingQuality / PairedReads return (rank_mq >= 0)."""
        self._check(self.lib.ma_b200_set_reported_only(self.h, 1 if on else 0))
        self._reported_only = bool(on)

    def gather_probe(self, buffer_bytes: int) -> float:
        """Measured GB/s of independent random 64-byte reads over a buffer of this size (seeding roofline)."""
        g = ctypes.c_double(0)
        self._check(self.lib.ma_b200_gather_probe(self.h, int(buffer_bytes), ctypes.byref(g)))
        return g.value

    # ---- banded DP ------------------------------------------------------------------------------------------
    def set_batch_split(self, reads_per_subbatch: int):
        """Sub-batch size of the pipelined align_batch (batches of >= 2 sub-batches are pipelined)."""
        self._check(self.lib.ma_b200_set_batch_split(self.h, int(reads_per_subbatch)))

    def ksw_set_extension_only(self, on: bool):
        """Early-termination mode for extension tasks: only max / max_q / max_t / CIGAR are defined."""
        self._check(self.lib.ma_b200_ksw_set_extension_only(self.h, 1 if on else 0))

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


def pack_reads(reads):
    """list of uint8 arrays (or a 2-D array) -> (concatenated bytes, int64 offsets[n+1])."""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, L = reads.shape
        return np.ascontiguousarray(reads, dtype=np.uint8).reshape(-1), np.arange(n + 1, dtype=np.int64) * L
    lens = np.array([len(r) for r in reads], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    data = np.concatenate([np.asarray(r, dtype=np.uint8) for r in reads]) if len(reads) else np.zeros(0, np.uint8)
    return data, off


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
