"""Shared test helpers: dump loader, oracle (CPU restatement) bindings, reference harness runner.

The oracle and oracle/_ref are CHECKERS: only tests, smoke() and bench.py's cpu_baseline may use them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DUMP = os.path.join(ORACLE_DIR, "_ref", "ref_dump")
GOLDEN = os.path.join(ROOT, "tests", "golden")

KSW_FIELDS = ["qlen", "tlen", "w", "zdrop", "flag", "max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte",
              "mte_q", "score", "n_cigar", "reach_end"]


def load_dump(path: str) -> dict:
    """Reads the MADUMP1 container written by oracle/ref_dump.cpp (named int64 arrays)."""
    out = {}
    with open(path, "rb") as f:
        assert f.readline() == b"MADUMP1\n"
        while True:
            line = f.readline()
            if not line:
                break
            name, n = line.decode().split()
            out[name] = np.frombuffer(f.read(int(n) * 8), dtype=np.int64).copy()
    return out


def have_ref() -> bool:
    return os.path.exists(REF_DUMP)


def run_ref(*args) -> str:
    return subprocess.check_output([REF_DUMP, *[str(a) for a in args]]).decode()


class OracleScore(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("match", "mismatch", "gap", "extend", "gap2", "extend2")]


class OracleKsw(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q",
                                            "score", "n_cigar", "reach_end")]


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libma_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
        _oracle = ctypes.CDLL(path)
    return _oracle


DEFAULT_SCORE = OracleScore(2, 4, 4, 2, 24, 1)


def oracle_ksw(q: np.ndarray, t: np.ndarray, w: int, zdrop: int, flag: int, score=DEFAULT_SCORE):
    """Returns (dict of kswcpp_extz_t fields, cigar uint32 array, cells)."""
    lib = oracle_lib()
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    ez = OracleKsw()
    cap = len(q) + len(t) + 8
    cig = np.zeros(cap, dtype=np.uint32)
    cells = ctypes.c_int64(0)
    rc = lib.ma_oracle_ksw(len(q), q.ctypes.data_as(ctypes.c_void_p), len(t), t.ctypes.data_as(ctypes.c_void_p),
                           ctypes.byref(score), w, zdrop, flag, ctypes.byref(ez),
                           cig.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(cells))
    assert rc == 0
    res = {n: getattr(ez, n) for n, _ in OracleKsw._fields_}
    return res, cig[:ez.n_cigar].copy(), cells.value


def split_ksw_dump(d: dict):
    """Yields (fields dict, query, target, cigar) for every logged kswcpp call of a ref_dump file."""
    calls = d["ksw_calls"].reshape(-1, 16)
    seq = d["ksw_seq"].astype(np.uint8)
    cig = d["ksw_cigar"].astype(np.uint32)
    so = co = 0
    for row in calls:
        f = dict(zip(KSW_FIELDS, (int(x) for x in row)))
        q = seq[so:so + f["qlen"]]
        t = seq[so + f["qlen"]:so + f["qlen"] + f["tlen"]]
        so += f["qlen"] + f["tlen"]
        c = cig[co:co + f["n_cigar"]]
        co += f["n_cigar"]
        yield f, q, t, c


def write_pairs(path: str, pairs) -> None:
    """pairs: iterable of (w, zdrop, flag, q, t) -> text format of `ref_dump ksw`."""
    with open(path, "w") as f:
        for w, zdrop, flag, q, t in pairs:
            qs = "".join(map(str, q.tolist())) or "-"
            ts = "".join(map(str, t.tolist())) or "-"
            f.write("%d %d %d %s %s\n" % (w, zdrop, flag, qs, ts))


def oracle_align_dump(prefix: str, reads_txt: str, preset: str, out: str, srand_base: int = -1, stages: int = 5,
                      overrides=None, min_genome_size: int = -1, params=None):
    """overrides: dict with any of bandwidth_ext, zdrop, padding, max_gap_area, min_bandwidth_gap;
    min_genome_size: "Minimum Genome Size for Heuristics" (-1: the preset's 10 M);
    params: dict of ma_b200_params field names -> values applied after the preset (PARAM_NAMES maps them to the
    reference's parameter names)."""
    lib = oracle_lib()
    lib.ma_oracle_set_params.argtypes = [ctypes.c_char_p]
    lib.ma_oracle_set_params(";".join("%s=%r" % kv for kv in (params or {}).items()).encode())
    lib.ma_oracle_set_min_genome_size.argtypes = [ctypes.c_longlong]
    lib.ma_oracle_set_min_genome_size(min_genome_size)
    keys = ["bandwidth_ext", "zdrop", "padding", "max_gap_area", "min_bandwidth_gap"]
    ov = (ctypes.c_int * 5)(*[int((overrides or {}).get(k, -1)) for k in keys])
    lib.ma_oracle_set_overrides(ov)
    err = ctypes.create_string_buffer(512)
    lib.ma_oracle_align_dump.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p,
                                         ctypes.c_longlong, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    rc = lib.ma_oracle_align_dump(prefix.encode(), reads_txt.encode(), preset.encode(), out.encode(), srand_base,
                                  stages, err, 512)
    if rc != 0:
        raise RuntimeError("oracle: " + err.value.decode())
    return load_dump(out)


# ma_b200_params field -> the reference's parameter name (libs/ms/inc/ms/util/parameter.h:621-880)
PARAM_NAMES = {
    "seeding_technique": "Seeding Technique", "min_seed_length": "Minimal Seed Length",
    "max_ambiguity": "Maximal Ambiguity", "seed_drop_min_size": "Seeding Drop-off A - Minimal Seed Size",
    "seed_drop_factor": "Seeding Drop-off B - Factor", "max_num_soc": "Maximal Number of SoCs",
    "min_num_soc": "Minimal Number of SoCs", "soc_score_drop": "SoC Score Drop-off",
    "harm_score_min": "Minimal Harmonization Score", "harm_score_min_rel": "Relative Minimal Harmonization Score",
    "score_diff_tolerance": "Harmonization Drop-off A - Score Difference",
    "max_score_lookahead": "Harmonization Drop-off B - Lookahead",
    "switch_qlen": "Harmonization Score Drop-off - Minimal Query Length",
    "max_delta_dist": "Artifact Filter A - Maximal Delta Distance",
    "min_delta_dist": "Artifact Filter B - Minimal Delta Distance",
    "optimistic_gap_estimation": "Pick Local Seed Set B - Optimistic Gap Estimation",
    "gap_cost_cutting": "Pick Local Seed Set A - Enabled", "max_gap_area": "Maximal Gap Size",
    "genome_size_disable": "Minimum Genome Size for Heuristics", "disable_heuristics": "Disable All Heuristics",
    "padding": "Padding", "bandwidth_ext": "Bandwidth for Extensions", "min_bandwidth_gap": "Minimal Bandwidth in Gaps",
    "zdrop": "Z Drop", "report_n": "Maximal Number of Reported Alignments",
    "min_alignment_score": "Minimal Alignment Score", "max_supplementary_per_prim": "Number Supplementary Alignments",
    "max_overlap_supplementary": "Maximal Supplementary Overlap", "paired_mean": "Mean Distance of Paired Reads",
    "paired_std": "Standard Deviation of Paired Reads", "paired_bonus": "Score Factor for Paired Reads",
}


def ref_param_env(params: dict) -> dict:
    """Environment for `ref_dump` that sets the same parameters on the reference (MA_REF_SET, by reference name)."""
    def text(k, v):
        if k == "seeding_technique":
            return "1" if v else "0"  # a choice parameter: index into { maxSpan, SMEMs, MEMs }
        if k in ("disable_heuristics", "optimistic_gap_estimation", "gap_cost_cutting"):
            return "true" if v else "false"
        return repr(v)
    return {"MA_REF_SET": ";".join("%s=%s" % (PARAM_NAMES[k], text(k, v)) for k, v in params.items())}
