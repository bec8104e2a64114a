"""pybind11 module (ma_b200/pybind) exposing the host-side module mirror under the reference's Python names
(libs/ma/src/util/export.cpp:38-67)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
import pipeline_common as PC


def load():
    subprocess.check_call(["make", "-s", "-C", os.path.join(H.ROOT, "ma_b200", "pybind")])
    sys.path.insert(0, os.path.join(H.ROOT, "ma_b200"))
    import ma_b200_py
    return ma_b200_py


def test_pybind_module_names_and_loud_failure_without_gpu():
    import torch
    m = load()
    for name in ("ParameterSetManager", "NucSeq", "Seed", "Seeds", "Segment", "Alignment", "FMIndex", "BinarySeeding",
                 "Harmonization", "NeedlemanWunsch", "MappingQuality", "PairedReads", "SmallInversions", "read_file",
                 "sam_header", "sam_records"):
        assert hasattr(m, name), name
    p = m.ParameterSetManager()
    p.set_selected("illumina_paired")
    assert p.use_paired_reads and p.by_name("max_ambiguity") == 500
    with pytest.raises(RuntimeError):
        p.set_selected("no such preset")
    q = m.NucSeq("ACGTNacgt")
    assert len(q) == 9 and str(q) == "ACGTNACGT"
    reads = m.read_file(os.path.join(H.GOLDEN, "gold_reads.fq"))
    parsed = [l.split("\t") for l in open(os.path.join(H.GOLDEN, "gold_reads_fq.parsed")).read().splitlines()]
    assert [(r.name, str(r)) for r in reads] == [(p[0], p[1]) for p in parsed]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            m.FMIndex(0)


@pytest.mark.gpu
def test_pybind_modules_match_reference_golden():
    m = load()
    gold = PC.load_gold("illumina")
    p = m.ParameterSetManager()
    p.set_selected("illumina")
    p.srand_base = PC.SRAND
    fm = m.FMIndex(0)
    fm.load(PC.GOLD_PREFIX)
    reads = [m.NucSeq(l.strip()) for l in open(PC.gold_reads("illumina")) if l.strip()]
    segs = m.BinarySeeding(p).execute(fm, reads)
    got = [[s.start, s.size, s.sa_start, s.sa_start_rev_comp, s.sa_size] for v in segs for s in v]
    assert np.array_equal(np.array(got, dtype=np.int64).reshape(-1), gold["seg"])
    alns = m.NeedlemanWunsch(p).execute(fm, reads)
    rows = [[a.begin_on_query, a.end_on_query, a.begin_on_ref(), a.end_on_ref(), a.get_score(), a.index_of_strip,
             a.length(), len(a.data)] for v in alns for a in v]
    assert np.array_equal(np.array(rows, dtype=np.int64).reshape(-1), gold["aln"])
    mq = m.MappingQuality(p).execute(fm, reads)
    exp = gold["mq"].reshape(-1, 3)
    k = 0
    for i, v in enumerate(mq):
        assert len(v) == gold["mq_off"][i + 1] - gold["mq_off"][i]
        for a in v:
            g = gold["aln"][8 * (gold["aln_off"][i] + exp[k][0]):][:8]
            assert (a.begin_on_query, a.end_on_query, a.get_score()) == (g[0], g[1], g[4])
            assert int(a.secondary) | int(a.supplementary) << 1 == exp[k][1]
            assert np.float64(a.mapping_quality).view(np.int64) == exp[k][2] or (
                np.isnan(a.mapping_quality) and np.isnan(np.int64(exp[k][2]).view(np.float64)))
            k += 1


@pytest.mark.gpu
def test_pybind_small_inversions_and_sam_text_match_reference():
    """read_file -> MappingQuality -> SmallInversions -> sam_records == the reference's SAM (gold_inv_default_z20.sam)."""
    m = load()
    p = m.ParameterSetManager()
    p.set_selected("default")
    p.srand_base = PC.SRAND
    p.z_drop_inversions = 20
    fm = m.FMIndex(0)
    fm.load(PC.GOLD_PREFIX)
    reads = m.read_file(os.path.join(H.GOLDEN, "gold_reads_inv.fa"))
    mq = m.MappingQuality(p).execute(fm, reads)
    inv = m.SmallInversions(p).execute(fm, mq, reads)
    assert sum(len(v) for v in inv) == sum(len(v) for v in mq) + 14
    text = m.sam_header(fm) + m.sam_records(fm, reads, inv)
    assert text == open(os.path.join(H.GOLDEN, "gold_inv_default_z20.sam")).read()
