"""The C++ host-side mirror of the reference's module interface (include/ma_b200_modules.hpp)."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
import pipeline_common as PC


def build(tmp_path):
    exe = str(tmp_path / "test_modules")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(H.ROOT, "tests", "cpp", "test_modules.cpp"),
                           "-L" + os.path.join(H.ROOT, "ma_b200"), "-lma_b200",
                           "-Wl,-rpath," + os.path.join(H.ROOT, "ma_b200")])
    return exe


def test_cpp_modules_compile_and_fail_loudly_without_gpu(tmp_path):
    import torch
    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([exe, PC.GOLD_PREFIX, PC.gold_reads("illumina"), "illumina", "1000"], capture_output=True)
    assert p.returncode == 3 and b"no CUDA device" in p.stderr


@pytest.mark.gpu
def test_cpp_aligner_matches_reference_golden(tmp_path):
    exe = build(tmp_path)
    out = subprocess.check_output([exe, PC.GOLD_PREFIX, PC.gold_reads("illumina"), "illumina", str(PC.SRAND)]).decode()
    gold = PC.load_gold("illumina")
    rows = [l.split() for l in out.strip().splitlines()]
    assert len(rows) == len(gold["aln"]) // 8
    k = 0
    for i in range(len(gold["aln_off"]) - 1):
        for j in range(gold["aln_off"][i], gold["aln_off"][i + 1]):
            g = gold["aln"][8 * j:8 * j + 8]
            r = rows[k]
            assert int(r[0]) == i
            assert [int(x) for x in r[1:8]] == [int(g[0]), int(g[1]), int(g[2]), int(g[3]), int(g[4]), int(g[5]),
                                                int(g[6])]
            runs = gold["alndata"][2 * gold["alndata_off"][j]:2 * gold["alndata_off"][j + 1]].reshape(-1, 2)
            assert [tuple(map(int, x.split(":"))) for x in r[8:]] == [tuple(map(int, x)) for x in runs]
            k += 1


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["illumina", "illuminapaired"])
def test_cpp_report_matches_reference_golden(tmp_path, preset):
    """Aligner::report = MappingQuality per read (single-end presets) / PairedReads per pair (paired preset)."""
    exe = build(tmp_path)
    out = subprocess.check_output([exe, PC.GOLD_PREFIX, PC.gold_reads(preset), preset, str(PC.SRAND), "report"]).decode()
    gold = PC.load_gold(preset)
    rows = [[int(x) for x in l.split()] for l in out.strip().splitlines()]
    exp = []
    if preset == "illuminapaired":
        pr = gold["pr"].reshape(-1, 4)
        for p in range(len(gold["pr_off"]) - 1):
            for mate, idx, flags, bits in pr[gold["pr_off"][p]:gold["pr_off"][p + 1]]:
                g = gold["aln"][8 * (gold["aln_off"][2 * p + mate] + idx):][:8]
                exp.append([p, int(mate), int(g[0]), int(g[1]), int(g[2]), int(g[3]), int(g[4]), int(flags), int(bits)])
    else:
        mq = gold["mq"].reshape(-1, 3)
        for i in range(len(gold["mq_off"]) - 1):
            for idx, flags, bits in mq[gold["mq_off"][i]:gold["mq_off"][i + 1]]:
                g = gold["aln"][8 * (gold["aln_off"][i] + idx):][:8]
                exp.append([i, 1, int(g[0]), int(g[1]), int(g[2]), int(g[3]), int(g[4]), int(flags), int(bits)])
    assert rows == exp


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["illumina", "illuminapaired", "pacbio"])
def test_cpp_sam_output_matches_reference_writers(tmp_path, preset):
    """SURVEY.md §8(f) N2 (writer side): include/ma_b200_sam.hpp against the SAM text of the reference's own
    FileWriter / PairedFileWriter (tests/golden/gold_<preset>.sam)."""
    exe = build(tmp_path)
    out = subprocess.check_output([exe, PC.GOLD_PREFIX, PC.gold_reads(preset), preset, str(PC.SRAND), "sam"]).decode()
    exp = open(os.path.join(H.GOLDEN, "gold_%s.sam" % preset)).read()
    got_lines, exp_lines = out.splitlines(), exp.splitlines()
    assert len(got_lines) == len(exp_lines)
    for i, (g, e) in enumerate(zip(got_lines, exp_lines)):
        assert g == e, (i, g[:120], e[:120])


@pytest.mark.gpu
@pytest.mark.parametrize("reads,preset,gold", [("gold_reads.fq", "illuminapaired", "gold_fq_illuminapaired.sam"),
                                               ("gold_reads.fa", "illumina", "gold_fa_illumina.sam")])
def test_cpp_fastq_to_sam_matches_reference(tmp_path, reads, preset, gold):
    """FASTQ / FASTA in, SAM out (names, qualities, mate fields) == reference FileReader -> path -> (Paired)FileWriter."""
    exe = build(tmp_path)
    out = subprocess.check_output([exe, PC.GOLD_PREFIX, os.path.join(H.GOLDEN, reads), preset, str(PC.SRAND), "sam"])
    assert out.decode() == open(os.path.join(H.GOLDEN, gold)).read()
