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
