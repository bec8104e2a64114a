"""SURVEY.md §8(f) N2, writer side: include/ma_b200_sam.hpp must print what the reference's FileWriter /
PairedFileWriter print (tests/golden/gold_<preset>.sam, written by the compiled reference). CPU only: the reported
alignments come from the golden dumps of the reference."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
import pipeline_common as PC


def records(preset, paired):
    g = PC.load_gold(preset)
    reads = [l.strip() for l in open(PC.gold_reads(preset)) if l.strip()]
    out = ["Q r%d %s" % (i, r or "-") for i, r in enumerate(reads)]

    def aln_row(read, idx):
        j = g["aln_off"][read] + idx
        a = g["aln"][8 * j:8 * j + 8]
        runs = g["alndata"][2 * g["alndata_off"][j]:2 * g["alndata_off"][j + 1]].reshape(-1, 2)
        return a, " ".join("%d:%d" % (t, n) for t, n in runs)

    if paired:
        pr = g["pr"].reshape(-1, 4)
        for p in range(len(g["pr_off"]) - 1):
            for mate, idx, flags, bits in pr[g["pr_off"][p]:g["pr_off"][p + 1]]:
                a, runs = aln_row(2 * p + mate, idx)
                out.append("A %d %d %d %d %d %d %d %d %d %d %d %s" % (p, 1 - mate, a[0], a[1], a[2], a[3], a[4], a[6],
                                                                      flags & 1, flags >> 1, bits, runs))
    else:
        mq = g["mq"].reshape(-1, 3)
        for i in range(len(g["mq_off"]) - 1):
            for idx, flags, bits in mq[g["mq_off"][i]:g["mq_off"][i + 1]]:
                a, runs = aln_row(i, idx)
                out.append("A %d 0 %d %d %d %d %d %d %d %d %d %s" % (i, a[0], a[1], a[2], a[3], a[4], a[6], flags & 1,
                                                                     flags >> 1, bits, runs))
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("preset", ["illumina", "illuminapaired", "pacbio"])
def test_sam_writer_matches_reference_writers(tmp_path, preset):
    exe = str(tmp_path / "test_sam")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(H.ROOT, "tests", "cpp", "test_sam.cpp"),
                           "-L" + os.path.join(H.ROOT, "ma_b200"), "-lma_b200",
                           "-Wl,-rpath," + os.path.join(H.ROOT, "ma_b200")])
    paired = preset == "illuminapaired"
    got = subprocess.run([exe, PC.GOLD_PREFIX, "1" if paired else "0"], input=records(preset, paired).encode(),
                         capture_output=True, check=True).stdout.decode()
    exp = open(os.path.join(H.GOLDEN, "gold_%s.sam" % preset)).read()
    gl, el = got.splitlines(), exp.splitlines()
    for i, (a, b) in enumerate(zip(gl, el)):
        assert a == b, (i, a[:140], b[:140])
    assert len(gl) == len(el)


@pytest.mark.parametrize("name", ["fq", "fa"])
def test_read_parser_matches_reference_file_reader(tmp_path, name):
    """FASTQ (multi-line records, CRLF, lower case, descriptions, blank lines) and FASTA (wrapped lines, IUPAC codes)
    parsed like the reference's FileReader (tests/golden/gold_reads_<name>.parsed written by `ref_dump reads`)."""
    exe = str(tmp_path / "test_reader")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(H.ROOT, "tests", "cpp", "test_reader.cpp"),
                           "-L" + os.path.join(H.ROOT, "ma_b200"), "-lma_b200",
                           "-Wl,-rpath," + os.path.join(H.ROOT, "ma_b200")])
    got = subprocess.check_output([exe, os.path.join(H.GOLDEN, "gold_reads.%s" % name)]).decode()
    assert got == open(os.path.join(H.GOLDEN, "gold_reads_%s.parsed" % name)).read()


def _build_reader2(tmp_path):
    exe = str(tmp_path / "test_reader2")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(H.ROOT, "tests", "cpp", "test_reader2.cpp"),
                           "-DMA_B200_WITH_ZLIB", "-lpthread", "-L" + os.path.join(H.ROOT, "ma_b200"), "-lma_b200", "-lz",
                           "-Wl,-rpath," + os.path.join(H.ROOT, "ma_b200")])
    return exe


@pytest.mark.parametrize("name", ["fq", "fa"])
def test_read_parser_two_pass_equals_serial_and_reference(tmp_path, name):
    """nextRecord() + parseRecord() on several threads (what maCMD_b200's reader does) == next() == the reference."""
    exe = _build_reader2(tmp_path)
    got = subprocess.check_output([exe, os.path.join(H.GOLDEN, "gold_reads.%s" % name)]).decode()
    assert got == open(os.path.join(H.GOLDEN, "gold_reads_%s.parsed" % name)).read()


@pytest.mark.parametrize("text,message", [(b"ACGT\n", b"FASTA/Q"), (b">empty\n\n>next\nACGT\n", b"found empty read"),
                                          (b"", None)])
def test_read_parser_rejects_malformed_input(tmp_path, text, message):
    """Error behaviour of FileReader::execute (fileReader.cpp:37-203): not FASTA/FASTQ, empty read; an empty file
    holds no reads."""
    exe = _build_reader2(tmp_path)
    f = tmp_path / "in.txt"
    f.write_bytes(text)
    r = subprocess.run([exe, str(f)], capture_output=True)
    if message is None:
        assert r.returncode == 0 and r.stdout == b""
    else:
        assert r.returncode == 3 and message in r.stderr, r.stderr


def test_read_parser_reads_gzip_input(tmp_path):
    """.gz input like the reference's GzFileStream (fileReader.h:286-400): same reads as from the plain file."""
    import gzip
    exe = _build_reader2(tmp_path)
    f = tmp_path / "reads.fq.gz"
    with gzip.open(f, "wb") as g:
        g.write(open(os.path.join(H.GOLDEN, "gold_reads.fq"), "rb").read())
    got = subprocess.check_output([exe, str(f)]).decode()
    assert got == open(os.path.join(H.GOLDEN, "gold_reads_fq.parsed")).read()


@pytest.mark.parametrize("preset,tag,flags", [("illumina", "x_soft", [0, 1, 0, 0]),
                                               ("illuminapaired", "nosec_soft", [1, 1, 1, 1]),
                                               ("pacbio", "x_nosupp", [0, 0, 0, 1])])
def test_sam_writer_options_match_reference_writers(tmp_path, preset, tag, flags):
    """The writers' options ("Use M in CIGAR" off -> =/X, "Soft clip", "Omit Secondary / Supplementary Alignments"):
    tests/golden/gold_<preset>_<tag>.sam written by the reference's FileWriter / PairedFileWriter with them set."""
    exe = str(tmp_path / "test_sam")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(H.ROOT, "tests", "cpp", "test_sam.cpp"),
                           "-L" + os.path.join(H.ROOT, "ma_b200"), "-lma_b200",
                           "-Wl,-rpath," + os.path.join(H.ROOT, "ma_b200")])
    paired = preset == "illuminapaired"
    got = subprocess.run([exe, PC.GOLD_PREFIX, "1" if paired else "0"] + [str(f) for f in flags],
                         input=records(preset, paired).encode(), capture_output=True, check=True).stdout.decode()
    exp = open(os.path.join(H.GOLDEN, "gold_%s_%s.sam" % (preset, tag))).read()
    gl, el = got.splitlines(), exp.splitlines()
    for i, (a, b) in enumerate(zip(gl, el)):
        assert a == b, (i, a[:160], b[:160])
    assert len(gl) == len(el)
