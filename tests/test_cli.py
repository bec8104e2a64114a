"""maCMD_b200 (ma_b200/cli): the reference CLI's options for the alignment path (cmdMa.cpp:252-431), FASTA/FASTQ in,
SAM out, compared with the SAM text the UNMODIFIED reference wrote for the same files (tests/golden/make_golden_reads.py:
FileReader -> modules -> FileWriter / PairedFileWriter)."""
import os
import subprocess

import pytest

import helpers as H
import pipeline_common as PC

CLI = os.path.join(H.ROOT, "ma_b200", "cli", "maCMD_b200")


def build_cli():
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(CLI)])
    return CLI


def test_cli_builds_and_rejects_bad_usage(tmp_path):
    """CPU: the front end links against the C-ABI library; bad options fail with a message, and without a CUDA device the
    program stops with an error (no CPU fallback)."""
    exe = build_cli()
    r = subprocess.run([exe, "--NoSuchOption", "1"], capture_output=True)
    assert r.returncode == 1 and b"unknown option" in r.stderr
    r = subprocess.run([exe, "-x", PC.GOLD_PREFIX], capture_output=True)
    assert r.returncode == 1 and b"usage" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "-x", PC.GOLD_PREFIX, "-i", os.path.join(H.GOLDEN, "gold_reads.fa")],
                           capture_output=True)
        assert r.returncode == 1 and b"no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1000000, 8])
def test_cli_interleaved_pairs_sam_matches_reference(tmp_path, batch):
    exe = build_cli()
    out = str(tmp_path / "o.sam")
    subprocess.check_call([exe, "-x", PC.GOLD_PREFIX, "-i", os.path.join(H.GOLDEN, "gold_reads.fq"), "-p",
                           "Illumina_Paired", "--Interleaved", "--Srand", str(PC.SRAND), "--Batch", str(batch),
                           "-o", out, "-t", "4"])
    assert open(out).read() == open(os.path.join(H.GOLDEN, "gold_fq_illuminapaired.sam")).read()


@pytest.mark.gpu
def test_cli_mate_files_switch_pairing_on(tmp_path):
    """-i reads -m mates with the unpaired Illumina presetting == Illumina_Paired (cmdMa.cpp:323-330); two input files."""
    exe = build_cli()
    recs = open(os.path.join(H.GOLDEN, "gold_reads.fq"), "rb").read().replace(b"\r\n", b"\n").replace(b"\n\n", b"\n")
    parts = [b"@pair" + r for r in recs.split(b"@pair")[1:]]
    assert len(parts) == 40
    names = []
    for k, (lo, hi) in enumerate([(0, 14), (14, 40)]):  # two files per mate: the comma separated list of the reference
        for m in (0, 1):
            fn = str(tmp_path / ("m%d_%d.fq" % (m, k)))
            open(fn, "wb").write(b"".join(parts[lo + m:hi:2]))
            names.append(fn)
    # the RANSAC stream of a read is srand(base + index in input order): pairs are read 2k (from -i), 2k+1 (from -m)
    out = subprocess.check_output([exe, "-x", PC.GOLD_PREFIX, "-i", names[0] + "," + names[2], "-m",
                                   names[1] + "," + names[3], "-p", "Illumina", "--Srand", str(PC.SRAND)])
    assert out.decode() == open(os.path.join(H.GOLDEN, "gold_fq_illuminapaired.sam")).read()


@pytest.mark.gpu
def test_cli_fasta_single_end_sam_matches_reference(tmp_path):
    exe = build_cli()
    out = subprocess.check_output([exe, "-x", PC.GOLD_PREFIX, "-i", os.path.join(H.GOLDEN, "gold_reads.fa"), "-p",
                                   "Illumina", "--Srand", str(PC.SRAND), "--Batch", "6"])
    assert out.decode() == open(os.path.join(H.GOLDEN, "gold_fa_illumina.sam")).read()


@pytest.mark.gpu
@pytest.mark.parametrize("preset,extra,gold", [("Default", ["--Z_Drop_Inversions", "20"], "gold_inv_default_z20.sam"),
                                               ("Illumina", [], "gold_inv_illumina.sam")])
def test_cli_small_inversions_sam_matches_reference(tmp_path, preset, extra, gold):
    """--Detect_Small_Inversions == the reference's SmallInversions between MappingQuality and the
    writer (host glue of include/ma_b200_modules.hpp, its DP calls as one ma_b200_ksw_batch on the GPU)."""
    exe = build_cli()
    out = subprocess.check_output([exe, "-x", PC.GOLD_PREFIX, "-i", os.path.join(H.GOLDEN, "gold_reads_inv.fa"), "-p",
                                   preset, "--Srand", str(PC.SRAND), "--Detect_Small_Inversions", "true"] + extra)
    exp = open(os.path.join(H.GOLDEN, gold)).read()
    assert sum(1 for l in exp.splitlines() if not l.startswith("@") and int(l.split("\t")[1]) & 0x800) >= 1
    assert out.decode() == exp


@pytest.mark.gpu
def test_cli_two_devices_keep_input_order(tmp_path):
    """--Devices 0,1: batches go to whichever GPU is free (index replicated, no collective); the SAM
    file is the same as with one device. Needs two GPUs (gpurun --gpus 2)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = build_cli()
    out = subprocess.check_output([exe, "-x", PC.GOLD_PREFIX, "-i", os.path.join(H.GOLDEN, "gold_reads.fq"), "-p",
                                   "Illumina_Paired", "--Interleaved", "--Srand", str(PC.SRAND), "--Batch", "4",
                                   "--Devices", "0,1"])
    assert out.decode() == open(os.path.join(H.GOLDEN, "gold_fq_illuminapaired.sam")).read()


@pytest.mark.gpu
@pytest.mark.parametrize("drop_from,message", [(1, b"fewer mates than reads"), (0, b"more mates than reads")])
def test_cli_unequal_mate_files_fail_loudly(tmp_path, drop_from, message):
    exe = build_cli()
    recs = open(os.path.join(H.GOLDEN, "gold_reads.fq"), "rb").read().replace(b"\r\n", b"\n").replace(b"\n\n", b"\n")
    parts = [b"@pair" + r for r in recs.split(b"@pair")[1:]]
    files = []
    for m in (0, 1):
        mine = parts[m::2]
        if m == drop_from:
            mine = mine[:-1]
        files.append(str(tmp_path / ("m%d.fq" % m)))
        open(files[-1], "wb").write(b"".join(mine))
    r = subprocess.run([exe, "-x", PC.GOLD_PREFIX, "-i", files[0], "-m", files[1], "-p", "Illumina", "-o",
                        str(tmp_path / "o.sam")], capture_output=True)
    assert r.returncode == 1 and message in r.stderr, r.stderr[-300:]


@pytest.mark.gpu
def test_cli_gzip_input(tmp_path):
    import gzip
    exe = build_cli()
    f = tmp_path / "reads.fq.gz"
    with gzip.open(f, "wb") as g:
        g.write(open(os.path.join(H.GOLDEN, "gold_reads.fq"), "rb").read())
    out = subprocess.check_output([exe, "-x", PC.GOLD_PREFIX, "-i", str(f), "-p", "Illumina_Paired", "--Interleaved",
                                   "--Srand", str(PC.SRAND)])
    assert out.decode() == open(os.path.join(H.GOLDEN, "gold_fq_illuminapaired.sam")).read()


@pytest.mark.gpu
def test_cli_writer_flags(tmp_path):
    """--Use_M_in_CIGAR false --Soft_clip == the reference's FileWriter with those flags (gold_illumina_x_soft.sam)."""
    exe = build_cli()
    fa = tmp_path / "r.fa"
    with open(fa, "w") as f:
        for i, l in enumerate(x.strip() for x in open(PC.gold_reads("illumina")) if x.strip()):
            f.write(">r%d\n%s\n" % (i, l))
    out = subprocess.check_output([exe, "-x", PC.GOLD_PREFIX, "-i", str(fa), "-p", "Illumina", "--Srand", str(PC.SRAND),
                                   "--Use_M_in_CIGAR", "false", "--Soft_clip"])
    assert out.decode() == open(os.path.join(H.GOLDEN, "gold_illumina_x_soft.sam")).read()


@pytest.mark.gpu
def test_cli_small_inversions_with_paired_reads(tmp_path):
    """Paired graph with "Detect Small Inversions" (export.cpp:176-184): MappingQuality on the device, SmallInversions
    per mate (host glue, DP on the device), PairedReads on the host (include/ma_b200_modules.hpp PairedReadsHost) — the
    inversion records enter the pairing and the pair's mapping quality."""
    exe = build_cli()
    out = subprocess.check_output([exe, "-x", PC.GOLD_PREFIX, "-i", os.path.join(H.GOLDEN, "gold_reads_inv_pairs.fa"),
                                   "-p", "Illumina_Paired", "--Interleaved", "--Srand", str(PC.SRAND),
                                   "--Detect_Small_Inversions", "--Z_Drop_Inversions", "20", "--Batch", "10"])
    assert out.decode() == open(os.path.join(H.GOLDEN, "gold_inv_illuminapaired_z20.sam")).read()
    # and it differs from the run without the module (the module is not a no-op on this set)
    plain = subprocess.check_output([exe, "-x", PC.GOLD_PREFIX, "-i", os.path.join(H.GOLDEN, "gold_reads_inv_pairs.fa"),
                                     "-p", "Illumina_Paired", "--Interleaved", "--Srand", str(PC.SRAND)])
    assert plain != out


@pytest.mark.gpu
def test_cli_create_index_writes_the_reference_files(tmp_path):
    """-X <fasta>,<folder>,<name> (cmdMa.cpp:332-345): the files of the reference's own builder for the golden genome,
    byte for byte (.ann up to its random seed field), and -x <name>.json loads them."""
    from ma_b200 import index
    exe = build_cli()
    ix = index.load_index(PC.GOLD_PREFIX)
    fwd = ix.forward_codes()
    with open(tmp_path / "genome.fa", "w") as f:
        for name, s, l in zip(ix.contig_names, ix.contig_start, ix.contig_len):
            seq = "".join("ACGT"[c] for c in fwd[int(s):int(s) + int(l)])
            f.write(">%s synthetic\n" % name)
            f.write("\n".join(seq[k:k + 70] for k in range(0, len(seq), 70)) + "\n")
    subprocess.check_call([exe, "-X", "%s,%s,%s" % (tmp_path / "genome.fa", tmp_path, "gold2")])
    for ext in (".bwt", ".sa", ".pac", ".amb"):
        assert open(str(tmp_path / "genome") + ext, "rb").read() == open(PC.GOLD_PREFIX + ext, "rb").read(), ext
    mine = open(str(tmp_path / "genome") + ".ann").read().split("\n")
    ref = open(PC.GOLD_PREFIX + ".ann").read().split("\n")
    assert mine[0].split()[:2] == ref[0].split()[:2] and mine[1:] == ref[1:]
    out = subprocess.check_output([exe, "-x", str(tmp_path / "gold2.json"), "-i", os.path.join(H.GOLDEN, "gold_reads.fa"),
                                   "-p", "Illumina", "--Srand", str(PC.SRAND)])
    assert out.decode() == open(os.path.join(H.GOLDEN, "gold_fa_illumina.sam")).read()
    r = subprocess.run([exe, "-X", "%s,%s,%s" % (os.path.join(H.GOLDEN, "gold_reads.fa"), tmp_path, "bad")],
                       capture_output=True)
    assert r.returncode == 1 and b"without N" in r.stderr  # IUPAC codes in that file
