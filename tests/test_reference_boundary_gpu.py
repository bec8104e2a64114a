"""The drop-in boundary compiled against the REAL reference: integration/gpu_align.h —
`GpuIndex`, `GpuAlign : libMS::Module<...>` and the per-read blocking batcher `GpuAlignPerRead` — is built by
integration/Makefile against /root/reference's headers (libs/ms/inc/ms/module/module.h:63-122) and linked with the
compiled reference (oracle/_ref/libma_ref.so) and libma_b200.so. integration/ref_gpu_sam wires it the way
setUpCompGraph does (libs/ma/src/util/export.cpp:99-126): the reference's FileReader / NucSeq and presets in front,
the reference's FileWriter / PairedFileWriter behind. The SAM text must equal the golden files the pure reference
chain wrote (tests/golden/make_golden_pipeline.py)."""
import os
import subprocess

import pytest

import helpers as H
import pipeline_common as PC

pytestmark = pytest.mark.gpu
EXE = os.path.join(H.ORACLE_DIR, "_ref", "ref_gpu_sam")


def _need_exe():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/ref_gpu_sam not built (needs the reference sources: make -C integration)")


@pytest.mark.parametrize("preset", ["illumina", "pacbio", "illuminapaired"])
def test_gpu_align_module_in_reference_graph_writes_reference_sam(preset, tmp_path):
    _need_exe()
    out = str(tmp_path / "o.sam")
    subprocess.check_call([EXE, "batch", PC.GOLD_PREFIX, PC.gold_reads(preset), preset, out, str(PC.SRAND)])
    assert open(out).read() == open(os.path.join(H.GOLDEN, "gold_%s.sam" % preset)).read()


def test_fasta_reads_through_reference_file_reader(tmp_path):
    _need_exe()
    out = str(tmp_path / "o.sam")
    subprocess.check_call([EXE, "batch", PC.GOLD_PREFIX, os.path.join(H.GOLDEN, "gold_reads.fa"), "illumina", out,
                           str(PC.SRAND)])
    assert open(out).read() == open(os.path.join(H.GOLDEN, "gold_fa_illumina.sam")).read()


def test_per_read_blocking_batcher_one_graph_thread_is_exact(tmp_path):
    """GpuAlignPerRead: execute( read ) per graph thread; with one thread the batches follow the file order, so the
    RANSAC streams (srand base + read index) and therefore the SAM file are the golden ones."""
    _need_exe()
    out = str(tmp_path / "o.sam")
    subprocess.check_call([EXE, "perread", PC.GOLD_PREFIX, PC.gold_reads("illumina"), "illumina", out, str(PC.SRAND),
                           "1", "16"])
    assert open(out).read() == open(os.path.join(H.GOLDEN, "gold_illumina.sam")).read()


def test_per_read_blocking_batcher_many_graph_threads(tmp_path):
    """Eight graph threads share the module (batches of up to 32 reads): every read is written exactly once and nobody
    waits for ever when the threads run out of reads at different times. (The order of the records and the RANSAC
    stream of a read depend on the thread interleaving, as the reference's own multi-threaded output order does.)"""
    _need_exe()
    out = str(tmp_path / "o.sam")
    subprocess.run([EXE, "perread", PC.GOLD_PREFIX, PC.gold_reads("illumina"), "illumina", out, str(PC.SRAND), "8",
                    "32"], check=True, timeout=120)
    gold = open(os.path.join(H.GOLDEN, "gold_illumina.sam")).read().splitlines()
    got = open(out).read().splitlines()
    primary = lambda lines: sorted(l.split("\t")[0] for l in lines if not l.startswith("@") and
                                   not int(l.split("\t")[1]) & 0x900)
    assert primary(got) == primary(gold)
    assert [l for l in got if l.startswith("@")] == [l for l in gold if l.startswith("@")]


def test_reference_computational_graph_with_gpu_module_one_thread_is_exact(tmp_path):
    """setUpCompGraphGpu: the reference's own graph (QueuePicker, Lock, FileReader, QueuePlacer, FileWriter,
    ProgressPrinter, UnLock wired with promiseMe and evaluated by BasePledge::simultaneousGet, as
    ExecutionContext::doAlign does, execution-context.h:291-406) with GpuAlignPerRead in place of the five CPU modules.
    One graph thread: batches follow the file order, the SAM file is the reference chain's."""
    _need_exe()
    out = str(tmp_path / "o.sam")
    subprocess.run([EXE, "graph", PC.GOLD_PREFIX, os.path.join(H.GOLDEN, "gold_reads.fa"), "illumina", out, str(PC.SRAND),
                    "1", "8"], check=True, timeout=120)
    assert open(out).read() == open(os.path.join(H.GOLDEN, "gold_fa_illumina.sam")).read()


def test_reference_computational_graph_with_gpu_module_many_threads(tmp_path):
    """Six graph threads on the shared module: every read is written exactly once, the run ends (no thread waits for a
    batch that never fills once the others have run out of reads)."""
    _need_exe()
    out = str(tmp_path / "o.sam")
    subprocess.run([EXE, "graph", PC.GOLD_PREFIX, os.path.join(H.GOLDEN, "gold_reads.fa"), "illumina", out, str(PC.SRAND),
                    "6", "16"], check=True, timeout=120)
    gold = open(os.path.join(H.GOLDEN, "gold_fa_illumina.sam")).read().splitlines()
    got = open(out).read().splitlines()
    primary = lambda lines: sorted(l.split("\t")[0] for l in lines if not l.startswith("@") and
                                   not int(l.split("\t")[1]) & 0x900)
    assert primary(got) == primary(gold)
