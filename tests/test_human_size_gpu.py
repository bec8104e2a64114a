"""BASELINE.json configs[2] / configs[3] on the human-sized genome (31 contigs x 100 Mbp, 6.2 G BWT symbols; SA values and
occurrence counts beyond 2^32): the index built on the GPU and the device path on it against the UNMODIFIED reference.

tests/golden/human_size_sha1.json (tests/golden/make_golden_human_size.py) holds, from the reference's own builder
(FMIndex(pPack) -> bwtLarge, 6 550 s on one host core) and its modules on that index:
  * the SHA-1 of the index arrays (BWT words with their occurrence blocks, SA samples, packed forward strand),
  * the SHA-1 of every stage's dump, the mapping qualities and the pairing of 3 000 simulated Illumina reads,
  * the SHA-1 of the SAM text of those reads (PairedFileWriter),
  * the same for 200 simulated 10 kbp PacBio reads (PacBio preset).
The GPU builder (3.7 s) must reproduce the index bit for bit, and the device path every dump. No tolerance."""
import hashlib
import json
import os

import numpy as np
import pytest

import helpers as H
import pipeline_common as PC
from ma_b200 import api, synth

pytestmark = pytest.mark.gpu
PIN_FILE = os.path.join(H.GOLDEN, "human_size_sha1.json")


def _sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a)).hexdigest()


def _sha1_i64(a):
    return hashlib.sha1(np.ascontiguousarray(np.asarray(a, dtype=np.int64)).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def human():
    if not os.path.exists(PIN_FILE):
        pytest.skip("golden hashes of the human-sized configuration not generated")
    pin = json.load(open(PIN_FILE))
    genome = synth.random_genome([pin["contig_len"]] * pin["n_contigs"], pin["genome_seed"])
    lens = np.array([len(c) for c in genome], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    fwd = np.concatenate(genome)
    ctx = api.Context(0, "illumina_paired")
    ctx.index_build(fwd, starts, lens)
    yield pin, genome, fwd, ctx
    ctx.close()


def test_gpu_built_human_index_equals_reference_builder(human):
    pin, genome, fwd, ctx = human
    ix = ctx.index_download()
    assert int(ix.primary) == pin["index"]["primary"] and [int(x) for x in ix.L2] == pin["index"]["L2"]
    assert int(ix.bwt.size) == pin["index"]["n_words"] and int(ix.sa.size) == pin["index"]["n_sa"]
    assert int(ix.sa.max()) >= 2 ** 32  # the regime the 100 Mbp configuration never reaches
    got = {"bwt": _sha1(ix.bwt), "sa": _sha1(ix.sa), "pac": _sha1(ix.pac[:(len(fwd) + 3) // 4])}
    assert got == pin["index"]["sha1"]


def test_illumina_sample_on_human_index_equals_reference(human):
    pin, genome, fwd, ctx = human
    m1, m2, *_ = synth.simulate_pairs(genome, pin["n_pairs_simulated"], 150, pin["read_seed"], flat=fwd)
    n = pin["n_reads"]
    reads = np.empty((n, 150), dtype=np.uint8)
    reads[0::2], reads[1::2] = m1[:n // 2], m2[:n // 2]
    p = api.preset("illumina_paired")
    p.srand_base = pin["srand_base"]
    ctx.set_params(p)
    got = PC.gpu_stage_dump(ctx, reads)
    got.update(PC.gpu_mapq_dump(ctx, reads, p))
    diff = [k for k, h in pin["sha1"].items() if _sha1_i64(got[k]) != h]
    assert not diff, diff


def test_pacbio_sample_on_human_index_equals_reference(human):
    pin, genome, fwd, ctx0 = human
    lp = pin["pacbio"]
    reads, *_ = synth.simulate_long_reads(genome, lp["n_reads"], lp["read_len"], lp["seed"], flat=fwd)
    # a second context on the same device would build the index again: switch the preset of the shared one instead
    p = api.preset("pacbio")
    p.srand_base = pin["srand_base"]
    ctx0.set_params(p)
    got = PC.gpu_stage_dump(ctx0, reads, keep_segments=16384)
    got.update(PC.gpu_mapq_dump(ctx0, reads, p))
    diff = [k for k, h in lp["sha1"].items() if _sha1_i64(got[k]) != h]
    assert not diff, diff
