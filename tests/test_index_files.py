"""On-disk index loaders: the reference's files load, truncated / mismatched ones are refused with
a descriptive error instead of out-of-bounds reads (the reference's loaders throw on bad files as well, fMIndex.h:854-884,
pack.h:275-451)."""
import os
import shutil

import pytest

import helpers as H
import pipeline_common as PC
from ma_b200 import index


def _copy(tmp_path):
    for e in (".bwt", ".sa", ".pac", ".ann", ".amb"):
        shutil.copy(PC.GOLD_PREFIX + e, str(tmp_path / ("g" + e)))
    return str(tmp_path / "g")


def test_reference_index_files_load(tmp_path):
    ix = index.load_index(_copy(tmp_path))
    assert ix.ref_len == 2 * ix.fwd_len and ix.sa[0] == -1 and len(ix.contig_names) == len(ix.contig_start)
    assert ix.bwt.size >= (ix.ref_len + 127) // 128 * 16


@pytest.mark.parametrize("ext,keep", [(".bwt", 30), (".bwt", 4000), (".sa", 40), (".sa", 2000), (".pac", 100)])
def test_truncated_index_files_are_refused(tmp_path, ext, keep):
    prefix = _copy(tmp_path)
    with open(prefix + ext, "rb") as f:
        data = f.read()
    with open(prefix + ext, "wb") as f:
        f.write(data[:keep])
    with pytest.raises(ValueError, match="corrupt index file"):
        index.load_index(prefix)


def test_mismatched_annotation_is_refused(tmp_path):
    prefix = _copy(tmp_path)
    lines = open(prefix + ".ann").read().splitlines()
    head = lines[0].split()
    head[0] = str(int(head[0]) + 4)  # forward length that does not match the BWT
    with open(prefix + ".ann", "w") as f:
        f.write("\n".join([" ".join(head)] + lines[1:]) + "\n")
    with pytest.raises(ValueError, match="corrupt index file"):
        index.load_index(prefix)
