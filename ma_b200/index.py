"""Readers / writers for the reference's on-disk index files (row N4 of SURVEY.md §8(f)).

  <prefix>.bwt  primary (i64), L2[1..4] (4 x u64), BWT words           FMIndex::vSaveBWT, fMIndex.h:515-526
  <prefix>.sa   primary, L2[1..4], sa_intv (i32), seq_len (u64), sa[1:] FMIndex::vSaveSuffixArray, fMIndex.h:529-549
  <prefix>.pac  2-bit packed forward strand (+ size bytes)               Pack::vStorePack, pack.h:181-222
  <prefix>.ann  contig table (text)                                       Pack::vStoreCollectionDescripton, pack.h:230-269
  <prefix>.amb  hole table (text)

The arrays are passed unchanged to ma_b200_index_upload: the HBM layout of the occurrence table is the reference's
64-byte block layout.
"""
from __future__ import annotations

import os

import numpy as np


class Index:
    def __init__(self):
        self.bwt = None          # uint32 words
        self.L2 = None           # int64[5]
        self.primary = 0
        self.ref_len = 0         # forward + reverse
        self.sa = None           # int64, sa[0] = -1
        self.sa_intv = 32
        self.pac = None          # uint8
        self.fwd_len = 0
        self.contig_names = []
        self.contig_start = None  # int64
        self.contig_len = None    # int64

    def forward_codes(self) -> np.ndarray:
        """Unpacks the forward strand (1 byte per base)."""
        n = self.fwd_len
        b = self.pac[:(n + 3) // 4]
        out = np.empty((len(b), 4), dtype=np.uint8)
        for j in range(4):
            out[:, j] = (b >> (6 - 2 * j)) & 3
        return out.reshape(-1)[:n]


def load_index(prefix: str) -> Index:
    ix = Index()
    with open(prefix + ".bwt", "rb") as f:
        raw = f.read()
    if len(raw) < 40 or (len(raw) - 40) % 4:
        raise ValueError("corrupt index file %s.bwt: shorter than its header or not a whole number of words" % prefix)
    ix.primary = int(np.frombuffer(raw[:8], dtype=np.int64)[0])
    ix.L2 = np.zeros(5, dtype=np.int64)
    ix.L2[1:] = np.frombuffer(raw[8:40], dtype=np.int64)
    ix.bwt = np.frombuffer(raw[40:], dtype=np.uint32).copy()
    ix.ref_len = int(ix.L2[4])
    with open(prefix + ".sa", "rb") as f:
        raw = f.read()
    if len(raw) < 52:
        raise ValueError("corrupt index file %s.sa: shorter than its header" % prefix)
    ix.sa_intv = int(np.frombuffer(raw[40:44], dtype=np.int32)[0])
    if ix.sa_intv <= 0 or ix.ref_len <= 0:
        raise ValueError("corrupt index files %s: sampling interval / reference length" % prefix)
    n_sa = (ix.ref_len + ix.sa_intv) // ix.sa_intv
    if len(raw) < 52 + 8 * (n_sa - 1) or ix.bwt.size < (ix.ref_len + 127) // 128 * 16:
        raise ValueError("corrupt index files %s: .sa / .bwt shorter than the reference length needs" % prefix)
    ix.sa = np.empty(n_sa, dtype=np.int64)
    ix.sa[0] = -1
    ix.sa[1:] = np.frombuffer(raw[52:52 + 8 * (n_sa - 1)], dtype=np.int64)
    with open(prefix + ".ann") as f:
        head = f.readline().split()
        ix.fwd_len, n_seq = int(head[0]), int(head[1])
        starts, lens = [], []
        for _ in range(n_seq):
            l1 = f.readline().split()
            ix.contig_names.append(l1[1])
            l2 = f.readline().split()
            starts.append(int(l2[0]))
            lens.append(int(l2[1]))
    ix.contig_start = np.array(starts, dtype=np.int64)
    ix.contig_len = np.array(lens, dtype=np.int64)
    with open(prefix + ".pac", "rb") as f:
        ix.pac = np.frombuffer(f.read(), dtype=np.uint8)[:(ix.fwd_len + 3) // 4].copy()
    if ix.ref_len != 2 * ix.fwd_len or len(ix.pac) < (ix.fwd_len + 3) // 4:
        raise ValueError("corrupt index files %s: .ann / .pac do not match the BWT" % prefix)
    return ix


def pack_forward(codes: np.ndarray) -> np.ndarray:
    """2-bit packing of Pack::vSetNucleotideOnPos (pack.h:162-167): base i in byte i>>2, shift (~i & 3) << 1."""
    n = len(codes)
    pad = (-n) % 4
    c = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4).astype(np.uint8)
    return ((c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]).astype(np.uint8)


def store_index(ix: Index, prefix: str) -> None:
    """Writes the index in the reference's file formats so that the unmodified reference can load it."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".bwt", "wb") as f:
        f.write(np.int64(ix.primary).tobytes())
        f.write(ix.L2[1:5].astype(np.int64).tobytes())
        f.write(ix.bwt.astype(np.uint32).tobytes())
    with open(prefix + ".sa", "wb") as f:
        f.write(np.int64(ix.primary).tobytes())
        f.write(ix.L2[1:5].astype(np.int64).tobytes())
        f.write(np.int32(ix.sa_intv).tobytes())
        f.write(np.uint64(ix.ref_len).tobytes())
        f.write(ix.sa[1:].astype(np.int64).tobytes())
    with open(prefix + ".pac", "wb") as f:
        f.write(ix.pac[:(ix.fwd_len + 3) // 4].tobytes())
        if ix.fwd_len % 4 == 0:
            f.write(b"\0")
        f.write(bytes([ix.fwd_len % 4]))
    with open(prefix + ".ann", "w") as f:
        f.write("%d %d %d\n" % (ix.fwd_len, len(ix.contig_names), 11))
        for name, s, l in zip(ix.contig_names, ix.contig_start, ix.contig_len):
            f.write("0 %s synthetic\n%d %d 0\n" % (name, s, l))
    with open(prefix + ".amb", "w") as f:
        f.write("%d %d 0\n" % (ix.fwd_len, len(ix.contig_names)))
