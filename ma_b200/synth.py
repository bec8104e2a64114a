"""Deterministic synthetic genomes and simulated reads for the BASELINE.json configs.

The reference ships no simulator; every test and the bench use this one generator so that
the oracle, the CUDA path and the judge see identical inputs.  Everything is numpy-vectorised so that the
100 Mbp / 2 M-read configuration is generated in seconds.

Nucleotide code: A=0 C=1 G=2 T=3 N=4 (the reference's NucSeq code, libs/ma/src/container/nucSeq.cpp:17-28).
"""
from __future__ import annotations

import numpy as np

_ALPHA = np.frombuffer(b"ACGTN", dtype=np.uint8)


def random_genome(contig_lengths, seed: int):
    """Returns a list of uint8 arrays (codes 0..3), i.i.d. uniform."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [rng.integers(0, 4, size=int(n), dtype=np.uint8) for n in contig_lengths]


def codes_to_text(codes: np.ndarray) -> str:
    return _ALPHA[codes].tobytes().decode("ascii")


def text_to_codes(text: str) -> np.ndarray:
    lut = np.full(256, 4, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
        lut[c + 32] = i
    return lut[np.frombuffer(text.encode("ascii"), dtype=np.uint8)]


def write_genome_txt(path: str, contigs, names=None) -> None:
    """Format read by `oracle/_ref/ref_dump index`: '>name' line, then the sequence on one line."""
    with open(path, "w") as f:
        for i, c in enumerate(contigs):
            f.write(">%s\n" % (names[i] if names else "chr%d" % (i + 1)))
            f.write(codes_to_text(c))
            f.write("\n")


def revcomp(codes: np.ndarray) -> np.ndarray:
    out = codes[..., ::-1].copy()
    m = out < 4
    out[m] = 3 - out[m]
    return out


def simulate_reads(contigs, n_reads: int, read_len: int, seed: int, sub_rate=0.008, ins_rate=0.001, del_rate=0.001,
                   strand_both=True, chunk=200_000, flat=None):
    """Single-end reads: uniform windows, 50 % reverse-complemented, per-base substitution / insertion / deletion.

    Returns (reads[n, read_len] uint8, contig_id[n], pos[n], is_rev[n]).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.array([len(c) for c in contigs], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)])
    genome = flat if flat is not None else (np.concatenate(contigs) if len(contigs) > 1 else contigs[0])
    pad = max(16, int(read_len * (del_rate + ins_rate) * 8) + 16)
    tlen = read_len + pad
    usable = np.maximum(lens - tlen, 1)
    reads = np.empty((n_reads, read_len), dtype=np.uint8)
    cid_all = np.empty(n_reads, dtype=np.int64)
    pos_all = np.empty(n_reads, dtype=np.int64)
    rev_all = np.empty(n_reads, dtype=bool)
    for lo in range(0, n_reads, chunk):
        n = min(chunk, n_reads - lo)
        cid = rng.choice(len(contigs), size=n, p=usable / usable.sum())
        pos = (rng.random(n) * usable[cid]).astype(np.int64)
        rev = rng.random(n) < 0.5 if strand_both else np.zeros(n, dtype=bool)
        idx = (starts[cid] + pos)[:, None] + np.arange(tlen)[None, :]
        tmpl = genome[np.minimum(idx, len(genome) - 1)]
        tmpl[rev] = revcomp(tmpl[rev])  # template of a reverse read = revcomp of the window
        u = rng.random((n, tlen))
        sub = u < sub_rate
        dele = (u >= sub_rate) & (u < sub_rate + del_rate)
        ins = (u >= sub_rate + del_rate) & (u < sub_rate + del_rate + ins_rate)
        shift = rng.integers(1, 4, size=(n, tlen), dtype=np.uint8)
        base = np.where(sub, (tmpl + shift) & 3, tmpl).astype(np.uint8)
        ins_base = rng.integers(0, 4, size=(n, tlen), dtype=np.uint8)
        # output length contributed by each template base: 0 (deleted), 1, or 2 (inserted base first)
        olen = np.where(dele, 0, np.where(ins, 2, 1)).astype(np.int32)
        end = np.cumsum(olen, axis=1)
        beg = end - olen
        out = np.full((n, read_len + 2), 4, dtype=np.uint8)
        rows = np.broadcast_to(np.arange(n)[:, None], (n, tlen))
        # inserted base goes to position beg, the template base to end-1
        m_ins = ins & (beg < read_len)
        out[rows[m_ins], beg[m_ins]] = ins_base[m_ins]
        m_b = (~dele) & (end - 1 < read_len)
        out[rows[m_b], (end - 1)[m_b]] = base[m_b]
        reads[lo:lo + n] = out[:, :read_len]
        cid_all[lo:lo + n] = cid
        pos_all[lo:lo + n] = pos
        rev_all[lo:lo + n] = rev
    assert (reads < 4).all()
    return reads, cid_all, pos_all, rev_all


def simulate_pairs(contigs, n_pairs: int, read_len: int, seed: int, sub_rate=0.01, indel_rate=0.01, ins_mean=400,
                   ins_sd=50, ins_min=300, ins_max=700, chunk=200_000, flat=None):
    """2x read_len FR pairs, insert size N(ins_mean, ins_sd) clipped to [ins_min, ins_max].

    Returns (mate1[n, L], mate2[n, L], contig_id, frag_pos, frag_len, frag_is_rev).  mate1 is the fragment's
    5' end on its strand, mate2 is the reverse complement of the fragment's 3' end.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.array([len(c) for c in contigs], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)])
    genome = flat if flat is not None else (np.concatenate(contigs) if len(contigs) > 1 else contigs[0])
    pad = 32
    usable = np.maximum(lens - ins_max - pad, 1)
    m1 = np.empty((n_pairs, read_len), dtype=np.uint8)
    m2 = np.empty((n_pairs, read_len), dtype=np.uint8)
    meta = [np.empty(n_pairs, dtype=np.int64) for _ in range(3)]
    rev_all = np.empty(n_pairs, dtype=bool)
    ir, dr = indel_rate / 2, indel_rate / 2

    def mutate(tmpl):
        n, tlen = tmpl.shape
        u = rng.random((n, tlen))
        sub = u < sub_rate
        dele = (u >= sub_rate) & (u < sub_rate + dr)
        ins = (u >= sub_rate + dr) & (u < sub_rate + dr + ir)
        shift = rng.integers(1, 4, size=(n, tlen), dtype=np.uint8)
        base = np.where(sub, (tmpl + shift) & 3, tmpl).astype(np.uint8)
        ins_base = rng.integers(0, 4, size=(n, tlen), dtype=np.uint8)
        olen = np.where(dele, 0, np.where(ins, 2, 1)).astype(np.int32)
        end = np.cumsum(olen, axis=1)
        beg = end - olen
        out = np.full((n, read_len + 2), 4, dtype=np.uint8)
        rows = np.broadcast_to(np.arange(n)[:, None], (n, tlen))
        m_ins = ins & (beg < read_len)
        out[rows[m_ins], beg[m_ins]] = ins_base[m_ins]
        m_b = (~dele) & (end - 1 < read_len)
        out[rows[m_b], (end - 1)[m_b]] = base[m_b]
        return out[:, :read_len]

    tlen = read_len + pad
    for lo in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - lo)
        cid = rng.choice(len(contigs), size=n, p=usable / usable.sum())
        pos = (rng.random(n) * usable[cid]).astype(np.int64)
        flen = np.clip(np.rint(rng.normal(ins_mean, ins_sd, n)), ins_min, ins_max).astype(np.int64)
        rev = rng.random(n) < 0.5
        g0 = starts[cid] + pos
        # left window (forward strand) and right window (revcomp of the fragment end), each tlen long
        left = genome[np.minimum(g0[:, None] + np.arange(tlen)[None, :], len(genome) - 1)]
        ridx = (g0 + flen - 1)[:, None] - np.arange(tlen)[None, :]
        right = 3 - genome[np.clip(ridx, 0, len(genome) - 1)]
        a = mutate(left)
        b = mutate(right.astype(np.uint8))
        # forward fragment: mate1 = left, mate2 = right; reverse fragment: swapped roles
        m1[lo:lo + n] = np.where(rev[:, None], b, a)
        m2[lo:lo + n] = np.where(rev[:, None], a, b)
        meta[0][lo:lo + n] = cid
        meta[1][lo:lo + n] = pos
        meta[2][lo:lo + n] = flen
        rev_all[lo:lo + n] = rev
    assert (m1 < 4).all() and (m2 < 4).all()
    return m1, m2, meta[0], meta[1], meta[2], rev_all


def simulate_long_reads(contigs, n_reads: int, read_len: int, seed: int, sub_rate=0.04, ins_rate=0.04, del_rate=0.04,
                        flat=None):
    """PacBio-like reads (config 4). Same generator as simulate_reads with higher rates."""
    return simulate_reads(contigs, n_reads, read_len, seed, sub_rate, ins_rate, del_rate, chunk=2000, flat=flat)


# ---- fixed read SETS that can be generated shard by shard (BASELINE configs[2] / configs[3]: one set of 10 M pairs /
# 100 k long reads split over 1, 2, 4 or 8 GPUs): block b of the set is an independent stream seeded with (seed, b),
# so a rank generates only the blocks its shard overlaps, in parallel worker processes.
PAIR_BLOCK = 250_000
LONG_BLOCK = 2_000
_BLOCK_CTX = None


def _block_seed(seed: int, b: int) -> int:
    return seed * 1_000_003 + b


def _pair_block(b):
    contigs, flat, read_len, seed = _BLOCK_CTX
    m1, m2, *_ = simulate_pairs(contigs, PAIR_BLOCK, read_len, _block_seed(seed, b), flat=flat)
    out = np.empty((2 * PAIR_BLOCK, read_len), dtype=np.uint8)  # mates interleaved: read 2i, 2i+1 = pair i
    out[0::2], out[1::2] = m1, m2
    return out


def _long_block(b):
    contigs, flat, read_len, seed = _BLOCK_CTX
    return simulate_long_reads(contigs, LONG_BLOCK, read_len, _block_seed(seed, b), flat=flat)[0]


def _blocks(fn, first, last, per_block, ctx, workers):
    """Units [first, last) of a block-structured set; `workers` forked processes (fork: the genome is shared)."""
    global _BLOCK_CTX
    _BLOCK_CTX = ctx
    bs = list(range(first // per_block, (last + per_block - 1) // per_block)) if last > first else []
    try:
        if workers > 1 and len(bs) > 1:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(min(workers, len(bs))) as pool:
                parts = pool.map(fn, bs)
        else:
            parts = [fn(b) for b in bs]
    finally:
        _BLOCK_CTX = None
    if not parts:
        return np.empty((0, ctx[2]), dtype=np.uint8)
    allr = np.concatenate(parts) if len(parts) > 1 else parts[0]
    return allr, bs[0] * per_block


def pair_set_reads(contigs, first_read: int, last_read: int, read_len: int, seed: int, workers: int = 1, flat=None):
    """Reads [first_read, last_read) (mates interleaved) of the fixed pair set (seed): block b = PAIR_BLOCK pairs."""
    if flat is None:
        flat = np.concatenate(contigs) if len(contigs) > 1 else contigs[0]
    r = _blocks(_pair_block, first_read // 2, (last_read + 1) // 2, PAIR_BLOCK, (contigs, flat, read_len, seed), workers)
    if isinstance(r, np.ndarray):
        return r
    allr, base_pair = r
    return allr[first_read - 2 * base_pair:last_read - 2 * base_pair]


def long_set_reads(contigs, first_read: int, last_read: int, read_len: int, seed: int, workers: int = 1, flat=None):
    """Reads [first_read, last_read) of the fixed long-read set (seed): block b = LONG_BLOCK reads."""
    if flat is None:
        flat = np.concatenate(contigs) if len(contigs) > 1 else contigs[0]
    r = _blocks(_long_block, first_read, last_read, LONG_BLOCK, (contigs, flat, read_len, seed), workers)
    if isinstance(r, np.ndarray):
        return r
    allr, base = r
    return allr[first_read - base:last_read - base]


def write_reads_txt(path: str, reads: np.ndarray) -> None:
    """Format read by `ref_dump align|bench`: one read per line."""
    txt = _ALPHA[reads]
    with open(path, "wb") as f:
        nl = np.full((txt.shape[0], 1), 10, dtype=np.uint8)
        f.write(np.concatenate([txt, nl], axis=1).tobytes())


def mutate_pairs_for_dp(n: int, length: int, divergence: float, seed: int):
    """DP-only sweep inputs (config 5): target random, query = mutated copy (1/3 subst, 1/3 ins, 1/3 del)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for _ in range(n):
        t = rng.integers(0, 4, size=length, dtype=np.uint8)
        u = rng.random(length)
        r = divergence / 3
        q = []
        for i in range(length):
            if u[i] < r:
                q.append((int(t[i]) + int(rng.integers(1, 4))) & 3)
            elif u[i] < 2 * r:
                continue
            elif u[i] < 3 * r:
                q.append(int(rng.integers(0, 4)))
                q.append(int(t[i]))
            else:
                q.append(int(t[i]))
        out.append((np.array(q, dtype=np.uint8), t))
    return out
