"""Seeded generator of DP-only test problems (BASELINE.json config 5 shapes, scaled down for tests)."""
import numpy as np

GLOBAL, EXT, EXT_RIGHT = 0, 0x40, 0x40 | 0x02 | 0x80


def mutate(rng, t, div):
    u = rng.random(len(t))
    sub = rng.integers(1, 4, size=len(t))
    ins = rng.integers(0, 4, size=len(t))
    q = []
    for i, b in enumerate(t):
        if u[i] < div / 3:
            q.append((int(b) + int(sub[i])) & 3)
        elif u[i] < 2 * div / 3:
            continue
        elif u[i] < div:
            q.append(int(ins[i]))
            q.append(int(b))
        else:
            q.append(int(b))
    return np.array(q, dtype=np.uint8)


def random_pairs(n, seed, lengths=(1, 2, 3, 5, 8, 15, 16, 17, 31, 33, 50, 100, 150, 300, 700, 1000, 1400, 2000),
                 with_n=True):
    """Mixed bag: global gap fills, left/right extensions with random tails, narrow-band (out-of-band) globals."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pairs = []
    for _ in range(n):
        L = int(rng.choice(lengths))
        t = rng.integers(0, 4, size=L, dtype=np.uint8)
        div = float(rng.choice([0.0, 0.01, 0.05, 0.15, 0.5]))
        q = mutate(rng, t, div)
        if len(q) == 0:
            q = np.array([1], dtype=np.uint8)
        mode = int(rng.integers(0, 4))
        if with_n and rng.random() < 0.1:
            q[rng.integers(0, len(q))] = 4
        if mode == 0:
            w, zd, fl = max(int(rng.choice([1, 5, 20, 64])), abs(len(t) - len(q)) + 10), -1, GLOBAL
        elif mode == 1:
            w, zd, fl = int(rng.choice([16, 64, 512])), 200, EXT
            t = np.concatenate([t, rng.integers(0, 4, size=int(rng.integers(0, 400)), dtype=np.uint8)])
        elif mode == 2:
            w, zd, fl = int(rng.choice([16, 64, 512])), int(rng.choice([20, 200])), EXT_RIGHT
            t = np.concatenate([t, rng.integers(0, 4, size=int(rng.integers(0, 400)), dtype=np.uint8)])
        else:
            w, zd, fl = int(rng.choice([3, 10, 30])), -1, GLOBAL
        pairs.append((w, zd, fl, q, t))
    return pairs


def sweep_pairs(n, length, w, mode, div, seed):
    """Config-5 point: n pairs of one (length, band, mode, divergence)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pairs = []
    for _ in range(n):
        t = rng.integers(0, 4, size=length, dtype=np.uint8)
        q = mutate(rng, t, div)
        if len(q) == 0:
            q = t[:1].copy()
        if mode == GLOBAL:
            ww = max(w, abs(len(t) - len(q)) + 10)  # ksw_simplified, needlemanWunsch.cpp:70-71
            pairs.append((ww, -1, GLOBAL, q, t))
        else:
            pairs.append((w, 200, mode, q, t))
    return pairs
