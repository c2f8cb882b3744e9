"""Seeded generators of kernel-level test cases shared by the CPU (oracle pinning) and GPU (parity) tests."""
import numpy as np

PARAMS = [  # (a, b, o_del, e_del, o_ins, e_ins)
    (1, 4, 6, 1, 6, 1), (1, 4, 6, 1, 6, 1), (2, 6, 8, 2, 10, 1), (1, 9, 1, 1, 1, 1), (1, 1, 1, 1, 1, 1), (3, 12, 18, 3, 18, 3), (1, 4, 10, 2, 4, 3),
]


def mutate(rng, t, sub, indel, n_rate=0.0):
    out = []
    i = 0
    while i < len(t):
        r = rng.random()
        if r < indel / 2:
            i += int(rng.integers(1, 6))
            continue
        if r < indel:
            out.extend(rng.integers(0, 4, size=int(rng.integers(1, 6))).tolist())
        c = int(t[i])
        if rng.random() < sub:
            c = (c + int(rng.integers(1, 4))) & 3
        if rng.random() < n_rate:
            c = 4
        out.append(c)
        i += 1
    return np.array(out, dtype=np.uint8)


def extend_cases(seed, n, max_q=150, max_t=400):
    """-> list of dict(q, t, params, w, end_bonus, zdrop, h0)"""
    rng = np.random.default_rng(seed)
    cases = []
    for k in range(n):
        tl = int(rng.integers(0, max_t + 1)) if rng.random() < 0.9 else int(rng.integers(0, 8))
        t = rng.integers(0, 4, size=tl).astype(np.uint8)
        style = rng.random()
        if style < 0.7:
            q = mutate(rng, t, rng.choice([0.0, 0.01, 0.05, 0.2]), rng.choice([0.0, 0.002, 0.02, 0.1]), rng.choice([0.0, 0.0, 0.01]))
        elif style < 0.85:
            q = rng.integers(0, 4, size=int(rng.integers(1, max_q + 1))).astype(np.uint8)
        else:  # good prefix then garbage (z-drop / clipping)
            cut = int(rng.integers(0, tl + 1))
            q = np.concatenate([mutate(rng, t[:cut], 0.01, 0.002), rng.integers(0, 4, size=int(rng.integers(0, 80))).astype(np.uint8)])
        ql = int(rng.integers(1, max_q + 1))
        q = q[:ql]
        if len(q) == 0:
            q = rng.integers(0, 4, size=1).astype(np.uint8)
        if rng.random() < 0.05 and len(t):
            t = t.copy(); t[rng.integers(0, len(t))] = 4
        cases.append(dict(q=q, t=t, params=PARAMS[int(rng.integers(0, len(PARAMS)))],
                          w=int(rng.choice([100, 100, 200, 50, 10, 3, 1])), end_bonus=int(rng.choice([5, 5, 0, 9])),
                          zdrop=int(rng.choice([100, 100, 0, 20, 5])), h0=int(rng.integers(1, 151))))
    return cases


def align_cases(seed, n, sixteen=False, gap_open_min=1):
    """mate-rescue-like local alignments: query 20..249 (8 bit) or up to 300 (16 bit), target window up to 900"""
    rng = np.random.default_rng(seed)
    cases = []
    pars = [p for p in PARAMS if p[2] >= gap_open_min and p[4] >= gap_open_min]
    for k in range(n):
        a, b, od, ed, oi, ei = pars[int(rng.integers(0, len(pars)))]
        tl = int(rng.integers(30, 900))
        t = rng.integers(0, 4, size=tl).astype(np.uint8)
        ql = int(rng.integers(20, 300 if sixteen else 250))
        if not sixteen:
            ql = min(ql, (249 // a))
        style = rng.random()
        if style < 0.75:
            s = int(rng.integers(0, max(1, tl - ql // 2)))
            q = mutate(rng, t[s:s + ql + 20], rng.choice([0.0, 0.02, 0.1]), rng.choice([0.0, 0.005, 0.05]), rng.choice([0.0, 0.01]))[:ql]
            if rng.random() < 0.3:  # second copy elsewhere -> score2
                s2 = int(rng.integers(0, max(1, tl - ql)))
                t[s2:s2 + len(q)] = mutate(rng, q, 0.05, 0.0)[:len(t[s2:s2 + len(q)])] if len(q) <= tl - s2 else t[s2:s2 + len(q)]
        else:
            q = rng.integers(0, 4, size=ql).astype(np.uint8)
        if len(q) < 5:
            q = rng.integers(0, 4, size=20).astype(np.uint8)
        minsc = int(rng.choice([19, 19, 30, 0])) * a
        xtra = 0x80000 | (0x40000 if rng.random() < 0.8 else 0) | minsc
        if not sixteen:
            xtra |= 0x10000
        cases.append(dict(q=q, t=t, params=(a, b, od, ed, oi, ei), xtra=xtra))
    return cases


def global_cases(seed, n, max_q=260):
    """banded global alignments: bands 1..200, tlen != qlen, gappy pairs whose CIGARs run to dozens of operations"""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        ql = int(rng.integers(1, max_q))
        t0 = rng.integers(0, 4, size=ql + 300, dtype=np.uint8)
        style = k % 4
        if style == 0:
            q = mutate(rng, t0[:ql], 0.03, 0.003, 0.0)
        elif style == 1:
            q = mutate(rng, t0[:ql], 0.06, 0.08, 0.01)           # many short indels: long CIGARs
        elif style == 2:
            q = mutate(rng, t0[:ql], 0.01, 0.0, 0.02)
            cut = int(rng.integers(0, len(q) + 1))
            q = np.concatenate([q[:cut], rng.integers(0, 4, size=int(rng.integers(1, 40)), dtype=np.uint8), q[cut:]])   # one long insertion
        else:
            q = rng.integers(0, 5, size=ql, dtype=np.uint8)      # unrelated
        q = q[:max_q]
        if len(q) == 0:
            q = np.zeros(1, np.uint8)
        w = int(rng.choice([1, 2, 3, 5, 10, 30, 60, 100, 200]))
        # (the band always reaches the last cell: |tlen - qlen| <= w, as bwa_gen_cigar2 guarantees - below that ksw_global2 tracks back
        #  through cells it never wrote)
        d = int(rng.integers(-min(w, len(q) - 1), w + 1)) if style != 3 else int(rng.integers(-min(3, w, len(q) - 1), min(3, w) + 1))
        tl = max(1, len(q) + d)
        t = t0[:tl]
        params = [(1, 4, 6, 1, 6, 1), (1, 4, 6, 1, 6, 1), (2, 5, 7, 2, 8, 1), (1, 3, 4, 2, 5, 3)][int(rng.integers(0, 4))]
        out.append(dict(q=q.astype(np.uint8), t=t.astype(np.uint8), w=w, params=params))
    return out
