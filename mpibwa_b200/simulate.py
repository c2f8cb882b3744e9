"""Synthetic references and wgsim-style paired reads for the benchmark configurations of BASELINE.json.

No dataset can be downloaded here and the image has no wgsim, so the workloads are generated: an i.i.d. uniform
ACGT reference with planted diverged repeats (so that the max_occ / repeat paths of the aligner fire) and read
pairs drawn from normally distributed inserts with substitutions, indels and a sprinkle of N bases.
Everything is seeded and vectorised with numpy; this is tooling, not part of the alignment hot path.
"""
from __future__ import annotations

import numpy as np

_BASES = np.frombuffer(b"ACGTN", dtype=np.uint8)


def make_reference(total_bp, n_contigs=4, seed=1, repeat_frac=0.05, repeat_div=0.01, rep_len=(200, 5000)):
    """-> (names, lengths, codes uint8[total_bp] in 0..3).  Contig i is named chr<i+1>."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, size=total_bp, dtype=np.uint8)
    planted = 0
    target = int(total_bp * repeat_frac)
    while planted < target and total_bp > 4 * rep_len[0]:
        ln = int(rng.integers(rep_len[0], min(rep_len[1], total_bp // 4) + 1))
        src = int(rng.integers(0, total_bp - ln))
        dst = int(rng.integers(0, total_bp - ln))
        seg = codes[src:src + ln].copy()
        if rng.random() < 0.5:
            seg = (3 - seg[::-1]).astype(np.uint8)
        mut = rng.random(ln) < repeat_div
        seg[mut] = (seg[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
        codes[dst:dst + ln] = seg
        planted += ln
    w = np.linspace(1.6, 0.6, n_contigs)
    lengths = np.floor(w / w.sum() * total_bp).astype(np.int64)
    lengths[-1] = total_bp - lengths[:-1].sum()
    names = ["chr%d" % (i + 1) for i in range(n_contigs)]
    return names, lengths, codes


def codes_to_fasta_contigs(names, lengths, codes, n_runs=0, seed=3):
    """-> contigs for index_build.write_fasta; optionally overwrite n_runs stretches with N (creates .amb holes)."""
    rng = np.random.default_rng(seed)
    out, off = [], 0
    for nm, ln in zip(names, lengths):
        seq = _BASES[codes[off:off + ln]].copy()
        for _ in range(n_runs):
            l = int(rng.integers(1, 200))
            if ln > l + 2:
                s = int(rng.integers(0, ln - l))
                seq[s:s + l] = ord("N")
        out.append((nm, "", seq.tobytes()))
        off += ln
    return out


def _apply_errors(rng, src, L, sub, indel, max_indel, n_rate):
    """src: uint8[n, W] source windows (codes 0..3), W >= L + slack.  Returns codes uint8[n, L] (4 = N)."""
    n, W = src.shape
    ins = np.zeros((n, L), dtype=bool)
    dele = np.zeros((n, L), dtype=np.int32)
    if indel > 0:
        ev = rng.random((n, L)) < indel
        ev[:, 0] = False
        is_ins = ev & (rng.random((n, L)) < 0.5)
        is_del = ev & ~is_ins
        glen = np.minimum(rng.geometric(0.7 if max_indel > 1 else 1.0, size=(n, L)), max_indel).astype(np.int32)
        dele = np.where(is_del, glen, 0).astype(np.int32)
        run = np.where(is_ins, glen, 0).astype(np.int32)
        cur = run
        for _ in range(max_indel):
            ins |= cur > 0
            nxt = np.zeros_like(cur)
            nxt[:, 1:] = np.maximum(cur[:, :-1] - 1, 0)
            cur = nxt
            if not cur.any():
                break
        dele[ins] = 0
    adv = (~ins).astype(np.int32)
    off = np.cumsum(adv, axis=1) - adv + np.cumsum(dele, axis=1)
    off = np.minimum(off, W - 1)
    out = np.take_along_axis(src, off, axis=1)
    if ins.any():
        out[ins] = rng.integers(0, 4, size=int(ins.sum()), dtype=np.uint8)
    if sub > 0:
        m = rng.random((n, L)) < sub
        out[m] = (out[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
    if n_rate > 0:
        out[rng.random((n, L)) < n_rate] = 4
    return out


def simulate_pairs(codes, lengths, n_pairs, read_len=150, ins_mean=400, ins_sd=50, sub=0.01, indel=0.001,
                   max_indel=1, n_rate=0.001, seed=2, trim_to=None, unmappable_frac=0.0, prefix="sim", batch=200000):
    """-> (fastq bytes R1, fastq bytes R2).  trim_to=(lo, hi): each read independently cut to U(lo, hi) bases."""
    rng = np.random.default_rng(seed)
    lengths = np.asarray(lengths, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    slack = 8 + (max_indel * 4 if indel > 0 else 0)
    W = read_len + slack
    min_ins = read_len + 10 + slack
    r1_parts, r2_parts = [], []
    digits = max(7, len(str(n_pairs)))
    for b0 in range(0, n_pairs, batch):
        n = min(batch, n_pairs - b0)
        isz = np.maximum(np.rint(rng.normal(ins_mean, ins_sd, size=n)).astype(np.int64), min_ins)
        ctg = rng.choice(len(lengths), size=n, p=lengths / lengths.sum())
        isz = np.minimum(isz, lengths[ctg] - 1)
        pos = starts[ctg] + (rng.random(n) * (lengths[ctg] - isz)).astype(np.int64)
        ar = np.arange(W, dtype=np.int64)
        left = codes[np.minimum(pos[:, None] + ar[None, :], len(codes) - 1)]
        right = 3 - codes[np.maximum((pos + isz - 1)[:, None] - ar[None, :], 0)]
        flip = rng.random(n) < 0.5
        a = np.where(flip[:, None], right, left).astype(np.uint8)
        b = np.where(flip[:, None], left, right).astype(np.uint8)
        r1 = _apply_errors(rng, a, read_len, sub, indel, max_indel, n_rate)
        r2 = _apply_errors(rng, b, read_len, sub, indel, max_indel, n_rate)
        if unmappable_frac > 0:
            um = rng.random(n) < unmappable_frac
            k = int(um.sum())
            if k:
                r2[um] = rng.integers(0, 4, size=(k, read_len), dtype=np.uint8)
        for which, (r, parts) in enumerate(((r1, r1_parts), (r2, r2_parts))):
            seq = _BASES[r]
            qual = np.where(r == 4, ord("#"), ord("F")).astype(np.uint8)
            names = np.char.add(np.char.add("@" + prefix, np.char.zfill(np.arange(b0, b0 + n).astype(str), digits)), "/%d" % (which + 1))
            nm = np.frombuffer("".join(names.tolist()).encode(), dtype=np.uint8).reshape(n, -1)
            if trim_to is None:
                rec = np.concatenate([nm, np.full((n, 1), 10, np.uint8), seq, np.frombuffer(b"\n+\n", np.uint8)[None, :].repeat(n, 0),
                                      qual, np.full((n, 1), 10, np.uint8)], axis=1)
                parts.append(rec.tobytes())
            else:
                ls = rng.integers(trim_to[0], trim_to[1] + 1, size=n)
                nmb, sb, qb = nm.tobytes(), seq.tobytes(), qual.tobytes()
                wn = nm.shape[1]
                parts.append(b"".join(nmb[i * wn:(i + 1) * wn] + b"\n" + sb[i * read_len:i * read_len + l] + b"\n+\n" +
                                      qb[i * read_len:i * read_len + l] + b"\n" for i, l in enumerate(ls.tolist())))
    return b"".join(r1_parts), b"".join(r2_parts)
