"""Index construction tooling: FASTA -> .pac/.ann/.amb/.bwt/.sa files bit-compatible with `bwa index` 0.7.17.

The reference ships no index builder (bwt_bwtgen / bwa_idx_build are declared in reference src/bwt.h:94-98 and
src/bwa.h:49 but defined nowhere) and expects the five files from an external `bwa index` (reference
docs/README.md:67).  The synthetic benchmark references need one, so this module writes the exact encodings the
reference *reads*:

  .pac  2 bits/base forward strand, MSB first (reference src/bntseq.c:224-225,309-322)
  .ann/.amb text tables (reference src/bntseq.c:66-96); N -> lrand48()&3 after srand48(11) (src/bntseq.c:261,290)
  .bwt  primary, L2[1..4], occ-interleaved BWT of fwd+revcomp text (reader: reference src/bwt.c:443-462)
  .sa   primary, L2[1..4], sa_intv, seq_len, samples (reader: reference src/bwt.c:421-441)

The suffix array is built by a radix-style sort of 31-mers (torch.sort; runs on the GPU when one is present)
followed by rounds that refine the groups of still-equal suffixes by their next 31-mer.  This is tooling, not
part of the alignment hot path.
"""
from __future__ import annotations

import ctypes
import os
import numpy as np
import torch

OCC_INTERVAL = 128
_NT4 = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _NT4[ord(_c)] = _i
    _NT4[ord(_c.lower())] = _i


def read_fasta(path):
    """-> list of (name, comment, sequence bytes)"""
    out, name, comment, chunks = [], None, "", []
    with open(path, "rb") as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    out.append((name, comment, b"".join(chunks)))
                head = line[1:].decode()
                parts = head.split(None, 1)
                name = parts[0] if parts else ""
                comment = parts[1] if len(parts) > 1 else ""
                chunks = []
            elif name is not None:
                chunks.append(line)
    if name is not None:
        out.append((name, comment, b"".join(chunks)))
    return out


def write_fasta(path, contigs, width=60):
    with open(path, "wb") as fh:
        for name, comment, seq in contigs:
            fh.write(b">" + name.encode() + ((b" " + comment.encode()) if comment else b"") + b"\n")
            arr = np.frombuffer(seq, dtype=np.uint8)
            n = len(arr)
            full = n // width * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = arr[:full].reshape(-1, width)
                body[:, width] = 10
                fh.write(body.tobytes())
            if n > full:
                fh.write(arr[full:].tobytes() + b"\n")


def _lrand48_stream(n):
    """n values of lrand48()&3 after srand48(11) (glibc), as the reference uses for N bases."""
    libc = ctypes.CDLL(None)
    libc.srand48(ctypes.c_long(11))
    libc.lrand48.restype = ctypes.c_long
    return np.fromiter((libc.lrand48() & 3 for _ in range(n)), dtype=np.uint8, count=n)


def encode_contigs(contigs):
    """-> (codes uint8[l_pac] with N replaced, anns, ambs) following add1() of reference src/bntseq.c:227-273"""
    anns, ambs, parts = [], [], []
    offset = 0
    n_total_n = sum(int((_NT4[np.frombuffer(s, dtype=np.uint8)] >= 4).sum()) for _, _, s in contigs)
    rnd = _lrand48_stream(n_total_n) if n_total_n else np.zeros(0, np.uint8)
    rpos = 0
    for name, comment, seq in contigs:
        raw = np.frombuffer(seq, dtype=np.uint8)
        c = _NT4[raw].copy()
        isn = c >= 4
        n_ambs = 0
        if isn.any():
            idx = np.flatnonzero(isn)
            # a hole continues while the same ambiguous character repeats at consecutive positions
            prev_same = np.zeros(len(idx), dtype=bool)
            prev_same[1:] = (idx[1:] == idx[:-1] + 1) & (raw[idx[1:]] == raw[idx[:-1]])
            # also the reference compares with `lasts`, the previous character whatever it was
            starts = np.flatnonzero(~prev_same)
            ends = np.append(starts[1:], len(idx))
            for s, e in zip(starts, ends):
                ambs.append((offset + int(idx[s]), int(e - s), chr(raw[idx[s]])))
            n_ambs = len(starts)
            c[idx] = rnd[rpos:rpos + len(idx)]
            rpos += len(idx)
        anns.append(dict(name=name, anno=comment if comment else "(null)", offset=offset, len=len(raw), n_ambs=n_ambs, gi=0))
        parts.append(c)
        offset += len(raw)
    codes = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    return codes, anns, ambs


def write_pac_ann_amb(prefix, codes, anns, ambs):
    l_pac = len(codes)
    pad = (-l_pac) % 4
    c = np.concatenate([codes, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    pac = (c[:, 0] << 6 | c[:, 1] << 4 | c[:, 2] << 2 | c[:, 3]).astype(np.uint8)
    with open(prefix + ".pac", "wb") as fh:
        fh.write(pac.tobytes())
        if l_pac % 4 == 0:
            fh.write(b"\0")
        fh.write(bytes([l_pac % 4]))
    with open(prefix + ".ann", "w") as fh:
        fh.write("%d %d %u\n" % (l_pac, len(anns), 11))
        for a in anns:
            fh.write("%d %s" % (a["gi"], a["name"]))
            fh.write(" %s\n" % a["anno"] if a["anno"] else "\n")
            fh.write("%d %d %d\n" % (a["offset"], a["len"], a["n_ambs"]))
    with open(prefix + ".amb", "w") as fh:
        fh.write("%d %d %u\n" % (l_pac, len(anns), len(ambs)))
        for off, ln, ch in ambs:
            fh.write("%d %d %c\n" % (off, ln, ch))


def _kmer31(text: torch.Tensor) -> torch.Tensor:
    """int64 key of the 31 bases starting at every position (zero padded past the end)"""
    n = text.numel()
    k = text.to(torch.int64)

    def shifted(x, d):
        out = torch.zeros_like(x)
        if d < n:
            out[: n - d] = x[d:]
        return out

    k = (k << 2) | shifted(k, 1)            # 2 bases
    k = (k << 4) | shifted(k, 2)            # 4
    k = (k << 8) | shifted(k, 4)            # 8
    k = (k << 16) | shifted(k, 8)           # 16 bases = 32 bits
    k = (k << 30) | (shifted(k, 16) >> 2)   # 31 bases = 62 bits
    return k


def suffix_array(codes: np.ndarray, device=None) -> np.ndarray:
    """Suffix array (int64[n]) of the text with an implicit sentinel smaller than every base."""
    n = len(codes)
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    text = torch.from_numpy(np.ascontiguousarray(codes)).to(device)
    K = 31
    key = _kmer31(text)
    ntr = min(K - 1, n)                      # suffixes shorter than K bases: positions n-ntr .. n-1
    order0 = torch.cat([torch.arange(n - 1, n - 1 - ntr, -1, device=device), torch.arange(0, n - ntr, device=device)])
    skey, perm = torch.sort(key[order0], stable=True)
    sa = order0[perm]
    del perm, order0
    trunc = sa > n - K
    same = torch.zeros(n, dtype=torch.bool, device=device)
    same[1:] = (skey[1:] == skey[:-1]) & ~trunc[1:] & ~trunc[:-1]
    del skey, trunc
    # group id = index of the first element of the run of equal keys
    idx = torch.arange(n, device=device)
    start = torch.where(same, torch.zeros_like(idx), idx)
    grp = torch.cummax(start, 0).values
    del start
    nxt = torch.zeros(n, dtype=torch.bool, device=device)
    nxt[:-1] = same[1:]
    active = torch.nonzero(same | nxt).flatten()
    del same, nxt, idx
    depth = K
    while active.numel() > 0:
        pos = sa[active]
        g = grp[active]
        p2 = pos + depth
        inb = p2 < n
        k2 = torch.where(inb, key[torch.clamp(p2, max=n - 1)], torch.zeros_like(p2))
        vlen = torch.clamp(n - p2, min=0, max=K)
        # order inside a group: by next 31-mer; among equal padded keys the shorter (truncated) suffix first
        o = torch.sort(vlen, stable=True).indices
        o = o[torch.sort(k2[o], stable=True).indices]
        o = o[torch.sort(g[o], stable=True).indices]
        pos, g, k2, vlen = pos[o], g[o], k2[o], vlen[o]
        sa[active] = pos
        m = active.numel()
        same = torch.zeros(m, dtype=torch.bool, device=device)
        same[1:] = (g[1:] == g[:-1]) & (k2[1:] == k2[:-1]) & (vlen[1:] == K) & (vlen[:-1] == K)
        ar = torch.arange(m, device=device)
        start = torch.where(same, torch.zeros_like(ar), ar)
        first = torch.cummax(start, 0).values
        grp[active] = active[first]
        nxt = torch.zeros(m, dtype=torch.bool, device=device)
        nxt[:-1] = same[1:]
        active = active[same | nxt]
        depth += K
    return sa.cpu().numpy()


def build_bwt_sa(codes_fwd: np.ndarray, sa_intv=32, device=None):
    """-> dict(primary, L2[5], seq_len, bwt uint32[bwt_size] (occ-interleaved), sa uint64[n_sa])"""
    l_pac = len(codes_fwd)
    text = np.concatenate([codes_fwd, (3 - codes_fwd[::-1]).astype(np.uint8)])
    n = len(text)
    sa = suffix_array(text, device)
    row0 = int(np.flatnonzero(sa == 0)[0])
    primary = row0 + 1
    # BWT column without the sentinel row
    prev = sa - 1
    bw = np.empty(n, dtype=np.uint8)
    bw[0] = text[n - 1]
    keep = np.ones(n, dtype=bool)
    keep[row0] = False
    bw[1:] = text[prev[keep]]
    cnt = np.bincount(text, minlength=4).astype(np.uint64)
    L2 = np.zeros(5, dtype=np.uint64)
    L2[1:] = np.cumsum(cnt)
    # 2-bit pack, first symbol in the top bits of each word
    n_words = (n + 15) >> 4
    padded = np.zeros(n_words * 16, dtype=np.uint32)
    padded[:n] = bw
    shifts = (30 - 2 * np.arange(16)).astype(np.uint32)
    words = np.bitwise_or.reduce(padded.reshape(-1, 16) << shifts, axis=1).astype(np.uint32)
    # occ interleave: per 128 symbols 4 x uint64 counts-before followed by the (up to) 8 symbol words
    n_blk = (n + OCC_INTERVAL - 1) // OCC_INTERVAL
    blk_pad = np.full(n_blk * OCC_INTERVAL, 4, dtype=np.uint8)
    blk_pad[:n] = bw
    blk = blk_pad.reshape(n_blk, OCC_INTERVAL)
    per = np.stack([(blk == c).sum(axis=1) for c in range(4)], axis=1).astype(np.uint64)
    before = np.zeros((n_blk + 1, 4), dtype=np.uint64)
    before[1:] = np.cumsum(per, axis=0)
    bwt_size = n_words + (n_blk + 1) * 8
    wpad = np.zeros(n_blk * 8, dtype=np.uint32)
    wpad[:n_words] = words
    rec = np.zeros((n_blk, 16), dtype=np.uint32)
    rec[:, :8] = before[:n_blk].view(np.uint32).reshape(n_blk, 8)
    rec[:, 8:] = wpad.reshape(n_blk, 8)
    flat = rec.reshape(-1)
    tail_words = n_words - (n_blk - 1) * 8 if n_blk else 0
    body = flat[: (n_blk - 1) * 16 + 8 + tail_words] if n_blk else flat[:0]
    bwt = np.concatenate([body, before[n_blk].view(np.uint32)])
    assert len(bwt) == bwt_size, (len(bwt), bwt_size)
    # sampled suffix array in the (n+1)-row coordinate system; row 0 is the sentinel suffix
    n_sa = (n + sa_intv) // sa_intv
    rows = np.arange(1, n_sa, dtype=np.int64) * sa_intv
    sa_s = np.empty(n_sa, dtype=np.uint64)
    sa_s[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    sa_s[1:] = sa[rows - 1].astype(np.uint64)
    return dict(primary=primary, L2=L2, seq_len=n, bwt=bwt, sa=sa_s, sa_intv=sa_intv, l_pac=l_pac)


def write_bwt_sa(prefix, idx):
    hdr = np.array([idx["primary"], *idx["L2"][1:5]], dtype=np.uint64)
    with open(prefix + ".bwt", "wb") as fh:
        fh.write(hdr.tobytes())
        fh.write(idx["bwt"].tobytes())
    with open(prefix + ".sa", "wb") as fh:
        fh.write(hdr.tobytes())
        fh.write(np.array([idx["sa_intv"], idx["seq_len"]], dtype=np.uint64).tobytes())
        fh.write(idx["sa"][1:].tobytes())


def build_index(fasta_path, prefix=None, sa_intv=32, device=None):
    """`bwa index` equivalent: writes <prefix>.{pac,ann,amb,bwt,sa}; returns the prefix."""
    prefix = prefix or fasta_path
    contigs = read_fasta(fasta_path)
    codes, anns, ambs = encode_contigs(contigs)
    write_pac_ann_amb(prefix, codes, anns, ambs)
    write_bwt_sa(prefix, build_bwt_sa(codes, sa_intv, device))
    return prefix


def build_index_from_codes(prefix, names, lengths, codes, sa_intv=32, device=None):
    """Index an N-free reference given directly as base codes (skips FASTA parsing); also writes <prefix> FASTA-less."""
    anns, off = [], 0
    for nm, ln in zip(names, lengths):
        anns.append(dict(name=nm, anno="(null)", offset=off, len=int(ln), n_ambs=0, gi=0))
        off += int(ln)
    assert off == len(codes)
    write_pac_ann_amb(prefix, codes, anns, [])
    write_bwt_sa(prefix, build_bwt_sa(codes, sa_intv, device))
    return prefix


if __name__ == "__main__":
    import sys
    build_index(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
