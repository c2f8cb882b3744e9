"""Index construction tooling: FASTA -> .pac/.ann/.amb/.bwt/.sa files bit-compatible with `bwa index` 0.7.17.

The reference ships no index builder (bwt_bwtgen / bwa_idx_build are declared in reference src/bwt.h:94-98 and
src/bwa.h:49 but defined nowhere) and expects the five files from an external `bwa index` (reference
docs/README.md:67).  The synthetic benchmark references need one, so this module writes the exact encodings the
reference *reads*:

  .pac  2 bits/base forward strand, MSB first (reference src/bntseq.c:224-225,309-322)
  .ann/.amb text tables (reference src/bntseq.c:66-96); N -> lrand48()&3 after srand48(11) (src/bntseq.c:261,290)
  .bwt  primary, L2[1..4], occ-interleaved BWT of fwd+revcomp text (reader: reference src/bwt.c:443-462)
  .sa   primary, L2[1..4], sa_intv, seq_len, samples (reader: reference src/bwt.c:421-441)

The suffix array is built bucket range by bucket range (12-base prefixes; a radix sort of the 31-base keys of one range -
torch.sort, on the GPU when one is present - followed by rounds that refine the groups of still-equal suffixes by their next
31 bases), so that memory stays a few arrays of one range whatever the reference size: a human-sized reference (2 x 3.1 G
suffixes) is built on one B200.  BWT symbols and sampled rows are taken from each slice as it is produced and the
occ-interleaved layout is packed on the device.  This is tooling, not part of the alignment hot path.
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
    with open(prefix + ".pac", "wb") as fh:
        step = 1 << 28                                       # (slices: a human-sized reference is 3.1 G codes)
        for b0 in range(0, l_pac, step):
            seg = codes[b0:min(l_pac, b0 + step)]
            pad = (-len(seg)) % 4
            c = (np.concatenate([seg, np.zeros(pad, np.uint8)]) if pad else seg).reshape(-1, 4)
            fh.write((c[:, 0] << 6 | c[:, 1] << 4 | c[:, 2] << 2 | c[:, 3]).astype(np.uint8).tobytes())
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


K = 31                      # bases per sort key (62 bits of an int64)
_PREFIX = 12                # bases of the bucket prefix (4^12 buckets)


class _Text:
    """2-bit text packed 32 bases per int64 word (first base in the top bits) with the key of any position two gathers away"""

    def __init__(self, codes: np.ndarray, device):
        self.n = n = len(codes)
        self.device = device
        nw = (n + 31) // 32 + 2
        self.words = torch.zeros(nw, dtype=torch.int64, device=device)
        self.text = torch.from_numpy(np.ascontiguousarray(codes)).to(device)
        sh = (62 - 2 * torch.arange(32, device=device)).to(torch.int64)
        step = 1 << 26                                       # bases per slice (32 x 8 bytes of scratch each)
        for b0 in range(0, n, step):
            b1 = min(n, b0 + step)
            m = (b1 - b0 + 31) // 32
            blk = torch.zeros(m * 32, dtype=torch.int64, device=device)
            blk[: b1 - b0] = self.text[b0:b1]
            self.words[b0 // 32: b0 // 32 + m] = (blk.view(m, 32) << sh).sum(dim=1)     # (disjoint bit fields: the sum is an OR)
            del blk

    def key(self, pos: torch.Tensor) -> torch.Tensor:
        """the K = 31 bases starting at pos as a non-negative int64 (zero padded past the end)"""
        pos = torch.clamp(pos, max=self.n)                   # (positions past the end read the zero padding)
        q = pos >> 5
        r = (pos & 31) << 1
        hi = self.words[q]
        lo = self.words[q + 1]
        a = (torch.bitwise_left_shift(hi, r) >> 2) & 0x3FFFFFFFFFFFFFFF
        sh = torch.clamp(66 - r, max=63)
        b = torch.where(r >= 4, torch.bitwise_right_shift(lo, sh) & (torch.bitwise_left_shift(torch.ones_like(r), torch.clamp(r - 2, min=0)) - 1),
                        torch.zeros_like(lo))
        return a | b


def _sa_batches(codes: np.ndarray, device=None, batch=None):
    """Suffix array of the text (implicit sentinel smaller than every base), in order, as a stream of slices.

    Bucketed: a histogram of the 12-base prefixes cuts the suffixes into ranges of at most `batch` suffixes; each range is
    sorted by its 31-base key (one radix sort: torch.sort), then the groups of still-equal suffixes are refined by their next
    31 bases until no group is left.  Peak memory is a few arrays of `batch` int64 beside the text, whatever the text size, so a
    human-sized reference (2 x 3.1 G suffixes) fits one B200.  Yields (first_row, sa_slice int64 tensor on `device`)."""
    n = len(codes)
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if batch is None:
        batch = 1 << 29 if str(device).startswith("cuda") else 1 << 24
    T = _Text(codes, device)
    if n == 0:
        return
    shift = 2 * (K - _PREFIX)
    nb = 1 << (2 * _PREFIX)
    step = 1 << 27
    # ---- histogram of the bucket prefixes
    hist = torch.zeros(nb, dtype=torch.int64, device=device)
    for b0 in range(0, n, step):
        pos = torch.arange(b0, min(n, b0 + step), device=device)
        hist += torch.bincount(T.key(pos) >> shift, minlength=nb)
        del pos
    cum = torch.cumsum(hist, 0).cpu().numpy()
    # ---- ranges of whole buckets of at most `batch` suffixes (a single bucket may exceed it: poly-A runs ...)
    bounds = [0]
    while bounds[-1] < nb:
        lo = bounds[-1]
        base = int(cum[lo - 1]) if lo else 0
        hi = int(np.searchsorted(cum, base + batch, side="right"))
        bounds.append(min(nb, max(hi, lo + 1)))
    ntr = min(K - 1, n)                                      # suffixes shorter than K bases: positions n-ntr .. n-1
    row = 0
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        m_expect = int(cum[hi - 1]) - (int(cum[lo - 1]) if lo else 0)
        if m_expect == 0:
            continue
        parts = []
        for b0 in range(0, n, step):
            pos = torch.arange(b0, min(n, b0 + step), device=device)
            kb = T.key(pos) >> shift
            parts.append(pos[(kb >= lo) & (kb < hi)])
            del pos, kb
        pos = torch.cat(parts)
        del parts
        # truncated suffixes first, shortest first: the stable sort keeps them in front of the full suffixes with the same padded key
        tr = pos >= n - ntr
        if bool(tr.any()):
            pos = torch.cat([torch.flip(pos[tr], [0]), pos[~tr]])
        del tr
        key = T.key(pos)
        skey, perm = torch.sort(key, stable=True)
        sa = pos[perm]
        del perm, key, pos
        m = sa.numel()
        trunc = sa > n - K
        same = torch.zeros(m, dtype=torch.bool, device=device)
        same[1:] = (skey[1:] == skey[:-1]) & ~trunc[1:] & ~trunc[:-1]
        del skey, trunc
        idx = torch.arange(m, device=device)
        grp = torch.cummax(torch.where(same, torch.zeros_like(idx), idx), 0).values      # group id = index of its first element
        nxt = torch.zeros(m, dtype=torch.bool, device=device)
        nxt[:-1] = same[1:]
        active = torch.nonzero(same | nxt).flatten()
        del same, nxt, idx
        depth = K
        while active.numel() > 0:
            p = sa[active]
            g = grp[active]
            p2 = p + depth
            k2 = T.key(p2)
            vlen = torch.clamp(n - p2, min=0, max=K)
            # order inside a group: by next 31 bases; among equal padded keys the shorter (truncated) suffix first
            o = torch.sort(vlen, stable=True).indices
            o = o[torch.sort(k2[o], stable=True).indices]
            o = o[torch.sort(g[o], stable=True).indices]
            p, g, k2, vlen = p[o], g[o], k2[o], vlen[o]
            sa[active] = p
            ma = active.numel()
            same = torch.zeros(ma, dtype=torch.bool, device=device)
            same[1:] = (g[1:] == g[:-1]) & (k2[1:] == k2[:-1]) & (vlen[1:] == K) & (vlen[:-1] == K)
            ar = torch.arange(ma, device=device)
            first = torch.cummax(torch.where(same, torch.zeros_like(ar), ar), 0).values
            grp[active] = active[first]
            nx = torch.zeros(ma, dtype=torch.bool, device=device)
            nx[:-1] = same[1:]
            active = active[same | nx]
            depth += K
        del grp
        assert m == m_expect
        yield row, sa
        row += m
        del sa


def suffix_array(codes: np.ndarray, device=None, batch=None) -> np.ndarray:
    """Suffix array (int64[n]) of the text with an implicit sentinel smaller than every base."""
    parts = [sa.cpu().numpy() for _, sa in _sa_batches(codes, device, batch)]
    return np.concatenate(parts) if parts else np.zeros(0, np.int64)


def build_bwt_sa(codes_fwd: np.ndarray, sa_intv=32, device=None, batch=None):
    """-> dict(primary, L2[5], seq_len, bwt uint32[bwt_size] (occ-interleaved), sa uint64[n_sa]).  The suffix array is consumed
    slice by slice (BWT symbols and the sampled rows are all that is kept), and the occ-interleaved layout is packed on the device."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    l_pac = len(codes_fwd)
    text_np = np.concatenate([codes_fwd, (3 - codes_fwd[::-1]).astype(np.uint8)])
    n = len(text_np)
    text = torch.from_numpy(text_np).to(device)
    # rows of the (n+1)-row matrix: row 0 is the sentinel suffix, row i+1 is suffix sa[i]; bw_full[row] = the base before the suffix
    bw_full = torch.empty(n + 1, dtype=torch.uint8, device=device)
    bw_full[0] = text[n - 1]
    n_sa = (n + sa_intv) // sa_intv
    sa_s = torch.zeros(n_sa, dtype=torch.int64, device=device)
    primary = -1
    for row, sa in _sa_batches(text_np, device, batch):
        m = sa.numel()
        z = torch.nonzero(sa == 0).flatten()
        if z.numel():
            primary = row + int(z[0]) + 1
        bw_full[row + 1: row + 1 + m] = text[torch.clamp(sa - 1, min=0)]
        # sampled rows: matrix row j * sa_intv holds suffix sa[j * sa_intv - 1]
        j0 = (row + 1 + sa_intv - 1) // sa_intv
        j1 = (row + m) // sa_intv
        if j1 >= j0 and j1 >= 1:
            j = torch.arange(max(j0, 1), j1 + 1, device=device)
            sa_s[j] = sa[j * sa_intv - 1 - row]
        del sa
    assert primary > 0
    # BWT column without the sentinel row
    bw = torch.cat([bw_full[:primary], bw_full[primary + 1:]])
    del bw_full
    cnt = torch.bincount(text.to(torch.int64) if n < (1 << 24) else text[: 1 << 24].to(torch.int64), minlength=4)
    if n >= (1 << 24):
        cnt = torch.zeros(4, dtype=torch.int64, device=device)
        for b0 in range(0, n, 1 << 28):
            cnt += torch.bincount(text[b0:b0 + (1 << 28)].to(torch.int64), minlength=4)
    L2 = np.zeros(5, dtype=np.uint64)
    L2[1:] = np.cumsum(cnt.cpu().numpy().astype(np.uint64))
    del text
    # occ interleave: per 128 symbols 4 x uint64 counts-before followed by the (up to) 8 words of 16 two-bit symbols (first
    # symbol in the top bits of each word); one more count record closes the array
    n_words = (n + 15) >> 4
    n_blk = (n + OCC_INTERVAL - 1) // OCC_INTERVAL
    bwt_size = n_words + (n_blk + 1) * 8
    rec = np.zeros((n_blk + 1, 16), dtype=np.uint32)
    sh16 = (30 - 2 * torch.arange(16, device=device)).to(torch.int64)
    run = torch.zeros(4, dtype=torch.int64, device=device)
    step = 1 << 20                                           # blocks per slice
    for k0 in range(0, n_blk, step):
        k1 = min(n_blk, k0 + step)
        sl = torch.full(((k1 - k0) * OCC_INTERVAL,), 4, dtype=torch.uint8, device=device)
        seg = bw[k0 * OCC_INTERVAL: min(n, k1 * OCC_INTERVAL)]
        sl[: seg.numel()] = seg
        blk = sl.view(k1 - k0, OCC_INTERVAL)
        per = torch.stack([(blk == c).sum(dim=1) for c in range(4)], dim=1).to(torch.int64)
        before = torch.cumsum(per, 0) - per + run
        run = run + per.sum(dim=0)
        sym = torch.where(sl < 4, sl, torch.zeros_like(sl)).to(torch.int64).view(-1, 16)
        words = (sym << sh16).sum(dim=1).view(k1 - k0, 8)
        out = torch.empty((k1 - k0, 16), dtype=torch.int64, device=device)
        out[:, 0:8:2] = before & 0xFFFFFFFF
        out[:, 1:8:2] = before >> 32
        out[:, 8:] = words
        rec[k0:k1] = out.cpu().numpy().astype(np.uint32)
        del sl, blk, per, before, sym, words, out
    rec[n_blk, 0:8:2] = (run & 0xFFFFFFFF).cpu().numpy().astype(np.uint32)
    rec[n_blk, 1:8:2] = (run >> 32).cpu().numpy().astype(np.uint32)
    flat = rec.reshape(-1)
    tail_words = n_words - (n_blk - 1) * 8 if n_blk else 0
    bwt = np.concatenate([flat[: (n_blk - 1) * 16 + 8 + tail_words] if n_blk else flat[:0], flat[n_blk * 16: n_blk * 16 + 8]])
    assert len(bwt) == bwt_size, (len(bwt), bwt_size)
    sa_np = sa_s.cpu().numpy().astype(np.uint64)
    sa_np[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    if str(device).startswith("cuda"):
        del bw, sa_s
        torch.cuda.empty_cache()
    return dict(primary=primary, L2=L2, seq_len=n, bwt=bwt, sa=sa_np, sa_intv=sa_intv, l_pac=l_pac)


def write_bwt_sa(prefix, idx):
    hdr = np.array([idx["primary"], *idx["L2"][1:5]], dtype=np.uint64)
    with open(prefix + ".bwt", "wb") as fh:
        fh.write(hdr.tobytes())
        idx["bwt"].tofile(fh)
    with open(prefix + ".sa", "wb") as fh:
        fh.write(hdr.tobytes())
        fh.write(np.array([idx["sa_intv"], idx["seq_len"]], dtype=np.uint64).tobytes())
        idx["sa"][1:].tofile(fh)


def build_index(fasta_path, prefix=None, sa_intv=32, device=None, batch=None):
    """`bwa index` equivalent: writes <prefix>.{pac,ann,amb,bwt,sa}; returns the prefix."""
    prefix = prefix or fasta_path
    contigs = read_fasta(fasta_path)
    codes, anns, ambs = encode_contigs(contigs)
    write_pac_ann_amb(prefix, codes, anns, ambs)
    write_bwt_sa(prefix, build_bwt_sa(codes, sa_intv, device, batch))
    return prefix


def build_index_from_codes(prefix, names, lengths, codes, sa_intv=32, device=None, batch=None):
    """Index an N-free reference given directly as base codes (skips FASTA parsing); also writes <prefix> FASTA-less."""
    anns, off = [], 0
    for nm, ln in zip(names, lengths):
        anns.append(dict(name=nm, anno="(null)", offset=off, len=int(ln), n_ambs=0, gi=0))
        off += int(ln)
    assert off == len(codes)
    write_pac_ann_amb(prefix, codes, anns, [])
    write_bwt_sa(prefix, build_bwt_sa(codes, sa_intv, device, batch))
    return prefix


if __name__ == "__main__":
    import sys
    build_index(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
