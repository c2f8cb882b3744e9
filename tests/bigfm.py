"""TEST INFRASTRUCTURE: a synthetic FM index with more than 2^32 BWT rows, written in the reference's file formats
(reference src/bwt.c:421-462, src/bntseq.c:100-166,224-225), so that the 33-/40-bit paths of the device's occ sectors
(mpibwa_b200/csrc/fm_kernels.h: ld_occ, occ4_sector) and of the packed interval lists of the seeding kernels are exercised on the
REAL kernels (tests/test_gpu_parity.py::test_fm_index_beyond_2_32).

The "BWT" is not the transform of any text: bwt_occ4 / bwt_extend / the SMEM sweeps are pure counting arithmetic over the symbol
array and its interleaved counts, which is all that is compared (against oracle/oracle_fmindex.c on the same files).  Rows are
one symbol `dom` except for windows of random symbols, one of them around the row where the count of `dom` crosses 2^32."""
import numpy as np


def _block_counts(words):
    """words: (n_blk, 8) uint32, 16 two-bit symbols per word, first symbol in the top bits -> (n_blk, 4) counts of A,C,G,T"""
    out = np.zeros((words.shape[0], 4), np.int64)
    w = words.astype(np.uint64)
    for s in range(16):
        sym = (w >> np.uint64(2 * s)) & np.uint64(3)
        for c in range(4):
            out[:, c] += (sym == c).sum(axis=1)
    return out


def write_index(prefix, dom=1, seq_len=(1 << 32) + (1 << 27), windows=3, win_blocks=4096, seed=5):
    assert seq_len % 256 == 0
    rng = np.random.default_rng(seed)
    n_blk = seq_len // 128
    fill = np.uint32(sum(dom << (2 * s) for s in range(16)))
    sym = np.full((n_blk, 8), fill, np.uint32)
    per = np.zeros((n_blk, 4), np.int64)
    per[:, dom] = 128
    cross = (1 << 32) // 128                              # block in which the count of `dom` is about to cross 2^32
    starts = [cross - win_blocks // 2] + [int(x) for x in rng.integers(1000, n_blk - win_blocks - 1000, size=windows - 1)]
    for b0 in starts:
        w = rng.integers(0, 1 << 32, size=(win_blocks, 8), dtype=np.uint64).astype(np.uint32)
        sym[b0:b0 + win_blocks] = w
        per[b0:b0 + win_blocks] = _block_counts(w)
    before = np.zeros((n_blk + 1, 4), np.uint64)
    np.cumsum(per, axis=0, out=before[1:].view(np.int64))
    total = before[n_blk].astype(np.int64)
    bwt = np.zeros(n_blk * 16 + 8, np.uint32)
    body = bwt[:n_blk * 16].reshape(n_blk, 16)
    body[:, :8] = before[:n_blk].view(np.uint32).reshape(n_blk, 8)
    body[:, 8:] = sym
    bwt[n_blk * 16:] = before[n_blk].view(np.uint32)
    primary = seq_len // 3 + 12345
    L2 = np.concatenate([[0], np.cumsum(total)]).astype(np.uint64)
    with open(prefix + ".bwt", "wb") as fh:
        np.array([primary], np.uint64).tofile(fh)
        L2[1:].tofile(fh)
        bwt.tofile(fh)
    sa_intv = 1 << 30
    n_sa = (seq_len + sa_intv) // sa_intv
    with open(prefix + ".sa", "wb") as fh:
        np.array([primary, 0, 0, 0, 0, sa_intv, seq_len], np.uint64).tofile(fh)
        np.arange(1, n_sa, dtype=np.uint64).tofile(fh)
    l_pac = seq_len // 2
    with open(prefix + ".pac", "wb") as fh:
        fh.write(bytes(l_pac // 4 + 2))
    half = l_pac // 2
    with open(prefix + ".ann", "w") as fh:
        fh.write("%d 2 11\n0 big1 (null)\n0 %d 0\n0 big2 (null)\n%d %d 0\n" % (l_pac, half, half, l_pac - half))
    with open(prefix + ".amb", "w") as fh:
        fh.write("%d 2 0\n" % l_pac)
    return dict(seq_len=seq_len, primary=primary, L2=[int(x) for x in L2], windows=[(b0 * 128, (b0 + win_blocks) * 128) for b0 in starts], dom=dom)
