"""Read sharding across ranks: one process per GPU, full index replica per GPU, chunks of the fastq
pair dealt out to ranks, no data-path collective.

The reference hosts deal chunks out through an MPI RMA fetch-and-add counter (reference src/mainParallel.c:1112-1119),
i.e. first come first served; a chunk is self-contained (mem_pestat is computed per chunk, src/bwamem.c:1226-1229), so
WHICH rank aligns a chunk never changes its SAM.  Here the deal is the deterministic round-robin `chunk c -> rank
c % world`, which needs no shared counter and keeps the per-chunk output reproducible."""
import ctypes as C


def chunks_for_rank(n_chunks, rank, world):
    """ids of the chunks rank `rank` of `world` aligns"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_chunks, world))


def align_sharded(aligner, fq1, fq2, K, rank, world, trimmed=False):
    """Parse + plan like the hosts, align only this rank's chunks.  -> {chunk id: SAM bytes}, n_chunks."""
    b1, s1, n1 = aligner.parse(fq1)
    b2, s2, n2 = (aligner.parse(fq2) if fq2 is not None else (None, None, n1))
    if n1 != n2:
        raise ValueError("mate files hold different numbers of reads")
    ends = aligner.plan(n1, s1, s2, K, trimmed)
    begs = [0] + ends[:-1]
    per_read = 2 if s2 else 1
    out = {}
    for c in chunks_for_rank(len(ends), rank, world):
        n_proc = begs[c] * per_read if trimmed else 0      # n_processed convention of the trimmed branch
        out[c] = aligner.align_chunk(s1, s2, begs[c], ends[c], n_proc)[1]
    aligner.lib.b200_free(s1)
    if s2:
        aligner.lib.b200_free(s2)
    return out, len(ends)
