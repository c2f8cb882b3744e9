"""World-size-2 run of the read-sharding logic over `gloo` on CPU: two ranks deal the chunks of one fastq pair between
them (mpibwa_b200.shard), every chunk's SAM must equal the single-process result, and the union must cover every chunk
exactly once.  The device stages are stood in for by tests/hostemu (test scaffold, never part of the product)."""
import hashlib
import os
import socket
import pytest
from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, libpath, idx, f1, f2, K, q):
    import torch.distributed as dist
    import mpibwa_b200 as M
    from mpibwa_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = M.load(libpath)
    al = M.Aligner(idx, device=0, n_threads=2, lib=lib, verbose=1)
    sams, n_chunks = shard.align_sharded(al, open(f1, "rb").read(), open(f2, "rb").read(), K, rank, world)
    mine = {c: hashlib.md5(s).hexdigest() for c, s in sams.items()}
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, n_chunks, mine))
    dist.barrier()
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


def _head(path, n_reads, out):
    with open(path, "rb") as fi:
        lines = fi.read().split(b"\n")[:4 * n_reads]
    with open(out, "wb") as fo:
        fo.write(b"\n".join(lines) + b"\n")
    return out


def test_chunks_for_rank_partition():
    from mpibwa_b200 import shard
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 16):
            seen = sorted(c for r in range(world) for c in shard.chunks_for_rank(n, r, world))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard.chunks_for_rank(4, 2, 2)


def test_two_ranks_gloo_match_single_process(hostemu_built, examples, tmp_path):
    import torch.multiprocessing as mp
    import mpibwa_b200 as M
    from mpibwa_b200 import shard
    libpath = os.path.join(hostemu_built, "libmpibwa_b200_hostemu.so")
    f1 = _head(examples["R1_10K"], 1000, str(tmp_path / "a.fq"))
    f2 = _head(examples["R2_10K"], 1000, str(tmp_path / "b.fq"))
    K = 60000                                                # -> 5 chunks
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, libpath, examples["idx"], f1, f2, K, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    lib = M.load(libpath)
    al = M.Aligner(examples["idx"], device=0, n_threads=2, lib=lib, verbose=1)
    single, n_chunks = shard.align_sharded(al, open(f1, "rb").read(), open(f2, "rb").read(), K, 0, 1)
    assert n_chunks >= 4
    merged = {}
    for rank, n, mine in gathered:
        assert n == n_chunks
        assert sorted(mine) == shard.chunks_for_rank(n_chunks, rank, 2)
        merged.update(mine)
    assert sorted(merged) == list(range(n_chunks))
    for c in range(n_chunks):
        assert merged[c] == hashlib.md5(single[c]).hexdigest(), c
