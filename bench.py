#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 alignment core (contract: see the task statement / DESIGN.md).

Workload (default: BASELINE.json configs[2]'s reference; --ref-bp 100000000 is configs[1]): synthetic 3.1 Gbp reference (24
contigs, i.i.d. ACGT + 5 % planted diverged repeats), wgsim-style 2x150 bp pairs (insert N(400,50), 1 % substitutions, 0.1 % indels,
0.1 % N), chunked with the reference hosts' rule at -K 100000000 (333 334 pairs per chunk).  A "step" is one mem_process_seqs call
on one chunk, made in its chunk-job form (b200_align_seqs_begin / b200_align_fastq_begin + b200_align_chunk_end,
include/mpibwa_b200.h) so that several chunks are in flight and their kernels share the SMs.  The timed region covers exactly K
chunks, first begin to last end; the K chunks' fastq bytes sit in private page-locked host buffers when it starts.

  value  read pairs/s through mem_process_seqs with the chunk's encoded reads already resident in HBM
  e2e    read pairs/s from raw fastq bytes in (page-locked) host memory to SAM bytes in host memory: b200_align_fastq_begin /
         b200_align_chunk_end, every H2D/D2H copy inside - the reference-facing call with host buffers
  roofline / kernels   per device stage: algorithmic work / CUDA-event kernel time vs the measured peak (seeding: the reference
         algorithm's occ-block touches, counted by a pass of the general kernel over the same chunk)
  parity               untimed leg: the SAM of one timed chunk per rank against the compiled reference's (md5); the run fails otherwise
  cpu_baseline         the compiled reference (oracle/_ref/ref_driver) on the box's host cores, bounded sample

`--impl reference` times the reference's own CPU mem_process_seqs (oracle/_ref, all host threads) on bounded
samples of the same workload.  Multi-GPU: one process per GPU (torchrun), full index replica per GPU, reads sharded
by chunk, no collective on the data path (weak scaling: every rank aligns its own chunks).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CACHE = os.environ.get("B200_BENCH_CACHE", "/tmp/b200_bench_cache")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=8)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--ref-bp", type=int, default=3_100_000_000,
                   help="synthetic reference size: 100000000 = BASELINE configs[1], 3100000000 = configs[2] (human-sized)")
    p.add_argument("--pairs", type=int, default=1_000_000, help="simulated pairs per rank")
    p.add_argument("--read-len", type=int, default=150)
    p.add_argument("-K", type=int, default=100_000_000, dest="K")
    p.add_argument("--ref-sample-pairs", type=int, default=333_334, help="pairs per step of the CPU reference arm (one chunk)")
    p.add_argument("--cpu-baseline-pairs", type=int, default=1_000_000, help="pairs of the cpu_baseline sample of the b200 arm")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--parity-chunks", type=int, default=1, help="chunks per rank whose SAM is compared (md5) with the compiled reference's, untimed")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------ workload

def workload_dir(args):
    d = os.path.join(CACHE, "ref%d_L%d" % (args.ref_bp, args.read_len))
    os.makedirs(d, exist_ok=True)
    return d


def n_contigs(args):
    return 24 if args.ref_bp >= 2_000_000_000 else 4          # (configs[2]: 24 contigs with human-like length ratios)


def ensure_index(args):
    """synthetic reference + bwa-compatible index files, cached under CACHE"""
    import numpy as np
    from mpibwa_b200 import simulate, index_build
    d = workload_dir(args)
    prefix = os.path.join(d, "ref.fa")
    if not os.path.exists(prefix + ".done"):
        t = time.time()
        names, lengths, codes = simulate.make_reference(args.ref_bp, n_contigs(args), seed=1)
        t1 = time.time()
        index_build.build_index_from_codes(prefix, names, lengths, codes)
        t2 = time.time()
        np.save(os.path.join(d, "codes.npy"), codes)
        np.save(os.path.join(d, "lengths.npy"), lengths)
        open(prefix + ".done", "w").write("ok")
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.empty_cache()          # the builder's scratch goes back to the device before the library allocates
        except ImportError:
            pass
        log("[bench] %d bp reference simulated in %.1f s, index built in %.1f s, files written in %.1f s"
            % (args.ref_bp, t1 - t, t2 - t1, time.time() - t2))
    return prefix


def ensure_reads(args, rank, n_pairs, tag="r"):
    import numpy as np
    from mpibwa_b200 import simulate
    d = workload_dir(args)
    f1 = os.path.join(d, "%s%d_n%d_1.fq" % (tag, rank, n_pairs))
    f2 = f1[:-4] + "2.fq"
    if not (os.path.exists(f1 + ".done")):
        t = time.time()
        codes = np.load(os.path.join(d, "codes.npy"), mmap_mode="r")
        lengths = np.load(os.path.join(d, "lengths.npy"))
        r1, r2 = simulate.simulate_pairs(np.asarray(codes), lengths, n_pairs, read_len=args.read_len, seed=2 + 1000 * rank)
        open(f1, "wb").write(r1)
        open(f2, "wb").write(r2)
        open(f1 + ".done", "w").write("ok")
        log("[bench] rank %d simulated %d pairs in %.1f s" % (rank, n_pairs, time.time() - t))
    return f1, f2


# ------------------------------------------------------------------------------------------------ clocks

class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.max_mhz, self.stop_flag = device, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv = None
            log("[bench] NVML unavailable:", e)

    def run(self):
        while not self.stop_flag and self.nv:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                m = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if m & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def result(self):
        self.stop_flag = True
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------ reference arm

def run_ref_driver(prefix, f1, f2, K, threads, digest=None):
    """runs oracle/_ref/ref_driver; digest: a hashlib object that is fed the SAM records it prints (else they are dropped)"""
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    cmd = [drv, "-t", str(threads), "-K", str(K), "-v", "1", prefix, f1, f2]
    if digest is None:
        r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
        err = r.stderr
    else:
        with subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE) as pr:
            t = threading.Thread(target=lambda: errbuf.append(pr.stderr.read()))
            errbuf = []
            t.start()
            n = 0
            for blk in iter(lambda: pr.stdout.read(1 << 22), b""):
                digest.update(blk)
                n += len(blk)
            t.join()
            if pr.wait() != 0:
                raise RuntimeError("ref_driver failed:\n" + errbuf[0].decode()[-2000:])
        err = errbuf[0].decode()
        digest.n_bytes = n
    chunk_sec = []
    for line in err.splitlines():
        if line.startswith("[ref_driver] chunk="):
            kv = dict(x.split("=") for x in line.split()[1:])
            chunk_sec.append((int(kv["reads"]), float(kv["sec"])))
        elif line.startswith("[ref_driver]"):
            kv = dict(x.split("=") for x in line.split()[1:])
            run_ref_driver.chunks = chunk_sec
            return int(kv["reads"]), float(kv["mem_process_seqs_sec"])
    raise RuntimeError("ref_driver printed no timing line:\n" + err[-2000:])


class Digest:
    """md5 + byte count of a SAM stream"""

    def __init__(self):
        import hashlib
        self.h, self.n_bytes = hashlib.md5(), 0

    def update(self, b):
        self.h.update(b)

    def hexdigest(self):
        return self.h.hexdigest()


def slice_fastq(src, dst, first_read, n_reads, rec_bytes):
    with open(src, "rb") as fi:
        fi.seek(first_read * rec_bytes)
        data = fi.read(n_reads * rec_bytes)
    with open(dst, "wb") as fo:
        fo.write(data)


def record_bytes(path):
    with open(path, "rb") as fh:
        return sum(len(fh.readline()) for _ in range(4))


def reference_arm(args, prefix):
    """the reference's own CPU implementation of the path, all host threads, bounded samples of the workload: ONE run of
    oracle/_ref/ref_driver (the index is loaded once) over W + K chunks of the rank-0 read set, chunked by the reference hosts' rule
    at the same -K; the first W chunks are the warm-up, the time of the K others is what is reported"""
    cores = os.cpu_count() or 1
    f1, f2 = ensure_reads(args, 0, args.pairs)
    rb = record_bytes(f1)
    n = min(args.ref_sample_pairs, args.pairs)
    d = workload_dir(args)
    s1, s2 = os.path.join(d, "refsample_1.fq"), os.path.join(d, "refsample_2.fq")
    total = args.warmup + args.steps
    with open(s1, "wb") as o1, open(s2, "wb") as o2:
        for step in range(total):
            first = (step * n) % max(1, args.pairs - n)
            for src, dst in ((f1, o1), (f2, o2)):
                with open(src, "rb") as fi:
                    fi.seek(first * rb)
                    dst.write(fi.read(n * rb))
    run_ref_driver(prefix, s1, s2, args.K, cores)
    chunks = run_ref_driver.chunks
    assert len(chunks) == total, (len(chunks), total)
    for step, (reads, sec) in enumerate(chunks):
        log("[bench] reference step %d: %d reads in %.3f s" % (step, reads, sec))
    total_pairs = sum(r for r, _ in chunks[args.warmup:]) // 2
    total_s = sum(t for _, t in chunks[args.warmup:])
    v = total_pairs / total_s
    sample = "%d pairs per step (one mem_process_seqs call, -t %d) of the same simulated read set" % (n, cores)
    return {"metric": "aligned 2x150bp read pairs/sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, "reference CPU mem_process_seqs (oracle/_ref, unmodified sources, gcc -O2)"),
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def workload_config(args, what):
    which = ("configs[1]" if args.ref_bp == 100_000_000 else "configs[2] reference (human-sized), read set of configs[1]" if args.ref_bp >= 3_000_000_000
             else "configs[1] read set against a reference %dx larger than configs[1]'s, beyond the 126 MB L2 (configs[2]'s 3.1 Gbp: --ref-bp 3100000000)" % (args.ref_bp // 100_000_000))
    return {"workload": "%s: %d synthetic 2x%dbp pairs per GPU vs synthetic %d bp reference (%d contigs, 5%% planted repeats), -K %d"
                        % (which, args.pairs, args.read_len, args.ref_bp, n_contigs(args), args.K),
            "step": "one mem_process_seqs call on one chunk (%d pairs at full size) in its chunk-job form, up to four chunks in flight, each as one batch per kernel" % ((args.K // 2) // args.read_len + 1),
            "path": what, "cache_policy": "steps cycle over the rank's %d chunks (consecutive steps never align the same chunk; with up to four chunks in "
                                          "flight a chunk may be in flight twice); index (%d MB; with the structures derived from it at upload - k-mer tables, suffix array + inverse, Bloom filters - 20 times that) + per-chunk buffers exceed the 126 MB L2; an L2-sized buffer "
                                          "is rewritten before each timed region" % (max(1, -(-args.pairs // ((args.K // 2) // args.read_len + 1))), int(args.ref_bp * 1.75e-6))}


# ------------------------------------------------------------------------------------------------ b200 arm

def int32_peak_gops(torch):
    """measured int32 add/max issue rate of this GPU (alu pipe), via the library's micro-benchmark"""
    import mpibwa_b200 as M
    lib = M.load()
    return lib.b200_int32_peak(torch.cuda.current_device())


def kernel_table(agg, K, i32_peak, hbm_peak, ref_ratio=1.0):
    """per device stage: algorithmic work / CUDA-event kernel time vs the measured peak.  ref_ratio: occ blocks the REFERENCE's
    seeding loops touch on these reads / occ blocks the sweep kernels touched or looked up (measured on the isolated chunk)"""
    kern = {}
    if agg["ms_k_extend_dp"] > 0:
        gcups = agg["extend_cells"] / agg["ms_k_extend_dp"] / 1e6
        kern["ksw_extend2"] = {"bound": "int32_issue", "ms_per_step": agg["ms_k_extend_dp"] / K, "cells_per_step": agg["extend_cells"] / K,
                               "jobs_per_step": agg["n_extend_jobs"] / K, "rounds_per_step": agg["n_extend_rounds"] / K,
                               "stage_ms_per_step": agg["ms_k_extend"] / K, "gcups": gcups, "gcups_whole_stage": agg["extend_cells"] / agg["ms_k_extend"] / 1e6,
                               "achieved": gcups * 14, "unit": "Gop/s (14 int32 ops per cell)", "peak": i32_peak,
                               "frac": (gcups * 14 / i32_peak) if i32_peak else None}
    if agg["ms_k_smem"] > 0:
        gbs = 64.0 * agg["fm_occ_blocks"] * ref_ratio / agg["ms_k_smem"] / 1e6
        kern["smem_seeding"] = {"bound": "hbm", "ms_per_step": agg["ms_k_smem"] / K, "bytes_per_step": 64.0 * agg["fm_occ_blocks"] * ref_ratio / K,
                                "touched_bytes_per_step": 64.0 * agg["fm_occ_blocks"] / K,
                                "achieved": gbs, "unit": "GB/s", "peak": hbm_peak, "frac": gbs / hbm_peak,
                                "bytes_note": "bytes_per_step = ALGORITHMIC traffic (SURVEY.md 8d): 64 B per distinct occ block of the reference layout per bwt_occ4 "
                                              "of the reference's seeding loops on these reads, counted by running the general kernel (the reference's loops, "
                                              "B200_SEED_KERNEL=lanes) over the isolated chunk; touched_bytes_per_step = the same unit over what the sweep kernels "
                                              "actually extended or looked up in the k-mer tables (they skip most of the backward rows)"}
    if agg["ms_k_sa"] > 0:
        b = 64.0 * agg["fm_sa_steps"] + 8.0 * agg["fm_sa_lookups"]
        gbs = b / agg["ms_k_sa"] / 1e6
        kern["sa_lookup"] = {"bound": "hbm", "ms_per_step": agg["ms_k_sa"] / K, "bytes_per_step": b / K, "achieved": gbs, "unit": "GB/s",
                             "peak": hbm_peak, "frac": gbs / hbm_peak, "lookups_per_step": agg["fm_sa_lookups"] / K,
                             "bytes_note": "64 B per bwt_invPsi step the kernel walked + 8 B per look-up; with the whole suffix array in HBM "
                                           "(the default) a look-up walks nothing: one 5-byte read"}
    if agg["ms_k_sw"] > 0:
        gc = agg["sw_cells"] / agg["ms_k_sw"] / 1e6
        kern["ksw_align2"] = {"bound": "int32_issue", "ms_per_step": agg["ms_k_sw"] / K, "cells_per_step": agg["sw_cells"] / K, "gcups": gc,
                              "achieved": gc * 11, "unit": "Gop/s (11 ops per cell)", "peak": i32_peak, "frac": (gc * 11 / i32_peak) if i32_peak else None}
    if agg["ms_k_global"] > 0:
        gc = agg["global_cells"] / agg["ms_k_global"] / 1e6
        kern["ksw_global2"] = {"bound": "int32_issue", "ms_per_step": agg["ms_k_global"] / K, "cells_per_step": agg["global_cells"] / K,
                               "jobs_per_step": agg["n_global_jobs"] / K, "host_fallbacks_per_step": agg.get("n_global_host", 0) / K,
                               "gcups": gc, "achieved": gc * 14,
                               "unit": "Gop/s (14 int32 ops per cell, src/ksw.c:546-566)", "peak": i32_peak, "frac": (gc * 14 / i32_peak) if i32_peak else None}
    if agg.get("ms_k_chain", 0) > 0:
        kern["chaining"] = {"bound": "latency (one read per lane, pointer chasing)", "ms_per_step": agg["ms_k_chain"] / K,
                            "seeds_per_step": agg["n_seeds"] / K, "chains_per_step": agg["n_chains"] / K}
    return kern


def gpu_locality(torch, device):
    """(numa node, set of CPUs) the PCI device of a GPU is local to, from sysfs; (-1, empty) when the platform does not say"""
    try:
        p = torch.cuda.get_device_properties(device)
        base = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        return node, cpus & os.sched_getaffinity(0)
    except Exception:
        return -1, set()


def bind_near_gpu(torch, dist, local_rank, world, threads_per_rank):
    """One MPI rank per GPU is normally started bound to the socket its GPU hangs on (mpirun --map-by / --bind-to, srun --cpu-bind);
    torchrun binds nothing, so a rank's threads and its page-locked buffers may sit on the other socket and every H2D / D2H copy of
    the chunk loop crosses the inter-socket link.  Each rank restricts itself to the CPUs local to its GPU - only if every rank's GPU
    reports a node, the GPUs sit on at least two nodes and each node has enough CPUs for the ranks on it; B200_BENCH_NUMA=0 disables."""
    if world <= 1 or os.environ.get("B200_BENCH_NUMA", "1") == "0":
        return None
    node, cpus = gpu_locality(torch, local_rank)
    mine = torch.tensor([node, len(cpus)], dtype=torch.int64, device="cuda")
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    every = [t.tolist() for t in every]
    nodes = [n for n, _ in every]
    ok = all(n >= 0 for n in nodes) and len(set(nodes)) >= 2
    for n, c in every:
        ok = ok and c >= nodes.count(n) * max(2, threads_per_rank)
    if ok:
        try:
            os.sched_setaffinity(0, cpus)
        except OSError:
            ok = False
    return {"bound": bool(ok), "node": node, "cpus": len(cpus), "nodes_of_ranks": nodes}


def main():
    # rank 0 prints exactly ONE line on stdout: everything libraries write there meanwhile (NCCL's version banner ...) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        prefix = ensure_index(args)
        emit(reference_arm(args, prefix))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import mpibwa_b200 as M
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    numa = bind_near_gpu(torch, dist, local_rank, world, max(1, (os.cpu_count() or 1) // world))
    if numa:
        log("[bench] rank %d: GPU %d on NUMA node %d with %d local CPUs -> %s" % (rank, local_rank, numa["node"], numa["cpus"], "bound" if numa["bound"] else "not bound"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        prefix = ensure_index(args)
    barrier()
    prefix = os.path.join(workload_dir(args), "ref.fa")
    f1, f2 = ensure_reads(args, rank, args.pairs)
    n_threads = max(1, (os.cpu_count() or 1) // world)
    if os.environ.get("B200_BENCH_THREADS"):        # (experiments: the host-thread budget a rank gets at a larger N, on one GPU)
        n_threads = int(os.environ["B200_BENCH_THREADS"])
    al = M.Aligner(prefix, device=local_rank, n_threads=n_threads, verbose=1)
    lib = al.lib
    fq1, fq2 = open(f1, "rb").read(), open(f2, "rb").read()
    rb1, rb2 = record_bytes(f1), record_bytes(f2)
    # plan chunks once (untimed) with the reference hosts' rule
    b1, s1, n1 = al.parse(fq1)
    b2, s2, n2 = al.parse(fq2)
    ends = al.plan(n1, s1, s2, args.K)
    lib.b200_free(s1); lib.b200_free(s2)
    del b1, b2
    chunks = []
    beg = 0
    for e in ends:
        chunks.append((beg, e))
        beg = e
    log("[bench] rank %d: %d pairs in %d chunks, %d host threads" % (rank, n1, len(chunks), n_threads))
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    max_pairs = max(e - b for b, e in chunks)
    N_SLOTS = int(os.environ.get("B200_INFLIGHT", "4"))      # chunks in flight (the library has eight chunk slots and runs B200_INFLIGHT jobs at a time)
    # the in-flight chunks need their own fastq buffers (the parse is in place and the job reads the records later)
    n_buf = N_SLOTS + 1

    def pinned_bytes(n):
        """page-locked host buffer from the library's pool (what an MPI host would read its fastq chunk into), as a numpy view"""
        ptr = lib.b200_big_alloc(n)
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr))

    read_bufs = [[pinned_bytes(max_pairs * rb + 1) for rb in (rb1, rb2)] for _ in range(n_buf)]
    buf_turn = [0]

    def chunk_bytes(c):
        """private, writable copy of the chunk's fastq bytes (what the MPI host gets from its file read); the parse is in place"""
        b, e = chunks[c % len(chunks)]
        bufs = read_bufs[buf_turn[0] % n_buf]
        buf_turn[0] += 1
        out = []
        for k, (fq, rb) in enumerate(((fq1, rb1), (fq2, rb2))):
            nb = (e - b) * rb
            a = bufs[k]
            a[:nb] = np.frombuffer(fq, dtype=np.uint8, count=nb, offset=b * rb)
            a[nb] = 0
            out.append((a, nb))
        return out[0], out[1], e - b

    STAT_KEYS = ("ms_k_chain", "n_seeds", "n_chains", "ms_k_smem", "ms_k_sa", "ms_k_extend", "ms_k_extend_dp", "n_extend_rounds", "ms_k_sw", "extend_cells", "n_extend_jobs", "sw_cells", "n_sw_jobs",
                 "fm_occ_blocks", "fm_sa_steps", "fm_sa_lookups", "n_launches", "h2d_bytes", "d2h_bytes", "ms_seed", "ms_chain_host",
                 "ms_extend", "ms_regs_host", "ms_rescue", "ms_sam_host", "ms_total", "n_intv",
                 "ms_sam_plan", "ms_global", "ms_k_global", "n_global_jobs", "global_cells", "n_global_host",
                 "ms_k_finish", "ms_upload", "ms_deliver", "n_patch_reads", "n_rescue_rounds", "n_global_rerun", "n_aln_slots", "sam_bytes")

    def e2e_begin(c, raw=None):
        """raw fastq bytes -> chunk job (b200_align_fastq_begin: parse in place, interleave, align, concatenate - all on the library's job thread)"""
        a1, a2, n = raw if raw is not None else chunk_bytes(c)
        job = lib.b200_align_fastq_begin(al.opt, al.idx, 0, C.c_void_p(a1[0].ctypes.data), a1[1], C.c_void_p(a2[0].ctypes.data), a2[1])
        return (job, a1, a2, n)

    def e2e_end(h, st, digest=None):
        job, a1, a2, n = h
        sam = C.c_void_p()
        sam_len = C.c_int64()
        lib.b200_align_chunk_end(job, C.byref(sam), C.byref(sam_len), C.byref(st))
        out_len = sam_len.value
        if digest is not None:
            digest.update(C.string_at(sam, out_len))
            digest.n_bytes += out_len
        lib.b200_free(sam)
        return n, out_len

    DEPTH = N_SLOTS  # chunks the host loop keeps begun ahead of the one it waits for (the library runs B200_INFLIGHT at a time)

    def e2e_run(first, count, on_stats=None, raw=None):
        """`count` chunks from fastq bytes in host memory to SAM bytes in host memory as chunk jobs: begin(i), ... end(i - DEPTH + 1).  raw: the chunks' private fastq buffers when they were filled before the timed region."""
        pairs = out_bytes = 0
        pending = []
        st = M.b200_stats_t()

        def finish():
            nonlocal pairs, out_bytes
            n, ob = e2e_end(pending.pop(0), st)
            pairs += n; out_bytes += ob
            if on_stats:
                on_stats(st.as_dict())

        tb = te = 0.0
        for s in range(count):
            t_ = time.time()
            pending.append(e2e_begin(first + s, raw[s] if raw else None))
            tb += time.time() - t_
            if len(pending) >= DEPTH:
                t_ = time.time()
                finish()
                te += time.time() - t_
        while pending:
            t_ = time.time()
            finish()
            te += time.time() - t_
        if os.environ.get("B200_BENCH_DEBUG"):
            log("[bench] e2e_run: %d chunks, parse+begin %.1f ms, end (wait + free) %.1f ms" % (count, 1e3 * tb, 1e3 * te))
        return pairs, out_bytes

    def resident_group(first, count, ev0, ev1, on_stats=None):
        """`count` (<= N_SLOTS) chunks parsed, encoded and resident in HBM before the timed region, then aligned as chunk
        jobs (b200_align_seqs_begin: mem_process_seqs with the chunk's SAM text as one buffer)"""
        held = []
        for s in range(count):
            a1, a2, n = chunk_bytes(first + s)
            k1, p1, m1 = al.parse_array(*a1)
            k2, p2, m2 = al.parse_array(*a2)
            seqs = lib.b200_chunk_seqs(m1, p1, p2)
            lib.b200_stage_reads(al.opt, al.idx, 2 * m1, seqs)
            held.append((seqs, p1, p2, m1, n))
        flush_buf.add_(1)
        torch.cuda.synchronize()
        ev0.record()
        jobs = [lib.b200_align_seqs_begin(al.opt, al.idx, 0, 2 * m1, seqs, None) for seqs, p1, p2, m1, n in held]
        st = M.b200_stats_t()
        sams = []
        for j in jobs:
            sam, sam_len = C.c_void_p(), C.c_int64()
            lib.b200_align_chunk_end(j, C.byref(sam), C.byref(sam_len), C.byref(st))
            sams.append(sam)
            if on_stats:
                on_stats(st.as_dict())
        ev1.record()
        torch.cuda.synchronize()
        pairs = 0
        for sam in sams:
            lib.b200_free(sam)
        for seqs, p1, p2, m1, n in held:
            lib.b200_free(seqs); lib.b200_free(p1); lib.b200_free(p2)
            pairs += n
        return pairs, ev0.elapsed_time(ev1)

    # ---- warm-up: W steps end to end, then one untimed resident group so that every chunk slot has its device buffers
    e2e_run(0, args.warmup)
    resident_group(0, N_SLOTS, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    # ---- timed: device-resident inputs
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    agg = {k: 0.0 for k in STAT_KEYS}

    def add_stats(st):
        for k in STAT_KEYS:
            agg[k] += st[k]

    res_pairs, res_ms = 0, 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    done = 0
    while done < args.steps:
        g = min(N_SLOTS, args.steps - done)
        n, ms = resident_group(args.warmup + done, g, ev0, ev1, add_stats)
        res_pairs += n
        res_ms += ms
        done += g
    barrier()
    # ---- timed: end to end from host fastq bytes to host SAM bytes
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    io = {"h2d": 0, "d2h": 0}

    def add_io(st):
        io["h2d"] += st["h2d_bytes"]
        io["d2h"] += st["d2h_bytes"]
        if os.environ.get("B200_BENCH_DEBUG"):
            log("[bench] e2e job: " + " ".join("%s %.1f" % (k[3:], st[k]) for k in ("ms_total", "ms_upload", "ms_seed", "ms_chain_host", "ms_extend", "ms_regs_host", "ms_rescue", "ms_sam_plan", "ms_global", "ms_sam_host", "ms_deliver", "ms_k_finish")))

    # the K chunks' fastq bytes sit in private host buffers (what the host's file read leaves) when the timed region starts
    # (as long as K private copies fit in 6 GB of page-locked memory; beyond that every step copies its chunk into one of
    # B200_INFLIGHT + 1 recycled buffers inside the timed region - a host-to-host copy the contract does not ask for and that
    # eight ranks on one box compete over)
    raw = None
    if args.steps * (max_pairs * (rb1 + rb2) + 2) <= 6 << 30:
        raw = []
        for s in range(args.steps):
            b, e = chunks[(args.warmup + s) % len(chunks)]
            pair = []
            for fq, rb in ((fq1, rb1), (fq2, rb2)):
                nb = (e - b) * rb
                a = pinned_bytes(nb + 1)
                a[:nb] = np.frombuffer(fq, dtype=np.uint8, count=nb, offset=b * rb)
                a[nb] = 0
                pair.append((a, nb))
            raw.append((pair[0], pair[1], e - b))
    barrier()
    flush_buf.add_(1)
    e0.record()
    t0 = time.time()
    e2e_pairs, sam_bytes = e2e_run(args.warmup, args.steps, add_io, raw)
    e1.record()
    barrier()
    h2d, d2h = io["h2d"], io["d2h"]
    e2e_ms = e0.elapsed_time(e1)
    wall_ms = 1e3 * (time.time() - t0)
    clocks = sampler.result()
    # ---- untimed parity leg: the SAM of the first timed chunk(s) of this rank, end to end through the same call, against the
    # compiled reference's (oracle/_ref/ref_driver) on the same fastq slice - md5 of the record bytes
    parity = {"chunks": 0, "identical": None, "checked_against": None}
    ref_drv_path = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if args.parity_chunks > 0 and os.path.exists(ref_drv_path):
        ident, nb = True, 0
        for pc in range(args.parity_chunks):
            c = args.warmup + pc
            b, e = chunks[c % len(chunks)]
            mine = Digest()
            st_ = M.b200_stats_t()
            e2e_end(e2e_begin(c), st_, mine)
            d = workload_dir(args)
            s1, s2 = os.path.join(d, "parity%d_1.fq" % rank), os.path.join(d, "parity%d_2.fq" % rank)
            slice_fastq(f1, s1, b, e - b, rb1)
            slice_fastq(f2, s2, b, e - b, rb2)
            theirs = Digest()
            run_ref_driver(prefix, s1, s2, args.K, n_threads, theirs)
            same = mine.hexdigest() == theirs.hexdigest() and mine.n_bytes == theirs.n_bytes
            log("[bench] rank %d parity chunk %d: %d pairs, %d SAM bytes, md5 %s vs reference %s -> %s"
                % (rank, c, e - b, mine.n_bytes, mine.hexdigest(), theirs.hexdigest(), "identical" if same else "DIFFERENT"))
            ident = ident and same
            nb += mine.n_bytes
        parity = {"chunks": args.parity_chunks, "identical": ident, "sam_bytes": nb,
                  "checked_against": "oracle/_ref/ref_driver (unmodified reference sources) on the same fastq slice, md5 of the SAM records"}
    # ---- untimed extra pass: ONE chunk alone on the device - kernel-isolated efficiency
    resident_group(args.warmup + args.steps, 1, ev0, ev1)          # first one grows the device buffers to whole-chunk size
    iso = []
    os.environ["B200_EXT_RECORD"] = "1"            # keep this chunk's ksw_extend2 job list for the one-batch replay below
    resident_group(args.warmup + args.steps + 1, 1, ev0, ev1, iso.append)
    os.environ.pop("B200_EXT_RECORD", None)
    st_iso = iso[0]
    # the algorithmic FM-index traffic of that chunk: the general seeding kernel runs the reference's loops extension for extension
    iso_ref = []
    os.environ["B200_SEED_KERNEL"] = "lanes"
    resident_group(args.warmup + args.steps + 1, 1, ev0, ev1, iso_ref.append)
    os.environ.pop("B200_SEED_KERNEL", None)
    ref_ratio = iso_ref[0]["fm_occ_blocks"] / max(1, st_iso["fm_occ_blocks"])
    # kernel-isolated ksw_extend2 as BASELINE configs[1] words it: the exact job list of the chunk, replayed as one batch
    rp_cells, rp_jobs = C.c_int64(), C.c_int64()
    lib.b200_ext_replay(al.opt, C.byref(rp_cells), C.byref(rp_jobs))          # (warm-up)
    rp_ms = lib.b200_ext_replay(al.opt, C.byref(rp_cells), C.byref(rp_jobs))

    # ---- reduce over ranks: MAX of times, SUM of pairs
    bad = 0.0 if parity["identical"] in (True, None) else 1.0
    t = torch.tensor([res_ms, e2e_ms, bad], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([res_pairs, e2e_pairs, parity["chunks"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    res_ms_max, e2e_ms_max, any_bad = t.tolist()
    res_pairs_all, e2e_pairs_all, parity_chunks_all = cnt.tolist()
    if parity["identical"] is not None:
        parity["identical"] = any_bad == 0.0
        parity["chunks"] = int(parity_chunks_all)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
    i32_peak = int32_peak_gops(torch)
    # what the device delivers to random 32-byte sectors of a table of the index's size (the FM-index kernels' pattern)
    occ_bytes = int(args.ref_bp)                           # occ sectors: 0.5 byte per BWT symbol, 2 symbols per base
    rnd_peak = lib.b200_hbm_random_sector_peak(torch.cuda.current_device(), max(occ_bytes, 1 << 26))
    K = args.steps
    kern = kernel_table(agg, K, i32_peak, hbm_peak, ref_ratio)
    kern_iso = kernel_table(st_iso, 1, i32_peak, hbm_peak, ref_ratio)
    # The chunks in flight share the SMs (B200_TURN=0, the default): the CUDA-event duration of a kernel in the timed region includes
    # the time it spent sharing the device with the kernels of other chunks, so the roofline of the dominant kernel is taken from
    # the pass that runs ONE chunk alone (same process, same inputs, CUDA events on the launching stream); `kernels` keeps the
    # timed-region durations.
    concurrent = os.environ.get("B200_TURN", "0") == "0"
    src = kern_iso if concurrent else kern
    dom = max(src, key=lambda k: src[k]["ms_per_step"]) if src else None
    roof = None
    if dom:
        kd = src[dom]
        traffic = ncu_detail = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get(dom)                          # DRAM bytes per launch from the committed ncu --set full capture
            ncu_detail = tj.get(dom + "_detail")           # ... and its DRAM GB/s / L2 hit rates (north_star asks for both beside the roofline)
        roof = {"kernel": dom, "bound": kd["bound"], "achieved": kd["achieved"], "peak": kd["peak"], "unit": kd["unit"], "frac": kd["frac"],
                "traffic": traffic, "ncu": ncu_detail, "algorithmic_bytes_per_launch": kd.get("bytes_per_step"), "touched_bytes_per_launch": kd.get("touched_bytes_per_step"),
                "bytes_note": kd.get("bytes_note"), "peak_source": hbm_src if kd["bound"] == "hbm" else "int32 add/max issue rate measured live by b200_int32_peak()",
                "random_sector_peak": {"value": rnd_peak, "unit": "GB/s", "table_bytes": occ_bytes,
                                       "frac_of_it": (kd["achieved"] / rnd_peak) if kd["bound"] == "hbm" and rnd_peak else None,
                                       "what": "measured live (b200_hbm_random_sector_peak): independent 256-bit loads of uniformly random 32-byte sectors of a table "
                                               "the size of the occ table - the access pattern of the seeding / SA kernels; `peak` above stays the streaming copy bandwidth"},
                "measured": ("one chunk alone on the device (untimed extra pass of this run; in the timed region up to %d chunks share the SMs "
                             "and the per-launch durations overlap)" % N_SLOTS) if concurrent else "timed region"}
    line = {
        "metric": "aligned 2x150bp read pairs/sec", "value": res_pairs_all / (res_ms_max * 1e-3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, "mem_process_seqs in its chunk-job form (b200_align_seqs_begin / b200_align_fastq_begin + b200_align_chunk_end) through the C ABI of libmpibwa_b200.so (ctypes), %d host threads per rank" % n_threads),
        "e2e": {"value": e2e_pairs_all / (e2e_ms_max * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d // args.steps,
                "d2h_bytes_per_step": d2h // args.steps, "sam_bytes_per_step": sam_bytes // args.steps, "wall_ms_rank0": wall_ms},
        "gpu_launches": int(agg["n_launches"]), "clocks": clocks, "roofline": roof, "kernels": kern,
        "kernels_isolated": kern_iso,
        "kernels_note": "kernels: CUDA-event kernel times inside the timed region (chunk jobs: one whole-chunk batch per kernel, several chunks "
                        "in flight sharing the SMs); kernels_isolated / roofline: one extra untimed chunk run alone, the kernel-isolated figures of "
                        "BASELINE configs[1]; ksw_extend2_gcups is the isolated one (DP kernels of all chain2aln rounds of the chunk), "
                        "ksw_extend2_gcups_one_batch / ksw_extend2_replay the same job list replayed as one batch",
        "ksw_extend2_gcups": kern_iso.get("ksw_extend2", {}).get("gcups"),
        "ksw_extend2_gcups_one_batch": (rp_cells.value / rp_ms / 1e6) if rp_ms > 0 else None,
        "ksw_extend2_replay": {"what": "every ksw_extend2 job of one chunk (recorded from the pipeline) as ONE batch through the DP kernels: no chain2aln rounds "
                                       "in between, so no launch with too few jobs to fill the chip",
                               "jobs": rp_jobs.value, "cells": rp_cells.value, "ms": rp_ms,
                               "gcups": (rp_cells.value / rp_ms / 1e6) if rp_ms > 0 else None,
                               "frac": (rp_cells.value / rp_ms / 1e6 * 14 / i32_peak) if rp_ms > 0 and i32_peak else None},
        "stage_ms_per_step": {k: agg[k] / K for k in ("ms_upload", "ms_seed", "ms_chain_host", "ms_extend", "ms_regs_host", "ms_rescue", "ms_sam_plan", "ms_global", "ms_sam_host", "ms_deliver", "ms_total")},
        "stage_ms_note": "walls per chunk job incl. waiting for the device turn: upload = encode + read text staging + H2D; seed; chain; extend; regs_host = de-duplication; rescue = insert-size statistics + mate rescue; sam_plan = pairing + record plan; global = CIGAR stage; sam_host = NM/MD + SAM text; deliver = D2H + hand-over (all but upload/deliver are device stages; the names are those of round 1's records)",
        "finish_kernels_ms_per_step": agg["ms_k_finish"] / K,
        "host_threads": n_threads, "numa_binding": numa, "parity": parity,
    }
    if world == 1 and not args.no_cpu_baseline and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_driver")):
        cores = os.cpu_count() or 1
        n = min(args.cpu_baseline_pairs, args.pairs)
        d = workload_dir(args)
        s1, s2 = os.path.join(d, "cpusample_1.fq"), os.path.join(d, "cpusample_2.fq")
        slice_fastq(f1, s1, 0, n, rb1)
        slice_fastq(f2, s2, 0, n, rb2)
        reads, sec = run_ref_driver(prefix, s1, s2, args.K, cores)
        line["cpu_baseline"] = {"value": (reads // 2) / sec, "unit": "pairs/s", "cores": cores, "kind": "reference",
                                "sample": "first %d pairs of the workload, chunked at -K %d, -t %d (%.2f s in mem_process_seqs)" % (n, args.K, cores, sec)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    if parity["identical"] is False:
        log("[bench] FAILED: the SAM of the benchmarked workload differs from the compiled reference's")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
