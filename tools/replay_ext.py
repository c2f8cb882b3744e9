#!/usr/bin/env python
"""One chunk of the bench workload through the pipeline with the ksw_extend2 job list recorded, then the one-batch replay
(b200_ext_replay) - the short program profiled for the kernel-isolated ksw_extend2 figure (profiles/)."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    args.pairs = 333_334
    prefix = bench.ensure_index(args)
    f1, f2 = bench.ensure_reads(args, 0, args.pairs)
    import mpibwa_b200 as M
    al = M.Aligner(prefix, device=0, n_threads=os.cpu_count() or 1, verbose=1)
    os.environ["B200_LANES"] = "1"
    os.environ["B200_EXT_RECORD"] = "1"
    al.align(open(f1, "rb").read(), open(f2, "rb").read(), K=1 << 40)
    cells, jobs = C.c_int64(), C.c_int64()
    for _ in range(3):
        ms = al.lib.b200_ext_replay(al.opt, C.byref(cells), C.byref(jobs))
        print("replay: %d jobs, %d cells, %.3f ms, %.1f GCUPS" % (jobs.value, cells.value, ms, cells.value / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
