#!/usr/bin/env python
"""Sweep an environment knob of the device stages over the bench workload and print the CUDA-event kernel times.
usage: tune_env.py NAME v1 v2 ...   (e.g. B200_EXT_WARP_MAX 1024 2048 4096; B200_TUNE_REF_BP picks the reference size); one whole chunk per setting, B200_LANES=1"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    name, values = sys.argv[1], sys.argv[2:]
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    args.pairs = 333_334
    if os.environ.get("B200_TUNE_REF_BP"):
        args.ref_bp = int(os.environ["B200_TUNE_REF_BP"])
    prefix = bench.ensure_index(args)
    f1, f2 = bench.ensure_reads(args, 0, args.pairs)
    import mpibwa_b200 as M
    al = M.Aligner(prefix, device=0, n_threads=os.cpu_count() or 1, verbose=1)
    fq1, fq2 = open(f1, "rb").read(), open(f2, "rb").read()
    os.environ["B200_LANES"] = "1"
    al.align(fq1, fq2, K=args.K)
    for v in ["(default)"] + values + ["(default)"]:
        if v == "(default)":
            os.environ.pop(name, None)
        else:
            os.environ[name] = v
        al.align(fq1, fq2, K=args.K)
        st = al.stats()
        print("   stage walls (one chunk alone): seed %.1f chain %.1f extend %.1f regs %.1f rescue %.1f sam %.1f [plan %.1f cigar %.1f]" % (
            st["ms_seed"], st["ms_chain_host"], st["ms_extend"], st["ms_regs_host"], st["ms_rescue"], st["ms_sam_host"], st["ms_sam_plan"], st["ms_global"]), flush=True)
        print("%s=%s  smem %.2f ms  sa %.2f  chain %.2f  ext_dp %.2f  ext_stage %.2f  sw %.2f  global %.2f  total %.1f" % (
            name, v, st["ms_k_smem"], st["ms_k_sa"], st["ms_k_chain"], st["ms_k_extend_dp"], st["ms_k_extend"], st["ms_k_sw"], st["ms_k_global"], st["ms_total"]), flush=True)


if __name__ == "__main__":
    main()
