"""Regenerates tests/golden/ksw_vectors.json and tests/golden/sam_md5.json from the UNMODIFIED reference compiled
into oracle/_ref (run in the build container where /root/reference exists: `python tests/golden/make_golden.py`)."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzzgen  # noqa: E402
import oracle_lib as OL  # noqa: E402


def main():
    ref = OL.Reference()
    out = {"extend": [], "align": []}
    for c in fuzzgen.extend_cases(101, 300, max_q=120, max_t=250):
        a, b, od, ed, oi, ei = c["params"]
        r = ref.extend(c["q"], c["t"], OL.default_mat(a, b), od, ed, oi, ei, c["w"], c["end_bonus"], c["zdrop"], c["h0"])
        out["extend"].append(dict(q=c["q"].tolist(), t=c["t"].tolist(), a=a, b=b, o_del=od, e_del=ed, o_ins=oi, e_ins=ei,
                                  w=c["w"], end_bonus=c["end_bonus"], zdrop=c["zdrop"], h0=c["h0"], out=list(r)))
    for sixteen in (False, True):
        for c in fuzzgen.align_cases(103 + sixteen, 60, sixteen):
            a, b, od, ed, oi, ei = c["params"]
            r = ref.align(c["q"], c["t"], OL.default_mat(a, b), od, ed, oi, ei, c["xtra"])
            out["align"].append(dict(q=c["q"].tolist(), t=c["t"].tolist(), a=a, b=b, o_del=od, e_del=ed, o_ins=oi, e_ins=ei,
                                     xtra=c["xtra"], out=list(r)))
    with open(os.path.join(HERE, "ksw_vectors.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    # SAM digests of the reference on its own example data (the four input shapes of tools/check_examples.sh)
    import gzip
    import tarfile
    md5 = {}
    with tempfile.TemporaryDirectory() as d:
        with tarfile.open(os.path.join(HERE, "examples", "hg19.small.tar.gz")) as tf:
            tf.extractall(d)
        idx = [os.path.join(b, f) for b, _, fs in os.walk(d) for f in fs if f.endswith(".fa")][0]
        fq = {}
        for key in ("R1_10K", "R2_10K", "R1_10K_TRIM", "R2_10K_TRIM"):
            fq[key] = os.path.join(d, key + ".fq")
            with gzip.open(os.path.join(HERE, "examples", "HCC1187C_%s.fastq.gz" % key), "rb") as fi, open(fq[key], "wb") as fo:
                fo.write(fi.read())
        drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
        runs = {"pe": ["-H", idx, fq["R1_10K"], fq["R2_10K"]], "pe_K": ["-K", "500000", idx, fq["R1_10K"], fq["R2_10K"]],
                "trim": ["-T", "-K", "700000", idx, fq["R1_10K_TRIM"], fq["R2_10K_TRIM"]], "se": ["-K", "300000", idx, fq["R1_10K"]]}
        for name, args in runs.items():
            sam = subprocess.run([drv, "-t", "8"] + args, capture_output=True, check=True).stdout
            md5[name] = dict(args=[a if not a.startswith(d) else os.path.basename(a) for a in args], md5=hashlib.md5(sam).hexdigest(),
                             lines=sam.count(b"\n"))
            if name == "pe":  # first 3000 pairs' records for the CPU-side host-pipeline test
                pass
        # head subset: first 1500 pairs, single chunk
        for key in ("R1_10K", "R2_10K"):
            with open(fq[key], "rb") as fi:
                lines = fi.read().split(b"\n")[:6000]
            with open(fq[key] + ".head", "wb") as fo:
                fo.write(b"\n".join(lines) + b"\n")
        sam = subprocess.run([drv, "-t", "8", idx, fq["R1_10K"] + ".head", fq["R2_10K"] + ".head"], capture_output=True, check=True).stdout
        md5["pe_head1500"] = dict(args=["idx", "R1_10K[:1500]", "R2_10K[:1500]"], md5=hashlib.md5(sam).hexdigest(), lines=sam.count(b"\n"))
        with gzip.open(os.path.join(HERE, "pe_head1500.sam.gz"), "wb") as fo:
            fo.write(sam)
    with open(os.path.join(HERE, "sam_md5.json"), "w") as fh:
        json.dump(md5, fh, indent=1)
    print(json.dumps(md5, indent=1))


if __name__ == "__main__":
    main()
