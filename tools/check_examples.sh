#!/bin/bash
# Byte-compare the SAM of the B200 core with the compiled reference (oracle/_ref) on the bundled example data.
# usage: tools/check_examples.sh [driver]   (default tools/b200_driver; tests pass tests/_build/b200_driver_hostemu)
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DRV=$(readlink -f "${1:-$ROOT/tools/b200_driver}")
REF=$ROOT/oracle/_ref/ref_driver
W=${TMPDIR:-/tmp}/b200_examples.$$
mkdir -p $W && cd $W
tar xzf $ROOT/tests/golden/examples/hg19.small.tar.gz
IDX=$(ls $W/*.fa $W/*/*.fa 2>/dev/null | head -1)
for f in R1_10K R2_10K R1_10K_TRIM R2_10K_TRIM; do gzip -dc $ROOT/tests/golden/examples/HCC1187C_$f.fastq.gz > $f.fq; done
fail=0
run() { # name, args...
	name=$1; shift
	$REF -t 8 "$@" > $name.ref.sam 2> $name.ref.log || { echo "REF FAILED $name"; fail=1; }
	$DRV -t 8 "$@" > $name.b200.sam 2> $name.b200.log || { echo "B200 FAILED $name"; tail -5 $name.b200.log; fail=1; }
	if cmp -s $name.ref.sam $name.b200.sam; then echo "OK   $name $(wc -l < $name.ref.sam) lines $(grep -h 'mem_process_seqs_sec' $name.b200.log | tail -1)";
	else echo "DIFF $name"; diff $name.ref.sam $name.b200.sam | head -6; fail=1; fi
}
run pe      -H $IDX R1_10K.fq R2_10K.fq
run pe_K    -K 500000 $IDX R1_10K.fq R2_10K.fq
run trim    -T -K 700000 $IDX R1_10K_TRIM.fq R2_10K_TRIM.fq
run se      -K 300000 $IDX R1_10K.fq
grep -h "Processed" pe.b200.log | tail -2
rm -rf $W
exit $fail
