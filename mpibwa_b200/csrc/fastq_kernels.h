// fastq_kernels.h - the chunk's raw fastq bytes -> the device's read tables, one task per read.
//
// Same result as the in-place parse of the mpiBWA hosts (reference src/mainParallel.c:1257-1304; hostshim.cpp:b200_fastq_parse):
// four lines per record, the name starts after '@', is cut at the first blank and loses a trailing "/[0-9]"; the bases become
// codes 0-4 (reference src/bwamem.c:1057-1058, nst_nt4_table); mates are interleaved 2i, 2i+1 (src/mainParallel.c:1271-1314).
// The bytes are uploaded once and stay in HBM: names and qualities are never copied, the SAM formatter reads them where they lie.
// Host/device code (tests/hostemu loops the same bodies).
#pragma once
#include <cstdint>
#include "finish_kernels.h"

namespace b200 {

B200_HD int fq_isspace(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

B200_HD uint8_t fq_code(uint8_t c)          // nst_nt4_table, reference src/bntseq.c:39-56; bytes below 4 are already codes
{
	if (c < 4) return c;
	switch (c | 0x20) { case 'a': return 0; case 'c': return 1; case 'g': return 2; case 't': return 3; default: return 4; }
}

struct FastqView {
	const char *text;                       // file 1 at [0, len1), file 2 at [len1, len1 + len2)
	const int64_t *nl[2];                   // positions of the line ends of either file (within that file)
	int64_t base[2];                        // where the file starts in text
	int paired;
};

// read r of the interleaved batch: its text table entry, its sequence line and length
B200_HD void fastq_read(const FastqView &v, int64_t r, ReadText *rt, int64_t *seq_at, int32_t *l_seq)
{
	const int f = v.paired ? (int)(r & 1) : 0;
	const int64_t rec = v.paired ? r >> 1 : r;
	const int64_t *nl = v.nl[f];
	const int64_t b = v.base[f];
	const int64_t l0 = rec * 4;
	const int64_t s0 = l0 ? nl[l0 - 1] + 1 : 0, e0 = nl[l0];           // name line
	const int64_t s1 = e0 + 1, e1 = nl[l0 + 1];                         // bases
	const int64_t s3 = nl[l0 + 2] + 1;                                  // qualities
	const char *t = v.text + b;
	int64_t z = s0;
	while (z < e0 && !fq_isspace((unsigned char)t[z])) ++z;
	int64_t name = s0 + 1;
	if (z - 2 > name && t[z - 2] == '/' && t[z - 1] >= '0' && t[z - 1] <= '9') z -= 2;
	rt->name_off = b + name; rt->name_len = (int32_t)(z > name ? z - name : 0);
	rt->qual_off = b + s3; rt->comment_off = -1; rt->comment_len = 0;
	*seq_at = b + s1; *l_seq = (int32_t)(e1 - s1);
}

} // namespace b200
