// hostshim.cpp - the few host-side lines of the mpiBWA mains that sit directly around mem_process_seqs, exported
// through the C ABI so that test and bench harnesses (Python/ctypes, tools/b200_driver.c) drive the alignment core
// exactly like the MPI hosts do:
//   b200_fastq_parse    in-place fastq parse into bseq1_t      reference src/mainParallel.c:1257-1304
//   b200_plan_chunks    "close the chunk when bases > maxsiz"  reference src/parallel_aux.c:1532-1549 (same-size
//                       pairs: R1 bases vs K/2), :1068-1082 (trimmed pairs: R1+R2 vs K), src/mainParallel.c:2773 (SE)
//   b200_align_chunk    interleave mates, mem_process_seqs, concatenate seqs[i].sam in input order
//                       reference src/mainParallel.c:1271-1314 and copy_buffer_thr :103-127
#include "../../include/mpibwa_b200.h"
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {

int64_t b200_fastq_parse(char *buf, int64_t len, bseq1_t **out)
{
	size_t n = 0, m = 0, line = 0;
	bseq1_t *s = nullptr;
	char *p = buf, *q = buf, *e = buf + len;
	while (q < e) {
		if (*q != '\n') { ++q; continue; }
		*q = 0;
		switch (line & 3) {
		case 0: {
			if (n == m) { m = m ? m << 1 : 1024; s = (bseq1_t *)realloc(s, m * sizeof(bseq1_t)); }
			memset(&s[n], 0, sizeof(bseq1_t));
			s[n].name = p + 1;
			char *t = p;
			while (*t && !isspace((unsigned char)*t)) ++t;
			if (t - 2 > s[n].name && *(t - 2) == '/' && isdigit((unsigned char)*(t - 1))) *(t - 2) = 0;
			if (*t) *t = 0;
			break;
		}
		case 1: s[n].seq = p; s[n].l_seq = (int)(q - p); break;
		case 2: break;
		case 3: s[n].qual = p; ++n; break;
		}
		p = ++q; ++line;
	}
	*out = s;
	return (int64_t)n;
}

int64_t b200_plan_chunks(int64_t n, const bseq1_t *s1, const bseq1_t *s2, int64_t K, int trimmed, int64_t **ends)
{
	std::vector<int64_t> e;
	const int64_t maxsiz = (s2 && !trimmed) ? K / 2 : K;
	int64_t bases = 0;
	for (int64_t i = 0; i < n; ++i) {
		bases += s1[i].l_seq;
		if (s2 && trimmed) bases += s2[i].l_seq;
		if (bases > maxsiz || i + 1 == n) { e.push_back(i + 1); bases = 0; }
	}
	*ends = (int64_t *)malloc((e.size() + 1) * sizeof(int64_t));
	memcpy(*ends, e.data(), e.size() * sizeof(int64_t));
	return (int64_t)e.size();
}

bseq1_t *b200_chunk_seqs(int64_t n, const bseq1_t *s1, const bseq1_t *s2)
{
	const int64_t total = s2 ? 2 * n : n;
	bseq1_t *seqs = (bseq1_t *)malloc((size_t)(total + 1) * sizeof(bseq1_t));
	for (int64_t i = 0; i < n; ++i) {
		if (s2) { seqs[2 * i] = s1[i]; seqs[2 * i + 1] = s2[i]; }
		else seqs[i] = s1[i];
	}
	return seqs;
}

int64_t b200_collect_sam(int64_t total, bseq1_t *seqs, char **sam)
{
	size_t sum = 0;
	std::vector<size_t> len((size_t)total);
	for (int64_t i = 0; i < total; ++i) { len[i] = seqs[i].sam ? strlen(seqs[i].sam) : 0; sum += len[i]; }
	char *buf = sam ? (char *)malloc(sum + 1) : nullptr, *w = buf;
	for (int64_t i = 0; i < total; ++i) {
		if (buf) { memcpy(w, seqs[i].sam, len[i]); w += len[i]; }
		free(seqs[i].sam);
		seqs[i].sam = nullptr;
	}
	if (buf) { *w = 0; *sam = buf; }
	return (int64_t)sum;
}

int64_t b200_align_chunk(const mem_opt_t *opt, const bwaidx_t *idx, int64_t n_processed, int64_t n, bseq1_t *s1, bseq1_t *s2,
                         char **sam, int64_t *sam_len)
{
	const int64_t total = s2 ? 2 * n : n;
	bseq1_t *seqs = b200_chunk_seqs(n, s1, s2);
	mem_process_seqs(opt, idx->bwt, idx->bns, idx->pac, n_processed, (int)total, seqs, nullptr);
	int64_t l = b200_collect_sam(total, seqs, sam);
	if (sam_len) *sam_len = l;
	free(seqs);
	return total;
}

void b200_free(void *p) { free(p); }

} // extern "C"
