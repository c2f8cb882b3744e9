/* TEST INFRASTRUCTURE ONLY - never linked into, imported by or executed from the product path.
 *
 * Serial driver around the UNMODIFIED reference library (oracle/_ref/libbwa_ref.so, compiled from
 * /root/reference/src by oracle/Makefile).  It restates the ~40 host lines of the MPI mains that matter
 * for parity (the mains themselves need <mpi.h>, which this image does not have):
 *   - in-place fastq parse into bseq1_t          (reference src/mainParallel.c:1257-1304)
 *   - chunk rule "close when bases > maxsiz"     (reference src/parallel_aux.c:1532-1549, 1068-1082;
 *                                                 src/mainParallel.c:947,1874,2773)
 *   - n_processed convention                     (reference src/mainParallel.c:1314,2355-2357,3093)
 * and then calls the reference's own mem_process_seqs (src/bwamem.c:1205) on every chunk.
 *
 * usage: ref_driver [-K bases] [-t threads] [-T] [-H] <idxprefix> <r1.fq> [r2.fq]
 *   -T  use the trimmed-pairs rule (R1+R2 bases against K, running n_processed)
 *   -H  print @SQ header lines first
 * Output: SAM records on stdout (no @PG line: it embeds argv in the reference).
 * Timing of the mem_process_seqs calls alone is printed to stderr, per chunk and in total, as
 *   [ref_driver] chunk=<k> reads=<n> sec=<s>
 *   [ref_driver] reads=<n> chunks=<c> mem_process_seqs_sec=<s>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <unistd.h>
#include <time.h>
#include "bwamem.h"
#include "bwa.h"

static char *slurp(const char *fn, size_t *len)
{
	FILE *fp = fopen(fn, "rb");
	char *buf;
	if (!fp) { perror(fn); exit(1); }
	fseek(fp, 0, SEEK_END); *len = ftell(fp); fseek(fp, 0, SEEK_SET);
	buf = malloc(*len + 1);
	if (fread(buf, 1, *len, fp) != *len) { perror("fread"); exit(1); }
	buf[*len] = 0;
	fclose(fp);
	return buf;
}

/* parse a whole fastq buffer in place; returns number of records */
static size_t parse_fastq(char *buf, size_t len, bseq1_t **out)
{
	size_t n = 0, m = 0, line = 0;
	bseq1_t *s = 0;
	char *p = buf, *q = buf, *e = buf + len;
	while (q < e) {
		if (*q != '\n') { ++q; continue; }
		*q = 0;
		switch (line & 3) {
		case 0:
			if (n == m) { m = m ? m << 1 : 1024; s = realloc(s, m * sizeof(bseq1_t)); }
			memset(&s[n], 0, sizeof(bseq1_t));
			s[n].name = p + 1;
			while (*p && !isspace((unsigned char)*p)) ++p;
			if (p - 2 > s[n].name && *(p-2) == '/' && isdigit((unsigned char)*(p-1))) *(p-2) = 0;
			if (*p) *p = 0;
			break;
		case 1: s[n].seq = p; s[n].l_seq = (int)(q - p); break;
		case 2: break;
		case 3: s[n].qual = p; ++n; break;
		}
		p = ++q; ++line;
	}
	*out = s;
	return n;
}


/* -o name=value[,name=value...]: set mem_opt_t fields by name (the scoring matrix is refilled afterwards) */
static void set_opts(mem_opt_t *opt, const char *spec)
{
	char *dup = strdup(spec), *tok, *save = 0;
	for (tok = strtok_r(dup, ",", &save); tok; tok = strtok_r(0, ",", &save)) {
		char *eq = strchr(tok, '=');
		if (!eq) { fprintf(stderr, "bad -o item %s\n", tok); exit(1); }
		*eq = 0;
		const char *v = eq + 1;
#define OPT_I(f) else if (strcmp(tok, #f) == 0) opt->f = atoi(v)
#define OPT_F(f) else if (strcmp(tok, #f) == 0) opt->f = (float)atof(v)
		if (0) {}
		OPT_I(a); OPT_I(b); OPT_I(o_del); OPT_I(e_del); OPT_I(o_ins); OPT_I(e_ins); OPT_I(pen_unpaired); OPT_I(pen_clip5); OPT_I(pen_clip3);
		OPT_I(w); OPT_I(zdrop); OPT_I(T); OPT_I(min_seed_len); OPT_I(min_chain_weight); OPT_I(max_chain_extend); OPT_I(split_width);
		OPT_I(max_occ); OPT_I(max_chain_gap); OPT_I(max_ins); OPT_I(max_matesw); OPT_I(max_XA_hits); OPT_I(max_XA_hits_alt);
		OPT_F(split_factor); OPT_F(mask_level); OPT_F(drop_ratio); OPT_F(XA_drop_ratio);
		else if (strcmp(tok, "flag") == 0) opt->flag |= atoi(v);
		else { fprintf(stderr, "unknown -o field %s\n", tok); exit(1); }
#undef OPT_I
#undef OPT_F
	}
	free(dup);
	bwa_fill_scmat(opt->a, opt->b, opt->mat);
}

/* -I avg,std,high,low: fixed insert-size distribution for the FR orientation, the others marked failed (bwa mem -I) */
static mem_pestat_t *parse_pes(const char *spec, mem_pestat_t pes[4])
{
	double avg, std; int high, low;
	if (sscanf(spec, "%lf,%lf,%d,%d", &avg, &std, &high, &low) != 4) { fprintf(stderr, "bad -I %s\n", spec); exit(1); }
	memset(pes, 0, 4 * sizeof(mem_pestat_t));
	pes[0].failed = pes[2].failed = pes[3].failed = 1;
	pes[1].avg = avg; pes[1].std = std; pes[1].high = high; pes[1].low = low;
	return pes;
}

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char **argv)
{
	int c, trimmed = 0, header = 0, n_threads = 1;
	mem_pestat_t pes_fixed[4], *pes0 = 0;
	long K = 0;
	mem_opt_t *opt = mem_opt_init();
	while ((c = getopt(argc, argv, "K:t:THv:o:I:")) >= 0) {
		if (c == 'K') K = atol(optarg);
		else if (c == 't') n_threads = atoi(optarg);
		else if (c == 'T') trimmed = 1;
		else if (c == 'H') header = 1;
		else if (c == 'v') bwa_verbose = atoi(optarg);
		else if (c == 'o') set_opts(opt, optarg);
		else if (c == 'I') pes0 = parse_pes(optarg, pes_fixed);
	}
	if (argc - optind < 2) { fprintf(stderr, "usage: ref_driver [-K n] [-t n] [-T] [-H] idx r1.fq [r2.fq]\n"); return 1; }
	opt->n_threads = n_threads;
	if (K <= 0) K = (long)opt->chunk_size * n_threads;
	int paired = argc - optind >= 3;
	if (paired) opt->flag |= MEM_F_PE;
	bwaidx_t *idx = bwa_idx_load(argv[optind], BWA_IDX_ALL);
	if (!idx) return 1;
	size_t l1, l2 = 0, n1, n2 = 0, i;
	char *b1 = slurp(argv[optind+1], &l1), *b2 = 0;
	bseq1_t *s1, *s2 = 0;
	n1 = parse_fastq(b1, l1, &s1);
	if (paired) { b2 = slurp(argv[optind+2], &l2); n2 = parse_fastq(b2, l2, &s2); if (n1 != n2) { fprintf(stderr, "unequal read counts\n"); return 1; } }
	if (header)
		for (i = 0; i < (size_t)idx->bns->n_seqs; ++i)
			printf("@SQ\tSN:%s\tLN:%d\n", idx->bns->anns[i].name, idx->bns->anns[i].len);
	long maxsiz = paired && !trimmed ? K / 2 : K;
	size_t beg = 0, n_chunks = 0;
	long bases = 0;
	int64_t n_processed = 0;
	double t_mem = 0;
	bseq1_t *seqs = malloc((paired ? 2 : 1) * n1 * sizeof(bseq1_t));
	for (i = 0; i < n1; ++i) {
		bases += s1[i].l_seq;
		if (paired && trimmed) bases += s2[i].l_seq;
		if (bases > maxsiz || i + 1 == n1) {
			size_t k, n = 0;
			for (k = beg; k <= i; ++k) {
				seqs[n++] = s1[k];
				if (paired) seqs[n++] = s2[k];
			}
			double t0 = now();
			mem_process_seqs(opt, idx->bwt, idx->bns, idx->pac, trimmed ? n_processed : 0, (int)n, seqs, pes0);
			t_mem += now() - t0;
			fprintf(stderr, "[ref_driver] chunk=%zu reads=%zu sec=%.4f\n", n_chunks, n, now() - t0);
			n_processed += n;
			for (k = 0; k < n; ++k) { fputs(seqs[k].sam, stdout); free(seqs[k].sam); }
			beg = i + 1; bases = 0; ++n_chunks;
		}
	}
	fprintf(stderr, "[ref_driver] reads=%zu chunks=%zu mem_process_seqs_sec=%.3f\n", (paired ? 2 : 1) * n1, n_chunks, t_mem);
	return 0;
}
