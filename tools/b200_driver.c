/* b200_driver - minimal single-process stand-in for the mpiBWA chunk loop, driving the B200 alignment core through
 * its C ABI exactly like the MPI hosts do (reference src/mainParallel.c:1146-1493): parse a fastq chunk in place into
 * bseq1_t (src/mainParallel.c:1257-1304), call mem_process_seqs, write seqs[i].sam in input order.
 * Chunks follow the reference rule "close when bases > maxsiz" (src/parallel_aux.c:1532-1549; maxsiz = K/2 of R1
 * for same-size pairs, K of R1+R2 for trimmed pairs with a running n_processed, K for single-end).
 * With -r RANK -n NRANKS chunk c is taken by rank c % NRANKS (the MPI hosts claim chunks from a shared counter;
 * a static round-robin keeps runs reproducible) and only that rank's records are written.
 *
 * With -P the loop keeps two chunks in flight through the chunk-job form of the call (b200_process_seqs_begin / _end):
 * "read chunk i+1; begin(i+1); end(i); write chunk i" - the patched host loop of INTEGRATION.md.
 *
 * With -C each chunk goes through b200_align_chunk (mates interleaved by the library, SAM returned as one buffer).
 *
 * usage: b200_driver [-K bases] [-t threads] [-T] [-H] [-P|-C] [-r rank -n nranks] [-d device] <idxprefix|.map> <r1.fq> [r2.fq]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <unistd.h>
#include <time.h>
#include "mpibwa_b200.h"

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
	int c, trimmed = 0, header = 0, n_threads = 1, rank = 0, nranks = 1, device = 0, pipelined = 0, route = 0;
	mem_pestat_t pes_fixed[4], *pes0 = 0;
	long K = 0;
	mem_opt_t *opt = mem_opt_init();
	while ((c = getopt(argc, argv, "K:t:THPCFv:r:n:d:o:I:R:")) >= 0) {
		if (c == 'K') K = atol(optarg);
		else if (c == 't') n_threads = atoi(optarg);
		else if (c == 'T') trimmed = 1;
		else if (c == 'H') header = 1;
		else if (c == 'P') pipelined = 1;
		else if (c == 'C') pipelined = 2;     /* b200_align_chunk: the library interleaves the mates and returns one SAM buffer */
		else if (c == 'F') pipelined = 3;     /* b200_align_fastq_begin: the chunk goes in as raw fastq bytes (parsed on the device) */
		else if (c == 'v') bwa_verbose = atoi(optarg);
		else if (c == 'r') rank = atoi(optarg);
		else if (c == 'n') nranks = atoi(optarg);
		else if (c == 'd') device = atoi(optarg);
		else if (c == 'o') set_opts(opt, optarg);
		else if (c == 'I') pes0 = parse_pes(optarg, pes_fixed);
		else if (c == 'R') route = atoi(optarg);   /* with -F: B200_ROUTE_* flags (per-chromosome routing of the chunk's lines) */
	}
	if (argc - optind < 2) { fprintf(stderr, "usage: b200_driver [-K n] [-t n] [-T] [-H] [-r rank -n nranks] [-d dev] idx r1.fq [r2.fq]\n"); return 1; }
	opt->n_threads = n_threads;
	if (K <= 0) K = (long)opt->chunk_size * n_threads;
	int paired = argc - optind >= 3;
	if (paired) opt->flag |= MEM_F_PE;
	bwaidx_t *idx;
	size_t la = strlen(argv[optind]);
	if (la > 4 && strcmp(argv[optind] + la - 4, ".map") == 0) {
		size_t l_mem;
		uint8_t *mem = (uint8_t *)slurp(argv[optind], &l_mem);
		idx = calloc(1, sizeof(bwaidx_t));
		bwa_mem2idx((int64_t)l_mem, mem, idx);
	} else idx = bwa_idx_load(argv[optind], BWA_IDX_ALL);
	if (!idx) return 1;
	b200_gpu_init(idx, device);
	size_t l1, l2 = 0, n1, n2 = 0, i;
	char *b1 = slurp(argv[optind+1], &l1), *b2 = 0, *raw1 = 0, *raw2 = 0;
	bseq1_t *s1, *s2 = 0;
	if (pipelined == 3) { raw1 = malloc(l1 + 1); memcpy(raw1, b1, l1 + 1); }       /* the parse below is in place: keep the bytes */
	n1 = parse_fastq(b1, l1, &s1);
	if (paired) {
		b2 = slurp(argv[optind+2], &l2);
		if (pipelined == 3) { raw2 = malloc(l2 + 1); memcpy(raw2, b2, l2 + 1); }
		n2 = parse_fastq(b2, l2, &s2);
		if (n1 != n2) { fprintf(stderr, "unequal read counts\n"); return 1; }
	}
	if (header && rank == 0)
		for (i = 0; i < (size_t)idx->bns->n_seqs; ++i)
			printf("@SQ\tSN:%s\tLN:%d\n", idx->bns->anns[i].name, idx->bns->anns[i].len);
	long maxsiz = paired && !trimmed ? K / 2 : K;
	size_t beg = 0, n_chunks = 0, mine = 0;
	long bases = 0;
	int64_t n_processed = 0;
	double t_mem = 0;
	bseq1_t *seqs = malloc((paired ? 2 : 1) * n1 * sizeof(bseq1_t));
	b200_job_t *prev_job = 0;                 /* -P: the chunk whose end() is still to come */
	bseq1_t *prev_seqs = 0;
	size_t prev_n = 0;
	for (i = 0; i < n1; ++i) {
		bases += s1[i].l_seq;
		if (paired && trimmed) bases += s2[i].l_seq;
		if (bases > maxsiz || i + 1 == n1) {
			if ((int)(n_chunks % nranks) == rank) {
				size_t k, n = 0;
				for (k = beg; k <= i; ++k) {
					seqs[n++] = s1[k];
					if (paired) seqs[n++] = s2[k];
				}
				double t0 = now();
				if (pipelined == 3) {
					/* byte range of the chunk's records in either file: from the '@' of its first record to the '@' of the next chunk's */
					size_t o1 = (size_t)(s1[beg].name - 1 - b1), e1 = i + 1 < n1 ? (size_t)(s1[i + 1].name - 1 - b1) : l1;
					size_t o2 = paired ? (size_t)(s2[beg].name - 1 - b2) : 0, e2 = paired ? (i + 1 < n1 ? (size_t)(s2[i + 1].name - 1 - b2) : l2) : 0;
					char *sam = 0; int64_t sam_len = 0;
					b200_set_routing(route);
					b200_job_t *job = b200_align_fastq_begin(opt, idx, trimmed ? n_processed : 0, raw1 + o1, (int64_t)(e1 - o1), paired ? raw2 + o2 : 0, (int64_t)(e2 - o2));
					if (!route) {
						n = (size_t)b200_align_chunk_end(job, &sam, &sam_len, 0);
						fwrite(sam, 1, (size_t)sam_len, stdout);
					} else {
						/* routed chunk: "@@CHUNK", then either the destinations' ranges ("@@DEST d bytes" + the bytes) or the text in
						 * input order followed by its line table ("@@LINE off len rid mate_rid read") */
						b200_sam_line_t *lines = 0; int64_t n_lines = 0, *dest_off = 0; int n_dest = 0, d;
						n = (size_t)b200_align_chunk_end_routed(job, &sam, &sam_len, &lines, &n_lines, &dest_off, &n_dest, 0);
						printf("@@CHUNK\t%lld\n", (long long)sam_len);
						if (route & B200_ROUTE_BY_CONTIG) {
							for (d = 0; d < n_dest; ++d)
								if (dest_off[d + 1] > dest_off[d]) {
									printf("@@DEST\t%d\t%lld\n", d, (long long)(dest_off[d + 1] - dest_off[d]));
									fwrite(sam + dest_off[d], 1, (size_t)(dest_off[d + 1] - dest_off[d]), stdout);
								}
						} else {
							int64_t l;
							fwrite(sam, 1, (size_t)sam_len, stdout);
							for (l = 0; l < n_lines; ++l)
								printf("@@LINE\t%lld\t%d\t%d\t%d\t%d\n", (long long)lines[l].off, lines[l].len, lines[l].rid, lines[l].mate_rid, lines[l].read);
						}
						free(lines); free(dest_off);
					}
					b200_free(sam);
				} else if (pipelined == 2) {
					char *sam = 0; int64_t sam_len = 0;
					n = (size_t)b200_align_chunk(opt, idx, trimmed ? n_processed : 0, (int64_t)(i - beg + 1), s1 + beg, paired ? s2 + beg : 0, &sam, &sam_len);
					fwrite(sam, 1, (size_t)sam_len, stdout);
					b200_free(sam);
				} else if (pipelined) {
					bseq1_t *cs = malloc(n * sizeof(bseq1_t));
					memcpy(cs, seqs, n * sizeof(bseq1_t));
					b200_job_t *job = b200_process_seqs_begin(opt, idx->bwt, idx->bns, idx->pac, trimmed ? n_processed : 0, (int)n, cs, pes0);
					if (prev_job) {
						b200_process_seqs_end(prev_job, 0);
						for (k = 0; k < prev_n; ++k) { fputs(prev_seqs[k].sam, stdout); free(prev_seqs[k].sam); }
						free(prev_seqs);
					}
					prev_job = job; prev_seqs = cs; prev_n = n;
				} else {
					mem_process_seqs(opt, idx->bwt, idx->bns, idx->pac, trimmed ? n_processed : 0, (int)n, seqs, pes0);
					for (k = 0; k < n; ++k) { fputs(seqs[k].sam, stdout); free(seqs[k].sam); }
				}
				t_mem += now() - t0;
				n_processed += n;
				mine += n;
			}
			beg = i + 1; bases = 0; ++n_chunks;
		}
	}
	if (prev_job) {
		size_t k;
		double t0 = now();
		b200_process_seqs_end(prev_job, 0);
		t_mem += now() - t0;
		for (k = 0; k < prev_n; ++k) { fputs(prev_seqs[k].sam, stdout); free(prev_seqs[k].sam); }
		free(prev_seqs);
	}
	fprintf(stderr, "[b200_driver] rank=%d reads=%zu chunks=%zu mem_process_seqs_sec=%.3f\n", rank, mine, n_chunks, t_mem);
	b200_gpu_release();
	return 0;
}
