// capi.cpp - the extern "C" surface declared in include/mpibwa_b200.h: option/header helpers, index loading and
// (de)serialisation in the reference's on-disk and `.map` formats, and the thin wrappers that route the
// reference's call surface into the batched device stages.
#include "host_align.h"
#include "stages.h"
#include "util.h"
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>

namespace b200 {
Engine *engine_for(const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac);
Engine *engine_current();
Engine *aux_acquire();
void aux_release();
void engine_select_device(int dev);
void engine_release();
ExtOpt make_ext_opt(const mem_opt_t *opt);
SwOpt make_sw_opt(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins);
SeedOpt make_seed_opt(const mem_opt_t *opt);
void process_seqs(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                  int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0);
void stage_reads(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int n, bseq1_t *seqs);
struct SeqJob;
SeqJob *process_seqs_begin(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int64_t n_processed, int n,
                           bseq1_t *seqs, const mem_pestat_t *pes0, void (*after)(void *, SeqJob *), void *arg, bool one_buffer,
                           char *fq1, int64_t len1, char *fq2, int64_t len2, int route);
int64_t job_n_reads(SeqJob *j);
int64_t job_take_lines(SeqJob *j, b200_sam_line_t **out);
int job_take_dest_off(SeqJob *j, int64_t **out);
int64_t job_take_sam(SeqJob *j, char **out);
void process_seqs_end(SeqJob *j, b200_stats_t *stats);
void last_stats(b200_stats_t *out);

static void die(const char *what, const char *arg)
{
	fprintf(stderr, "[mpibwa_b200] %s%s%s\n", what, arg ? ": " : "", arg ? arg : "");
	abort();
}

// The single-job / caller-batch entry points run on an engine of their own, one call at a time, and take the device turn like a
// chunk job's stage does: they are safe to call from several threads and while b200_process_seqs_begin jobs are in flight.
struct AuxGuard {
	Engine *e;
	AuxGuard() : e(aux_acquire()) { if (!e) die("no index on the device: call b200_gpu_init() or mem_process_seqs() first", nullptr); }
	~AuxGuard() { aux_release(); }
	AuxGuard(const AuxGuard &) = delete;
	operator Engine *() const { return e; }
};
} // namespace b200

using namespace b200;

extern "C" {

int bwa_verbose = 3;
char bwa_rg_id[256];
char *bwa_pg = 0;

const char *b200_version(void) { return "mpibwa_b200 0.1 (BWA-MEM 0.7.17 semantics, mpiBWA 1.5.5 call surface)"; }

/* ------------------------------------------------------------------ options and header helpers */

void bwa_fill_scmat(int a, int b, int8_t mat[25])
{
	int k = 0;
	for (int i = 0; i < 4; ++i) {
		for (int j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? a : -b);
		mat[k++] = -1;
	}
	for (int j = 0; j < 5; ++j) mat[k++] = -1;
}

mem_opt_t *mem_opt_init(void)
{
	mem_opt_t *o = (mem_opt_t *)calloc(1, sizeof(mem_opt_t));
	o->flag = 0;
	o->a = 1; o->b = 4;
	o->o_del = o->o_ins = 6;
	o->e_del = o->e_ins = 1;
	o->w = 100;
	o->T = 30;
	o->zdrop = 100;
	o->pen_unpaired = 17;
	o->pen_clip5 = o->pen_clip3 = 5;
	o->max_mem_intv = 20;
	o->min_seed_len = 19;
	o->split_width = 10;
	o->max_occ = 500;
	o->max_chain_gap = 10000;
	o->max_ins = 10000;
	o->mask_level = 0.50;
	o->drop_ratio = 0.50;
	o->XA_drop_ratio = 0.80;
	o->split_factor = 1.5;
	o->chunk_size = 10000000;
	o->n_threads = 1;
	o->max_XA_hits = 5;
	o->max_XA_hits_alt = 200;
	o->max_matesw = 50;
	o->mask_level_redun = 0.95;
	o->min_chain_weight = 0;
	o->max_chain_extend = 1 << 30;
	o->mapQ_coef_len = 50; o->mapQ_coef_fac = log(o->mapQ_coef_len);
	bwa_fill_scmat(o->a, o->b, o->mat);
	return o;
}

static char *unescape_in_place(char *s)
{
	char *p, *q;
	for (p = q = s; *p; ++p) {
		if (*p == '\\') {
			++p;
			if (*p == 't') *q++ = '\t';
			else if (*p == 'n') *q++ = '\n';
			else if (*p == 'r') *q++ = '\r';
			else if (*p == '\\') *q++ = '\\';
		} else *q++ = *p;
	}
	*q = 0;
	return s;
}

char *bwa_set_rg(const char *s)
{
	char *p, *q, *r, *rg_line = 0;
	memset(bwa_rg_id, 0, 256);
	if (strstr(s, "@RG") != s) {
		if (bwa_verbose >= 1) fprintf(stderr, "[E::%s] the read group line is not started with @RG\n", __func__);
		return 0;
	}
	if (strstr(s, "\t") != NULL) {
		if (bwa_verbose >= 1) fprintf(stderr, "[E::%s] the read group line contained literal <tab> characters -- replace with escaped tabs: \\t\n", __func__);
		return 0;
	}
	rg_line = strdup(s);
	unescape_in_place(rg_line);
	if ((p = strstr(rg_line, "\tID:")) == 0) {
		if (bwa_verbose >= 1) fprintf(stderr, "[E::%s] no ID within the read group line\n", __func__);
		free(rg_line);
		return 0;
	}
	p += 4;
	for (q = p; *q && *q != '\t' && *q != '\n'; ++q) {}
	if (q - p + 1 > 256) {
		if (bwa_verbose >= 1) fprintf(stderr, "[E::%s] @RG:ID is longer than 255 characters\n", __func__);
		free(rg_line);
		return 0;
	}
	for (q = p, r = bwa_rg_id; *q && *q != '\t' && *q != '\n'; ++q) *r++ = *q;
	return rg_line;
}

char *bwa_insert_header(const char *s, char *hdr)
{
	int len = 0;
	if (s == 0 || s[0] != '@') return hdr;
	if (hdr) {
		len = (int)strlen(hdr);
		hdr = (char *)realloc(hdr, len + strlen(s) + 2);
		hdr[len++] = '\n';
		strcpy(hdr + len, s);
	} else hdr = strdup(s);
	unescape_in_place(hdr + len);
	return hdr;
}

/* ------------------------------------------------------------------ index files (.bwt .sa .ann .amb .pac) and the .map image */

static FILE *must_open(const std::string &fn, const char *mode)
{
	FILE *fp = fopen(fn.c_str(), mode);
	if (!fp) die("cannot open", fn.c_str());
	return fp;
}
static void must_read(void *dst, size_t sz, size_t n, FILE *fp, const std::string &fn)
{
	const size_t blk = (size_t)1 << 26;
	size_t total = sz * n, done = 0;
	while (done < total) {
		size_t want = total - done < blk ? total - done : blk;
		if (fread((char *)dst + done, 1, want, fp) != want) die("short read", fn.c_str());
		done += want;
	}
}

static bwt_t *load_bwt(const std::string &prefix)
{
	std::string fn = prefix + ".bwt";
	FILE *fp = must_open(fn, "rb");
	bwt_t *bwt = (bwt_t *)calloc(1, sizeof(bwt_t));
	fseek(fp, 0, SEEK_END);
	bwt->bwt_size = ((bwtint_t)ftell(fp) - sizeof(bwtint_t) * 5) >> 2;
	bwt->bwt = (uint32_t *)calloc(bwt->bwt_size, 4);
	fseek(fp, 0, SEEK_SET);
	must_read(&bwt->primary, sizeof(bwtint_t), 1, fp, fn);
	must_read(bwt->L2 + 1, sizeof(bwtint_t), 4, fp, fn);
	must_read(bwt->bwt, 4, bwt->bwt_size, fp, fn);
	bwt->seq_len = bwt->L2[4];
	fclose(fp);
	for (int i = 0; i != 256; ++i) {          // byte -> per-symbol counts (reference src/bwt.c:42-51); kept for ABI
		uint32_t x = 0;
		for (int j = 0; j != 4; ++j)
			x |= (((i & 3) == j) + ((i >> 2 & 3) == j) + ((i >> 4 & 3) == j) + (i >> 6 == j)) << (j << 3);
		bwt->cnt_table[i] = x;
	}
	fn = prefix + ".sa";
	fp = must_open(fn, "rb");
	bwtint_t primary, skipped[4], seq_len, sa_intv;
	must_read(&primary, sizeof(bwtint_t), 1, fp, fn);
	if (primary != bwt->primary) die("SA-BWT inconsistency: primary is not the same", fn.c_str());
	must_read(skipped, sizeof(bwtint_t), 4, fp, fn);
	must_read(&sa_intv, sizeof(bwtint_t), 1, fp, fn);
	must_read(&seq_len, sizeof(bwtint_t), 1, fp, fn);
	if (seq_len != bwt->seq_len) die("SA-BWT inconsistency: seq_len is not the same", fn.c_str());
	bwt->sa_intv = (int)sa_intv;
	bwt->n_sa = (bwt->seq_len + bwt->sa_intv) / bwt->sa_intv;
	bwt->sa = (bwtint_t *)calloc(bwt->n_sa, sizeof(bwtint_t));
	bwt->sa[0] = (bwtint_t)-1;
	must_read(bwt->sa + 1, sizeof(bwtint_t), bwt->n_sa - 1, fp, fn);
	fclose(fp);
	return bwt;
}

// .ann / .amb text (reference src/bntseq.c:100-166)
static bntseq_t *load_bns(const std::string &prefix)
{
	bntseq_t *bns = (bntseq_t *)calloc(1, sizeof(bntseq_t));
	char buf[8192];
	long long xx;
	std::string fn = prefix + ".ann";
	FILE *fp = must_open(fn, "r");
	if (fscanf(fp, "%lld%d%u", &xx, &bns->n_seqs, &bns->seed) != 3) die("malformed header", fn.c_str());
	bns->l_pac = xx;
	bns->anns = (bntann1_t *)calloc(bns->n_seqs, sizeof(bntann1_t));
	for (int i = 0; i < bns->n_seqs; ++i) {
		bntann1_t *p = bns->anns + i;
		int c, n_read;
		if (fscanf(fp, "%u%8191s", &p->gi, buf) != 2) die("malformed record", fn.c_str());
		p->name = strdup(buf);
		std::string anno;
		while ((c = fgetc(fp)) != '\n' && c != EOF) anno.push_back((char)c);
		if (anno.size() > 1 && anno != " (null)") p->anno = strdup(anno.c_str() + 1);   // skip the leading space (reference src/bntseq.c:131-132)
		else p->anno = strdup("");
		n_read = fscanf(fp, "%lld%d%d", &xx, &p->len, &p->n_ambs);
		if (n_read != 3) die("malformed record", fn.c_str());
		p->offset = xx;
	}
	fclose(fp);
	fn = prefix + ".amb";
	fp = must_open(fn, "r");
	int32_t n_seqs;
	if (fscanf(fp, "%lld%d%d", &xx, &n_seqs, &bns->n_holes) != 3) die("malformed header", fn.c_str());
	if (xx != bns->l_pac || n_seqs != bns->n_seqs) die("inconsistent .ann and .amb files", fn.c_str());
	bns->ambs = bns->n_holes ? (bntamb1_t *)calloc(bns->n_holes, sizeof(bntamb1_t)) : 0;
	for (int i = 0; i < bns->n_holes; ++i) {
		bntamb1_t *p = bns->ambs + i;
		char c[2];
		if (fscanf(fp, "%lld%d%1s", &xx, &p->len, c) != 3) die("malformed record", fn.c_str());
		p->offset = xx;
		p->amb = c[0];
	}
	fclose(fp);
	// ALT contigs: names listed in <prefix>.alt (reference src/bntseq.c:169-197)
	fn = prefix + ".alt";
	if ((fp = fopen(fn.c_str(), "r")) != 0) {
		while (fscanf(fp, "%8191s", buf) == 1) {
			if (buf[0] != '@')
				for (int i = 0; i < bns->n_seqs; ++i)
					if (strcmp(bns->anns[i].name, buf) == 0) { bns->anns[i].is_alt = 1; break; }
			int c;
			while ((c = fgetc(fp)) != '\n' && c != EOF) {}
		}
		fclose(fp);
	}
	return bns;
}

bwaidx_t *bwa_idx_load(const char *hint, int which)
{
	std::string prefix(hint);
	FILE *fp;
	if ((fp = fopen((prefix + ".64.bwt").c_str(), "rb")) != 0) { fclose(fp); prefix += ".64"; }
	else if ((fp = fopen((prefix + ".bwt").c_str(), "rb")) != 0) fclose(fp);
	else {
		if (bwa_verbose >= 1) fprintf(stderr, "[E::%s] fail to locate the index files\n", __func__);
		return 0;
	}
	bwaidx_t *idx = (bwaidx_t *)calloc(1, sizeof(bwaidx_t));
	if (which & BWA_IDX_BWT) idx->bwt = load_bwt(prefix);
	if (which & BWA_IDX_BNS) {
		idx->bns = load_bns(prefix);
		if (which & BWA_IDX_PAC) {
			std::string fn = prefix + ".pac";
			fp = must_open(fn, "rb");
			idx->pac = (uint8_t *)calloc(idx->bns->l_pac / 4 + 1, 1);
			must_read(idx->pac, 1, idx->bns->l_pac / 4 + 1, fp, fn);
			fclose(fp);
		}
	}
	return idx;
}

void bwa_idx_destroy(bwaidx_t *idx)
{
	if (idx == 0) return;
	if (idx->mem == 0) {
		if (idx->bwt) { free(idx->bwt->sa); free(idx->bwt->bwt); free(idx->bwt); }
		if (idx->bns) {
			for (int i = 0; i < idx->bns->n_seqs; ++i) { free(idx->bns->anns[i].name); free(idx->bns->anns[i].anno); }
			free(idx->bns->anns); free(idx->bns->ambs); free(idx->bns);
		}
		free(idx->pac);
	} else if (!idx->is_shm) free(idx->mem);   // bwt/bns/pac live inside the image
	free(idx);
}

// `.map` image: bwt_t | bwt[] | sa[] | bntseq_t | ambs[] | anns[] | (name\0 anno\0)* | pac[]   
int bwa_mem2idx(int64_t l_mem, uint8_t *mem, bwaidx_t *idx)
{
	int64_t k = 0;
	idx->bwt = (bwt_t *)mem; k += sizeof(bwt_t);
	idx->bwt->bwt = (uint32_t *)(mem + k); k += idx->bwt->bwt_size * 4;
	idx->bwt->sa = (bwtint_t *)(mem + k); k += idx->bwt->n_sa * sizeof(bwtint_t);
	idx->bns = (bntseq_t *)(mem + k); k += sizeof(bntseq_t);
	idx->bns->ambs = (bntamb1_t *)(mem + k); k += idx->bns->n_holes * sizeof(bntamb1_t);
	idx->bns->anns = (bntann1_t *)(mem + k); k += idx->bns->n_seqs * sizeof(bntann1_t);
	for (int i = 0; i < idx->bns->n_seqs; ++i) {
		idx->bns->anns[i].name = (char *)(mem + k); k += strlen(idx->bns->anns[i].name) + 1;
		idx->bns->anns[i].anno = (char *)(mem + k); k += strlen(idx->bns->anns[i].anno) + 1;
	}
	idx->pac = (uint8_t *)(mem + k); k += idx->bns->l_pac / 4 + 1;
	if (k != l_mem) die("bwa_mem2idx: image length mismatch", nullptr);
	idx->l_mem = k; idx->mem = mem;
	return 0;
}

int bwa_idx2mem(bwaidx_t *idx)
{
	const bwt_t *bwt = idx->bwt;
	const bntseq_t *bns = idx->bns;
	int64_t total = sizeof(bwt_t) + bwt->bwt_size * 4 + bwt->n_sa * sizeof(bwtint_t) + sizeof(bntseq_t)
	              + bns->n_holes * sizeof(bntamb1_t) + bns->n_seqs * sizeof(bntann1_t) + bns->l_pac / 4 + 1;
	for (int i = 0; i < bns->n_seqs; ++i) total += strlen(bns->anns[i].name) + strlen(bns->anns[i].anno) + 2;
	uint8_t *mem = (uint8_t *)malloc(total);
	int64_t k = 0;
	auto put = [&](const void *p, int64_t n) { memcpy(mem + k, p, n); k += n; };
	put(bwt, sizeof(bwt_t));
	put(bwt->bwt, bwt->bwt_size * 4);
	put(bwt->sa, bwt->n_sa * sizeof(bwtint_t));
	put(bns, sizeof(bntseq_t));
	put(bns->ambs, bns->n_holes * sizeof(bntamb1_t));
	put(bns->anns, bns->n_seqs * sizeof(bntann1_t));
	for (int i = 0; i < bns->n_seqs; ++i) {
		put(bns->anns[i].name, strlen(bns->anns[i].name) + 1);
		put(bns->anns[i].anno, strlen(bns->anns[i].anno) + 1);
	}
	put(idx->pac, bns->l_pac / 4 + 1);
	// release the separately allocated pieces, then re-point into the image
	free(idx->bwt->sa); free(idx->bwt->bwt); free(idx->bwt);
	for (int i = 0; i < idx->bns->n_seqs; ++i) { free(idx->bns->anns[i].name); free(idx->bns->anns[i].anno); }
	free(idx->bns->anns); free(idx->bns->ambs); free(idx->bns);
	free(idx->pac);
	idx->bwt = 0; idx->bns = 0; idx->pac = 0;
	return bwa_mem2idx(k, mem, idx);
}

/* ------------------------------------------------------------------ hot path and B200 additions */

void mem_process_seqs(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                      int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0)
{
	process_seqs(opt, bwt, bns, pac, n_processed, n, seqs, pes0);
}

int b200_device_count(void) { return engine_device_count(); }

int b200_gpu_init(const bwaidx_t *idx, int device)
{
	engine_select_device(device);
	engine_for(idx->bwt, idx->bns, idx->pac);
	return 0;
}

void b200_gpu_release(void) { engine_release(); }

void b200_stage_reads(const mem_opt_t *opt, const bwaidx_t *idx, int n, bseq1_t *seqs)
{
	stage_reads(opt, idx->bwt, idx->bns, idx->pac, n, seqs);
}

/* chunk jobs: the asynchronous form of mem_process_seqs / b200_align_chunk (see include/mpibwa_b200.h) */
struct b200_job {
	SeqJob *job;
	bseq1_t *seqs; int64_t total;       // b200_align_chunk_begin: the interleaved mates and the SAM collected by the job thread
	char *sam; int64_t sam_len;
	b200_sam_line_t *lines = nullptr; int64_t n_lines = 0; int64_t *dest_off = nullptr; int n_dest = 0;   // b200_set_routing products
	int n_threads;
	char *fq[2]; int64_t fq_len[2];     // b200_align_fastq_begin: the raw fastq buffers, parsed in place by the job thread
	bseq1_t *mate[2];
};

static std::atomic<int> g_route{0};
void b200_set_routing(int flags) { g_route = flags; }
static void take_routing(b200_job *x, SeqJob *self)
{
	x->n_lines = job_take_lines(self, &x->lines);
	x->n_dest = job_take_dest_off(self, &x->dest_off);
}

b200_job_t *b200_process_seqs_begin(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                                    int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0)
{
	b200_job *j = new b200_job();
	j->seqs = nullptr; j->total = 0; j->sam = nullptr; j->sam_len = 0;
	j->job = process_seqs_begin(opt, bwt, bns, pac, n_processed, n, seqs, pes0, nullptr, nullptr, false, nullptr, 0, nullptr, 0, 0);
	return j;
}

void b200_process_seqs_end(b200_job_t *j, b200_stats_t *stats)
{
	process_seqs_end(j->job, stats);
	delete j;
}

b200_job_t *b200_align_chunk_begin(const mem_opt_t *opt, const bwaidx_t *idx, int64_t n_processed, int64_t n, bseq1_t *s1, bseq1_t *s2)
{
	b200_job *j = new b200_job();
	j->total = s2 ? 2 * n : n;
	j->seqs = b200_chunk_seqs(n, s1, s2);
	j->sam = nullptr; j->sam_len = 0;
	j->n_threads = opt->n_threads;
	// the chunk's text comes back from the device as one buffer (no malloc per read)
	j->job = process_seqs_begin(opt, idx->bwt, idx->bns, idx->pac, n_processed, (int)j->total, j->seqs, nullptr,
		[](void *p, SeqJob *self) { b200_job *x = (b200_job *)p; x->sam_len = job_take_sam(self, &x->sam); take_routing(x, self); free(x->seqs); x->seqs = nullptr; }, j, true, nullptr, 0, nullptr, 0, g_route);
	return j;
}

b200_job_t *b200_align_seqs_begin(const mem_opt_t *opt, const bwaidx_t *idx, int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0)
{
	b200_job *j = new b200_job();
	j->total = n; j->seqs = nullptr; j->sam = nullptr; j->sam_len = 0; j->n_threads = opt->n_threads;
	j->job = process_seqs_begin(opt, idx->bwt, idx->bns, idx->pac, n_processed, n, seqs, pes0,
		[](void *p, SeqJob *self) { b200_job *x = (b200_job *)p; x->sam_len = job_take_sam(self, &x->sam); take_routing(x, self); }, j, true, nullptr, 0, nullptr, 0, g_route);
	return j;
}

b200_job_t *b200_align_fastq_begin(const mem_opt_t *opt, const bwaidx_t *idx, int64_t n_processed, char *fq1, int64_t len1, char *fq2, int64_t len2)
{
	b200_job *j = new b200_job();
	j->total = 0; j->seqs = nullptr; j->sam = nullptr; j->sam_len = 0; j->n_threads = opt->n_threads;
	j->fq[0] = fq1; j->fq[1] = fq2; j->fq_len[0] = len1; j->fq_len[1] = len2; j->mate[0] = j->mate[1] = nullptr;
	// the job thread uploads the raw bytes; the device parses, interleaves and encodes them (fastq_kernels.h)
	j->job = process_seqs_begin(opt, idx->bwt, idx->bns, idx->pac, n_processed, 0, nullptr, nullptr,
		[](void *p, SeqJob *self) { b200_job *x = (b200_job *)p; x->sam_len = job_take_sam(self, &x->sam); take_routing(x, self); x->total = job_n_reads(self); }, j, true,
		fq1, len1, fq2, fq2 ? len2 : 0, g_route);
	return j;
}

int64_t b200_align_chunk_end(b200_job_t *j, char **sam, int64_t *sam_len, b200_stats_t *stats)
{
	process_seqs_end(j->job, stats);
	const int64_t total = j->total;
	if (sam) *sam = j->sam; else b200_free(j->sam);      // (pool buffer: never free())
	if (sam_len) *sam_len = j->sam_len;
	free(j->lines); free(j->dest_off);
	delete j;
	return total;
}

int64_t b200_align_chunk_end_routed(b200_job_t *j, char **sam, int64_t *sam_len, b200_sam_line_t **lines, int64_t *n_lines,
                                    int64_t **dest_off, int *n_dest, b200_stats_t *stats)
{
	process_seqs_end(j->job, stats);
	j->job = nullptr;
	if (lines) { *lines = j->lines; j->lines = nullptr; }
	if (n_lines) *n_lines = j->n_lines;
	if (dest_off) { *dest_off = j->dest_off; j->dest_off = nullptr; }
	if (n_dest) *n_dest = j->n_dest;
	const int64_t total = j->total;
	if (sam) *sam = j->sam; else b200_free(j->sam);
	if (sam_len) *sam_len = j->sam_len;
	free(j->lines); free(j->dest_off);
	delete j;
	return total;
}

int64_t b200_align_chunk(const mem_opt_t *opt, const bwaidx_t *idx, int64_t n_processed, int64_t n, bseq1_t *s1, bseq1_t *s2,
                         char **sam, int64_t *sam_len)
{
	return b200_align_chunk_end(b200_align_chunk_begin(opt, idx, n_processed, n, s1, s2), sam, sam_len, nullptr);
}

double b200_ext_replay(const mem_opt_t *opt, int64_t *cells, int64_t *n_jobs)
{
	return stage_extend_replay(engine_current(), make_ext_opt(opt), cells, n_jobs);    // (bench tool: replays what the primary engine recorded; no job may be in flight)
}

void b200_get_stats(b200_stats_t *out)
{
	if (engine_current()) last_stats(out);
	else memset(out, 0, sizeof *out);
}

static ExtOpt ext_opt_from(const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int zdrop)
{
	ExtOpt e;
	memset(&e, 0, sizeof e);
	e.o_del = o_del; e.e_del = e_del; e.o_ins = o_ins; e.e_ins = e_ins; e.zdrop = zdrop;
	e.max_sc = 0;
	for (int i = 0; i < 25; ++i) { e.mat[i] = mat[i]; if (mat[i] > e.max_sc) e.max_sc = mat[i]; }
	e.a = mat[0];
	return e;
}

int b200_ksw_extend2_batch(int64_t n_jobs, b200_extend_job_t *jobs, const uint8_t *query, int64_t query_bytes,
                           const uint8_t *target, int64_t target_bytes, const int8_t mat[25],
                           int o_del, int e_del, int o_ins, int e_ins, int zdrop)
{
	AuxGuard eng;
	stage_extend_bytes(eng, ext_opt_from(mat, o_del, e_del, o_ins, e_ins, zdrop), n_jobs, jobs,
	                   query, query_bytes, target, target_bytes);
	return 0;
}

int b200_ksw_align2_batch(int64_t n_jobs, b200_align_job_t *jobs, const uint8_t *query, int64_t query_bytes,
                          const uint8_t *target, int64_t target_bytes, const int8_t mat[25],
                          int o_del, int e_del, int o_ins, int e_ins)
{
	AuxGuard eng;
	stage_sw_bytes(eng, make_sw_opt(mat, o_del, e_del, o_ins, e_ins), n_jobs, jobs, query, query_bytes, target, target_bytes);
	return 0;
}

int b200_ksw_global2_batch(int64_t n_jobs, b200_global_job_t *jobs, const uint8_t *query, int64_t query_bytes,
                           const uint8_t *target, int64_t target_bytes, const int8_t mat[25],
                           int o_del, int e_del, int o_ins, int e_ins, uint32_t **cigar)
{
	GlobalOpt go;
	go.o_del = o_del; go.e_del = e_del; go.o_ins = o_ins; go.e_ins = e_ins; go.a = mat[0]; go.w_max = 0x3fffffff;
	memcpy(go.mat, mat, 25);
	std::vector<uint32_t> cg;
	{ AuxGuard eng; stage_global_batch(eng, go, n_jobs, jobs, query, query_bytes, target, target_bytes, cg); }
	*cigar = (uint32_t *)malloc(cg.size() * 4 + 4);
	memcpy(*cigar, cg.data(), cg.size() * 4);
	return 0;
}

int b200_collect_intv_batch(const mem_opt_t *opt, int n_reads, const int64_t *off, const uint8_t *seq,
                            bwtintv_t **intv, int64_t **intv_off)
{
	std::vector<int64_t> io;
	std::vector<Intv> iv;
	{ AuxGuard eng; stage_collect_intv(eng, make_seed_opt(opt), n_reads, off, seq, io, iv); }
	*intv_off = (int64_t *)malloc(io.size() * sizeof(int64_t));
	memcpy(*intv_off, io.data(), io.size() * sizeof(int64_t));
	*intv = (bwtintv_t *)malloc((iv.size() + 1) * sizeof(bwtintv_t));
	memcpy(*intv, iv.data(), iv.size() * sizeof(bwtintv_t));
	return 0;
}

int b200_bwt_sa_batch(int64_t n, const bwtint_t *k, bwtint_t *sa)
{
	AuxGuard eng;
	stage_sa(eng, n, k, sa);
	return 0;
}

/* ------------------------------------------------------------------ the reference's inner call surface, as batches of one */

int ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore, int *max_off)
{
	if (m != 5) die("ksw_extend2: only the 5-letter nucleotide alphabet is supported", nullptr);
	if (h0 <= 0) die("ksw_extend2: h0 must be positive", nullptr);
	b200_extend_job_t j;
	memset(&j, 0, sizeof j);
	j.qlen = qlen; j.tlen = tlen; j.q_off = 0; j.t_off = 0; j.h0 = h0; j.w = w; j.end_bonus = end_bonus;
	b200_ksw_extend2_batch(1, &j, query, qlen, target, tlen, mat, o_del, e_del, o_ins, e_ins, zdrop);
	if (qle) *qle = j.qle;
	if (tle) *tle = j.tle;
	if (gtle) *gtle = j.gtle;
	if (gscore) *gscore = j.gscore;
	if (max_off) *max_off = j.max_off;
	return j.score;
}

kswr_t ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
                  int o_del, int e_del, int o_ins, int e_ins, int xtra, kswq_t **qry)
{
	if (m != 5) die("ksw_align2: only the 5-letter nucleotide alphabet is supported", nullptr);
	if (qry && *qry) die("ksw_align2: cached query profiles are not supported (pass qry = NULL)", nullptr);
	b200_align_job_t j;
	memset(&j, 0, sizeof j);
	j.qlen = qlen; j.tlen = tlen; j.xtra = xtra;
	b200_ksw_align2_batch(1, &j, query, qlen, target, tlen, mat, o_del, e_del, o_ins, e_ins);
	return j.r;
}

int ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t **cigar)
{
	if (m != 5) die("ksw_global2: only the 5-letter nucleotide alphabet is supported", nullptr);
	GlobalOpt go;
	go.o_del = o_del; go.e_del = e_del; go.o_ins = o_ins; go.e_ins = e_ins; go.a = mat[0]; go.w_max = 0x3fffffff;
	memcpy(go.mat, mat, 25);
	std::vector<uint32_t> cg;
	int score;
	{ AuxGuard eng; score = stage_global_bytes(eng, go, qlen, query, tlen, target, w, (n_cigar && cigar) ? &cg : nullptr); }
	if (n_cigar) *n_cigar = 0;
	if (n_cigar && cigar) {
		*n_cigar = (int)cg.size();
		*cigar = (uint32_t *)malloc(cg.size() * 4 + 4);
		memcpy(*cigar, cg.data(), cg.size() * 4);
	}
	return score;
}

void bwt_extend(const bwt_t *bwt, const bwtintv_t *ik, bwtintv_t ok[4], int is_back)
{
	(void)bwt;
	Intv in = { ik->x[0], ik->x[1], ik->x[2], ik->info }, out[4];
	{ AuxGuard eng; stage_fm_extend(eng, in, out, is_back); }
	for (int i = 0; i < 4; ++i) { ok[i].x[0] = out[i].x0; ok[i].x[1] = out[i].x1; ok[i].x[2] = out[i].x2; }
}

bwtint_t bwt_sa(const bwt_t *bwt, bwtint_t k)
{
	(void)bwt;
	bwtint_t r;
	AuxGuard eng;
	stage_sa(eng, 1, &k, &r);
	return r;
}

void mem_chain2aln(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_query,
                   const uint8_t *query, const mem_chain_t *c, mem_alnreg_v *av)
{
	(void)pac;
	if (c->n == 0) return;
	AuxGuard eng;
	// one read, one chain; regions already in av take part in the containment test, so they are passed along
	HChain hc;
	hc.rid = c->rid; hc.frac_rep = c->frac_rep; hc.pos = c->pos;
	for (int i = 0; i < c->n; ++i) hc.seeds.push_back({c->seeds[i].rbeg, c->seeds[i].qbeg, c->seeds[i].len, c->seeds[i].score});
	int64_t rmax[2];
	chain_window(opt, bns, l_query, hc, rmax);
	std::vector<int64_t> off = {0, l_query};
	std::vector<uint8_t> codes(query, query + l_query);
	codes.resize(l_query + 8);
	stage_upload_reads(eng, 1, off.data(), codes.data());
	std::vector<int32_t> chain_off = {0, 1};
	std::vector<DChain> dc(1);
	std::vector<DSeed> ds;
	std::vector<int32_t> srt;
	std::vector<uint64_t> key;
	// earlier regions are encoded as a leading pseudo-chain-free prefix: the device task receives them pre-filled
	dc[0].seed_beg = 0; dc[0].n_seeds = c->n; dc[0].rid = c->rid; dc[0].frac_rep = c->frac_rep;
	dc[0].rmax0 = rmax[0]; dc[0].rmax1 = rmax[1];
	for (int i = 0; i < c->n; ++i) {
		ds.push_back({c->seeds[i].rbeg, c->seeds[i].qbeg, c->seeds[i].len, c->seeds[i].score, 0});
		key.push_back((uint64_t)c->seeds[i].score << 32 | (uint32_t)i);
	}
	std::sort(key.begin(), key.end());
	for (uint64_t k : key) srt.push_back((int32_t)(uint32_t)k);
	if (av->n != 0) die("mem_chain2aln: a non-empty region list must go through mem_process_seqs", nullptr);
	ExtIn xin = { 1, chain_off.data(), dc.data(), 1, ds.data(), (int64_t)ds.size(), srt.data() };
	ExtRegs xr;
	stage_extend(eng, make_ext_opt(opt), xin);
	stage_extend_download(eng, xr);
	const DReg *regs = xr.regs;
	const int64_t *reg_off = xr.reg_off;
	for (int i = 0; i < (int)reg_off[1]; ++i) {
		if (av->n == av->m) { av->m = av->m ? av->m << 1 : 2; av->a = (mem_alnreg_t *)realloc(av->a, av->m * sizeof(mem_alnreg_t)); }
		mem_alnreg_t *a = &av->a[av->n++];
		memset(a, 0, sizeof *a);
		const DReg &d = regs[i];
		a->rb = d.rb; a->re = d.re; a->qb = d.qb; a->qe = d.qe; a->rid = d.rid; a->score = d.score; a->truesc = d.truesc;
		a->w = d.w; a->seedcov = d.seedcov; a->seedlen0 = d.seedlen0; a->frac_rep = d.frac_rep;
	}
}

int bwt_smem1(const bwt_t *bwt, int len, const uint8_t *q, int x, int min_intv, bwtintv_v *mem, bwtintv_v *tmpvec[2])
{
	(void)bwt; (void)tmpvec;                  // (the scratch vectors of the reference live on the device here)
	std::vector<Intv> out;
	int ret;
	{ AuxGuard eng; ret = stage_smem1(eng, len, q, x, min_intv < 1 ? 1 : (uint64_t)min_intv, out); }
	mem->n = 0;
	if (out.size() > mem->m) { mem->m = out.size(); mem->a = (bwtintv_t *)realloc(mem->a, mem->m * sizeof(bwtintv_t)); }
	for (const Intv &v : out) { bwtintv_t &o = mem->a[mem->n++]; o.x[0] = v.x0; o.x[1] = v.x1; o.x[2] = v.x2; o.info = v.info; }
	return ret;
}

void b200_get_aux_stats(b200_stats_t *out)
{
	AuxGuard eng;
	*out = engine_stats(eng);
}

} // extern "C"
