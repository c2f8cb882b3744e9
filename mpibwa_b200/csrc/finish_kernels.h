// finish_kernels.h - per-task bodies of the "finish" stages: everything mem_process_seqs does after seed extension,
// one device task per read, pair, region or output record (SURVEY.md row f3 and the host residue of rows a2/a12/a15):
//
//   regs_dedup        mem_sort_dedup_patch + mem_patch_reg            reference src/bwamem.c:406-489
//   pestat_gather     candidate insert sizes of mem_pestat            reference src/bwamem_pair.c:46-66 (the statistics
//                     themselves - quartiles, mean, std.dev in double - stay on the host: a few microseconds per chunk)
//   rescue_*          window planning and the sequential insert/skip logic of mem_matesw
//                                                                     reference src/bwamem_pair.c:111-180, 262-275
//   pair_decide       mem_mark_primary_se, mem_pair, mapQ, the record plan of mem_sam_pe / mem_reg2sam
//                                                                     reference src/bwamem.c:493-558, 952-1049,
//                                                                     src/bwamem_pair.c:182-243, 277-393, src/bwamem_extra.c:98-140
//   aln_finish        the part of mem_reg2aln after ksw_global2: NM / MD (bwa_gen_cigar2), position, clipping
//                                                                     reference src/bwa.c:167-207, src/bwamem.c:1123-1159
//   sam_format        mem_aln2sam                                     reference src/bwamem.c:825-946
//
// Bit-exactness.  Integer work is restated literally.  The floating-point decisions are IEEE operations in the reference's
// types and order (the library is compiled with -fmad=false so that nvcc never contracts a*b+c), and the three libm calls
// whose device versions differ from glibc in the last bit are replaced by tables computed on the HOST with glibc:
// log(i) for integer i (mapQ) and, per chunk, .721*log(2*erfc(|d-avg|/std/sqrt2))*a for every integer
// insert size d in [low, high] of each orientation (mem_pair).
// klib's unstable introsort is restated swap for swap (idx_introsort) wherever keys can tie.
//
// Host/device code: tests/hostemu loops the same bodies on the CPU.
#pragma once
#include <cstdint>
#include "../../include/mpibwa_b200.h"
#include "fm_kernels.h"
#include "ext_kernels.h"
#include "sw_kernels.h"
#include "global_kernels.h"

namespace b200 {

typedef mem_alnreg_t Reg;                   // 88 bytes, the reference's layout (src/bwamem.h:59-77)

struct SwJob {
	int64_t rb;             // first reference position of the target window (forward+reverse coordinate)
	int32_t tlen;
	int32_t read;           // index of the query read in the current batch
	int32_t is_rev;         // use the reverse complement of the read as the query
	int32_t xtra;
	int32_t q_beg, q_len;   // sub-range of the read used as query (whole read for mate rescue)
};

B200_HD uint64_t fin_mix64(uint64_t key)      // Thomas Wang's mix, reference src/utils.h:98-109
{
	key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
	key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
	return key;
}

/* ---------------------------------------------------------------- tables and shared context */

struct FinTables {
	const double *logtab; int n_log;            // logtab[i] = log((double)i), glibc
	mem_pestat_t pes[4];
	const double *pair_tab; int64_t pair_off[4]; // pair_tab[pair_off[d] + dist - pes[d].low], dist in [low, high]
};

struct ReadText { int64_t name_off, qual_off, comment_off; int32_t name_len, comment_len; };     // qual_off / comment_off < 0: none

struct FinCtx {
	mem_opt_t opt;
	FmView fm;
	const uint8_t *ctg_alt;
	const int64_t *ctg_name_off; const char *ctg_names;       // contig names, back to back (no terminators)
	const int64_t *ctg_anno_off; const char *ctg_annos;
	int n_reads, pe;
	int64_t n_processed;
	const int64_t *off; const uint8_t *codes;                 // resident reads (codes 0-4)
	const ReadText *rtext; const char *text;                  // names, qualities, comments
	char rg_id[256]; int rg_len;
	int32_t *err;                                             // [0]: first error code (FIN_ERR_*), [1]: its argument
};
enum { FIN_ERR_LOGTAB = 1, FIN_ERR_NAMES = 2, FIN_ERR_PAIRTAB = 3 };

B200_HD void fin_fail(const FinCtx &cx, int code, int arg)
{
#if defined(__CUDA_ARCH__)
	if (atomicCAS(cx.err, 0, code) == 0) cx.err[1] = arg;
#else
	if (cx.err[0] == 0) { cx.err[0] = code; cx.err[1] = arg; }
#endif
}

B200_HD double fin_log(const FinCtx &cx, const FinTables &tb, int64_t i)
{
	if (i < 0 || i >= tb.n_log) { fin_fail(cx, FIN_ERR_LOGTAB, (int)i); return 0.; }
	return tb.logtab[i];
}

/* ---------------------------------------------------------------- klib's introsort on an index array */

// ix[0..n) is sorted so that lt(ix[i+1], ix[i]) never holds, with exactly the swaps ks_introsort (reference
// src/ksort.h:162-214) performs on the elements themselves: median of three on ranges above 16, comb sort when the depth
// budget runs out, one final insertion sort.  lt(i, j) compares ELEMENTS i and j.
template <class LT>
B200_HD void idx_insertion(int32_t *ix, int s, int t, LT lt)
{
	for (int i = s + 1; i < t; ++i)
		for (int j = i; j > s && lt(ix[j], ix[j - 1]); --j) { const int32_t x = ix[j]; ix[j] = ix[j - 1]; ix[j - 1] = x; }
}

template <class LT>
B200_HD void idx_comb(int32_t *ix, int s, int n, LT lt)
{
	const double shrink = 1.2473309501039786540366528676643;
	int gap = n;
	bool swapped;
	do {
		if (gap > 2) {
			gap = (int)(gap / shrink);
			if (gap == 9 || gap == 10) gap = 11;
		}
		swapped = false;
		for (int i = s; i < s + n - gap; ++i) {
			const int j = i + gap;
			if (lt(ix[j], ix[i])) { const int32_t x = ix[i]; ix[i] = ix[j]; ix[j] = x; swapped = true; }
		}
	} while (swapped || gap > 2);
	if (gap != 1) idx_insertion(ix, s, s + n, lt);
}

template <class LT>
B200_HDN void idx_introsort(int32_t *ix, int n, LT lt)
{
	if (n < 1) return;
	if (n == 2) { if (lt(ix[1], ix[0])) { const int32_t x = ix[0]; ix[0] = ix[1]; ix[1] = x; } return; }
	int d;
	for (d = 2; (1ul << d) < (unsigned long)n; ++d) {}
	int stk_lo[72], stk_hi[72], stk_d[72], sp = 0;
	int s = 0, t = n - 1;
	d <<= 1;
	for (;;) {
		if (s < t) {
			if (--d == 0) { idx_comb(ix, s, t - s + 1, lt); t = s; continue; }
			int i = s, j = t, k = i + ((j - i) >> 1) + 1;
			if (lt(ix[k], ix[i])) { if (lt(ix[k], ix[j])) k = j; }
			else k = lt(ix[j], ix[i]) ? i : j;
			const int32_t pivot = ix[k];
			if (k != t) { const int32_t x = ix[k]; ix[k] = ix[t]; ix[t] = x; }
			for (;;) {
				do ++i; while (lt(ix[i], pivot));
				do --j; while (i <= j && lt(pivot, ix[j]));
				if (j <= i) break;
				const int32_t x = ix[i]; ix[i] = ix[j]; ix[j] = x;
			}
			{ const int32_t x = ix[i]; ix[i] = ix[t]; ix[t] = x; }
			if (i - s > t - i) {
				if (i - s > 16) { stk_lo[sp] = s; stk_hi[sp] = i - 1; stk_d[sp] = d; ++sp; }
				s = t - i > 16 ? i + 1 : t;
			} else {
				if (t - i > 16) { stk_lo[sp] = i + 1; stk_hi[sp] = t; stk_d[sp] = d; ++sp; }
				t = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (sp == 0) { idx_insertion(ix, 0, n, lt); return; }
			--sp; s = stk_lo[sp]; t = stk_hi[sp]; d = stk_d[sp];
		}
	}
}

// sorts the n regions at a[] by lt with ks_introsort's permutation (tmp: n spare regions, ix: n spare ints)
template <class LT>
B200_HD void regs_sort(Reg *a, int n, Reg *tmp, int32_t *ix, LT lt)
{
	if (n < 2) return;
	for (int i = 0; i < n; ++i) ix[i] = i;
	idx_introsort(ix, n, [&](int x, int y) { return lt(a[x], a[y]); });
	bool moved = false;
	for (int i = 0; i < n; ++i) if (ix[i] != i) { moved = true; break; }
	if (!moved) return;
	for (int i = 0; i < n; ++i) tmp[i] = a[i];
	for (int i = 0; i < n; ++i) a[i] = tmp[ix[i]];
}

/* ---------------------------------------------------------------- banded global alignment, score only (mem_patch_reg) */

// ksw_global2 without the direction matrix (reference src/ksw.c:504-606 with n_cigar_ == 0); ROW: H/E row accessor over
// qlen + 1 columns, SEQ: oriented query/target accessor (global_kernels.h)
template <class ROW, class SEQ>
B200_HDN int global_score(const GlobalOpt &o, const SEQ &s, int tlen, int w, ROW eh)
{
	const int qlen = s.l_query;
	const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
	int j;
	eh.set_h(0, 0); eh.set_e(0, B200_GLOBAL_MINUS_INF);
	for (j = 1; j <= qlen && j <= w; ++j) { eh.set_h(j, -(o.o_ins + o.e_ins * j)); eh.set_e(j, B200_GLOBAL_MINUS_INF); }
	for (; j <= qlen; ++j) { eh.set_h(j, B200_GLOBAL_MINUS_INF); eh.set_e(j, B200_GLOBAL_MINUS_INF); }
	for (int i = 0; i < tlen; ++i) {
		int32_t f = B200_GLOBAL_MINUS_INF, h1, t;
		const auto mrow = s.trow(o, i);
		const int beg = i > w ? i - w : 0;
		const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
		h1 = beg == 0 ? -(o.o_del + o.e_del * (i + 1)) : B200_GLOBAL_MINUS_INF;
		for (j = beg; j < end; ++j) {
			int32_t h, m = eh.h(j), e = eh.e(j);
			eh.set_h(j, h1);
			m += s.sub(mrow, j);
			h = m >= e ? m : e;
			h = h >= f ? h : f;
			h1 = h;
			t = m - oe_del;
			e -= o.e_del;
			e = e > t ? e : t;
			eh.set_e(j, e);
			t = m - oe_ins;
			f -= o.e_ins;
			f = f > t ? f : t;
		}
		eh.set_h(end, h1); eh.set_e(end, B200_GLOBAL_MINUS_INF);
	}
	return eh.h(qlen);
}

B200_HD GlobalOpt fin_global_opt(const mem_opt_t &opt)
{
	GlobalOpt go;
	go.o_del = opt.o_del; go.e_del = opt.e_del; go.o_ins = opt.o_ins; go.e_ins = opt.e_ins; go.a = opt.a; go.w_max = opt.w << 2;
	for (int i = 0; i < 25; ++i) go.mat[i] = opt.mat[i];
	return go;
}

// The score bwa_gen_cigar2 returns when asked for no CIGAR (mem_patch_reg's use, reference src/bwa.c:121-166).
// rows: H/E row scratch for l_query + 1 columns, or null - then a region that needs the banded DP sets *need_dp and 0 comes back.
B200_HDN int fin_patch_score(const FinCtx &cx, const uint8_t *query, int l_query, int64_t rb, int64_t re, int w_, int32_t *rows, int64_t stride,
                            bool *need_dp)
{
	const int64_t l_pac = cx.fm.l_pac;
	if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
	if (rb < 0 || re > l_pac << 1) return 0;                // the fetch would come back shorter than the interval
	const GlobalOpt go = fin_global_opt(cx.opt);
	GlobalJob jb;
	jb.rb = rb; jb.re = re; jb.qb = 0; jb.qe = l_query;
	const GlobalSeqs s = global_seqs(cx.fm.pac, l_pac, query, jb);
	const int rlen = (int)(re - rb);
	if (l_query == rlen && w_ == 0) {
		int sc = 0;
		for (int i = 0; i < l_query; ++i) sc += go.mat[s.ta(i) * 5 + s.qa(i)];
		return sc;
	}
	if (!rows) { *need_dp = true; return 0; }
	const int w = global_band(go, l_query, rlen, w_);
	GlobalRow eh = { rows, stride };
	return global_score(go, s, rlen, w, eh);
}

/* ---------------------------------------------------------------- mem_sort_dedup_patch */

// mem_patch_reg, reference src/bwamem.c:406-435.  patch = false: the call sites that pass bns = NULL (mate rescue).
B200_HD int fin_patch_reg(const FinCtx &cx, bool patch, const uint8_t *query, const Reg *a, const Reg *b, int *_w, int32_t *rows, int64_t stride,
                          bool *need_dp)
{
	const mem_opt_t &opt = cx.opt;
	const int64_t l_pac = cx.fm.l_pac;
	if (!patch) return 0;
	if (a->rb < l_pac && b->rb >= l_pac) return 0;
	if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
	int w = (int)((a->re - b->rb) - (a->qe - b->qb));
	w = w > 0 ? w : -w;
	double r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
	r = r > 0. ? r : -r;
	if (a->re < b->rb || a->qe < b->qb) {
		if (w > opt.w << 1 || r >= 0.05f) return 0;
	} else if (w > opt.w << 2 || r >= 0.05f * 2) return 0;
	w += a->w + b->w;
	w = w < opt.w << 2 ? w : opt.w << 2;
	const int score = fin_patch_score(cx, query + a->qb, b->qe - a->qb, a->rb, b->re, w, rows, stride, need_dp);
	if (*need_dp) return 0;
	const int q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
	const int r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
	if ((double)score / (q_s > r_s ? q_s : r_s) < 0.90f) return 0;
	*_w = w;
	return score;
}

// mem_sort_dedup_patch, reference src/bwamem.c:437-489, in place on a[0..n); returns the new count.  When the read needs a
// banded DP for mem_patch_reg and rows == null, *need_dp is set and the contents of a[] are unspecified (the caller reruns
// the read from its input with row scratch).
B200_HDN int regs_dedup(const FinCtx &cx, bool patch, const uint8_t *query, int n, Reg *a, Reg *tmp, int32_t *ix, int32_t *rows, int64_t stride,
                        bool *need_dp)
{
	const mem_opt_t &opt = cx.opt;
	int m, i, j;
	if (n <= 1) return n;
	regs_sort(a, n, tmp, ix, [](const Reg &x, const Reg &y) { return x.re < y.re; });
	for (i = 0; i < n; ++i) a[i].n_comp = 1;
	for (i = 1; i < n; ++i) {
		Reg *p = &a[i];
		if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + opt.max_chain_gap) continue;
		for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt.max_chain_gap; --j) {
			Reg *q = &a[j];
			int64_t orr, oq, mr, mq;
			int score, w;
			if (q->qe == q->qb) continue;
			orr = q->re - p->rb;
			oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
			mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
			mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
			if (orr > opt.mask_level_redun * mr && oq > opt.mask_level_redun * mq) {
				if (p->score < q->score) { p->qe = p->qb; break; }
				else q->qe = q->qb;
			} else if (q->rb < p->rb && (score = fin_patch_reg(cx, patch, query, q, p, &w, rows, stride, need_dp)) > 0) {
				p->n_comp += q->n_comp + 1;
				p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
				p->sub = p->sub > q->sub ? p->sub : q->sub;
				p->csub = p->csub > q->csub ? p->csub : q->csub;
				p->qb = q->qb; p->rb = q->rb;
				p->truesc = p->score = score;
				p->w = w;
				q->qb = q->qe;
			}
			if (*need_dp) return 0;
		}
	}
	for (i = 0, m = 0; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
	n = m;
	regs_sort(a, n, tmp, ix, [](const Reg &x, const Reg &y) {
		return x.score > y.score || (x.score == y.score && (x.rb < y.rb || (x.rb == y.rb && x.qb < y.qb)));
	});
	for (i = 1; i < n; ++i)
		if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb)
			a[i].qe = a[i].qb;
	for (i = 1, m = 1; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
	return m;
}

// the regions of one read as mem_chain2aln left them -> mem_align1_core's result (reference src/bwamem.c:1073-1085)
B200_HD int regs_from_ext(const FinCtx &cx, const uint8_t *query, int n, const DReg *in, Reg *a, Reg *tmp, int32_t *ix, int32_t *rows, int64_t stride,
                          bool *need_dp)
{
	for (int k = 0; k < n; ++k) {
		const DReg d = in[k];
		Reg r;
		r.rb = d.rb; r.re = d.re; r.qb = d.qb; r.qe = d.qe; r.rid = d.rid; r.score = d.score; r.truesc = d.truesc;
		r.sub = 0; r.alt_sc = 0; r.csub = 0; r.sub_n = 0; r.w = d.w; r.seedcov = d.seedcov; r.secondary = 0; r.secondary_all = 0;
		r.seedlen0 = d.seedlen0; r.n_comp = 0; r.is_alt = 0; r.frac_rep = d.frac_rep; r.hash = 0;
		a[k] = r;
	}
	n = regs_dedup(cx, true, query, n, a, tmp, ix, rows, stride, need_dp);
	if (*need_dp) return 0;
	for (int k = 0; k < n; ++k)
		if (a[k].rid >= 0 && cx.ctg_alt[a[k].rid]) a[k].is_alt = 1;
	return n;
}

/* ---------------------------------------------------------------- insert-size candidates (mem_pestat) */

B200_HD int fin_infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)      // mem_infer_dir, reference src/bwamem_pair.c:23-30
{
	const int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
	const int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
	*dist = p2 > b1 ? p2 - b1 : b1 - p2;
	return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

B200_HD int fin_unique_sub(const mem_opt_t &opt, int n, const Reg *r)            // cal_sub, reference src/bwamem_pair.c:32-44
{
	int j;
	for (j = 1; j < n; ++j) {
		const int b_max = r[j].qb > r[0].qb ? r[j].qb : r[0].qb;
		const int e_min = r[j].qe < r[0].qe ? r[j].qe : r[0].qe;
		if (e_min > b_max) {
			const int min_l = r[j].qe - r[j].qb < r[0].qe - r[0].qb ? r[j].qe - r[j].qb : r[0].qe - r[0].qb;
			if (e_min - b_max >= min_l * opt.mask_level) break;
		}
	}
	return j < n ? r[j].score : opt.min_seed_len * opt.a;
}

// returns dir << 32 | insert size of the pair's candidate, or 0 when it has none
B200_HD uint64_t pestat_candidate(const FinCtx &cx, int n0, const Reg *r0, int n1, const Reg *r1)
{
	if (n0 == 0 || n1 == 0) return 0;
	if (fin_unique_sub(cx.opt, n0, r0) > 0.8 * r0[0].score) return 0;
	if (fin_unique_sub(cx.opt, n1, r1) > 0.8 * r1[0].score) return 0;
	if (r0[0].rid != r1[0].rid) return 0;
	int64_t is;
	const int dir = fin_infer_dir(cx.fm.l_pac, r0[0].rb, r1[0].rb, &is);
	if (is && is <= cx.opt.max_ins) return (uint64_t)dir << 32 | (uint64_t)is;
	return 0;
}

/* ---------------------------------------------------------------- mate rescue */

B200_HD void fin_clip_window(const FmView &fm, int64_t *beg, int64_t mid, int64_t *end, int *rid)   // the window part of bns_fetch_seq
{
	int is_rev;
	if (*end < *beg) { const int64_t t = *beg; *beg = *end; *end = t; }
	*rid = fm_pos2rid(fm, fm_depos(fm, mid, &is_rev));
	int64_t far_beg = fm.ctg_off[*rid], far_end = far_beg + fm.ctg_len[*rid];
	if (is_rev) { const int64_t t = far_beg; far_beg = (fm.l_pac << 1) - far_end; far_end = (fm.l_pac << 1) - t; }
	*beg = *beg > far_beg ? *beg : far_beg;
	*end = *end < far_end ? *end : far_end;
}

// window of mem_matesw for orientation r; false when the reference would not run SW
B200_HD bool rescue_window(const FinCtx &cx, const mem_pestat_t *pes, const Reg *a, int l_ms, int r, int64_t *rb_, int64_t *re_, int *is_rev_)
{
	const int64_t l_pac = cx.fm.l_pac;
	const int is_rev = (r >> 1 != (r & 1)), is_larger = !(r >> 1);
	int rid = -1;
	int64_t rb, re;
	if (!is_rev) {
		rb = is_larger ? a->rb + pes[r].low : a->rb - pes[r].high;
		re = (is_larger ? a->rb + pes[r].high : a->rb - pes[r].low) + l_ms;
	} else {
		rb = (is_larger ? a->rb + pes[r].low : a->rb - pes[r].high) - l_ms;
		re = is_larger ? a->rb + pes[r].high : a->rb - pes[r].low;
	}
	if (rb < 0) rb = 0;
	if (re > l_pac << 1) re = l_pac << 1;
	if (rb >= re) return false;
	fin_clip_window(cx.fm, &rb, (rb + re) >> 1, &re, &rid);
	*rb_ = rb; *re_ = re; *is_rev_ = is_rev;
	return a->rid == rid && re - rb >= cx.opt.min_seed_len;
}

B200_HD void rescue_skip_mask(const FinCtx &cx, const mem_pestat_t *pes, const Reg *a, int n_ma, const Reg *ma, int skip[4])
{
	for (int r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
	for (int i = 0; i < n_ma; ++i) {
		int64_t dist;
		const int r = fin_infer_dir(cx.fm.l_pac, a->rb, ma[i].rb, &dist);
		if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
	}
}

// the anchors of one end: the first max_matesw regions within pen_unpaired of the best (reference src/bwamem_pair.c:262-270)
B200_HD int rescue_n_anchors(const mem_opt_t &opt, int n, const Reg *a)
{
	int nb = 0;
	for (int j = 0; j < n && nb < opt.max_matesw; ++j)
		if (a[j].score >= a[0].score - opt.pen_unpaired) ++nb;
	return nb;
}

#define RESCUE_KEY(end, anchor, r) ((int32_t)((end) << 30 | (anchor) << 2 | (r)))

// round 0 of a pair: every (end, anchor, orientation) not ruled out by the mate's regions before rescue.  With jobs == null
// only counts.  Returns the number of jobs.
B200_HD int rescue_plan_pair(const FinCtx &cx, const mem_pestat_t *pes, int64_t p, const int n[2], const Reg *const a[2], SwJob *jobs, int32_t *keys)
{
	const mem_opt_t &opt = cx.opt;
	const int xtra_base = KSW_XSUBO | KSW_XSTART | (opt.min_seed_len * opt.a);
	int nj = 0;
	for (int i = 0; i < 2; ++i) {
		const Reg *ai = a[i], *ma = a[!i];
		const int l_ms = (int)(cx.off[(p << 1 | !i) + 1] - cx.off[p << 1 | !i]);
		int nb = 0;
		for (int j = 0; j < n[i] && nb < opt.max_matesw; ++j) {
			if (ai[j].score < ai[0].score - opt.pen_unpaired) continue;
			int skip[4];
			rescue_skip_mask(cx, pes, &ai[j], n[!i], ma, skip);
			for (int r = 0; r < 4; ++r) {
				int64_t rb, re;
				int is_rev;
				if (skip[r] || !rescue_window(cx, pes, &ai[j], l_ms, r, &rb, &re, &is_rev)) continue;
				if (jobs) {
					SwJob jb;
					jb.rb = rb; jb.tlen = (int)(re - rb); jb.read = (int32_t)(p << 1 | !i); jb.is_rev = is_rev;
					jb.xtra = xtra_base | (l_ms * opt.a < 250 ? KSW_XBYTE : 0);
					jb.q_beg = 0; jb.q_len = l_ms;
					jobs[nj] = jb; keys[nj] = RESCUE_KEY(i, nb, r);
				}
				++nj;
			}
			++nb;
		}
	}
	return nj;
}

// the job for one (end, anchor, orientation) that the replay found missing
B200_HD void rescue_extra_job(const FinCtx &cx, const mem_pestat_t *pes, int64_t p, int32_t key, const int n[2], const Reg *const a[2], SwJob *job)
{
	const mem_opt_t &opt = cx.opt;
	const int end = (uint32_t)key >> 30, anchor = (key >> 2) & 0xfffffff, r = key & 3;
	const Reg *ai = a[end], *an = nullptr;
	int nb = -1;
	for (int j = 0; j < n[end]; ++j) {
		if (ai[j].score < ai[0].score - opt.pen_unpaired) continue;
		if (++nb == anchor) { an = &ai[j]; break; }
	}
	const int l_ms = (int)(cx.off[(p << 1 | !end) + 1] - cx.off[p << 1 | !end]);
	int64_t rb = 0, re = 0;
	int is_rev = 0;
	rescue_window(cx, pes, an, l_ms, r, &rb, &re, &is_rev);
	job->rb = rb; job->tlen = (int)(re - rb); job->read = (int32_t)(p << 1 | !end); job->is_rev = is_rev;
	job->xtra = KSW_XSUBO | KSW_XSTART | (opt.min_seed_len * opt.a) | (l_ms * opt.a < 250 ? KSW_XBYTE : 0);
	job->q_beg = 0; job->q_len = l_ms;
}

// results known for one pair: its round-0 jobs [jb, jb + jn) plus a chain of extras (next[] links, -1 ends)
struct RescueHave {
	const SwJob *jobs; const int32_t *keys; const SwRes *res; const int32_t *next;
	int64_t jb; int jn; int32_t extra;
	B200_HD int64_t find(int32_t key) const
	{
		for (int64_t x = jb; x < jb + jn; ++x) if (keys[x] == key) return x;
		for (int32_t x = extra; x >= 0; x = next[x]) if (keys[x] == key) return x;
		return -1;
	}
};

// Replays mem_sam_pe's rescue loop (reference src/bwamem_pair.c:262-275 around mem_matesw :111-180) for one pair with the
// precomputed SW results.  src[i]: the regions of end i before rescue (immutable), w[i]: working lists (capacity: n + 4 per
// anchor of the other end), initialised here.  Returns -1 when done (nw[] = final counts) or the key of the first result
// the reference would compute that is not in `have`.
B200_HDN int32_t rescue_replay_pair(const FinCtx &cx, const mem_pestat_t *pes, int64_t p, const int n[2], const Reg *const src[2], Reg *const w[2],
                                   int nw[2], Reg *tmp, int32_t *ix, const RescueHave &have)
{
	const mem_opt_t &opt = cx.opt;
	const int64_t l_pac = cx.fm.l_pac;
	for (int i = 0; i < 2; ++i) { for (int k = 0; k < n[i]; ++k) w[i][k] = src[i][k]; nw[i] = n[i]; }
	bool dummy = false;
	for (int i = 0; i < 2; ++i) {
		const int l_ms = (int)(cx.off[(p << 1 | !i) + 1] - cx.off[p << 1 | !i]);
		Reg *ma = w[!i];
		int &n_ma = nw[!i];
		int nb = 0;
		for (int j = 0; j < n[i] && nb < opt.max_matesw; ++j) {
			if (src[i][j].score < src[i][0].score - opt.pen_unpaired) continue;
			const Reg *a = &src[i][j];
			int skip[4], nn = 0;
			rescue_skip_mask(cx, pes, a, n_ma, ma, skip);
			if (skip[0] + skip[1] + skip[2] + skip[3] != 4) {
				for (int r = 0; r < 4; ++r) {
					if (skip[r]) continue;
					int64_t rb, re;
					int is_rev;
					if (rescue_window(cx, pes, a, l_ms, r, &rb, &re, &is_rev)) {
						const int32_t key = RESCUE_KEY(i, nb, r);
						const int64_t x = have.find(key);
						if (x < 0) return key;
						const SwRes aln = have.res[x];
						if (aln.score >= opt.min_seed_len && aln.qb >= 0) {
							Reg b;
							b.rb = is_rev ? (l_pac << 1) - (rb + aln.te + 1) : rb + aln.tb;
							b.re = is_rev ? (l_pac << 1) - (rb + aln.tb) : rb + aln.te + 1;
							b.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
							b.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
							b.rid = a->rid; b.score = aln.score; b.truesc = 0; b.sub = 0; b.alt_sc = 0; b.csub = aln.score2; b.sub_n = 0; b.w = 0;
							b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
							b.secondary = -1; b.secondary_all = 0; b.seedlen0 = 0; b.n_comp = 0; b.is_alt = a->is_alt; b.frac_rep = 0; b.hash = 0;
							int at;
							for (at = 0; at < n_ma; ++at) if (ma[at].score < b.score) break;
							for (int k = n_ma; k > at; --k) ma[k] = ma[k - 1];
							ma[at] = b;
							++n_ma;
						}
						++nn;
					}
					if (nn) n_ma = regs_dedup(cx, false, nullptr, n_ma, ma, tmp, ix, nullptr, 0, &dummy);
				}
			}
			++nb;
		}
	}
	return -1;
}

/* ---------------------------------------------------------------- primary marking, mapQ, pairing */

B200_HD void fin_mark_primary_core(const mem_opt_t &opt, int n, Reg *a, int32_t *z)       // reference src/bwamem.c:493-521
{
	int tmp = opt.a + opt.b, nz = 0;
	tmp = opt.o_del + opt.e_del > tmp ? opt.o_del + opt.e_del : tmp;
	tmp = opt.o_ins + opt.e_ins > tmp ? opt.o_ins + opt.e_ins : tmp;
	z[nz++] = 0;
	for (int i = 1; i < n; ++i) {
		int k;
		for (k = 0; k < nz; ++k) {
			const int j = z[k];
			const int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
			const int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
			if (e_min > b_max) {
				const int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
				if (e_min - b_max >= min_l * opt.mask_level) {
					if (a[j].sub == 0) a[j].sub = a[i].score;
					if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
					break;
				}
			}
		}
		if (k == nz) z[nz++] = i;
		else a[i].secondary = z[k];
	}
}

// mem_mark_primary_se, reference src/bwamem.c:523-558 (both sorts have total orders: the hash is a bijection of the index)
B200_HDN int fin_mark_primary_se(const mem_opt_t &opt, int n, Reg *a, int64_t id, Reg *tmp, int32_t *ix, int32_t *z)
{
	int i, n_pri;
	if (n == 0) return 0;
	for (i = n_pri = 0; i < n; ++i) {
		a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = fin_mix64((uint64_t)(id + i));
		if (!a[i].is_alt) ++n_pri;
	}
	regs_sort(a, n, tmp, ix, [](const Reg &x, const Reg &y) {
		return x.score > y.score || (x.score == y.score && (x.is_alt < y.is_alt || (x.is_alt == y.is_alt && x.hash < y.hash)));
	});
	fin_mark_primary_core(opt, n, a, z);
	for (i = 0; i < n; ++i) {
		Reg *p = &a[i];
		p->secondary_all = i;
		if (!p->is_alt && p->secondary >= 0 && a[p->secondary].is_alt) p->alt_sc = a[p->secondary].score;
	}
	if (n_pri >= 0 && n_pri < n) {
		if (n_pri > 0)
			regs_sort(a, n, tmp, ix, [](const Reg &x, const Reg &y) {
				return x.is_alt < y.is_alt || (x.is_alt == y.is_alt && (x.score > y.score || (x.score == y.score && x.hash < y.hash)));
			});
		for (i = 0; i < n; ++i) z[a[i].secondary_all] = i;
		for (i = 0; i < n; ++i) {
			if (a[i].secondary >= 0) {
				a[i].secondary_all = z[a[i].secondary];
				if (a[i].is_alt) a[i].secondary = 0x7fffffff;
			} else a[i].secondary_all = -1;
		}
		if (n_pri > 0) {
			for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
			fin_mark_primary_core(opt, n_pri, a, z);
		}
	} else {
		for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
	}
	return n_pri;
}

B200_HD void fin_reorder_primary5(int T, int n, Reg *a)          // mem_reorder_primary5, reference src/bwamem.c:978-1000
{
	int n_pri = 0, left_st = 0x7fffffff, left_k = -1;
	for (int k = 0; k < n; ++k)
		if (a[k].secondary < 0 && !a[k].is_alt && a[k].score >= T) ++n_pri;
	if (n_pri <= 1) return;
	for (int k = 0; k < n; ++k) {
		const Reg *p = &a[k];
		if (p->secondary >= 0 || p->is_alt || p->score < T) continue;
		if (p->qb < left_st) { left_st = p->qb; left_k = k; }
	}
	if (left_k == 0) return;
	const Reg t = a[0]; a[0] = a[left_k]; a[left_k] = t;
	for (int k = 1; k < n; ++k) {
		Reg *p = &a[k];
		if (p->secondary == 0) p->secondary = left_k;
		else if (p->secondary == left_k) p->secondary = 0;
		if (p->secondary_all == 0) p->secondary_all = left_k;
		else if (p->secondary_all == left_k) p->secondary_all = 0;
	}
}

// mem_approx_mapq_se, reference src/bwamem.c:952-976
B200_HD int fin_mapq_se(const FinCtx &cx, const FinTables &tb, const Reg *a)
{
	const mem_opt_t &opt = cx.opt;
	int mapq, l, sub = a->sub ? a->sub : opt.min_seed_len * opt.a;
	double identity;
	sub = a->csub > sub ? a->csub : sub;
	if (sub >= a->score) return 0;
	l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int)(a->re - a->rb);
	identity = 1. - (double)(l * opt.a - a->score) / (opt.a + opt.b) / l;
	if (a->score == 0) {
		mapq = 0;
	} else if (opt.mapQ_coef_len > 0) {
		double tmp;
		tmp = l < opt.mapQ_coef_len ? 1. : opt.mapQ_coef_fac / fin_log(cx, tb, l);
		tmp *= identity * identity;
		mapq = (int)(6.02 * (a->score - sub) / opt.a * tmp * tmp + .499);
	} else {
		mapq = (int)(30.0 * (1. - (double)sub / a->score) * fin_log(cx, tb, a->seedcov) + .499);
		mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
	}
	if (a->sub_n > 0) mapq -= (int)(4.343 * fin_log(cx, tb, a->sub_n + 1) + .499);
	if (mapq > 60) mapq = 60;
	if (mapq < 0) mapq = 0;
	mapq = (int)(mapq * (1. - a->frac_rep) + .499);
	return mapq;
}

struct FinPair64 { uint64_t x, y; };
B200_HD bool fin_pair_lt(const FinPair64 &a, const FinPair64 &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }

// mem_pair, reference src/bwamem_pair.c:182-243.  v: scratch for n_pri[0] + n_pri[1] keys, ix: as many ints.
// Both of the reference's sorts run over distinct keys, so any sort gives its order; the list of candidate pairs (u) is not
// materialised: the best, the second best and the count within `tmp` of the second best come from two passes over the
// same enumeration.
B200_HDN int fin_pair_ends(const FinCtx &cx, const FinTables &tb, const Reg *const a[2], const int n_pri[2], int id, int *sub, int *n_sub, int z[2],
                          FinPair64 *v, FinPair64 *vtmp, int32_t *ix)
{
	const mem_opt_t &opt = cx.opt;
	const int64_t l_pac = cx.fm.l_pac;
	int nv = 0;
	for (int r = 0; r < 2; ++r)
		for (int i = 0; i < n_pri[r]; ++i) {
			const Reg *e = &a[r][i];
			FinPair64 key;
			key.x = e->rb < l_pac ? e->rb : (l_pac << 1) - 1 - e->rb;
			key.x = (uint64_t)e->rid << 32 | (key.x - cx.fm.ctg_off[e->rid]);
			key.y = (uint64_t)e->score << 32 | i << 2 | (e->rb >= l_pac) << 1 | r;
			v[nv++] = key;
		}
	if (nv > 1) {
		for (int i = 0; i < nv; ++i) ix[i] = i;
		idx_introsort(ix, nv, [&](int x, int y) { return fin_pair_lt(v[x], v[y]); });
		for (int i = 0; i < nv; ++i) vtmp[i] = v[i];
		for (int i = 0; i < nv; ++i) v[i] = vtmp[ix[i]];
	}
	const uint64_t idmask = (uint64_t)(int64_t)(int32_t)((uint32_t)id << 8);     // "p->y ^ id<<8" with an int id
	FinPair64 best = { 0, 0 }, second = { 0, 0 };
	int64_t un = 0;
	int n_close = 0;
	for (int pass = 0; pass < 2; ++pass) {
		int y[4] = { -1, -1, -1, -1 };
		if (pass == 1 && un < 2) break;
		const int64_t sub_q = (int64_t)(second.x >> 32);
		int tmp = opt.a + opt.b;
		tmp = tmp > opt.o_del + opt.e_del ? tmp : opt.o_del + opt.e_del;
		tmp = tmp > opt.o_ins + opt.e_ins ? tmp : opt.o_ins + opt.e_ins;
		for (int i = 0; i < nv; ++i) {
			for (int r = 0; r < 2; ++r) {
				const int dir = r << 1 | (int)(v[i].y >> 1 & 1);
				if (tb.pes[dir].failed) continue;
				const int which = r << 1 | (int)((v[i].y & 1) ^ 1);
				if (y[which] < 0) continue;
				for (int k = y[which]; k >= 0; --k) {
					if ((int)(v[k].y & 3) != which) continue;
					const int64_t dist = (int64_t)v[i].x - (int64_t)v[k].x;
					if (dist > tb.pes[dir].high) break;
					if (dist < tb.pes[dir].low) continue;
					const double t = tb.pair_tab[tb.pair_off[dir] + dist - tb.pes[dir].low];
					int q = (int)((double)((v[i].y >> 32) + (v[k].y >> 32)) + t + .499);
					if (q < 0) q = 0;
					FinPair64 pr;
					pr.y = (uint64_t)k << 32 | (uint32_t)i;
					pr.x = (uint64_t)q << 32 | (fin_mix64(pr.y ^ idmask) & 0xffffffffU);
					if (pass == 0) {
						if (un == 0 || fin_pair_lt(best, pr)) { second = best; best = pr; }
						else if (un == 1 || fin_pair_lt(second, pr)) second = pr;
						++un;
					} else if (!(pr.x == best.x && pr.y == best.y) && sub_q - (int64_t)(pr.x >> 32) <= tmp) ++n_close;
				}
			}
			y[v[i].y & 3] = i;
		}
	}
	if (un == 0) { *sub = 0; *n_sub = 0; return 0; }
	{
		const int i = (int)(best.y >> 32), k = (int)(best.y << 32 >> 32);
		z[v[i].y & 1] = (int)(v[i].y << 32 >> 34);
		z[v[k].y & 1] = (int)(v[k].y << 32 >> 34);
	}
	*sub = un > 1 ? (int)(second.x >> 32) : 0;
	*n_sub = un > 1 ? n_close : 0;
	return (int)(best.x >> 32);
}

/* ---------------------------------------------------------------- the record plan (mem_sam_pe / mem_reg2sam) */

// one SAM line of a read
struct SamRec {
	int32_t reg;            // region of the read's list it reports, -1: the read is unmapped
	int32_t flag;           // flag bits before mem_aln2sam adds the pair bits
	int32_t mapq;
	int32_t sub;            // XS (-1: not printed)
};

// per region: does an output line need its alignment (REG_ALN), and of which region's XA tag is it a member (xa_of >= 0)
enum { REG_ALN = 1 };

B200_HD int fin_pri_idx(double XA_drop_ratio, const Reg *a, int i)      // get_pri_idx, reference src/bwamem_extra.c:98-104
{
	const int k = a[i].secondary_all;
	if (k >= 0 && a[i].score >= a[k].score * XA_drop_ratio) return k;
	return -1;
}

// mem_gen_alt's membership (reference src/bwamem_extra.c:106-140): xa_of[i] = the region whose XA tag lists region i, or -1;
// members are flagged for alignment.  cnt / has_alt: n spare ints each.
B200_HD void fin_plan_xa(const mem_opt_t &opt, int n, const Reg *a, int32_t *xa_of, uint8_t *need, int32_t *cnt, int32_t *has_alt)
{
	for (int i = 0; i < n; ++i) { xa_of[i] = -1; cnt[i] = 0; has_alt[i] = 0; }
	if (opt.flag & MEM_F_ALL) return;
	for (int i = 0; i < n; ++i) {
		const int r = fin_pri_idx(opt.XA_drop_ratio, a, i);
		if (r >= 0) { ++cnt[r]; if (a[i].is_alt) has_alt[r] = 1; }
	}
	for (int i = 0; i < n; ++i) {
		const int r = fin_pri_idx(opt.XA_drop_ratio, a, i);
		if (r < 0) continue;
		if (cnt[r] > opt.max_XA_hits_alt || (!has_alt[r] && cnt[r] > opt.max_XA_hits)) continue;
		xa_of[i] = r; need[i] |= REG_ALN;
	}
}

B200_HD bool fin_reg_mapped(const Reg *p) { return !(p->rb < 0 || p->re < 0); }

// mem_reg2sam's selection (reference src/bwamem.c:1003-1049); returns the number of records written to rec[]
B200_HD int fin_plan_reg2sam(const FinCtx &cx, const FinTables &tb, int n, const Reg *a, int extra_flag, SamRec *rec, uint8_t *need)
{
	const mem_opt_t &opt = cx.opt;
	int l = 0;
	for (int k = 0; k < n; ++k) {
		const Reg *p = &a[k];
		if (p->score < opt.T) continue;
		if (p->secondary >= 0 && (p->is_alt || !(opt.flag & MEM_F_ALL))) continue;
		if (p->secondary >= 0 && p->secondary < 0x7fffffff && p->score < a[p->secondary].score * opt.drop_ratio) continue;
		SamRec q;
		if (fin_reg_mapped(p)) {
			q.reg = k; q.flag = p->secondary >= 0 ? 0x100 : 0;
			q.mapq = p->secondary < 0 ? (fin_mapq_se(cx, tb, p) & 0xff) : 0;
			q.sub = p->sub > p->csub ? p->sub : p->csub;
			need[k] |= REG_ALN;
		} else { q.reg = -1; q.flag = 0x4; q.mapq = 0; q.sub = 0; }
		q.flag |= extra_flag;
		if (p->secondary >= 0) q.sub = -1;
		if (l && p->secondary < 0) q.flag |= (opt.flag & MEM_F_NO_MULTI) ? 0x10000 : 0x800;
		if (!(opt.flag & MEM_F_KEEP_SUPP_MAPQ) && l && !p->is_alt && q.mapq > rec[0].mapq) q.mapq = rec[0].mapq;
		rec[l++] = q;
	}
	if (l == 0) {
		SamRec q;
		q.reg = -1; q.flag = 0x4 | extra_flag; q.mapq = 0; q.sub = 0;
		rec[l++] = q;
	}
	return l;
}

#define FIN_RAW_MAPQ(diff, a) ((int)(6.02 * (diff) / (a) + .499))

// Scratch of one pair: everything is indexed like the pair's region slots (cap = slots of both reads together).
struct PairScratch { Reg *tmp; int32_t *ix, *z, *cnt, *has_alt; FinPair64 *v, *vtmp; };

// The part of mem_sam_pe after mate rescue (reference src/bwamem_pair.c:277-393) up to, but not including, the alignments
// and the text: final region state, which regions need an alignment, the output records of both reads and each read's mate
// region.  a[i]: regions of read i (n[i] of them), xa_of / need: per region, rec[i]: records of read i, mate_reg[i]: region
// of the MATE whose alignment is read i's mate (-1: mate unmapped).
B200_HDN void pair_decide(const FinCtx &cx, const FinTables &tb, uint64_t id, const int n[2], Reg *const a[2], int32_t *const xa_of[2],
                         uint8_t *const need[2], SamRec *const rec[2], int n_rec[2], int mate_reg[2], const PairScratch &S)
{
	const mem_opt_t &opt = cx.opt;
	int i, j, z[2] = { 0, 0 }, o, subo = 0, n_sub = 0, extra_flag = 1, n_pri[2];
	for (i = 0; i < 2; ++i) {
		for (j = 0; j < n[i]; ++j) { need[i][j] = 0; xa_of[i][j] = -1; }
		n_pri[i] = fin_mark_primary_se(opt, n[i], a[i], (int64_t)(id << 1 | (uint64_t)i), S.tmp, S.ix, S.z);
	}
	if (opt.flag & MEM_F_PRIMARY5) { fin_reorder_primary5(opt.T, n[0], a[0]); fin_reorder_primary5(opt.T, n[1], a[1]); }
	o = 0;
	if (!(opt.flag & MEM_F_NOPAIRING) && n_pri[0] && n_pri[1])
		o = fin_pair_ends(cx, tb, a, n_pri, (int)id, &subo, &n_sub, z, S.v, S.vtmp, S.ix);
	if (o > 0) {
		int is_multi[2], q_pe, score_un, q_se[2];
		for (i = 0; i < 2; ++i) {
			for (j = 1; j < n_pri[i]; ++j)
				if (a[i][j].secondary < 0 && a[i][j].score >= opt.T) break;
			is_multi[i] = j < n_pri[i] ? 1 : 0;
		}
		if (!(is_multi[0] || is_multi[1])) {
			score_un = a[0][0].score + a[1][0].score - opt.pen_unpaired;
			subo = subo > score_un ? subo : score_un;
			q_pe = FIN_RAW_MAPQ(o - subo, opt.a);
			if (n_sub > 0) q_pe -= (int)(4.343 * fin_log(cx, tb, n_sub + 1) + .499);
			if (q_pe < 0) q_pe = 0;
			if (q_pe > 60) q_pe = 60;
			q_pe = (int)(q_pe * (1. - .5 * (a[0][0].frac_rep + a[1][0].frac_rep)) + .499);
			if (o > score_un) {
				Reg *c[2];
				c[0] = &a[0][z[0]]; c[1] = &a[1][z[1]];
				for (i = 0; i < 2; ++i) {
					if (c[i]->secondary >= 0) { c[i]->sub = a[i][c[i]->secondary].score; c[i]->secondary = -2; }
					q_se[i] = fin_mapq_se(cx, tb, c[i]);
				}
				q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
				q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
				extra_flag |= 2;
				q_se[0] = q_se[0] < FIN_RAW_MAPQ(c[0]->score - c[0]->csub, opt.a) ? q_se[0] : FIN_RAW_MAPQ(c[0]->score - c[0]->csub, opt.a);
				q_se[1] = q_se[1] < FIN_RAW_MAPQ(c[1]->score - c[1]->csub, opt.a) ? q_se[1] : FIN_RAW_MAPQ(c[1]->score - c[1]->csub, opt.a);
			} else {
				z[0] = z[1] = 0;
				q_se[0] = fin_mapq_se(cx, tb, &a[0][0]);
				q_se[1] = fin_mapq_se(cx, tb, &a[1][0]);
			}
			for (i = 0; i < 2; ++i) {
				const int k = a[i][z[i]].secondary_all;
				if (k >= 0 && k < n_pri[i]) {
					for (j = 0; j < n[i]; ++j)
						if (a[i][j].secondary_all == k || j == k) a[i][j].secondary_all = z[i];
					a[i][z[i]].secondary_all = -1;
				}
			}
			for (i = 0; i < 2; ++i) fin_plan_xa(opt, n[i], a[i], xa_of[i], need[i], S.cnt, S.has_alt);
			for (i = 0; i < 2; ++i) {
				const Reg *p = &a[i][z[i]];
				SamRec h;
				// (a region with a negative coordinate would print as unmapped; mem_chain2aln and mem_matesw never produce one)
				h.reg = fin_reg_mapped(p) ? z[i] : -1;
				h.flag = (h.reg < 0 ? 0x4 : (p->secondary >= 0 ? 0x100 : 0)) | 0x40 << i | extra_flag;
				h.mapq = q_se[i] & 0xff;
				h.sub = h.reg < 0 ? 0 : (p->sub > p->csub ? p->sub : p->csub);
				if (h.reg >= 0) need[i][z[i]] |= REG_ALN;
				rec[i][0] = h; n_rec[i] = 1;
				mate_reg[!i] = h.reg;
				if (n_pri[i] < n[i]) {
					const Reg *g = &a[i][n_pri[i]];
					if (g->score < opt.T || g->secondary >= 0 || !g->is_alt) continue;
					SamRec q;
					q.reg = fin_reg_mapped(g) ? n_pri[i] : -1;
					q.flag = (q.reg < 0 ? 0x4 : 0) | 0x800 | 0x40 << i | extra_flag;
					q.mapq = q.reg < 0 ? 0 : (fin_mapq_se(cx, tb, g) & 0xff);
					q.sub = q.reg < 0 ? 0 : (g->sub > g->csub ? g->sub : g->csub);
					if (q.reg >= 0) need[i][n_pri[i]] |= REG_ALN;
					rec[i][1] = q; n_rec[i] = 2;
				}
			}
			return;
		}
	}
	// no_pairing
	int which[2], h_rid[2];
	for (i = 0; i < 2; ++i) {
		which[i] = -1;
		if (n[i] > 0) {
			if (a[i][0].score >= opt.T) which[i] = 0;
			else if (n_pri[i] < n[i] && a[i][n_pri[i]].score >= opt.T) which[i] = n_pri[i];
		}
		if (which[i] >= 0 && !fin_reg_mapped(&a[i][which[i]])) which[i] = -1;
		h_rid[i] = -1;
		if (which[i] >= 0) {
			const Reg *p = &a[i][which[i]];
			int is_rev;
			h_rid[i] = fm_pos2rid(cx.fm, fm_depos(cx.fm, p->rb < cx.fm.l_pac ? p->rb : p->re - 1, &is_rev));
			need[i][which[i]] |= REG_ALN;
		}
		mate_reg[!i] = which[i];
	}
	if (!(opt.flag & MEM_F_NOPAIRING) && h_rid[0] == h_rid[1] && h_rid[0] >= 0) {
		int64_t dist;
		const int d = fin_infer_dir(cx.fm.l_pac, a[0][0].rb, a[1][0].rb, &dist);
		if (!tb.pes[d].failed && dist >= tb.pes[d].low && dist <= tb.pes[d].high) extra_flag |= 2;
	}
	for (i = 0; i < 2; ++i) {
		fin_plan_xa(opt, n[i], a[i], xa_of[i], need[i], S.cnt, S.has_alt);
		n_rec[i] = fin_plan_reg2sam(cx, tb, n[i], a[i], (i ? 0x81 : 0x41) | extra_flag, rec[i], need[i]);
	}
}

// single-end reads: worker2's else branch, reference src/bwamem.c:1191-1196
B200_HD void single_decide(const FinCtx &cx, const FinTables &tb, int64_t id, int n, Reg *a, int32_t *xa_of, uint8_t *need, SamRec *rec, int *n_rec,
                           const PairScratch &S)
{
	for (int j = 0; j < n; ++j) { need[j] = 0; xa_of[j] = -1; }
	fin_mark_primary_se(cx.opt, n, a, id, S.tmp, S.ix, S.z);
	if (cx.opt.flag & MEM_F_PRIMARY5) fin_reorder_primary5(cx.opt.T, n, a);
	fin_plan_xa(cx.opt, n, a, xa_of, need, S.cnt, S.has_alt);
	*n_rec = fin_plan_reg2sam(cx, tb, n, a, 0, rec, need);
}

/* ---------------------------------------------------------------- alignments of the flagged regions */

B200_HD int fin_infer_bw(int l1, int l2, int score, int a, int q, int r)       // infer_bw, reference src/bwamem.c:1002-1009 (bwa 0.7.17 numbering: :1092)
{
	if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
	int w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
	const int d = l1 > l2 ? l1 - l2 : l2 - l1;
	if (w < d) w = d;
	return w;
}

B200_HD int fin_first_band(const mem_opt_t &opt, const Reg *ar)                // reference src/bwamem.c:1107-1111
{
	const int tmp = fin_infer_bw(ar->qe - ar->qb, (int)(ar->re - ar->rb), ar->truesc, opt.a, opt.o_del, opt.e_del);
	int w2 = fin_infer_bw(ar->qe - ar->qb, (int)(ar->re - ar->rb), ar->truesc, opt.a, opt.o_ins, opt.e_ins);
	w2 = w2 > tmp ? w2 : tmp;
	if (w2 > opt.w) w2 = w2 < ar->w ? w2 : ar->w;
	return w2;
}

// the alignment of one flagged region, final form (mem_aln_t minus the per-record fields)
struct AlnRes {
	int64_t pos;            // position on the contig, 0-based
	int64_t cig_off;        // first operation in the CIGAR arena (room in front for the 5' clip)
	int64_t md_off;         // MD text in the MD arena
	int32_t rid, n_cigar, NM, md_len;
	int32_t is_rev, ref_len;        // ref_len: reference bases the CIGAR spans (TLEN)
};

// does the region go through the banded DP (else: equal lengths and an inferred band of 0, one M operation)?
B200_HD bool aln_needs_dp(const FinCtx &cx, const Reg *ar, int *w2_)
{
	const int w2 = fin_first_band(cx.opt, ar);
	*w2_ = w2;
	const int64_t l_pac = cx.fm.l_pac;
	if (!(ar->qe > ar->qb && ar->re <= l_pac << 1 && ar->rb < ar->re && !(ar->rb < l_pac && ar->re > l_pac))) return false;
	return global_needs_dp(ar->qe - ar->qb, ar->re - ar->rb, w2 < cx.opt.w << 2 ? w2 : cx.opt.w << 2);
}

// capacity, in operations, of a region's CIGAR strip: one per base of query and reference plus the two clips
B200_HD int64_t aln_cigar_cap(const Reg *ar) { return (int64_t)(ar->qe - ar->qb) + (ar->re - ar->rb) + 4; }
// capacity of its MD strip: a number (as many characters as matches it counts, or "0") and a letter per reference base, a number
// and '^' per operation
B200_HD int64_t aln_md_cap(const Reg *ar) { return 2 * (ar->re - ar->rb) + 2 * aln_cigar_cap(ar) + 16; }

template <class SINK>
B200_HD void fin_put_uint(SINK &s, uint64_t x)
{
	char buf[20];
	int l = 0;
	do { buf[l++] = (char)('0' + x % 10); x /= 10; } while (x);
	while (l > 0) s.put(buf[--l]);
}
template <class SINK>
B200_HD void fin_put_int(SINK &s, int64_t c)
{
	if (c < 0) { s.put('-'); fin_put_uint(s, (uint64_t)(-c)); }
	else fin_put_uint(s, (uint64_t)c);
}

struct MdSink { char *p; int n; B200_HD void put(char c) { p[n++] = c; } };

// NM / MD of bwa_gen_cigar2 (reference src/bwa.c:167-207) and the tail of mem_reg2aln (src/bwamem.c:1123-1159).
// cig: the n_cigar operations ksw_global2 produced (or the single M of the gap-free path), stored at arena[cig_off + 1 ..] so
// that the 5' clip fits in front.
B200_HDN void aln_finish(const FinCtx &cx, const Reg *ar, int l_query, const uint8_t *query, int n_cigar, uint32_t *arena, int64_t cig_slot,
                        char *md_arena, int64_t md_off, AlnRes *out)
{
	const int64_t l_pac = cx.fm.l_pac;
	uint32_t *cig = arena + cig_slot + 1;
	GlobalJob jb;
	jb.rb = ar->rb; jb.re = ar->re; jb.qb = ar->qb; jb.qe = ar->qe;
	const GlobalSeqs s = global_seqs(cx.fm.pac, l_pac, query + ar->qb, jb);
	int NM = -1;
	MdSink md = { md_arena + md_off, 0 };
	if (n_cigar > 0) {
		int x = 0, y = 0, u = 0, n_mm = 0, n_gap = 0;
		const char *int2base = ar->rb < l_pac ? "ACGTN" : "TGCAN";
		for (int k = 0; k < n_cigar; ++k) {
			const int op = cig[k] & 0xf, len = (int)(cig[k] >> 4);
			if (op == 0) {
				for (int i = 0; i < len; ++i) {
					const int t = s.ta(y + i);
					if (s.qa(x + i) != t) { fin_put_uint(md, (uint64_t)u); md.put(int2base[t]); ++n_mm; u = 0; }
					else ++u;
				}
				x += len; y += len;
			} else if (op == 2) {
				if (k > 0 && k < n_cigar - 1) {
					fin_put_uint(md, (uint64_t)u); md.put('^');
					for (int i = 0; i < len; ++i) md.put(int2base[s.ta(y + i)]);
					u = 0; n_gap += len;
				}
				y += len;
			} else if (op == 1) { x += len; n_gap += len; }
		}
		fin_put_uint(md, (uint64_t)u);
		NM = n_mm + n_gap;
	}
	int is_rev;
	int64_t pos = fm_depos(cx.fm, ar->rb < l_pac ? ar->rb : ar->re - 1, &is_rev);
	int first = 0;
	if (n_cigar > 0) {
		if ((cig[0] & 0xf) == 2) { pos += cig[0] >> 4; first = 1; --n_cigar; }
		else if ((cig[n_cigar - 1] & 0xf) == 2) --n_cigar;
	}
	int64_t c0 = cig_slot + 1 + first;
	if (ar->qb != 0 || ar->qe != l_query) {
		const int clip5 = is_rev ? l_query - ar->qe : ar->qb;
		const int clip3 = is_rev ? ar->qb : l_query - ar->qe;
		if (clip5) { --c0; arena[c0] = (uint32_t)clip5 << 4 | 3; ++n_cigar; }
		if (clip3) { arena[c0 + n_cigar] = (uint32_t)clip3 << 4 | 3; ++n_cigar; }
	}
	int ref_len = 0;
	for (int k = 0; k < n_cigar; ++k) { const int op = arena[c0 + k] & 0xf; if (op == 0 || op == 2) ref_len += (int)(arena[c0 + k] >> 4); }
	out->rid = fm_pos2rid(cx.fm, pos);
	out->pos = pos - cx.fm.ctg_off[out->rid];
	out->cig_off = c0; out->n_cigar = n_cigar; out->NM = (int32_t)((uint32_t)NM & 0x3fffff); out->md_off = md_off; out->md_len = md.n;
	out->is_rev = is_rev; out->ref_len = ref_len;
}

/* ---------------------------------------------------------------- SAM text (mem_aln2sam) */

struct CountSink {
	int64_t n = 0;
	B200_HD void put(char) { ++n; }
	B200_HD void puts(const char *, int64_t l) { n += l; }
};

// byte writer that gathers eight bytes per store once the destination is 8-byte aligned
struct WriteSink {
	char *p;
	uint64_t acc = 0; int na = 0;
	B200_HD explicit WriteSink(char *dst) : p(dst) {}
	B200_HD void put(char c)
	{
		if (na == 0 && ((uintptr_t)p & 7)) { *p++ = c; return; }
		acc |= (uint64_t)(uint8_t)c << (na << 3);
		if (++na == 8) { *reinterpret_cast<uint64_t *>(p) = acc; p += 8; acc = 0; na = 0; }
	}
	B200_HD void puts(const char *s, int64_t l) { for (int64_t i = 0; i < l; ++i) put(s[i]); }
	B200_HD void flush() { for (int i = 0; i < na; ++i) p[i] = (char)(acc >> (i << 3)); p += na; acc = 0; na = 0; }
};

template <class SINK> B200_HD void fin_put_lit(SINK &s, const char *lit) { for (; *lit; ++lit) s.put(*lit); }

// "%.3f" of (double)num / den for the pa:f tag: the exact binary value rounded to three decimals, half to even (glibc)
template <class SINK>
B200_HD void fin_put_ratio3(SINK &s, int num, int den)
{
	const double q = (double)num / den;
	union { double d; uint64_t u; } cv;
	cv.d = q < 0 ? -q : q;
	const int e = (int)(cv.u >> 52 & 0x7ff);
	uint64_t m = cv.u & 0xfffffffffffffull;
	uint64_t ip;                                 // round(|q| * 1000)
	if (e == 0) ip = 0;                          // zero or subnormal
	else {
		m |= 1ull << 52;
		const int k = 1075 - e;                  // |q| = m / 2^k
		if (k <= 0) ip = 0xffffffffffffffffull;  // (beyond 2^52: not reachable with int scores)
		else if (k >= 64) ip = 0;                // |q| < 2^-11: rounds to 0.000
		else {
			// m * 1000 as a 128-bit product (m < 2^53, so the high word is tiny)
			const uint64_t lo = m * 1000ull;
			const uint64_t hi = (uint64_t)((((m >> 32) * 1000ull) + (((m & 0xffffffffull) * 1000ull) >> 32)) >> 32);
			const uint64_t whole = k == 0 ? lo : (lo >> k) | (k < 64 && hi ? hi << (64 - k) : 0);
			const uint64_t rem = lo & ((1ull << k) - 1), half = 1ull << (k - 1);
			ip = whole + ((rem > half || (rem == half && (whole & 1))) ? 1 : 0);
		}
	}
	if (q < 0 && ip) s.put('-');
	fin_put_uint(s, ip / 1000);
	s.put('.');
	s.put((char)('0' + ip / 100 % 10)); s.put((char)('0' + ip / 10 % 10)); s.put((char)('0' + ip % 10));
}

struct SamView {            // what the formatter reads; all arrays live in HBM
	const Reg *regs; const int64_t *roff; const int32_t *nreg;       // final regions of read r: regs[roff[r] .. roff[r] + nreg[r])
	const int32_t *xa_of; const int32_t *aln_slot;                   // per region slot: XA owner, index into aln[] (-1: none)
	const AlnRes *aln; const uint32_t *cig; const char *md;
	const SamRec *recs; const int32_t *nrec; const int32_t *mate_reg; // records of read r at recs[roff[r] + r ..]; mate_reg: -2 = single-end
};

template <class SINK>
B200_HD void fin_put_cigar(const FinCtx &cx, SINK &s, const SamView &V, const AlnRes *p, int n_cigar, int is_alt, int which)
{
	if (!n_cigar) { s.put('*'); return; }
	for (int i = 0; i < n_cigar; ++i) {
		const uint32_t v = V.cig[p->cig_off + i];
		int c = v & 0xf;
		if (!(cx.opt.flag & MEM_F_SOFTCLIP) && !is_alt && (c == 3 || c == 4)) c = which ? 4 : 3;
		fin_put_uint(s, v >> 4); s.put("MIDSH"[c]);
	}
}

template <class SINK>
B200_HD void fin_put_ctg(const FinCtx &cx, SINK &s, int rid) { s.puts(cx.ctg_names + cx.ctg_name_off[rid], cx.ctg_name_off[rid + 1] - cx.ctg_name_off[rid]); }

// one SAM line: record `which` of read r (mem_aln2sam, reference src/bwamem.c:825-946).  *rid_out / *mrid_out: the contigs printed
// as RNAME and RNEXT (-1: '*') - the two fields the per-chromosome hosts parse back out of every line to route it
// (reference src/mainParallelByChromosome.c:1395-1457).
template <class SINK>
B200_HDN void sam_format(const FinCtx &cx, const SamView &V, int64_t r, int which, SINK &s, int *rid_out = nullptr, int *mrid_out = nullptr)
{
	const mem_opt_t &opt = cx.opt;
	const int64_t base = V.roff[r];
	const SamRec *list = V.recs + base + r;
	const int n_list = V.nrec[r];
	const SamRec rc = list[which];
	const Reg *reg = rc.reg >= 0 ? &V.regs[base + rc.reg] : nullptr;
	const AlnRes *pa = rc.reg >= 0 ? &V.aln[V.aln_slot[base + rc.reg]] : nullptr;
	const int mreg = V.mate_reg[r];
	const bool has_mate = mreg != -2;
	const int64_t mr = r ^ 1;
	const AlnRes *ma = (has_mate && mreg >= 0) ? &V.aln[V.aln_slot[V.roff[mr] + mreg]] : nullptr;
	const int m_is_alt = (has_mate && mreg >= 0) ? V.regs[V.roff[mr] + mreg].is_alt : 0;
	// the copies mem_aln2sam edits
	int p_rid = pa ? pa->rid : -1, p_ncig = pa ? pa->n_cigar : 0, p_rev = pa ? pa->is_rev : 0;
	int64_t p_pos = pa ? pa->pos : -1;
	int m_rid = ma ? ma->rid : -1, m_ncig = ma ? ma->n_cigar : 0, m_rev = ma ? ma->is_rev : 0;
	int64_t m_pos = ma ? ma->pos : -1;
	const int p_is_alt = reg ? reg->is_alt : 0;
	int flag = rc.flag;
	flag |= has_mate ? 0x1 : 0;
	flag |= p_rid < 0 ? 0x4 : 0;
	flag |= has_mate && m_rid < 0 ? 0x8 : 0;
	if (p_rid < 0 && has_mate && m_rid >= 0) { p_rid = m_rid; p_pos = m_pos; p_rev = m_rev; p_ncig = 0; }
	if (has_mate && m_rid < 0 && p_rid >= 0) { m_rid = p_rid; m_pos = p_pos; m_rev = p_rev; m_ncig = 0; }
	flag |= p_rev ? 0x10 : 0;
	flag |= has_mate && m_rev ? 0x20 : 0;

	if (rid_out) { *rid_out = p_rid; *mrid_out = has_mate ? m_rid : -1; }
	const ReadText rt = cx.rtext[r];
	s.puts(cx.text + rt.name_off, rt.name_len); s.put('\t');
	fin_put_int(s, (flag & 0xffff) | (flag & 0x10000 ? 0x100 : 0)); s.put('\t');
	if (p_rid >= 0) {
		fin_put_ctg(cx, s, p_rid); s.put('\t');
		fin_put_int(s, p_pos + 1); s.put('\t');
		fin_put_int(s, rc.mapq); s.put('\t');
		fin_put_cigar(cx, s, V, pa, p_ncig, p_is_alt, which);
	} else fin_put_lit(s, "*\t0\t0\t*");
	s.put('\t');

	if (has_mate && m_rid >= 0) {
		if (p_rid == m_rid) s.put('=');
		else fin_put_ctg(cx, s, m_rid);
		s.put('\t');
		fin_put_int(s, m_pos + 1); s.put('\t');
		if (p_rid == m_rid) {
			const int64_t p0 = p_pos + (p_rev ? (p_ncig ? pa->ref_len : 0) - 1 : 0);
			const int64_t p1 = m_pos + (m_rev ? (m_ncig ? ma->ref_len : 0) - 1 : 0);
			if (m_ncig == 0 || p_ncig == 0) s.put('0');
			else fin_put_int(s, -(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0)));
		} else s.put('0');
	} else fin_put_lit(s, "*\t0\t0");
	s.put('\t');

	const int l_seq = (int)(cx.off[r + 1] - cx.off[r]);
	const uint8_t *seq = cx.codes + cx.off[r];
	const char *qual = rt.qual_off >= 0 ? cx.text + rt.qual_off : nullptr;
	if (flag & 0x100) {
		fin_put_lit(s, "*\t*");
	} else {
		int qb = 0, qe = l_seq;
		if (p_ncig && which && !(opt.flag & MEM_F_SOFTCLIP) && !p_is_alt) {
			const uint32_t c0 = V.cig[pa->cig_off], c1 = V.cig[pa->cig_off + p_ncig - 1];
			if (!p_rev) {
				if ((c0 & 0xf) == 4 || (c0 & 0xf) == 3) qb += c0 >> 4;
				if ((c1 & 0xf) == 4 || (c1 & 0xf) == 3) qe -= c1 >> 4;
			} else {
				if ((c0 & 0xf) == 4 || (c0 & 0xf) == 3) qe -= c0 >> 4;
				if ((c1 & 0xf) == 4 || (c1 & 0xf) == 3) qb += c1 >> 4;
			}
		}
		if (!p_rev) {
			for (int i = qb; i < qe; ++i) s.put("ACGTN"[seq[i]]);
			s.put('\t');
			if (qual) { if (qe > qb) s.puts(qual + qb, qe - qb); }
			else s.put('*');
		} else {
			for (int i = qe - 1; i >= qb; --i) s.put("TGCAN"[seq[i]]);
			s.put('\t');
			if (qual) { for (int i = qe - 1; i >= qb; --i) s.put(qual[i]); }
			else s.put('*');
		}
	}

	if (p_ncig) {
		fin_put_lit(s, "\tNM:i:"); fin_put_int(s, pa->NM);
		fin_put_lit(s, "\tMD:Z:"); s.puts(V.md + pa->md_off, pa->md_len);
	}
	if (has_mate && m_ncig) { fin_put_lit(s, "\tMC:Z:"); fin_put_cigar(cx, s, V, ma, m_ncig, m_is_alt, which); }
	const int score = reg ? reg->score : 0;
	if (score >= 0) { fin_put_lit(s, "\tAS:i:"); fin_put_int(s, score); }
	if (rc.sub >= 0) { fin_put_lit(s, "\tXS:i:"); fin_put_int(s, rc.sub); }
	if (cx.rg_len) { fin_put_lit(s, "\tRG:Z:"); s.puts(cx.rg_id, cx.rg_len); }
	if (!(flag & 0x100)) {
		int i;
		for (i = 0; i < n_list; ++i)
			if (i != which && !(list[i].flag & 0x100)) break;
		if (i < n_list) {
			fin_put_lit(s, "\tSA:Z:");
			for (i = 0; i < n_list; ++i) {
				const SamRec &o = list[i];
				if (i == which || (o.flag & 0x100) || o.reg < 0) continue;
				const AlnRes *oa = &V.aln[V.aln_slot[base + o.reg]];          // (a read with two records has no unmapped one)
				fin_put_ctg(cx, s, oa->rid); s.put(',');
				fin_put_int(s, oa->pos + 1); s.put(',');
				s.put("+-"[oa->is_rev]); s.put(',');
				for (int k = 0; k < oa->n_cigar; ++k) { const uint32_t c = V.cig[oa->cig_off + k]; fin_put_uint(s, c >> 4); s.put("MIDSH"[c & 0xf]); }
				s.put(','); fin_put_int(s, o.mapq);
				s.put(','); fin_put_int(s, oa->NM);
				s.put(';');
			}
		}
		const int alt_sc = reg ? reg->alt_sc : 0;
		if (alt_sc > 0) { fin_put_lit(s, "\tpa:f:"); fin_put_ratio3(s, score, alt_sc); }
	}
	if (rc.reg >= 0 && !(opt.flag & MEM_F_ALL)) {           // XA: the members planned by fin_plan_xa, in region order
		bool any = false;
		const int n = V.nreg[r];
		for (int i = 0; i < n; ++i) {
			if (V.xa_of[base + i] != rc.reg) continue;
			if (!any) { fin_put_lit(s, "\tXA:Z:"); any = true; }
			const AlnRes *t = &V.aln[V.aln_slot[base + i]];
			fin_put_ctg(cx, s, t->rid);
			s.put(','); s.put("+-"[t->is_rev]); fin_put_int(s, t->pos + 1);
			s.put(',');
			for (int k = 0; k < t->n_cigar; ++k) { const uint32_t c = V.cig[t->cig_off + k]; fin_put_uint(s, c >> 4); s.put("MIDSHN"[c & 0xf]); }
			s.put(','); fin_put_int(s, t->NM);
			s.put(';');
		}
	}
	if (rt.comment_off >= 0) { s.put('\t'); s.puts(cx.text + rt.comment_off, rt.comment_len); }
	if ((opt.flag & MEM_F_REF_HDR) && p_rid >= 0 && cx.ctg_anno_off[p_rid + 1] > cx.ctg_anno_off[p_rid]) {
		fin_put_lit(s, "\tXR:Z:");
		const char *an = cx.ctg_annos + cx.ctg_anno_off[p_rid];
		const int64_t l = cx.ctg_anno_off[p_rid + 1] - cx.ctg_anno_off[p_rid];
		for (int64_t i = 0; i < l; ++i) s.put(an[i] == '\t' ? ' ' : an[i]);
	}
	s.put('\n');
}

// all records of read r
template <class SINK>
B200_HD void sam_format_read(const FinCtx &cx, const SamView &V, int64_t r, SINK &s)
{
	const int n = V.nrec[r];
	for (int w = 0; w < n; ++w) sam_format(cx, V, r, w, s);
}

// one entry per SAM line of the chunk, in output order: where the line is in the text and which contigs it names
struct SamLine { int64_t off; int32_t len, rid, mate_rid, read; };

// mates must carry the same name (reference src/bwamem_pair.c:360)
B200_HD bool fin_names_differ(const FinCtx &cx, int64_t r0)
{
	const ReadText a = cx.rtext[r0], b = cx.rtext[r0 + 1];
	if (a.name_len != b.name_len) return true;
	for (int i = 0; i < a.name_len; ++i) if (cx.text[a.name_off + i] != cx.text[b.name_off + i]) return true;
	return false;
}

} // namespace b200
