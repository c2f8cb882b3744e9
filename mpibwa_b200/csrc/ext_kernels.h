// ext_kernels.h - per-thread task bodies of the seed-extension stage.
//
//   extend_core     banded affine-gap extension with BWA's z-drop and adaptive band; bit-exact restatement of the
//                   recurrences of reference src/ksw.c:380-479 (see SURVEY.md A.1 for the rules that bite:
//                   a zero diagonal cannot restart, gaps open from M, ties in the row maximum go to the larger j,
//                   stale cells outside the shrinking band keep their old values).
//   chain2aln_need_extension   the per-seed containment test of reference src/bwamem.c:671-706 (the device walks a read's
//                   chains and seeds in the reference's order in ext_rounds.cuh: the test of seed k needs the regions of
//                   the earlier seeds).
//
// The DP state row is accessed through an accessor object so that the same code runs over shared memory,
// interleaved global scratch or a plain host array.
#pragma once
#include <cstdint>
#include "fm_kernels.h"

namespace b200 {

struct ExtOpt {
	int a, b, o_del, e_del, o_ins, e_ins, w, zdrop, pen_clip5, pen_clip3;
	int max_sc;             // largest entry of mat
	int8_t mat[25];
};

struct DSeed { int64_t rbeg; int32_t qbeg, len, score, pad; };

struct DChain {
	int64_t rmax0, rmax1;   // reference window after contig clipping (bns_fetch_seq semantics)
	int32_t seed_beg, n_seeds;
	int32_t rid;
	float frac_rep;
};

struct DReg {               // the fields mem_chain2aln writes; everything else of mem_alnreg_t is zero
	int64_t rb, re;
	int32_t qb, qe;
	int32_t rid, score, truesc, w;
	int32_t seedcov, seedlen0;
	float frac_rep;
	int32_t pad;
};

struct ExtOut { int score, qle, tle, gtle, gscore, max_off; };

// query accessors: codes 0..4
struct QFwd { const uint8_t *p; B200_HD int operator()(int j) const { return p[j]; } };
struct QRev { const uint8_t *p; B200_HD int operator()(int j) const { return p[-j]; } };      // p points at the first base
// target accessors
struct TBytes { const uint8_t *p; B200_HD int operator()(int i) const { return p[i]; } };
struct TPacFwd { const uint8_t *pac; int64_t l_pac, beg; B200_HD int operator()(int i) const { return fm_base(pac, l_pac, beg + i); } };
struct TPacRev { const uint8_t *pac; int64_t l_pac, beg; B200_HD int operator()(int i) const { return fm_base(pac, l_pac, beg - i); } };

// plain {h,e} row in memory with a stride (stride 1 = host array, stride = #threads = interleaved scratch)
struct EhStrided {
	int32_t *base; int64_t stride;
	B200_HD void get(int j, int &h, int &e) const { const int32_t *p = base + (int64_t)j * 2 * stride; h = p[0]; e = p[stride]; }
	B200_HD void set_h(int j, int h) const { base[(int64_t)j * 2 * stride] = h; }
	B200_HD void set_e(int j, int e) const { base[(int64_t)j * 2 * stride + stride] = e; }
	B200_HD void set(int j, int h, int e) const { int32_t *p = base + (int64_t)j * 2 * stride; p[0] = h; p[stride] = e; }
};

template <class QA, class TA, class EH>
B200_HDN void extend_core(int qlen, QA query, int tlen, TA target, const ExtOpt &o, int w, int end_bonus, int h0,
                          EH eh, ExtOut *out, int64_t *cells)
{
	const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
	int i, j, beg, end, max, max_i, max_j, max_ie, gscore, max_off;
	// row -1
	eh.set(0, h0, 0);
	if (qlen >= 1) eh.set(1, h0 > oe_ins ? h0 - oe_ins : 0, 0);
	{
		int prev = h0 > oe_ins ? h0 - oe_ins : 0;
		for (j = 2; j <= qlen && prev > o.e_ins; ++j) { prev -= o.e_ins; eh.set(j, prev, 0); }
		for (; j <= qlen; ++j) eh.set(j, 0, 0);
	}
	// band clamp
	{
		int max_ins = (int)((double)(qlen * o.max_sc + end_bonus - o.o_ins) / o.e_ins + 1.);
		max_ins = max_ins > 1 ? max_ins : 1;
		w = w < max_ins ? w : max_ins;
		int max_del = (int)((double)(qlen * o.max_sc + end_bonus - o.o_del) / o.e_del + 1.);
		max_del = max_del > 1 ? max_del : 1;
		w = w < max_del ? w : max_del;
	}
	max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
	beg = 0; end = qlen;
	int64_t ncell = 0;
	for (i = 0; i < tlen; ++i) {
		int f = 0, h1, m = 0, mj = -1;
		const int8_t *mrow = o.mat + target(i) * 5;
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		if (beg == 0) { h1 = h0 - (o.o_del + o.e_del * (i + 1)); if (h1 < 0) h1 = 0; }
		else h1 = 0;
		for (j = beg; j < end; ++j) {
			int M, e, h, t;
			eh.get(j, M, e);
			eh.set_h(j, h1);
			M = M ? M + mrow[query(j)] : 0;
			h = M > e ? M : e;
			h = h > f ? h : f;
			h1 = h;
			mj = m > h ? mj : j;
			m = m > h ? m : h;
			t = M - oe_del; t = t > 0 ? t : 0;
			e -= o.e_del; e = e > t ? e : t;
			eh.set_e(j, e);
			t = M - oe_ins; t = t > 0 ? t : 0;
			f -= o.e_ins; f = f > t ? f : t;
		}
		if (end > beg) ncell += end - beg;
		eh.set(end, h1, 0);
		if (j == qlen) {
			max_ie = gscore > h1 ? max_ie : i;
			gscore = gscore > h1 ? gscore : h1;
		}
		if (m == 0) break;
		if (m > max) {
			max = m; max_i = i; max_j = mj;
			int d = mj - i; d = d < 0 ? -d : d;
			max_off = max_off > d ? max_off : d;
		} else if (o.zdrop > 0) {
			if (i - max_i > mj - max_j) {
				if (max - m - ((i - max_i) - (mj - max_j)) * o.e_del > o.zdrop) break;
			} else {
				if (max - m - ((mj - max_j) - (i - max_i)) * o.e_ins > o.zdrop) break;
			}
		}
		int hh, ee;
		for (j = beg; j < end; ++j) { eh.get(j, hh, ee); if (hh != 0 || ee != 0) break; }
		beg = j;
		for (j = end; j >= beg; --j) { eh.get(j, hh, ee); if (hh != 0 || ee != 0) break; }
		end = j + 2 < qlen ? j + 2 : qlen;
	}
	out->score = max; out->qle = max_j + 1; out->tle = max_i + 1; out->gtle = max_ie + 1;
	out->gscore = gscore; out->max_off = max_off;
	if (cells) *cells += ncell;
}

B200_HD int cal_max_gap(const ExtOpt &o, int qlen)
{
	int l_del = (int)((double)(qlen * o.a - o.o_del) / o.e_del + 1.);
	int l_ins = (int)((double)(qlen * o.a - o.o_ins) / o.e_ins + 1.);
	int l = l_del > l_ins ? l_del : l_ins;
	l = l > 1 ? l : 1;
	return l < o.w << 1 ? l : o.w << 1;
}

// Decide whether seed `k` (position in the score-sorted order `srt`) of chain c still needs extension given the
// regions regs[0..n_av) already made for this read.  Returns 1 to extend, 0 to skip (and marks srt[k] = -1).
B200_HDN int chain2aln_need_extension(const ExtOpt &o, int l_query, const DChain &c, const DSeed *seeds, int32_t *srt,
                                      int k, const DReg *regs, int n_av)
{
	const DSeed &s = seeds[srt[k]];
	int i;
	for (i = 0; i < n_av; ++i) {
		const DReg &p = regs[i];
		int64_t rd;
		int qd, w, max_gap;
		if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
		if (s.len - p.seedlen0 > .1 * l_query) continue;
		qd = s.qbeg - p.qb; rd = s.rbeg - p.rb;
		max_gap = cal_max_gap(o, qd < rd ? qd : (int)rd);
		w = max_gap < p.w ? max_gap : p.w;
		if (qd - rd < w && rd - qd < w) break;
		qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
		max_gap = cal_max_gap(o, qd < rd ? qd : (int)rd);
		w = max_gap < p.w ? max_gap : p.w;
		if (qd - rd < w && rd - qd < w) break;
	}
	if (i == n_av) return 1;
	for (i = k + 1; i < c.n_seeds; ++i) {
		if (srt[i] < 0) continue;
		const DSeed &t = seeds[srt[i]];
		if (t.len < s.len * .95) continue;
		if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
		if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
	}
	if (i == c.n_seeds) { srt[k] = -1; return 0; }
	return 1;
}

} // namespace b200
