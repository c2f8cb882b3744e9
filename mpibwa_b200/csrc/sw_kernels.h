// sw_kernels.h - per-thread task body of the local Smith-Waterman used by mate rescue (and seed filtering).
//
// Scalar restatement of the OBSERVABLE behaviour of reference src/ksw.c:63-365 (ksw_qinit / ksw_u8 / ksw_i16 /
// ksw_align2), rules in SURVEY.md A.2:
//   * the query is padded to a multiple of 16 (8-bit scores) or 8 (16-bit scores) columns that score 0 and take
//     part in the row maximum;
//   * the first gap costs o+e and opens from H; E,F,H are clamped at 0;
//   * 8-bit mode works on values biased by `shift` that saturate at 255;
//   * te = first row reaching the best score (strict >), qe = smallest query index attaining it in that row;
//   * rows whose maximum is >= minsc feed the merged-run list that gives score2/te2;
//   * KSW_XSTART runs a second pass over the reversed prefixes with KSW_XSTOP|score.
// Equivalence of the plain row order used here with the striped lazy-F order of the reference holds for gap-open
// penalties >= 1 (the reference's own early exit assumes it) and is fuzz-checked against oracle/_ref in tests.
#pragma once
#include <cstdint>
#include "fm_kernels.h"

namespace b200 {

struct SwOpt {
	int o_del, e_del, o_ins, e_ins;
	int max_sc, shift;      // largest entry of mat; -smallest entry (8-bit bias)
	int8_t mat[25];
};

struct SwRes { int score, te, qe, score2, te2, tb, qb; };

// smallest strip width (query columns per lane) of the one-warp-per-job kernel (sw_warp_kernel.cuh) that holds the job, or 0
// when it must take the general (one thread per job) kernel
B200_HD int sw_warp_class(int qlen, int xtra)
{
	const int p = (xtra & 0x10000) ? 16 : 8;
	const int qpad = (qlen + p - 1) / p * p;
	return qpad <= 64 ? 2 : qpad <= 128 ? 4 : qpad <= 160 ? 5 : qpad <= 256 ? 8 : 0;
}

struct Row16 {
	uint16_t *base; int64_t stride;
	B200_HD int get(int q) const { return base[(int64_t)q * stride]; }
	B200_HD void set(int q, int v) const { base[(int64_t)q * stride] = (uint16_t)v; }
};
struct List64 {
	uint64_t *base; int64_t stride;
	B200_HD uint64_t get(int k) const { return base[(int64_t)k * stride]; }
	B200_HD void set(int k, uint64_t v) const { base[(int64_t)k * stride] = v; }
};

// query accessors returning codes 0..4
struct SQFwd { const uint8_t *p; B200_HD int operator()(int j) const { return p[j]; } };
struct SQRevComp { const uint8_t *p; int l; B200_HD int operator()(int j) const { int c = p[l - 1 - j]; return c < 4 ? 3 - c : 4; } };
template <class QA> struct SQFlip { QA q; int last; B200_HD int operator()(int j) const { return q(last - j); } };
template <class TA> struct STFlip { TA t; int te; B200_HD int operator()(int i) const { return i <= te ? t(te - i) : t(i); } };
struct STBytes { const uint8_t *p; B200_HD int operator()(int i) const { return p[i]; } };
struct STPac { const uint8_t *pac; int64_t l_pac, beg; B200_HD int operator()(int i) const { return fm_base(pac, l_pac, beg + i); } };

// one pass; size = 1 (8-bit semantics) or 2 (16-bit semantics)
template <class QA, class TA>
B200_HDN void sw_pass(int qlen, QA query, int tlen, TA target, const SwOpt &o, int size, int minsc, int endsc,
                      Row16 H, Row16 E, List64 b, SwRes *r, int64_t *cells)
{
	const int p = size == 1 ? 16 : 8;
	const int slen = (qlen + p - 1) / p, qpad = slen * p;
	const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
	const int cap = 255 - o.shift;            // largest representable 8-bit score
	int gmax = 0, te = -1, qe = 0, n_b = 0, i, q;
	r->score = 0; r->te = -1; r->qe = -1; r->score2 = -1; r->te2 = -1; r->tb = -1; r->qb = -1;
	for (q = 0; q < qpad; ++q) { H.set(q, 0); E.set(q, 0); }
	int64_t rows = 0;
	for (i = 0; i < tlen; ++i) {
		const int8_t *mrow = o.mat + target(i) * 5;
		int f = 0, diag = 0, imax = 0, iq = 0;
		for (q = 0; q < qpad; ++q) {
			int h = diag + (q < qlen ? mrow[query(q)] : 0);
			int e = E.get(q);
			if (size == 1) { if (h > cap) h = cap; }
			if (h < 0) h = 0;
			h = h > e ? h : e;
			h = h > f ? h : f;
			diag = H.get(q);
			H.set(q, h);
			if (h > imax) { imax = h; iq = q; }
			int t = h - oe_del; t = t > 0 ? t : 0;
			e -= o.e_del; e = e > t ? e : t;
			E.set(q, e);
			t = h - oe_ins; t = t > 0 ? t : 0;
			f -= o.e_ins; f = f > t ? f : t;
		}
		++rows;
		if (imax >= minsc) {
			if (n_b == 0 || (int32_t)b.get(n_b - 1) + 1 != i) b.set(n_b++, (uint64_t)imax << 32 | (uint32_t)i);
			else if ((int)(b.get(n_b - 1) >> 32) < imax) b.set(n_b - 1, (uint64_t)imax << 32 | (uint32_t)i);
		}
		if (imax > gmax) {
			gmax = imax; te = i; qe = iq;
			if (size == 1) { if (gmax + o.shift >= 255 || gmax >= endsc) break; }
			else if (gmax >= endsc) break;
		}
	}
	if (cells) *cells += rows * qpad;
	if (size == 1) r->score = gmax + o.shift < 255 ? gmax : 255;
	else r->score = gmax;
	r->te = te;
	if (size == 2 || r->score != 255) {
		r->qe = qe;
		if (n_b > 0) {
			int d = (r->score + o.max_sc - 1) / o.max_sc;
			int low = te - d, high = te + d;
			for (int k = 0; k < n_b; ++k) {
				uint64_t v = b.get(k);
				int e = (int32_t)v;
				if ((e < low || e > high) && (int)(v >> 32) > r->score2) { r->score2 = (int)(v >> 32); r->te2 = e; }
			}
		}
	}
}

// ksw_align2 with qry == NULL
template <class QA, class TA>
B200_HDN void sw_align(int qlen, QA query, int tlen, TA target, const SwOpt &o, int xtra,
                       Row16 H, Row16 E, List64 b, SwRes *r, int64_t *cells)
{
	const int size = (xtra & 0x10000) ? 1 : 2;
	const int minsc = (xtra & 0x40000) ? (xtra & 0xffff) : 0x10000;
	const int endsc = (xtra & 0x20000) ? (xtra & 0xffff) : 0x10000;
	sw_pass(qlen, query, tlen, target, o, size, minsc, endsc, H, E, b, r, cells);
	if ((xtra & 0x80000) == 0 || ((xtra & 0x40000) && r->score < (xtra & 0xffff))) return;
	if (r->qe < 0) return;                    // 8-bit overflow: the reference is undefined from here on
	SwRes rr;
	SQFlip<QA> q2 = { query, r->qe };
	STFlip<TA> t2 = { target, r->te };
	sw_pass(r->qe + 1, q2, tlen, t2, o, size, 0x10000, r->score & 0xffff, H, E, b, &rr, cells);
	if (r->score == rr.score) { r->tb = r->te - rr.te; r->qb = r->qe - rr.qe; }
}

} // namespace b200
