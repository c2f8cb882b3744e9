// global_kernels.h - per-thread task body of the CIGAR stage: the alignment part of mem_reg2aln (reference
// src/bwamem.c:1107-1122: up to three band-doubling calls of bwa_gen_cigar2, src/bwa.c:121-166) down to ksw_global2
// (src/ksw.c:504-606): banded global affine-gap DP with a 6-bit direction matrix and its traceback.
//
// One device task = one region of one read.  The query is a slice of the read resident in HBM, the target comes
// straight from the 2-bit reference (both read backwards for reverse-strand regions, which is what the reference's
// in-place reversals amount to), the H/E row sits in thread-interleaved scratch and the direction matrix in a
// per-task strip of global memory.  The task returns the score and the CIGAR; NM and MD are derived from the CIGAR on
// the host (a 150-step compare loop).  Host/device code: tests/hostemu runs the same body on the CPU.
#pragma once
#include <cstdint>
#include "fm_kernels.h"

namespace b200 {


struct GlobalOpt {
	int o_del, e_del, o_ins, e_ins;
	int a;                  // match score (opt->a): "score < truesc - a" retry rule
	int w_max;              // opt->w << 2
	int8_t mat[25];
};

struct GlobalJob {
	int64_t rb, re;         // reference interval in the forward+reverse coordinate (one strand, inside one contig)
	int64_t zoff;           // offset of this task's direction-matrix strip
	int32_t read, qb, qe;   // query = read[qb, qe)
	int32_t w2;             // first band width to try (inferred from the region's score, src/bwamem.c:1107-1111)
	int32_t truesc;
	int32_t wmax;           // widest band any of the (up to three) tries can use - sizes the row buffer
	int64_t cig_off;        // where the CIGAR goes in the arena (room for qlen + tlen operations)
};

struct GlobalRes {
	int32_t score;
	int32_t n_cigar;        // -2 (device-internal): a retry needed a wider row window than the launch had - rerun in a wider class
	int32_t n_tries, pad;
};

#define B200_GLOBAL_MINUS_INF (-0x40000000)
#define B200_GLOBAL_RAW ((int32_t)0x80000000)     // GlobalJob.truesc: w2 is the band itself, no bwa_gen_cigar2 clamp, no retries

// band actually used by bwa_gen_cigar2 for a requested width w_ (reference src/bwa.c:151-160)
B200_HD int global_band(const GlobalOpt &o, int l_query, int rlen, int w_)
{
	int max_ins = (int)((double)(((l_query + 1) >> 1) * o.mat[0] - o.o_ins) / o.e_ins + 1.);
	int max_del = (int)((double)(((l_query + 1) >> 1) * o.mat[0] - o.o_del) / o.e_del + 1.);
	int max_gap = max_ins > max_del ? max_ins : max_del;
	max_gap = max_gap > 1 ? max_gap : 1;
	int d = rlen - l_query; d = d < 0 ? -d : d;
	int w = (max_gap + d + 1) >> 1;
	w = w < w_ ? w : w_;
	int min_w = d + 3;
	return w > min_w ? w : min_w;
}

// bytes of direction matrix the task may need: the widest of its (up to three) tries; *wmax = that band
B200_HD int64_t global_z_need(const GlobalOpt &o, int l_query, int rlen, int w2, int *wmax)
{
	int64_t need = 0;
	*wmax = 0;
	for (int i = 0; i < 3; ++i) {
		w2 = w2 < o.w_max ? w2 : o.w_max;
		const int w = global_band(o, l_query, rlen, w2);
		const int n_col = l_query < 2 * w + 1 ? l_query : 2 * w + 1;
		const int64_t b = (int64_t)((n_col + 3) & ~3) * rlen;      // rows of the direction matrix are padded to whole words
		need = need > b ? need : b;
		*wmax = *wmax > w ? *wmax : w;
		if (w2 == o.w_max) break;
		w2 <<= 1;
	}
	return need;
}

// does mem_reg2aln need a DP at all for this region?  (equal lengths and an inferred band of 0 take the no-gap path)
B200_HD bool global_needs_dp(int l_query, int64_t rlen, int w2) { return !(l_query == rlen && w2 == 0); }

struct GlobalRow {          // H/E row, int32, one column = two words at stride
	int32_t *base; int64_t stride;
	B200_HD int32_t h(int j) const { return base[(int64_t)j * 2 * stride]; }
	B200_HD int32_t e(int j) const { return base[(int64_t)j * 2 * stride + stride]; }
	B200_HD void set_h(int j, int32_t v) const { base[(int64_t)j * 2 * stride] = v; }
	B200_HD void set_e(int j, int32_t v) const { base[(int64_t)j * 2 * stride + stride] = v; }
};

struct GlobalSeqs {         // oriented views of query and target
	const uint8_t *q; int l_query;
	const uint8_t *pac; int64_t l_pac, rb, re; int rev;
	B200_HD int qa(int j) const { return rev ? q[l_query - 1 - j] : q[j]; }
	B200_HD int ta(int i) const { return fm_base(pac, l_pac, rev ? re - 1 - i : rb + i); }
	B200_HD const int8_t *trow(const GlobalOpt &o, int i) const { return o.mat + ta(i) * 5; }
	B200_HD int sub(const int8_t *row, int j) const { return row[qa(j)]; }
};

// ksw_global2 with traceback.  Returns the score; the CIGAR goes to cigar[] (room for qlen + tlen operations), its length to *n_cigar_.
// ROW: H/E row accessor; SEQ: oriented query/target accessor (qa, ta, l_query).  z must be 4-byte aligned.
template <class ROW, class SEQ>
B200_HDN int global_dp(const GlobalOpt &o, const SEQ &s, int tlen, int w, ROW eh, uint8_t *z, uint32_t *cigar, int *n_cigar_,
                       int64_t *cells)
{
	const int qlen = s.l_query;
	const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
	const int n_col = ((qlen < 2 * w + 1 ? qlen : 2 * w + 1) + 3) & ~3;   // row stride of the direction matrix
	int j;
	eh.set_h(0, 0); eh.set_e(0, B200_GLOBAL_MINUS_INF);
	for (j = 1; j <= qlen && j <= w; ++j) { eh.set_h(j, -(o.o_ins + o.e_ins * j)); eh.set_e(j, B200_GLOBAL_MINUS_INF); }
	// Columns beyond w need no initial value: column i+w is first read in row i, after row i-1 has written it as its
	// right edge (eh[end]); and the last row always reaches column qlen because the band is at least |tlen-qlen|+3 wide.
	// This is what lets the device keep only a band-wide circular window of the row.
	int64_t ncell = 0;
	for (int i = 0; i < tlen; ++i) {
		int32_t f = B200_GLOBAL_MINUS_INF, h1, t;
		const auto mrow = s.trow(o, i);                // substitution scores of target base i
		const int beg = i > w ? i - w : 0;
		const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
		h1 = beg == 0 ? -(o.o_del + o.e_del * (i + 1)) : B200_GLOBAL_MINUS_INF;
		uint32_t *zi = reinterpret_cast<uint32_t *>(z + (int64_t)i * n_col);
		uint32_t zacc = 0;
		for (j = beg; j < end; ++j) {
			int32_t h, m = eh.h(j), e = eh.e(j);
			uint8_t d;
			eh.set_h(j, h1);
			m += s.sub(mrow, j);
			d = m >= e ? 0 : 1;
			h = m >= e ? m : e;
			d = h >= f ? d : 2;
			h = h >= f ? h : f;
			h1 = h;
			t = m - oe_del;
			e -= o.e_del;
			d |= e > t ? 1 << 2 : 0;
			e = e > t ? e : t;
			eh.set_e(j, e);
			t = m - oe_ins;
			f -= o.e_ins;
			d |= f > t ? 2 << 4 : 0;
			f = f > t ? f : t;
			const int zk = j - beg;                    // four 6-bit direction codes per stored word
			zacc |= (uint32_t)d << ((zk & 3) << 3);
			if ((zk & 3) == 3) { zi[zk >> 2] = zacc; zacc = 0; }
		}
		if ((end - beg) & 3) zi[(end - beg) >> 2] = zacc;
		if (end > beg) ncell += end - beg;
		eh.set_h(end, h1); eh.set_e(end, B200_GLOBAL_MINUS_INF);
	}
	if (cells) *cells += ncell;
	const int score = eh.h(qlen);
	// traceback (operations come out last to first)
	int n = 0, which = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
	int last_op = -1;
	uint32_t cur = 0;                              // the operation being grown lives in a register
#define B200_PUSH(op_, len_) do { \
		if (last_op != (op_)) { if (last_op >= 0) cigar[n++] = cur; cur = (uint32_t)(len_) << 4 | (uint32_t)(op_); last_op = (op_); } \
		else cur += (uint32_t)(len_) << 4; } while (0)
	while (i >= 0 && k >= 0) {
		which = z[(int64_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
		if (which == 0) { B200_PUSH(0, 1); --i; --k; }
		else if (which == 1) { B200_PUSH(2, 1); --i; }
		else { B200_PUSH(1, 1); --k; }
	}
	if (i >= 0) B200_PUSH(2, i + 1);
	if (k >= 0) B200_PUSH(1, k + 1);
#undef B200_PUSH
	if (last_op >= 0) cigar[n++] = cur;
	for (int x = 0; x < n >> 1; ++x) { const uint32_t tmp = cigar[x]; cigar[x] = cigar[n - 1 - x]; cigar[n - 1 - x] = tmp; }
	*n_cigar_ = n;
	return score;
}

// the band-doubling loop of mem_reg2aln around bwa_gen_cigar2 (reference src/bwamem.c:1112-1122)
// max_band: widest band the row accessor can hold (0x7fffffff: any)
template <class ROW, class SEQ>
B200_HDN void global_task(const GlobalOpt &o, const SEQ &s, const GlobalJob &jb, ROW eh, uint8_t *z, uint32_t *cigar, GlobalRes *out, int64_t *cells,
                          int max_band = 0x7fffffff)
{
	const int l_query = s.l_query, rlen = (int)(jb.re - jb.rb);
	int w2 = jb.w2, last_sc = -(1 << 30), score = 0, n_cigar = 0, i = 0;
	if (jb.truesc == B200_GLOBAL_RAW) {             // one plain ksw_global2 call with the caller's band (b200_ksw_global2_batch)
		if (w2 > max_band) { out->score = 0; out->n_cigar = -2; out->n_tries = 0; out->pad = 0; return; }
		out->score = global_dp(o, s, rlen, w2, eh, z, cigar, &n_cigar, cells);
		out->n_cigar = n_cigar; out->n_tries = 1; out->pad = 0;
		return;
	}
	do {
		w2 = w2 < o.w_max ? w2 : o.w_max;
		if (l_query == rlen && w2 == 0) {
			cigar[0] = (uint32_t)l_query << 4;
			n_cigar = 1;
			score = 0;
			for (int x = 0; x < l_query; ++x) score += s.sub(s.trow(o, x), x);
		} else {
			const int w = global_band(o, l_query, rlen, w2);
			if (w > max_band) { out->score = 0; out->n_cigar = -2; out->n_tries = i; out->pad = 0; return; }
			score = global_dp(o, s, rlen, w, eh, z, cigar, &n_cigar, cells);
		}
		if (score == last_sc || w2 == o.w_max) { ++i; break; }
		last_sc = score;
		w2 <<= 1;
	} while (++i < 3 && score < jb.truesc - o.a);
	out->score = score; out->n_cigar = n_cigar; out->n_tries = i; out->pad = 0;
}

B200_HD GlobalSeqs global_seqs(const uint8_t *pac, int64_t l_pac, const uint8_t *q, const GlobalJob &jb)
{
	GlobalSeqs s;
	s.q = q; s.l_query = jb.qe - jb.qb; s.pac = pac; s.l_pac = l_pac; s.rb = jb.rb; s.re = jb.re; s.rev = jb.rb >= l_pac;
	return s;
}

} // namespace b200
