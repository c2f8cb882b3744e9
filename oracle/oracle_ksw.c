/* TEST INFRASTRUCTURE ONLY - see oracle.h.  Scalar restatements of the two Smith-Waterman kernels. */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* Banded seed extension, reference src/ksw.c:380-479.
 * Row state: H[j] = H(i-1, j-1) as seen by row i, E[j] = E(i, j).  Cells outside the (shrinking) band keep whatever
 * an earlier row left there, exactly like the reference's eh[] array. */
void orc_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t mat[25],
                     int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, orc_ext_t *out)
{
	int *H = calloc(qlen + 2, sizeof(int)), *E = calloc(qlen + 2, sizeof(int));
	int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int i, j, k, best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0, beg = 0, end = qlen, top = 0;
	int64_t cells = 0;
	/* first row (src/ksw.c:395-397) */
	H[0] = h0;
	if (qlen >= 1) H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
	for (j = 2; j <= qlen && H[j - 1] > e_ins; ++j) H[j] = H[j - 1] - e_ins;
	/* band clamp (src/ksw.c:399-407) */
	for (k = 0; k < 25; ++k) top = imax(top, mat[k]);
	w = imin(w, imax(1, (int)((double)(qlen * top + end_bonus - o_ins) / e_ins + 1.)));
	w = imin(w, imax(1, (int)((double)(qlen * top + end_bonus - o_del) / e_del + 1.)));
	for (i = 0; i < tlen; ++i) {
		int f = 0, left, rowmax = 0, rowmax_j = -1;
		const int8_t *s = mat + 5 * target[i];
		beg = imax(beg, i - w);
		end = imin(imin(end, i + w + 1), qlen);
		left = beg == 0 ? imax(0, h0 - (o_del + e_del * (i + 1))) : 0;          /* src/ksw.c:420-423 */
		for (j = beg; j < end; ++j) {
			int diag = H[j], e = E[j], m, h;
			H[j] = left;
			m = diag ? diag + s[query[j]] : 0;                                   /* a zero diagonal cannot restart */
			h = imax(imax(m, e), f);
			left = h;
			if (!(rowmax > h)) rowmax_j = j;                                     /* ties go to the larger j */
			rowmax = imax(rowmax, h);
			E[j] = imax(e - e_del, imax(m - oe_del, 0));                          /* gaps open from M */
			f = imax(f - e_ins, imax(m - oe_ins, 0));
		}
		if (end > beg) cells += end - beg;
		H[end] = left; E[end] = 0;
		if (j == qlen) {                                                          /* src/ksw.c:449-453: ties -> later row */
			if (!(gscore > left)) best_ie = i;
			gscore = imax(gscore, left);
		}
		if (rowmax == 0) break;
		if (rowmax > best) {
			best = rowmax; best_i = i; best_j = rowmax_j;
			max_off = imax(max_off, abs(rowmax_j - i));
		} else if (zdrop > 0) {                                                   /* src/ksw.c:458-464 */
			int di = i - best_i, dj = rowmax_j - best_j;
			if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
			else if (best - rowmax - (dj - di) * e_ins > zdrop) break;
		}
		for (j = beg; j < end && H[j] == 0 && E[j] == 0; ++j) {}                  /* src/ksw.c:466-469 */
		beg = j;
		for (j = end; j >= beg && H[j] == 0 && E[j] == 0; --j) {}
		end = imin(j + 2, qlen);
	}
	out->score = best; out->qle = best_j + 1; out->tle = best_i + 1; out->gtle = best_ie + 1;
	out->gscore = gscore; out->max_off = max_off; out->cells = cells;
	free(H); free(E);
}

/* One pass of the local alignment in plain row order.  The observable rules of the striped SSE2 code
 * (src/ksw.c:111-334): query padded to a multiple of 16 (8-bit) / 8 (16-bit) zero-scoring columns that take part in
 * row maxima; first gap costs o+e and opens from H; everything clamped at 0; 8-bit scores saturate at 255-shift;
 * te = first row with the strict maximum; qe = smallest column attaining it; rows >= minsc feed a merged-run list. */
typedef struct { int score, te, qe, score2, te2; } pass_t;

static void sw_pass(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t mat[25], int o_del,
                    int e_del, int o_ins, int e_ins, int size, int minsc, int endsc, pass_t *r, int64_t *cells)
{
	int lanes = size == 1 ? 16 : 8, qpad = (qlen + lanes - 1) / lanes * lanes;
	int *H = calloc(qpad + 1, sizeof(int)), *E = calloc(qpad + 1, sizeof(int));
	int *b_sc = malloc((tlen + 1) * sizeof(int)), *b_te = malloc((tlen + 1) * sizeof(int)), n_b = 0;
	int oe_del = o_del + e_del, oe_ins = o_ins + e_ins, lo = 127, top = 0, shift, cap, k, i, q, gmax = 0, te = -1, qe = 0;
	for (k = 0; k < 25; ++k) { lo = imin(lo, mat[k]); top = imax(top, mat[k]); }
	shift = -lo; cap = 255 - shift;
	for (i = 0; i < tlen; ++i) {
		const int8_t *s = mat + 5 * target[i];
		int f = 0, diag = 0, rmax = 0, rq = 0;
		for (q = 0; q < qpad; ++q) {
			int h = diag + (q < qlen ? s[query[q]] : 0);
			if (size == 1 && h > cap) h = cap;
			h = imax(imax(imax(h, 0), E[q]), f);
			diag = H[q]; H[q] = h;
			if (h > rmax) { rmax = h; rq = q; }
			E[q] = imax(E[q] - e_del, imax(h - oe_del, 0));
			f = imax(f - e_ins, imax(h - oe_ins, 0));
		}
		*cells += qpad;
		if (rmax >= minsc) {                                                      /* src/ksw.c:192-200 */
			if (n_b == 0 || b_te[n_b - 1] + 1 != i) { b_sc[n_b] = rmax; b_te[n_b++] = i; }
			else if (b_sc[n_b - 1] < rmax) { b_sc[n_b - 1] = rmax; b_te[n_b - 1] = i; }
		}
		if (rmax > gmax) {                                                        /* src/ksw.c:201-206 */
			gmax = rmax; te = i; qe = rq;
			if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) break;
		}
	}
	r->score = (size == 1 && gmax + shift >= 255) ? 255 : gmax;
	r->te = te; r->qe = -1; r->score2 = -1; r->te2 = -1;
	if (size == 2 || r->score != 255) {
		r->qe = qe;
		if (n_b > 0) {                                                            /* src/ksw.c:218-226 */
			int d = (r->score + top - 1) / top;
			for (k = 0; k < n_b; ++k)
				if ((b_te[k] < te - d || b_te[k] > te + d) && b_sc[k] > r->score2) { r->score2 = b_sc[k]; r->te2 = b_te[k]; }
		}
	}
	free(H); free(E); free(b_sc); free(b_te);
}

/* reference src/ksw.c:343-365 */
void orc_ksw_align2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t mat[25],
                    int o_del, int e_del, int o_ins, int e_ins, int xtra, orc_aln_t *out)
{
	int size = (xtra & 0x10000) ? 1 : 2;
	int minsc = (xtra & 0x40000) ? (xtra & 0xffff) : 0x10000;
	int endsc = (xtra & 0x20000) ? (xtra & 0xffff) : 0x10000;
	pass_t r, rr;
	out->cells = 0;
	sw_pass(qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, size, minsc, endsc, &r, &out->cells);
	out->score = r.score; out->te = r.te; out->qe = r.qe; out->score2 = r.score2; out->te2 = r.te2; out->tb = out->qb = -1;
	if (!(xtra & 0x80000) || ((xtra & 0x40000) && r.score < (xtra & 0xffff)) || r.qe < 0) return;
	{
		uint8_t *rq = malloc(r.qe + 1), *rt = malloc(r.te + 1);
		int k;
		for (k = 0; k <= r.qe; ++k) rq[k] = query[r.qe - k];
		for (k = 0; k <= r.te; ++k) rt[k] = target[r.te - k];
		sw_pass(r.qe + 1, rq, r.te + 1, rt, mat, o_del, e_del, o_ins, e_ins, size, 0x10000, r.score & 0xffff, &rr, &out->cells);
		if (rr.score == r.score) { out->tb = r.te - rr.te; out->qb = r.qe - rr.qe; }
		free(rq); free(rt);
	}
}

/* ---------------------------------------------------------------------------------------------------------------
 * ksw_global2, reference src/ksw.c:504-606: banded global alignment with affine gaps (separate deletion / insertion costs)
 * and its traceback.  Restated over FULL matrices (the reference rolls one row): cell (i, j), i = target row, j = query
 * column, both 0-based, lives inside the band when i - w <= j <= i + w.
 *   M(i,j)   = H(i-1,j-1) + S(i,j)                      H(-1,-1) = 0, H(-1,j) = -(o_ins + e_ins (j+1)) for j < w,
 *   H(i,j)   = max{M, E(i,j), F(i,j)}                   H(i,-1) = -(o_del + e_del (i+1)) while the band touches column 0
 *   E(i+1,j) = max{M - o_del, E(i,j)} - e_del           ties: M before E before F for H; "continue the gap" only when strictly
 *   F(i,j+1) = max{M - o_ins, F(i,j)} - e_ins           better than opening it (reference :546-566)
 * The traceback state machine (which = 0 M, 1 E / deletion, 2 F / insertion) reads, per cell, where H came from and whether
 * the E / F of the NEXT cell continues a gap (reference :586-599).  Returns the score; *n_cigar / cigar (caller frees) as
 * the reference; *cells = band cells computed. */
#define ORC_NEG (-0x40000000)
int orc_ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t mat[25],
                    int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t **cigar, int64_t *cells)
{
	const long W = qlen + 2, R = tlen + 3;
	int32_t *H = malloc(sizeof(int32_t) * W * R), *E = malloc(sizeof(int32_t) * W * R);
	uint8_t *src = malloc(W * R), *e_ext = malloc(W * R), *f_ext = malloc(W * R);
	int i, j, score;
	int64_t n_cells = 0;
#define AT(a, i, j) a[((long)(i) + 1) * W + (j) + 1]          /* indices from -1 */
	for (i = -1; i <= tlen; ++i) for (j = -1; j < qlen; ++j) { AT(H, i, j) = ORC_NEG; AT(E, i, j) = ORC_NEG; }
	AT(H, -1, -1) = 0;
	for (j = 0; j < qlen && j + 1 <= w; ++j) AT(H, -1, j) = -(o_ins + e_ins * (j + 1));
	for (i = 0; i < tlen; ++i) {
		const int beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
		int32_t F = ORC_NEG;
		if (beg == 0) AT(H, i, -1) = -(o_del + e_del * (i + 1));
		for (j = beg; j < end; ++j) {
			const int32_t M = AT(H, i - 1, j - 1) + mat[target[i] * 5 + query[j]];
			const int32_t e = AT(E, i, j);
			int32_t h = M, t;
			uint8_t s = 0;
			if (e > h) { h = e; s = 1; }
			if (F > h) { h = F; s = 2; }
			AT(H, i, j) = h; AT(src, i, j) = s;
			t = M - (o_del + e_del);
			AT(e_ext, i, j) = e - e_del > t;
			AT(E, i + 1, j) = e - e_del > t ? e - e_del : t;
			t = M - (o_ins + e_ins);
			AT(f_ext, i, j) = F - e_ins > t;
			F = F - e_ins > t ? F - e_ins : t;
			++n_cells;
		}
	}
	score = qlen > 0 ? (tlen > 0 ? AT(H, tlen - 1, qlen - 1) : AT(H, -1, qlen - 1)) : (tlen > 0 ? AT(H, tlen - 1, -1) : 0);
	if (cells) *cells = n_cells;
	if (n_cigar) *n_cigar = 0;
	if (n_cigar && cigar) {
		uint32_t *cg = malloc(sizeof(uint32_t) * (qlen + tlen + 2));
		int n = 0, which = 0, k;
		i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
#define PUSH(op, len) do { if (n && (cg[n - 1] & 0xf) == (uint32_t)(op)) cg[n - 1] += (uint32_t)(len) << 4; else cg[n++] = (uint32_t)(len) << 4 | (op); } while (0)
		while (i >= 0 && k >= 0) {
			/* in state 0 the cell says where H came from; in state 1 (2) whether the gap that ENDS here was opened here or continues */
			if (which == 0) which = AT(src, i, k);
			else if (which == 1) which = AT(e_ext, i, k) ? 1 : 0;
			else which = AT(f_ext, i, k) ? 2 : 0;
			if (which == 0) { PUSH(0, 1); --i; --k; }
			else if (which == 1) { PUSH(2, 1); --i; }
			else { PUSH(1, 1); --k; }
		}
		if (i >= 0) PUSH(2, i + 1);
		if (k >= 0) PUSH(1, k + 1);
#undef PUSH
		for (i = 0; i < n >> 1; ++i) { uint32_t t = cg[i]; cg[i] = cg[n - 1 - i]; cg[n - 1 - i] = t; }
		*n_cigar = n; *cigar = cg;
	}
#undef AT
	free(H); free(E); free(src); free(e_ext); free(f_ext);
	return score;
}
