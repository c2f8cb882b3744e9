/* TEST INFRASTRUCTURE ONLY - see oracle.h.  FM-index search restated with the simplest possible arithmetic
 * (symbol-by-symbol rank inside a block instead of the reference's byte look-up table). */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* symbol at position p of the stored BWT string (sentinel removed), reference src/bwt.h:72-78 */
static int bwt_sym(const orc_fm_t *fm, uint64_t p)
{
	const uint32_t *blk = fm->bwt + (p >> 7) * 16 + 8;
	uint32_t w = blk[(p & 127) >> 4];
	return (w >> (2 * (15 - (p & 15)))) & 3;
}

/* number of each symbol in BWT rows [0, k] (k in the coordinate that includes the sentinel row)
 * reference src/bwt.c:169-186; k == -1 gives zeros */
void orc_occ4(const orc_fm_t *fm, uint64_t k, uint64_t cnt[4])
{
	uint64_t p, first;
	int c;
	if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	if (k >= fm->primary) --k;                                  /* the sentinel row holds no symbol */
	first = k >> 7 << 7;
	for (c = 0; c < 4; ++c) cnt[c] = ((const uint64_t *)(fm->bwt + (k >> 7) * 16))[c];
	for (p = first; p <= k; ++p) ++cnt[bwt_sym(fm, p)];
}

/* reference src/bwt.c:262-275 */
void orc_extend(const orc_fm_t *fm, const orc_intv_t *ik, orc_intv_t ok[4], int is_back)
{
	uint64_t lo[4], hi[4], same = is_back ? ik->x0 : ik->x1, other = is_back ? ik->x1 : ik->x0, acc;
	int c;
	orc_occ4(fm, same - 1, lo);
	orc_occ4(fm, same - 1 + ik->x2, hi);
	for (c = 0; c < 4; ++c) {
		uint64_t v = fm->L2[c] + 1 + lo[c];
		if (is_back) ok[c].x0 = v; else ok[c].x1 = v;
		ok[c].x2 = hi[c] - lo[c];
	}
	/* the other strand's interval: T first, then G, C, A stacked above it; +1 when the sentinel is inside */
	acc = other + (same <= fm->primary && same + ik->x2 - 1 >= fm->primary ? 1 : 0);
	for (c = 3; c >= 0; --c) {
		if (is_back) ok[c].x1 = acc; else ok[c].x0 = acc;
		acc += ok[c].x2;
	}
}

static void set_intv(const orc_fm_t *fm, int c, orc_intv_t *ik)
{
	ik->x0 = fm->L2[c] + 1; ik->x2 = fm->L2[c + 1] - fm->L2[c]; ik->x1 = fm->L2[3 - c] + 1; ik->info = 0;
}

/* All SMEMs covering position x, reference src/bwt.c:289-351 (max_intv = 0 as on the mem path) */
int orc_smem1(const orc_fm_t *fm, int len, const uint8_t *q, int x, uint64_t min_intv, orc_intv_t *mem, int *n_mem)
{
	orc_intv_t *fwd = malloc((len + 1) * sizeof(orc_intv_t)), *nxt = malloc((len + 1) * sizeof(orc_intv_t));
	orc_intv_t ik, ok[4];
	int n_fwd = 0, n_nxt, i, j, ret, n = 0;
	*n_mem = 0;
	if (q[x] > 3) { free(fwd); free(nxt); return x + 1; }
	if (min_intv < 1) min_intv = 1;
	set_intv(fm, q[x], &ik);
	ik.info = x + 1;
	/* forward: remember the interval each time its size is about to shrink */
	for (i = x + 1; i < len; ++i) {
		if (q[i] > 3) { fwd[n_fwd++] = ik; break; }
		orc_extend(fm, &ik, ok, 0);
		if (ok[3 - q[i]].x2 != ik.x2) {
			fwd[n_fwd++] = ik;
			if (ok[3 - q[i]].x2 < min_intv) break;
		}
		ik = ok[3 - q[i]]; ik.info = i + 1;
	}
	if (i == len) fwd[n_fwd++] = ik;
	ret = (int)fwd[n_fwd - 1].info;                              /* end of the longest forward match */
	/* longest first */
	for (j = 0; j < n_fwd / 2; ++j) { orc_intv_t t = fwd[j]; fwd[j] = fwd[n_fwd - 1 - j]; fwd[n_fwd - 1 - j] = t; }
	/* backward: extend all candidates by q[i]; a candidate that dies is a SMEM if it is the longest still alive
	 * and not contained in the previously emitted one */
	for (i = x - 1; i >= -1; --i) {
		int c = (i < 0 || q[i] > 3) ? -1 : q[i];
		n_nxt = 0;
		for (j = 0; j < n_fwd; ++j) {
			if (c >= 0) orc_extend(fm, &fwd[j], ok, 1);
			if (c < 0 || ok[c].x2 < min_intv) {
				if (n_nxt == 0 && (n == 0 || (uint64_t)(i + 1) < (mem[n - 1].info >> 32))) {
					mem[n] = fwd[j];
					mem[n].info |= (uint64_t)(i + 1) << 32;
					++n;
				}
			} else if (n_nxt == 0 || ok[c].x2 != nxt[n_nxt - 1].x2) {
				ok[c].info = fwd[j].info;
				nxt[n_nxt++] = ok[c];
			}
		}
		if (n_nxt == 0) break;
		{ orc_intv_t *t = fwd; fwd = nxt; nxt = t; n_fwd = n_nxt; }
	}
	for (j = 0; j < n / 2; ++j) { orc_intv_t t = mem[j]; mem[j] = mem[n - 1 - j]; mem[n - 1 - j] = t; }
	*n_mem = n;
	free(fwd); free(nxt);
	return ret;
}

/* reference src/bwt.c:358-379 */
int orc_seed_strategy1(const orc_fm_t *fm, int len, const uint8_t *q, int x, int min_len, int max_intv, orc_intv_t *mem)
{
	orc_intv_t ik, ok[4];
	int i;
	memset(mem, 0, sizeof *mem);
	if (q[x] > 3) return x + 1;
	set_intv(fm, q[x], &ik);
	for (i = x + 1; i < len; ++i) {
		if (q[i] > 3) return i + 1;
		orc_extend(fm, &ik, ok, 0);
		if (ok[3 - q[i]].x2 < (uint64_t)max_intv && i - x >= min_len) {
			*mem = ok[3 - q[i]];
			mem->info = (uint64_t)x << 32 | (uint32_t)(i + 1);
			return i + 1;
		}
		ik = ok[3 - q[i]];
	}
	return len;
}

static int cmp_info(const void *a, const void *b)
{
	uint64_t x = ((const orc_intv_t *)a)->info, y = ((const orc_intv_t *)b)->info;
	return x < y ? -1 : x > y;
}

/* reference src/bwamem.c:114-162.  Equal keys are identical intervals, so any sort gives the reference's list. */
int orc_collect_intv(const orc_fm_t *fm, int min_seed_len, float split_factor, int split_width, int max_mem_intv,
                     int len, const uint8_t *seq, orc_intv_t *out)
{
	orc_intv_t *tmp = malloc((len + 1) * sizeof(orc_intv_t));
	int split_len = (int)(min_seed_len * split_factor + .499), n = 0, n_tmp, x = 0, i, k, old_n;
	if (len < min_seed_len) { free(tmp); return 0; }
	while (x < len) {
		if (seq[x] > 3) { ++x; continue; }
		x = orc_smem1(fm, len, seq, x, 1, tmp, &n_tmp);
		for (i = 0; i < n_tmp; ++i)
			if ((int)(uint32_t)tmp[i].info - (int)(tmp[i].info >> 32) >= min_seed_len) out[n++] = tmp[i];
	}
	old_n = n;
	for (k = 0; k < old_n; ++k) {
		int start = (int)(out[k].info >> 32), end = (int)(uint32_t)out[k].info;
		if (end - start < split_len || out[k].x2 > (uint64_t)split_width) continue;
		orc_smem1(fm, len, seq, (start + end) >> 1, out[k].x2 + 1, tmp, &n_tmp);
		for (i = 0; i < n_tmp; ++i)
			if ((int)(uint32_t)tmp[i].info - (int)(tmp[i].info >> 32) >= min_seed_len) out[n++] = tmp[i];
	}
	if (max_mem_intv > 0) {
		x = 0;
		while (x < len) {
			orc_intv_t m;
			if (seq[x] > 3) { ++x; continue; }
			x = orc_seed_strategy1(fm, len, seq, x, min_seed_len, max_mem_intv, &m);
			if (m.x2 > 0) out[n++] = m;
		}
	}
	qsort(out, n, sizeof(orc_intv_t), cmp_info);
	free(tmp);
	return n;
}

/* SA[k] by walking the LF mapping to a sampled row, reference src/bwt.c:53-59,86-96 */
uint64_t orc_sa(const orc_fm_t *fm, uint64_t k)
{
	uint64_t steps = 0;
	while (k % (uint64_t)fm->sa_intv != 0) {
		++steps;
		if (k == fm->primary) k = 0;
		else {
			uint64_t cnt[4];
			int c = bwt_sym(fm, k - (k > fm->primary));
			orc_occ4(fm, k, cnt);
			k = fm->L2[c] + cnt[c];
		}
	}
	return steps + fm->sa[k / (uint64_t)fm->sa_intv];
}
