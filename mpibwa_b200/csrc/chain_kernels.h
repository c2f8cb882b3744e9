// chain_kernels.h - per-read task bodies of the chaining stage: mem_chain, mem_chain_flt and the
// flattening that feeds chain2aln, as one device task per read.
//
//   chain_build_filter   reference src/bwamem.c:251-315 (mem_chain: seeds in look-up order are merged into the chain
//                        found by the B-tree's `lower`, else open a new chain) followed by src/bwamem.c:327-385
//                        (mem_chain_flt: weights, ks_introsort by weight, overlap filter).  The two parity hazards are
//                        restated node for node / swap for swap: the order-5 kbtree (with duplicate chain positions both
//                        the predecessor that `lower` finds and the in-order traversal depend on where the nodes split,
//                        reference src/kbtree.h) and klib's unstable introsort (reference src/ksort.h:162-214), whose
//                        permutation of equal weights decides which chain survives.
//   chain_emit           writes the kept chains in chain2aln's input form: seeds in chain order, rmax window
//                        (src/bwamem.c:621-660), seed order by score (src/bwamem.c:664-669).
//
// A read's working set lives in planes of a per-batch scratch indexed like its seed list (plane p, seed slot i ->
// scr[p * n_total + i]): a read with n seeds can open at most n chains.  The B-tree nodes come from a per-read pool of
// n/2 + 2 nodes (a t = 5 tree over n keys never has more than n/4 + 1).
// The same code runs in tests/hostemu, where it is cross-checked against the host chaining of host_align.cpp.
#pragma once
#include <cstdint>
#include "ext_kernels.h"

namespace b200 {

struct SeedRec { int64_t rbeg; uint16_t qbeg, len; int32_t rid; };   // 16 bytes; reads are shorter than 65536 bases

struct ChainOpt {               // the subset of mem_opt_t the chaining stage reads
	int a, o_del, e_del, o_ins, e_ins, w;
	int min_seed_len, max_chain_gap, min_chain_weight, max_chain_extend;
	float mask_level, drop_ratio;
};

struct BtNode { int32_t n, internal; int32_t key[9]; int32_t child[10]; };     // order-5 node: up to 9 keys

enum { CH_NEXT = 0, CH_FIRST, CH_LAST, CH_ORD, CH_KEYLO, CH_KEYHI, CH_FLTFIRST, CH_KEPT, CH_LIST, CH_N_PLANES };

struct ChainScratch {
	int32_t *scr; int64_t n_total;      // CH_N_PLANES planes of n_total int32
	BtNode *nodes;                      // node pool of the read at nodes + (base >> 1) + 2 * r
	const uint8_t *ctg_alt;             // per contig: is_alt
	B200_HD int32_t &at(int plane, int64_t i) const { return scr[(int64_t)plane * n_total + i]; }
};

/* ---------------------------------------------------------------- the order-5 B-tree of mem_chain (reference src/kbtree.h) */

struct ChainTree {
	BtNode *nd; int n_nodes, root, n_keys;
	const SeedRec *seeds; const int32_t *first;      // position of chain handle h = seeds[first[h]].rbeg

	B200_HD int64_t P(int h) const { return seeds[first[h]].rbeg; }
	B200_HD int new_node(int internal)
	{
		BtNode &z = nd[n_nodes];
		z.n = 0; z.internal = internal;
		for (int i = 0; i < 10; ++i) z.child[i] = -1;
		return n_nodes++;
	}
	B200_HD void init(BtNode *pool, const SeedRec *s, const int32_t *f) { nd = pool; seeds = s; first = f; n_nodes = 0; n_keys = 0; root = new_node(0); }
	// index of the first key equal to p (r = 0) or of the last key below it (r > 0); -1 if all keys are above
	B200_HD int locate(const BtNode &x, int64_t p, int *r) const
	{
		if (x.n == 0) return -1;
		int lo = 0, hi = x.n;
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (P(x.key[mid]) < p) lo = mid + 1; else hi = mid;
		}
		if (lo == x.n) { *r = 1; return x.n - 1; }
		const int64_t q = P(x.key[lo]);
		const int c = (p > q) - (p < q);
		*r = c;
		return c < 0 ? lo - 1 : lo;
	}
	// handle of the closest chain at or below p (the `lower` of kb_intervalp), -1 if none
	B200_HD int lower(int64_t p) const
	{
		int x = root, low = -1;
		while (x >= 0) {
			int r = 0;
			const int i = locate(nd[x], p, &r);
			if (i >= 0 && r == 0) return nd[x].key[i];
			if (i >= 0) low = nd[x].key[i];
			if (!nd[x].internal) return low;
			x = nd[x].child[i + 1];
		}
		return low;
	}
	B200_HD void split(int xi, int i, int yi)
	{
		const int zi = new_node(nd[yi].internal);
		BtNode &x = nd[xi], &y = nd[yi], &z = nd[zi];
		z.n = 4;
		for (int k = 0; k < 4; ++k) z.key[k] = y.key[5 + k];
		if (y.internal) for (int k = 0; k < 5; ++k) z.child[k] = y.child[5 + k];
		y.n = 4;
		for (int k = x.n; k > i; --k) x.child[k + 1] = x.child[k];
		x.child[i + 1] = zi;
		for (int k = x.n - 1; k >= i; --k) x.key[k + 1] = x.key[k];
		x.key[i] = y.key[4];
		++x.n;
	}
	B200_HD void insert(int handle)
	{
		++n_keys;
		int r = root;
		if (nd[r].n == 9) {
			const int s = new_node(1);
			nd[s].child[0] = r;
			root = s;
			split(s, 0, r);
			r = s;
		}
		const int64_t p = P(handle);
		int xi = r, dummy;
		for (;;) {
			if (!nd[xi].internal) {
				BtNode &x = nd[xi];
				const int i = locate(x, p, &dummy);
				for (int k = x.n - 1; k > i; --k) x.key[k + 1] = x.key[k];
				x.key[i + 1] = handle;
				++x.n;
				return;
			}
			int i = locate(nd[xi], p, &dummy) + 1;
			if (nd[nd[xi].child[i]].n == 9) {
				split(xi, i, nd[xi].child[i]);
				if (p > P(nd[xi].key[i])) ++i;
			}
			xi = nd[xi].child[i];
		}
	}
	// in-order traversal into out[]; returns the number of handles
	B200_HD int in_order(int32_t *out, int64_t stride_unused = 0) const
	{
		(void)stride_unused;
		int sp = 0, n = 0;
		int stk_node[16], stk_i[16];        // height of a t = 5 tree over 2^31 keys is below 16
		stk_node[0] = root; stk_i[0] = 0;
		while (sp >= 0) {
			const BtNode &x = nd[stk_node[sp]];
			const int i = stk_i[sp];
			if (x.internal) {
				// visit child i, then key i
				if (i <= x.n) {
					if (i > 0) out[n++] = x.key[i - 1];
					stk_i[sp] = i + 1;
					++sp; stk_node[sp] = x.child[i]; stk_i[sp] = 0;
				} else --sp;
			} else {
				for (int k = 0; k < x.n; ++k) out[n++] = x.key[k];
				--sp;
			}
		}
		return n;
	}
};

/* ---------------------------------------------------------------- klib's introsort on (weight, index) keys, descending weight */

B200_HD bool ck_lt(uint64_t x, uint64_t y) { return (int32_t)(x >> 32) > (int32_t)(y >> 32); }
B200_HD void ck_swap(uint64_t &x, uint64_t &y) { const uint64_t t = x; x = y; y = t; }

// keys as two int32 planes would double every access; they are stored as uint64 pairs in two adjacent planes instead
struct KeyArr {
	int32_t *lo, *hi;
	B200_HD uint64_t get(int i) const { return (uint64_t)(uint32_t)hi[i] << 32 | (uint32_t)lo[i]; }
	B200_HD void set(int i, uint64_t v) const { lo[i] = (int32_t)(uint32_t)v; hi[i] = (int32_t)(uint32_t)(v >> 32); }
	B200_HD void swap(int i, int j) const { const uint64_t a = get(i), b = get(j); set(i, b); set(j, a); }
};

B200_HD void ck_insertion(const KeyArr &a, int s, int t)       // [s, t)
{
	for (int i = s + 1; i < t; ++i)
		for (int j = i; j > s && ck_lt(a.get(j), a.get(j - 1)); --j) a.swap(j, j - 1);
}

B200_HD void ck_comb(const KeyArr &a, int s, int n)
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
			if (ck_lt(a.get(j), a.get(i))) { a.swap(i, j); swapped = true; }
		}
	} while (swapped || gap > 2);
	if (gap != 1) ck_insertion(a, s, s + n);
}

B200_HD void ck_introsort(const KeyArr &a, int n)
{
	if (n < 1) return;
	if (n == 2) { if (ck_lt(a.get(1), a.get(0))) a.swap(0, 1); return; }
	int d;
	for (d = 2; (1ul << d) < (unsigned long)n; ++d) {}
	int stk_lo[72], stk_hi[72], stk_d[72], sp = 0;
	int s = 0, t = n - 1;
	d <<= 1;
	for (;;) {
		if (s < t) {
			if (--d == 0) { ck_comb(a, s, t - s + 1); t = s; continue; }
			int i = s, j = t, k = i + ((j - i) >> 1) + 1;
			if (ck_lt(a.get(k), a.get(i))) { if (ck_lt(a.get(k), a.get(j))) k = j; }
			else k = ck_lt(a.get(j), a.get(i)) ? i : j;
			const uint64_t pivot = a.get(k);
			if (k != t) a.swap(k, t);
			for (;;) {
				do ++i; while (ck_lt(a.get(i), pivot));
				do --j; while (i <= j && ck_lt(pivot, a.get(j)));
				if (j <= i) break;
				a.swap(i, j);
			}
			a.swap(i, t);
			if (i - s > t - i) {
				if (i - s > 16) { stk_lo[sp] = s; stk_hi[sp] = i - 1; stk_d[sp] = d; ++sp; }
				s = t - i > 16 ? i + 1 : t;
			} else {
				if (t - i > 16) { stk_lo[sp] = i + 1; stk_hi[sp] = t; stk_d[sp] = d; ++sp; }
				t = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (sp == 0) { ck_insertion(a, 0, n); return; }
			--sp; s = stk_lo[sp]; t = stk_hi[sp]; d = stk_d[sp];
		}
	}
}

/* ---------------------------------------------------------------- mem_chain + mem_chain_flt for one read */

// Returns the number of kept chains; their handles, in mem_chain_flt's output order, are left in plane CH_LIST and the
// total number of their seeds in *n_kept_seeds.  Seeds of chain h: CH_FIRST[h], then CH_NEXT links.
B200_HD int chain_build_filter(const ChainOpt &co, int64_t l_pac, const ChainScratch &S, int r, int l_seq, const SeedRec *seeds_all,
                               int64_t base, int n_in, int *n_kept_seeds)
{
	*n_kept_seeds = 0;
	if (l_seq < co.min_seed_len || n_in == 0) return 0;
	const SeedRec *seeds = seeds_all + base;
	int32_t *nxt = &S.at(CH_NEXT, base), *first = &S.at(CH_FIRST, base), *last = &S.at(CH_LAST, base), *ord = &S.at(CH_ORD, base);
	ChainTree tree;
	tree.init(S.nodes + (base >> 1) + 2 * (int64_t)r, seeds, first);
	int n_pool = 0;
	for (int i = 0; i < n_in; ++i) {
		const SeedRec p = seeds[i];
		if (p.rid < 0) continue;
		bool add = true;
		if (tree.n_keys) {
			const int h = tree.lower(p.rbeg);
			if (h >= 0) {                                   // test_and_merge, reference src/bwamem.c:190-211
				const SeedRec f = seeds[first[h]], l = seeds[last[h]];
				const int64_t qend = (int64_t)l.qbeg + l.len, rend = l.rbeg + l.len;
				if (p.rid == f.rid) {
					if (p.qbeg >= f.qbeg && (int64_t)p.qbeg + p.len <= qend && p.rbeg >= f.rbeg && p.rbeg + p.len <= rend) add = false;   // contained
					else if (!((l.rbeg < l_pac || f.rbeg < l_pac) && p.rbeg >= l_pac)) {
						const int64_t x = (int64_t)p.qbeg - l.qbeg, y = p.rbeg - l.rbeg;
						if (y >= 0 && x - y <= co.w && y - x <= co.w && x - l.len < co.max_chain_gap && y - l.len < co.max_chain_gap) {
							nxt[last[h]] = i; nxt[i] = -1; last[h] = i;
							add = false;
						}
					}
				}
			}
		}
		if (add) {
			const int h = n_pool++;
			first[h] = i; last[h] = i; nxt[i] = -1;
			tree.insert(h);
		}
	}
	const int n_chains = tree.in_order(ord);
	// ---- mem_chain_flt
	KeyArr keys = { &S.at(CH_KEYLO, base), &S.at(CH_KEYHI, base) };
	int n_chn = 0;
	for (int c = 0; c < n_chains; ++c) {
		const int h = ord[c];
		int64_t end = 0;
		int w = 0, tmp;
		for (int j = first[h]; j >= 0; j = nxt[j]) {            // mem_chain_weight, reference src/bwamem.c:213-232
			const SeedRec s = seeds[j];
			if (s.qbeg >= end) w += s.len;
			else if ((int64_t)s.qbeg + s.len > end) w += (int)(s.qbeg + s.len - end);
			end = end > (int64_t)s.qbeg + s.len ? end : (int64_t)s.qbeg + s.len;
		}
		tmp = w; w = 0; end = 0;
		for (int j = first[h]; j >= 0; j = nxt[j]) {
			const SeedRec s = seeds[j];
			if (s.rbeg >= end) w += s.len;
			else if (s.rbeg + s.len > end) w += (int)(s.rbeg + s.len - end);
			end = end > s.rbeg + s.len ? end : s.rbeg + s.len;
		}
		w = w < tmp ? w : tmp;
		w = w < 1 << 30 ? w : (1 << 30) - 1;
		if (w >= co.min_chain_weight) keys.set(n_chn++, (uint64_t)(uint32_t)w << 32 | (uint32_t)h);
	}
	if (n_chn == 0) return 0;
	ck_introsort(keys, n_chn);
	// sorted chain i: handle keys.lo[i], weight keys.hi[i]
	int32_t *flt_first = &S.at(CH_FLTFIRST, base), *kept = &S.at(CH_KEPT, base), *list = &S.at(CH_LIST, base);
	for (int i = 0; i < n_chn; ++i) { flt_first[i] = -1; kept[i] = 0; }
	int n_list = 0;
	kept[0] = 3;
	list[n_list++] = 0;
	for (int i = 1; i < n_chn; ++i) {
		const int hi_ = keys.lo[i], wi = keys.hi[i];
		const int bi = seeds[first[hi_]].qbeg, ei = seeds[last[hi_]].qbeg + seeds[last[hi_]].len;
		const int alt_i = S.ctg_alt[seeds[first[hi_]].rid];
		int large_ovlp = 0, k;
		for (k = 0; k < n_list; ++k) {
			const int j = list[k], hj = keys.lo[j], wj = keys.hi[j];
			const int bj = seeds[first[hj]].qbeg, ej = seeds[last[hj]].qbeg + seeds[last[hj]].len;
			const int b_max = bj > bi ? bj : bi, e_min = ej < ei ? ej : ei;
			if (e_min > b_max && (!S.ctg_alt[seeds[first[hj]].rid] || alt_i)) {
				const int li = ei - bi, lj = ej - bj, min_l = li < lj ? li : lj;
				if (e_min - b_max >= min_l * co.mask_level && min_l < co.max_chain_gap) {
					large_ovlp = 1;
					if (flt_first[j] < 0) flt_first[j] = i;
					if (wi < wj * co.drop_ratio && wj - wi >= co.min_seed_len << 1) break;
				}
			}
		}
		if (k == n_list) { list[n_list++] = i; kept[i] = large_ovlp ? 2 : 3; }
	}
	for (int k = 0; k < n_list; ++k) { const int f = flt_first[list[k]]; if (f >= 0) kept[f] = 1; }
	int i, k;
	for (i = k = 0; i < n_chn; ++i) {
		if (kept[i] == 0 || kept[i] == 3) continue;
		if (++k >= co.max_chain_extend) break;
	}
	for (; i < n_chn; ++i) if (kept[i] < 3) kept[i] = 0;
	int m = 0, ns = 0;
	for (i = 0; i < n_chn; ++i) {
		if (kept[i] == 0) continue;
		const int h = keys.lo[i];
		list[m++] = h;                                  // (m <= i: the kept-list entries read above are no longer needed)
		for (int j = first[h]; j >= 0; j = nxt[j]) ++ns;
	}
	*n_kept_seeds = ns;
	return m;
}

/* ---------------------------------------------------------------- flattening for chain2aln */

B200_HD int chain_max_gap(const ChainOpt &co, int qlen)            // cal_max_gap, reference src/bwamem.c:621-628
{
	const int l_del = (int)((double)(qlen * co.a - co.o_del) / co.e_del + 1.);
	const int l_ins = (int)((double)(qlen * co.a - co.o_ins) / co.e_ins + 1.);
	int l = l_del > l_ins ? l_del : l_ins;
	l = l > 1 ? l : 1;
	return l < co.w << 1 ? l : co.w << 1;
}

// writes the n_kept chains of the read at chains[c0 ..] with their seeds at dseeds[s0 ..] and the by-score order in srt[s0 ..]
B200_HD void chain_emit(const ChainOpt &co, const FmView &fm, const ChainScratch &S, int l_seq, int l_rep, const SeedRec *seeds_all,
                        int64_t base, int n_kept, int64_t c0, int64_t s0, DChain *chains, DSeed *dseeds, int32_t *srt)
{
	const SeedRec *seeds = seeds_all + base;
	const int32_t *nxt = &S.at(CH_NEXT, base), *first = &S.at(CH_FIRST, base), *list = &S.at(CH_LIST, base);
	const int64_t l_pac = fm.l_pac;
	const float frac = (float)l_rep / l_seq;
	int64_t si = s0;
	for (int c = 0; c < n_kept; ++c) {
		const int h = list[c];
		DChain d;
		d.seed_beg = (int32_t)si; d.rid = seeds[first[h]].rid; d.frac_rep = frac;
		int64_t rmax0 = l_pac << 1, rmax1 = 0;
		int n = 0;
		for (int j = first[h]; j >= 0; j = nxt[j], ++n) {
			const SeedRec t = seeds[j];
			DSeed o;
			o.rbeg = t.rbeg; o.qbeg = t.qbeg; o.len = t.len; o.score = t.len; o.pad = 0;
			dseeds[si + n] = o;
			const int64_t b = t.rbeg - (t.qbeg + chain_max_gap(co, t.qbeg));
			const int64_t e = t.rbeg + t.len + ((l_seq - t.qbeg - t.len) + chain_max_gap(co, l_seq - t.qbeg - t.len));
			rmax0 = rmax0 < b ? rmax0 : b;
			rmax1 = rmax1 > e ? rmax1 : e;
			// seeds in ascending (score, position in chain) order: insertion into the sorted prefix
			int k = n;
			while (k > 0 && dseeds[si + srt[si + k - 1]].score > o.score) { srt[si + k] = srt[si + k - 1]; --k; }
			srt[si + k] = n;
		}
		d.n_seeds = n;
		rmax0 = rmax0 > 0 ? rmax0 : 0;
		rmax1 = rmax1 < l_pac << 1 ? rmax1 : l_pac << 1;
		const int64_t rbeg0 = seeds[first[h]].rbeg;
		if (rmax0 < l_pac && l_pac < rmax1) {
			if (rbeg0 < l_pac) rmax1 = l_pac; else rmax0 = l_pac;
		}
		// clip to the contig of the first seed (bns_fetch_seq, reference src/bntseq.c:421-446)
		{
			int is_rev;
			if (rmax1 < rmax0) { const int64_t t = rmax0; rmax0 = rmax1; rmax1 = t; }
			const int rid = fm_pos2rid(fm, fm_depos(fm, rbeg0, &is_rev));
			int64_t far_beg = fm.ctg_off[rid], far_end = far_beg + fm.ctg_len[rid];
			if (is_rev) { const int64_t t = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - t; }
			rmax0 = rmax0 > far_beg ? rmax0 : far_beg;
			rmax1 = rmax1 < far_end ? rmax1 : far_end;
		}
		d.rmax0 = rmax0; d.rmax1 = rmax1;
		chains[c0 + c] = d;
		si += n;
	}
}

} // namespace b200
