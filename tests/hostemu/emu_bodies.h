// TEST SCAFFOLD (tests/hostemu only): straightforward one-read-at-a-time restatements that the CPU emulation of the device
// stages runs and cross-checks the real kernels' state machines against.  Nothing in mpibwa_b200/ includes this file.
//   fm_seed_strategy1 / fm_collect_intv   reference src/bwt.c:358-379, src/bwamem.c:114-162 (the product runs smem_sweeps.cuh /
//                                         smem_kernel.cuh; the emulation compares their interval sets with this on every read)
//   chain2aln_extend_seed / chain2aln_read   reference src/bwamem.c:632-786 as one loop per read (the product runs the rounds
//                                         of ext_rounds.cuh)
#pragma once
#include "../../mpibwa_b200/csrc/fm_kernels.h"
#include "../../mpibwa_b200/csrc/ext_kernels.h"

namespace b200 {

// forward-only greedy seed; m->x2 == 0 when nothing was found. Returns the next x.
B200_HDN int fm_seed_strategy1(const FmView &fm, int len, const uint8_t *q, int x, int min_len, int max_intv,
                               Intv *m, int64_t *n_blocks)
{
	Intv ik, ok[4];
	m->x0 = m->x1 = m->x2 = m->info = 0;
	if (q[x] > 3) return x + 1;
	fm_set_intv(fm, q[x], ik);
	for (int i = x + 1; i < len; ++i) {
		if (q[i] < 4) {
			int c = 3 - q[i];
			fm_extend(fm, ik, ok, 0, n_blocks);
			if (ok[c].x2 < (uint64_t)max_intv && i - x >= min_len) {
				*m = ok[c];
				m->info = (uint64_t)x << 32 | (uint32_t)(i + 1);
				return i + 1;
			}
			ik = ok[c];
		} else return i + 1;
	}
	return len;
}

// The three seeding passes of one read.  out[0..cap) receives the interval list sorted by info.
// scratch = 3*(len+1) Intv.  Returns the number of intervals, or -(needed) when cap is too small.
B200_HDN int fm_collect_intv(const FmView &fm, const SeedOpt &so, int len, const uint8_t *seq,
                             Intv *out, int cap, Intv *scratch, int64_t *n_blocks)
{
	Intv *mem1 = scratch, *ta = scratch + (len + 1), *tb = scratch + 2 * (len + 1);
	int n = 0, n1, x = 0;
	while (x < len) {                             // pass 1: all SMEMs
		if (seq[x] < 4) {
			x = fm_smem1(fm, len, seq, x, 1, mem1, &n1, ta, tb, n_blocks);
			for (int i = 0; i < n1; ++i) {
				int slen = (int)(uint32_t)mem1[i].info - (int)(mem1[i].info >> 32);
				if (slen >= so.min_seed_len) { if (n < cap) out[n] = mem1[i]; ++n; }
			}
		} else ++x;
	}
	int old_n = n < cap ? n : cap;                // pass 2: re-seed inside long, rare SMEMs
	if (n <= cap)
	for (int k = 0; k < old_n; ++k) {
		const Intv p = out[k];
		int start = (int)(p.info >> 32), end = (int)(int32_t)p.info;
		if (end - start < so.split_len || p.x2 > (uint64_t)so.split_width) continue;
		fm_smem1(fm, len, seq, (start + end) >> 1, p.x2 + 1, mem1, &n1, ta, tb, n_blocks);
		for (int i = 0; i < n1; ++i) {
			int slen = (int)(uint32_t)mem1[i].info - (int)(mem1[i].info >> 32);
			if (slen >= so.min_seed_len) { if (n < cap) out[n] = mem1[i]; ++n; }
		}
	}
	if (so.max_mem_intv > 0) {                    // pass 3: LAST-like greedy seeds
		x = 0;
		while (x < len) {
			if (seq[x] < 4) {
				Intv m;
				x = fm_seed_strategy1(fm, len, seq, x, so.min_seed_len, so.max_mem_intv, &m, n_blocks);
				if (m.x2 > 0) { if (n < cap) out[n] = m; ++n; }
			} else ++x;
		}
	}
	if (n > cap) return -n;
	for (int i = 1; i < n; ++i) {                 // order by (start,end); equal keys are identical intervals
		Intv v = out[i];
		int j = i - 1;
		while (j >= 0 && out[j].info > v.info) { out[j + 1] = out[j]; --j; }
		out[j + 1] = v;
	}
	return n;
}

// Left + right extension of one seed with the band-doubling retry; fills *a.
template <class EH>
B200_HDN void chain2aln_extend_seed(const ExtOpt &o, const uint8_t *pac, int64_t l_pac, int l_query, const uint8_t *query,
                                    const DChain &c, const DSeed *seeds, const DSeed &s, EH eh, DReg *a,
                                    int64_t *cells, int *n_calls)
{
	int aw0 = o.w, aw1 = o.w;
	a->w = o.w; a->score = a->truesc = -1; a->rid = c.rid;
	a->qb = a->qe = 0; a->rb = a->re = 0; a->seedcov = 0; a->seedlen0 = 0; a->pad = 0;
	if (s.qbeg) {
		ExtOut x; x.score = -1; x.qle = x.tle = x.gtle = 0; x.gscore = -1; x.max_off = 0;
		int tlen = (int)(s.rbeg - c.rmax0);
		QRev qa = { query + s.qbeg - 1 };
		TPacRev ta = { pac, l_pac, s.rbeg - 1 };
		for (int i = 0; i < 2; ++i) {
			int prev = a->score;
			aw0 = o.w << i;
			extend_core(s.qbeg, qa, tlen, ta, o, aw0, o.pen_clip5, s.len * o.a, eh, &x, cells);
			if (n_calls) ++*n_calls;
			a->score = x.score;
			if (a->score == prev || x.max_off < (aw0 >> 1) + (aw0 >> 2)) break;
		}
		if (x.gscore <= 0 || x.gscore <= a->score - o.pen_clip5) {
			a->qb = s.qbeg - x.qle; a->rb = s.rbeg - x.tle;
			a->truesc = a->score;
		} else {
			a->qb = 0; a->rb = s.rbeg - x.gtle;
			a->truesc = x.gscore;
		}
	} else { a->score = a->truesc = s.len * o.a; a->qb = 0; a->rb = s.rbeg; }

	if (s.qbeg + s.len != l_query) {
		ExtOut x; x.score = -1; x.qle = x.tle = x.gtle = 0; x.gscore = -1; x.max_off = 0;
		int sc0 = a->score;
		int qe = s.qbeg + s.len;
		int64_t re = s.rbeg + s.len;
		int tlen = (int)(c.rmax1 - re);
		QFwd qa = { query + qe };
		TPacFwd ta = { pac, l_pac, re };
		for (int i = 0; i < 2; ++i) {
			int prev = a->score;
			aw1 = o.w << i;
			extend_core(l_query - qe, qa, tlen, ta, o, aw1, o.pen_clip3, sc0, eh, &x, cells);
			if (n_calls) ++*n_calls;
			a->score = x.score;
			if (a->score == prev || x.max_off < (aw1 >> 1) + (aw1 >> 2)) break;
		}
		if (x.gscore <= 0 || x.gscore <= a->score - o.pen_clip3) {
			a->qe = qe + x.qle; a->re = re + x.tle;
			a->truesc += a->score - sc0;
		} else {
			a->qe = l_query; a->re = re + x.gtle;
			a->truesc += x.gscore - sc0;
		}
	} else { a->qe = l_query; a->re = s.rbeg + s.len; }

	int cov = 0;
	for (int i = 0; i < c.n_seeds; ++i) {
		const DSeed &t = seeds[i];
		if (t.qbeg >= a->qb && t.qbeg + t.len <= a->qe && t.rbeg >= a->rb && t.rbeg + t.len <= a->re) cov += t.len;
	}
	a->seedcov = cov;
	a->w = aw0 > aw1 ? aw0 : aw1;
	a->seedlen0 = s.len;
	a->frac_rep = c.frac_rep;
}

// All chains of one read, in order.  seeds/srt are the chain-local arrays (indexing by c.seed_beg is done here).
template <class EH>
B200_HDN int chain2aln_read(const ExtOpt &o, const uint8_t *pac, int64_t l_pac, int l_query, const uint8_t *query,
                            const DChain *chains, int n_chains, const DSeed *all_seeds, int32_t *all_srt,
                            EH eh, DReg *regs, int64_t *cells, int *n_calls)
{
	int n_av = 0;
	for (int ci = 0; ci < n_chains; ++ci) {
		const DChain &c = chains[ci];
		if (c.n_seeds == 0) continue;
		const DSeed *seeds = all_seeds + c.seed_beg;
		int32_t *srt = all_srt + c.seed_beg;
		for (int k = c.n_seeds - 1; k >= 0; --k) {
			if (!chain2aln_need_extension(o, l_query, c, seeds, srt, k, regs, n_av)) continue;
			chain2aln_extend_seed(o, pac, l_pac, l_query, query, c, seeds, seeds[srt[k]], eh, &regs[n_av], cells, n_calls);
			++n_av;
		}
	}
	return n_av;
}

} // namespace b200
