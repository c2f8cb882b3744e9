// host_align.cpp - host-resident BWA-MEM logic (see host_align.h).  All arithmetic that decides output bytes
// (float/double compares, (int)(x+.499) roundings, unstable sort ties) is kept in the reference's types and order.
#include "host_align.h"
#include <atomic>
#include "util.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <algorithm>

namespace b200 {

const unsigned char kNt4[256] = {
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4,
	4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4
};

/* ------------------------------------------------------------------ reference coordinates */

int bns_pos2rid_h(const bntseq_t *bns, int64_t pos_f)
{
	if (pos_f >= bns->l_pac) return -1;
	int left = 0, mid = 0, right = bns->n_seqs;
	while (left < right) {
		mid = (left + right) >> 1;
		if (pos_f >= bns->anns[mid].offset) {
			if (mid == bns->n_seqs - 1) break;
			if (pos_f < bns->anns[mid + 1].offset) break;
			left = mid + 1;
		} else right = mid;
	}
	return mid;
}

int bns_intv2rid_h(const bntseq_t *bns, int64_t rb, int64_t re)
{
	int is_rev;
	if (rb < bns->l_pac && re > bns->l_pac) return -2;
	int rid_b = bns_pos2rid_h(bns, bns_depos_h(bns, rb, &is_rev));
	int rid_e = rb < re ? bns_pos2rid_h(bns, bns_depos_h(bns, re - 1, &is_rev)) : rid_b;
	return rid_b == rid_e ? rid_b : -1;
}

void bns_clip_window(const bntseq_t *bns, int64_t *beg, int64_t mid, int64_t *end, int *rid)
{
	int is_rev;
	if (*end < *beg) std::swap(*beg, *end);
	*rid = bns_pos2rid_h(bns, bns_depos_h(bns, mid, &is_rev));
	int64_t far_beg = bns->anns[*rid].offset;
	int64_t far_end = far_beg + bns->anns[*rid].len;
	if (is_rev) {
		int64_t t = far_beg;
		far_beg = (bns->l_pac << 1) - far_end;
		far_end = (bns->l_pac << 1) - t;
	}
	*beg = *beg > far_beg ? *beg : far_beg;
	*end = *end < far_end ? *end : far_end;
}

/* ------------------------------------------------------------------ chaining */

// reference src/bwamem.c:190-211
static bool chain_absorbs(const mem_opt_t *opt, int64_t l_pac, HChain &c, const HSeed &p, int seed_rid)
{
	const HSeed &last = c.seeds.back(), &first = c.seeds.front();
	int64_t qend = last.qbeg + last.len, rend = last.rbeg + last.len;
	if (seed_rid != c.rid) return false;
	if (p.qbeg >= first.qbeg && p.qbeg + p.len <= qend && p.rbeg >= first.rbeg && p.rbeg + p.len <= rend)
		return true;
	if ((last.rbeg < l_pac || first.rbeg < l_pac) && p.rbeg >= l_pac) return false;
	int64_t x = p.qbeg - last.qbeg, y = p.rbeg - last.rbeg;
	if (y >= 0 && x - y <= opt->w && y - x <= opt->w && x - last.len < opt->max_chain_gap && y - last.len < opt->max_chain_gap) {
		c.seeds.push_back(p);
		return true;
	}
	return false;
}

void build_chains(const mem_opt_t *opt, const bntseq_t *bns, int l_seq, const SeedRec *seeds, int64_t n_seeds,
                  int l_rep, std::vector<HChain> &out)
{
	out.clear();
	if (l_seq < opt->min_seed_len) return;
	std::vector<HChain> pool;
	std::vector<int64_t> pos;
	PosTree tree(&pos);
	const int64_t l_pac = bns->l_pac;
	for (int64_t i = 0; i < n_seeds; ++i) {
		const SeedRec &r = seeds[i];
		if (r.rid < 0) continue;
		HSeed s = { r.rbeg, r.qbeg, r.len, r.len };
		bool add = true;
		if (tree.size()) {
			int lower = tree.lower(s.rbeg);
			if (lower >= 0 && chain_absorbs(opt, l_pac, pool[lower], s, r.rid)) add = false;
		}
		if (add) {
			pool.emplace_back();
			HChain &c = pool.back();
			c.seeds.reserve(4);
			c.seeds.push_back(s);
			c.rid = r.rid; c.pos = s.rbeg;
			c.is_alt = !!bns->anns[r.rid].is_alt;
			pos.push_back(s.rbeg);
			tree.insert((int)pool.size() - 1);
		}
	}
	out.reserve(pool.size());
	float frac = (float)l_rep / l_seq;
	tree.in_order([&](int h) { out.push_back(std::move(pool[h])); out.back().frac_rep = frac; });
}

// reference src/bwamem.c:213-232
static int chain_weight(const HChain &c)
{
	int64_t end;
	int w = 0, tmp;
	size_t j;
	for (j = 0, end = 0; j < c.seeds.size(); ++j) {
		const HSeed &s = c.seeds[j];
		if (s.qbeg >= end) w += s.len;
		else if (s.qbeg + s.len > end) w += s.qbeg + s.len - end;
		end = end > s.qbeg + s.len ? end : s.qbeg + s.len;
	}
	tmp = w; w = 0;
	for (j = 0, end = 0; j < c.seeds.size(); ++j) {
		const HSeed &s = c.seeds[j];
		if (s.rbeg >= end) w += s.len;
		else if (s.rbeg + s.len > end) w += s.rbeg + s.len - end;
		end = end > s.rbeg + s.len ? end : s.rbeg + s.len;
	}
	w = w < tmp ? w : tmp;
	return w < 1 << 30 ? w : (1 << 30) - 1;
}

static inline int chn_beg(const HChain &c) { return c.seeds.front().qbeg; }
static inline int chn_end(const HChain &c) { return c.seeds.back().qbeg + c.seeds.back().len; }

void filter_chains(const mem_opt_t *opt, std::vector<HChain> &a)
{
	if (a.empty()) return;
	struct Key { int w, idx; };
	std::vector<Key> keys;
	keys.reserve(a.size());
	for (size_t i = 0; i < a.size(); ++i) {
		HChain &c = a[i];
		c.first = -1; c.kept = 0;
		c.w = chain_weight(c);
		if (c.w >= opt->min_chain_weight) keys.push_back({c.w, (int)i});
	}
	tie_sort(keys, [](const Key &x, const Key &y) { return x.w > y.w; });
	std::vector<HChain> s;
	s.reserve(keys.size());
	for (const Key &k : keys) s.push_back(std::move(a[k.idx]));
	a.swap(s);
	int n_chn = (int)a.size();
	if (n_chn == 0) return;   // the reference would touch a[0] here; only reachable with min_chain_weight > 0
	std::vector<int> chains;
	a[0].kept = 3;
	chains.push_back(0);
	for (int i = 1; i < n_chn; ++i) {
		int large_ovlp = 0;
		size_t k;
		for (k = 0; k < chains.size(); ++k) {
			int j = chains[k];
			int b_max = chn_beg(a[j]) > chn_beg(a[i]) ? chn_beg(a[j]) : chn_beg(a[i]);
			int e_min = chn_end(a[j]) < chn_end(a[i]) ? chn_end(a[j]) : chn_end(a[i]);
			if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
				int li = chn_end(a[i]) - chn_beg(a[i]);
				int lj = chn_end(a[j]) - chn_beg(a[j]);
				int min_l = li < lj ? li : lj;
				if (e_min - b_max >= min_l * opt->mask_level && min_l < opt->max_chain_gap) {
					large_ovlp = 1;
					if (a[j].first < 0) a[j].first = i;
					if (a[i].w < a[j].w * opt->drop_ratio && a[j].w - a[i].w >= opt->min_seed_len << 1)
						break;
				}
			}
		}
		if (k == chains.size()) {
			chains.push_back(i);
			a[i].kept = large_ovlp ? 2 : 3;
		}
	}
	for (size_t i = 0; i < chains.size(); ++i) {
		HChain &c = a[chains[i]];
		if (c.first >= 0) a[c.first].kept = 1;
	}
	int i, k;
	for (i = k = 0; i < n_chn; ++i) {
		if (a[i].kept == 0 || a[i].kept == 3) continue;
		if (++k >= opt->max_chain_extend) break;
	}
	for (; i < n_chn; ++i)
		if (a[i].kept < 3) a[i].kept = 0;
	size_t m = 0;
	for (i = 0; i < n_chn; ++i)
		if (a[i].kept != 0) { if ((size_t)i != m) a[m] = std::move(a[i]); ++m; }
	a.resize(m);
}

static inline int max_gap_for(const mem_opt_t *opt, int qlen)
{
	int l_del = (int)((double)(qlen * opt->a - opt->o_del) / opt->e_del + 1.);
	int l_ins = (int)((double)(qlen * opt->a - opt->o_ins) / opt->e_ins + 1.);
	int l = l_del > l_ins ? l_del : l_ins;
	l = l > 1 ? l : 1;
	return l < opt->w << 1 ? l : opt->w << 1;
}

void chain_window(const mem_opt_t *opt, const bntseq_t *bns, int l_query, const HChain &c, int64_t rmax[2])
{
	const int64_t l_pac = bns->l_pac;
	rmax[0] = l_pac << 1; rmax[1] = 0;
	for (const HSeed &t : c.seeds) {
		int64_t b = t.rbeg - (t.qbeg + max_gap_for(opt, t.qbeg));
		int64_t e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + max_gap_for(opt, l_query - t.qbeg - t.len));
		rmax[0] = rmax[0] < b ? rmax[0] : b;
		rmax[1] = rmax[1] > e ? rmax[1] : e;
	}
	rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
	rmax[1] = rmax[1] < l_pac << 1 ? rmax[1] : l_pac << 1;
	if (rmax[0] < l_pac && l_pac < rmax[1]) {
		if (c.seeds[0].rbeg < l_pac) rmax[1] = l_pac;
		else rmax[0] = l_pac;
	}
	int rid;
	bns_clip_window(bns, &rmax[0], c.seeds[0].rbeg, &rmax[1], &rid);
}

/* ------------------------------------------------------------------ paired-end statistics and pairing */

// mem_pestat's arithmetic (reference src/bwamem_pair.c:67-109) over the candidate insert sizes that the device gathered
// (finish_kernels.h: pestat_candidate; cand[p] = orientation << 32 | insert size, 0 = the pair has none)
void pestat_from_candidates(const mem_opt_t *opt, int64_t n, const uint64_t *cand, mem_pestat_t pes[4])
{
	int d, max;
	std::vector<uint64_t> isize[4];
	memset(pes, 0, 4 * sizeof(mem_pestat_t));
	for (int64_t i = 0; i < n; ++i)
		if (cand[i]) isize[cand[i] >> 32 & 3].push_back(cand[i] & 0xffffffffu);      // (the order does not matter: every use below goes through the sorted array)
	if (bwa_verbose >= 3)
		fprintf(stderr, "[M::%s] # candidate unique pairs for (FF, FR, RF, RR): (%ld, %ld, %ld, %ld)\n", "mem_pestat",
		        (long)isize[0].size(), (long)isize[1].size(), (long)isize[2].size(), (long)isize[3].size());
	for (d = 0; d < 4; ++d) {
		mem_pestat_t *r = &pes[d];
		std::vector<uint64_t> &q = isize[d];
		int p25, p50, p75, x;
		if (q.size() < 10) {
			fprintf(stderr, "[M::%s] skip orientation %c%c as there are not enough pairs\n", "mem_pestat", "FR"[d >> 1 & 1], "FR"[d & 1]);
			r->failed = 1;
			continue;
		} else fprintf(stderr, "[M::%s] analyzing insert size distribution for orientation %c%c...\n", "mem_pestat", "FR"[d >> 1 & 1], "FR"[d & 1]);
		// ascending order (the sums below are taken in it).  Every value is in [1, max_ins]: a counting sort does in a
		// millisecond what a comparison sort of a third of a million values does in twenty, on the critical path of the call
		if (opt->max_ins > 0 && opt->max_ins <= 1 << 22) {
			std::vector<uint32_t> cnt((size_t)opt->max_ins + 2, 0);
			for (uint64_t v : q) ++cnt[v];
			size_t at = 0;
			for (size_t v = 0; v < cnt.size(); ++v)
				for (uint32_t c = cnt[v]; c > 0; --c) q[at++] = v;
		} else std::sort(q.begin(), q.end());
		p25 = (int)q[(int)(.25 * q.size() + .499)];
		p50 = (int)q[(int)(.50 * q.size() + .499)];
		p75 = (int)q[(int)(.75 * q.size() + .499)];
		r->low = (int)(p25 - 2.0 * (p75 - p25) + .499);
		if (r->low < 1) r->low = 1;
		r->high = (int)(p75 + 2.0 * (p75 - p25) + .499);
		fprintf(stderr, "[M::%s] (25, 50, 75) percentile: (%d, %d, %d)\n", "mem_pestat", p25, p50, p75);
		fprintf(stderr, "[M::%s] low and high boundaries for computing mean and std.dev: (%d, %d)\n", "mem_pestat", r->low, r->high);
		size_t k;
		for (k = 0, x = 0, r->avg = 0; k < q.size(); ++k)
			if (q[k] >= (uint64_t)r->low && q[k] <= (uint64_t)r->high) { r->avg += q[k]; ++x; }
		r->avg /= x;
		for (k = 0, r->std = 0; k < q.size(); ++k)
			if (q[k] >= (uint64_t)r->low && q[k] <= (uint64_t)r->high)
				r->std += (q[k] - r->avg) * (q[k] - r->avg);
		r->std = sqrt(r->std / x);
		fprintf(stderr, "[M::%s] mean and std.dev: (%.2f, %.2f)\n", "mem_pestat", r->avg, r->std);
		r->low = (int)(p25 - 3.0 * (p75 - p25) + .499);
		r->high = (int)(p75 + 3.0 * (p75 - p25) + .499);
		if (r->low > r->avg - 4.0 * r->std) r->low = (int)(r->avg - 4.0 * r->std + .499);
		if (r->high < r->avg + 4.0 * r->std) r->high = (int)(r->avg + 4.0 * r->std + .499);
		if (r->low < 1) r->low = 1;
		fprintf(stderr, "[M::%s] low and high boundaries for proper pairs: (%d, %d)\n", "mem_pestat", r->low, r->high);
	}
	for (d = 0, max = 0; d < 4; ++d)
		max = max > (int)isize[d].size() ? max : (int)isize[d].size();
	for (d = 0; d < 4; ++d)
		if (pes[d].failed == 0 && isize[d].size() < max * 0.05) {
			pes[d].failed = 1;
			fprintf(stderr, "[M::%s] skip orientation %c%c\n", "mem_pestat", "FR"[d >> 1 & 1], "FR"[d & 1]);
		}
}

} // namespace b200
