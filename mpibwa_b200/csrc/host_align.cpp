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

AlignCtx &align_ctx()
{
	static thread_local AlignCtx ctx;
	return ctx;
}

/* ------------------------------------------------------------------ cycle accounting (B200_HOST_PROF) */
static const bool g_prof_on = getenv("B200_HOST_PROF") != nullptr;
static std::atomic<uint64_t> g_prof_cyc[HP_N], g_prof_cnt[HP_N];
struct ProfTls { uint64_t cyc[HP_N], cnt[HP_N]; };
static thread_local ProfTls t_prof;
struct ProfScope {
	int k; uint64_t t0;
	explicit ProfScope(int k_) : k(k_), t0(g_prof_on ? __builtin_ia32_rdtsc() : 0) {}
	~ProfScope() { if (g_prof_on) { ProfTls &p = t_prof; p.cyc[k] += __builtin_ia32_rdtsc() - t0; ++p.cnt[k]; } }
};
void host_prof_flush()
{
	if (!g_prof_on) return;
	ProfTls &p = t_prof;
	for (int k = 0; k < HP_N; ++k) { g_prof_cyc[k] += p.cyc[k]; g_prof_cnt[k] += p.cnt[k]; p.cyc[k] = p.cnt[k] = 0; }
}
void host_prof_report(const char *what)
{
	if (!g_prof_on) return;
	static const char *name[HP_N] = { "sam_pe_finish", "mark_primary", "pair_ends", "gen_alt", "reg2aln", "gen_cigar", "aln2sam", "dup_cstr", "reg2sam" };
	fprintf(stderr, "[host_prof] %s:", what);
	for (int k = 0; k < HP_N; ++k) fprintf(stderr, " %s %.1f Mcyc/%lluk", name[k], g_prof_cyc[k].exchange(0) * 1e-6, (unsigned long long)(g_prof_cnt[k].exchange(0) / 1000));
	fprintf(stderr, "\n");
}

char *dup_cstr(const std::string &s)
{
	ProfScope ps(HP_DUP);
	char *p = (char *)malloc(s.size() + 1);
	memcpy(p, s.data(), s.size());
	p[s.size()] = 0;
	return p;
}

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

static inline int pac_base(const uint8_t *pac, int64_t l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }

// four base codes per pac byte, first base in the low byte of the word (forward) / complemented and reversed (reverse strand)
static const struct PacLut {
	uint32_t fwd[256], rev[256];
	PacLut()
	{
		for (int b = 0; b < 256; ++b) {
			const uint32_t c0 = b >> 6 & 3, c1 = b >> 4 & 3, c2 = b >> 2 & 3, c3 = b & 3;
			fwd[b] = c0 | c1 << 8 | c2 << 16 | c3 << 24;
			rev[b] = (3 - c3) | (3 - c2) << 8 | (3 - c1) << 16 | (3 - c0) << 24;
		}
	}
} kPacLut;

// bns_get_seq into a caller buffer of end - beg bytes (0 <= beg <= end <= 2 l_pac); returns the number of bases written
static int64_t bns_get_seq_raw(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, uint8_t *d)
{
	int64_t l = 0;
	if (beg >= l_pac || end <= l_pac) {
		if (beg >= l_pac) {                        // reverse strand: complement of forward [beg_f+1, end_f], last base first
			int64_t beg_f = (l_pac << 1) - 1 - end, end_f = (l_pac << 1) - 1 - beg;
			int64_t k = end_f;
			for (; k > beg_f && (k & 3) != 3; --k) d[l++] = 3 - pac_base(pac, k);
			for (; k - 3 > beg_f; k -= 4) { const uint32_t w = kPacLut.rev[pac[k >> 2]]; memcpy(d + l, &w, 4); l += 4; }
			for (; k > beg_f; --k) d[l++] = 3 - pac_base(pac, k);
		} else {
			int64_t k = beg;
			for (; k < end && (k & 3); ++k) d[l++] = pac_base(pac, k);
			for (; k + 4 <= end; k += 4) { const uint32_t w = kPacLut.fwd[pac[k >> 2]]; memcpy(d + l, &w, 4); l += 4; }
			for (; k < end; ++k) d[l++] = pac_base(pac, k);
		}
	}
	return l;
}

void bns_get_seq_h(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, std::vector<uint8_t> &seq)
{
	seq.clear();
	if (end < beg) std::swap(beg, end);
	if (end > l_pac << 1) end = l_pac << 1;
	if (beg < 0) beg = 0;
	if (beg >= l_pac || end <= l_pac) {
		seq.resize(end - beg);
		bns_get_seq_raw(l_pac, pac, beg, end, seq.data());
	}
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

/* ------------------------------------------------------------------ global alignment + CIGAR */

static const int kMinusInf = -0x40000000;

int global_align(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                 int o_del, int e_del, int o_ins, int e_ins, int w, std::vector<uint32_t> *cigar)
{
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	struct Cell { int32_t h, e; };
	std::vector<Cell> eh(qlen + 1);
	std::vector<int8_t> qp((size_t)qlen * 5);
	std::vector<uint8_t> z;
	if (cigar) { cigar->clear(); z.resize((size_t)n_col * tlen); }
	for (int k = 0, i = 0; k < 5; ++k) {
		const int8_t *p = mat + k * 5;
		for (int j = 0; j < qlen; ++j) qp[i++] = p[query[j]];
	}
	int j;
	eh[0].h = 0; eh[0].e = kMinusInf;
	for (j = 1; j <= qlen && j <= w; ++j) { eh[j].h = -(o_ins + e_ins * j); eh[j].e = kMinusInf; }
	for (; j <= qlen; ++j) eh[j].h = eh[j].e = kMinusInf;
	for (int i = 0; i < tlen; ++i) {
		int32_t f = kMinusInf, h1, t;
		const int8_t *q = &qp[(size_t)target[i] * qlen];
		int beg = i > w ? i - w : 0;
		int end = i + w + 1 < qlen ? i + w + 1 : qlen;
		h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : kMinusInf;
		uint8_t *zi = cigar ? &z[(size_t)i * n_col] : nullptr;
		for (j = beg; j < end; ++j) {
			Cell *p = &eh[j];
			int32_t h, m = p->h, e = p->e;
			uint8_t d;
			p->h = h1;
			m += q[j];
			d = m >= e ? 0 : 1;
			h = m >= e ? m : e;
			d = h >= f ? d : 2;
			h = h >= f ? h : f;
			h1 = h;
			t = m - oe_del;
			e -= e_del;
			d |= e > t ? 1 << 2 : 0;
			e = e > t ? e : t;
			p->e = e;
			t = m - oe_ins;
			f -= e_ins;
			d |= f > t ? 2 << 4 : 0;
			f = f > t ? f : t;
			if (zi) zi[j - beg] = d;
		}
		eh[end].h = h1; eh[end].e = kMinusInf;
	}
	int score = eh[qlen].h;
	if (cigar) {
		int which = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
		auto push = [&](int op, int len) {
			if (cigar->empty() || op != (int)(cigar->back() & 0xf)) cigar->push_back((uint32_t)len << 4 | op);
			else cigar->back() += (uint32_t)len << 4;
		};
		while (i >= 0 && k >= 0) {
			which = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
			if (which == 0) { push(0, 1); --i; --k; }
			else if (which == 1) { push(2, 1); --i; }
			else { push(1, 1); --k; }
		}
		if (i >= 0) push(2, i + 1);
		if (k >= 0) push(1, k + 1);
		std::reverse(cigar->begin(), cigar->end());
	}
	return score;
}

static void put_int(std::string &s, long c)
{
	char buf[32];
	int l = 0;
	if (c == 0) { s.push_back('0'); return; }
	unsigned long x = c < 0 ? (unsigned long)(-c) : (unsigned long)c;
	for (; x > 0; x /= 10) buf[l++] = (char)(x % 10 + '0');
	if (c < 0) buf[l++] = '-';
	while (l > 0) s.push_back(buf[--l]);
}

bool gen_cigar(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac,
               const uint8_t *pac, int l_query, uint8_t *query, int64_t rb, int64_t re, int *score,
               std::vector<uint32_t> *cigar, int *NM, std::string *md, const GlobalRes *pre)
{
	ProfScope ps(HP_GEN_CIGAR);
	if (cigar) cigar->clear();
	if (NM) *NM = -1;
	if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return false;
	std::vector<uint8_t> rseq_heap;
	uint8_t rseq_stack[1024 + 8];
	uint8_t *rseq = rseq_stack;
	int64_t rlen;
	bool rseq_reversed = false;
	if (re - rb <= 1024 && rb >= 0 && re <= l_pac << 1) {      // the common case: no heap traffic for the window
		if (rb >= l_pac) {
			// the reverse-strand window is needed back to front (below): that is the complement of the forward strand read
			// in ascending order, fetched directly instead of fetching and reversing
			const int64_t beg_f = (l_pac << 1) - 1 - re, end_f = (l_pac << 1) - 1 - rb;
			int64_t l = 0, k = beg_f + 1;
			for (; k <= end_f && (k & 3); ++k) rseq_stack[l++] = 3 - pac_base(pac, k);
			for (; k + 4 <= end_f + 1; k += 4) { const uint32_t w = 0x03030303u - kPacLut.fwd[pac[k >> 2]]; memcpy(rseq_stack + l, &w, 4); l += 4; }
			for (; k <= end_f; ++k) rseq_stack[l++] = 3 - pac_base(pac, k);
			rlen = l;
			rseq_reversed = true;
		} else rlen = bns_get_seq_raw(l_pac, pac, rb, re, rseq_stack);
	} else {
		bns_get_seq_h(l_pac, pac, rb, re, rseq_heap);
		rseq_heap.resize(rseq_heap.size() + 8);
		rseq = rseq_heap.data();
		rlen = (int64_t)rseq_heap.size() - 8;
	}
	if (re - rb != rlen) return false;
	if (rb >= l_pac) {
		std::reverse(query, query + l_query);
		if (!rseq_reversed) std::reverse(rseq, rseq + rlen);
	}
	if (pre) {                                    // alignment already done by the CIGAR stage on the device
		cigar->assign(pre->cigar, pre->cigar + pre->n_cigar);
		*score = pre->score;
	} else if (l_query == re - rb && w_ == 0) {
		if (cigar) cigar->push_back((uint32_t)l_query << 4 | 0);
		int sc = 0, i = 0;
		if (mat[0] == mat[6] && mat[0] == mat[12] && mat[0] == mat[18])      // eight equal non-N bases at a time
			for (; i + 8 <= l_query; i += 8) {
				uint64_t qa, ra;
				memcpy(&qa, query + i, 8); memcpy(&ra, rseq + i, 8);
				if (qa == ra && !(qa & 0x0404040404040404ull)) sc += 8 * mat[0];
				else for (int k = i; k < i + 8; ++k) sc += mat[rseq[k] * 5 + query[k]];
			}
		for (; i < l_query; ++i) sc += mat[rseq[i] * 5 + query[i]];
		*score = sc;
	} else {
		int w, max_gap, max_ins, max_del, min_w;
		max_ins = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_ins) / e_ins + 1.);
		max_del = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_del) / e_del + 1.);
		max_gap = max_ins > max_del ? max_ins : max_del;
		max_gap = max_gap > 1 ? max_gap : 1;
		w = (max_gap + abs((int)rlen - l_query) + 1) >> 1;
		w = w < w_ ? w : w_;
		min_w = abs((int)rlen - l_query) + 3;
		w = w > min_w ? w : min_w;
		*score = global_align(l_query, query, (int)rlen, rseq, mat, o_del, e_del, o_ins, e_ins, w, cigar);
	}
	if (NM && cigar) {
		int x = 0, y = 0, u = 0, n_mm = 0, n_gap = 0;
		const char *int2base = rb < l_pac ? "ACGTN" : "TGCAN";
		md->clear();
		const int n_cigar = (int)cigar->size();
		for (int k = 0; k < n_cigar; ++k) {
			int op = (*cigar)[k] & 0xf, len = (*cigar)[k] >> 4;
			if (op == 0) {
				for (int i = 0; i < len; ++i) {
					if (i + 8 <= len) {                   // eight equal bases at a time (most of a read matches)
						uint64_t qa, ra;
						memcpy(&qa, query + x + i, 8); memcpy(&ra, rseq + y + i, 8);
						if (qa == ra) { u += 8; i += 7; continue; }
					}
					if (query[x + i] != rseq[y + i]) {
						put_int(*md, u);
						md->push_back(int2base[rseq[y + i]]);
						++n_mm; u = 0;
					} else ++u;
				}
				x += len; y += len;
			} else if (op == 2) {
				if (k > 0 && k < n_cigar - 1) {
					put_int(*md, u); md->push_back('^');
					for (int i = 0; i < len; ++i) md->push_back(int2base[rseq[y + i]]);
					u = 0; n_gap += len;
				}
				y += len;
			} else if (op == 1) { x += len; n_gap += len; }
		}
		put_int(*md, u);
		*NM = n_mm + n_gap;
	}
	if (rb >= l_pac) std::reverse(query, query + l_query);
	return true;
}

/* ------------------------------------------------------------------ region de-duplication */

static const float kPatchMaxRBw = 0.05f, kPatchMinScRatio = 0.90f;

static int patch_reg(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, uint8_t *query,
                     const mem_alnreg_t *a, const mem_alnreg_t *b, int *_w)
{
	int w, score, q_s, r_s;
	double r;
	if (bns == 0 || pac == 0 || query == 0) return 0;
	if (a->rb < bns->l_pac && b->rb >= bns->l_pac) return 0;
	if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
	w = (int)((a->re - b->rb) - (a->qe - b->qb));
	w = w > 0 ? w : -w;
	r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
	r = r > 0. ? r : -r;
	if (a->re < b->rb || a->qe < b->qb) {
		if (w > opt->w << 1 || r >= kPatchMaxRBw) return 0;
	} else if (w > opt->w << 2 || r >= kPatchMaxRBw * 2) return 0;
	w += a->w + b->w;
	w = w < opt->w << 2 ? w : opt->w << 2;
	score = 0;
	gen_cigar(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w, bns->l_pac, pac, b->qe - a->qb,
	          query + a->qb, a->rb, b->re, &score, nullptr, nullptr, nullptr);
	q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
	r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
	if ((double)score / (q_s > r_s ? q_s : r_s) < kPatchMinScRatio) return 0;
	*_w = w;
	return score;
}

int sort_dedup_patch(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, uint8_t *query, int n, mem_alnreg_t *a)
{
	int m, i, j;
	if (n <= 1) return n;
	tie_sort((size_t)n, a, [](const mem_alnreg_t &x, const mem_alnreg_t &y) { return x.re < y.re; });
	for (i = 0; i < n; ++i) a[i].n_comp = 1;
	for (i = 1; i < n; ++i) {
		mem_alnreg_t *p = &a[i];
		if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + opt->max_chain_gap) continue;
		for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt->max_chain_gap; --j) {
			mem_alnreg_t *q = &a[j];
			int64_t orr, oq, mr, mq;
			int score, w;
			if (q->qe == q->qb) continue;
			orr = q->re - p->rb;
			oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
			mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
			mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
			if (orr > opt->mask_level_redun * mr && oq > opt->mask_level_redun * mq) {
				if (p->score < q->score) { p->qe = p->qb; break; }
				else q->qe = q->qb;
			} else if (q->rb < p->rb && (score = patch_reg(opt, bns, pac, query, q, p, &w)) > 0) {
				p->n_comp += q->n_comp + 1;
				p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
				p->sub = p->sub > q->sub ? p->sub : q->sub;
				p->csub = p->csub > q->csub ? p->csub : q->csub;
				p->qb = q->qb; p->rb = q->rb;
				p->truesc = p->score = score;
				p->w = w;
				q->qb = q->qe;
			}
		}
	}
	for (i = 0, m = 0; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
	n = m;
	tie_sort((size_t)n, a, [](const mem_alnreg_t &x, const mem_alnreg_t &y) {
		return x.score > y.score || (x.score == y.score && (x.rb < y.rb || (x.rb == y.rb && x.qb < y.qb)));
	});
	for (i = 1; i < n; ++i)
		if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb)
			a[i].qe = a[i].qb;
	for (i = 1, m = 1; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
	return m;
}

/* ------------------------------------------------------------------ primary / secondary marking, mapQ */

static void mark_primary_core(const mem_opt_t *opt, int n, mem_alnreg_t *a, std::vector<int> &z)
{
	int i, tmp;
	size_t k;
	tmp = opt->a + opt->b;
	tmp = opt->o_del + opt->e_del > tmp ? opt->o_del + opt->e_del : tmp;
	tmp = opt->o_ins + opt->e_ins > tmp ? opt->o_ins + opt->e_ins : tmp;
	z.clear();
	z.push_back(0);
	for (i = 1; i < n; ++i) {
		for (k = 0; k < z.size(); ++k) {
			int j = z[k];
			int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
			int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
			if (e_min > b_max) {
				int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
				if (e_min - b_max >= min_l * opt->mask_level) {
					if (a[j].sub == 0) a[j].sub = a[i].score;
					if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt))
						++a[j].sub_n;
					break;
				}
			}
		}
		if (k == z.size()) z.push_back(i);
		else a[i].secondary = z[k];
	}
}

int mark_primary_se(const mem_opt_t *opt, int n, mem_alnreg_t *a, int64_t id)
{
	int i, n_pri;
	std::vector<int> z;
	if (n == 0) return 0;
	for (i = n_pri = 0; i < n; ++i) {
		a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = mix64(id + i);
		if (!a[i].is_alt) ++n_pri;
	}
	tie_sort((size_t)n, a, [](const mem_alnreg_t &x, const mem_alnreg_t &y) {
		return x.score > y.score || (x.score == y.score && (x.is_alt < y.is_alt || (x.is_alt == y.is_alt && x.hash < y.hash)));
	});
	mark_primary_core(opt, n, a, z);
	for (i = 0; i < n; ++i) {
		mem_alnreg_t *p = &a[i];
		p->secondary_all = i;
		if (!p->is_alt && p->secondary >= 0 && a[p->secondary].is_alt)
			p->alt_sc = a[p->secondary].score;
	}
	if (n_pri >= 0 && n_pri < n) {
		z.resize(n);
		if (n_pri > 0)
			tie_sort((size_t)n, a, [](const mem_alnreg_t &x, const mem_alnreg_t &y) {
				return x.is_alt < y.is_alt || (x.is_alt == y.is_alt && (x.score > y.score || (x.score == y.score && x.hash < y.hash)));
			});
		for (i = 0; i < n; ++i) z[a[i].secondary_all] = i;
		for (i = 0; i < n; ++i) {
			if (a[i].secondary >= 0) {
				a[i].secondary_all = z[a[i].secondary];
				if (a[i].is_alt) a[i].secondary = INT_MAX;
			} else a[i].secondary_all = -1;
		}
		if (n_pri > 0) {
			for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
			mark_primary_core(opt, n_pri, a, z);
		}
	} else {
		for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
	}
	return n_pri;
}

void reorder_primary5(int T, RegVec &a)
{
	int n_pri = 0, left_st = INT_MAX, left_k = -1;
	for (size_t k = 0; k < a.size(); ++k)
		if (a[k].secondary < 0 && !a[k].is_alt && a[k].score >= T) ++n_pri;
	if (n_pri <= 1) return;
	for (size_t k = 0; k < a.size(); ++k) {
		mem_alnreg_t *p = &a[k];
		if (p->secondary >= 0 || p->is_alt || p->score < T) continue;
		if (p->qb < left_st) { left_st = p->qb; left_k = (int)k; }
	}
	if (left_k == 0) return;
	std::swap(a[0], a[left_k]);
	for (size_t k = 1; k < a.size(); ++k) {
		mem_alnreg_t *p = &a[k];
		if (p->secondary == 0) p->secondary = left_k;
		else if (p->secondary == left_k) p->secondary = 0;
		if (p->secondary_all == 0) p->secondary_all = left_k;
		else if (p->secondary_all == left_k) p->secondary_all = 0;
	}
}

int approx_mapq_se(const mem_opt_t *opt, const mem_alnreg_t *a)
{
	int mapq, l, sub = a->sub ? a->sub : opt->min_seed_len * opt->a;
	double identity;
	sub = a->csub > sub ? a->csub : sub;
	if (sub >= a->score) return 0;
	l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int)(a->re - a->rb);
	identity = 1. - (double)(l * opt->a - a->score) / (opt->a + opt->b) / l;
	if (a->score == 0) {
		mapq = 0;
	} else if (opt->mapQ_coef_len > 0) {
		double tmp;
		tmp = l < opt->mapQ_coef_len ? 1. : opt->mapQ_coef_fac / log(l);
		tmp *= identity * identity;
		mapq = (int)(6.02 * (a->score - sub) / opt->a * tmp * tmp + .499);
	} else {
		mapq = (int)(30.0 * (1. - (double)sub / a->score) * log(a->seedcov) + .499);
		mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
	}
	if (a->sub_n > 0) mapq -= (int)(4.343 * log(a->sub_n + 1) + .499);
	if (mapq > 60) mapq = 60;
	if (mapq < 0) mapq = 0;
	mapq = (int)(mapq * (1. - a->frac_rep) + .499);
	return mapq;
}

/* ------------------------------------------------------------------ paired-end statistics and pairing */

int infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
	int64_t p2;
	int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
	p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
	*dist = p2 > b1 ? p2 - b1 : b1 - p2;
	return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

static int unique_sub(const mem_opt_t *opt, const RegVec &r)
{
	size_t j;
	for (j = 1; j < r.size(); ++j) {
		int b_max = r[j].qb > r[0].qb ? r[j].qb : r[0].qb;
		int e_min = r[j].qe < r[0].qe ? r[j].qe : r[0].qe;
		if (e_min > b_max) {
			int min_l = r[j].qe - r[j].qb < r[0].qe - r[0].qb ? r[j].qe - r[j].qb : r[0].qe - r[0].qb;
			if (e_min - b_max >= min_l * opt->mask_level) break;
		}
	}
	return j < r.size() ? r[j].score : opt->min_seed_len * opt->a;
}

void pestat(const mem_opt_t *opt, int64_t l_pac, int n, const RegVec *regs, mem_pestat_t pes[4])
{
	int d, max;
	std::vector<uint64_t> isize[4];
	memset(pes, 0, 4 * sizeof(mem_pestat_t));
	{	// candidate insert sizes, gathered by the worker threads; the order does not matter because every use below goes
		// through the sorted array
		const int nt = opt->n_threads > 0 ? opt->n_threads : 1;
		std::vector<std::vector<uint64_t>> part((size_t)nt * 4);
		parallel_for(nt, n >> 1, 4096, [&](int tid, int64_t b, int64_t e) {
			for (int64_t i = b; i < e; ++i) {
				int64_t is;
				const RegVec &r0 = regs[i << 1 | 0], &r1 = regs[i << 1 | 1];
				if (r0.empty() || r1.empty()) continue;
				if (unique_sub(opt, r0) > 0.8 * r0[0].score) continue;
				if (unique_sub(opt, r1) > 0.8 * r1[0].score) continue;
				if (r0[0].rid != r1[0].rid) continue;
				const int dir = infer_dir(l_pac, r0[0].rb, r1[0].rb, &is);
				if (is && is <= opt->max_ins) part[(size_t)tid * 4 + dir].push_back(is);
			}
		});
		for (d = 0; d < 4; ++d)
			for (int t = 0; t < nt; ++t) isize[d].insert(isize[d].end(), part[(size_t)t * 4 + d].begin(), part[(size_t)t * 4 + d].end());
	}
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

struct Pair64 { uint64_t x, y; };
static inline bool pair_lt(const Pair64 &a, const Pair64 &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }

int pair_ends(const mem_opt_t *opt, const bntseq_t *bns, const mem_pestat_t pes[4], RegVec a[2], int id,
              int *sub, int *n_sub, int z[2], int n_pri[2])
{
	std::vector<Pair64> v, u;
	int r, i, k, y[4], ret;
	int64_t l_pac = bns->l_pac;
	for (r = 0; r < 2; ++r) {
		for (i = 0; i < n_pri[r]; ++i) {
			Pair64 key;
			mem_alnreg_t *e = &a[r][i];
			key.x = e->rb < l_pac ? e->rb : (l_pac << 1) - 1 - e->rb;
			key.x = (uint64_t)e->rid << 32 | (key.x - bns->anns[e->rid].offset);
			key.y = (uint64_t)e->score << 32 | i << 2 | (e->rb >= l_pac) << 1 | r;
			v.push_back(key);
		}
	}
	tie_sort(v, pair_lt);
	y[0] = y[1] = y[2] = y[3] = -1;
	for (i = 0; i < (int)v.size(); ++i) {
		for (r = 0; r < 2; ++r) {
			int dir = r << 1 | (v[i].y >> 1 & 1), which;
			if (pes[dir].failed) continue;
			which = r << 1 | ((v[i].y & 1) ^ 1);
			if (y[which] < 0) continue;
			for (k = y[which]; k >= 0; --k) {
				int64_t dist;
				int q;
				double ns;
				if ((int)(v[k].y & 3) != which) continue;
				dist = (int64_t)v[i].x - v[k].x;
				if (dist > pes[dir].high) break;
				if (dist < pes[dir].low) continue;
				ns = (dist - pes[dir].avg) / pes[dir].std;
				q = (int)((v[i].y >> 32) + (v[k].y >> 32) + .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt->a + .499);
				if (q < 0) q = 0;
				Pair64 p;
				p.y = (uint64_t)k << 32 | i;
				p.x = (uint64_t)q << 32 | (mix64(p.y ^ id << 8) & 0xffffffffU);
				u.push_back(p);
			}
		}
		y[v[i].y & 3] = i;
	}
	if (!u.empty()) {
		int tmp = opt->a + opt->b;
		tmp = tmp > opt->o_del + opt->e_del ? tmp : opt->o_del + opt->e_del;
		tmp = tmp > opt->o_ins + opt->e_ins ? tmp : opt->o_ins + opt->e_ins;
		tie_sort(u, pair_lt);
		size_t un = u.size();
		i = (int)(u[un - 1].y >> 32); k = (int)(u[un - 1].y << 32 >> 32);
		z[v[i].y & 1] = (int)(v[i].y << 32 >> 34);
		z[v[k].y & 1] = (int)(v[k].y << 32 >> 34);
		ret = (int)(u[un - 1].x >> 32);
		*sub = un > 1 ? (int)(u[un - 2].x >> 32) : 0;
		for (i = (int)((long)un - 2), *n_sub = 0; i >= 0; --i)
			if (*sub - (int)(u[i].x >> 32) <= tmp) ++*n_sub;
	} else { ret = 0; *sub = 0; *n_sub = 0; }
	return ret;
}

/* ------------------------------------------------------------------ region -> alignment record -> SAM */

static inline int infer_bw(int l1, int l2, int score, int a, int q, int r)
{
	int w;
	if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
	w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
	if (w < abs(l1 - l2)) w = abs(l1 - l2);
	return w;
}

static int reg_first_band(const mem_opt_t *opt, const mem_alnreg_t *ar)       // reference src/bwamem.c:1107-1111
{
	int tmp = infer_bw(ar->qe - ar->qb, (int)(ar->re - ar->rb), ar->truesc, opt->a, opt->o_del, opt->e_del);
	int w2 = infer_bw(ar->qe - ar->qb, (int)(ar->re - ar->rb), ar->truesc, opt->a, opt->o_ins, opt->e_ins);
	w2 = w2 > tmp ? w2 : tmp;
	if (w2 > opt->w) w2 = w2 < ar->w ? w2 : ar->w;
	return w2;
}

bool reg_global_job(const mem_opt_t *opt, const bntseq_t *bns, const mem_alnreg_t *ar, int read, GlobalJob *j)
{
	if (ar->rb < 0 || ar->re < 0) return false;
	const int w2 = reg_first_band(opt, ar);
	if (!(ar->qe > ar->qb && ar->re <= bns->l_pac << 1 && ar->rb < ar->re && !(ar->rb < bns->l_pac && ar->re > bns->l_pac) &&
	      global_needs_dp(ar->qe - ar->qb, ar->re - ar->rb, w2 < opt->w << 2 ? w2 : opt->w << 2))) return false;
	j->rb = ar->rb; j->re = ar->re; j->zoff = 0; j->qb = ar->qb; j->qe = ar->qe; j->w2 = w2; j->truesc = ar->truesc; j->wmax = 0;
	j->read = read;
	return true;
}

void reg2aln(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_query, const char *query_,
             const mem_alnreg_t *ar, Aln *out)
{
	ProfScope ps(HP_REG2ALN);
	Aln &a = *out;
	a = Aln();
	a.pos = 0; a.rid = 0;
	if (ar == 0 || ar->rb < 0 || ar->re < 0) {
		a.rid = -1; a.pos = -1; a.flag |= 0x4;
		return;
	}
	int i, w2, qb, qe, NM = -1, score = 0, is_rev, last_sc = -(1 << 30);
	int64_t pos, rb, re;
	qb = ar->qb; qe = ar->qe;
	rb = ar->rb; re = ar->re;
	// gen_cigar reverses the query of a reverse-strand region in place (and restores it), so those work on a copy; an
	// already encoded read (first byte a code: all are) with a forward-strand region is used where it lies
	uint8_t query_stack[512];
	std::vector<uint8_t> query_heap;
	uint8_t *query = query_stack;
	if (l_query > 0 && (uint8_t)query_[0] < 5 && rb < bns->l_pac) query = (uint8_t *)const_cast<char *>(query_);
	else {
		if (l_query > (int)sizeof query_stack) { query_heap.resize(l_query); query = query_heap.data(); }
		for (i = 0; i < l_query; ++i)
			query[i] = query_[i] < 5 ? query_[i] : kNt4[(uint8_t)query_[i]];
	}
	a.mapq = ar->secondary < 0 ? (approx_mapq_se(opt, ar) & 0xff) : 0;
	if (ar->secondary >= 0) a.flag |= 0x100;
	w2 = reg_first_band(opt, ar);
	AlignCtx &cx = align_ctx();
	bool done = false;
	GlobalJob want;
	if (cx.mode == AlignCtx::LOOKUP && reg_global_job(opt, bns, ar, query_ == cx.seq_ptr[0] ? cx.read_idx[0] : cx.read_idx[1], &want)) {
		for (int k = 0; k < cx.n_jobs; ++k) {
			const GlobalJob &j = cx.jobs[k];
			if (j.read != want.read || j.rb != want.rb || j.re != want.re || j.qb != want.qb || j.qe != want.qe || j.w2 != want.w2 || j.truesc != want.truesc) continue;
			const GlobalRes &g = cx.res[k];
			if (g.n_cigar >= 0) {
				gen_cigar(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, 0, bns->l_pac, pac, qe - qb,
				          &query[qb], rb, re, &score, &a.cigar, &NM, &a.md, &g);
				done = true;
			}                                           // else: too many CIGAR operations for the result record - align here
			break;
		}
	}
	i = 0;
	if (!done && cx.mode == AlignCtx::LOOKUP && global_needs_dp(qe - qb, re - rb, w2 < opt->w << 2 ? w2 : opt->w << 2)) ++cx.n_host_dp;
	if (!done) do {
		w2 = w2 < opt->w << 2 ? w2 : opt->w << 2;
		gen_cigar(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w2, bns->l_pac, pac, qe - qb,
		          &query[qb], rb, re, &score, &a.cigar, &NM, &a.md, nullptr);
		if (score == last_sc || w2 == opt->w << 2) break;
		last_sc = score;
		w2 <<= 1;
	} while (++i < 3 && score < ar->truesc - opt->a);
	a.NM = (uint32_t)NM & 0x3fffff;
	pos = bns_depos_h(bns, rb < bns->l_pac ? rb : re - 1, &is_rev);
	a.is_rev = is_rev;
	if (!a.cigar.empty()) {
		if ((a.cigar[0] & 0xf) == 2) {
			pos += a.cigar[0] >> 4;
			a.cigar.erase(a.cigar.begin());
		} else if ((a.cigar.back() & 0xf) == 2) {
			a.cigar.pop_back();
		}
	}
	if (qb != 0 || qe != l_query) {
		int clip5 = is_rev ? l_query - qe : qb;
		int clip3 = is_rev ? qb : l_query - qe;
		if (clip5) a.cigar.insert(a.cigar.begin(), (uint32_t)clip5 << 4 | 3);
		if (clip3) a.cigar.push_back((uint32_t)clip3 << 4 | 3);
	}
	a.rid = bns_pos2rid_h(bns, pos);
	a.pos = pos - bns->anns[a.rid].offset;
	a.score = ar->score; a.sub = ar->sub > ar->csub ? ar->sub : ar->csub;
	a.is_alt = ar->is_alt; a.alt_sc = ar->alt_sc;
}

static inline int pri_idx(double XA_drop_ratio, const mem_alnreg_t *a, int i)
{
	int k = a[i].secondary_all;
	if (k >= 0 && a[i].score >= a[k].score * XA_drop_ratio) return k;
	return -1;
}

bool gen_alt(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const RegVec &a, int l_query,
             const char *query, std::vector<std::string> &XA)
{
	ProfScope ps(HP_GEN_ALT);
	const int n = (int)a.size();
	int i, r, tot = 0;
	XA.clear();
	// most reads have no secondary hit within XA_drop_ratio of its primary: find that out before allocating anything
	for (i = 0; i < n; ++i) if (pri_idx(opt->XA_drop_ratio, a.data(), i) >= 0) break;
	if (i == n) return false;
	std::vector<int> cnt(n, 0);
	std::vector<char> has_alt(n, 0);
	for (i = 0; i < n; ++i) {
		r = pri_idx(opt->XA_drop_ratio, a.data(), i);
		if (r >= 0) {
			++cnt[r]; ++tot;
			if (a[i].is_alt) has_alt[r] = 1;
		}
	}
	if (tot == 0) return false;
	XA.assign(n, std::string());
	for (i = 0; i < n; ++i) {
		if ((r = pri_idx(opt->XA_drop_ratio, a.data(), i)) < 0) continue;
		if (cnt[r] > opt->max_XA_hits_alt || (!has_alt[r] && cnt[r] > opt->max_XA_hits)) continue;
		Aln t;
		reg2aln(opt, bns, pac, l_query, query, &a[i], &t);
		std::string &s = XA[r];
		s += bns->anns[t.rid].name;
		s.push_back(','); s.push_back("+-"[t.is_rev]); put_int(s, t.pos + 1);
		s.push_back(',');
		for (uint32_t c : t.cigar) { put_int(s, c >> 4); s.push_back("MIDSHN"[c & 0xf]); }
		s.push_back(','); put_int(s, (int)t.NM);
		s.push_back(';');
	}
	return true;
}

static inline int ref_len_of(const Aln &p)
{
	int l = 0;
	for (uint32_t c : p.cigar) { int op = c & 0xf; if (op == 0 || op == 2) l += c >> 4; }
	return l;
}

struct AlnView {   // scalar copy of an Aln that aln2sam may edit (the reference edits a struct copy)
	int64_t pos; int rid, flag; uint32_t is_rev, is_alt, mapq, NM; int n_cigar; const Aln *src; int score, sub, alt_sc;
	explicit AlnView(const Aln &a) : pos(a.pos), rid(a.rid), flag(a.flag), is_rev(a.is_rev), is_alt(a.is_alt), mapq(a.mapq),
		NM(a.NM), n_cigar((int)a.cigar.size()), src(&a), score(a.score), sub(a.sub), alt_sc(a.alt_sc) {}
};

static inline void put_cigar(const mem_opt_t *opt, const AlnView &p, std::string &str, int which)
{
	if (p.n_cigar) {
		for (int i = 0; i < p.n_cigar; ++i) {
			uint32_t v = p.src->cigar[i];
			int c = v & 0xf;
			if (!(opt->flag & MEM_F_SOFTCLIP) && !p.is_alt && (c == 3 || c == 4))
				c = which ? 4 : 3;
			put_int(str, v >> 4); str.push_back("MIDSH"[c]);
		}
	} else str.push_back('*');
}

void aln2sam(const mem_opt_t *opt, const bntseq_t *bns, std::string &str, const bseq1_t *s, int n, const Aln *list,
             int which, const Aln *m_)
{
	ProfScope ps(HP_ALN2SAM);
	AlnView p(list[which]);
	AlnView mt(m_ ? *m_ : list[which]);
	AlnView *m = m_ ? &mt : nullptr;
	p.flag |= m ? 0x1 : 0;
	p.flag |= p.rid < 0 ? 0x4 : 0;
	p.flag |= m && m->rid < 0 ? 0x8 : 0;
	if (p.rid < 0 && m && m->rid >= 0) { p.rid = m->rid; p.pos = m->pos; p.is_rev = m->is_rev; p.n_cigar = 0; }
	if (m && m->rid < 0 && p.rid >= 0) { m->rid = p.rid; m->pos = p.pos; m->is_rev = p.is_rev; m->n_cigar = 0; }
	p.flag |= p.is_rev ? 0x10 : 0;
	p.flag |= m && m->is_rev ? 0x20 : 0;

	str += s->name; str.push_back('\t');
	put_int(str, (p.flag & 0xffff) | (p.flag & 0x10000 ? 0x100 : 0)); str.push_back('\t');
	if (p.rid >= 0) {
		str += bns->anns[p.rid].name; str.push_back('\t');
		put_int(str, p.pos + 1); str.push_back('\t');
		put_int(str, (int)p.mapq); str.push_back('\t');
		put_cigar(opt, p, str, which);
	} else str += "*\t0\t0\t*";
	str.push_back('\t');

	if (m && m->rid >= 0) {
		if (p.rid == m->rid) str.push_back('=');
		else str += bns->anns[m->rid].name;
		str.push_back('\t');
		put_int(str, m->pos + 1); str.push_back('\t');
		if (p.rid == m->rid) {
			// n_cigar may have been zeroed above; the reference then sums over zero operations
			int64_t p0 = p.pos + (p.is_rev ? (p.n_cigar ? ref_len_of(*p.src) : 0) - 1 : 0);
			int64_t p1 = m->pos + (m->is_rev ? (m->n_cigar ? ref_len_of(*m->src) : 0) - 1 : 0);
			if (m->n_cigar == 0 || p.n_cigar == 0) str.push_back('0');
			else put_int(str, -(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0)));
		} else str.push_back('0');
	} else str += "*\t0\t0";
	str.push_back('\t');

	const uint32_t *cig = p.src->cigar.data();
	if (p.flag & 0x100) {
		str += "*\t*";
	} else if (!p.is_rev) {
		int qb = 0, qe = s->l_seq;
		if (p.n_cigar && which && !(opt->flag & MEM_F_SOFTCLIP) && !p.is_alt) {
			if ((cig[0] & 0xf) == 4 || (cig[0] & 0xf) == 3) qb += cig[0] >> 4;
			if ((cig[p.n_cigar - 1] & 0xf) == 4 || (cig[p.n_cigar - 1] & 0xf) == 3) qe -= cig[p.n_cigar - 1] >> 4;
		}
		if (qe > qb) {
			const size_t at = str.size();
			str.resize(at + (size_t)(qe - qb));
			char *d = &str[at];
			for (int i = qb; i < qe; ++i) *d++ = "ACGTN"[(int)s->seq[i]];
		}
		str.push_back('\t');
		if (s->qual) str.append(s->qual + qb, qe > qb ? qe - qb : 0);
		else str.push_back('*');
	} else {
		int qb = 0, qe = s->l_seq;
		if (p.n_cigar && which && !(opt->flag & MEM_F_SOFTCLIP) && !p.is_alt) {
			if ((cig[0] & 0xf) == 4 || (cig[0] & 0xf) == 3) qe -= cig[0] >> 4;
			if ((cig[p.n_cigar - 1] & 0xf) == 4 || (cig[p.n_cigar - 1] & 0xf) == 3) qb += cig[p.n_cigar - 1] >> 4;
		}
		const size_t nq = qe > qb ? (size_t)(qe - qb) : 0;
		size_t at = str.size();
		str.resize(at + nq);
		char *d = &str[at];
		for (int i = qe - 1; i >= qb; --i) *d++ = "TGCAN"[(int)s->seq[i]];
		str.push_back('\t');
		if (s->qual) {
			at = str.size();
			str.resize(at + nq);
			d = &str[at];
			for (int i = qe - 1; i >= qb; --i) *d++ = s->qual[i];
		} else str.push_back('*');
	}

	if (p.n_cigar) {
		str += "\tNM:i:"; put_int(str, (int)p.NM);
		str += "\tMD:Z:"; str += p.src->md;
	}
	if (m && m->n_cigar) { str += "\tMC:Z:"; put_cigar(opt, *m, str, which); }
	if (p.score >= 0) { str += "\tAS:i:"; put_int(str, p.score); }
	if (p.sub >= 0) { str += "\tXS:i:"; put_int(str, p.sub); }
	if (bwa_rg_id[0]) { str += "\tRG:Z:"; str += bwa_rg_id; }
	if (!(p.flag & 0x100)) {
		int i;
		for (i = 0; i < n; ++i)
			if (i != which && !(list[i].flag & 0x100)) break;
		if (i < n) {
			str += "\tSA:Z:";
			for (i = 0; i < n; ++i) {
				const Aln *r = &list[i];
				if (i == which || (r->flag & 0x100)) continue;
				str += bns->anns[r->rid].name; str.push_back(',');
				put_int(str, r->pos + 1); str.push_back(',');
				str.push_back("+-"[r->is_rev]); str.push_back(',');
				for (uint32_t c : r->cigar) { put_int(str, c >> 4); str.push_back("MIDSH"[c & 0xf]); }
				str.push_back(','); put_int(str, (int)r->mapq);
				str.push_back(','); put_int(str, (int)r->NM);
				str.push_back(';');
			}
		}
		if (p.alt_sc > 0) {
			char buf[64];
			snprintf(buf, sizeof buf, "\tpa:f:%.3f", (double)p.score / p.alt_sc);
			str += buf;
		}
	}
	if (p.src->XA && !p.src->XA->empty()) { str += "\tXA:Z:"; str += *p.src->XA; }
	if (s->comment) { str.push_back('\t'); str += s->comment; }
	if ((opt->flag & MEM_F_REF_HDR) && p.rid >= 0 && bns->anns[p.rid].anno != 0 && bns->anns[p.rid].anno[0] != 0) {
		str += "\tXR:Z:";
		size_t from = str.size();
		str += bns->anns[p.rid].anno;
		for (size_t i = from; i < str.size(); ++i)
			if (str[i] == '\t') str[i] = ' ';
	}
	str.push_back('\n');
}

void reg2sam(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, bseq1_t *s, RegVec &a, int extra_flag,
             const Aln *m)
{
	ProfScope ps(HP_REG2SAM);
	std::string local, *sink = align_ctx().sink;
	std::string &str = sink ? *sink : local;
	std::vector<Aln> aa;
	std::vector<std::string> XA;
	bool has_xa = false;
	int l = 0;
	if (!(opt->flag & MEM_F_ALL))
		has_xa = gen_alt(opt, bns, pac, a, s->l_seq, s->seq, XA);
	aa.reserve(a.size());
	for (size_t k = 0; k < a.size(); ++k) {
		mem_alnreg_t *p = &a[k];
		if (p->score < opt->T) continue;
		if (p->secondary >= 0 && (p->is_alt || !(opt->flag & MEM_F_ALL))) continue;
		if (p->secondary >= 0 && p->secondary < INT_MAX && p->score < a[p->secondary].score * opt->drop_ratio) continue;
		aa.emplace_back();
		Aln *q = &aa.back();
		reg2aln(opt, bns, pac, s->l_seq, s->seq, p, q);
		q->XA = has_xa ? &XA[k] : nullptr;
		q->flag |= extra_flag;
		if (p->secondary >= 0) q->sub = -1;
		if (l && p->secondary < 0)
			q->flag |= (opt->flag & MEM_F_NO_MULTI) ? 0x10000 : 0x800;
		if (!(opt->flag & MEM_F_KEEP_SUPP_MAPQ) && l && !p->is_alt && q->mapq > aa[0].mapq)
			q->mapq = aa[0].mapq;
		++l;
	}
	if (aa.empty()) {
		Aln t;
		reg2aln(opt, bns, pac, s->l_seq, s->seq, 0, &t);
		t.flag |= extra_flag;
		aln2sam(opt, bns, str, s, 1, &t, 0, m);
	} else {
		for (size_t k = 0; k < aa.size(); ++k)
			aln2sam(opt, bns, str, s, (int)aa.size(), aa.data(), (int)k, m);
	}
	s->sam = sink ? nullptr : dup_cstr(str);
}

#define RAW_MAPQ(diff, a) ((int)(6.02 * (diff) / (a) + .499))

void sam_pe_finish(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4],
                   uint64_t id, bseq1_t s[2], RegVec a[2])
{
	ProfScope ps(HP_SAM_PE);
	int i, j, z[2], o, subo, n_sub, extra_flag = 1, n_pri[2];
	Aln h[2], g[2];
	std::vector<Aln> aa[2];
	{
		ProfScope pm(HP_MARK_PRIMARY);
		n_pri[0] = mark_primary_se(opt, (int)a[0].size(), a[0].data(), id << 1 | 0);
		n_pri[1] = mark_primary_se(opt, (int)a[1].size(), a[1].data(), id << 1 | 1);
	}
	if (opt->flag & MEM_F_PRIMARY5) {
		reorder_primary5(opt->T, a[0]);
		reorder_primary5(opt->T, a[1]);
	}
	bool paired_out = false;
	o = 0;
	if (!(opt->flag & MEM_F_NOPAIRING) && n_pri[0] && n_pri[1]) { ProfScope pp(HP_PAIR); o = pair_ends(opt, bns, pes, a, (int)id, &subo, &n_sub, z, n_pri); }
	if (o > 0) {
		int is_multi[2], q_pe, score_un, q_se[2];
		for (i = 0; i < 2; ++i) {
			for (j = 1; j < n_pri[i]; ++j)
				if (a[i][j].secondary < 0 && a[i][j].score >= opt->T) break;
			is_multi[i] = j < n_pri[i] ? 1 : 0;
		}
		if (!(is_multi[0] || is_multi[1])) {
			paired_out = true;
			score_un = a[0][0].score + a[1][0].score - opt->pen_unpaired;
			subo = subo > score_un ? subo : score_un;
			q_pe = RAW_MAPQ(o - subo, opt->a);
			if (n_sub > 0) q_pe -= (int)(4.343 * log(n_sub + 1) + .499);
			if (q_pe < 0) q_pe = 0;
			if (q_pe > 60) q_pe = 60;
			q_pe = (int)(q_pe * (1. - .5 * (a[0][0].frac_rep + a[1][0].frac_rep)) + .499);
			if (o > score_un) {
				mem_alnreg_t *c[2];
				c[0] = &a[0][z[0]]; c[1] = &a[1][z[1]];
				for (i = 0; i < 2; ++i) {
					if (c[i]->secondary >= 0) { c[i]->sub = a[i][c[i]->secondary].score; c[i]->secondary = -2; }
					q_se[i] = approx_mapq_se(opt, c[i]);
				}
				q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
				q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
				extra_flag |= 2;
				q_se[0] = q_se[0] < RAW_MAPQ(c[0]->score - c[0]->csub, opt->a) ? q_se[0] : RAW_MAPQ(c[0]->score - c[0]->csub, opt->a);
				q_se[1] = q_se[1] < RAW_MAPQ(c[1]->score - c[1]->csub, opt->a) ? q_se[1] : RAW_MAPQ(c[1]->score - c[1]->csub, opt->a);
			} else {
				z[0] = z[1] = 0;
				q_se[0] = approx_mapq_se(opt, &a[0][0]);
				q_se[1] = approx_mapq_se(opt, &a[1][0]);
			}
			for (i = 0; i < 2; ++i) {
				int k = a[i][z[i]].secondary_all;
				if (k >= 0 && k < n_pri[i]) {
					for (j = 0; j < (int)a[i].size(); ++j)
						if (a[i][j].secondary_all == k || j == k)
							a[i][j].secondary_all = z[i];
					a[i][z[i]].secondary_all = -1;
				}
			}
			std::vector<std::string> XA[2];
			bool has_xa[2] = { false, false };
			if (!(opt->flag & MEM_F_ALL))
				for (i = 0; i < 2; ++i)
					has_xa[i] = gen_alt(opt, bns, pac, a[i], s[i].l_seq, s[i].seq, XA[i]);
			for (i = 0; i < 2; ++i) {
				reg2aln(opt, bns, pac, s[i].l_seq, s[i].seq, &a[i][z[i]], &h[i]);
				h[i].mapq = (uint32_t)q_se[i] & 0xff;
				h[i].flag |= 0x40 << i | extra_flag;
				h[i].XA = has_xa[i] ? &XA[i][z[i]] : nullptr;
				if (n_pri[i] < (int)a[i].size()) {
					mem_alnreg_t *p = &a[i][n_pri[i]];
					if (p->score < opt->T || p->secondary >= 0 || !p->is_alt) continue;
					reg2aln(opt, bns, pac, s[i].l_seq, s[i].seq, p, &g[i]);
					g[i].flag |= 0x800 | 0x40 << i | extra_flag;
					g[i].XA = has_xa[i] ? &XA[i][n_pri[i]] : nullptr;
					aa[i].push_back(h[i]);                  // (only an ALT supplementary record needs the two-entry list)
					aa[i].push_back(std::move(g[i]));
				}
			}
			// records go straight into the block buffer of the sweep when there is one, else into seqs[i].sam
			std::string local, *sink = align_ctx().sink;
			std::string &str = sink ? *sink : local;
			if (!sink) str.reserve(640);
			for (i = 0; i < 2; ++i) {
				if (aa[i].empty()) aln2sam(opt, bns, str, &s[i], 1, &h[i], 0, &h[!i]);
				else for (j = 0; j < (int)aa[i].size(); ++j) aln2sam(opt, bns, str, &s[i], (int)aa[i].size(), aa[i].data(), j, &h[!i]);
				if (!sink) { s[i].sam = dup_cstr(str); str.clear(); } else s[i].sam = nullptr;
			}
			if (strcmp(s[0].name, s[1].name) != 0) {
				fprintf(stderr, "[mem_sam_pe] paired reads have different names: \"%s\", \"%s\"\n", s[0].name, s[1].name);
				abort();
			}
		}
	}
	if (paired_out) return;
	// no_pairing
	for (i = 0; i < 2; ++i) {
		int which = -1;
		if (!a[i].empty()) {
			if (a[i][0].score >= opt->T) which = 0;
			else if (n_pri[i] < (int)a[i].size() && a[i][n_pri[i]].score >= opt->T)
				which = n_pri[i];
		}
		if (which >= 0) reg2aln(opt, bns, pac, s[i].l_seq, s[i].seq, &a[i][which], &h[i]);
		else reg2aln(opt, bns, pac, s[i].l_seq, s[i].seq, 0, &h[i]);
	}
	if (!(opt->flag & MEM_F_NOPAIRING) && h[0].rid == h[1].rid && h[0].rid >= 0) {
		int64_t dist;
		int d = infer_dir(bns->l_pac, a[0][0].rb, a[1][0].rb, &dist);
		if (!pes[d].failed && dist >= pes[d].low && dist <= pes[d].high) extra_flag |= 2;
	}
	reg2sam(opt, bns, pac, &s[0], a[0], 0x41 | extra_flag, &h[1]);
	reg2sam(opt, bns, pac, &s[1], a[1], 0x81 | extra_flag, &h[0]);
	if (strcmp(s[0].name, s[1].name) != 0) {
		fprintf(stderr, "[mem_sam_pe] paired reads have different names: \"%s\", \"%s\"\n", s[0].name, s[1].name);
		abort();
	}
}

} // namespace b200
