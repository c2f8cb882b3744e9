// smem_kernel.cuh - SMEM seeding (mem_collect_intv, reference src/bwamem.c:114-162 over bwt_smem1a / bwt_seed_strategy1,
// src/bwt.c:289-379) as a per-lane state machine whose only expensive step is ONE bwt_extend per loop iteration.
//
// Why a state machine: the three seeding passes of a read are a few hundred *dependent* bi-directional extensions
// (each = two 64-byte occ blocks of the FM-index, fetched with independent 128-bit loads) wrapped in irregular list
// logic.  Running the reference's nested loops one read per thread leaves every lane of a warp in a different loop
// nest (v1: 8.8 of 32 lanes active per instruction) and forces the per-read interval lists into thread-local arrays
// in HBM (v1: 47.8 GB written per chunk).  Here
//   * every lane owns one read and a small explicit state; each trip of the warp loop issues exactly one extension for
//     every live lane (the block loads of all lanes are in flight together), then a short state transition;
//   * the interval list of bwt_smem1a lives in shared memory.  ONE list suffices: the backward sweep compacts `prev`
//     into `curr` in place (the write cursor never passes the read cursor) and the list is kept top-aligned and walked
//     downwards so that the reference's "reverse curr" step disappears; entries are packed to 16 bytes (three 33-bit
//     interval words + the end position) and laid out [entry][word][thread] (bank = lane).  Entries beyond the shared
//     quota spill to a per-thread global strip (only very repetitive reads get there);
//   * finished lanes fetch the next read from a global counter (persistent grid, no tail of idle lanes);
//   * SMEMs are appended to the read's output strip as they are found; the sort by (start,end) that closes
//     mem_collect_intv is done by rank in k_compact_intv (equal keys are identical intervals).
//
// The code is host/device so that tests/hostemu can run the very same state machine on the CPU against the
// straightforward restatement in fm_kernels.h (and through it against the oracle).
#pragma once
#include "fm_kernels.h"

namespace b200 {

struct alignas(16) Q4 { uint32_t x, y, z, w; };

B200_HD Q4 ld_q4(const uint32_t *p)
{
#if defined(__CUDA_ARCH__)
	const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
	Q4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
	return r;
#else
	return *reinterpret_cast<const Q4 *>(p);
#endif
}

// fm.L2[i] without a dynamically indexed kernel parameter (which would be copied to local memory)
B200_HD uint64_t l2_at(const FmView &fm, int i)
{
	return i == 0 ? fm.L2[0] : i == 1 ? fm.L2[1] : i == 2 ? fm.L2[2] : i == 3 ? fm.L2[3] : fm.L2[4];
}

// the two occ sectors a backward (is_back) or forward extension of the interval reads
B200_HD void fm_extend_load(const FmView &fm, uint64_t x0, uint64_t x1, uint64_t x2, int is_back, OccRaw &rk, OccRaw &rl)
{
	const uint64_t base = is_back ? x0 : x1;
	const uint64_t k = base - 1, l = base - 1 + x2;
	const uint64_t ka = k == (uint64_t)-1 ? 0 : k - (k >= fm.primary), la = l == (uint64_t)-1 ? 0 : l - (l >= fm.primary);
	rk = ld_occ(fm, ka >> 6); rl = ld_occ(fm, la >> 6);
}

// bwt_extend (reference src/bwt.c:262-275) returning only the interval of base c, from the two occ sectors (one 256-bit load each)
B200_HD void fm_extend_use(const FmView &fm, uint64_t x0, uint64_t x1, uint64_t x2, int is_back, int c, const OccRaw &rk, const OccRaw &rl,
                           uint64_t &o0, uint64_t &o1, uint64_t &o2, int64_t &n_blocks)
{
	const uint64_t base = is_back ? x0 : x1, other = is_back ? x1 : x0;
	const uint64_t k = base - 1, l = base - 1 + x2;
	const bool kz = k == (uint64_t)-1, lz = l == (uint64_t)-1;
	const uint64_t ka = kz ? 0 : k - (k >= fm.primary), la = lz ? 0 : l - (l >= fm.primary);
	uint64_t tk[4], tl[4];
	occ4_sector(rk, ka, tk);
	occ4_sector(rl, la, tl);
	if (kz) tk[0] = tk[1] = tk[2] = tk[3] = 0;
	if (lz) tl[0] = tl[1] = tl[2] = tl[3] = 0;
	n_blocks += (kz ? 0 : 1) + ((!lz && (kz || (la >> 7) != (ka >> 7))) ? 1 : 0);
	const uint64_t s1 = tl[1] - tk[1], s2 = tl[2] - tk[2], s3 = tl[3] - tk[3];
	const uint64_t tkc = c == 0 ? tk[0] : c == 1 ? tk[1] : c == 2 ? tk[2] : tk[3];
	const uint64_t tlc = c == 0 ? tl[0] : c == 1 ? tl[1] : c == 2 ? tl[2] : tl[3];
	const uint64_t nb = l2_at(fm, c) + 1 + tkc;
	uint64_t oth = other + ((base <= fm.primary && base + x2 - 1 >= fm.primary) ? 1 : 0);
	oth += (c < 3 ? s3 : 0) + (c < 2 ? s2 : 0) + (c < 1 ? s1 : 0);
	o2 = tlc - tkc;
	if (is_back) { o0 = nb; o1 = oth; } else { o1 = nb; o0 = oth; }
}

B200_HD void fm_extend_sel(const FmView &fm, uint64_t x0, uint64_t x1, uint64_t x2, int is_back, int c,
                           uint64_t &o0, uint64_t &o1, uint64_t &o2, int64_t &n_blocks)
{
	OccRaw rk, rl;
	fm_extend_load(fm, x0, x1, x2, is_back, rk, rl);
	fm_extend_use(fm, x0, x1, x2, is_back, c, rk, rl, o0, o1, o2, n_blocks);
}

/* ---------------------------------------------------------------- k-mer tables
 * The bi-interval of a pattern is a function of the pattern alone, whichever way bwt_extend (reference src/bwt.c:262-275) reached
 * it (the reference's text is its own reverse complement, so the "other" coordinate it carries along is the interval of the reverse
 * complement).  With 180 GB of HBM the intervals of EVERY pattern of up to kmax bases (15 for a human-sized reference: 22.9 GB) are
 * tabulated once when the index is uploaded, by the same extension routine, level L from level L-1, so that the sweeps can replace
 * an extension whose result is a pattern of at most kmax bases - two dependent random occ sectors - by ONE random sector that
 * depends on nothing but the read.  Entries of absent patterns hold exactly what the reference's extensions of an empty interval
 * produce (size 0 with real coordinates): the greedy pass keeps extending those (src/bwt.c:367-376).
 * An entry also carries the number of reference-layout occ blocks (the unit of SURVEY.md 8d's algorithmic traffic) that the
 * reference's forward extensions from the first base to this pattern touch, so that the traffic counter stays the reference's. */
B200_HD uint64_t ktab_off(int L) { return (((uint64_t)1 << (2 * L)) - 4) / 3; }       // entries of lengths 1 .. L-1
B200_HD uint64_t ktab_entries(int kmax) { return kmax > 0 ? ktab_off(kmax + 1) : 0; }
// default depth: the longest patterns of which a text of seq_len bases still holds most (4^kmax <= 2 * seq_len), at most 15
// (22.9 GB of table; 16 would be 91.6 GB)
inline int ktab_default_kmax(uint64_t seq_len)
{
	int k = 0;
	while (k < 15 && ((uint64_t)1 << (2 * (k + 1))) <= 2 * seq_len) ++k;
	return k;
}
B200_HD Q4 ktab_pack(uint64_t x0, uint64_t x1, uint64_t x2, int blocks)
{
	Q4 v;
	v.x = (uint32_t)x0; v.y = (uint32_t)x1; v.z = (uint32_t)x2;
	v.w = (uint32_t)blocks | (uint32_t)(x0 >> 32) << 29 | (uint32_t)(x1 >> 32) << 30 | (uint32_t)(x2 >> 32) << 31;
	return v;
}
// the sector that holds entry (L, idx) and which half of it
B200_HD const uint32_t *ktab_sector(const FmView &fm, int L, uint32_t idx, int &half)
{
	const uint64_t e = ktab_off(L) + idx;
	half = (int)(e & 1);
	return fm.ktab + ((e >> 1) << 3);
}
B200_HD void ktab_unpack(const OccRaw &r, int half, uint64_t &x0, uint64_t &x1, uint64_t &x2, int &blocks)
{
	const uint32_t a = half ? r.w[4] : r.w[0], b = half ? r.w[5] : r.w[1], c = half ? r.w[6] : r.w[2], w = half ? r.w[7] : r.w[3];
	x0 = (uint64_t)(w >> 29 & 1u) << 32 | a;
	x1 = (uint64_t)(w >> 30 & 1u) << 32 | b;
	x2 = (uint64_t)(w >> 31) << 32 | c;
	blocks = (int)(w & 0xffu);
}
// reference-layout occ blocks one bwt_extend of the interval touches (base = the coordinate it extends on); see fm_extend_use
B200_HD int fm_extend_blocks(const FmView &fm, uint64_t base, uint64_t x2)
{
	const uint64_t k = base - 1, l = base - 1 + x2;
	const bool kz = k == (uint64_t)-1, lz = l == (uint64_t)-1;
	const uint64_t ka = kz ? 0 : k - (k >= fm.primary), la = lz ? 0 : l - (l >= fm.primary);
	return (kz ? 0 : 1) + ((!lz && (kz || (la >> 7) != (ka >> 7))) ? 1 : 0);
}
// entry `idx` of level L (L >= 1) of the table from level L-1: the forward extension of the parent pattern by the base idx & 3
B200_HD Q4 ktab_make(const FmView &fm, int L, uint32_t idx)
{
	const int b = (int)(idx & 3u);
	if (L == 1) return ktab_pack(l2_at(fm, b) + 1, l2_at(fm, 3 - b) + 1, l2_at(fm, b + 1) - l2_at(fm, b), 0);
	int half, pb;
	const uint32_t *ps = ktab_sector(fm, L - 1, idx >> 2, half);
	OccRaw r;
	for (int i = 0; i < 8; ++i) r.w[i] = ps[i];
	uint64_t p0, p1, p2, o0, o1, o2;
	ktab_unpack(r, half, p0, p1, p2, pb);
	// (a symbol array that is not the transform of a text - a damaged index - can lead an interval outside the table: keep the loads inside)
	if (p1 < 1 || p1 - 1 + p2 > fm.seq_len) return ktab_pack(1, 1, 0, pb);
	int64_t nb = 0;
	fm_extend_sel(fm, p0, p1, p2, 0, 3 - b, o0, o1, o2, nb);
	return ktab_pack(o0, o1, o2, pb + (int)nb);
}
// a step of a sweep that is served by the table (tab) or by an extension: the loads ...
B200_HD void fm_step_load(const FmView &fm, bool tab, int L, uint32_t idx, uint64_t x0, uint64_t x1, uint64_t x2, int is_back, OccRaw &rk, OccRaw &rl, int &half)
{
	if (tab) { rk = ld_sector(ktab_sector(fm, L, idx, half)); return; }
	half = 0;
	fm_extend_load(fm, x0, x1, x2, is_back, rk, rl);
}
// ... and the result.  n_blocks counts what the reference's extension of (x0, x1, x2) touches either way.
B200_HD void fm_step_use(const FmView &fm, bool tab, int half, uint64_t x0, uint64_t x1, uint64_t x2, int is_back, int c, const OccRaw &rk, const OccRaw &rl,
                         uint64_t &o0, uint64_t &o1, uint64_t &o2, int64_t &n_blocks)
{
	if (tab) {
		int tb;
		ktab_unpack(rk, half, o0, o1, o2, tb);
		n_blocks += fm_extend_blocks(fm, is_back ? x0 : x1, x2);
		return;
	}
	fm_extend_use(fm, x0, x1, x2, is_back, c, rk, rl, o0, o1, o2, n_blocks);
}

/* ---------------------------------------------------------------- packed reads
 * The sweeps ask for WINDOWS of a read (the bases of a table look-up, of a Bloom filter word, of a comparison with the text).
 * Gathering them byte by byte is a loop only the asking lane runs while its warp waits; from a 2-bit copy of the read a window is
 * two loads and a shift.  One chunk-wide pass (k_pack_reads) packs every read: 32 bases per 64-bit word, first base most
 * significant, an ambiguous base stored as 0 and flagged in the mask word that follows (bit 31 - (t & 31) of its low half);
 * `stride` word pairs per read, the pairs past the read's end are zero. */
struct PackedRead {
	const uint64_t *w; int n_words;          // word pair i: w[2 * i] bases, w[2 * i + 1] mask
	B200_HD uint64_t bases_at(int i) const { return i >= 0 && i < n_words ? w[2 * i] : 0; }
	B200_HD uint64_t mask_at(int i) const { return i >= 0 && i < n_words ? w[2 * i + 1] : 0; }
	// the L <= 32 bases from position s on (positions outside the read give 0), right-aligned
	B200_HD uint64_t window(int s, int L) const
	{
		const int i = s >> 5, o = s & 31;
		const uint64_t hi = bases_at(i), lo = bases_at(i + 1);
		const uint64_t v = o ? hi << (2 * o) | lo >> (64 - 2 * o) : hi;
		return v >> (64 - 2 * L);
	}
	// ambiguity flags of the 32 positions from s on, position s in bit 31
	B200_HD uint32_t flags(int s) const
	{
		const int i = s >> 5, o = s & 31;
		const uint32_t hi = (uint32_t)mask_at(i), lo = (uint32_t)mask_at(i + 1);
		return o ? hi << o | lo >> (32 - o) : hi;
	}
	// number of plain bases from s on, at most L <= 32 (the caller bounds L by the read's end)
	B200_HD int plain_run(int s, int L) const
	{
		const uint32_t f = flags(s) & (L >= 32 ? 0xffffffffu : ~(0xffffffffu >> L));
#if defined(__CUDA_ARCH__)
		const int n = f ? __clz((int)f) : 32;
#else
		const int n = f ? __builtin_clz(f) : 32;
#endif
		return n < L ? n : L;
	}
	// position of the first ambiguous base at or after s, or `end`
	B200_HD int next_flag_from(int s, int end) const
	{
		for (int p = s; p < end; p += 32) {
			const int n = plain_run(p, end - p < 32 ? end - p : 32);
			if (n < 32 || p + n >= end) return p + n < end ? p + n : end;
		}
		return end;
	}
	// position of the last ambiguous base before x, or -1
	B200_HD int last_flag_before(int x) const
	{
		for (int i = (x - 1) >> 5; i >= 0; --i) {
			uint32_t f = (uint32_t)mask_at(i);
			if (i == (x - 1) >> 5 && ((x - 1) & 31) != 31) f &= ~(0xffffffffu >> (((x - 1) & 31) + 1));
			if (f) {
#if defined(__CUDA_ARCH__)
				return (i << 5) + 31 - (__ffs((int)f) - 1);
#else
				return (i << 5) + 31 - __builtin_ctz(f);
#endif
			}
		}
		return -1;
	}
};
// word pair `i` of a read's packed copy from its byte codes (0-3 plain, > 3 ambiguous)
B200_HD void pack_read_word(const uint8_t *q, int len, int i, uint64_t &bases, uint64_t &mask)
{
	bases = 0; mask = 0;
	for (int t = 0; t < 32; ++t) {
		const int p = (i << 5) + t;
		const int c = p < len ? q[p] : 0;
		bases = bases << 2 | (uint64_t)(c > 3 ? 0 : c);
		mask = mask << 1 | (uint64_t)(p < len && c > 3 ? 1 : 0);
	}
}
B200_HD int packed_words_for(int max_len) { return (max_len + 31) / 32 + 1; }

// 32 bases of the 2-bit text from forward position f on (f >= 0; bytes past the array's end must be readable: it is padded), left-aligned
B200_HD uint64_t pac_window32(const uint8_t *pac, int64_t f)
{
	const uint8_t *p = pac + (f >> 2);
	uint64_t v = 0;
	for (int k = 0; k < 8; ++k) v = v << 8 | p[k];
	const int o = (int)(f & 3);
	return o ? v << (2 * o) | (uint64_t)p[8] >> (8 - 2 * o) : v;
}
// the bases of a 64-bit word (32 of them) in reverse order
B200_HD uint64_t rev_bases32(uint64_t v)
{
#if defined(__CUDA_ARCH__)
	v = __brevll(v);
#else
	v = (v >> 32) | (v << 32);
	v = (v & 0xffff0000ffff0000ull) >> 16 | (v & 0x0000ffff0000ffffull) << 16;
	v = (v & 0xff00ff00ff00ff00ull) >> 8 | (v & 0x00ff00ff00ff00ffull) << 8;
	v = (v & 0xf0f0f0f0f0f0f0f0ull) >> 4 | (v & 0x0f0f0f0f0f0f0f0full) << 4;
	v = (v & 0xccccccccccccccccull) >> 2 | (v & 0x3333333333333333ull) << 2;
	v = (v & 0xaaaaaaaaaaaaaaaaull) >> 1 | (v & 0x5555555555555555ull) << 1;
#endif
	return (v & 0xaaaaaaaaaaaaaaaaull) >> 1 | (v & 0x5555555555555555ull) << 1;     // (bits reversed: put each base's two bits back in order)
}
// How far does the text go on like the read?  The read's bases from position i on (pr, at most n of them, all plain) against
// 3 - T[p], 3 - T[p - 1], ... - the reverse-complement strand of the text read downwards from p (see the unique walk in
// smem_sweeps.cuh) - 32 bases per round; returns the number of bases that agree (at most n, at most p + 1).
B200_HD int text_match_down(const uint8_t *pac, int64_t l_pac, int64_t p, const PackedRead &pr, int i, int n)
{
	int m = 0;
	while (m < n && p >= 0) {
		int avail;               // text bases this round can compare, at most 32
		uint64_t t;
		if (p >= l_pac) {        // 3 - T[p - u] = pac[f + u], f = 2 l_pac - 1 - p: the forward strand read upwards
			const int64_t f = (l_pac << 1) - 1 - p;
			avail = l_pac - f < 32 ? (int)(l_pac - f) : 32;
			t = pac_window32(pac, f);
		} else {                 // 3 - pac[p - u]: the forward strand read downwards, complemented
			avail = p + 1 < 32 ? (int)(p + 1) : 32;
			const int64_t s = p - 31;
			t = s >= 0 ? ~rev_bases32(pac_window32(pac, s)) : ~rev_bases32(pac_window32(pac, 0) >> (2 * (int)-s));
		}
		int want = n - m < 32 ? n - m : 32;
		if (avail < want) want = avail;
		const uint64_t r = pr.window(i + m, 32) << 0;
		const uint64_t x = (r ^ (t >> 0)) & (want >= 32 ? ~0ull : ~(~0ull >> (2 * want)));
		// r is right-aligned for L = 32 (a whole word), t left-aligned: both hold their first base in the top two bits
		int eq;
#if defined(__CUDA_ARCH__)
		eq = x ? __clzll((long long)x) >> 1 : want;
#else
		eq = x ? __builtin_clzll(x) >> 1 : want;
#endif
		if (eq > want) eq = want;
		m += eq; p -= eq;
		if (eq < want) break;                  // (a round cut short by the strand boundary goes on across it)
	}
	return m;
}

// The interval list of one lane.  Entry k < quota lives in shared memory (sh[(k*4 + word) * stride]), the rest in the
// lane's global strip (spill[(k - quota) * sstride]).  Values up to 2^33-1, end positions up to 2^29-1.
struct SeedList {
	uint32_t *sh; int stride, quota;
	Q4 *spill; int64_t sstride;
	B200_HD void set(int k, uint64_t x0, uint64_t x1, uint64_t x2, int end) const
	{
		Q4 v;
		v.x = (uint32_t)x0; v.y = (uint32_t)x1; v.z = (uint32_t)x2;
		v.w = (uint32_t)end | (uint32_t)(x0 >> 32) << 29 | (uint32_t)(x1 >> 32) << 30 | (uint32_t)(x2 >> 32) << 31;
		if (k < quota) {
			uint32_t *p = sh + (size_t)(k * 4) * stride;
			p[0] = v.x; p[stride] = v.y; p[2 * stride] = v.z; p[3 * stride] = v.w;
		} else spill[(int64_t)(k - quota) * sstride] = v;
	}
	B200_HD void get(int k, uint64_t &x0, uint64_t &x1, uint64_t &x2, int &end) const
	{
		Q4 v;
		if (k < quota) {
			const uint32_t *p = sh + (size_t)(k * 4) * stride;
			v.x = p[0]; v.y = p[stride]; v.z = p[2 * stride]; v.w = p[3 * stride];
		} else v = spill[(int64_t)(k - quota) * sstride];
		x0 = (uint64_t)(v.w >> 29 & 1u) << 32 | v.x;
		x1 = (uint64_t)(v.w >> 30 & 1u) << 32 | v.y;
		x2 = (uint64_t)(v.w >> 31) << 32 | v.z;
		end = (int)(v.w & 0x1fffffffu);
	}
};

struct SeedLane {
	enum { P1_NEXT, FWD, BWD_ROW, BWD, SMEM_END, P2_NEXT, P3_NEXT, P3, FINISH };
	// the read
	int len; const uint8_t *q; Intv *outp;
	// machine
	int st, pass, x, sx, i, c, is_back;
	uint64_t k0, k1, k2; int kend;        // interval to extend next (forward: ik; backward: the entry being extended)
	uint64_t min_intv, last_x2;
	int n_list, n_prev, j, n_curr;
	int nm, last_start, ret;
	int n_out, old_n, k2i;

	B200_HD void begin(const SeedOpt &so, int len_, const uint8_t *q_, Intv *outp_)
	{
		len = len_; q = q_; outp = outp_;
		n_out = 0; x = 0; pass = 1;
		st = len >= so.min_seed_len ? P1_NEXT : FINISH;
	}
	B200_HD void set_intv(const FmView &fm, int b)
	{
		k0 = l2_at(fm, b) + 1; k2 = l2_at(fm, b + 1) - l2_at(fm, b); k1 = l2_at(fm, 3 - b) + 1;
	}
	B200_HD void start_smem(const FmView &fm, int x_, uint64_t mi)
	{
		set_intv(fm, q[x_]);
		kend = x_ + 1; min_intv = mi < 1 ? 1 : mi;
		n_list = 0; nm = 0; last_start = 0; sx = x_; i = x_ + 1; st = FWD;
	}
	B200_HD void emit(const SeedOpt &so, int cap, uint64_t p0, uint64_t p1, uint64_t p2, int start, int end, bool filter)
	{
		if (filter && end - start < so.min_seed_len) return;
		if (n_out < cap) { Intv v; v.x0 = p0; v.x1 = p1; v.x2 = p2; v.info = (uint64_t)start << 32 | (uint32_t)end; outp[n_out] = v; }
		++n_out;
	}
	B200_HD void fwd_done(const SeedList &L)
	{
		L.set(n_list++, k0, k1, k2, kend);        // the interval that could not be extended any further
		ret = kend; n_prev = n_list; i = sx - 1; st = BWD_ROW;
	}

	// run the cheap transitions until the lane needs an extension (true; inputs in k0,k1,k2,is_back,c) or its read is done
	B200_HD bool advance(const FmView &fm, const SeedOpt &so, int cap, const SeedList &L)
	{
		for (;;) {
			switch (st) {
			case P1_NEXT:
				if (x >= len) { pass = 2; old_n = n_out <= cap ? n_out : 0; k2i = 0; st = P2_NEXT; break; }
				if (q[x] > 3) { ++x; break; }
				start_smem(fm, x, 1);
				break;
			case FWD:
				if (i < len && q[i] < 4) { c = 3 - q[i]; is_back = 0; return true; }
				fwd_done(L);                          // end of the read or an ambiguous base
				break;
			case BWD_ROW: {
				const int cc = i < 0 ? -1 : (q[i] < 4 ? (int)q[i] : -1);
				if (cc < 0) {                         // nothing can be extended: only the longest entry may be reported
					if (nm == 0 || i + 1 < last_start) {
						L.get(n_list - 1, k0, k1, k2, kend);
						emit(so, cap, k0, k1, k2, i + 1, kend, true);
						last_start = i + 1; ++nm;
					}
					st = SMEM_END;
					break;
				}
				c = cc; j = 0; n_curr = 0; st = BWD;
				break;
			}
			case BWD:
				L.get(n_list - 1 - j, k0, k1, k2, kend);
				is_back = 1;
				return true;
			case SMEM_END:
				if (pass == 1) { x = ret; st = P1_NEXT; }
				else { ++k2i; st = P2_NEXT; }
				break;
			case P2_NEXT: {
				if (k2i >= old_n) { pass = 3; x = 0; st = so.max_mem_intv > 0 ? P3_NEXT : FINISH; break; }
				const Intv p = outp[k2i];
				const int start = (int)(p.info >> 32), end = (int)(int32_t)p.info;
				if (end - start < so.split_len || p.x2 > (uint64_t)so.split_width) { ++k2i; break; }
				const int mid = (start + end) >> 1;
				if (q[mid] > 3) { ++k2i; break; }
				start_smem(fm, mid, p.x2 + 1);
				break;
			}
			case P3_NEXT:
				if (x >= len) { st = FINISH; break; }
				if (q[x] > 3) { ++x; break; }
				set_intv(fm, q[x]);
				sx = x; i = x + 1; st = P3;
				break;
			case P3:
				if (i >= len) { x = len; st = P3_NEXT; break; }
				if (q[i] > 3) { x = i + 1; st = P3_NEXT; break; }
				c = 3 - q[i]; is_back = 0;
				return true;
			default:
				return false;
			}
		}
	}

	// Fast path for the three steady states (forward sweep, backward sweep incl. the change of row, greedy pass): digests
	// the result of the extension and sets up the next one without leaving straight-line code.  Returns 0 when the next
	// extension is ready, 1 when the rarer transition must go through consume() + advance() (state untouched), 2 when
	// only advance() is needed (state already updated as consume() would have).
	B200_HD int fast_step(const SeedOpt &so, const SeedList &L, uint64_t o0, uint64_t o1, uint64_t o2)
	{
		if (st == BWD) {
			const bool live = o2 >= min_intv;
			if (!live && n_curr == 0 && (nm == 0 || i + 1 < last_start)) return 1;      // an SMEM is reported
			if (live && (n_curr == 0 || o2 != last_x2)) {
				L.set(n_list - 1 - n_curr, o0, o1, o2, kend);
				++n_curr; last_x2 = o2;
			}
			if (++j == n_prev) {
				if (n_curr == 0) { st = SMEM_END; return 2; }
				n_prev = n_curr; --i;
				if (i < 0 || q[i] > 3) { st = BWD_ROW; return 2; }
				c = q[i]; j = 0; n_curr = 0;
			}
			L.get(n_list - 1 - j, k0, k1, k2, kend);
			return 0;
		}
		const int ni = i + 1;
		if (ni >= len) return 1;
		const int qn = q[ni];
		if (qn > 3) return 1;
		if (st == FWD) {
			if (o2 != k2) {
				if (o2 < min_intv) return 1;
				L.set(n_list++, k0, k1, k2, kend);
			}
			kend = ni;
		} else if (o2 < (uint64_t)so.max_mem_intv && i - sx >= so.min_seed_len) return 1;   // P3: a seed is reported
		k0 = o0; k1 = o1; k2 = o2; i = ni; c = 3 - qn;
		return 0;
	}

	// digest the result of the extension requested by advance()
	B200_HD void consume(const SeedOpt &so, int cap, const SeedList &L, uint64_t o0, uint64_t o1, uint64_t o2)
	{
		if (st == FWD) {
			if (o2 != k2) {
				if (o2 < min_intv) { fwd_done(L); return; }
				L.set(n_list++, k0, k1, k2, kend);
			}
			k0 = o0; k1 = o1; k2 = o2; kend = i + 1; ++i;
		} else if (st == BWD) {
			if (o2 < min_intv) {
				if (n_curr == 0 && (nm == 0 || i + 1 < last_start)) {
					emit(so, cap, k0, k1, k2, i + 1, kend, true);
					last_start = i + 1; ++nm;
				}
			} else if (n_curr == 0 || o2 != last_x2) {
				L.set(n_list - 1 - n_curr, o0, o1, o2, kend);
				++n_curr; last_x2 = o2;
			}
			if (++j == n_prev) {
				if (n_curr == 0) st = SMEM_END;
				else { n_prev = n_curr; --i; st = BWD_ROW; }
			}
		} else {                                      // P3 (reference src/bwt.c:367-376)
			if (o2 < (uint64_t)so.max_mem_intv && i - sx >= so.min_seed_len) {
				if (o2 > 0) emit(so, cap, o0, o1, o2, sx, i + 1, false);
				x = i + 1; st = P3_NEXT;
			} else { k0 = o0; k1 = o1; k2 = o2; ++i; }
		}
	}
};

#if defined(__CUDACC__)
// Persistent kernel: every lane pulls reads from *next_read until none are left.
// n_intv[r] = number of intervals of read r (unsorted, in out[r*cap ..]), or -(needed) when cap was too small.
__global__ void __launch_bounds__(128) k_seed_lanes(FmView fm, SeedOpt so, int n_reads, const int64_t *__restrict__ off,
                                                    const uint8_t *__restrict__ codes, Intv *out, int cap, int quota, Q4 *spill,
                                                    int32_t *n_intv, int *next_read, int *worst, unsigned long long *occ_blocks,
                                                    const int32_t *__restrict__ only_neg)     // non-null: only reads r with only_neg[r] < 0
{
	extern __shared__ uint32_t seed_sh[];
	SeedList L;
	L.sh = seed_sh + threadIdx.x; L.stride = 128; L.quota = quota;
	L.sstride = (int64_t)gridDim.x * blockDim.x;
	L.spill = spill + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	SeedLane ln;
	ln.st = SeedLane::FINISH; ln.n_out = 0;
	int r = -1;
	bool need = false, drained = false;
	int64_t blocks = 0;
	for (;;) {
		while (!need && !drained) {
			if (r >= 0) {
				n_intv[r] = ln.n_out > cap ? -ln.n_out : ln.n_out;
				if (ln.n_out > cap) atomicMax(worst, ln.n_out);
			}
			r = atomicAdd(next_read, 1);
			if (r >= n_reads) { r = -1; drained = true; break; }
			if (only_neg && only_neg[r] >= 0) { r = -1; continue; }
			ln.begin(so, (int)(off[r + 1] - off[r]), codes + off[r], out + (int64_t)r * cap);
			need = ln.advance(fm, so, cap, L);
		}
		if (!__any_sync(0xffffffffu, need)) break;
		if (need) {
			uint64_t o0, o1, o2;
			fm_extend_sel(fm, ln.k0, ln.k1, ln.k2, ln.is_back, ln.c, o0, o1, o2, blocks);
			const int slow = ln.fast_step(so, L, o0, o1, o2);
			if (slow) {
				if (slow == 1) ln.consume(so, cap, L, o0, o1, o2);
				need = ln.advance(fm, so, cap, L);
			}
		}
	}
	for (int o = 16; o > 0; o >>= 1) blocks += __shfl_down_sync(0xffffffffu, blocks, o);
	if ((threadIdx.x & 31) == 0 && blocks) atomicAdd(occ_blocks, (unsigned long long)blocks);
}
#endif

} // namespace b200
