// smem_sweeps.cuh - SMEM seeding (mem_collect_intv, reference src/bwamem.c:114-162) as four sweeps: the throughput path of the
// seeding stage.
//
// 1. The split.  In the per-read state machine of smem_kernel.cuh the lanes of a warp sit in different states (forward sweep,
//    backward sweep, greedy pass, transitions).  The dependencies of mem_collect_intv allow a coarser split:
//      * the start of the next bwt_smem1a call is the END of the longest forward match (ret = curr[0].info after the reversal,
//        src/bwt.c:319-322), i.e. it is known after the FORWARD sweep alone - so all forward sweeps of pass 1 of a read run back
//        to back, each leaving its interval list (src/bwt.c:304-318) in a per-read strip in HBM;
//      * the greedy third pass (bwt_seed_strategy1) depends on nothing but the read;
//      * every backward sweep (src/bwt.c:326-345) depends only on its own list;
//      * pass 2 re-seeds inside the long, rare SMEMs that pass 1 reported: same two sweeps again.
//    So: k_sweep_fwd<1> (pass-1 forward sweeps + pass 3), k_sweep_bwd, k_sweep_fwd<2>, k_sweep_bwd.  Lanes are persistent and pull
//    reads from a counter; every trip of the warp loop is one memory access per lane.
// 2. Touching less.  With extensions only, these kernels sit on the DRAM random-access roofline of a human-sized index.  What
//    they do now instead of most extensions (every structure is a pure function of the index, built at upload):
//      * k-mer interval tables: a step whose result is a pattern of at most kmax bases is one look-up (FwdLane tab 1-3, BwdLane tab 1);
//      * backward sweeps as independent entry chains with the report rule taken from the chains' ends (BwdLane);
//      * Bloom filters over the text's min_seed_len-mers in front of every chain (BwdLane tab 2);
//      * unique walks: a one-row interval is compared with the 2-bit text through the whole suffix array and its inverse (FwdLane tab 4, 5);
//      * packed 2-bit copies of the reads, so that the windows those look-ups need are word operations (PackedRead).
//    The interval lists and SMEMs that come out are the reference's, bit for bit.
// A read whose strip would overflow, or whose backward chains exceed their work budget, is flagged and redone by the general
// state-machine kernel (k_seed_lanes: the reference's loops, no tables), which has no such limits.
// Host/device code: tests/hostemu runs the same four sweeps on the CPU against the plain restatement on every read.
#pragma once
#include "smem_kernel.cuh"

namespace b200 {

// strip of one read: records of 16 bytes; a sweep is a header {n, x, min_intv lo, min_intv hi} followed by its n entries
// (packed like SeedList entries) in push order
struct SweepStrip {
	Q4 *base; int cap;
	B200_HD static Q4 pack(uint64_t x0, uint64_t x1, uint64_t x2, int end)
	{
		Q4 v;
		v.x = (uint32_t)x0; v.y = (uint32_t)x1; v.z = (uint32_t)x2;
		v.w = (uint32_t)end | (uint32_t)(x0 >> 32) << 29 | (uint32_t)(x1 >> 32) << 30 | (uint32_t)(x2 >> 32) << 31;
		return v;
	}
};

struct FwdLane {
	enum { NEXT, FWD, FWD_JUMP, P3_NEXT, P3, DONE };
	int len; const uint8_t *q; PackedRead pr; Intv *outp; Q4 *strip; int strip_cap;
	int mode;                   // 1: pass-1 sweeps then pass 3; 2: pass-2 sweeps
	int st, x, sx, i, c;
	uint64_t k0, k1, k2, min_intv; int kend;
	int wpos, hdr_pos, n, n_sweeps, over;
	int n_out, old_n, k2i;
	// k-mer tables (smem_kernel.cuh): tab != 0 = the pending step is a table look-up of the klen bases packed in W instead of an
	// extension; 1 = one step of a forward sweep (pattern q[sx .. i]), 2 = the first klen-1 steps of the greedy pass at once
	//   3 = the first kj bases of a forward sweep at once (its shorter prefixes become IMPLICIT entries of the list: the backward
	//       chains take them from the tables, see BwdLane); 4 / 5 = the suffix array / inverse suffix array read of a unique walk
	int kmax, kj, tab, klen; uint32_t W;
	int n_impl;                 // implicit entries of the current sweep: forward extents 1 .. n_impl
	int uw_ok, uw_n; uint64_t t1;

	// k_first: first interval of `outp` that pass 2 may re-seed (the greedy seeds of pass 3 come before it and are not candidates)
	B200_HD void begin(const SeedOpt &so, const FmView &fm, int mode_, int len_, const uint8_t *q_, const PackedRead &pr_, Intv *outp_, Q4 *strip_, int strip_cap_, int n_out_, int k_first)
	{
		pr = pr_;
		kmax = fm.kmax; kj = fm.kmax < so.min_seed_len ? fm.kmax : so.min_seed_len; mode = mode_;
		uw_ok = fm.sa5 != nullptr && fm.isa5 != nullptr; uw_n = 0; t1 = 0; n_impl = 0; len = len_; q = q_; outp = outp_; strip = strip_; strip_cap = strip_cap_;
		n_out = n_out_; old_n = n_out_; k2i = k_first;
		x = 0; wpos = 0; n_sweeps = 0; over = 0;
		tab = 0; klen = 0; W = 0;
		st = len >= so.min_seed_len ? NEXT : DONE;
	}
	// the extension by base qn = q[i] of the sweep that began at sx is the next step: through the table while the pattern is short
	B200_HD void want(int qn)
	{
		c = 3 - qn;
		tab = 0;
		if (st == FWD && i - sx < kmax) { W = W << 2 | (uint32_t)qn; klen = i - sx + 1; tab = 1; }
	}
	B200_HD void set_intv(const FmView &fm, int b) { k0 = l2_at(fm, b) + 1; k2 = l2_at(fm, b + 1) - l2_at(fm, b); k1 = l2_at(fm, 3 - b) + 1; }
	B200_HD void push()
	{
		if (wpos < strip_cap) strip[wpos] = SweepStrip::pack(k0, k1, k2, kend); else over = 1;
		++wpos; ++n;
	}
	B200_HD void start_sweep(const FmView &fm, int x_, uint64_t mi)
	{
		set_intv(fm, q[x_]);
		kend = x_ + 1; min_intv = mi < 1 ? 1 : mi; sx = x_; i = x_ + 1; n = 0;
		hdr_pos = wpos++;
		W = q[x_];
		st = FWD; n_impl = 0;
		// the first kj bases at once when they are plain bases of the read: one look-up instead of kj - 1 steps
		if (kj >= 2 && x_ + kj <= len && pr.plain_run(x_, kj) == kj) { W = (uint32_t)pr.window(x_, kj); klen = kj; tab = 3; st = FWD_JUMP; }
	}
	B200_HD void end_sweep()      // the interval that could not be extended further closes the list; its end is the next x
	{
		push();
		if (hdr_pos < strip_cap) { Q4 h; h.x = (uint32_t)n | (uint32_t)n_impl << 16; h.y = (uint32_t)sx; h.z = (uint32_t)min_intv; h.w = (uint32_t)(min_intv >> 32); strip[hdr_pos] = h; }
		else over = 1;
		++n_sweeps;
		x = kend;
		st = NEXT;
	}
	B200_HD void emit3(int cap, uint64_t p0, uint64_t p1, uint64_t p2, int start, int end)
	{
		if (n_out < cap) { Intv v; v.x0 = p0; v.x1 = p1; v.x2 = p2; v.info = (uint64_t)start << 32 | (uint32_t)end; outp[n_out] = v; }
		++n_out;
	}
	// slow path: runs until the lane needs an extension (true) or has finished its read (false)
	B200_HD bool advance(const FmView &fm, const SeedOpt &so)
	{
		for (;;) {
			switch (st) {
			case NEXT:
				if (mode == 1) {
					if (x >= len) { x = 0; st = so.max_mem_intv > 0 ? P3_NEXT : DONE; break; }
					if (q[x] > 3) { ++x; break; }
					start_sweep(fm, x, 1);
				} else {
					if (k2i >= old_n) { st = DONE; break; }
					const Intv p = outp[k2i++];
					const int start = (int)(p.info >> 32), end = (int)(int32_t)p.info;
					if (end - start < so.split_len || p.x2 > (uint64_t)so.split_width) break;
					const int mid = (start + end) >> 1;
					if (q[mid] > 3) break;
					start_sweep(fm, mid, p.x2 + 1);
				}
				break;
			case FWD:
				if (i < len && q[i] < 4) { want(q[i]); return true; }
				end_sweep();
				break;
			case FWD_JUMP:
				return true;                                  // (tab == 3: set up by start_sweep)
			case P3_NEXT: {
				if (x >= len) { st = DONE; break; }
				if (q[x] > 3) { ++x; break; }
				set_intv(fm, q[x]);
				sx = x; i = x + 1; st = P3;
				// no seed can be reported before the pattern has min_seed_len + 1 bases (src/bwt.c:369): the run of plain bases at sx,
				// up to min(kmax, min_seed_len) of them, is ONE table look-up instead of an extension per base
				const int run = kj < len - sx ? kj : len - sx;
				const int m = run >= 2 ? pr.plain_run(sx, run) : 0;
				if (m >= 2) { W = (uint32_t)pr.window(sx, m); klen = m; i = sx + m - 1; tab = 2; c = 3 - q[i]; return true; }
				break;
			}
			case P3:
				if (i >= len) { st = DONE; break; }
				if (q[i] > 3) { x = i + 1; st = P3_NEXT; break; }
				want(q[i]);
				return true;
			default:
				return false;
			}
		}
	}
	// digest one extension; true = the next extension is already set up, false = go through advance()
	B200_HD bool step(const FmView &fm, const SeedOpt &so, int cap, uint64_t o0, uint64_t o1, uint64_t o2)
	{
		if (st == FWD_JUMP) {
			st = FWD; tab = 0;
			if (o2 >= min_intv) { k0 = o0; k1 = o1; k2 = o2; i = sx + kj; kend = i; n_impl = kj - 1; }
			else W = q[sx];                                  // the match ends inside the first kj bases: step through them (i, k* are still the first base's)
		} else if (tab == 4) {
			// unique walk: the pattern's single occurrence is at a known place of the text, so the rest of the forward sweep is a
			// comparison of the read with the reference (its reverse-complement strand read downwards, see fwd_lane_fetch)
			t1 = o0;
			const int m = t1 >= 1 ? text_match_down(fm.pac, fm.l_pac, (int64_t)t1 - 1, pr, i, pr.next_flag_from(i, len) - i) : 0;
			if (m == 0) { tab = 0; end_sweep(); return false; }   // (q[i] is a plain base: the text differs there, or ends)
			uw_n = m; t1 -= m; tab = 5;
			return true;
		} else if (tab == 5) {
			k1 = o0; i += uw_n; kend = i; tab = 0;
			end_sweep();                                      // whatever stopped the comparison stops the sweep (src/bwt.c:304-318)
			return false;
		} else if (st == FWD) {
			if (o2 != k2) {
				if (o2 < min_intv) { end_sweep(); return false; }
				push();
			}
			k0 = o0; k1 = o1; k2 = o2; kend = i + 1; ++i;
			if (k2 == 1 && min_intv == 1 && uw_ok && i < len && q[i] < 4) { tab = 4; return true; }
		} else {                                              // P3 (reference src/bwt.c:367-376)
			if (o2 < (uint64_t)so.max_mem_intv && i - sx >= so.min_seed_len) {
				if (o2 > 0) emit3(cap, o0, o1, o2, sx, i + 1);
				x = i + 1; st = P3_NEXT;
				return false;
			}
			k0 = o0; k1 = o1; k2 = o2; ++i;
		}
		if (i < len) { const int qn = q[i]; if (qn < 4) { want(qn); return true; } }
		return false;
	}
};

// one trip of a forward lane: the pending step's result and the reference's occ-block count for it
B200_HD void fwd_lane_fetch(const FmView &fm, const FwdLane &ln, uint64_t &o0, uint64_t &o1, uint64_t &o2, int64_t &blocks)
{
	if (ln.tab >= 4) {
		// Unique walk (tab 4, 5).  Once the interval holds ONE row the rest of a forward sweep asks, base after base, whether the
		// text goes on like the read; the reference answers each with an occ look-up of the reverse complement's row k1.  With
		// the whole suffix array and its inverse in HBM the row is turned into its text position once (tab 4), the bases are
		// compared in the 2-bit text itself (step()), and the position where the comparison stops is turned back into the row
		// (tab 5) - the forward coordinate k0 does not move while the single occurrence goes on matching.
		o0 = sa5_read(ln.tab == 4 ? fm.sa5 : fm.isa5, ln.tab == 4 ? ln.k1 : ln.t1); o1 = 0; o2 = 0;
		blocks += 1;
		return;
	}
	OccRaw rk, rl;
	int half;
	fm_step_load(fm, ln.tab != 0, ln.klen, ln.W, ln.k0, ln.k1, ln.k2, 0, rk, rl, half);
	if (ln.tab >= 2) { int tb; ktab_unpack(rk, half, o0, o1, o2, tb); blocks += ln.tab == 2 ? tb : 1; }
	else fm_step_use(fm, ln.tab != 0, half, ln.k0, ln.k1, ln.k2, 0, ln.c, rk, rl, o0, o1, o2, blocks);
}

// Backward sweeps (reference src/bwt.c:326-345) WITHOUT the rows.  The reference extends every entry of the forward list one base
// at a time, row by row, and reports an entry when it dies in a row in which no longer-forward entry survived before it.  An
// entry's occurrences contain those of every entry that reaches further forward, so with
//     b_k = the number of bases entry k can be extended backward by before its interval falls under min_intv
//           (or the read's start / an ambiguous base stops it: b_lim),
// b_k never decreases from the longest-forward entry to the shortest, and entry k is reported exactly when b_k is GREATER than the
// b of the next-longer entry (equal b: the longer one was reported in that row first and `i + 1 < last_start` fails; entries the
// reference merges because their sizes coincide have identical occurrences, hence equal b, and are never reported).  So the
// entries are independent chains and nothing needs b_k but the report decision - and a report shorter than min_seed_len is dropped
// (src/bwamem.c:130).  With the k-mer tables an entry of L0 < kj = min(kmax, min_seed_len) forward bases starts its chain with ONE
// look-up of the kj-base pattern that ends where the entry ends: if that pattern is absent the entry dies short of kj bases - it
// cannot be reported, and every shorter-forward entry that is still alive at its own kj-base pattern reaches further back than it
// did (the report rule holds without knowing where it died); if present, the chain goes on from there by extensions.  Per sweep:
// one look-up per entry plus the few extensions past kj bases, instead of the whole triangle of rows.
struct BwdLane {
	enum { NEXT, ENTRY, STEP, DONE };
	int len; const uint8_t *q; PackedRead pr; Intv *outp; const Q4 *strip;
	int st, c, rpos, sweeps_left;
	uint64_t k0, k1, k2, min_intv; int kend;      // the entry's interval as extended so far
	int n_list, n_impl, j, x, b, b_lim, b_prev, n_out;   // n_impl: implicit entries after the n_list stored ones (forward extents n_impl .. 1)
	int kmax, kj, tab, klen; uint32_t W;          // tab == 1: the pending step is a look-up of the klen bases in W; 2: a Bloom filter word
	uint64_t PH, PL;                              // bases q[x-32 .. x) and q[x .. x+32), two bits each, first base most significant
	int bk; uint64_t bv;                          // Bloom filter over the text's bk-mers (0: none); bv = hash of the window being asked for
	// The reference merges entries whose sizes coincide in a row; independent chains walk such entries separately.  That is a few
	// steps per entry on ordinary reads, but entries that survive TOGETHER for many bases (a read whose flank matches another copy
	// of a repeat) would multiply the reference's work: a read that needs more than `budget` extensions in one kernel is handed to
	// the general state machine (k_seed_lanes: the reference's rows) like a read whose strip overflowed.
	int steps, budget, over;
	// What stands in for the merge: the sizes the last WALKED entry had at the first TRAJ offsets of its chain (tr[cur]).  A
	// shorter-forward entry whose size at one of those offsets is the same has the same occurrences from there on (its set contains
	// the walked entry's), so it ends where the walked entry ended and is not reported: its chain stops at once.
	enum { TRAJ = 8 };
	uint32_t *tr; int tr_stride;                  // tr[(buf * TRAJ + t) * tr_stride]: buf `cur` = last walked entry, the other = the chain being walked
	int cur, tr_base, tr_n, tr_b, nw_base, nw_n;  // offsets tr_base .. tr_base + tr_n - 1 hold sizes; tr_b = where that entry ended

	B200_HD bool same_as_walked(int off, uint64_t size) const
	{
		const int t = off - tr_base;
		return t >= 0 && t < tr_n && size < 0xffffffffu && tr[(cur * TRAJ + t) * tr_stride] == (uint32_t)size;
	}
	B200_HD void note(int off, uint64_t size)      // the chain being walked is at `off` with `size` occurrences
	{
		if (nw_n == 0) nw_base = off;
		if (off - nw_base == nw_n && nw_n < TRAJ) { tr[((cur ^ 1) * TRAJ + nw_n) * tr_stride] = size < 0xffffffffu ? (uint32_t)size : 0xffffffffu; ++nw_n; }
	}

	B200_HD void begin(const SeedOpt &so, int kmax_, int bk_, int len_, const uint8_t *q_, const PackedRead &pr_, Intv *outp_, const Q4 *strip_, int n_sweeps, int n_out_, uint32_t *tr_, int tr_stride_)
	{
		pr = pr_;
		tr = tr_; tr_stride = tr_stride_; cur = 0; tr_n = 0; tr_base = 0; tr_b = 0; nw_n = 0; nw_base = 0;
		kmax = kmax_; kj = kmax_ < so.min_seed_len ? kmax_ : so.min_seed_len; tab = 0; klen = 0; W = 0; PH = PL = 0; bk = bk_; bv = 0;
		len = len_; q = q_; outp = outp_; strip = strip_; rpos = 0; sweeps_left = n_sweeps; n_out = n_out_;
		steps = 0; budget = 4 * len_ + 64; over = 0;
		st = NEXT;
	}
	B200_HD void emit(const SeedOpt &so, int cap, int start)
	{
		if (kend - start >= so.min_seed_len) {
			if (n_out < cap) { Intv v; v.x0 = k0; v.x1 = k1; v.x2 = k2; v.info = (uint64_t)start << 32 | (uint32_t)kend; outp[n_out] = v; }
			++n_out;
		}
	}
	// the entry's chain ended after b backward bases
	B200_HD void finish(const SeedOpt &so, int cap)
	{
		if (b > b_prev) { emit(so, cap, x - b); b_prev = b; }
		cur ^= 1; tr_base = nw_base; tr_n = nw_n; tr_b = b;       // its sizes are what the following entries are compared with
		++j; st = ENTRY;
	}
	// the entry has the occurrences of the last walked entry: it ends where that one ended, unreported
	B200_HD void merged() { b_prev = tr_b; ++j; st = ENTRY; }
	B200_HD bool advance(const SeedOpt &so, int cap)
	{
		for (;;) {
			switch (st) {
			case NEXT: {
				if (sweeps_left == 0) { st = DONE; return false; }
				--sweeps_left;
				const Q4 h = strip[rpos];
				++rpos;                                           // rpos -> first entry of the sweep
				n_list = (int)(h.x & 0xffffu); n_impl = (int)(h.x >> 16); x = (int)h.y; min_intv = (uint64_t)h.w << 32 | h.z;
				b_lim = x - 1 - pr.last_flag_before(x);
				PH = PL = 0;
				if (kj || bk) {
					const int span = bk > kj ? bk : kj;               // (no window asked for below reaches further from x than this)
					PH = pr.window(x - span, span);
					PL = pr.window(x, span) << (2 * (32 - span));
				}
				j = 0; b_prev = -1; tr_n = 0;
				st = ENTRY;
				break;
			}
			case ENTRY: {
				if (j == n_list + n_impl || (j >= n_list && b_lim == 0)) { rpos += n_list; st = NEXT; break; }
				if (j < n_list) {
					const Q4 v = strip[rpos + n_list - 1 - j];        // longest-forward entry first
					k0 = (uint64_t)(v.w >> 29 & 1u) << 32 | v.x;
					k1 = (uint64_t)(v.w >> 30 & 1u) << 32 | v.y;
					k2 = (uint64_t)(v.w >> 31) << 32 | v.z;
					kend = (int)(v.w & 0x1fffffffu);
				} else kend = x + n_impl - (j - n_list);              // an implicit entry: a prefix shorter than kj bases, taken from the tables below
				b = 0; nw_n = 0;
				if (b_lim == 0) { finish(so, cap); break; }
				const int l0 = kend - x;
				if (bk && l0 < bk) {
					// can the entry grow to bk (<= min_seed_len) bases at all?  Not if the read's start or an ambiguous base is nearer
					// than that, and not if the text does not hold the bk-base window that ends where the entry ends (min_intv >= 2,
					// the re-seeding pass: does not hold it more than once)
					if (bk - l0 > b_lim) { b_prev = -1; ++j; break; }
					bv = bloom_mix(window(kend, bk));
					tab = 2; st = STEP;
					return true;
				}
				if (!start_chain()) break;
				return true;
			}
			default:
				return false;
			}
		}
	}
	// the bases q[end - L .. end) as a number (first base most significant), end in (x, x + 32], L <= 32 bases of which at most 32 before x
	B200_HD uint64_t window(int end, int L) const
	{
		const int l0 = end - x, nh = L - l0;
		const uint64_t lo = PL >> (2 * (32 - l0));
		const uint64_t hi = nh <= 0 ? 0 : nh >= 32 ? PH : PH & (((uint64_t)1 << (2 * nh)) - 1);
		return (nh > 0 ? hi << (2 * l0) : 0) | (nh >= 0 ? lo : lo >> (2 * -nh));
	}
	// first step of the entry's chain: the kj-base look-up or, for an entry that is longer already, its first extension;
	// false = the entry was merged into the last walked one
	B200_HD bool start_chain()
	{
		const int l0 = kend - x;
		if (l0 < kj) {
			b = kj - l0 < b_lim ? kj - l0 : b_lim;
			klen = l0 + b;
			W = (uint32_t)window(kend, klen);
			tab = 1;
		} else {
			if (same_as_walked(0, k2)) { merged(); return false; }
			note(0, k2);
			tab = 0; c = q[x - 1];
		}
		st = STEP;
		return true;
	}
	// digest the look-up or the extension; true = the chain's next extension is set up, false = advance() needed
	B200_HD bool step(const SeedOpt &so, int cap, uint64_t o0, uint64_t o1, uint64_t o2)
	{
		if (tab == 2) {
			tab = 0;
			const uint64_t m = bloom_bits(bv);
			if ((o0 & m) != m) { b_prev = -1; ++j; st = ENTRY; return false; }     // certainly absent: cannot reach min_seed_len bases
			return start_chain();
		}
		if (tab) {
			tab = 0;
			if (o2 < min_intv) { b_prev = -1; ++j; st = ENTRY; return false; }     // dies short of kj bases: no report, see above
			k0 = o0; k1 = o1; k2 = o2;
		} else {
			if (o2 < min_intv) { finish(so, cap); return false; }
			k0 = o0; k1 = o1; k2 = o2; ++b;
			if (++steps > budget) { over = 1; st = DONE; return false; }
		}
		if (same_as_walked(b, k2)) { merged(); return false; }
		note(b, k2);
		if (b == b_lim) { finish(so, cap); return false; }
		c = q[x - b - 1];
		return true;
	}
};

// one trip of a backward lane
B200_HD void bwd_lane_fetch(const FmView &fm, const BwdLane &ln, uint64_t &o0, uint64_t &o1, uint64_t &o2, int64_t &blocks)
{
	if (ln.tab == 2) {
#if defined(__CUDA_ARCH__)
		o0 = __ldg(fm.bloom + 2 * (ln.bv & fm.bloom_mask) + (ln.min_intv >= 2 ? 1 : 0));
#else
		o0 = fm.bloom[2 * (ln.bv & fm.bloom_mask) + (ln.min_intv >= 2 ? 1 : 0)];
#endif
		o1 = o2 = 0;
		blocks += 1;
		return;
	}
	OccRaw rk, rl;
	int half;
	fm_step_load(fm, ln.tab != 0, ln.klen, ln.W, ln.k0, ln.k1, ln.k2, 1, rk, rl, half);
	if (ln.tab) { int tb; ktab_unpack(rk, half, o0, o1, o2, tb); blocks += 1; }
	else fm_extend_use(fm, ln.k0, ln.k1, ln.k2, 1, ln.c, rk, rl, o0, o1, o2, blocks);
}

#if defined(__CUDACC__)
struct SweepArgs {
	FmView fm; SeedOpt so;
	int n_reads; const int64_t *off; const uint8_t *codes;
	Intv *out; int cap;
	const uint64_t *pk; int pk_stride;  // packed copy of read r (smem_kernel.cuh PackedRead): pk_stride word pairs at pk + 2 * r * pk_stride
	Q4 *strips; int strip_cap;      // strip of read r at strips + r * strip_cap
	int32_t *n_intv;                // running count of reported intervals per read (may exceed cap: overflow)
	int32_t *n_first;               // number of pass-3 seeds at the head of a read's output (pass 2 skips them)
	int32_t *n_sweeps;              // sweeps left in the strip by the forward kernel; -1 = strip overflow, read needs the general kernel
	int *next_read;                 // one counter per kernel of the sequence (the kernel is told which)
	int *worst;                     // largest interval count over the reads whose output overflowed cap
	int *n_over;                    // number of reads handed to the general kernel
	unsigned long long *occ_blocks;
};

// MINB: resident blocks per SM the kernel is compiled for (the register budget follows: 9 -> 56 registers, 12 -> 40, 16 -> 32; what
// does not fit spills to thread-local memory - the cold part of the lane state - in exchange for more extensions in flight per SM)
// the packed copies of the reads (PackedRead), one word pair per thread
__global__ void k_pack_reads(int n_reads, const int64_t *__restrict__ off, const uint8_t *__restrict__ codes, int stride, uint64_t *pk)
{
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (int64_t)n_reads * stride) return;
	const int r = (int)(t / stride), i = (int)(t % stride);
	uint64_t b, m;
	pack_read_word(codes + off[r], (int)(off[r + 1] - off[r]), i, b, m);
	pk[2 * t] = b; pk[2 * t + 1] = m;
}

template <int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) k_sweep_fwd(SweepArgs a)
{
	FwdLane ln;
	ln.st = FwdLane::DONE; ln.n_out = 0; ln.n_sweeps = 0; ln.over = 0;
	int r = -1;
	bool need = false, drained = false;
	int64_t blocks = 0;
	for (;;) {
		while (!need && !drained) {
			if (r >= 0) {
				a.n_intv[r] = ln.n_out; a.n_sweeps[r] = ln.over ? -1 : ln.n_sweeps; if (MODE == 1) a.n_first[r] = ln.n_out;
				if (ln.over) atomicAdd(a.n_over, 1);
				if (ln.n_out > a.cap) atomicMax(a.worst, ln.n_out);
			}
			r = atomicAdd(a.next_read, 1);
			if (r >= a.n_reads) { r = -1; drained = true; break; }
			if (MODE == 2 && a.n_sweeps[r] < 0) { r = -1; continue; }       // already handed to the general kernel
			const PackedRead pr = { a.pk + (int64_t)r * a.pk_stride * 2, a.pk_stride };
			ln.begin(a.so, a.fm, MODE, (int)(a.off[r + 1] - a.off[r]), a.codes + a.off[r], pr, a.out + (int64_t)r * a.cap,
			         a.strips + (int64_t)r * a.strip_cap, a.strip_cap, MODE == 1 ? 0 : a.n_intv[r], MODE == 1 ? 0 : a.n_first[r]);
			if (MODE == 2 && ln.n_out > a.cap) ln.st = FwdLane::DONE;        // output overflow: the whole batch is rerun anyway
			need = ln.advance(a.fm, a.so);
		}
		if (!__any_sync(0xffffffffu, need)) break;
		if (need) {
			uint64_t o0, o1, o2;
			fwd_lane_fetch(a.fm, ln, o0, o1, o2, blocks);
			if (!ln.step(a.fm, a.so, a.cap, o0, o1, o2)) need = ln.advance(a.fm, a.so);
		}
	}
	for (int o = 16; o > 0; o >>= 1) blocks += __shfl_down_sync(0xffffffffu, blocks, o);
	if ((threadIdx.x & 31) == 0 && blocks) atomicAdd(a.occ_blocks, (unsigned long long)blocks);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_sweep_bwd(SweepArgs a)
{
	__shared__ uint32_t traj_sh[2 * BwdLane::TRAJ * 128];
	BwdLane ln;
	ln.st = BwdLane::DONE; ln.n_out = 0;
	int r = -1;
	bool need = false, drained = false;
	int64_t blocks = 0;
	for (;;) {
		while (!need && !drained) {
			if (r >= 0) {
				a.n_intv[r] = ln.n_out;
				if (ln.over) { a.n_sweeps[r] = -1; atomicAdd(a.n_over, 1); }
				else if (ln.n_out > a.cap) atomicMax(a.worst, ln.n_out);
			}
			r = atomicAdd(a.next_read, 1);
			if (r >= a.n_reads) { r = -1; drained = true; break; }
			const int ns = a.n_sweeps[r];
			if (ns <= 0) { r = -1; continue; }
			const PackedRead pr = { a.pk + (int64_t)r * a.pk_stride * 2, a.pk_stride };
			ln.begin(a.so, a.fm.kmax, a.fm.bloom && a.so.min_seed_len >= a.fm.bloom_k ? a.fm.bloom_k : 0, (int)(a.off[r + 1] - a.off[r]), a.codes + a.off[r], pr, a.out + (int64_t)r * a.cap, a.strips + (int64_t)r * a.strip_cap, ns, a.n_intv[r], traj_sh + threadIdx.x, 128);
			need = ln.advance(a.so, a.cap);
		}
		if (!__any_sync(0xffffffffu, need)) break;
		if (need) {
			uint64_t o0, o1, o2;
			bwd_lane_fetch(a.fm, ln, o0, o1, o2, blocks);
			if (!ln.step(a.so, a.cap, o0, o1, o2)) need = ln.advance(a.so, a.cap);
		}
	}
	for (int o = 16; o > 0; o >>= 1) blocks += __shfl_down_sync(0xffffffffu, blocks, o);
	if ((threadIdx.x & 31) == 0 && blocks) atomicAdd(a.occ_blocks, (unsigned long long)blocks);
}
#endif

} // namespace b200
