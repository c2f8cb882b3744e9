// finish_stage.h - the stages of mem_process_seqs after seed extension, as ONE sequence of batched tasks over data that never
// leaves the device: region de-duplication, insert-size candidates, mate rescue (plan -> ksw_align2 batch -> replay), primary
// marking / pairing / mapQ / record plan, CIGAR stage (ksw_global2 + traceback), NM/MD, SAM text.  The only host work in the
// sequence is mem_pestat's arithmetic on the candidate list (double precision, a few microseconds) and a handful of counter
// reads that size the next launch.
//
// The sequence is written once, over a small backend interface (BK), and instantiated twice:
//   stages_cuda.cu       BK = the CUDA engine: every run() is a kernel launch (k_task<TASK>), scans and sorts are CUB
//   tests/hostemu        BK = plain loops on the CPU (test scaffold for GPU-less containers; never part of the product)
//
// BK provides:
//   T *buf<T>(id, n)                     scratch for n elements, valid until the same id is asked for again (contents not kept)
//   T *grow<T>(id, n, keep)              same, keeping the first `keep` elements
//   void run(n, task)                    task(i) for every i in [0, n)
//   void scan(in, out, n)                out[0..n) = exclusive prefix sums (int64) of the n int32 at in[]
//   void sort_pairs(key, val, n)         sort (uint32 key, int32 val) pairs by ascending key
//   void zero(p, bytes), int64 get64(p), int32 get32(p), void upload(dst, src, bytes), void download(dst, src, bytes)
//   void sw_launch(so, src, order, cnt[5], max_t, max_q)            the ksw_align2 kernels over classified jobs
//   void global_launch(go, jobs, order, cnt[6], qmax[6], z, cig, res, squeeze_window)   the ksw_global2 kernels
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "finish_kernels.h"

namespace b200 {

#if defined(__CUDA_ARCH__)
#define FIN_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define FIN_ATOMIC_MAX(p, v) atomicMax((p), (v))
#else
#define FIN_ATOMIC_ADD(p, v) __sync_fetch_and_add((p), (v))
static inline int fin_host_atomic_max(int32_t *p, int32_t v) { int32_t o = *p; if (v > o) *p = v; return o; }
#define FIN_ATOMIC_MAX(p, v) fin_host_atomic_max((p), (v))
#endif

enum FinBuf {
	FB_R1 = 0, FB_TMP, FB_IX, FB_N1, FB_PATCH, FB_CTR, FB_ROWS, FB_CAND, FB_NJOB, FB_NANCH, FB_CAP, FB_ROFF, FB_JOFF, FB_R, FB_NREG,
	FB_JOBS, FB_KEYS, FB_RES, FB_NEXT, FB_HEAD, FB_PEND0, FB_PEND1, FB_MISS, FB_ORDER, FB_Z, FB_CNT, FB_HASALT, FB_V, FB_VTMP, FB_XAOF, FB_NEED,
	FB_RECS, FB_NREC, FB_MATE, FB_CN_ALN, FB_CN_DP, FB_CN_CIG, FB_CN_MD, FB_CN_Z, FB_O_ALN, FB_O_DP, FB_O_CIG, FB_O_MD, FB_O_Z, FB_ALNSLOT,
	FB_SLOTREG, FB_SLOTJOB, FB_SLOTCIG, FB_SLOTMD, FB_GJOBS, FB_GRES, FB_GKEY, FB_GSEL, FB_GSEL2, FB_GZ, FB_CIG, FB_MD, FB_ALN, FB_LEN, FB_SAMOFF, FB_SAM,
	FB_PAIRTAB, FB_LINEOFF, FB_LINES, FB_RKEY, FB_RVAL, FB_RLEN, FB_RTOFF, FB_DESTOFF, FB_ROUTED, FB_N
};

inline SwOpt fin_sw_opt(const mem_opt_t &opt)
{
	SwOpt s;
	s.o_del = opt.o_del; s.e_del = opt.e_del; s.o_ins = opt.o_ins; s.e_ins = opt.e_ins;
	int mn = 127, mx = 0;
	for (int i = 0; i < 25; ++i) { s.mat[i] = opt.mat[i]; if (opt.mat[i] < mn) mn = opt.mat[i]; if (opt.mat[i] > mx) mx = opt.mat[i]; }
	s.max_sc = mx; s.shift = (256 - (mn & 0xff)) & 0xff;
	return s;
}

/* ---------------------------------------------------------------- tasks */

struct DedupTask {          // one read: regions as mem_chain2aln left them -> mem_align1_core's list (first pass: no row scratch)
	FinCtx cx; const int64_t *xoff; const DReg *xregs; Reg *r1, *tmp; int32_t *ix, *n1, *patch_list, *ctr;
	B200_HD void operator()(int64_t r) const
	{
		const int64_t b = xoff[r];
		const int n = (int)(xoff[r + 1] - b);
		bool need_dp = false;
		n1[r] = n ? regs_from_ext(cx, cx.codes + cx.off[r], n, xregs + b, r1 + b, tmp + b, ix + b, nullptr, 0, &need_dp) : 0;
		if (need_dp) patch_list[FIN_ATOMIC_ADD(ctr, 1)] = (int32_t)r;
	}
};

struct DedupPatchTask {     // the reads whose mem_patch_reg needs a banded DP, rerun with row scratch
	FinCtx cx; const int64_t *xoff; const DReg *xregs; Reg *r1, *tmp; int32_t *ix, *n1; const int32_t *patch_list; int32_t *rows; int64_t stride;
	B200_HD void operator()(int64_t t) const
	{
		const int64_t r = patch_list[t], b = xoff[r];
		bool need_dp = false;
		n1[r] = regs_from_ext(cx, cx.codes + cx.off[r], (int)(xoff[r + 1] - b), xregs + b, r1 + b, tmp + b, ix + b, rows + t, stride, &need_dp);
	}
};

struct PestatTask {
	FinCtx cx; const int64_t *xoff; const Reg *r1; const int32_t *n1; uint64_t *cand;
	B200_HD void operator()(int64_t p) const
	{
		cand[p] = pestat_candidate(cx, n1[p << 1], r1 + xoff[p << 1], n1[p << 1 | 1], r1 + xoff[p << 1 | 1]);
	}
};

struct RescueCountTask {    // per pair: round-0 job count, anchors per end, region capacity of both reads
	FinCtx cx; FinTables tb; const int64_t *xoff; const Reg *r1; const int32_t *n1; int32_t *njob, *cap; int rescue;
	B200_HD void operator()(int64_t p) const
	{
		const int n[2] = { n1[p << 1], n1[p << 1 | 1] };
		const Reg *const a[2] = { r1 + xoff[p << 1], r1 + xoff[p << 1 | 1] };
		if (!rescue) { njob[p] = 0; cap[p << 1] = n[0]; cap[p << 1 | 1] = n[1]; return; }
		njob[p] = rescue_plan_pair(cx, tb.pes, p, n, a, nullptr, nullptr);
		cap[p << 1] = n[0] + 4 * rescue_n_anchors(cx.opt, n[1], a[1]);
		cap[p << 1 | 1] = n[1] + 4 * rescue_n_anchors(cx.opt, n[0], a[0]);
	}
};

struct RescueEmitTask {
	FinCtx cx; FinTables tb; const int64_t *xoff; const Reg *r1; const int32_t *n1; const int64_t *joff; SwJob *jobs; int32_t *keys, *head;
	B200_HD void operator()(int64_t p) const
	{
		const int n[2] = { n1[p << 1], n1[p << 1 | 1] };
		const Reg *const a[2] = { r1 + xoff[p << 1], r1 + xoff[p << 1 | 1] };
		head[p] = -1;
		if (joff[p + 1] > joff[p]) rescue_plan_pair(cx, tb.pes, p, n, a, jobs + joff[p], keys + joff[p]);
	}
};

struct SwClassTask {        // strip-width class of every job (sw_warp_kernel.cuh), class counts, longest target / general-path query
	const SwJob *jobs; int32_t *cls, *ctr;      // ctr[0..4] counts, ctr[5] max tlen, ctr[6] max qlen of class 0
	B200_HD void operator()(int64_t i) const
	{
		const SwJob j = jobs[i];
		const int c = sw_warp_class(j.q_len, j.xtra);
		const int k = c == 0 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 5 ? 3 : 4;
		cls[i] = k;
		FIN_ATOMIC_ADD(ctr + k, 1);
		FIN_ATOMIC_MAX(ctr + 5, j.tlen);
		if (k == 0) FIN_ATOMIC_MAX(ctr + 6, j.q_len);
	}
};
struct SwScatterTask {
	const int32_t *cls; int32_t *cur, *order;   // cur[k]: next free position of class k (starts at the class base)
	B200_HD void operator()(int64_t i) const { order[FIN_ATOMIC_ADD(cur + cls[i], 1)] = (int32_t)i; }
};

struct RescueReplayTask {   // one pair: mem_matesw's insert/skip logic over the SW results; a pair that misses a result is queued
	FinCtx cx; FinTables tb; const int64_t *xoff; const Reg *r1; const int32_t *n1; const int64_t *roff; Reg *R, *tmp; int32_t *ix, *nreg;
	const int64_t *joff; const SwJob *jobs; const int32_t *keys; const SwRes *res; const int32_t *next, *head;
	const int32_t *pend_in; int32_t *pend_out, *miss, *ctr;         // pend_in == null: every pair
	int hide;                                                       // (tests) pretend the last `hide` round-0 results of every pair are missing
	B200_HD void operator()(int64_t t) const
	{
		const int64_t p = pend_in ? pend_in[t] : t;
		const int n[2] = { n1[p << 1], n1[p << 1 | 1] };
		const Reg *const src[2] = { r1 + xoff[p << 1], r1 + xoff[p << 1 | 1] };
		Reg *const w[2] = { R + roff[p << 1], R + roff[p << 1 | 1] };
		int nw[2];
		int jn = (int)(joff[p + 1] - joff[p]);
		jn = jn > hide ? jn - hide : 0;
		RescueHave have = { jobs, keys, res, next, joff[p], jn, head[p] };
		// (scratch: the pair's own region slots in tmp / ix - a list never outgrows its capacity)
		const int32_t key = rescue_replay_pair(cx, tb.pes, p, n, src, w, nw, tmp + roff[p << 1], ix + roff[p << 1], have);
		if (key < 0) { nreg[p << 1] = nw[0]; nreg[p << 1 | 1] = nw[1]; }
		else { pend_out[FIN_ATOMIC_ADD(ctr, 1)] = (int32_t)p; miss[p] = key; }
	}
};

struct RescueExtraTask {    // one more job for every queued pair
	FinCtx cx; FinTables tb; const int64_t *xoff; const Reg *r1; const int32_t *n1; const int32_t *pend, *miss; SwJob *jobs; int32_t *keys, *next, *head;
	int64_t j0;
	B200_HD void operator()(int64_t t) const
	{
		const int64_t p = pend[t], x = j0 + t;
		const int n[2] = { n1[p << 1], n1[p << 1 | 1] };
		const Reg *const a[2] = { r1 + xoff[p << 1], r1 + xoff[p << 1 | 1] };
		rescue_extra_job(cx, tb.pes, p, miss[p], n, a, &jobs[x]);
		keys[x] = miss[p]; next[x] = head[p]; head[p] = (int32_t)x;
	}
};

struct CopyRegsTask {       // no rescue: the working lists are the de-duplicated ones
	const int64_t *xoff, *roff; const Reg *r1; const int32_t *n1; Reg *R; int32_t *nreg;
	B200_HD void operator()(int64_t r) const
	{
		const int n = n1[r];
		for (int k = 0; k < n; ++k) R[roff[r] + k] = r1[xoff[r] + k];
		nreg[r] = n;
	}
};

struct DecideTask {         // one pair (or one single-end read): final region state and the record plan
	FinCtx cx; FinTables tb; const int64_t *roff; Reg *R, *tmp; int32_t *ix, *z, *cnt, *has_alt; FinPair64 *v, *vtmp;
	int32_t *nreg, *xa_of; uint8_t *need; SamRec *recs; int32_t *nrec, *mate;
	B200_HD void operator()(int64_t u) const
	{
		if (cx.pe) {
			const int64_t r0 = u << 1, b = roff[r0];
			const int n[2] = { nreg[r0], nreg[r0 + 1] };
			Reg *const a[2] = { R + roff[r0], R + roff[r0 + 1] };
			int32_t *const xo[2] = { xa_of + roff[r0], xa_of + roff[r0 + 1] };
			uint8_t *const nd[2] = { need + roff[r0], need + roff[r0 + 1] };
			SamRec *const rc[2] = { recs + roff[r0] + r0, recs + roff[r0 + 1] + r0 + 1 };
			int n_rec[2] = { 0, 0 }, mate_reg[2] = { -1, -1 };
			const PairScratch S = { tmp + b, ix + b, z + b, cnt + b, has_alt + b, v + b, vtmp + b };
			if (fin_names_differ(cx, r0)) fin_fail(cx, FIN_ERR_NAMES, (int)r0);
			pair_decide(cx, tb, (uint64_t)((cx.n_processed >> 1) + u), n, a, xo, nd, rc, n_rec, mate_reg, S);
			nrec[r0] = n_rec[0]; nrec[r0 + 1] = n_rec[1]; mate[r0] = mate_reg[0]; mate[r0 + 1] = mate_reg[1];
		} else {
			const int64_t b = roff[u];
			const PairScratch S = { tmp + b, ix + b, z + b, cnt + b, has_alt + b, v + b, vtmp + b };
			int n_rec = 0;
			single_decide(cx, tb, cx.n_processed + u, nreg[u], R + b, xa_of + b, need + b, recs + b + u, &n_rec, S);
			nrec[u] = n_rec; mate[u] = -2;
		}
	}
};

struct AlnCountTask {       // per read: flagged regions, those that need the DP, and the arena room they take
	FinCtx cx; const int64_t *roff; const Reg *R; const int32_t *nreg; const uint8_t *need; int32_t *c_aln, *c_dp, *c_cig, *c_md, *c_z;
	B200_HD void operator()(int64_t r) const
	{
		const GlobalOpt go = fin_global_opt(cx.opt);
		int na = 0, nd = 0;
		int64_t cig = 0, md = 0, z = 0;
		for (int k = 0; k < nreg[r]; ++k) {
			if (!(need[roff[r] + k] & REG_ALN)) continue;
			const Reg *ar = &R[roff[r] + k];
			++na; cig += aln_cigar_cap(ar); md += aln_md_cap(ar);
			int w2, wmax;
			if (aln_needs_dp(cx, ar, &w2)) { ++nd; z += (global_z_need(go, ar->qe - ar->qb, (int)(ar->re - ar->rb), w2, &wmax) + 15) & ~(int64_t)15; }
		}
		c_aln[r] = na; c_dp[r] = nd; c_cig[r] = (int32_t)cig; c_md[r] = (int32_t)md; c_z[r] = (int32_t)z;
	}
};

struct AlnEmitTask {        // per read: alignment slots of the flagged regions and the jobs of the CIGAR stage
	FinCtx cx; const int64_t *roff; const Reg *R; const int32_t *nreg; const uint8_t *need;
	const int64_t *o_aln, *o_dp, *o_cig, *o_md, *o_z;
	int32_t *aln_slot; int64_t *slot_reg; int32_t *slot_job; int64_t *slot_cig, *slot_md; GlobalJob *gjobs;
	B200_HD void operator()(int64_t r) const
	{
		const GlobalOpt go = fin_global_opt(cx.opt);
		int64_t slot = o_aln[r], job = o_dp[r], cig = o_cig[r], md = o_md[r], z = o_z[r];
		for (int k = 0; k < nreg[r]; ++k) {
			const int64_t g = roff[r] + k;
			aln_slot[g] = -1;
			if (!(need[g] & REG_ALN)) continue;
			const Reg *ar = &R[g];
			aln_slot[g] = (int32_t)slot; slot_reg[slot] = g; slot_cig[slot] = cig; slot_md[slot] = md; slot_job[slot] = -1;
			int w2, wmax;
			if (aln_needs_dp(cx, ar, &w2)) {
				GlobalJob j;
				j.rb = ar->rb; j.re = ar->re; j.zoff = z; j.read = (int32_t)r; j.qb = ar->qb; j.qe = ar->qe; j.w2 = w2; j.truesc = ar->truesc;
				z += (global_z_need(go, ar->qe - ar->qb, (int)(ar->re - ar->rb), w2, &wmax) + 15) & ~(int64_t)15;
				j.wmax = wmax; j.cig_off = cig + 1;
				gjobs[job] = j; slot_job[slot] = (int32_t)job; ++job;
			}
			cig += aln_cigar_cap(ar); md += aln_md_cap(ar); ++slot;
		}
	}
};

static const int kGlobalClsS[5] = { 32, 64, 128, 256, 512 };
struct GlobalClassTask {    // class (row window), band and length key of every selected job; class counts and longest query per class
	GlobalOpt go; const GlobalJob *jobs; const int32_t *sel; uint32_t *key; int32_t *ctr; int pass, squeeze;     // ctr[0..5] counts, ctr[6..11] qmax
	B200_HD void operator()(int64_t x) const
	{
		const GlobalJob j = jobs[sel[x]];
		const int ql = j.qe - j.qb, rl = (int)(j.re - j.rb);
		int band = j.truesc == B200_GLOBAL_RAW ? j.w2 : pass == 0 ? global_band(go, ql, rl, j.w2 < go.w_max ? j.w2 : go.w_max) : j.wmax;
		if (pass == 0 && squeeze) band >>= 2;
		const int need = 2 * band + 2;
		int k = 5;
		if (ql <= 256) { k = need <= 32 ? 0 : need <= 64 ? 1 : need <= 128 ? 2 : need <= 256 ? 3 : need <= 512 ? 4 : 5; }
		FIN_ATOMIC_ADD(ctr + k, 1);
		FIN_ATOMIC_MAX(ctr + 6 + k, ql);
		key[x] = (uint32_t)((k * 256 + (band < 255 ? band : 255)) * 64 + (rl >> 4 < 63 ? rl >> 4 : 63));
	}
};
struct IotaTask { int32_t *p; B200_HD void operator()(int64_t i) const { p[i] = (int32_t)i; } };
struct GlobalRerunTask {    // jobs whose retry outgrew the row window of their class
	const GlobalRes *res; int32_t *sel2, *ctr;
	B200_HD void operator()(int64_t i) const { if (res[i].n_cigar == -2) sel2[FIN_ATOMIC_ADD(ctr, 1)] = (int32_t)i; }
};

struct AlnFinishTask {      // per alignment slot: CIGAR (from the DP or the gap-free path) -> NM, MD, position, clips
	FinCtx cx; const Reg *R; const int64_t *slot_reg; const int32_t *slot_job; const int64_t *slot_cig, *slot_md; const GlobalRes *gres;
	const int64_t *roff; int n_reads; uint32_t *cig; char *md; AlnRes *aln;
	B200_HD void operator()(int64_t s) const
	{
		const int64_t g = slot_reg[s];
		const Reg *ar = &R[g];
		// the read of region slot g: the last r with roff[r] <= g
		int64_t lo = 0, hi = n_reads;
		while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (roff[mid] <= g) lo = mid; else hi = mid; }
		const int64_t r = lo;
		const int l_query = (int)(cx.off[r + 1] - cx.off[r]);
		int n_cigar;
		if (slot_job[s] >= 0) n_cigar = gres[slot_job[s]].n_cigar;
		else {
			const int64_t l_pac = cx.fm.l_pac;
			const bool ok = ar->qe > ar->qb && ar->rb < ar->re && !(ar->rb < l_pac && ar->re > l_pac) && ar->rb >= 0 && ar->re <= l_pac << 1;
			n_cigar = ok ? 1 : 0;
			cig[slot_cig[s] + 1] = (uint32_t)(ar->qe - ar->qb) << 4;
		}
		aln_finish(cx, ar, l_query, cx.codes + cx.off[r], n_cigar, cig, slot_cig[s], md, slot_md[s], &aln[s]);
	}
};

struct SamCountTask {
	FinCtx cx; SamView V; int32_t *len;
	B200_HD void operator()(int64_t r) const { CountSink s; sam_format_read(cx, V, r, s); len[r] = (int32_t)s.n; }
};
struct SamWriteTask {
	FinCtx cx; SamView V; const int64_t *sam_off; char *sam;
	const int64_t *line_off; SamLine *lines;      // (optional) routing table: lines of read r at lines[line_off[r] ..]
	B200_HD void operator()(int64_t r) const
	{
		WriteSink s(sam + sam_off[r]);
		if (!lines) { sam_format_read(cx, V, r, s); s.flush(); return; }
		const int n = V.nrec[r];
		int64_t at = sam_off[r];
		for (int w = 0; w < n; ++w) {
			int rid, mrid;
			sam_format(cx, V, r, w, s, &rid, &mrid);
			const int64_t end = (int64_t)(s.p - sam) + s.na;
			SamLine L;
			L.off = at; L.len = (int32_t)(end - at); L.rid = rid; L.mate_rid = mrid; L.read = (int32_t)r;
			lines[line_off[r] + w] = L;
			at = end;
		}
		s.flush();
	}
};

/* ---------------------------------------------------------------- per-chromosome routing of the SAM lines
 * (reference src/mainParallelByChromosome.c:1395-1457 parses RNAME / RNEXT back out of every line and copies the line into the buffer
 * of its contig, or of "unmapped"; in the branch without fixmate a line whose mate sits on another contig goes into "discordant" AS
 * WELL.  Here the writer of the line knows both contigs: a stable sort of (destination, line) pairs, a prefix sum of the line lengths
 * in that order and one copy per pair build the same buffers - destinations 0 .. n_ctg-1, n_ctg = discordant, n_ctg+1 = unmapped -
 * back to back, lines of a destination in output order.) */
enum { ROUTE_LINES = 1, ROUTE_BY_CONTIG = 2, ROUTE_DISCORDANT = 4 };
struct RouteKeyTask {
	const SamLine *lines; int64_t n_lines; int n_ctg, with_disc; uint32_t *key; int32_t *val;
	B200_HD void operator()(int64_t i) const
	{
		const SamLine L = lines[i];
		key[i] = L.rid < 0 ? (uint32_t)n_ctg + 1 : (uint32_t)L.rid;
		val[i] = (int32_t)i;
		const bool disc = with_disc && L.rid >= 0 && L.mate_rid >= 0 && L.rid != L.mate_rid;
		key[n_lines + i] = disc ? (uint32_t)n_ctg : (uint32_t)n_ctg + 2;      // (n_ctg + 2: not a destination, sorts last, length 0)
		val[n_lines + i] = (int32_t)i;
	}
};
struct RouteLenTask {
	const SamLine *lines; const uint32_t *key; const int32_t *val; int n_dest; int32_t *len;
	B200_HD void operator()(int64_t j) const { len[j] = key[j] < (uint32_t)n_dest ? lines[val[j]].len : 0; }
};
struct RouteBoundsTask {       // j in [0, m]: destinations in (key[j-1], key[j]] start at roff[j]
	const uint32_t *key; const int64_t *roff; int64_t m; int n_dest; int64_t *dest_off;
	B200_HD void operator()(int64_t j) const
	{
		const int64_t lo = j == 0 ? -1 : (key[j - 1] < (uint32_t)n_dest ? (int64_t)key[j - 1] : n_dest);
		const int64_t hi = j == m ? n_dest : (key[j] < (uint32_t)n_dest ? (int64_t)key[j] : n_dest);
		for (int64_t d = lo + 1; d <= hi; ++d) dest_off[d] = roff[j];
	}
};
struct RouteCopyTask {
	const SamLine *lines; const uint32_t *key; const int32_t *val; const int64_t *roff; int n_dest; const char *sam; char *routed;
	B200_HD void operator()(int64_t j) const
	{
		if (key[j] >= (uint32_t)n_dest) return;
		const SamLine L = lines[val[j]];
		WriteSink s(routed + roff[j]);
		s.puts(sam + L.off, L.len);
		s.flush();
	}
};

/* ---------------------------------------------------------------- host side of mem_pestat (reference src/bwamem_pair.c:67-109) */

// statistics from the candidates (dir << 32 | insert size; 0 = none), with the reference's progress lines on stderr
void pestat_from_candidates(const mem_opt_t *opt, int64_t n, const uint64_t *cand, mem_pestat_t pes[4]);

// .721 * log(2 erfc(|d - avg| / std / sqrt 2)) * a for every integer insert size d in [low, high] of each orientation (glibc)
inline void build_pair_tables(const mem_opt_t &opt, const mem_pestat_t pes[4], std::vector<double> &tab, int64_t off[4])
{
	tab.clear();
	for (int d = 0; d < 4; ++d) {
		off[d] = (int64_t)tab.size();
		if (pes[d].failed || pes[d].high < pes[d].low) continue;
		for (int64_t dist = pes[d].low; dist <= pes[d].high; ++dist) {
			const double ns = (dist - pes[d].avg) / pes[d].std;
			tab.push_back(.721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt.a);
		}
	}
	if (tab.empty()) tab.push_back(0.);
}

/* ---------------------------------------------------------------- the sequence */

struct FinishIn {
	const DReg *xregs; const int64_t *xoff;     // regions of the extension stage, compacted per read (device memory)
	const mem_pestat_t *pes0;                   // caller-given insert-size statistics, or null
	int max_len;
	const double *logtab; int n_log;            // device memory
	int route;                                  // ROUTE_* flags: also build the per-line routing table / the text grouped by contig
};
struct FinishOut {
	const char *sam;            // device memory: the chunk's SAM text, records of read r at [sam_off[r], sam_off[r+1])
	const int64_t *sam_off;     // device memory, n_reads + 1
	int64_t sam_bytes;
	const SamLine *lines;       // device memory (when asked for): one entry per SAM line, in output order
	int64_t n_lines;
	const char *routed;         // device memory (ROUTE_BY_CONTIG): the lines grouped by destination, destination d at [dest_off[d], dest_off[d+1])
	const int64_t *dest_off;    // n_ctg + 3 entries
	int64_t routed_bytes;
};

template <class BK>
void finish_run(BK &bk, FinCtx cx, const FinishIn &in, FinishOut &out, b200_stats_t &st, double (*clock_ms)())
{
	const mem_opt_t &opt = cx.opt;
	const int n = cx.n_reads;
	const int64_t n_units = cx.pe ? n >> 1 : n;
	int32_t *ctr = bk.template buf<int32_t>(FB_CTR, 64);
	bk.zero(ctr, 64 * sizeof(int32_t));
	cx.err = ctr + 32;
	double t0 = clock_ms(), t1;

	// ---- mem_sort_dedup_patch per read
	const int64_t n_x = bk.get64(in.xoff + n);
	Reg *r1 = bk.template buf<Reg>(FB_R1, n_x + 1);
	int32_t *n1 = bk.template buf<int32_t>(FB_N1, n + 1);
	{
		Reg *tmp = bk.template buf<Reg>(FB_TMP, n_x + 1);
		int32_t *ix = bk.template buf<int32_t>(FB_IX, n_x + 1);
		int32_t *patch_list = bk.template buf<int32_t>(FB_PATCH, n + 1);
		bk.run(n, DedupTask{ cx, in.xoff, in.xregs, r1, tmp, ix, n1, patch_list, ctr });
		const int np = bk.get32(ctr);
		if (np > 0) {
			int32_t *rows = bk.template buf<int32_t>(FB_ROWS, (size_t)np * 2 * (in.max_len + 2));
			bk.run(np, DedupPatchTask{ cx, in.xoff, in.xregs, r1, tmp, ix, n1, patch_list, rows, np });
		}
		st.n_patch_reads += np;
	}
	t1 = clock_ms(); st.ms_regs_host += t1 - t0; t0 = t1;

	// ---- insert-size statistics: candidates on the device, mem_pestat's arithmetic and the mem_pair table on the host
	FinTables tb;
	memset(&tb, 0, sizeof tb);
	tb.logtab = in.logtab; tb.n_log = in.n_log;
	{
		std::vector<double> tab;
		if (cx.pe) {
			if (in.pes0) memcpy(tb.pes, in.pes0, 4 * sizeof(mem_pestat_t));
			else {
				uint64_t *cand = bk.template buf<uint64_t>(FB_CAND, n_units + 1);
				bk.run(n_units, PestatTask{ cx, in.xoff, r1, n1, cand });
				std::vector<uint64_t> h(n_units);
				bk.download(h.data(), cand, sizeof(uint64_t) * n_units);
				pestat_from_candidates(&opt, n_units, h.data(), tb.pes);
			}
			for (int d = 0; d < 4; ++d)
				if (!tb.pes[d].failed && (int64_t)tb.pes[d].high - tb.pes[d].low > ((int64_t)1 << 26)) { fprintf(stderr, "[mpibwa_b200] insert-size range of %lld values is not supported\n", (long long)tb.pes[d].high - tb.pes[d].low); abort(); }
		} else for (int d = 0; d < 4; ++d) tb.pes[d].failed = 1;
		build_pair_tables(opt, tb.pes, tab, tb.pair_off);
		double *d_tab = bk.template buf<double>(FB_PAIRTAB, tab.size());
		bk.upload(d_tab, tab.data(), sizeof(double) * tab.size());
		tb.pair_tab = d_tab;
	}

	// ---- mate rescue
	const bool rescue = cx.pe && !(opt.flag & MEM_F_NO_RESCUE);
	int32_t *cap = bk.template buf<int32_t>(FB_CAP, n + 1);
	int64_t *roff = bk.template buf<int64_t>(FB_ROFF, n + 1);
	int32_t *nreg = bk.template buf<int32_t>(FB_NREG, n + 1);
	Reg *R = nullptr, *tmp = nullptr;
	int32_t *ix = nullptr;
	int64_t n_slots = 0;
	if (cx.pe) {
		int32_t *njob = bk.template buf<int32_t>(FB_NJOB, n_units + 1);
		int64_t *joff = bk.template buf<int64_t>(FB_JOFF, n_units + 1);
		bk.run(n_units, RescueCountTask{ cx, tb, in.xoff, r1, n1, njob, cap, rescue ? 1 : 0 });
		bk.zero(cap + n, sizeof(int32_t)); bk.zero(njob + n_units, sizeof(int32_t));
		bk.scan(cap, roff, n + 1);
		bk.scan(njob, joff, n_units + 1);
		n_slots = bk.get64(roff + n);
		const int64_t J0 = bk.get64(joff + n_units);
		R = bk.template buf<Reg>(FB_R, n_slots + 1);
		tmp = bk.template buf<Reg>(FB_TMP, n_slots + 1);
		ix = bk.template buf<int32_t>(FB_IX, n_slots + 1);
		if (rescue) {
			int64_t jcap = J0 + 1024;
			SwJob *jobs = bk.template buf<SwJob>(FB_JOBS, jcap);
			int32_t *keys = bk.template buf<int32_t>(FB_KEYS, jcap), *next = bk.template buf<int32_t>(FB_NEXT, jcap);
			SwRes *res = bk.template buf<SwRes>(FB_RES, jcap);
			int32_t *head = bk.template buf<int32_t>(FB_HEAD, n_units + 1), *miss = bk.template buf<int32_t>(FB_MISS, n_units + 1);
			int32_t *pend[2] = { bk.template buf<int32_t>(FB_PEND0, n_units + 1), bk.template buf<int32_t>(FB_PEND1, n_units + 1) };
			bk.run(n_units, RescueEmitTask{ cx, tb, in.xoff, r1, n1, joff, jobs, keys, head });
			const SwOpt so = fin_sw_opt(opt);
			auto run_sw = [&](int64_t j0, int64_t nj) {
				if (nj <= 0) return;
				int32_t *cls = bk.template buf<int32_t>(FB_Z, nj);
				int32_t *order = bk.template buf<int32_t>(FB_ORDER, nj);
				bk.zero(ctr + 8, 16 * sizeof(int32_t));
				bk.run(nj, SwClassTask{ jobs + j0, cls, ctr + 8 });
				int32_t h[8];
				bk.download(h, ctr + 8, sizeof h);
				int32_t base[5];
				base[0] = 0;
				for (int k = 1; k < 5; ++k) base[k] = base[k - 1] + h[k - 1];
				bk.upload(ctr + 16, base, sizeof base);
				bk.run(nj, SwScatterTask{ cls, ctr + 16, order });
				bk.sw_launch(so, jobs + j0, res + j0, order, h, h[5], h[6]);
				st.n_sw_jobs += nj;
			};
			run_sw(0, J0);
			// round 0 replays every pair; a pair that misses a result (a rescued hit displaced the region that had ruled an
			// orientation out) gets one more job per round
			int64_t jn = J0, n_pend = 0;
			bk.zero(ctr + 1, sizeof(int32_t));
			const int hide = getenv("B200_RESCUE_HIDE") ? atoi(getenv("B200_RESCUE_HIDE")) : 0;     // (tests: forces the extra rounds)
			bk.run(n_units, RescueReplayTask{ cx, tb, in.xoff, r1, n1, roff, R, tmp, ix, nreg, joff, jobs, keys, res, next, head, nullptr, pend[0], miss, ctr + 1, hide });
			n_pend = bk.get32(ctr + 1);
			int cur = 0;
			while (n_pend > 0) {
				if (jn + n_pend > jcap) {
					const int64_t want = jn + n_pend + (jn >> 2) + 1024;
					jobs = bk.template grow<SwJob>(FB_JOBS, want, jn); keys = bk.template grow<int32_t>(FB_KEYS, want, jn);
					next = bk.template grow<int32_t>(FB_NEXT, want, jn); res = bk.template grow<SwRes>(FB_RES, want, jn);
					jcap = want;
				}
				bk.run(n_pend, RescueExtraTask{ cx, tb, in.xoff, r1, n1, pend[cur], miss, jobs, keys, next, head, jn });
				run_sw(jn, n_pend);
				jn += n_pend;
				bk.zero(ctr + 1, sizeof(int32_t));
				bk.run(n_pend, RescueReplayTask{ cx, tb, in.xoff, r1, n1, roff, R, tmp, ix, nreg, joff, jobs, keys, res, next, head, pend[cur], pend[cur ^ 1], miss, ctr + 1, hide });
				n_pend = bk.get32(ctr + 1);
				cur ^= 1;
				++st.n_rescue_rounds;
			}
		} else bk.run(n, CopyRegsTask{ in.xoff, roff, r1, n1, R, nreg });
	} else {
		// single-end: the de-duplicated lists are final
		bk.zero(n1 + n, sizeof(int32_t));
		bk.scan(n1, roff, n + 1);
		n_slots = bk.get64(roff + n);
		R = bk.template buf<Reg>(FB_R, n_slots + 1);
		tmp = bk.template buf<Reg>(FB_TMP, n_slots + 1);
		ix = bk.template buf<int32_t>(FB_IX, n_slots + 1);
		bk.run(n, CopyRegsTask{ in.xoff, roff, r1, n1, R, nreg });
	}
	t1 = clock_ms(); st.ms_rescue += t1 - t0; t0 = t1;

	// ---- primary marking, pairing, mapQ, record plan
	int32_t *xa_of = bk.template buf<int32_t>(FB_XAOF, n_slots + 1);
	uint8_t *need = bk.template buf<uint8_t>(FB_NEED, n_slots + 1);
	SamRec *recs = bk.template buf<SamRec>(FB_RECS, n_slots + n + 1);
	int32_t *nrec = bk.template buf<int32_t>(FB_NREC, n + 1), *mate = bk.template buf<int32_t>(FB_MATE, n + 1);
	{
		int32_t *z = bk.template buf<int32_t>(FB_Z, n_slots + 1), *cnt = bk.template buf<int32_t>(FB_CNT, n_slots + 1);
		int32_t *has_alt = bk.template buf<int32_t>(FB_HASALT, n_slots + 1);
		FinPair64 *v = bk.template buf<FinPair64>(FB_V, n_slots + 1), *vtmp = bk.template buf<FinPair64>(FB_VTMP, n_slots + 1);
		bk.run(n_units, DecideTask{ cx, tb, roff, R, tmp, ix, z, cnt, has_alt, v, vtmp, nreg, xa_of, need, recs, nrec, mate });
	}
	t1 = clock_ms(); st.ms_sam_plan += t1 - t0; t0 = t1;

	// ---- alignments of the flagged regions: slots, arenas, jobs of the CIGAR stage
	int32_t *c_aln = bk.template buf<int32_t>(FB_CN_ALN, n + 1), *c_dp = bk.template buf<int32_t>(FB_CN_DP, n + 1), *c_cig = bk.template buf<int32_t>(FB_CN_CIG, n + 1);
	int32_t *c_md = bk.template buf<int32_t>(FB_CN_MD, n + 1), *c_z = bk.template buf<int32_t>(FB_CN_Z, n + 1);
	int64_t *o_aln = bk.template buf<int64_t>(FB_O_ALN, n + 1), *o_dp = bk.template buf<int64_t>(FB_O_DP, n + 1), *o_cig = bk.template buf<int64_t>(FB_O_CIG, n + 1);
	int64_t *o_md = bk.template buf<int64_t>(FB_O_MD, n + 1), *o_z = bk.template buf<int64_t>(FB_O_Z, n + 1);
	bk.run(n, AlnCountTask{ cx, roff, R, nreg, need, c_aln, c_dp, c_cig, c_md, c_z });
	int32_t *cs[5] = { c_aln, c_dp, c_cig, c_md, c_z };
	int64_t *os[5] = { o_aln, o_dp, o_cig, o_md, o_z };
	int64_t tot[5];
	for (int k = 0; k < 5; ++k) { bk.zero(cs[k] + n, sizeof(int32_t)); bk.scan(cs[k], os[k], n + 1); }
	for (int k = 0; k < 5; ++k) tot[k] = bk.get64(os[k] + n);
	const int64_t n_aln = tot[0], n_dp = tot[1];
	int32_t *aln_slot = bk.template buf<int32_t>(FB_ALNSLOT, n_slots + 1);
	int64_t *slot_reg = bk.template buf<int64_t>(FB_SLOTREG, n_aln + 1), *slot_cig = bk.template buf<int64_t>(FB_SLOTCIG, n_aln + 1), *slot_md = bk.template buf<int64_t>(FB_SLOTMD, n_aln + 1);
	int32_t *slot_job = bk.template buf<int32_t>(FB_SLOTJOB, n_aln + 1);
	GlobalJob *gjobs = bk.template buf<GlobalJob>(FB_GJOBS, n_dp + 1);
	GlobalRes *gres = bk.template buf<GlobalRes>(FB_GRES, n_dp + 1);
	uint32_t *cig = bk.template buf<uint32_t>(FB_CIG, tot[2] + 4);
	char *md = bk.template buf<char>(FB_MD, tot[3] + 4);
	AlnRes *aln = bk.template buf<AlnRes>(FB_ALN, n_aln + 1);
	bk.run(n, AlnEmitTask{ cx, roff, R, nreg, need, o_aln, o_dp, o_cig, o_md, o_z, aln_slot, slot_reg, slot_job, slot_cig, slot_md, gjobs });
	t1 = clock_ms(); st.ms_sam_plan += t1 - t0; t0 = t1;

	// ---- CIGAR stage: classes by row window; regions whose band-doubling retry outgrows the window rerun in a wider class
	if (n_dp > 0) {
		const GlobalOpt go = fin_global_opt(opt);
		uint8_t *z = bk.template buf<uint8_t>(FB_GZ, (size_t)tot[4] + 64);
		uint32_t *key = bk.template buf<uint32_t>(FB_GKEY, n_dp);
		int32_t *sel = bk.template buf<int32_t>(FB_GSEL, n_dp), *sel2 = bk.template buf<int32_t>(FB_GSEL2, n_dp);
		bk.run(n_dp, IotaTask{ sel });
		int64_t m = n_dp;
		const int squeeze = getenv("B200_GLOBAL_SQUEEZE") != nullptr;
		for (int pass = 0; pass < 2 && m > 0; ++pass) {
			int32_t *sl = pass == 0 ? sel : sel2;
			bk.zero(ctr + 8, 16 * sizeof(int32_t));
			bk.run(m, GlobalClassTask{ go, gjobs, sl, key, ctr + 8, pass, squeeze });
			bk.sort_pairs(key, sl, m);
			int32_t h[12];
			bk.download(h, ctr + 8, sizeof h);
			bk.global_launch(go, gjobs, sl, h, h + 6, z, cig, gres);
			if (pass == 0) {
				bk.zero(ctr + 2, sizeof(int32_t));
				bk.run(n_dp, GlobalRerunTask{ gres, sel2, ctr + 2 });
				m = bk.get32(ctr + 2);
				st.n_global_rerun += m;
			}
		}
		st.n_global_jobs += n_dp;
	}
	t1 = clock_ms(); st.ms_global += t1 - t0; t0 = t1;

	// ---- NM / MD / position / clips, then the SAM text at prefix-sum offsets
	bk.run(n_aln, AlnFinishTask{ cx, R, slot_reg, slot_job, slot_cig, slot_md, gres, roff, n, cig, md, aln });
	SamView V = { R, roff, nreg, xa_of, aln_slot, aln, cig, md, recs, nrec, mate };
	int32_t *len = bk.template buf<int32_t>(FB_LEN, n + 1);
	int64_t *sam_off = bk.template buf<int64_t>(FB_SAMOFF, n + 1);
	bk.run(n, SamCountTask{ cx, V, len });
	bk.zero(len + n, sizeof(int32_t));
	bk.scan(len, sam_off, n + 1);
	const int64_t total = bk.get64(sam_off + n);
	char *sam = bk.template buf<char>(FB_SAM, total + 16);
	int64_t *line_off = nullptr;
	SamLine *lines = nullptr;
	int64_t n_lines = 0;
	if (in.route) {
		line_off = bk.template buf<int64_t>(FB_LINEOFF, n + 1);
		bk.zero(nrec + n, sizeof(int32_t));
		bk.scan(nrec, line_off, n + 1);
		n_lines = bk.get64(line_off + n);
		lines = bk.template buf<SamLine>(FB_LINES, n_lines + 1);
	}
	bk.run(n, SamWriteTask{ cx, V, sam_off, sam, line_off, lines });
	int32_t err[2];
	bk.download(err, ctr + 32, sizeof err);
	if (err[0] == FIN_ERR_NAMES) { fprintf(stderr, "[mem_sam_pe] paired reads have different names (reads %d and %d of the chunk)\n", err[1], err[1] + 1); abort(); }
	if (err[0]) { fprintf(stderr, "[mpibwa_b200] finish stage: table range exceeded (code %d, value %d)\n", err[0], err[1]); abort(); }
	t1 = clock_ms(); st.ms_sam_host += t1 - t0; t0 = t1;
	out.routed = nullptr; out.dest_off = nullptr; out.routed_bytes = 0;
	if ((in.route & ROUTE_BY_CONTIG) && n_lines > 0) {
		const int n_ctg = cx.fm.n_ctg, n_dest = n_ctg + 2;
		if (n_dest + 1 >= (1 << 20)) { fprintf(stderr, "[mpibwa_b200] per-contig routing: too many contigs (%d)\n", n_ctg); abort(); }
		const int64_t m = 2 * n_lines;
		uint32_t *key = bk.template buf<uint32_t>(FB_RKEY, m);
		int32_t *val = bk.template buf<int32_t>(FB_RVAL, m);
		bk.run(n_lines, RouteKeyTask{ lines, n_lines, n_ctg, (in.route & ROUTE_DISCORDANT) ? 1 : 0, key, val });
		bk.sort_pairs(key, val, m);
		int32_t *rlen = bk.template buf<int32_t>(FB_RLEN, m + 1);
		int64_t *roff = bk.template buf<int64_t>(FB_RTOFF, m + 1);
		bk.run(m, RouteLenTask{ lines, key, val, n_dest, rlen });
		bk.zero(rlen + m, sizeof(int32_t));
		bk.scan(rlen, roff, m + 1);
		const int64_t rtotal = bk.get64(roff + m);
		int64_t *dest_off = bk.template buf<int64_t>(FB_DESTOFF, n_dest + 1);
		bk.run(m + 1, RouteBoundsTask{ key, roff, m, n_dest, dest_off });
		char *routed = bk.template buf<char>(FB_ROUTED, rtotal + 16);
		bk.run(m, RouteCopyTask{ lines, key, val, roff, n_dest, sam, routed });
		out.routed = routed; out.dest_off = dest_off; out.routed_bytes = rtotal;
		t1 = clock_ms(); st.ms_sam_host += t1 - t0; t0 = t1;
	}
	st.n_aln_slots += n_aln;
	out.sam = sam; out.sam_off = sam_off; out.sam_bytes = total; out.lines = lines; out.n_lines = n_lines;
}

} // namespace b200
