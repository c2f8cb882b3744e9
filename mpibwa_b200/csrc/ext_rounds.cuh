// ext_rounds.cuh - seed extension (mem_chain2aln + ksw_extend2, reference src/bwamem.c:632-786, src/ksw.c:380-479)
// as a device-resident state machine plus a batched DP kernel.
//
// Why not one thread per read (v1) and why not one warp per anti-diagonal: BWA's band is re-derived after every row
// from the zero structure of the *completed* row (src/ksw.c:466-469) and cells outside the band keep stale values,
// so a row cannot start before the previous one has finished - a wavefront would have to speculate and roll back.
// The jobs are also small (median query 48 columns, ~45 live cells per row).  What maps well is ONE DP JOB PER LANE:
// every lane runs the exact scalar recurrence on its own job, the H/E row lives in shared memory as packed 16-bit
// pairs laid out [column][lane] (bank = lane, conflict-free for any per-lane column), the query is staged in shared
// memory as bytes, substitution scores come from one PRMT over a 2-register score row, and warps are formed from
// jobs SORTED by (query length, target length) so that lanes stay in step.
//
//   k_ext_advance   one thread per read with work left: finishes the region of the seed whose DP just returned,
//                   walks the read's chains/seeds in the reference's order applying the containment tests
//                   (src/bwamem.c:671-706) until the next seed that needs extending, and emits ONE DP job
//                   (left or right side) for it.  The skip decision needs the regions found so far, hence rounds.
//   k_ext_dp        the batched ksw_extend2 (with the caller's band-doubling retry, src/bwamem.c:723-734,751-762).
//   k_ext_dp_big    same recurrence through extend_core with the row in global memory, for jobs that do not fit
//                   the fast path (very long queries or scores that overflow 15 bits).
#pragma once
#include "ext_kernels.h"

namespace b200 {

struct ExtJob {                 // one pending ksw_extend2 call (+ retry) of a read
	int64_t qaddr;              // index into the code buffer of query base 0
	int64_t f0;                 // forward-strand coordinate of target base 0 (comp 0/1) or offset into a byte buffer (comp 2)
	int32_t qstep, fstep;       // +-1 walking directions
	int32_t comp;               // 0: target from pac; 1: from pac, complemented (reverse strand); 2: target bytes (codes 0-4)
	int32_t qlen, tlen, h0, prev, bonus;
	int32_t w0;                 // 0: mem_chain2aln's rule (band eo.w, one retry with 2*eo.w); > 0: exactly one call with band w0
	int32_t score, qle, tle, gtle, gscore, aw, max_off;   // results
	int32_t pad;
};

struct ExtState {               // per read
	int32_t ci, k;              // cursor: chain index (relative to the read) and position in the chain's score order
	int32_t n_av;               // regions emitted so far
	int32_t phase;              // 0 search, 1 left DP pending, 2 right DP pending, 3 finished
	int32_t aw0, aw1, sc0, seed;
	DReg a;                     // region under construction
};

#define EXT_N_CLASS 8
__device__ __constant__ int c_ext_class_cap[EXT_N_CLASS] = { 32, 64, 96, 128, 160, 256, 704, 0x7fffffff };

__device__ __forceinline__ int ext_class_of(int qlen, int h0, int max_sc)
{
	if ((long long)h0 + (long long)qlen * max_sc >= 32768) return EXT_N_CLASS - 1;
	int c = 0;
	while (qlen > c_ext_class_cap[c]) ++c;
	return c;
}

// ---------------------------------------------------------------- the state machine

__device__ __forceinline__ void ext_emit_job(ExtJob *jb, const uint8_t *, int64_t l_pac, int64_t qaddr, int qstep, int64_t p0, int pstep,
                                             int qlen, int tlen, int h0, int prev, int bonus)
{
	const bool fwd = p0 < l_pac;                  // p0: coordinate of target base 0 in [0, 2*l_pac); windows never bridge strands
	jb->qaddr = qaddr; jb->qstep = qstep;
	jb->f0 = fwd ? p0 : (l_pac << 1) - 1 - p0; jb->fstep = fwd ? pstep : -pstep; jb->comp = fwd ? 0 : 1;
	jb->qlen = qlen; jb->tlen = tlen; jb->h0 = h0; jb->prev = prev; jb->bonus = bonus; jb->w0 = 0;
}

// Advances the chain2aln walk of read r until it needs a DP (returns true; the job is in *jb) or is finished.
__device__ __forceinline__ bool ext_advance_read(const ExtOpt &eo, int64_t l_pac, int r, const int64_t *__restrict__ off,
                                                 const int32_t *__restrict__ chain_off, const DChain *__restrict__ chains,
                                                 const DSeed *__restrict__ seeds, int32_t *srt, ExtState &st, ExtJob *jb, DReg *regs,
                                                 int32_t *n_regs)
{
	const int c0 = chain_off[r], nc = chain_off[r + 1] - c0;
	const int l_query = (int)(off[r + 1] - off[r]);
	const int64_t qbase = off[r];
	DReg *out = regs + chains[c0].seed_beg;
	bool emitted = false;
	for (;;) {
		const DChain &c = chains[c0 + st.ci];
		const DSeed *cs = seeds + c.seed_beg;
		if (st.phase == 1) {                          // left extension returned (reference src/bwamem.c:723-743)
			const DSeed &s = cs[st.seed];
			st.a.score = jb->score; st.aw0 = jb->aw;
			if (jb->gscore <= 0 || jb->gscore <= st.a.score - eo.pen_clip5) {
				st.a.qb = s.qbeg - jb->qle; st.a.rb = s.rbeg - jb->tle; st.a.truesc = st.a.score;
			} else { st.a.qb = 0; st.a.rb = s.rbeg - jb->gtle; st.a.truesc = jb->gscore; }
			st.phase = 4;                             // -> right side
		} else if (st.phase == 2) {                   // right extension returned (src/bwamem.c:751-771)
			const DSeed &s = cs[st.seed];
			const int qe = s.qbeg + s.len;
			const int64_t re = s.rbeg + s.len;
			st.a.score = jb->score; st.aw1 = jb->aw;
			if (jb->gscore <= 0 || jb->gscore <= st.a.score - eo.pen_clip3) {
				st.a.qe = qe + jb->qle; st.a.re = re + jb->tle; st.a.truesc += st.a.score - st.sc0;
			} else { st.a.qe = l_query; st.a.re = re + jb->gtle; st.a.truesc += jb->gscore - st.sc0; }
			st.phase = 5;                             // -> finish the region
		} else if (st.phase == 0) {                   // look for the next seed that needs extending
			bool found = false;
			while (st.ci < nc) {
				const DChain &cc = chains[c0 + st.ci];
				if (st.k < 0) { ++st.ci; if (st.ci < nc) st.k = chains[c0 + st.ci].n_seeds - 1; continue; }
				if (cc.n_seeds > 0 && chain2aln_need_extension(eo, l_query, cc, seeds + cc.seed_beg, srt + cc.seed_beg, st.k, out, st.n_av)) { found = true; break; }
				--st.k;
			}
			if (!found) { st.phase = 3; n_regs[r] = st.n_av; break; }
			const DChain &cc = chains[c0 + st.ci];
			st.seed = srt[cc.seed_beg + st.k];
			const DSeed &s = seeds[cc.seed_beg + st.seed];
			st.aw0 = st.aw1 = eo.w;
			st.a.w = eo.w; st.a.score = st.a.truesc = -1; st.a.rid = cc.rid;
			st.a.qb = st.a.qe = 0; st.a.rb = st.a.re = 0; st.a.seedcov = 0; st.a.seedlen0 = 0; st.a.pad = 0; st.a.frac_rep = 0;
			if (s.qbeg) {
				ext_emit_job(jb, nullptr, l_pac, qbase + s.qbeg - 1, -1, s.rbeg - 1, -1, s.qbeg, (int)(s.rbeg - cc.rmax0), s.len * eo.a, -1, eo.pen_clip5);
				st.phase = 1; emitted = true; break;
			}
			st.a.score = st.a.truesc = s.len * eo.a; st.a.qb = 0; st.a.rb = s.rbeg;
			st.phase = 4;
		} else if (st.phase == 4) {                   // right side of the current seed
			const DSeed &s = cs[st.seed];
			if (s.qbeg + s.len != l_query) {
				const int qe = s.qbeg + s.len;
				const int64_t re = s.rbeg + s.len;
				st.sc0 = st.a.score;
				ext_emit_job(jb, nullptr, l_pac, qbase + qe, 1, re, 1, l_query - qe, (int)(c.rmax1 - re), st.sc0, st.a.score, eo.pen_clip3);
				st.phase = 2; emitted = true; break;
			}
			st.a.qe = l_query; st.a.re = s.rbeg + s.len;
			st.phase = 5;
		} else {                                      // phase 5: seedcov, band, bookkeeping (src/bwamem.c:774-783)
			const DSeed &s = cs[st.seed];
			int cov = 0;
			for (int i = 0; i < c.n_seeds; ++i) {
				const DSeed &u = cs[i];
				if (u.qbeg >= st.a.qb && u.qbeg + u.len <= st.a.qe && u.rbeg >= st.a.rb && u.rbeg + u.len <= st.a.re) cov += u.len;
			}
			st.a.seedcov = cov;
			st.a.w = st.aw0 > st.aw1 ? st.aw0 : st.aw1;
			st.a.seedlen0 = s.len;
			st.a.frac_rep = c.frac_rep;
			out[st.n_av++] = st.a;
			--st.k;
			st.phase = 0;
		}
	}
	return emitted;
}

__global__ void __launch_bounds__(128) k_ext_advance(ExtOpt eo, int64_t l_pac, int n_active, const int32_t *__restrict__ active,
                                                     const int64_t *__restrict__ off, const int32_t *__restrict__ chain_off,
                                                     const DChain *__restrict__ chains, const DSeed *__restrict__ seeds, int32_t *srt,
                                                     ExtState *state, ExtJob *jobs, DReg *regs, int32_t *n_regs,
                                                     int32_t *next_active, uint32_t *next_key, int32_t *counters /* [0]=n_next, [1..]=class hist */)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_active) return;
	const int r = active[t];
	ExtState st = state[r];
	ExtJob *jb = &jobs[r];
	const bool emitted = ext_advance_read(eo, l_pac, r, off, chain_off, chains, seeds, srt, st, jb, regs, n_regs);
	state[r] = st;
	if (emitted) {
		int pos = atomicAdd(&counters[0], 1);
		int cls = ext_class_of(jb->qlen, jb->h0, eo.max_sc);
		atomicAdd(&counters[1 + cls], 1);
		next_active[pos] = r;
		int tl = jb->tlen > 0xffff ? 0xffff : jb->tlen;
		int ql = jb->qlen > 0x3fff ? 0x3fff : jb->qlen;
		next_key[pos] = ((uint32_t)cls << 28) | ((uint32_t)(ql & 0xfff) << 16) | (uint32_t)tl;
	}
}

__global__ void k_ext_init(int n_reads, const int32_t *__restrict__ chain_off, const DChain *__restrict__ chains, ExtState *state,
                           int32_t *n_regs, int32_t *active, int32_t *counters)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int c0 = chain_off[r], nc = chain_off[r + 1] - c0;
	n_regs[r] = 0;
	if (nc <= 0) return;
	ExtState st;
	st.ci = 0; st.k = chains[c0].n_seeds - 1; st.n_av = 0; st.phase = 0; st.aw0 = st.aw1 = st.sc0 = st.seed = 0;
	st.a = DReg();
	state[r] = st;
	active[atomicAdd(&counters[0], 1)] = r;
}

// gather the regions of every read into one compact array (reg_off = exclusive scan of n_regs)
__global__ void k_ext_gather(int n_reads, const int32_t *__restrict__ chain_off, const DChain *__restrict__ chains,
                             const int32_t *__restrict__ n_regs, const int64_t *__restrict__ reg_off, const DReg *__restrict__ regs, DReg *out)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int n = n_regs[r];
	if (n == 0) return;
	const DReg *src = regs + chains[chain_off[r]].seed_beg;
	DReg *dst = out + reg_off[r];
	for (int i = 0; i < n; ++i) dst[i] = src[i];
}

// ---------------------------------------------------------------- the DP

__device__ __forceinline__ int prmt_score(uint32_t lo, uint32_t hi, uint32_t sel)
{
	int s;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(s) : "r"(lo), "r"(hi), "r"(sel));
	return s;
}

__device__ __forceinline__ int pac_fbase(const uint8_t *__restrict__ pac, int64_t f) { return pac[f >> 2] >> ((~f & 3) << 1) & 3; }

// target base of a job: 2-bit reference (forward / complemented) or a byte buffer of codes 0-4
__device__ __forceinline__ int ext_tbase(const uint8_t *__restrict__ src, int64_t f, int comp)
{
	if (comp == 2) return src[f];
	const int b = pac_fbase(src, f);
	return comp ? 3 - b : b;
}

// The target of a job is a window of the 2-bit reference at a random place (775 MB for a human-sized reference: a DRAM miss per
// 128-byte line), read one base per DP row.  Asking for its lines up front turns the serialised misses of the row loop - where the
// whole warp waits for the lane that crossed into a new line - into one overlapped batch behind the query staging.
// `first`, `step`: which of the window's lines this thread asks for (all of them: 0, 1; one warp per job: lane, 32).
__device__ __forceinline__ void ext_prefetch_target(const uint8_t *__restrict__ pac, int64_t f0, int fstep, int tlen, int comp, int first, int step)
{
	if (comp == 2 || tlen <= 0) return;
	const int64_t fe = f0 + (int64_t)fstep * (tlen - 1);
	const int64_t lo = (f0 < fe ? f0 : fe) >> 2, hi = (f0 < fe ? fe : f0) >> 2;
	for (int64_t a = (lo & ~(int64_t)127) + (int64_t)first * 128; a <= hi; a += (int64_t)step * 128)
		asm volatile("prefetch.global.L1 [%0];" :: "l"(pac + a));
}

// Query bytes live in shared memory as [column/4][lane][column%4] (bank = lane for every lane/column combination).
__device__ __forceinline__ int ext_qidx(int j) { return ((j >> 2) << 7) + (j & 3); }

// row -1 of ksw_extend2 (src/ksw.c:395-397) and the band clamp (src/ksw.c:399-407)
__device__ __forceinline__ int ext_init_row(const ExtOpt &o, const ExtJob &jb, int w, uint32_t *S)
{
	const int oe_ins = o.o_ins + o.e_ins;
	int v = jb.h0 > oe_ins ? jb.h0 - oe_ins : 0;
	S[0] = (uint32_t)jb.h0;
	for (int j = 1; j <= jb.qlen; ++j) { S[j * 32] = (uint32_t)v; v = v > o.e_ins ? v - o.e_ins : 0; }
	int max_ins = (int)((double)(jb.qlen * o.max_sc + jb.bonus - o.o_ins) / o.e_ins + 1.);
	max_ins = max_ins > 1 ? max_ins : 1;
	w = w < max_ins ? w : max_ins;
	int max_del = (int)((double)(jb.qlen * o.max_sc + jb.bonus - o.o_del) / o.e_del + 1.);
	max_del = max_del > 1 ? max_del : 1;
	return w < max_del ? w : max_del;
}

__device__ __forceinline__ void ext_fill_score_rows(const ExtOpt &eo, uint32_t *sc_lo, uint32_t *sc_hi)
{
	if (threadIdx.x < 5) {
		const int8_t *m = eo.mat + threadIdx.x * 5;
		sc_lo[threadIdx.x] = (uint32_t)(uint8_t)m[0] | (uint32_t)(uint8_t)m[1] << 8 | (uint32_t)(uint8_t)m[2] << 16 | (uint32_t)(uint8_t)m[3] << 24;
		sc_hi[threadIdx.x] = (uint32_t)(uint8_t)m[4];
	}
}

// Fast path: one job per lane.  The row loop is warp-synchronous: every iteration each live lane computes one row of its own
// job, then the warp reconverges, so that the per-row prologue/epilogue is issued once per warp-row and only the cell loop runs
// with per-lane trip counts.  Lanes are PERSISTENT: a job ends where z-drop, an all-zero row or the target's end stop it
// (src/ksw.c:455-465) - a spurious 19-mer seed of a human-sized index dies after a handful of rows, a true flank runs all of them -
// so no sort key keeps 32 lanes in step; a lane whose job is over takes the next one of its size class from a counter (the order
// is largest first), staged and initialised while the others wait, once `refill` lanes of the warp are idle (16: measured best of 1 / 8 / 16 /
// 32; the chunk's job list as one batch 6.35 -> 5.53 ms against lanes that keep their first job only).
__global__ void __launch_bounds__(64) k_ext_dp(ExtOpt eo, const uint8_t *__restrict__ pac, const uint8_t *__restrict__ codes,
                                               ExtJob *jobs, const int32_t *__restrict__ order, int n, int qcap, unsigned long long *cells_out,
                                               unsigned long long *calls_out, int *next, int refill)
{
	extern __shared__ uint32_t smem[];
	__shared__ uint32_t sc_lo[5], sc_hi[5];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int qpad = (qcap + 4) & ~3;
	const int per_warp_words = (qcap + 1) * 32 + qpad * 8;
	uint32_t *S = smem + (size_t)wib * per_warp_words + lane;
	uint8_t *Q = (uint8_t *)(smem + (size_t)wib * per_warp_words + (qcap + 1) * 32) + lane * 4;
	ext_fill_score_rows(eo, sc_lo, sc_hi);
	__syncthreads();
	const int e_del = eo.e_del, e_ins = eo.e_ins, oe_del = eo.o_del + eo.e_del, oe_ins = eo.o_ins + eo.e_ins;
	long long cells = 0;
	int calls = 0;
	ExtJob *jp = nullptr;
	ExtJob jb;
	jb.qlen = 0; jb.tlen = 0; jb.h0 = 1; jb.prev = -1; jb.bonus = 0; jb.f0 = 0; jb.fstep = 0; jb.comp = 0; jb.qaddr = 0; jb.qstep = 0; jb.w0 = 0;
	bool alive = false, drained = false;
	// per-attempt DP state
	int aw = eo.w, attempt = 0, prev_score = -1, w = 0;
	int i = 0, beg = 0, end = 0, max = 0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
	int64_t f = 0;
	for (;;) {
		const unsigned idle = __ballot_sync(0xffffffffu, !alive && !drained);
		if (idle && (__popc(idle) >= refill || !__any_sync(0xffffffffu, alive))) {
			if (!alive && !drained) {
				const int t = atomicAdd(next, 1);
				if (t < n) {
					jp = &jobs[order[t]]; jb = *jp; alive = true;
					ext_prefetch_target(pac, jb.f0, jb.fstep, jb.tlen, jb.comp, 0, 1);
					for (int j = 0; j < jb.qlen; ++j) Q[ext_qidx(j)] = codes[jb.qaddr + (int64_t)jb.qstep * j];
					aw = jb.w0 > 0 ? jb.w0 : eo.w; attempt = jb.w0 > 0 ? 1 : 0; prev_score = jb.prev;
					w = ext_init_row(eo, jb, aw, S);
					i = 0; beg = 0; end = jb.qlen; max = jb.h0; max_i = -1; max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
					f = jb.f0;
				} else drained = true;
			}
			__syncwarp();
		}
		if (!__any_sync(0xffffffffu, alive)) break;
		if (alive) {
			bool done = i >= jb.tlen;
			if (!done) {
				const int tb = ext_tbase(pac, f, jb.comp);
				const uint32_t lo = sc_lo[tb], hi = sc_hi[tb];
				int fgap = 0, h1, hj = -1, j;
				if (beg < i - w) beg = i - w;
				if (end > i + w + 1) end = i + w + 1;
				if (end > jb.qlen) end = jb.qlen;
				if (beg == 0) { h1 = jb.h0 - (eo.o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
				else h1 = 0;
#define B200_EXT_CELL(J_, Q_) do { \
					const uint32_t wd = S[(J_) * 32]; \
					const int sc = prmt_score(lo, hi, (uint32_t)(Q_) * 0x1111u + 0x8880u); \
					const int hd = (int)(wd & 0xffffu); \
					int e = (int)(wd >> 16); \
					int M = hd + sc; \
					M = hd ? M : 0; \
					const int h = ::max(::max(M, e), fgap); \
					hj = ::max(hj, h * 65536 + (J_)); \
					e = ::max(::max(e - e_del, M - oe_del), 0); \
					fgap = ::max(::max(fgap - e_ins, M - oe_ins), 0); \
					S[(J_) * 32] = (uint32_t)(e * 65536 + h1); \
					h1 = h; } while (0)
				// columns up to the next multiple of four, whole groups of four (one 32-bit load brings their query codes), rest
				for (j = beg; j < end && (j & 3); ++j) B200_EXT_CELL(j, Q[ext_qidx(j)]);
				for (; j + 4 <= end; j += 4) {
					const uint32_t qw = *reinterpret_cast<const uint32_t *>(Q + ((j >> 2) << 7));
					B200_EXT_CELL(j, qw & 0xffu);
					B200_EXT_CELL(j + 1, (qw >> 8) & 0xffu);
					B200_EXT_CELL(j + 2, (qw >> 16) & 0xffu);
					B200_EXT_CELL(j + 3, qw >> 24);
				}
				for (; j < end; ++j) B200_EXT_CELL(j, Q[ext_qidx(j)]);
#undef B200_EXT_CELL
				if (end > beg) cells += end - beg;
				S[end * 32] = (uint32_t)h1;
				if (j == jb.qlen) {
					max_ie = gscore > h1 ? max_ie : i;
					gscore = gscore > h1 ? gscore : h1;
				}
				const int m = hj < 0 ? 0 : hj >> 16, mj = hj < 0 ? -1 : hj & 0xffff;
				if (m == 0) done = true;
				else {
					if (m > max) {
						max = m; max_i = i; max_j = mj;
						int d = mj - i; d = d < 0 ? -d : d;
						max_off = max_off > d ? max_off : d;
					} else if (eo.zdrop > 0) {
						if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > eo.zdrop) done = true; }
						else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > eo.zdrop) done = true; }
					}
					if (!done) {
						for (j = beg; j < end && S[j * 32] == 0; ++j) {}
						beg = j;
						for (j = end; j >= beg && S[j * 32] == 0; --j) {}
						end = j + 2 < jb.qlen ? j + 2 : jb.qlen;
						++i; f += jb.fstep;
						if (i >= jb.tlen) done = true;
					}
				}
			}
			if (done) {                                     // this ksw_extend2 call is over (src/bwamem.c:723-734,751-762)
				++calls;
				const int score = max;
				if (attempt == 0 && !(score == prev_score || max_off < (aw >> 1) + (aw >> 2))) {
					attempt = 1; prev_score = score; aw = eo.w << 1;
					w = ext_init_row(eo, jb, aw, S);
					i = 0; beg = 0; end = jb.qlen; max = jb.h0; max_i = max_j = max_ie = -1; gscore = -1; max_off = 0; f = jb.f0;
				} else {
					jp->score = score; jp->qle = max_j + 1; jp->tle = max_i + 1; jp->gtle = max_ie + 1; jp->gscore = gscore; jp->aw = aw;
					jp->max_off = max_off;
					alive = false;
				}
			}
		}
		__syncwarp();
	}
	for (int o = 16; o > 0; o >>= 1) { cells += __shfl_down_sync(0xffffffffu, cells, o); calls += __shfl_down_sync(0xffffffffu, calls, o); }
	if (lane == 0) { if (cells) atomicAdd(cells_out, (unsigned long long)cells); if (calls) atomicAdd(calls_out, (unsigned long long)calls); }
}

// general path: any query length / score range, row state in global memory (int32, lane-interleaved)
struct QStep { const uint8_t *p; int64_t step; __device__ __forceinline__ int operator()(int j) const { return p[step * j]; } };
struct TStep { const uint8_t *pac; int64_t f0; int fstep, comp; __device__ __forceinline__ int operator()(int i) const { return ext_tbase(pac, f0 + (int64_t)fstep * i, comp); } };

__global__ void __launch_bounds__(128) k_ext_dp_big(ExtOpt eo, const uint8_t *__restrict__ pac, const uint8_t *__restrict__ codes,
                                                    ExtJob *jobs, const int32_t *__restrict__ order, int n, int32_t *eh, int64_t stride,
                                                    unsigned long long *cells_out, unsigned long long *calls_out)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	int64_t cells = 0;
	int calls = 0;
	if (t < n) {
		ExtJob *jp = &jobs[order[t]];
		const ExtJob jb = *jp;
		EhStrided acc = { eh + t, stride };
		QStep qa = { codes + jb.qaddr, jb.qstep };
		TStep ta = { pac, jb.f0, jb.fstep, jb.comp };
		ExtOut x; x.score = -1; x.qle = x.tle = x.gtle = 0; x.gscore = -1; x.max_off = 0;
		int score = jb.prev, aw = eo.w;
		for (int it = 0; it < 2; ++it) {
			const int prev = score;
			aw = jb.w0 > 0 ? jb.w0 : eo.w << it;
			extend_core(jb.qlen, qa, jb.tlen, ta, eo, aw, jb.bonus, jb.h0, acc, &x, &cells);
			++calls;
			score = x.score;
			if (jb.w0 > 0 || score == prev || x.max_off < (aw >> 1) + (aw >> 2)) break;
		}
		jp->score = score; jp->qle = x.qle; jp->tle = x.tle; jp->gtle = x.gtle; jp->gscore = x.gscore; jp->aw = aw; jp->max_off = x.max_off;
	}
	long long c = cells;
	for (int o = 16; o > 0; o >>= 1) { c += __shfl_down_sync(0xffffffffu, c, o); calls += __shfl_down_sync(0xffffffffu, calls, o); }
	if ((threadIdx.x & 31) == 0) { if (c) atomicAdd(cells_out, (unsigned long long)c); if (calls) atomicAdd(calls_out, (unsigned long long)calls); }
}

// ---------------------------------------------------------------- warp-cooperative DP (latency path)
//
// One WARP per ksw_extend2 call: the 32 lanes take 32 adjacent columns of the band at a time.  Within a row the only
// left-to-right dependency is F, and because BWA opens gaps from M (src/ksw.c:444-446: t = M - oe_ins) F is a pure
// max-plus prefix over g_j = max(M_j - oe_ins, 0):  F_j = max_{k<j} (g_k - (j-1-k) e_ins)  - one 5-step shuffle scan of
// a_j = g_j + j e_ins per 32 columns.  H(i,j-1) reaches column j by one shuffle, the row maximum (ties to the larger j)
// by one max-reduction of (h << 32 | j), and the band shrink (first / last non-zero entry of the finished row,
// src/ksw.c:466-469) by ballots.  H and E rows live in shared memory as int32 (no 15-bit limit).  Per row this costs
// ~100-150 issue slots per warp instead of ~30 per CELL per lane in k_ext_dp, i.e. ~8x less latency per job at ~3x the
// instruction count: it is used where a round has too few jobs to fill the chip (the long-query classes and the tail of
// the round sequence), where the latency of the longest job, not throughput, sets the time.
struct ExtWarpScratch { int32_t *H, *E; uint8_t *Q; };

__device__ __forceinline__ ExtWarpScratch ext_warp_scratch(uint32_t *smem, int wib, int qcap)
{
	const int words = 2 * (qcap + 2) + ((qcap + 4) >> 2);
	ExtWarpScratch s;
	s.H = (int32_t *)(smem + (size_t)wib * words);
	s.E = s.H + (qcap + 2);
	s.Q = (uint8_t *)(s.E + (qcap + 2));
	return s;
}
static inline size_t ext_warp_smem_bytes(int warps, int qcap) { return (size_t)warps * (2 * (qcap + 2) + ((qcap + 4) >> 2)) * 4; }

// runs the job (both band attempts of the caller, src/bwamem.c:723-734,751-762) and leaves the results in jb
__device__ void ext_dp_warp(const ExtOpt &eo, const uint32_t *__restrict__ sc_lo, const uint32_t *__restrict__ sc_hi,
                            const uint8_t *__restrict__ pac, const uint8_t *__restrict__ codes, ExtJob &jb, const ExtWarpScratch &S,
                            long long &cells, int &calls)
{
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int qlen = jb.qlen, tlen = jb.tlen, h0 = jb.h0;
	const int e_del = eo.e_del, e_ins = eo.e_ins, oe_del = eo.o_del + eo.e_del, oe_ins = eo.o_ins + eo.e_ins;
	ext_prefetch_target(pac, jb.f0, jb.fstep, tlen, jb.comp, lane, 32);
	for (int j = lane; j < qlen; j += 32) S.Q[j] = codes[jb.qaddr + (int64_t)jb.qstep * j];
	const bool small = (long long)h0 + (long long)qlen * eo.max_sc < 32768 && qlen < 65536;
	int prev_score = jb.prev, aw = eo.w, score = 0;
	int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
	for (int attempt = 0; attempt < 2; ++attempt) {
		aw = jb.w0 > 0 ? jb.w0 : eo.w << attempt;
		__syncwarp();
		for (int j = lane; j <= qlen; j += 32) {           // row -1 (src/ksw.c:395-397)
			int v = j == 0 ? h0 : h0 - oe_ins - (j - 1) * e_ins;
			S.H[j] = v > 0 ? v : 0; S.E[j] = 0;
		}
		int w = aw;                                        // band clamp (src/ksw.c:399-407)
		{
			int max_ins = (int)((double)(qlen * eo.max_sc + jb.bonus - eo.o_ins) / eo.e_ins + 1.);
			max_ins = max_ins > 1 ? max_ins : 1;
			w = w < max_ins ? w : max_ins;
			int max_del = (int)((double)(qlen * eo.max_sc + jb.bonus - eo.o_del) / eo.e_del + 1.);
			max_del = max_del > 1 ? max_del : 1;
			w = w < max_del ? w : max_del;
		}
		__syncwarp();
		max = h0; max_i = -1; max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
		int beg = 0, end = qlen;
		int64_t f = jb.f0;
		int tb_next = tlen > 0 ? ext_tbase(pac, f, jb.comp) : 0;
		for (int i = 0; i < tlen; ++i) {
			const int tb = tb_next;
			f += jb.fstep;
			if (i + 1 < tlen) tb_next = ext_tbase(pac, f, jb.comp);
			const uint32_t lo = sc_lo[tb], hi = sc_hi[tb];
			if (beg < i - w) beg = i - w;
			if (end > i + w + 1) end = i + w + 1;
			if (end > qlen) end = qlen;
			int h1 = 0;
			if (beg == 0) { h1 = h0 - (eo.o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
			int fcar = 0, hcar = h1;                       // F entering / H left of the first column of the chunk
			long long best = -1;
			int best32 = -1;
			int first_nz = 0x7fffffff, last_nz = -1, hlast = h1;
			for (int cb = beg; cb < end; cb += 32) {
				const int j = cb + lane;
				const bool act = j < end;
				const int hd = act ? S.H[j] : 0;
				int e = act ? S.E[j] : 0;
				const int q = act ? S.Q[j] : 4;
				int M = hd + prmt_score(lo, hi, (uint32_t)q * 0x1111u + 0x8880u);
				M = hd ? M : 0;
				// F by max-plus scan; the carry enters as a virtual column cb-1
				int a = act ? ::max(M - oe_ins, 0) + j * e_ins : (int)0x80000000;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, a, o); if (lane >= o) a = ::max(a, v); }
				const int acar = fcar + (cb - 1) * e_ins;
				int x = __shfl_up_sync(FULL, a, 1);
				x = lane == 0 ? acar : ::max(x, acar);
				const int fj = x - (j - 1) * e_ins;
				const int h = ::max(::max(M, e), fj);
				e = ::max(::max(e - e_del, M - oe_del), 0);
				int hl = __shfl_up_sync(FULL, h, 1);
				if (lane == 0) hl = hcar;
				if (act) {
					S.H[j] = hl; S.E[j] = e;
					if (small) { const int key = h * 65536 + j; best32 = ::max(best32, key); }
					else { const long long key = (long long)h << 32 | (unsigned)j; best = key > best ? key : best; }
				}
				const unsigned nz = __ballot_sync(FULL, act && (hl | e) != 0);
				if (nz) {
					if (first_nz == 0x7fffffff) first_nz = cb + __ffs(nz) - 1;
					last_nz = cb + 31 - __clz(nz);
				}
				const int nlast = ::min(end - cb, 32) - 1;    // lane of the last active column of this chunk
				hlast = __shfl_sync(FULL, h, nlast);
				hcar = hlast;
				fcar = ::max(__shfl_sync(FULL, a, 31), acar) - (cb + 31) * e_ins;   // only used when the chunk is full
			}
			if (end > beg) cells += lane == 0 ? end - beg : 0;
			if (lane == 0) { S.H[end] = hlast; S.E[end] = 0; }
			if (hlast != 0) last_nz = end;
			if ((beg < end ? end : beg) == qlen) {
				max_ie = gscore > hlast ? max_ie : i;
				gscore = gscore > hlast ? gscore : hlast;
			}
			int m, mj;
			if (small) {                                   // scores below 2^15: (h << 16 | j) fits one register
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) best32 = ::max(best32, __shfl_xor_sync(FULL, best32, o));
				m = best32 < 0 ? 0 : best32 >> 16; mj = best32 < 0 ? -1 : best32 & 0xffff;
			} else {
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) { const long long v = __shfl_xor_sync(FULL, best, o); best = v > best ? v : best; }
				m = best < 0 ? 0 : (int)(best >> 32); mj = best < 0 ? -1 : (int)(uint32_t)best;
			}
			if (m == 0) break;
			if (m > max) {
				max = m; max_i = i; max_j = mj;
				int d = mj - i; d = d < 0 ? -d : d;
				max_off = max_off > d ? max_off : d;
			} else if (eo.zdrop > 0) {
				if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > eo.zdrop) break; }
				else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > eo.zdrop) break; }
			}
			// band for the next row: first / last non-zero entry of the finished row (src/ksw.c:466-469)
			const int nbeg = first_nz != 0x7fffffff ? first_nz : end;
			const int jl = last_nz >= nbeg ? last_nz : nbeg - 1;
			beg = nbeg;
			end = jl + 2 < qlen ? jl + 2 : qlen;
			__syncwarp();
		}
		++calls;
		score = max;
		if (jb.w0 == 0 && attempt == 0 && !(score == prev_score || max_off < (aw >> 1) + (aw >> 2))) { prev_score = score; continue; }
		break;
	}
	jb.score = score; jb.qle = max_j + 1; jb.tle = max_i + 1; jb.gtle = max_ie + 1; jb.gscore = gscore; jb.aw = aw; jb.max_off = max_off;
}

// one warp per job of a round
__global__ void __launch_bounds__(128) k_ext_dp_warp(ExtOpt eo, const uint8_t *__restrict__ pac, const uint8_t *__restrict__ codes,
                                                     ExtJob *jobs, const int32_t *__restrict__ order, int n, int qcap,
                                                     unsigned long long *cells_out, unsigned long long *calls_out)
{
	extern __shared__ uint32_t smem[];
	__shared__ uint32_t sc_lo[5], sc_hi[5];
	ext_fill_score_rows(eo, sc_lo, sc_hi);
	__syncthreads();
	const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wid = blockIdx.x * (blockDim.x >> 5) + wib;
	if (wid >= n) return;
	ExtJob *jp = &jobs[order[wid]];
	ExtJob jb = *jp;
	long long cells = 0;
	int calls = 0;
	ext_dp_warp(eo, sc_lo, sc_hi, pac, codes, jb, ext_warp_scratch(smem, wib, qcap), cells, calls);
	if (lane == 0) {
		jp->score = jb.score; jp->qle = jb.qle; jp->tle = jb.tle; jp->gtle = jb.gtle; jp->gscore = jb.gscore; jp->aw = jb.aw;
		jp->max_off = jb.max_off;
		if (cells) atomicAdd(cells_out, (unsigned long long)cells);
		atomicAdd(calls_out, (unsigned long long)calls);
	}
}

// Tail of the round sequence: one warp per read that still has work walks the rest of its chains on its own -
// lane 0 runs the chain2aln state machine, the warp runs every DP it asks for - so the few reads with many chains
// no longer cost a launch + sort + host round trip per extension.
__global__ void __launch_bounds__(128) k_ext_tail(ExtOpt eo, int64_t l_pac, const uint8_t *__restrict__ pac, const uint8_t *__restrict__ codes,
                                                  int n_active, const int32_t *__restrict__ active, const int64_t *__restrict__ off,
                                                  const int32_t *__restrict__ chain_off, const DChain *__restrict__ chains,
                                                  const DSeed *__restrict__ seeds, int32_t *srt, ExtState *state, ExtJob *jobs, DReg *regs,
                                                  int32_t *n_regs, int qcap, unsigned long long *cells_out, unsigned long long *calls_out)
{
	extern __shared__ uint32_t smem[];
	__shared__ uint32_t sc_lo[5], sc_hi[5];
	ext_fill_score_rows(eo, sc_lo, sc_hi);
	__syncthreads();
	const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wid = blockIdx.x * (blockDim.x >> 5) + wib;
	if (wid >= n_active) return;
	const int r = active[wid];
	const ExtWarpScratch S = ext_warp_scratch(smem, wib, qcap);
	ExtJob *jp = &jobs[r];
	long long cells = 0;
	int calls = 0;
	ExtState st;
	if (lane == 0) st = state[r];
	for (;;) {
		int emitted = 0;
		if (lane == 0) emitted = ext_advance_read(eo, l_pac, r, off, chain_off, chains, seeds, srt, st, jp, regs, n_regs) ? 1 : 0;
		emitted = __shfl_sync(0xffffffffu, emitted, 0);
		if (!emitted) break;
		__syncwarp();
		ExtJob jb = *jp;                                   // written by lane 0 just above
		ext_dp_warp(eo, sc_lo, sc_hi, pac, codes, jb, S, cells, calls);
		if (lane == 0) { jp->score = jb.score; jp->qle = jb.qle; jp->tle = jb.tle; jp->gtle = jb.gtle; jp->gscore = jb.gscore; jp->aw = jb.aw; }
		__syncwarp();
	}
	if (lane == 0) {
		state[r] = st;
		if (cells) atomicAdd(cells_out, (unsigned long long)cells);
		if (calls) atomicAdd(calls_out, (unsigned long long)calls);
	}
}

} // namespace b200
