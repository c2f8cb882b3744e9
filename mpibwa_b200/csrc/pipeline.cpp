// pipeline.cpp - mem_process_seqs: the reference's per-read thread loops (src/bwamem.c:1205-1234) re-organised
// into batched stages so that the FM-index search, the seed extension and the mate-rescue Smith-Waterman of a
// whole chunk run as device kernels, with the pointer-chasing steps in between on host threads:
//
//   encode -> [GPU] seeding + SA look-up -> chaining/filtering -> [GPU] chain2aln (ksw_extend2)
//          -> dedup/patch -> insert-size statistics (chunk-global) -> [GPU] mate-rescue ksw_align2 -> replay
//          -> primary marking, pairing, mapQ, CIGAR, SAM text
//
// Results do not depend on opt->n_threads or on batching: the extension of a seed does not depend
// on the regions found so far (only the decision to extend does, and that decision is taken on the device in the
// reference's order), and a rescue alignment depends only on (anchor, orientation, mate), so it can be computed
// up front and the sequential insert/skip logic of mem_matesw replayed afterwards.
#include "host_align.h"
#include "stages.h"
#include "util.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <unordered_map>
#include <algorithm>
#include <atomic>
#include <thread>
#include <condition_variable>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <set>
#include <sys/mman.h>

namespace b200 {

static double now_ms()
{
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

/* ------------------------------------------------------------------ engine registry */

static std::mutex g_mu;
static std::vector<Engine *> g_engines;     // [0] owns the index in HBM; the others are clones that share it (one per sub-batch lane)
static const void *g_engine_key = nullptr;
static int g_device = -1, g_engine_dev = -1;     // requested device (-1: from the launcher's local rank), device of g_engines
// One lane at a time drives the GPU (the host stages of the other lanes and jobs overlap it).  The lock is handed to the waiter
// of the OLDEST chunk job: a chunk that has reached its last device stage does not queue behind the seeding of three
// younger chunks, which keeps the latency of a job - and the ramp of a short run - down.
class DeviceTurn {
public:
	void lock(uint64_t ticket)
	{
		std::unique_lock<std::mutex> lk(mu_);
		auto it = waiting_.insert(ticket);
		cv_.wait(lk, [&] { return !held_ && *waiting_.begin() == ticket; });
		waiting_.erase(it);
		held_ = true;
	}
	void unlock()
	{
		{ std::lock_guard<std::mutex> lk(mu_); held_ = false; }
		cv_.notify_all();
	}
private:
	std::mutex mu_;
	std::condition_variable cv_;
	std::multiset<uint64_t> waiting_;
	bool held_ = false;
};
static DeviceTurn g_gpu_turn;
struct DeviceTurnGuard { uint64_t t; explicit DeviceTurnGuard(uint64_t t_) : t(t_) { g_gpu_turn.lock(t); } ~DeviceTurnGuard() { g_gpu_turn.unlock(); } };

// engine of the single-job C wrappers (ksw_extend2, bwt_sa, ... in capi.cpp): a clone of its own, so that those calls never
// touch the resident reads, scratch buffers or counters of a chunk job; serialised by g_aux_mu and the device turn (AuxGuard)
static Engine *g_aux = nullptr;
static std::mutex g_aux_mu;
static int g_running = 0;                    // chunk jobs past their turn gate (guarded by g_slot_mu)
static std::mutex g_slot_mu;

static void destroy_engines_locked()
{
	if (g_aux) { engine_destroy(g_aux); g_aux = nullptr; }
	for (size_t k = g_engines.size(); k-- > 0;) engine_destroy(g_engines[k]);
	g_engines.clear(); g_engine_key = nullptr;
}

// local rank of this process: torchrun, Open MPI, MVAPICH2, Intel MPI / MPICH (hydra), Slurm - the reference hosts are
// launched with mpirun or srun, not torchrun
static int local_rank_from_env()
{
	static const char *names[] = { "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID" };
	for (const char *n : names) { const char *v = getenv(n); if (v && *v) return atoi(v); }
	return 0;
}

Engine *engine_for(const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac)
{
	std::lock_guard<std::mutex> lk(g_mu);
	int dev = g_device;
	if (dev < 0) {
		dev = local_rank_from_env();
		int nd = engine_device_count();
		if (nd > 0) dev %= nd;
	}
	if (!g_engines.empty() && g_engine_key == (const void *)bwt->bwt && g_engine_dev == dev) return g_engines[0];
	if (!g_engines.empty()) {
		std::lock_guard<std::mutex> sl(g_slot_mu);
		if (g_running > 0) { fprintf(stderr, "[mpibwa_b200] the index or device was changed while chunk jobs are in flight\n"); abort(); }
	}
	destroy_engines_locked();
	g_engines.push_back(engine_create(bwt, bns, pac, dev));
	g_engine_key = (const void *)bwt->bwt;
	g_engine_dev = dev;
	return g_engines[0];
}

// engine of lane k of chunk slot `slot` (slot 0, lane 0 is the primary engine; all others are clones sharing its index)
static const int MAX_LANES = 4, N_SLOTS = 4;
static Engine *engine_lane(int slot, int k)
{
	std::lock_guard<std::mutex> lk(g_mu);
	const int idx = slot * MAX_LANES + k;
	while ((int)g_engines.size() <= idx) g_engines.push_back(engine_clone(g_engines[0]));
	return g_engines[idx];
}

// Chunk slots: a slot is a set of lane engines that one mem_process_seqs call occupies from start to end.  Several calls
// may be in flight (process_seqs_begin / _end): up to B200_INFLIGHT (default 4) run at once, in submission order, so that
// the device stages of chunk i+1 run under the host stages (rescue replay, pairing, SAM text) of chunk i.
struct Slot { bool busy = false; const void *staged_key = nullptr; int staged_n = 0, staged_lanes = 0; int64_t staged_bases = 0; };
static Slot g_slots[N_SLOTS];
static std::condition_variable g_slot_cv;
static uint64_t g_ticket_next = 0, g_ticket_serving = 0;
static b200_stats_t g_last_stats;

void engine_select_device(int dev) { g_device = dev; }
void engine_release()
{
	std::lock_guard<std::mutex> lk(g_mu);
	{
		std::lock_guard<std::mutex> sl(g_slot_mu);
		if (g_running > 0) { fprintf(stderr, "[mpibwa_b200] b200_gpu_release() while chunk jobs are in flight\n"); abort(); }
	}
	destroy_engines_locked();
}
Engine *engine_current() { return g_engines.empty() ? nullptr : g_engines[0]; }

// AuxGuard: exclusive use of the wrappers' engine and of the device for the lifetime of the guard.  The wrappers queue behind
// every waiting chunk job (ticket = max).
Engine *aux_acquire()
{
	g_aux_mu.lock();
	{
		std::lock_guard<std::mutex> lk(g_mu);
		if (g_engines.empty()) { g_aux_mu.unlock(); return nullptr; }
		if (!g_aux) g_aux = engine_clone(g_engines[0]);
	}
	g_gpu_turn.lock(~(uint64_t)0);
	return g_aux;
}
void aux_release() { g_gpu_turn.unlock(); g_aux_mu.unlock(); }
void last_stats(b200_stats_t *out) { std::lock_guard<std::mutex> lk(g_slot_mu); *out = g_last_stats; }

/* ------------------------------------------------------------------ option packing */

ExtOpt make_ext_opt(const mem_opt_t *opt)
{
	ExtOpt e;
	e.a = opt->a; e.b = opt->b; e.o_del = opt->o_del; e.e_del = opt->e_del; e.o_ins = opt->o_ins; e.e_ins = opt->e_ins;
	e.w = opt->w; e.zdrop = opt->zdrop; e.pen_clip5 = opt->pen_clip5; e.pen_clip3 = opt->pen_clip3;
	e.max_sc = 0;
	for (int i = 0; i < 25; ++i) { e.mat[i] = opt->mat[i]; if (opt->mat[i] > e.max_sc) e.max_sc = opt->mat[i]; }
	return e;
}

SwOpt make_sw_opt(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins)
{
	SwOpt s;
	s.o_del = o_del; s.e_del = e_del; s.o_ins = o_ins; s.e_ins = e_ins;
	int mn = 127, mx = 0;
	for (int i = 0; i < 25; ++i) { s.mat[i] = mat[i]; if (mat[i] < mn) mn = mat[i]; if (mat[i] > mx) mx = mat[i]; }
	s.max_sc = mx; s.shift = (256 - (mn & 0xff)) & 0xff;
	return s;
}

ChainOpt make_chain_opt(const mem_opt_t *opt)
{
	ChainOpt c;
	c.a = opt->a; c.o_del = opt->o_del; c.e_del = opt->e_del; c.o_ins = opt->o_ins; c.e_ins = opt->e_ins; c.w = opt->w;
	c.min_seed_len = opt->min_seed_len; c.max_chain_gap = opt->max_chain_gap; c.min_chain_weight = opt->min_chain_weight;
	c.max_chain_extend = opt->max_chain_extend; c.mask_level = opt->mask_level; c.drop_ratio = opt->drop_ratio;
	return c;
}

// mem_flt_chained_seeds (reference src/bwamem.c:598-615) returns at once when min_l > MEM_SHORT_EXT * l_query, i.e. for
// every read shorter than ~730 bases with default options; batches with a read to which it applies take the host chaining
static bool seed_sw_filter_applies(const mem_opt_t *opt, int l_query)
{
	double min_l = opt->min_chain_weight ? 1.1f * opt->min_chain_weight : 5.5f * log(l_query);
	return !(min_l > 0.05f * l_query);
}

SeedOpt make_seed_opt(const mem_opt_t *opt)
{
	SeedOpt s;
	s.min_seed_len = opt->min_seed_len;
	s.split_len = (int)(opt->min_seed_len * opt->split_factor + .499);
	s.split_width = opt->split_width;
	s.max_occ = opt->max_occ;
	s.max_mem_intv = (int)opt->max_mem_intv;
	return s;
}

/* ------------------------------------------------------------------ mate rescue (reference src/bwamem_pair.c:111-180) */

struct RescueKey { int end, anchor, r; };
struct RescueRes { RescueKey key; int64_t rb; SwRes res; };

// window of mem_matesw for orientation r; returns false when the reference would not run SW
static bool rescue_window(const mem_opt_t *opt, const bntseq_t *bns, const mem_pestat_t pes[4], const mem_alnreg_t *a,
                          int l_ms, int r, int64_t *rb_, int64_t *re_, int *is_rev_)
{
	const int64_t l_pac = bns->l_pac;
	int is_rev = (r >> 1 != (r & 1)), is_larger = !(r >> 1), rid = -1;
	int64_t rb, re;
	if (!is_rev) {
		rb = is_larger ? a->rb + pes[r].low : a->rb - pes[r].high;
		re = (is_larger ? a->rb + pes[r].high : a->rb - pes[r].low) + l_ms;
	} else {
		rb = (is_larger ? a->rb + pes[r].low : a->rb - pes[r].high) - l_ms;
		re = is_larger ? a->rb + pes[r].high : a->rb - pes[r].low;
	}
	if (rb < 0) rb = 0;
	if (re > l_pac << 1) re = l_pac << 1;
	if (rb >= re) return false;
	bns_clip_window(bns, &rb, (rb + re) >> 1, &re, &rid);
	*rb_ = rb; *re_ = re; *is_rev_ = is_rev;
	return a->rid == rid && re - rb >= opt->min_seed_len;
}

static void rescue_skip_mask(const mem_pestat_t pes[4], int64_t l_pac, const mem_alnreg_t *a, const RegVec &ma, int skip[4])
{
	for (int r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
	for (size_t i = 0; i < ma.size(); ++i) {
		int64_t dist;
		int r = infer_dir(l_pac, a->rb, ma[i].rb, &dist);
		if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
	}
}

// Replays mem_matesw for one anchor with precomputed SW results.  Returns false (and reports the missing job)
// when a result that the reference would compute here is not in `have`.
static bool rescue_replay_anchor(const mem_opt_t *opt, const bntseq_t *bns, const mem_pestat_t pes[4],
                                 const mem_alnreg_t *a, int l_ms, RegVec &ma, int end, int anchor,
                                 const std::vector<RescueRes> &have, RescueKey *missing)
{
	const int64_t l_pac = bns->l_pac;
	int skip[4], n = 0;
	rescue_skip_mask(pes, l_pac, a, ma, skip);
	if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return true;
	for (int r = 0; r < 4; ++r) {
		if (skip[r]) continue;
		int64_t rb, re;
		int is_rev;
		if (rescue_window(opt, bns, pes, a, l_ms, r, &rb, &re, &is_rev)) {
			const RescueRes *res = nullptr;
			for (const RescueRes &h : have)
				if (h.key.end == end && h.key.anchor == anchor && h.key.r == r) { res = &h; break; }
			if (!res) { missing->end = end; missing->anchor = anchor; missing->r = r; return false; }
			const SwRes &aln = res->res;
			if (aln.score >= opt->min_seed_len && aln.qb >= 0) {
				mem_alnreg_t b;
				memset(&b, 0, sizeof b);
				b.rid = a->rid;
				b.is_alt = a->is_alt;
				b.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
				b.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
				b.rb = is_rev ? (l_pac << 1) - (rb + aln.te + 1) : rb + aln.tb;
				b.re = is_rev ? (l_pac << 1) - (rb + aln.tb) : rb + aln.te + 1;
				b.score = aln.score;
				b.csub = aln.score2;
				b.secondary = -1;
				b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
				ma.push_back(b);
				size_t i, tmp;
				for (i = 0; i < ma.size() - 1; ++i)
					if (ma[i].score < b.score) break;
				tmp = i;
				for (i = ma.size() - 1; i > tmp; --i) ma[i] = ma[i - 1];
				ma[i] = b;
			}
			++n;
		}
		if (n) ma.resize(sort_dedup_patch(opt, 0, 0, 0, (int)ma.size(), ma.data()));
	}
	return true;
}

/* ------------------------------------------------------------------ the hot path */


// A chunk is cut into two (optionally four) sub-batches ("lanes") of whole pairs.  Each lane has its own engine (stream, scratch,
// resident reads; the index is shared) and runs seeding -> chaining -> extension -> regions, and later rescue -> SAM, on
// its own; two driver threads walk the lanes so that one lane's host stage overlaps the other lane's device stage.
struct Lane { Engine *eng; int r0, n; };

// SAM text of a chunk as blocks of consecutive records (b200_align_chunk: the caller wants one buffer, so the sweep appends
// the records of a block of pairs to one string instead of malloc()ing seqs[i].sam per read and concatenating afterwards)
struct SamBlocks { std::vector<std::vector<std::string>> lane; };      // [lane][block], in input order

static std::vector<Lane> make_lanes(int n, int slot, int want_default)
{
	// A synchronous mem_process_seqs call runs its chunk as two lanes: enough to overlap one lane's host stage with the
	// other's device stage while the kernels still see half a chunk per launch.  Chunk jobs (process_seqs_begin) overlap
	// whole chunks instead and run ONE lane, so every kernel sees the whole chunk (the DP rounds have a latency floor: half
	// batches cost the extension kernels a third of their efficiency).  B200_LANES / B200_LANE_MIN (reads per lane) override.
	const int want = getenv("B200_LANES") ? atoi(getenv("B200_LANES")) : want_default;
	const int lane_min = getenv("B200_LANE_MIN") ? atoi(getenv("B200_LANE_MIN")) : 65536;
	int k = want >= 4 ? 4 : want >= 2 ? 2 : 1;
	while (k > 1 && n < k * lane_min) k >>= 1;
	std::vector<Lane> lanes;
	int r0 = 0;
	for (int i = 0; i < k; ++i) {
		int r1 = i + 1 == k ? n : (int)(((int64_t)n * (i + 1) / k) & ~1ll);
		lanes.push_back({ engine_lane(slot, i), r0, r1 - r0 });
		r0 = r1;
	}
	return lanes;
}

// encode the reads of one lane in place, flatten them and make them resident in HBM
static int64_t stage_lane_reads(const mem_opt_t *opt, const Lane &L, bseq1_t *seqs_all)
{
	const int n = L.n;
	bseq1_t *seqs = seqs_all + L.r0;
	const int nt = opt->n_threads > 0 ? opt->n_threads : 1;
	std::vector<int64_t> off(n + 1);
	off[0] = 0;
	for (int i = 0; i < n; ++i) off[i + 1] = off[i] + seqs[i].l_seq;
	uint8_t *codes = stage_read_buffer(L.eng, off[n] + 16);
	parallel_for(nt, n, 4096, [&](int, int64_t b, int64_t e) {
		for (int64_t i = b; i < e; ++i) {
			char *s = seqs[i].seq;
			uint8_t *d = &codes[off[i]];
			const int l = seqs[i].l_seq;
			int j = 0;
#if defined(__SSE2__) && !defined(B200_NO_SSE_ENCODE)
			// sixteen bases at a time while they are all A/C/G/T in either case: code = ((c >> 1) ^ (c >> 2)) & 3;
			// anything else (N, IUPAC codes, bytes that are already codes) goes through the table below
			const __m128i lower = _mm_set1_epi8(0x20), three = _mm_set1_epi8(3);
			for (; j + 16 <= l; j += 16) {
				const __m128i c = _mm_loadu_si128((const __m128i *)(s + j)), u = _mm_or_si128(c, lower);
				const __m128i ok = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('a')), _mm_cmpeq_epi8(u, _mm_set1_epi8('c'))),
				                                _mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('g')), _mm_cmpeq_epi8(u, _mm_set1_epi8('t'))));
				if (_mm_movemask_epi8(ok) != 0xffff) break;
				const __m128i code = _mm_and_si128(_mm_xor_si128(_mm_srli_epi16(c, 1), _mm_srli_epi16(c, 2)), three);
				_mm_storeu_si128((__m128i *)(s + j), code);
				_mm_storeu_si128((__m128i *)(d + j), code);
			}
#endif
			for (; j < l; ++j) {
				s[j] = s[j] < 4 ? s[j] : (char)kNt4[(uint8_t)s[j]];
				d[j] = (uint8_t)s[j];
			}
		}
	});
	stage_upload_reads(L.eng, n, off.data(), codes);
	return off[n];
}

// b200_stage_reads(): make the reads of a coming mem_process_seqs call resident ahead of time (in a free chunk slot,
// which the call for the same `seqs` then takes over)
void stage_reads(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int n, bseq1_t *seqs)
{
	engine_for(bwt, bns, pac);
	int slot = -1;
	{
		std::unique_lock<std::mutex> lk(g_slot_mu);
		g_slot_cv.wait(lk, [&] { for (int k = 0; k < N_SLOTS; ++k) if (!g_slots[k].busy && !g_slots[k].staged_key) { slot = k; return true; } return false; });
		g_slots[slot].busy = true;
	}
	int64_t bases = 0;
	std::vector<Lane> lanes = make_lanes(n, slot, 1);
	for (const Lane &L : lanes) bases += stage_lane_reads(opt, L, seqs);
	std::lock_guard<std::mutex> lk(g_slot_mu);
	g_slots[slot].busy = false; g_slots[slot].staged_key = (const void *)seqs; g_slots[slot].staged_n = n; g_slots[slot].staged_bases = bases;
	g_slots[slot].staged_lanes = (int)lanes.size();
	g_slot_cv.notify_all();
}

template <class F>
static void drive_lanes(std::vector<Lane> &lanes, F body)
{
	std::atomic<int> next(0);
	auto run = [&]() { for (;;) { int k = next.fetch_add(1); if (k >= (int)lanes.size()) break; body(lanes[k]); } };
	if (lanes.size() > 1) { std::thread other(run); run(); other.join(); }
	else run();
}

#define GPU_STAGE(call) do { DeviceTurnGuard gpu_lk(ticket); call; } while (0)

static void process_seqs_slot(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                              int64_t n_processed_all, int n_all, bseq1_t *seqs_all, const mem_pestat_t *pes0,
                              int slot, int want_lanes, bool staged, int64_t staged_bases, b200_stats_t *stats_out, SamBlocks *sam_blocks,
                              uint64_t ticket)
{
	engine_for(bwt, bns, pac);
	std::vector<Lane> lanes = make_lanes(n_all, slot, want_lanes);
	if (sam_blocks) sam_blocks->lane.resize(lanes.size());
	if (staged && (int)lanes.size() != want_lanes) { fprintf(stderr, "[mpibwa_b200] staged reads do not match the lane split of the call\n"); abort(); }
	for (const Lane &L : lanes) memset(static_cast<b200_stats_t *>(&engine_stats(L.eng)), 0, sizeof(b200_stats_t));
	const int nt = opt->n_threads > 0 ? opt->n_threads : 1;
	const double t_start = now_ms();
	const bool pe = (opt->flag & MEM_F_PE) != 0;
	const int64_t l_pac = bns->l_pac;
	std::vector<RegVec> regs_all(n_all);

	// ================= phase 1, per lane: reads -> seeds -> chains -> regions
	drive_lanes(lanes, [&](Lane &L) {
	Engine *eng = L.eng;
	Stats &st = engine_stats(eng);
	const int n = L.n;
	bseq1_t *seqs = seqs_all + L.r0;
	RegVec *regs = regs_all.data() + L.r0;
	double t0 = now_ms(), t1;
	// ---- encode (reference src/bwamem.c:1057-1058), flatten and upload - unless b200_stage_reads() already did
	if (!staged) st.n_bases = stage_lane_reads(opt, L, seqs_all);       // (own stream and buffers: no need to hold the device)
	st.n_reads = n;

	// ---- chaining on the device unless a read is long enough for mem_flt_chained_seeds (B200_CHAIN=host forces the host path)
	bool dev_chain = !(getenv("B200_CHAIN") && !strcmp(getenv("B200_CHAIN"), "host"));
	{
		std::vector<int8_t> len_rule(4096, 0);
		for (int i = 0; i < n && dev_chain; ++i) {
			const int l = seqs[i].l_seq;
			if (l < opt->min_seed_len) continue;
			int8_t rule = l < 4096 ? len_rule[l] : 0;
			if (rule == 0) { rule = seed_sw_filter_applies(opt, l) ? 2 : 1; if (l < 4096) len_rule[l] = rule; }
			if (rule == 2) dev_chain = false;
		}
	}

	// ---- seeding on the device
	SeedOut sd;
	const bool chain_check = dev_chain && getenv("B200_CHAIN") && !strcmp(getenv("B200_CHAIN"), "check");   // run both, compare, abort on a difference
	GPU_STAGE(stage_seed(eng, make_seed_opt(opt), sd, dev_chain && !chain_check));
	t1 = now_ms(); st.ms_seed = t1 - t0; t0 = t1;
	st.n_seeds = sd.n_seeds;

	ExtIn xin;
	std::vector<int32_t> chk_co, chk_srt;
	std::vector<DChain> chk_ch;
	std::vector<DSeed> chk_se;
	if (dev_chain) {
		GPU_STAGE(stage_chain(eng, make_chain_opt(opt), xin, chain_check));
		st.n_chains = xin.n_chains;
		if (chain_check) {
			chk_co.assign(xin.chain_off, xin.chain_off + n + 1); chk_ch.assign(xin.chains, xin.chains + xin.n_chains);
			chk_se.assign(xin.seeds, xin.seeds + xin.n_seeds); chk_srt.assign(xin.srt, xin.srt + xin.n_seeds);
		}
	}
	if (!dev_chain || chain_check) {
	// ---- chaining + chain filtering on host threads
	std::vector<std::vector<HChain>> chains(n);
	std::vector<int32_t> n_chain_of(n), n_seed_of(n);
	parallel_for(nt, n, 512, [&](int, int64_t b, int64_t e) {
		for (int64_t i = b; i < e; ++i) {
			build_chains(opt, bns, seqs[i].l_seq, sd.seeds + sd.seed_off[i], sd.seed_off[i + 1] - sd.seed_off[i], sd.l_rep[i], chains[i]);
			filter_chains(opt, chains[i]);
			int32_t ns = 0;
			for (auto &c : chains[i]) ns += (int32_t)c.seeds.size();
			n_chain_of[i] = (int32_t)chains[i].size(); n_seed_of[i] = ns;
		}
	});
	if (getenv("B200_DEBUG")) fprintf(stderr, "[chain] build+filter %.1f ms\n", now_ms() - t0);

	// ---- mem_flt_chained_seeds (reference src/bwamem.c:571-615): only reads of >= ~730 bp get here
	{
		std::vector<SwJob> jobs;
		std::vector<HSeed *> owner;
		std::vector<int> long_reads;
		std::vector<int8_t> len_rule(4096, 0);                   // per read length: 1 = the filter does not apply, 2 = it does
		for (int i = 0; i < n; ++i) {
			int l_query = seqs[i].l_seq;
			if (l_query <= 0 || n_chain_of[i] == 0) continue;
			int8_t rule = l_query < 4096 ? len_rule[l_query] : 0;
			if (rule == 0) {
				double min_l = opt->min_chain_weight ? 1.1f * opt->min_chain_weight : 5.5f * log(l_query);
				rule = min_l > 0.05f * l_query ? 1 : 2;
				if (l_query < 4096) len_rule[l_query] = rule;
			}
			if (rule == 1) continue;
			long_reads.push_back(i);
			for (auto &c : chains[i])
				for (HSeed &s : c.seeds) {
					s.score = -1;                       // mem_seed_sw's "no need to do SW"
					if (s.len >= 200) continue;
					int qb = s.qbeg, qe = s.qbeg + s.len, rid;
					int64_t rb = s.rbeg, re = s.rbeg + s.len, mid = (rb + re) >> 1;
					qb -= 50; qb = qb > 0 ? qb : 0;
					qe += 50; qe = qe < l_query ? qe : l_query;
					rb -= 50; rb = rb > 0 ? rb : 0;
					re += 50; re = re < l_pac << 1 ? re : l_pac << 1;
					if (rb < l_pac && l_pac < re) { if (mid < l_pac) re = l_pac; else rb = l_pac; }
					if (qe - qb >= 200 || re - rb >= 200) continue;
					bns_clip_window(bns, &rb, mid, &re, &rid);
					SwJob j;
					j.rb = rb; j.tlen = (int)(re - rb); j.read = i; j.is_rev = 0; j.xtra = KSW_XSTART; j.q_beg = qb; j.q_len = qe - qb;
					jobs.push_back(j);
					owner.push_back(&s);
				}
		}
		if (!jobs.empty()) {
			std::vector<SwRes> res;
			GPU_STAGE(stage_sw(eng, make_sw_opt(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins), jobs, res));
			for (size_t x = 0; x < owner.size(); ++x) owner[x]->score = res[x].score;
		}
		for (int i : long_reads) {
			int l_query = seqs[i].l_seq;
			double min_l = opt->min_chain_weight ? 1.1f * opt->min_chain_weight : 5.5f * log(l_query);
			int min_HSP_score = (int)(opt->a * min_l + .499);
			for (auto &c : chains[i]) {
				size_t k = 0;
				for (size_t j = 0; j < c.seeds.size(); ++j) {
					HSeed s = c.seeds[j];
					if (s.score < 0 || s.score >= min_HSP_score) {
						s.score = s.score < 0 ? s.len * opt->a : s.score;
						c.seeds[k++] = s;
					}
				}
				c.seeds.resize(k);
			}
			int32_t ns = 0;
			for (auto &c : chains[i]) ns += (int32_t)c.seeds.size();
			n_seed_of[i] = ns;
		}
	}

	// ---- flatten chains for the extension stage (two sweeps: sizes, then a parallel fill at the prefix offsets)
	int32_t *chain_off = (int32_t *)stage_pinned(eng, PIN_CHAIN_OFF, sizeof(int32_t) * (n + 1));
	DChain *dchains = nullptr;
	DSeed *dseeds = nullptr;
	int32_t *srt = nullptr;
	{
		std::vector<int64_t> seed_at(n + 1);
		int64_t nc = 0, ns = 0;
		for (int i = 0; i < n; ++i) { chain_off[i] = (int32_t)nc; seed_at[i] = ns; nc += n_chain_of[i]; ns += n_seed_of[i]; }
		chain_off[n] = (int32_t)nc; seed_at[n] = ns;
		if (nc > 0x7fffffffLL || ns > 0x7fffffffLL) { fprintf(stderr, "[mpibwa_b200] too many chains/seeds in one chunk\n"); abort(); }
		dchains = (DChain *)stage_pinned(eng, PIN_CHAINS, sizeof(DChain) * (nc + 1));
		dseeds = (DSeed *)stage_pinned(eng, PIN_DSEEDS, sizeof(DSeed) * (ns + 1));
		srt = (int32_t *)stage_pinned(eng, PIN_SRT, sizeof(int32_t) * (ns + 1));
		xin.n_reads = n; xin.chain_off = chain_off; xin.chains = dchains; xin.n_chains = nc; xin.seeds = dseeds; xin.n_seeds = ns; xin.srt = srt;
		parallel_for(nt, n, 2048, [&](int, int64_t b, int64_t e) {
			std::vector<uint64_t> key;
			for (int64_t i = b; i < e; ++i) {
				int64_t ci = chain_off[i], si = seed_at[i];
				for (auto &c : chains[i]) {
					DChain d;
					d.seed_beg = (int32_t)si; d.n_seeds = (int32_t)c.seeds.size();
					d.rid = c.rid; d.frac_rep = c.frac_rep; d.rmax0 = d.rmax1 = 0;
					if (!c.seeds.empty()) {
						int64_t rmax[2];
						chain_window(opt, bns, seqs[i].l_seq, c, rmax);
						d.rmax0 = rmax[0]; d.rmax1 = rmax[1];
					}
					key.resize(c.seeds.size());
					for (size_t k = 0; k < c.seeds.size(); ++k) {
						const HSeed &s = c.seeds[k];
						dseeds[si + k] = {s.rbeg, s.qbeg, s.len, s.score, 0};
						key[k] = (uint64_t)s.score << 32 | k;
					}
					std::sort(key.begin(), key.end());
					for (size_t k = 0; k < key.size(); ++k) srt[si + k] = (int32_t)(uint32_t)key[k];
					dchains[ci++] = d;
					si += (int64_t)c.seeds.size();
				}
				std::vector<HChain>().swap(chains[i]);      // release in parallel
			}
		});
		st.n_chains = nc;
	}
	if (getenv("B200_DEBUG")) fprintf(stderr, "[chain] +flatten %.1f ms\n", now_ms() - t0);
	chains.clear(); chains.shrink_to_fit();
	xin.on_device = false;
	if (chain_check) {
		bool same = (int64_t)chk_ch.size() == xin.n_chains && (int64_t)chk_se.size() == xin.n_seeds && !memcmp(chk_co.data(), chain_off, sizeof(int32_t) * (n + 1));
		for (int64_t k = 0; same && k < xin.n_chains; ++k) {
			const DChain &x = chk_ch[k], &y = dchains[k];
			same = x.rmax0 == y.rmax0 && x.rmax1 == y.rmax1 && x.seed_beg == y.seed_beg && x.n_seeds == y.n_seeds && x.rid == y.rid && !memcmp(&x.frac_rep, &y.frac_rep, 4);
		}
		for (int64_t k = 0; same && k < xin.n_seeds; ++k) {
			const DSeed &x = chk_se[k], &y = dseeds[k];
			same = x.rbeg == y.rbeg && x.qbeg == y.qbeg && x.len == y.len && x.score == y.score && chk_srt[k] == srt[k];
		}
		if (!same) { fprintf(stderr, "[mpibwa_b200] B200_CHAIN=check: the chaining stage disagrees with the host chaining (%lld vs %lld chains, %lld vs %lld seeds)\n",
			(long long)chk_ch.size(), (long long)xin.n_chains, (long long)chk_se.size(), (long long)xin.n_seeds); abort(); }
	}
	}
	t1 = now_ms(); st.ms_chain_host = t1 - t0; t0 = t1;

	// ---- chain2aln / ksw_extend2 on the device
	ExtRegs xr;
	GPU_STAGE(stage_extend(eng, make_ext_opt(opt), xin, xr));
	const DReg *dregs = xr.regs;
	const int64_t *reg_off = xr.reg_off;
	t1 = now_ms(); st.ms_extend = t1 - t0; t0 = t1;

	// ---- mem_sort_dedup_patch + ALT marking (reference src/bwamem.c:1073-1085)
	parallel_for(nt, n, 512, [&](int, int64_t b, int64_t e) {
		for (int64_t i = b; i < e; ++i) {
			int nr = (int)(reg_off[i + 1] - reg_off[i]);
			if (nr == 0) continue;
			int64_t base = reg_off[i];
			RegVec &rv = regs[i];
			rv.resize(nr);
			for (int k = 0; k < nr; ++k) {
				const DReg &d = dregs[base + k];
				mem_alnreg_t &a = rv[k];
				memset(&a, 0, sizeof a);
				a.rb = d.rb; a.re = d.re; a.qb = d.qb; a.qe = d.qe; a.rid = d.rid; a.score = d.score; a.truesc = d.truesc;
				a.w = d.w; a.seedcov = d.seedcov; a.seedlen0 = d.seedlen0; a.frac_rep = d.frac_rep;
			}
			rv.resize(sort_dedup_patch(opt, bns, pac, (uint8_t *)seqs[i].seq, nr, rv.data()));
			for (auto &p : rv)
				if (p.rid >= 0 && bns->anns[p.rid].is_alt) p.is_alt = 1;
		}
	});
	t1 = now_ms(); st.ms_regs_host = t1 - t0; t0 = t1;
	});

	// ================= insert-size statistics: the one chunk-global reduction (reference src/bwamem.c:1226-1229)
	mem_pestat_t pes[4];
	const double t_pes = now_ms();
	if (pe) {
		if (pes0) memcpy(pes, pes0, 4 * sizeof(mem_pestat_t));
		else pestat(opt, l_pac, n_all, regs_all.data(), pes);
	}
	const double ms_pestat = now_ms() - t_pes;

	// ================= phase 2, per lane: mate rescue -> pairing, mapQ, CIGAR, SAM text
	drive_lanes(lanes, [&](Lane &L) {
	Engine *eng = L.eng;
	Stats &st = engine_stats(eng);
	const int n = L.n;
	bseq1_t *seqs = seqs_all + L.r0;
	RegVec *regs = regs_all.data() + L.r0;
	const int64_t n_processed = n_processed_all + L.r0;
	double t0 = now_ms(), t1;

	// ---- mate rescue: SW jobs on the device, then the sequential insert/skip logic replayed per pair
	if (pe && !(opt->flag & MEM_F_NO_RESCUE)) {
		const int n_pairs = n >> 1;
		const int xtra_base = KSW_XSUBO | KSW_XSTART | (opt->min_seed_len * opt->a);
		SwOpt so = make_sw_opt(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins);
		std::vector<int> pending;
		std::vector<std::vector<RescueKey>> want(n_pairs);           // jobs to run this round, by pair
		std::unordered_map<int, std::vector<RescueRes>> known;       // results carried over by deferred pairs
		// round 0: every (anchor, orientation) not ruled out by the mate's pre-rescue regions
		parallel_for(nt, n_pairs, 1024, [&](int, int64_t b, int64_t e) {
			for (int64_t p = b; p < e; ++p) {
				for (int i = 0; i < 2; ++i) {
					const RegVec &ai = regs[p << 1 | i], &ma = regs[p << 1 | !i];
					int l_ms = seqs[p << 1 | !i].l_seq;
					int nb = 0;
					for (size_t j = 0; j < ai.size() && nb < opt->max_matesw; ++j) {
						if (ai[j].score < ai[0].score - opt->pen_unpaired) continue;
						int skip[4];
						rescue_skip_mask(pes, l_pac, &ai[j], ma, skip);
						for (int r = 0; r < 4; ++r) {
							int64_t rb, re; int is_rev;
							if (!skip[r] && rescue_window(opt, bns, pes, &ai[j], l_ms, r, &rb, &re, &is_rev))
								want[p].push_back({i, nb, r});
						}
						++nb;
					}
				}
			}
		});
		// a pair none of whose anchors asks for an alignment cannot change in mem_matesw: only the others are replayed
		for (int i = 0; i < n_pairs; ++i) if (!want[i].empty()) pending.push_back(i);
		while (!pending.empty()) {
			std::vector<SwJob> jobs;
			std::vector<RescueKey> job_key;
			std::vector<int64_t> first(pending.size() + 1, 0);
			for (size_t x = 0; x < pending.size(); ++x) {
				int p = pending[x];
				for (const RescueKey &k : want[p]) {
					const RegVec &ai = regs[p << 1 | k.end];
					int nb = -1; const mem_alnreg_t *a = nullptr;
					for (size_t j = 0; j < ai.size(); ++j) {
						if (ai[j].score < ai[0].score - opt->pen_unpaired) continue;
						if (++nb == k.anchor) { a = &ai[j]; break; }
					}
					int l_ms = seqs[p << 1 | !k.end].l_seq;
					int64_t rb, re; int is_rev;
					rescue_window(opt, bns, pes, a, l_ms, k.r, &rb, &re, &is_rev);
					SwJob j;
					j.rb = rb; j.tlen = (int)(re - rb); j.read = p << 1 | !k.end; j.is_rev = is_rev;
					j.xtra = xtra_base | (l_ms * opt->a < 250 ? KSW_XBYTE : 0);
					j.q_beg = 0; j.q_len = l_ms;
					jobs.push_back(j); job_key.push_back(k);
				}
				first[x + 1] = (int64_t)jobs.size();
			}
			std::vector<SwRes> res;
			if (!jobs.empty()) GPU_STAGE(stage_sw(eng, so, jobs, res));
			std::vector<char> done(pending.size(), 0);
			std::vector<RescueKey> miss(pending.size());
			std::vector<std::vector<RescueRes>> carry(pending.size());
			parallel_for(nt, (int64_t)pending.size(), 1024, [&](int, int64_t b, int64_t e) {
				std::vector<RescueRes> have;
				for (int64_t x = b; x < e; ++x) {
					int p = pending[x];
					have.clear();
					auto it = known.find(p);            // read-only during this loop
					if (it != known.end()) have = it->second;
					for (int64_t y = first[x]; y < first[x + 1]; ++y) have.push_back({job_key[y], jobs[y].rb, res[y]});
					RegVec a[2] = { regs[p << 1], regs[p << 1 | 1] };
					RegVec anchors[2];
					for (int i = 0; i < 2; ++i)
						for (size_t j = 0; j < a[i].size(); ++j)
							if (a[i][j].score >= a[i][0].score - opt->pen_unpaired) anchors[i].push_back(a[i][j]);
					bool ok = true;
					for (int i = 0; i < 2 && ok; ++i)
						for (size_t j = 0; j < anchors[i].size() && (int)j < opt->max_matesw && ok; ++j)
							ok = rescue_replay_anchor(opt, bns, pes, &anchors[i][j], seqs[p << 1 | !i].l_seq, a[!i], i, (int)j, have, &miss[x]);
					if (ok) { regs[p << 1].swap(a[0]); regs[p << 1 | 1].swap(a[1]); done[x] = 1; }
					else carry[x].swap(have);
				}
			});
			std::vector<int> again;
			for (size_t x = 0; x < pending.size(); ++x) {
				if (done[x]) continue;
				int p = pending[x];
				known[p].swap(carry[x]);
				want[p].assign(1, miss[x]);
				again.push_back(p);
			}
			pending.swap(again);
		}
	}
	t1 = now_ms(); st.ms_rescue = t1 - t0; t0 = t1;

	// ---- primary marking, pairing, mapQ, CIGAR and SAM text (reference worker2, src/bwamem.c:1187-1203).
	// First every region that mem_reg2aln may be asked about is queued for the CIGAR stage on the device (a function of
	// the region alone: no dry run of the pairing logic), then the sweep over the pairs looks its alignments up.
	{
		const int64_t n_units = pe ? n >> 1 : n;
		const int per = pe ? 2 : 1;
		auto run_unit = [&](int64_t u, RegVec *r) {
			if (pe) sam_pe_finish(opt, bns, pac, pes, (uint64_t)((n_processed >> 1) + u), &seqs[u << 1], r);
			else {
				mark_primary_se(opt, (int)r[0].size(), r[0].data(), n_processed + u);
				if (opt->flag & MEM_F_PRIMARY5) reorder_primary5(opt->T, r[0]);
				reg2sam(opt, bns, pac, &seqs[u], r[0], 0, 0);
			}
		};
		struct UnitJobs { int32_t tid, start, count; };
		std::vector<UnitJobs> uj(n_units);
		std::vector<std::vector<GlobalJob>> tjobs(nt);
		parallel_for(nt, n_units, 1024, [&](int tid, int64_t b, int64_t e) {
			std::vector<GlobalJob> &out = tjobs[tid];
			for (int64_t u = b; u < e; ++u) {
				const int32_t start = (int32_t)out.size();
				for (int k = 0; k < per; ++k) {
					const RegVec &rv = regs[u * per + k];
					// a region below the output threshold, or far below the read's best hit (never a primary, never in XA), is
					// not worth a device job; should the sweep ask for it after all, reg2aln aligns it itself
					int best = 0;
					for (const mem_alnreg_t &a : rv) best = a.score > best ? a.score : best;
					for (const mem_alnreg_t &a : rv) {
						if (a.score < opt->T || a.score < best * opt->XA_drop_ratio - opt->pen_unpaired) continue;
						GlobalJob j;
						if (reg_global_job(opt, bns, &a, (int)(u * per + k), &j)) out.push_back(j);
					}
				}
				uj[u] = { tid, start, (int32_t)out.size() - start };
			}
		});
		std::vector<int64_t> tbase(nt + 1, 0);
		for (int t = 0; t < nt; ++t) tbase[t + 1] = tbase[t] + (int64_t)tjobs[t].size();
		std::vector<GlobalJob> gjobs(tbase[nt]);
		GlobalOpt go;
		go.o_del = opt->o_del; go.e_del = opt->e_del; go.o_ins = opt->o_ins; go.e_ins = opt->e_ins; go.a = opt->a; go.w_max = opt->w << 2;
		memcpy(go.mat, opt->mat, 25);
		int64_t zb = 0;
		for (int t = 0; t < nt; ++t)
			for (size_t k = 0; k < tjobs[t].size(); ++k) {
				GlobalJob j = tjobs[t][k];
				j.zoff = zb;
				zb += (global_z_need(go, j.qe - j.qb, (int)(j.re - j.rb), j.w2, &j.wmax) + 15) & ~(int64_t)15;
				gjobs[tbase[t] + (int64_t)k] = j;
			}
		tjobs.clear();
		double tg = now_ms();
		st.ms_sam_plan = tg - t0;
		const GlobalRes *gres = nullptr;
		if (!gjobs.empty()) GPU_STAGE(gres = stage_global(eng, go, gjobs, zb));
		st.ms_global = now_ms() - tg;
		std::atomic<int64_t> n_host_dp(0);
		const int64_t sweep_grain = 256;
		std::vector<std::string> *blocks = nullptr;
		if (sam_blocks) {
			blocks = &sam_blocks->lane[&L - lanes.data()];
			blocks->resize((size_t)((n_units + sweep_grain - 1) / sweep_grain));
			for (std::string &b : *blocks) b.clear();                          // (keeps the capacity)
		}
		parallel_for(nt, n_units, sweep_grain, [&](int, int64_t b, int64_t e) {
			AlignCtx &cx = align_ctx();
			cx.mode = AlignCtx::LOOKUP; cx.n_host_dp = 0;
			if (blocks) { cx.sink = &(*blocks)[(size_t)(b / sweep_grain)]; cx.sink->reserve((size_t)(e - b) * per * 448); }
			// a region's reference window is a random 40-byte read of the 2-bit reference (a DRAM miss per region, the largest
			// single cost of the sweep once the arithmetic was trimmed): touch the windows of a pair a few pairs ahead
			auto prefetch_unit = [&](int64_t v) {
				for (int k = 0; k < per; ++k) {
					const RegVec &rv = regs[v * per + k];
					for (size_t x = 0; x < rv.size() && x < 3; ++x) {
						const int64_t p = rv[x].rb < l_pac ? rv[x].rb : (l_pac << 1) - rv[x].re;
						if (p >= 0 && p < l_pac) { __builtin_prefetch(pac + (p >> 2)); __builtin_prefetch(pac + (p >> 2) + 64); }
					}
				}
			};
			for (int64_t v = b; v < e && v < b + 4; ++v) prefetch_unit(v);
			for (int64_t u = b; u < e; ++u) {
				if (u + 4 < e) prefetch_unit(u + 4);
				const int64_t first = tbase[uj[u].tid] + uj[u].start;
				cx.jobs = gjobs.data() + first; cx.res = gres ? gres + first : nullptr; cx.n_jobs = uj[u].count;
				for (int k = 0; k < per; ++k) { cx.seq_ptr[k] = seqs[u * per + k].seq; cx.read_idx[k] = (int)(u * per + k); }
				run_unit(u, &regs[u * per]);
			}
			cx.mode = AlignCtx::DIRECT; cx.jobs = nullptr; cx.res = nullptr; cx.n_jobs = 0; cx.sink = nullptr;
			n_host_dp += cx.n_host_dp;
			host_prof_flush();
		});
		st.n_global_host = n_host_dp;
	}
	t1 = now_ms(); st.ms_sam_host = t1 - t0;
	});

	// ================= merge the per-lane counters into the primary engine's record (b200_get_stats reads it)
	Stats &st = engine_stats(lanes[0].eng);
	for (size_t k = 1; k < lanes.size(); ++k) {
		const b200_stats_t &o = engine_stats(lanes[k].eng);
		st.ms_seed += o.ms_seed; st.ms_sa += o.ms_sa; st.ms_chain_host += o.ms_chain_host; st.ms_extend += o.ms_extend;
		st.ms_regs_host += o.ms_regs_host; st.ms_rescue += o.ms_rescue; st.ms_sam_host += o.ms_sam_host;
		st.ms_k_smem += o.ms_k_smem; st.ms_k_sa += o.ms_k_sa; st.ms_k_extend += o.ms_k_extend; st.ms_k_sw += o.ms_k_sw; st.ms_k_global += o.ms_k_global;
		st.n_reads += o.n_reads; st.n_bases += o.n_bases; st.n_intv += o.n_intv; st.n_seeds += o.n_seeds; st.n_chains += o.n_chains;
		st.n_extend_jobs += o.n_extend_jobs; st.extend_cells += o.extend_cells; st.n_sw_jobs += o.n_sw_jobs; st.sw_cells += o.sw_cells;
		st.n_global_jobs += o.n_global_jobs; st.global_cells += o.global_cells;
		st.fm_occ_blocks += o.fm_occ_blocks; st.fm_sa_steps += o.fm_sa_steps; st.fm_sa_lookups += o.fm_sa_lookups;
		st.n_launches += o.n_launches; st.h2d_bytes += o.h2d_bytes; st.d2h_bytes += o.d2h_bytes;
		st.ms_k_extend_dp += o.ms_k_extend_dp; st.n_extend_rounds += o.n_extend_rounds;
		st.ms_sam_plan += o.ms_sam_plan; st.ms_global += o.ms_global; st.ms_k_chain += o.ms_k_chain; st.n_global_host += o.n_global_host;
	}
	st.ms_rescue += ms_pestat;
	if (staged) st.n_bases = staged_bases;
	st.ms_total = now_ms() - t_start;
	if (bwa_verbose >= 3)
		fprintf(stderr, "[M::%s] Processed %d reads in %.3f real sec (%s, %d lane%s; stage walls summed over lanes: seed %.0f ms, chain %.0f, extend %.0f, regs %.0f, rescue %.0f, sam %.0f [plan %.0f, cigar stage %.0f])\n",
		        "mem_process_seqs", n_all, st.ms_total * 1e-3, engine_kind(), (int)lanes.size(), lanes.size() > 1 ? "s" : "", st.ms_seed, st.ms_chain_host,
		        st.ms_extend, st.ms_regs_host, st.ms_rescue, st.ms_sam_host, st.ms_sam_plan, st.ms_global);
	if (stats_out) *stats_out = st;
	host_prof_report("SAM sweep");
	std::lock_guard<std::mutex> lk(g_slot_mu);
	g_last_stats = st;
}

/* ------------------------------------------------------------------ chunk jobs */

struct SeqJob {
	std::thread th;
	b200_stats_t stats;
	SamBlocks *sam = nullptr;        // the slot's block buffers while the job holds the slot
};

// block buffers live with the slot and keep their capacity from chunk to chunk (no 200 KB allocations per block per chunk)
static SamBlocks g_slot_sam[N_SLOTS];

extern "C" void *b200_big_alloc(size_t bytes);

// concatenates the SAM blocks of a finished job into one malloc()ed, NUL-terminated buffer (parallel copy); returns its length
int64_t job_take_sam(SeqJob *j, int n_threads, char **out)
{
	std::vector<const std::string *> parts;
	for (auto &ln : j->sam->lane) for (auto &b : ln) parts.push_back(&b);
	std::vector<size_t> at(parts.size() + 1, 0);
	for (size_t k = 0; k < parts.size(); ++k) at[k + 1] = at[k] + parts[k]->size();
	char *buf = (char *)b200_big_alloc(at.back() + 1);      // recycled through b200_free()
#if defined(MADV_HUGEPAGE)
	if (at.back() >= ((size_t)8 << 20)) {       // hundreds of MB touched once: ask for huge pages instead of 65 k page faults
		const uintptr_t lo = ((uintptr_t)buf + ((size_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1), hi = ((uintptr_t)buf + at.back()) & ~(((uintptr_t)2 << 20) - 1);
		if (hi > lo) madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
	}
#endif
	parallel_for(n_threads, (int64_t)parts.size(), 16, [&](int, int64_t b, int64_t e) {
		for (int64_t k = b; k < e; ++k) memcpy(buf + at[k], parts[k]->data(), parts[k]->size());
	});
	buf[at.back()] = 0;
	*out = buf;
	return (int64_t)at.back();
}

SeqJob *process_seqs_begin(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                           int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0,
                           void (*after)(void *, SeqJob *), void *arg, int want_lanes, bool sam_as_blocks,
                           void (*before)(void *, bseq1_t **, int *))
{
	engine_for(bwt, bns, pac);
	SeqJob *j = new SeqJob();
	memset(&j->stats, 0, sizeof j->stats);
	int slot = -1;
	bool staged = false;
	int64_t staged_bases = 0;
	uint64_t ticket;
	{
		std::unique_lock<std::mutex> lk(g_slot_mu);
		for (int k = 0; k < N_SLOTS && seqs; ++k)
			if (!g_slots[k].busy && g_slots[k].staged_key == (const void *)seqs && g_slots[k].staged_n == n) { slot = k; staged = true; staged_bases = g_slots[k].staged_bases; want_lanes = g_slots[k].staged_lanes; }
		if (slot < 0)
			g_slot_cv.wait(lk, [&] { for (int k = 0; k < N_SLOTS; ++k) if (!g_slots[k].busy && !g_slots[k].staged_key) { slot = k; return true; } return false; });
		g_slots[slot].busy = true; g_slots[slot].staged_key = nullptr;
		ticket = g_ticket_next++;
	}
	const mem_pestat_t *pes = pes0;
	j->th = std::thread([=]() {
		static const int limit = getenv("B200_INFLIGHT") ? std::max(1, atoi(getenv("B200_INFLIGHT"))) : 4;
		bseq1_t *seqs_ = seqs;
		int n_ = n;
		if (before) before(arg, &seqs_, &n_);     // (b200_align_fastq_begin: parse + interleave on the job thread, before the job's turn)
		{
			std::unique_lock<std::mutex> lk(g_slot_mu);
			g_slot_cv.wait(lk, [&] { return ticket == g_ticket_serving && g_running < limit; });
			++g_running; ++g_ticket_serving;
			g_slot_cv.notify_all();
		}
		j->sam = sam_as_blocks ? &g_slot_sam[slot] : nullptr;
		process_seqs_slot(opt, bwt, bns, pac, n_processed, n_, seqs_, pes, slot, want_lanes, staged, staged_bases, &j->stats, j->sam, ticket);
		if (after) after(arg, j);                // (SAM concatenation out of the slot's block buffers: host work that overlaps the next chunk)
		{
			std::lock_guard<std::mutex> lk(g_slot_mu);
			--g_running; g_slots[slot].busy = false;
			g_slot_cv.notify_all();
		}
	});
	return j;
}

void process_seqs_end(SeqJob *j, b200_stats_t *stats)
{
	j->th.join();
	if (stats) *stats = j->stats;
	delete j;
}

void process_seqs(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                  int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0)
{
	process_seqs_end(process_seqs_begin(opt, bwt, bns, pac, n_processed, n, seqs, pes0, nullptr, nullptr, 2, false, nullptr), nullptr);
}

} // namespace b200
