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
// How the device stages of the chunks in flight share the GPU (B200_TURN):
//   0 (default)  no serialisation: every chunk job drives its own stream and the kernels of up to B200_INFLIGHT chunks share the
//                SMs.  The stages are a mix of latency-bound kernels (seeding, SA look-up: half of the issue slots idle) and
//                launches too small to fill 148 SMs (late extension rounds, rescue, CIGAR retries); interleaved, they fill each
//                other's holes - measured 8.6 M pairs/s end to end against 7.7 M with one chunk on the device at a time.
//   1            one turn for everything, handed to the oldest waiting chunk (round 1's behaviour: lowest latency per chunk)
//   2            two turns: FM-index stages (seeding, SA, chaining) and DP / finish stages - measured 7.8 M
static DeviceTurn g_gpu_turn[2];
static const int g_turn_mode = getenv("B200_TURN") ? atoi(getenv("B200_TURN")) : 0;
enum { TURN_FM = 0, TURN_DP = 1 };
struct DeviceTurnGuard {
	uint64_t t; int k;
	DeviceTurnGuard(uint64_t t_, int kind) : t(t_), k(g_turn_mode == 2 ? kind : 0) { if (g_turn_mode) g_gpu_turn[k].lock(t); }
	~DeviceTurnGuard() { if (g_turn_mode) g_gpu_turn[k].unlock(); }
};

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
static const int MAX_LANES = 1, N_SLOTS = 8;
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
struct Slot { bool busy = false; const void *staged_key = nullptr; int staged_n = 0; int64_t staged_bases = 0; };
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
	if (g_turn_mode) { g_gpu_turn[0].lock(~(uint64_t)0); if (g_turn_mode == 2) g_gpu_turn[1].lock(~(uint64_t)0); }
	return g_aux;
}
void aux_release()
{
	if (g_turn_mode) { if (g_turn_mode == 2) g_gpu_turn[1].unlock(); g_gpu_turn[0].unlock(); }
	g_aux_mu.unlock();
}
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

/* ------------------------------------------------------------------ the hot path */

// One chunk = one engine (stream, scratch, resident reads; the index is shared): every kernel sees the whole chunk.
struct Lane { Engine *eng; int r0, n; };

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
	// names, qualities and comments for the SAM text: gathered into one page-locked buffer at prefix offsets
	ReadText *rt = (ReadText *)stage_pinned(L.eng, PIN_RTEXT, sizeof(ReadText) * (n + 1));
	std::vector<int64_t> toff(n + 1);
	toff[0] = 0;
	for (int i = 0; i < n; ++i) {
		const bseq1_t &q = seqs[i];
		const int64_t nl = q.name ? (int64_t)strlen(q.name) : 0, cl = q.comment ? (int64_t)strlen(q.comment) : 0, ql = q.qual ? q.l_seq : 0;
		rt[i].name_len = (int32_t)nl; rt[i].comment_len = (int32_t)cl;
		rt[i].name_off = toff[i]; rt[i].qual_off = q.qual ? toff[i] + nl : -1; rt[i].comment_off = q.comment ? toff[i] + nl + ql : -1;
		toff[i + 1] = toff[i] + nl + ql + cl;
	}
	char *text = (char *)stage_pinned(L.eng, PIN_TEXT, (size_t)toff[n] + 16);
	parallel_for(nt, n, 4096, [&](int, int64_t b, int64_t e) {
		for (int64_t i = b; i < e; ++i) {
			const bseq1_t &q = seqs[i];
			char *d = text + toff[i];
			if (rt[i].name_len) memcpy(d, q.name, rt[i].name_len);
			if (q.qual) memcpy(d + rt[i].name_len, q.qual, q.l_seq);
			if (rt[i].comment_len) memcpy(d + rt[i].name_len + (q.qual ? q.l_seq : 0), q.comment, rt[i].comment_len);
		}
	});
	stage_upload_text(L.eng, n, rt, text, toff[n]);
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
	const Lane L = { engine_lane(slot, 0), 0, n };
	const int64_t bases = stage_lane_reads(opt, L, seqs);
	std::lock_guard<std::mutex> lk(g_slot_mu);
	g_slots[slot].busy = false; g_slots[slot].staged_key = (const void *)seqs; g_slots[slot].staged_n = n; g_slots[slot].staged_bases = bases;
	g_slot_cv.notify_all();
}

#define GPU_STAGE(kind, call) do { DeviceTurnGuard gpu_lk(ticket, kind); call; } while (0)

// where the SAM text of the chunk goes: seqs[i].sam (mem_process_seqs' contract: one malloc()ed string per read) or ONE buffer
// from the recycling pool of b200_big_alloc (b200_align_chunk / _fastq: the caller wants the chunk's text, not 667 k strings)
struct SamDest {
	bool one_buffer = false; char *sam = nullptr; int64_t sam_len = 0;
	int route = 0; b200_sam_line_t *lines = nullptr; int64_t n_lines = 0;     // (one-buffer jobs, B200_ROUTE_*: the per-line table, malloc()ed)
	int64_t *dest_off = nullptr; int n_dest = 0;                              // (B200_ROUTE_BY_CONTIG: sam is grouped by destination; n_dest + 1 offsets, malloc()ed)
};
// the chunk as raw fastq bytes (b200_align_fastq_begin): parsed on the device; seqs == null then
struct FastqSrc { char *fq[2]; int64_t len[2]; };
extern "C" int64_t b200_fastq_parse(char *buf, int64_t len, bseq1_t **out);
extern "C" bseq1_t *b200_chunk_seqs(int64_t n, const bseq1_t *s1, const bseq1_t *s2);

extern "C" void *b200_big_alloc(size_t bytes);

static void process_seqs_slot(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                              int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0,
                              int slot, bool staged, int64_t staged_bases, b200_stats_t *stats_out, SamDest *dest, uint64_t ticket,
                              const FastqSrc *fq = nullptr, int64_t *n_out = nullptr)
{
	engine_for(bwt, bns, pac);
	Engine *eng = engine_lane(slot, 0);
	Stats &st = engine_stats(eng);
	memset(static_cast<b200_stats_t *>(&st), 0, sizeof(b200_stats_t));
	const int nt = opt->n_threads > 0 ? opt->n_threads : 1;
	const double t_start = now_ms();
	const int64_t l_pac = bns->l_pac;
	double t0 = now_ms(), t1;

	// ---- the reads: raw fastq bytes parsed, interleaved and encoded on the device, or bseq1_t records encoded here (reference
	// src/bwamem.c:1057-1058) and uploaded with their text - unless b200_stage_reads() already did.  (Own stream and buffers: no
	// need to hold the device.)
	bseq1_t *own_seqs = nullptr;
	int max_len = -1;
	if (fq) {
		FastqInfo fi;
		stage_upload_fastq(eng, fq->fq[0], fq->len[0], fq->fq[1], fq->len[1], &fi);
		n = fi.n_reads; st.n_bases = fi.n_bases; max_len = fi.max_len;
		if (max_len >= opt->min_seed_len && seed_sw_filter_applies(opt, max_len)) {
			// reads long enough for mem_flt_chained_seeds take the host chaining, which works on bseq1_t records: parse here after all
			bseq1_t *m[2] = { nullptr, nullptr };
			const int64_t n1 = b200_fastq_parse(fq->fq[0], fq->len[0], &m[0]);
			if (fq->fq[1]) b200_fastq_parse(fq->fq[1], fq->len[1], &m[1]);
			seqs = own_seqs = b200_chunk_seqs(n1, m[0], m[1]);
			free(m[0]); free(m[1]);
			fq = nullptr;
		}
		if (n_out) *n_out = n;
	}
	const Lane L = { eng, 0, n };
	if (!fq) {
		if (!staged) st.n_bases = stage_lane_reads(opt, L, seqs);
		else st.n_bases = staged_bases;
	}
	st.n_reads = n;
	t1 = now_ms(); st.ms_upload = t1 - t0; t0 = t1;

	// ---- chaining on the device unless a read is long enough for mem_flt_chained_seeds (B200_CHAIN=host forces the host path)
	bool dev_chain = !(getenv("B200_CHAIN") && !strcmp(getenv("B200_CHAIN"), "host"));
	{
		std::vector<int8_t> len_rule(4096, 0);
		for (int i = 0; i < n && dev_chain && seqs; ++i) {
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
	GPU_STAGE(TURN_FM, stage_seed(eng, make_seed_opt(opt), sd, dev_chain && !chain_check));
	t1 = now_ms(); st.ms_seed = t1 - t0; t0 = t1;
	st.n_seeds = sd.n_seeds;

	ExtIn xin;
	std::vector<int32_t> chk_co, chk_srt;
	std::vector<DChain> chk_ch;
	std::vector<DSeed> chk_se;
	if (dev_chain) {
		GPU_STAGE(TURN_FM, stage_chain(eng, make_chain_opt(opt), xin, chain_check));
		st.n_chains = xin.n_chains;
		if (chain_check) {
			chk_co.assign(xin.chain_off, xin.chain_off + n + 1); chk_ch.assign(xin.chains, xin.chains + xin.n_chains);
			chk_se.assign(xin.seeds, xin.seeds + xin.n_seeds); chk_srt.assign(xin.srt, xin.srt + xin.n_seeds);
		}
	}
	if (!dev_chain || chain_check) {
	// ---- chaining + chain filtering on host threads (reads of >= ~730 bases: mem_flt_chained_seeds applies; and the checking mode)
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
			GPU_STAGE(TURN_DP, stage_sw(eng, make_sw_opt(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins), jobs, res));
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

	// ---- chain2aln / ksw_extend2 on the device; the regions stay there
	GPU_STAGE(TURN_DP, stage_extend(eng, make_ext_opt(opt), xin));
	t1 = now_ms(); st.ms_extend = t1 - t0; t0 = t1;

	// ---- everything else of mem_process_seqs (reference src/bwamem.c:1073-1085, 1187-1203, 1226-1229): de-duplication,
	// insert-size statistics, mate rescue, primary marking, pairing, mapQ, CIGARs, NM/MD and the SAM text - finish_stage.h
	FinishArgs fa;
	fa.opt = opt; fa.pes0 = (opt->flag & MEM_F_PE) ? pes0 : nullptr; fa.n_processed = n_processed; fa.rg_id = bwa_rg_id;
	fa.want_offsets = !dest->one_buffer;
	fa.route = dest->one_buffer ? dest->route : 0;
	fa.alloc = dest->one_buffer ? b200_big_alloc : nullptr;
	GPU_STAGE(TURN_DP, stage_finish(eng, fa));
	SamChunk sc;
	stage_fetch_sam(eng, fa, sc);           // (D2H on the engine's own stream: the next chunk's kernels run meanwhile)
	t0 = now_ms();
	if (dest->one_buffer) {
		dest->sam = sc.sam; dest->sam_len = sc.bytes;
		static_assert(sizeof(b200_sam_line_t) == sizeof(SamLine), "routing table entry layout");
		if (fa.route & B200_ROUTE_BY_CONTIG) {
			dest->n_dest = sc.n_dest;
			dest->dest_off = (int64_t *)calloc((size_t)sc.n_dest + 1, sizeof(int64_t));
			if (sc.dest_off) memcpy(dest->dest_off, sc.dest_off, sizeof(int64_t) * ((size_t)sc.n_dest + 1));
		}
		else if (fa.route) {
			dest->n_lines = sc.n_lines;
			dest->lines = (b200_sam_line_t *)malloc(sizeof(b200_sam_line_t) * (size_t)(sc.n_lines + 1));
			if (sc.n_lines) memcpy(dest->lines, sc.lines, sizeof(SamLine) * (size_t)sc.n_lines);
		}
	}
	else {
		// mem_process_seqs' contract: one malloc()ed, NUL-terminated string per read (the host free()s them)
		parallel_for(nt, n, 2048, [&](int, int64_t b, int64_t e) {
			for (int64_t i = b; i < e; ++i) {
				const int64_t l = sc.sam_off[i + 1] - sc.sam_off[i];
				char *p = (char *)malloc((size_t)l + 1);
				memcpy(p, sc.sam + sc.sam_off[i], (size_t)l);
				p[l] = 0;
				seqs[i].sam = p;
			}
		});
	}
	free(own_seqs);
	st.ms_deliver += now_ms() - t0;
	st.ms_total = now_ms() - t_start;
	if (bwa_verbose >= 3)
		fprintf(stderr, "[M::%s] Processed %d reads in %.3f real sec (%s; stage walls: upload %.0f ms, seed %.0f, chain %.0f, extend %.0f, dedup %.0f, rescue %.0f, pairing %.0f, cigar %.0f, text %.0f, deliver %.0f; %lld reads re-patched with a DP, %lld extra rescue rounds, %lld of %lld CIGAR jobs rerun wider)\n",
		        "mem_process_seqs", n, st.ms_total * 1e-3, engine_kind(), st.ms_upload, st.ms_seed, st.ms_chain_host,
		        st.ms_extend, st.ms_regs_host, st.ms_rescue, st.ms_sam_plan, st.ms_global, st.ms_sam_host, st.ms_deliver,
		        (long long)st.n_patch_reads, (long long)st.n_rescue_rounds, (long long)st.n_global_rerun, (long long)st.n_global_jobs);
	if (stats_out) *stats_out = st;
	std::lock_guard<std::mutex> lk(g_slot_mu);
	g_last_stats = st;
}

/* ------------------------------------------------------------------ chunk jobs */

struct SeqJob {
	std::thread th;
	b200_stats_t stats;
	SamDest dest;
	int64_t n_reads = 0;       // (fastq jobs: known once the device has parsed the bytes)
};
int64_t job_n_reads(SeqJob *j) { return j->n_reads; }

// the chunk's SAM text of a finished one-buffer job (from b200_big_alloc: release with b200_free)
int64_t job_take_sam(SeqJob *j, char **out)
{
	*out = j->dest.sam;
	j->dest.sam = nullptr;
	return j->dest.sam_len;
}
int64_t job_take_lines(SeqJob *j, b200_sam_line_t **out)
{
	*out = j->dest.lines;
	j->dest.lines = nullptr;
	return j->dest.n_lines;
}
int job_take_dest_off(SeqJob *j, int64_t **out)
{
	*out = j->dest.dest_off;
	j->dest.dest_off = nullptr;
	return j->dest.n_dest;
}

SeqJob *process_seqs_begin(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                           int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0,
                           void (*after)(void *, SeqJob *), void *arg, bool one_buffer,
                           char *fq1, int64_t len1, char *fq2, int64_t len2, int route)
{
	engine_for(bwt, bns, pac);
	SeqJob *j = new SeqJob();
	j->dest.route = route;
	memset(&j->stats, 0, sizeof j->stats);
	int slot = -1;
	bool staged = false;
	int64_t staged_bases = 0;
	uint64_t ticket;
	{
		std::unique_lock<std::mutex> lk(g_slot_mu);
		for (int k = 0; k < N_SLOTS && seqs; ++k)
			if (!g_slots[k].busy && g_slots[k].staged_key == (const void *)seqs && g_slots[k].staged_n == n) { slot = k; staged = true; staged_bases = g_slots[k].staged_bases; }
		if (slot < 0)
			g_slot_cv.wait(lk, [&] { for (int k = 0; k < N_SLOTS; ++k) if (!g_slots[k].busy && !g_slots[k].staged_key) { slot = k; return true; } return false; });
		g_slots[slot].busy = true; g_slots[slot].staged_key = nullptr;
		ticket = g_ticket_next++;
	}
	const mem_pestat_t *pes = pes0;
	j->th = std::thread([=]() {
		static const int limit = getenv("B200_INFLIGHT") ? std::max(1, atoi(getenv("B200_INFLIGHT"))) : 4;
		const FastqSrc fqs = { { fq1, fq2 }, { len1, len2 } };
		{
			std::unique_lock<std::mutex> lk(g_slot_mu);
			g_slot_cv.wait(lk, [&] { return ticket == g_ticket_serving && g_running < limit; });
			++g_running; ++g_ticket_serving;
			g_slot_cv.notify_all();
		}
		j->dest.one_buffer = one_buffer;
		process_seqs_slot(opt, bwt, bns, pac, n_processed, n, seqs, pes, slot, staged, staged_bases, &j->stats, &j->dest, ticket,
		                  fq1 ? &fqs : nullptr, &j->n_reads);
		if (after) after(arg, j);
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
	process_seqs_end(process_seqs_begin(opt, bwt, bns, pac, n_processed, n, seqs, pes0, nullptr, nullptr, false, nullptr, 0, nullptr, 0, 0), nullptr);
}

} // namespace b200
