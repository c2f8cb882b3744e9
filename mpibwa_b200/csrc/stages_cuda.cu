// stages_cuda.cu - the device stages of the alignment core (sm_100a).  The only implementation of stages.h that is
// linked into libmpibwa_b200.so; there is no CPU execution path (every entry aborts when CUDA is unusable).
//
// Data resident in HBM for the life of the engine (uploaded once from the .map image / bwa_idx_load result):
//   occ   the reference's occ-interleaved BWT (src/bwt.h:72-78) re-blocked at upload into 32-byte occ sectors (fm_kernels.h)
//   sa    suffix-array samples (every sa_intv-th row)                                 reference src/bwt.c:86-96
//   pac   2-bit forward strand                                                         reference src/bntseq.c:224-225
//   contig offset/length/ALT tables                                                    reference src/bntseq.h:44-51
// An L2 access-policy window (persisting) is laid over the occ sectors for the seeding kernels.
//
// Per chunk: encoded reads are uploaded once (stage_upload_reads) and stay resident for seeding, chaining, extension, mate
// rescue and the CIGAR stage; the kernels live in smem_sweeps.cuh / smem_kernel.cuh (seeding), chain_kernels.h (chaining),
// ext_rounds.cuh (ksw_extend2), sw_warp_kernel.cuh (ksw_align2) and global_kernels.h (ksw_global2 + traceback).
#include "stages.h"
#include "util.h"
#include "ext_rounds.cuh"
#include "smem_kernel.cuh"
#include "smem_sweeps.cuh"
#include "sw_warp_kernel.cuh"
#include "finish_stage.h"
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cmath>
#include <chrono>
#include <algorithm>

namespace b200 {

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	fprintf(stderr, "[mpibwa_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
	abort(); } } while (0)

static void die(const char *msg) { fprintf(stderr, "[mpibwa_b200] %s\n", msg); abort(); }

// Growable device buffer of an engine.  Growth is stream-ordered (cudaMallocAsync / cudaFreeAsync on the engine's stream, pool
// memory kept cached): a chunk that needs a larger scratch buffer does not synchronise the device under the other chunks in flight,
// which cudaFree would.  Every buffer registers with the engine under construction (EngineBufs), which binds the stream.
struct DevBuf;
static thread_local std::vector<DevBuf *> *tl_buf_registry = nullptr;
struct DevBuf {
	void *p = nullptr; size_t cap = 0;
	cudaStream_t st = nullptr; bool bound = false;
	DevBuf() { if (tl_buf_registry) tl_buf_registry->push_back(this); }
	void *need(size_t bytes)
	{
		if (bytes > cap) {
			size_t n = bytes + (bytes >> 2) + 256;
			if (bound) {
				if (p) CK(cudaFreeAsync(p, st));
				CK(cudaMallocAsync(&p, n, st));
			} else {
				if (p) CK(cudaFree(p));
				CK(cudaMalloc(&p, n));
			}
			cap = n;
		}
		return p;
	}
	template <class T> T *as(size_t n) { return (T *)need(n * sizeof(T)); }
	void release() { if (p) { if (bound) cudaFreeAsync(p, st); else cudaFree(p); } p = nullptr; cap = 0; }
};
struct EngineBufs {         // base of Engine: constructed before the members, so that their constructors find the registry
	std::vector<DevBuf *> all_bufs;
	EngineBufs() { tl_buf_registry = &all_bufs; }
};

// growable page-locked host buffer (results that the host threads read right after a D2H copy)
struct PinBuf {
	void *p = nullptr; size_t cap = 0;
	void *need(size_t bytes, size_t keep = 0)
	{
		if (bytes > cap) {
			size_t n = bytes + (bytes >> 2) + 4096;
			void *q = nullptr;
			CK(cudaHostAlloc(&q, n, cudaHostAllocDefault));
			if (p) { if (keep) memcpy(q, p, keep); CK(cudaFreeHost(p)); }
			p = q; cap = n;
		}
		return p;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Counters { unsigned long long occ_blocks, sa_steps, ext_cells, ext_calls, sw_cells, global_cells; };

class Engine : public EngineBufs {
public:
	Engine() { tl_buf_registry = nullptr; }
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_sync = nullptr;   // ev_sync: blocking-sync event (host threads sleep instead of spinning)
	FmView fm;                     // device pointers
	void *d_bwt = nullptr, *d_sa = nullptr, *d_pac = nullptr, *d_ctg_off = nullptr, *d_ctg_len = nullptr, *d_ctg_alt = nullptr;
	bool seeds_resident = false;   // the seed list of the whole batch is still in b_seeds / b_seedoff / b_lrep
	int64_t n_seeds_resident = 0;
	bool owns_index = true;        // false for clones made by engine_clone()
	size_t bwt_bytes = 0;
	Stats stats;
	// resident reads of the current chunk
	int n_reads = 0, max_len = 0;
	std::vector<int64_t> h_off;
	DevBuf d_off, d_codes;
	// scratch
	DevBuf b_strips, b_nfirst, b_nsweeps, b_pk;
	DevBuf b_chscr, b_chnodes, b_chnc, b_chns, b_chcoff, b_chsoff;
	DevBuf b_xrec; int64_t n_rec = 0;      // B200_EXT_RECORD: every ksw_extend2 job of the last stage_extend call (for the one-batch replay)
	DevBuf b_intv, b_scr, b_nintv, b_ioff, b_civ, b_slots, b_soff, b_seeds, b_lrep, b_seedoff, b_cub, b_wide;
	DevBuf b_chain_off, b_chains, b_dseeds, b_srt, b_regs, b_nregs, b_eh;
	DevBuf b_jobs, b_res, b_h, b_e, b_b, b_q, b_t;
	DevBuf b_xstate, b_xjobs, b_xact0, b_xact1, b_xkey, b_xkey2, b_xord, b_xctr, b_xout;
	PinBuf h_seeds, h_seed_off, h_lrep, h_codes, h_slot[PIN_N_SLOTS];
	DevBuf b_grow;
	// finish stages (finish_stage.h): scratch by FinBuf id, the read text, contig names, the log table, the SAM text on the host
	DevBuf fb[FB_N], d_rtext, d_text;
	void *d_ctg_name_off = nullptr, *d_ctg_names = nullptr, *d_ctg_anno_off = nullptr, *d_ctg_annos = nullptr, *d_logtab = nullptr, *d_ktab = nullptr, *d_sa5 = nullptr, *d_isa5 = nullptr, *d_bloom = nullptr;
	int n_log = 0;
	PinBuf h_sam, h_sam_off, h_lines, h_dest_off;
	FinishOut fin_out;
	double ms_task = 0, ms_task_text = 0;       // CUDA-event time of the finish-stage kernels of the current call
	std::vector<cudaEvent_t> ev_pool;
	static const int N_SIDE = 8;
	cudaStream_t side[N_SIDE];
	cudaEvent_t ev_fork = nullptr, ev_join[N_SIDE];
	Counters *d_cnt = nullptr;
	// per-device launch configuration, set up once per engine (kernel attributes and occupancy belong to the device)
	int n_sm = 0, seed_blocks_per_sm = 0, bwd_blocks_per_sm = 0, fwd_blocks_per_sm = 0;
	bool ext_attr_set = false, global_attr_set = false;

	void tic() { CK(cudaEventRecord(ev0, stream)); }
	double toc()
	{
		float ms = 0;
		CK(cudaEventRecord(ev1, stream));
		CK(cudaEventSynchronize(ev1));
		CK(cudaEventElapsedTime(&ms, ev0, ev1));
		return ms;
	}
	void h2d(void *dst, const void *src, size_t n)
	{
		if (!n) return;
		CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, stream));
		stats.h2d_bytes += (int64_t)n;
	}
	void d2h(void *dst, const void *src, size_t n)
	{
		if (!n) return;
		CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, stream));
		stats.d2h_bytes += (int64_t)n;
	}
	// waits for the stream without spinning: with several chunk jobs per process and one process per GPU the job threads outnumber
	// the cores a rank has, and a spinning waiter takes the core the next chunk's launches need
	void sync() { CK(cudaEventRecord(ev_sync, stream)); CK(cudaEventSynchronize(ev_sync)); }
	void zero_counters() { CK(cudaMemsetAsync(d_cnt, 0, sizeof(Counters), stream)); }
	Counters read_counters()
	{
		Counters c;
		CK(cudaMemcpyAsync(&c, d_cnt, sizeof c, cudaMemcpyDeviceToHost, stream));
		sync();
		return c;
	}
};

int engine_device_count()
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}
const char *engine_kind() { return "cuda"; }
Stats &engine_stats(Engine *e) { return e->stats; }

static void engine_set_l2_window(Engine *e);
static void engine_make_streams(Engine *e);

// re-blocks the reference's occ-interleaved BWT into occ sectors (fm_kernels.h), one sector per thread
__global__ void k_occ_convert(const uint32_t *__restrict__ ref_bwt, uint64_t n_sec, uint32_t *occ)
{
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_sec) return;
	uint32_t w[8];
	occ_convert_block(ref_bwt, b, w);
	uint4 *o = reinterpret_cast<uint4 *>(occ + (b << 3));
	o[0] = make_uint4(w[0], w[1], w[2], w[3]);
	o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// level L of the k-mer interval tables from level L-1 (smem_kernel.cuh), one entry per thread
__global__ void k_ktab_level(FmView fm, int L, Q4 *tab)
{
	const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >> (2 * L)) return;
	tab[ktab_off(L) + idx] = ktab_make(fm, L, (uint32_t)idx);
}

// the whole suffix array from its samples (fm_kernels.h), one sampled row per thread
__global__ void k_sa5_expand(FmView fm, uint64_t n_sa, uint8_t *sa5, uint8_t *isa5)
{
	const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j < n_sa) sa5_expand(fm, j, sa5, isa5);
}

// Bloom filter over the text's K-mers (fm_kernels.h), one text position per thread
__global__ void k_bloom_build(const uint8_t *__restrict__ pac, int64_t l_pac, int64_t n_pos, int K, uint64_t mask, unsigned long long *bloom)
{
	const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_pos) return;
	uint64_t word, bits;
	bloom_of_text(pac, l_pac, p, K, mask, word, bits);
	const unsigned long long old = atomicOr(bloom + 2 * word, (unsigned long long)bits);
	if ((old & bits) == bits) atomicOr(bloom + 2 * word + 1, (unsigned long long)bits);     // seen before (or looks like it): "more than once"
}

Engine *engine_create(const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int device)
{
	int nd = engine_device_count();
	if (nd <= 0) die("no usable CUDA device: the alignment core has no CPU path (sm_100a kernels only)");
	if (device < 0 || device >= nd) die("requested CUDA device does not exist");
	Engine *e = new Engine();
	e->device = device;
	CK(cudaSetDevice(device));
	engine_make_streams(e);
	const size_t ref_bytes = (size_t)bwt->bwt_size * 4;
	const uint64_t n_sec = (bwt->seq_len >> 6) + 2;
	e->bwt_bytes = (size_t)n_sec * 32;
	size_t sa_bytes = (size_t)bwt->n_sa * 8, pac_bytes = (size_t)(bns->l_pac / 4 + 1);
	CK(cudaMalloc(&e->d_bwt, e->bwt_bytes));
	CK(cudaMalloc(&e->d_sa, sa_bytes));
	CK(cudaMalloc(&e->d_pac, pac_bytes + 16));
	CK(cudaMalloc(&e->d_ctg_off, sizeof(int64_t) * bns->n_seqs));
	CK(cudaMalloc(&e->d_ctg_len, sizeof(int32_t) * bns->n_seqs));
	{	// upload the reference layout to a temporary, re-block it into occ sectors on the device, drop the temporary
		const size_t padded = ((n_sec >> 1) + 2) * 64;
		void *d_ref = nullptr;
		CK(cudaMalloc(&d_ref, padded));
		CK(cudaMemset(d_ref, 0, padded));
		CK(cudaMemcpy(d_ref, bwt->bwt, ref_bytes, cudaMemcpyHostToDevice));
		k_occ_convert<<<(unsigned)((n_sec + 255) / 256), 256>>>((const uint32_t *)d_ref, n_sec, (uint32_t *)e->d_bwt);
		CK(cudaGetLastError());
		CK(cudaDeviceSynchronize());
		CK(cudaFree(d_ref));
	}
	CK(cudaMemcpy(e->d_sa, bwt->sa, sa_bytes, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(e->d_pac, pac, pac_bytes, cudaMemcpyHostToDevice));
	CK(cudaMemset((char *)e->d_pac + pac_bytes, 0, 16));
	std::vector<int64_t> co(bns->n_seqs);
	std::vector<int32_t> cl(bns->n_seqs);
	for (int i = 0; i < bns->n_seqs; ++i) { co[i] = bns->anns[i].offset; cl[i] = bns->anns[i].len; }
	CK(cudaMemcpy(e->d_ctg_off, co.data(), co.size() * 8, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(e->d_ctg_len, cl.data(), cl.size() * 4, cudaMemcpyHostToDevice));
	std::vector<uint8_t> alt(bns->n_seqs + 1, 0);
	for (int i = 0; i < bns->n_seqs; ++i) alt[i] = bns->anns[i].is_alt ? 1 : 0;
	CK(cudaMalloc(&e->d_ctg_alt, alt.size()));
	CK(cudaMemcpy(e->d_ctg_alt, alt.data(), alt.size(), cudaMemcpyHostToDevice));
	{	// contig names and annotations (SAM text), back to back with offset tables
		std::vector<int64_t> no(bns->n_seqs + 1, 0), ao(bns->n_seqs + 1, 0);
		std::string names, annos;
		for (int i = 0; i < bns->n_seqs; ++i) {
			names += bns->anns[i].name; no[i + 1] = (int64_t)names.size();
			if (bns->anns[i].anno) annos += bns->anns[i].anno;
			ao[i + 1] = (int64_t)annos.size();
		}
		CK(cudaMalloc(&e->d_ctg_name_off, no.size() * 8)); CK(cudaMalloc(&e->d_ctg_names, names.size() + 8));
		CK(cudaMalloc(&e->d_ctg_anno_off, ao.size() * 8)); CK(cudaMalloc(&e->d_ctg_annos, annos.size() + 8));
		CK(cudaMemcpy(e->d_ctg_name_off, no.data(), no.size() * 8, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(e->d_ctg_names, names.data(), names.size(), cudaMemcpyHostToDevice));
		CK(cudaMemcpy(e->d_ctg_anno_off, ao.data(), ao.size() * 8, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(e->d_ctg_annos, annos.data(), annos.size(), cudaMemcpyHostToDevice));
	}
	{	// log(i) by glibc for every integer the mapQ formulas can ask for (finish_kernels.h: bit-exactness)
		e->n_log = 1 << 20;
		std::vector<double> lt(e->n_log);
		for (int i = 0; i < e->n_log; ++i) lt[i] = log((double)i);
		CK(cudaMalloc(&e->d_logtab, sizeof(double) * e->n_log));
		CK(cudaMemcpy(e->d_logtab, lt.data(), sizeof(double) * e->n_log, cudaMemcpyHostToDevice));
	}
	FmView &fm = e->fm;
	fm.occ = (const uint32_t *)e->d_bwt; fm.sa = (const uint64_t *)e->d_sa;
	fm.primary = bwt->primary;
	for (int i = 0; i < 5; ++i) fm.L2[i] = bwt->L2[i];
	fm.seq_len = bwt->seq_len; fm.sa_intv = bwt->sa_intv;
	fm.pac = (const uint8_t *)e->d_pac; fm.l_pac = bns->l_pac;
	fm.ctg_off = (const int64_t *)e->d_ctg_off; fm.ctg_len = (const int32_t *)e->d_ctg_len; fm.n_ctg = bns->n_seqs;
	if (fm.sa_intv & (fm.sa_intv - 1)) die("suffix-array sampling interval must be a power of two");
	if (fm.seq_len >> 33) die("references beyond 2^33 BWT symbols (4.29 Gbp) are not supported by the packed seeding lists");
	/* Derived index structures, in the order of what they buy per byte: k-mer interval tables, the whole suffix
	 * array, its inverse, the Bloom filters.  Each is allocated only if it leaves `reserve` of the device's memory free for the chunk
	 * slots (B200_TABLE_RESERVE_GB, default 72 GB or 40 % of a smaller device): a table that does not fit is skipped (the k-mer
	 * tables: made shallower) and the kernels take the path without it. */
	size_t total_b = 0;
	{ size_t f = 0; CK(cudaMemGetInfo(&f, &total_b)); }
	const size_t reserve = std::min<size_t>((size_t)((getenv("B200_TABLE_RESERVE_GB") ? atof(getenv("B200_TABLE_RESERVE_GB")) : 72.0) * 1e9), (size_t)(total_b * 0.4));
	auto fits = [&](size_t bytes) { size_t f = 0, t = 0; CK(cudaMemGetInfo(&f, &t)); return bytes + reserve <= f; };
	auto ms_since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
	const bool dbg = getenv("B200_DEBUG") != nullptr;
	fm.ktab = nullptr; fm.kmax = 0;
	{	// k-mer interval tables: every pattern of up to kmax bases, built level by level with the seeding kernels' own extension
		// (B200_KMER_MAX: depth override, 0 = none)
		int kmax = ktab_default_kmax(fm.seq_len);
		if (getenv("B200_KMER_MAX")) kmax = std::max(0, std::min(16, atoi(getenv("B200_KMER_MAX"))));
		while (kmax > 0 && !fits(ktab_entries(kmax) * sizeof(Q4) + 64)) --kmax;
		if (kmax > 0) {
			const auto t0 = std::chrono::steady_clock::now();
			CK(cudaMalloc(&e->d_ktab, ktab_entries(kmax) * sizeof(Q4) + 64));
			fm.ktab = (const uint32_t *)e->d_ktab;
			for (int L = 1; L <= kmax; ++L) {
				const uint64_t n = (uint64_t)1 << (2 * L);
				k_ktab_level<<<(unsigned)((n + 255) / 256), 256>>>(fm, L, (Q4 *)e->d_ktab);
				CK(cudaGetLastError());
			}
			CK(cudaDeviceSynchronize());
			fm.kmax = kmax;
			if (dbg) fprintf(stderr, "[mpibwa_b200] k-mer interval tables: patterns of up to %d bases, %.2f GB, built in %.0f ms\n", kmax, ktab_entries(kmax) * sizeof(Q4) / 1e9, ms_since(t0));
		}
	}
	fm.sa5 = nullptr; fm.isa5 = nullptr;
	const int sa_full = getenv("B200_SA_FULL") ? atoi(getenv("B200_SA_FULL")) : 2;
	if (sa_full > 0) {
		// the whole suffix array and its inverse, five bytes per row / position, expanded from the samples (B200_SA_FULL=0: keep the
		// samples only, 1: no inverse)
		const size_t bytes = ((size_t)fm.seq_len + 1) * 5 + 16;
		if (fits(bytes) && (fm.seq_len >> 40) == 0 && (uint64_t)bwt->n_sa == (fm.seq_len + fm.sa_intv) / fm.sa_intv) {
			const auto t0 = std::chrono::steady_clock::now();
			CK(cudaMalloc(&e->d_sa5, bytes));
			const bool inverse = sa_full > 1 && fits(bytes);
			if (inverse) { CK(cudaMalloc(&e->d_isa5, bytes)); CK(cudaMemset(e->d_isa5, 0, bytes)); }
			k_sa5_expand<<<(unsigned)(((uint64_t)bwt->n_sa + 127) / 128), 128>>>(fm, (uint64_t)bwt->n_sa, (uint8_t *)e->d_sa5, (uint8_t *)e->d_isa5);
			CK(cudaGetLastError());
			CK(cudaDeviceSynchronize());
			fm.sa5 = (const uint8_t *)e->d_sa5; fm.isa5 = (const uint8_t *)e->d_isa5;
			if (dbg) fprintf(stderr, "[mpibwa_b200] whole suffix array%s: %.2f GB, expanded from the samples in %.0f ms\n", inverse ? " and its inverse" : "", (inverse ? 2 : 1) * bytes / 1e9, ms_since(t0));
		}
	}
	fm.bloom = nullptr; fm.bloom_mask = 0; fm.bloom_k = 0;
	if (!(getenv("B200_BLOOM") && atoi(getenv("B200_BLOOM")) == 0) && (int64_t)fm.seq_len > 64) {
		// Bloom filters over the text's 19-mers (the default min_seed_len; used whenever the caller's is at least that)
		const int K = 19;
		const uint64_t n_words = bloom_words_for(fm.seq_len);
		if (fits(n_words * 16)) {
			const auto t0 = std::chrono::steady_clock::now();
			CK(cudaMalloc(&e->d_bloom, n_words * 16));
			CK(cudaMemset(e->d_bloom, 0, n_words * 16));
			const int64_t n_pos = (int64_t)fm.seq_len - K + 1;
			k_bloom_build<<<(unsigned)((n_pos + 255) / 256), 256>>>(fm.pac, fm.l_pac, n_pos, K, n_words - 1, (unsigned long long *)e->d_bloom);
			CK(cudaGetLastError());
			CK(cudaDeviceSynchronize());
			fm.bloom = (const uint64_t *)e->d_bloom; fm.bloom_mask = n_words - 1; fm.bloom_k = K;
			if (dbg) fprintf(stderr, "[mpibwa_b200] Bloom filters over the text's %d-mers (present / more than once): %.2f GB, built in %.0f ms\n", K, n_words * 16 / 1e9, ms_since(t0));
		}
	}

	engine_set_l2_window(e);
	return e;
}

// L2 persistence window over the occ/BWT blocks (north star: "L2-persistence windows for the hot Occ blocks")
static void engine_set_l2_window(Engine *e)
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, e->device));
	// B200_L2_WINDOW: 0 = no access-policy window, 1 = persisting share + streaming misses, 2 = persisting share + normal misses
	// Default 0: measured on 100 Mbp / 1 Gbp / 3.1 Gbp references the window COSTS 10-15 % of the end-to-end rate (seeding 16.6 -> 14.1,
	// 26.1 -> 23.7, 31.4 -> 28.1 ms; CIGAR stage 4.4 -> 3.2 ms): a window over a multi-GB table marks a random few percent of its lines
	// persisting, not the hot ones, and the persisting carve-out takes L2 away from everything else (profiles/r02_summary.md).
	static const int mode = getenv("B200_L2_WINDOW") ? atoi(getenv("B200_L2_WINDOW")) : 0;
	if (mode && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
		size_t persist = (size_t)prop.persistingL2CacheMaxSize;
		CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
		cudaStreamAttrValue attr;
		memset(&attr, 0, sizeof attr);
		size_t win = std::min(e->bwt_bytes, (size_t)prop.accessPolicyMaxWindowSize);
		attr.accessPolicyWindow.base_ptr = e->d_bwt;
		attr.accessPolicyWindow.num_bytes = win;
		attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)win);
		attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		attr.accessPolicyWindow.missProp = mode == 2 ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
		CK(cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
	}
}

static void engine_make_streams(Engine *e)
{
	CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
	{	// scratch comes from the device's default memory pool and stays cached there
		cudaMemPool_t mp;
		CK(cudaDeviceGetDefaultMemPool(&mp, e->device));
		uint64_t keep = ~0ull;
		CK(cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep));
		for (DevBuf *b : e->all_bufs) { b->st = e->stream; b->bound = true; }
	}
	CK(cudaEventCreateWithFlags(&e->ev0, cudaEventBlockingSync));
	CK(cudaEventCreateWithFlags(&e->ev1, cudaEventBlockingSync));
	CK(cudaEventCreateWithFlags(&e->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
	for (int i = 0; i < Engine::N_SIDE; ++i) { CK(cudaStreamCreateWithFlags(&e->side[i], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming)); }
	memset(static_cast<b200_stats_t *>(&e->stats), 0, sizeof(b200_stats_t));
	CK(cudaMalloc(&e->d_cnt, sizeof(Counters)));
}

Engine *engine_clone(Engine *base)
{
	Engine *e = new Engine();
	e->device = base->device;
	CK(cudaSetDevice(e->device));
	engine_make_streams(e);
	e->owns_index = false;
	e->d_bwt = base->d_bwt; e->d_sa = base->d_sa; e->d_pac = base->d_pac; e->d_ctg_off = base->d_ctg_off; e->d_ctg_len = base->d_ctg_len; e->d_ctg_alt = base->d_ctg_alt;
	e->d_ctg_name_off = base->d_ctg_name_off; e->d_ctg_names = base->d_ctg_names; e->d_ctg_anno_off = base->d_ctg_anno_off; e->d_ctg_annos = base->d_ctg_annos;
	e->d_logtab = base->d_logtab; e->n_log = base->n_log; e->d_ktab = base->d_ktab; e->d_sa5 = base->d_sa5; e->d_isa5 = base->d_isa5; e->d_bloom = base->d_bloom;
	e->bwt_bytes = base->bwt_bytes;
	e->fm = base->fm;
	engine_set_l2_window(e);
	return e;
}

void engine_destroy(Engine *e)
{
	if (!e) return;
	cudaSetDevice(e->device);
	cudaStreamSynchronize(e->stream);
	for (DevBuf *b : e->all_bufs) b->release();
	cudaStreamSynchronize(e->stream);
	e->h_seeds.release(); e->h_seed_off.release(); e->h_lrep.release(); e->h_codes.release();
	for (int i = 0; i < PIN_N_SLOTS; ++i) e->h_slot[i].release();
	e->h_sam.release(); e->h_sam_off.release(); e->h_lines.release(); e->h_dest_off.release();
	if (e->owns_index) {
		cudaFree(e->d_bwt); cudaFree(e->d_sa); cudaFree(e->d_pac); cudaFree(e->d_ctg_off); cudaFree(e->d_ctg_len); cudaFree(e->d_ctg_alt);
		cudaFree(e->d_ctg_name_off); cudaFree(e->d_ctg_names); cudaFree(e->d_ctg_anno_off); cudaFree(e->d_ctg_annos); cudaFree(e->d_logtab); cudaFree(e->d_ktab); cudaFree(e->d_sa5); cudaFree(e->d_isa5); cudaFree(e->d_bloom);
	}
	cudaFree(e->d_cnt);
	cudaEventDestroy(e->ev0); cudaEventDestroy(e->ev1); cudaEventDestroy(e->ev_fork); cudaEventDestroy(e->ev_sync);
	for (int i = 0; i < Engine::N_SIDE; ++i) { cudaStreamDestroy(e->side[i]); cudaEventDestroy(e->ev_join[i]); }
	for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
	cudaStreamDestroy(e->stream);
	delete e;
}

/* ------------------------------------------------------------------ small device helpers */

// asks for the 128-byte lines of the 2-bit reference that hold the window [rb, re) of the forward+reverse coordinate (one strand),
// ahead of a DP that reads one base per row: see ext_prefetch_target (ext_rounds.cuh).  first / step: this thread's share of the lines.
__device__ __forceinline__ void prefetch_ref_window(const uint8_t *__restrict__ pac, int64_t l_pac, int64_t rb, int64_t re, int first, int step)
{
	if (re <= rb) return;
	const int64_t fb = rb < l_pac ? rb : (l_pac << 1) - re, fe = rb < l_pac ? re - 1 : (l_pac << 1) - 1 - rb;
	for (int64_t a = ((fb >> 2) & ~(int64_t)127) + (int64_t)first * 128; a <= (fe >> 2); a += (int64_t)step * 128)
		asm volatile("prefetch.global.L1 [%0];" :: "l"(pac + a));
}

__device__ __forceinline__ void warp_add(unsigned long long *dst, long long v)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, (unsigned long long)v);
}

static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

/* ------------------------------------------------------------------ seeding */

// per read: order the interval list by (start,end) (the ks_introsort that closes mem_collect_intv, reference
// src/bwamem.c:160; equal keys are identical intervals, so a rank is enough), compact it, count the SA slots of every
// interval and compute l_rep (reference src/bwamem.c:261-269)
__global__ void __launch_bounds__(128) k_compact_intv(SeedOpt so, int n_reads, const Intv *__restrict__ in, int cap, const int32_t *__restrict__ n_intv,
                                                      const int64_t *__restrict__ ioff, Intv *civ, int32_t *slots, int32_t *l_rep)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int n = n_intv[r];
	const Intv *p = in + (int64_t)r * cap;
	const int64_t o = ioff[r];
	for (int i = 0; i < n; ++i) {
		const Intv v = p[i];
		int rank = 0;
		for (int j = 0; j < n; ++j) {
			const uint64_t u = p[j].info;
			rank += (u < v.info) || (u == v.info && j < i);
		}
		civ[o + rank] = v;
		slots[o + rank] = seed_slots(v.x2, so.max_occ);
	}
	if (!l_rep) return;
	int b = 0, e = 0, rep = 0;
	for (int i = 0; i < n; ++i) {
		const Intv v = civ[o + i];
		if (v.x2 <= (uint64_t)so.max_occ) continue;
		const int sb = (int)(v.info >> 32), se = (int)(uint32_t)v.info;
		if (sb > e) { rep += e - b; b = sb; e = se; }
		else e = e > se ? e : se;
	}
	rep += e - b;
	l_rep[r] = rep;
}

// one thread per suffix-array look-up
__global__ void __launch_bounds__(256) k_sa_seeds(FmView fm, SeedOpt so, int64_t n_slots, int64_t n_intv, const Intv *__restrict__ civ,
                                                   const int64_t *__restrict__ soff, SeedRec *seeds, Counters *cnt)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	long long steps_total = 0;
	if (t < n_slots) {
		int64_t lo = 0, hi = n_intv;            // largest i with soff[i] <= t
		while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (soff[mid] <= t) lo = mid; else hi = mid; }
		Intv p = civ[lo];
		int c = (int)(t - soff[lo]);
		uint64_t step = seed_step(p.x2, so.max_occ);
		int steps;
		SeedRec s;
		s.rbeg = (int64_t)fm_sa(fm, p.x0 + (uint64_t)c * step, &steps);
		s.qbeg = (uint16_t)(p.info >> 32);
		s.len = (uint16_t)((uint32_t)p.info - (uint32_t)(p.info >> 32));
		s.rid = fm_intv2rid(fm, s.rbeg, s.rbeg + s.len);
		seeds[t] = s;
		steps_total = steps;
	}
	warp_add(&cnt->sa_steps, steps_total);
}

__global__ void k_gather_seed_off(int n_reads, const int64_t *__restrict__ ioff, const int64_t *__restrict__ soff, int64_t *seed_off, int64_t base)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r <= n_reads) seed_off[r] = base + soff[ioff[r]];
}

__global__ void k_sa_plain(FmView fm, int64_t n, const uint64_t *__restrict__ k, uint64_t *sa)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) sa[t] = fm_sa(fm, k[t], nullptr);
}

__global__ void k_fm_extend1(FmView fm, Intv ik, Intv *ok, int is_back)
{
	Intv o[4];
	fm_extend(fm, ik, o, is_back, nullptr);
	for (int i = 0; i < 4; ++i) ok[i] = o[i];
}

__global__ void k_widen(int64_t n, const int32_t *__restrict__ in, int64_t *out)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) out[t] = in[t];
}

// exclusive prefix sum of n int32 counts into int64 offsets (cub::DeviceScan on a widened copy)
static void exclusive_scan(Engine *e, const int32_t *in, int64_t *out, int64_t n)
{
	int64_t *wide = e->b_wide.as<int64_t>(n);
	k_widen<<<grid_for(n, 256), 256, 0, e->stream>>>(n, in, wide);
	CK(cudaGetLastError());
	size_t tmp = 0;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, wide, out, n, e->stream));
	void *d = e->b_cub.need(tmp);
	CK(cub::DeviceScan::ExclusiveSum(d, tmp, wide, out, n, e->stream));
	e->stats.n_launches += 2;
}

uint8_t *stage_read_buffer(Engine *e, int64_t bytes)
{
	CK(cudaSetDevice(e->device));
	return (uint8_t *)e->h_codes.need((size_t)bytes);
}

void stage_upload_reads(Engine *e, int n_reads, const int64_t *off, const uint8_t *codes)
{
	CK(cudaSetDevice(e->device));
	e->n_reads = n_reads;
	e->h_off.assign(off, off + n_reads + 1);
	e->max_len = 0;
	for (int i = 0; i < n_reads; ++i) e->max_len = std::max(e->max_len, (int)(off[i + 1] - off[i]));
	e->h2d(e->d_off.as<int64_t>(n_reads + 1), off, sizeof(int64_t) * (n_reads + 1));
	e->h2d(e->d_codes.as<uint8_t>(off[n_reads] + 16), codes, (size_t)off[n_reads]);
	e->sync();
}

// steps A-C for reads [r0, r1): leaves the compacted interval list in b_civ, slot counts in b_slots, offsets in b_ioff
static int64_t run_collect(Engine *e, const SeedOpt &so, int r0, int r1, const int64_t *d_off, const uint8_t *d_codes, int max_len,
                           bool want_lrep)
{
	const int n = r1 - r0;
	int &blocks_per_sm = e->seed_blocks_per_sm, &n_sm = e->n_sm;
	int &bwd_blocks_per_sm = e->bwd_blocks_per_sm, &fwd_blocks_per_sm = e->fwd_blocks_per_sm;
	// (tuning knobs, tools/tune_env.py: entries of an interval list kept in shared memory; resident blocks the sweeps are compiled for)
	static const int quota_env = getenv("B200_SEED_QUOTA") ? atoi(getenv("B200_SEED_QUOTA")) : 16;
	static const int fwd_minb = getenv("B200_FWD_MINB") ? atoi(getenv("B200_FWD_MINB")) : 9;
	static const int bwd_minb = getenv("B200_BWD_MINB") ? atoi(getenv("B200_BWD_MINB")) : 6;
	const int threads = 128, quota = quota_env;
	const size_t sh_bytes = (size_t)threads * quota * 16;               // interval lists
	if (!blocks_per_sm) {
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, e->device));
		n_sm = prop.multiProcessorCount;
		CK(cudaFuncSetAttribute(k_seed_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_bytes));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_seed_lanes, threads, sh_bytes));
		if (blocks_per_sm < 1) die("k_seed_lanes does not fit an SM");
	}
	const int grid = std::min(n_sm * blocks_per_sm, grid_for(n, threads));
	int cap = std::min(max_len + 32, 64);
	int32_t *n_intv = e->b_nintv.as<int32_t>(n + 1);
	int *ctr = e->b_xctr.as<int>(16);
	const int spill_per = std::max(0, max_len + 2 - quota);
	Q4 *spill = e->b_scr.as<Q4>((size_t)grid * threads * spill_per + 1);
	// throughput path: the four homogeneous sweeps of smem_sweeps.cuh; the general state machine redoes the (rare) reads
	// whose sweep strip overflowed.  B200_SEED_KERNEL=lanes forces the general kernel for every read (parity tests).
	const bool use_sweeps = !(getenv("B200_SEED_KERNEL") && !strcmp(getenv("B200_SEED_KERNEL"), "lanes"));
	typedef void (*FwdK)(SweepArgs);
	typedef void (*BwdK)(SweepArgs);
	const FwdK fwd1 = fwd_minb >= 12 ? k_sweep_fwd<1, 12> : fwd_minb >= 9 ? k_sweep_fwd<1, 9> : fwd_minb >= 8 ? k_sweep_fwd<1, 8> : k_sweep_fwd<1, 6>;
	const FwdK fwd2 = fwd_minb >= 12 ? k_sweep_fwd<2, 12> : fwd_minb >= 9 ? k_sweep_fwd<2, 9> : fwd_minb >= 8 ? k_sweep_fwd<2, 8> : k_sweep_fwd<2, 6>;
	const BwdK bwd = bwd_minb >= 12 ? k_sweep_bwd<12> : bwd_minb >= 9 ? k_sweep_bwd<9> : k_sweep_bwd<6>;
	if (use_sweeps && !bwd_blocks_per_sm) {
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bwd_blocks_per_sm, bwd, threads, 0));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fwd_blocks_per_sm, fwd1, threads, 0));
		if (bwd_blocks_per_sm < 1 || fwd_blocks_per_sm < 1) die("sweep kernels do not fit an SM");
		if (getenv("B200_DEBUG")) fprintf(stderr, "[seed] sweeps: %d forward, %d backward blocks per SM (quota %d)\n", fwd_blocks_per_sm, bwd_blocks_per_sm, quota);
	}
	const int strip_cap = getenv("B200_SEED_STRIP") ? atoi(getenv("B200_SEED_STRIP")) : 3 * max_len + 8;   // (override: tests force overflows)
	for (;;) {
		Intv *out = e->b_intv.as<Intv>((size_t)n * cap);
		CK(cudaMemsetAsync(ctr, 0, 8 * sizeof(int), e->stream));
		int h[8];
		if (use_sweeps) {
			SweepArgs a;
			a.fm = e->fm; a.so = so; a.n_reads = n; a.off = d_off + r0; a.codes = d_codes; a.out = out; a.cap = cap;
			a.strips = e->b_strips.as<Q4>((size_t)n * strip_cap + 1); a.strip_cap = strip_cap;
			a.pk_stride = packed_words_for(max_len);
			uint64_t *pk = e->b_pk.as<uint64_t>((size_t)n * a.pk_stride * 2 + 2);
			a.pk = pk;
			k_pack_reads<<<grid_for((int64_t)n * a.pk_stride, 256), 256, 0, e->stream>>>(n, d_off + r0, d_codes, a.pk_stride, pk);
			e->stats.n_launches += 1;
			a.n_intv = n_intv; a.n_first = e->b_nfirst.as<int32_t>(n + 1); a.n_sweeps = e->b_nsweeps.as<int32_t>(n + 1);
			a.worst = ctr + 1; a.n_over = ctr + 2; a.occ_blocks = &e->d_cnt->occ_blocks;
			// B200_SEED_FILL < 1: the persistent sweeps take only that share of the blocks an SM could hold, leaving registers and
			// shared memory for the kernels of the other chunks in flight (the sweeps are bound by random DRAM sectors, not by issue)
			static const double fill = getenv("B200_SEED_FILL") ? atof(getenv("B200_SEED_FILL")) : 1.0;
			const int fpb = std::max(1, (int)(fwd_blocks_per_sm * fill + .5)), bpb = std::max(1, (int)(bwd_blocks_per_sm * fill + .5));
			const int gf = std::min(n_sm * fpb, grid_for(n, threads)), gb = std::min(n_sm * bpb, grid_for(n, threads));
			a.next_read = ctr + 3; fwd1<<<gf, threads, 0, e->stream>>>(a);
			a.next_read = ctr + 4; bwd<<<gb, threads, 0, e->stream>>>(a);
			a.next_read = ctr + 5; fwd2<<<gf, threads, 0, e->stream>>>(a);
			a.next_read = ctr + 6; bwd<<<gb, threads, 0, e->stream>>>(a);
			CK(cudaGetLastError());
			e->stats.n_launches += 4;
			CK(cudaMemcpyAsync(h, ctr, sizeof h, cudaMemcpyDeviceToHost, e->stream));
			e->sync();
			if (h[2] > 0) {                     // strip overflow / chain budget: those reads go through the general state machine
				if (getenv("B200_DEBUG")) fprintf(stderr, "[seed] %d of %d reads redone by the general kernel\n", h[2], n);
				k_seed_lanes<<<grid, threads, sh_bytes, e->stream>>>(e->fm, so, n, d_off + r0, d_codes, out, cap, quota, spill, n_intv, ctr, ctr + 1,
					&e->d_cnt->occ_blocks, a.n_sweeps);
				CK(cudaGetLastError());
				e->stats.n_launches += 1;
				CK(cudaMemcpyAsync(h, ctr, sizeof h, cudaMemcpyDeviceToHost, e->stream));
				e->sync();
			}
		} else {
			k_seed_lanes<<<grid, threads, sh_bytes, e->stream>>>(e->fm, so, n, d_off + r0, d_codes, out, cap, quota, spill, n_intv, ctr, ctr + 1,
				&e->d_cnt->occ_blocks, nullptr);
			CK(cudaGetLastError());
			e->stats.n_launches += 1;
			CK(cudaMemcpyAsync(h, ctr, sizeof h, cudaMemcpyDeviceToHost, e->stream));
			e->sync();
		}
		if (h[1] <= cap) break;                 // no read needed more than cap intervals
		cap = h[1] + 8;                         // rare (very repetitive reads): run the sub-batch again with room for the worst
	}
	CK(cudaMemsetAsync(n_intv + n, 0, sizeof(int32_t), e->stream));
	int64_t *ioff = e->b_ioff.as<int64_t>(n + 1);
	exclusive_scan(e, n_intv, ioff, n + 1);
	int64_t total = 0;
	e->d2h(&total, ioff + n, sizeof(int64_t));
	e->sync();
	Intv *civ = e->b_civ.as<Intv>(total + 1);
	int32_t *slots = e->b_slots.as<int32_t>(total + 1);
	int32_t *lrep = want_lrep ? e->b_lrep.as<int32_t>(n) : nullptr;
	k_compact_intv<<<grid_for(n, 128), 128, 0, e->stream>>>(so, n, (const Intv *)e->b_intv.p, cap, n_intv, ioff, civ, slots, lrep);
	CK(cudaGetLastError());
	e->stats.n_launches += 1;
	return total;
}

static int seed_sub_batch(int max_len)
{
	// output strip per read: up to len+32 intervals of 32 bytes (64 in the common case); keep one sub-batch under ~12 GB
	size_t per = (size_t)(max_len + 40) * sizeof(Intv);
	size_t n = ((size_t)12 << 30) / per;
	return (int)std::max<size_t>(1024, std::min<size_t>(n, 1 << 20));
}

void stage_collect_intv(Engine *e, const SeedOpt &so, int n_reads, const int64_t *off, const uint8_t *codes,
                        std::vector<int64_t> &intv_off, std::vector<Intv> &intv)
{
	stage_upload_reads(e, n_reads, off, codes);
	intv_off.assign(n_reads + 1, 0);
	intv.clear();
	e->zero_counters();
	const int sub = seed_sub_batch(e->max_len);
	for (int r0 = 0; r0 < n_reads; r0 += sub) {
		int r1 = std::min(n_reads, r0 + sub), n = r1 - r0;
		int64_t total = run_collect(e, so, r0, r1, (const int64_t *)e->d_off.p, (const uint8_t *)e->d_codes.p, e->max_len, false);
		std::vector<int64_t> io(n + 1);
		e->d2h(io.data(), e->b_ioff.p, sizeof(int64_t) * (n + 1));
		size_t base = intv.size();
		intv.resize(base + total);
		e->d2h(intv.data() + base, e->b_civ.p, sizeof(Intv) * total);
		e->sync();
		for (int i = 0; i <= n; ++i) intv_off[r0 + i] = (int64_t)base + io[i];
	}
	Counters c = e->read_counters();
	e->stats.fm_occ_blocks += (int64_t)c.occ_blocks;
	e->stats.n_intv += (int64_t)intv.size();
}

void stage_seed(Engine *e, const SeedOpt &so, SeedOut &out, bool keep_on_device)
{
	CK(cudaSetDevice(e->device));
	const int n_reads = e->n_reads;
	if (e->max_len > 0xffff) die("reads of 65536 bases or more are not supported by the seeding stage");
	int64_t *seed_off = (int64_t *)e->h_seed_off.need(sizeof(int64_t) * (n_reads + 1));
	int32_t *l_rep = (int32_t *)e->h_lrep.need(sizeof(int32_t) * (n_reads + 1));
	seed_off[0] = 0;
	int64_t n_seeds = 0;
	e->zero_counters();
	const int sub = seed_sub_batch(e->max_len);
	const bool resident = keep_on_device && n_reads <= sub;     // one sub-batch: its seed list can stay where it is
	double ms_smem = 0, ms_sa = 0;
	for (int r0 = 0; r0 < n_reads; r0 += sub) {
		int r1 = std::min(n_reads, r0 + sub), n = r1 - r0;
		e->tic();
		int64_t n_intv = run_collect(e, so, r0, r1, (const int64_t *)e->d_off.p, (const uint8_t *)e->d_codes.p, e->max_len, true);
		ms_smem += e->toc();
		e->stats.n_intv += n_intv;
		e->tic();
		int64_t *soff = e->b_soff.as<int64_t>(n_intv + 2);
		CK(cudaMemsetAsync((int32_t *)e->b_slots.p + n_intv, 0, sizeof(int32_t), e->stream));
		exclusive_scan(e, (const int32_t *)e->b_slots.p, soff, n_intv + 1);
		int64_t n_slots = 0;
		e->d2h(&n_slots, soff + n_intv, sizeof(int64_t));
		e->sync();
		SeedRec *d_seeds = e->b_seeds.as<SeedRec>(n_slots + 1);
		if (n_slots > 0) {
			k_sa_seeds<<<grid_for(n_slots, 256), 256, 0, e->stream>>>(e->fm, so, n_slots, n_intv, (const Intv *)e->b_civ.p, soff, d_seeds, e->d_cnt);
			CK(cudaGetLastError());
			e->stats.n_launches += 1;
		}
		int64_t *d_seed_off = e->b_seedoff.as<int64_t>(n + 1);
		k_gather_seed_off<<<grid_for(n + 1, 256), 256, 0, e->stream>>>(n, (const int64_t *)e->b_ioff.p, soff, d_seed_off, n_seeds);
		CK(cudaGetLastError());
		e->stats.n_launches += 1;
		ms_sa += e->toc();
		if (!resident) {
			SeedRec *seeds = (SeedRec *)e->h_seeds.need(sizeof(SeedRec) * (n_seeds + n_slots + 1), sizeof(SeedRec) * n_seeds);
			e->d2h(seeds + n_seeds, d_seeds, sizeof(SeedRec) * n_slots);
			e->d2h(seed_off + r0, d_seed_off, sizeof(int64_t) * (n + 1));
			e->d2h(l_rep + r0, e->b_lrep.p, sizeof(int32_t) * n);
			e->sync();
		}
		n_seeds += n_slots;
		e->stats.fm_sa_lookups += n_slots;
	}
	Counters c = e->read_counters();
	e->stats.fm_occ_blocks += (int64_t)c.occ_blocks;
	e->stats.fm_sa_steps += (int64_t)c.sa_steps;
	e->stats.ms_k_smem += ms_smem;
	e->stats.ms_k_sa += ms_sa;
	out.seed_off = seed_off; out.seeds = (const SeedRec *)e->h_seeds.p; out.l_rep = l_rep; out.n_seeds = n_seeds;
	e->seeds_resident = resident; e->n_seeds_resident = n_seeds;
	if (resident) { out.seed_off = nullptr; out.seeds = nullptr; out.l_rep = nullptr; }
}

/* ------------------------------------------------------------------ chaining (row f2) */

__global__ void __launch_bounds__(128) k_chain_build(ChainOpt co, int64_t l_pac, ChainScratch S, int n_reads, const int64_t *__restrict__ off,
                                                     const SeedRec *__restrict__ seeds, const int64_t *__restrict__ seed_off, int32_t *n_kc, int32_t *n_ks)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int64_t base = seed_off[r];
	int ns = 0;
	n_kc[r] = chain_build_filter(co, l_pac, S, r, (int)(off[r + 1] - off[r]), seeds, base, (int)(seed_off[r + 1] - base), &ns);
	n_ks[r] = ns;
}

__global__ void __launch_bounds__(128) k_chain_emit(ChainOpt co, FmView fm, ChainScratch S, int n_reads, const int64_t *__restrict__ off,
                                                    const SeedRec *__restrict__ seeds, const int64_t *__restrict__ seed_off, const int32_t *__restrict__ l_rep,
                                                    const int32_t *__restrict__ n_kc, const int64_t *__restrict__ coff, const int64_t *__restrict__ soff,
                                                    int32_t *chain_off, DChain *chains, DSeed *dseeds, int32_t *srt)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r > n_reads) return;
	chain_off[r] = (int32_t)coff[r];
	if (r == n_reads || n_kc[r] == 0) return;
	chain_emit(co, fm, S, (int)(off[r + 1] - off[r]), l_rep[r], seeds, seed_off[r], n_kc[r], coff[r], soff[r], chains, dseeds, srt);
}

void stage_chain(Engine *e, const ChainOpt &co, ExtIn &in, bool download)
{
	CK(cudaSetDevice(e->device));
	const int n = e->n_reads;
	const int64_t n_in = e->n_seeds_resident;
	if (!e->seeds_resident) {                   // several seeding sub-batches: bring the concatenated host copy back
		e->h2d(e->b_seeds.as<SeedRec>(n_in + 1), e->h_seeds.p, sizeof(SeedRec) * n_in);
		e->h2d(e->b_seedoff.as<int64_t>(n + 1), e->h_seed_off.p, sizeof(int64_t) * (n + 1));
		e->h2d(e->b_lrep.as<int32_t>(n + 1), e->h_lrep.p, sizeof(int32_t) * n);
	}
	e->tic();
	ChainScratch S;
	S.scr = e->b_chscr.as<int32_t>((size_t)CH_N_PLANES * (n_in + 1)); S.n_total = n_in + 1;
	S.nodes = e->b_chnodes.as<BtNode>((size_t)(n_in >> 1) + 2 * (size_t)n + 4);
	S.ctg_alt = (const uint8_t *)e->d_ctg_alt;
	int32_t *n_kc = e->b_chnc.as<int32_t>(n + 1), *n_ks = e->b_chns.as<int32_t>(n + 1);
	int64_t *coff = e->b_chcoff.as<int64_t>(n + 1), *soff = e->b_chsoff.as<int64_t>(n + 1);
	const SeedRec *seeds = (const SeedRec *)e->b_seeds.p;
	const int64_t *seed_off = (const int64_t *)e->b_seedoff.p;
	CK(cudaMemsetAsync(n_kc + n, 0, sizeof(int32_t), e->stream));
	CK(cudaMemsetAsync(n_ks + n, 0, sizeof(int32_t), e->stream));
	k_chain_build<<<grid_for(n, 128), 128, 0, e->stream>>>(co, e->fm.l_pac, S, n, (const int64_t *)e->d_off.p, seeds, seed_off, n_kc, n_ks);
	CK(cudaGetLastError());
	exclusive_scan(e, n_kc, coff, n + 1);
	exclusive_scan(e, n_ks, soff, n + 1);
	int64_t tot[2] = { 0, 0 };
	e->d2h(&tot[0], coff + n, sizeof(int64_t));
	e->d2h(&tot[1], soff + n, sizeof(int64_t));
	e->sync();
	if (tot[0] > 0x7fffffffLL || tot[1] > 0x7fffffffLL) die("too many chains/seeds in one batch");
	int32_t *d_co = e->b_chain_off.as<int32_t>(n + 1);
	DChain *d_ch = e->b_chains.as<DChain>(tot[0] + 1);
	DSeed *d_se = e->b_dseeds.as<DSeed>(tot[1] + 1);
	int32_t *d_srt = e->b_srt.as<int32_t>(tot[1] + 1);
	k_chain_emit<<<grid_for(n + 1, 128), 128, 0, e->stream>>>(co, e->fm, S, n, (const int64_t *)e->d_off.p, seeds, seed_off, (const int32_t *)e->b_lrep.p,
		n_kc, coff, soff, d_co, d_ch, d_se, d_srt);
	CK(cudaGetLastError());
	e->stats.n_launches += 2;
	e->stats.ms_k_chain += e->toc();
	in.n_reads = n; in.chain_off = nullptr; in.chains = nullptr; in.seeds = nullptr; in.srt = nullptr;
	in.n_chains = tot[0]; in.n_seeds = tot[1]; in.on_device = true;
	if (download) {
		int32_t *h_co = (int32_t *)stage_pinned(e, PIN_CHAIN_OFF, sizeof(int32_t) * (n + 1));
		DChain *h_ch = (DChain *)stage_pinned(e, PIN_CHAINS, sizeof(DChain) * (tot[0] + 1));
		DSeed *h_se = (DSeed *)stage_pinned(e, PIN_DSEEDS, sizeof(DSeed) * (tot[1] + 1));
		int32_t *h_srt = (int32_t *)stage_pinned(e, PIN_SRT, sizeof(int32_t) * (tot[1] + 1));
		e->d2h(h_co, d_co, sizeof(int32_t) * (n + 1));
		e->d2h(h_ch, d_ch, sizeof(DChain) * tot[0]);
		e->d2h(h_se, d_se, sizeof(DSeed) * tot[1]);
		e->d2h(h_srt, d_srt, sizeof(int32_t) * tot[1]);
		e->sync();
		in.chain_off = h_co; in.chains = h_ch; in.seeds = h_se; in.srt = h_srt;
	}
}

void stage_sa(Engine *e, int64_t n, const uint64_t *k, uint64_t *sa)
{
	CK(cudaSetDevice(e->device));
	if (n <= 0) return;
	uint64_t *dk = e->b_q.as<uint64_t>(n), *ds = e->b_t.as<uint64_t>(n);
	e->h2d(dk, k, sizeof(uint64_t) * n);
	k_sa_plain<<<grid_for(n, 256), 256, 0, e->stream>>>(e->fm, n, dk, ds);
	CK(cudaGetLastError());
	e->stats.n_launches += 1;
	e->d2h(sa, ds, sizeof(uint64_t) * n);
	e->sync();
}

__global__ void k_smem1_one(FmView fm, int len, const uint8_t *q, int x, uint64_t min_intv, Intv *mem, Intv *a, Intv *b, int32_t *out)
{
	int n = 0;
	out[0] = fm_smem1(fm, len, q, x, min_intv, mem, &n, a, b, nullptr);
	out[1] = n;
}

int stage_smem1(Engine *e, int len, const uint8_t *q, int x, uint64_t min_intv, std::vector<Intv> &mem)
{
	CK(cudaSetDevice(e->device));
	uint8_t *dq = e->b_q.as<uint8_t>(len + 16);
	Intv *d = e->b_t.as<Intv>((size_t)3 * (len + 1));
	int32_t *d_out = e->b_xctr.as<int32_t>(16);
	e->h2d(dq, q, len);
	k_smem1_one<<<1, 1, 0, e->stream>>>(e->fm, len, dq, x, min_intv, d, d + (len + 1), d + 2 * (len + 1), d_out);
	CK(cudaGetLastError());
	e->stats.n_launches += 1;
	int32_t h[2];
	e->d2h(h, d_out, sizeof h);
	e->sync();
	mem.resize(h[1]);
	if (h[1] > 0) { e->d2h(mem.data(), d, sizeof(Intv) * h[1]); e->sync(); }
	return h[0];
}

void stage_fm_extend(Engine *e, const Intv &ik, Intv ok[4], int is_back)
{
	CK(cudaSetDevice(e->device));
	Intv *d = e->b_t.as<Intv>(4);
	k_fm_extend1<<<1, 1, 0, e->stream>>>(e->fm, ik, d, is_back);
	CK(cudaGetLastError());
	e->stats.n_launches += 1;
	e->d2h(ok, d, sizeof(Intv) * 4);
	e->sync();
}

/* ------------------------------------------------------------------ extension */

// Rounds of (advance -> sort jobs by size -> batched DP) until every read has walked all its chains; see ext_rounds.cuh.
void *stage_pinned(Engine *e, int slot, size_t bytes)
{
	CK(cudaSetDevice(e->device));
	return e->h_slot[slot].need(bytes);
}

__global__ void k_ext_record(int n, const int32_t *__restrict__ active, const ExtJob *__restrict__ jobs, ExtJob *rec)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) rec[t] = jobs[active[t]];
}

// k_ext_dp over the n jobs of size class c in order[]: a persistent grid (as many blocks as the device can hold at that class's
// shared-memory footprint) whose lanes take jobs from *next (zeroed by the caller)
static void launch_ext_dp(Engine *e, cudaStream_t st, const ExtOpt &eo, const uint8_t *pac, const uint8_t *codes, ExtJob *jobs, const int32_t *order,
                          int n, int c, int qcap, unsigned long long *d_cells, unsigned long long *d_calls, int *next)
{
	static const int refill = getenv("B200_EXT_REFILL") ? std::max(1, std::min(32, atoi(getenv("B200_EXT_REFILL")))) : 16;
	const int threads = c == EXT_N_CLASS - 2 ? 32 : 64;
	const size_t per_warp = ((size_t)(qcap + 1) * 32 + (size_t)((qcap + 4) & ~3) * 8) * 4;
	const size_t smem = per_warp * (threads / 32);
	int n_sm = 0;
	CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, e->device));
	const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / threads, (size_t)(227 * 1024) / (smem + 1024)));
	const int grid = std::min(grid_for(n, threads), n_sm * per_sm);
	k_ext_dp<<<grid, threads, smem, st>>>(eo, pac, codes, jobs, order, n, qcap, d_cells, d_calls, next, refill);
}

static void ext_set_attrs(Engine *e)
{
	if (e->ext_attr_set) return;
	CK(cudaFuncSetAttribute(k_ext_dp, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
	CK(cudaFuncSetAttribute(k_ext_dp_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CK(cudaFuncSetAttribute(k_ext_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	e->ext_attr_set = true;
}

void stage_extend(Engine *e, const ExtOpt &eo, const ExtIn &in)
{
	CK(cudaSetDevice(e->device));
	const int n = in.n_reads;
	if (n != e->n_reads) die("stage_extend: chain table does not match the uploaded reads");
	if (n == 0) { CK(cudaMemsetAsync(e->b_soff.as<int64_t>(2), 0, sizeof(int64_t) * 2, e->stream)); return; }
	e->zero_counters();
	int32_t *d_co = e->b_chain_off.as<int32_t>(n + 1);
	DChain *d_ch = e->b_chains.as<DChain>(in.n_chains + 1);
	DSeed *d_se = e->b_dseeds.as<DSeed>(in.n_seeds + 1);
	int32_t *d_srt = e->b_srt.as<int32_t>(in.n_seeds + 1);
	DReg *d_regs = e->b_regs.as<DReg>(in.n_seeds + 1);
	int32_t *d_nr = e->b_nregs.as<int32_t>(n + 1);
	ExtState *d_state = e->b_xstate.as<ExtState>(n);
	ExtJob *d_jobs = e->b_xjobs.as<ExtJob>(n);
	int32_t *d_act[2] = { e->b_xact0.as<int32_t>(n), e->b_xact1.as<int32_t>(n) };
	uint32_t *d_key = e->b_xkey.as<uint32_t>(n), *d_key2 = e->b_xkey2.as<uint32_t>(n);
	int32_t *d_ord = e->b_xord.as<int32_t>(n);
	int32_t *d_ctr = e->b_xctr.as<int32_t>(16);
	if (!in.on_device) {
		e->h2d(d_co, in.chain_off, sizeof(int32_t) * (n + 1));
		e->h2d(d_ch, in.chains, sizeof(DChain) * in.n_chains);
		e->h2d(d_se, in.seeds, sizeof(DSeed) * in.n_seeds);
		e->h2d(d_srt, in.srt, sizeof(int32_t) * in.n_seeds);
	}
	ext_set_attrs(e);
	size_t sort_tmp = 0;
	CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_tmp, d_key, d_key2, d_act[0], d_ord, n, 0, 31, e->stream));
	void *d_sort_tmp = e->b_cub.need(sort_tmp);
	const int class_cap[EXT_N_CLASS] = { 32, 64, 96, 128, 160, 256, 704, 0x7fffffff };
	// warp-cooperative kernels (latency path): warps per block such that the rows of the longest query fit shared memory
	auto warps_for = [](int qcap) { return ext_warp_smem_bytes(4, qcap) <= 200 * 1024 ? 4 : ext_warp_smem_bytes(1, qcap) <= 200 * 1024 ? 1 : 0; };
	// measured on the bench workload (tools/tune_env.py): a class is better off with one warp per job below ~2 k jobs;
	// the fused tail kernel takes over once fewer than TAIL_READS_MAX reads still walk their chains
	const int WARP_JOBS_MAX = getenv("B200_EXT_WARP_MAX") ? atoi(getenv("B200_EXT_WARP_MAX")) : 2048;
	const int TAIL_READS_MAX = getenv("B200_EXT_TAIL_MAX") ? atoi(getenv("B200_EXT_TAIL_MAX")) : 50000;
	const int tail_warps = warps_for(e->max_len);
	const char *dbg = getenv("B200_DEBUG");
	const bool record = getenv("B200_EXT_RECORD") != nullptr;
	const int64_t rec_cap = record ? 2 * in.n_seeds + n : 0;
	ExtJob *d_rec = record ? e->b_xrec.as<ExtJob>((size_t)rec_cap + 1) : nullptr;
	if (record) e->n_rec = 0;
	int32_t ctr[16];
	unsigned long long *d_cells = &e->d_cnt->ext_cells, *d_calls = &e->d_cnt->ext_calls;
	float ms_dp = 0;
	e->tic();
	CK(cudaMemsetAsync(d_ctr, 0, 16 * sizeof(int32_t), e->stream));
	k_ext_init<<<grid_for(n, 256), 256, 0, e->stream>>>(n, d_co, d_ch, d_state, d_nr, d_act[0], d_ctr);
	CK(cudaGetLastError());
	e->stats.n_launches += 1;
	CK(cudaMemcpyAsync(ctr, d_ctr, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
	e->sync();
	int n_active = ctr[0], cur = 0, rounds = 0;
	size_t ev_used = 0;
	auto ev_pair = [&]() {
		if (ev_used + 2 > e->ev_pool.size()) { e->ev_pool.resize(ev_used + 2); CK(cudaEventCreate(&e->ev_pool[ev_used])); CK(cudaEventCreate(&e->ev_pool[ev_used + 1])); }
	};
	while (n_active > 0) {
		if (rounds > 0 && tail_warps && n_active <= TAIL_READS_MAX) {
			// few reads left: one warp per read finishes the walk on the device (advance + DP fused)
			ev_pair();
			CK(cudaEventRecord(e->ev_pool[ev_used], e->stream));
			const int threads = 32 * tail_warps;
			k_ext_tail<<<grid_for(n_active, tail_warps), threads, ext_warp_smem_bytes(tail_warps, e->max_len), e->stream>>>(eo, e->fm.l_pac, e->fm.pac,
				(const uint8_t *)e->d_codes.p, n_active, d_act[cur], (const int64_t *)e->d_off.p, d_co, d_ch, d_se, d_srt, d_state, d_jobs, d_regs, d_nr,
				e->max_len, d_cells, d_calls);
			CK(cudaGetLastError());
			CK(cudaEventRecord(e->ev_pool[ev_used + 1], e->stream));
			ev_used += 2;
			e->stats.n_launches += 1;
			if (dbg) fprintf(stderr, "[ext] tail kernel over %d reads\n", n_active);
			++rounds;
			break;
		}
		CK(cudaMemsetAsync(d_ctr, 0, 16 * sizeof(int32_t), e->stream));
		k_ext_advance<<<grid_for(n_active, 128), 128, 0, e->stream>>>(eo, e->fm.l_pac, n_active, d_act[cur], (const int64_t *)e->d_off.p, d_co, d_ch,
			d_se, d_srt, d_state, d_jobs, d_regs, d_nr, d_act[cur ^ 1], d_key, d_ctr);
		CK(cudaGetLastError());
		e->stats.n_launches += 1;
		CK(cudaMemcpyAsync(ctr, d_ctr, (1 + EXT_N_CLASS) * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
		e->sync();
		const int n_jobs = ctr[0];
		if (dbg)
			fprintf(stderr, "[ext] round %d active %d jobs %d classes %d %d %d %d %d %d %d %d\n", rounds, n_active, n_jobs, ctr[1], ctr[2], ctr[3],
			        ctr[4], ctr[5], ctr[6], ctr[7], ctr[8]);
		if (n_jobs == 0) break;
		if (record) {
			if (e->n_rec + n_jobs <= rec_cap) k_ext_record<<<grid_for(n_jobs, 256), 256, 0, e->stream>>>(n_jobs, d_act[cur ^ 1], d_jobs, d_rec + e->n_rec);
			e->n_rec = std::min<int64_t>(e->n_rec + n_jobs, rec_cap);
		}
		CK(cub::DeviceRadixSort::SortPairsDescending(d_sort_tmp, sort_tmp, d_key, d_key2, d_act[cur ^ 1], d_ord, n_jobs, 0, 31, e->stream));
		e->stats.n_launches += 3;
		ev_pair();
		CK(cudaEventRecord(e->ev_pool[ev_used], e->stream));
		// classes run concurrently on side streams (each launch has its own shared-memory footprint); largest jobs first
		int pos = 0, k = 0;
		CK(cudaEventRecord(e->ev_fork, e->stream));
		for (int c = EXT_N_CLASS - 1; c >= 0; --c) {
			const int cnt = ctr[1 + c];
			if (cnt == 0) continue;
			cudaStream_t st = e->side[k % Engine::N_SIDE];
			CK(cudaStreamWaitEvent(st, e->ev_fork, 0));
			const int qcap = c < EXT_N_CLASS - 1 ? class_cap[c] : e->max_len;
			const int wpb = warps_for(qcap);
			if (wpb && (cnt < WARP_JOBS_MAX || c >= EXT_N_CLASS - 3)) {
				k_ext_dp_warp<<<grid_for(cnt, wpb), 32 * wpb, ext_warp_smem_bytes(wpb, qcap), st>>>(eo, e->fm.pac, (const uint8_t *)e->d_codes.p,
					d_jobs, d_ord + pos, cnt, qcap, d_cells, d_calls);
			} else if (c < EXT_N_CLASS - 1) {
				launch_ext_dp(e, st, eo, e->fm.pac, (const uint8_t *)e->d_codes.p, d_jobs, d_ord + pos, cnt, c, qcap, d_cells, d_calls, d_ctr + 9 + c);
			} else {
				int64_t stride = ((int64_t)cnt + 31) & ~31ll;
				int32_t *d_eh = e->b_eh.as<int32_t>((size_t)stride * 2 * (e->max_len + 2));
				k_ext_dp_big<<<grid_for(cnt, 128), 128, 0, st>>>(eo, e->fm.pac, (const uint8_t *)e->d_codes.p, d_jobs, d_ord + pos, cnt,
					d_eh, stride, d_cells, d_calls);
			}
			CK(cudaGetLastError());
			CK(cudaEventRecord(e->ev_join[k % Engine::N_SIDE], st));
			e->stats.n_launches += 1;
			pos += cnt;
			++k;
		}
		for (int q = 0; q < k && q < Engine::N_SIDE; ++q) CK(cudaStreamWaitEvent(e->stream, e->ev_join[q], 0));
		CK(cudaEventRecord(e->ev_pool[ev_used + 1], e->stream));
		ev_used += 2;
		cur ^= 1;
		n_active = n_jobs;
		++rounds;
	}
	// compact the regions: reg_off = exclusive scan of n_regs
	CK(cudaMemsetAsync(d_nr + n, 0, sizeof(int32_t), e->stream));
	int64_t *d_roff = e->b_soff.as<int64_t>(n + 2);
	exclusive_scan(e, d_nr, d_roff, n + 1);
	int64_t total = 0;
	e->d2h(&total, d_roff + n, sizeof(int64_t));
	e->sync();
	DReg *d_out = e->b_xout.as<DReg>(total + 1);
	k_ext_gather<<<grid_for(n, 256), 256, 0, e->stream>>>(n, d_co, d_ch, d_nr, d_roff, d_regs, d_out);
	CK(cudaGetLastError());
	e->stats.n_launches += 1;
	e->stats.ms_k_extend += e->toc();
	for (size_t i = 0; i < ev_used; i += 2) {
		float ms;
		CK(cudaEventElapsedTime(&ms, e->ev_pool[i], e->ev_pool[i + 1]));
		ms_dp += ms;
		if (dbg) fprintf(stderr, "[ext] DP of round %d: %.3f ms\n", (int)(i / 2), ms);
	}
	e->stats.ms_k_extend_dp += ms_dp;
	e->stats.n_extend_rounds += rounds;
	Counters c = e->read_counters();
	e->stats.extend_cells += (int64_t)c.ext_cells;
	e->stats.n_extend_jobs += (int64_t)c.ext_calls;
}

void stage_extend_download(Engine *e, ExtRegs &out)
{
	CK(cudaSetDevice(e->device));
	const int n = e->n_reads;
	int64_t *reg_off = (int64_t *)e->h_slot[PIN_REG_OFF].need(sizeof(int64_t) * (n + 2));
	e->d2h(reg_off, e->b_soff.p, sizeof(int64_t) * (n + 1));
	e->sync();
	DReg *regs = (DReg *)e->h_slot[PIN_REGS].need(sizeof(DReg) * (reg_off[n] + 1));
	e->d2h(regs, e->b_xout.p, sizeof(DReg) * reg_off[n]);
	e->sync();
	out.regs = regs; out.reg_off = reg_off;
}

// Kernel-isolated ksw_extend2 (BASELINE configs[1]: "the exact ksw_extend2 job list ... dump once, replay on GPU"): all the jobs the
// last stage_extend call recorded (B200_EXT_RECORD) as ONE batch through the same DP kernels - no chain2aln rounds in between, so
// no round runs with too few jobs to fill the chip.  Returns the DP time in ms; *cells = reference cell count of the batch.
double stage_extend_replay(Engine *e, const ExtOpt &eo, int64_t *cells, int64_t *n_jobs)
{
	CK(cudaSetDevice(e->device));
	const int64_t n = e->n_rec;
	*cells = 0; *n_jobs = n;
	if (n == 0) return 0;
	ExtJob *d_rec = (ExtJob *)e->b_xrec.p;
	std::vector<ExtJob> hj(n);
	CK(cudaMemcpy(hj.data(), d_rec, sizeof(ExtJob) * n, cudaMemcpyDeviceToHost));
	const int class_cap[EXT_N_CLASS] = { 32, 64, 96, 128, 160, 256, 704, 0x7fffffff };
	std::vector<uint64_t> key(n);
	int64_t cnt[EXT_N_CLASS] = { 0 };
	for (int64_t i = 0; i < n; ++i) {
		const ExtJob &j = hj[i];
		int c = 0;
		if ((long long)j.h0 + (long long)j.qlen * eo.max_sc >= 32768) c = EXT_N_CLASS - 1;
		else while (j.qlen > class_cap[c]) ++c;
		++cnt[c];
		const uint64_t k32 = ((uint64_t)c << 28) | ((uint64_t)(std::min(j.qlen, 0x3fff) & 0xfff) << 16) | (uint64_t)std::min(j.tlen, 0xffff);
		key[i] = k32 << 32 | (uint64_t)i;
	}
	std::sort(key.begin(), key.end(), std::greater<uint64_t>());          // like the pipeline: largest class, longest query first
	std::vector<int32_t> order(n);
	for (int64_t i = 0; i < n; ++i) order[i] = (int32_t)(uint32_t)key[i];
	int32_t *d_ord = e->b_xord.as<int32_t>(n);
	CK(cudaMemcpy(d_ord, order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
	ext_set_attrs(e);
	auto warps_for = [](int qcap) { return ext_warp_smem_bytes(4, qcap) <= 200 * 1024 ? 4 : ext_warp_smem_bytes(1, qcap) <= 200 * 1024 ? 1 : 0; };
	e->zero_counters();
	int *d_next = e->b_xctr.as<int>(16);
	CK(cudaMemsetAsync(d_next, 0, 16 * sizeof(int), e->stream));
	e->sync();
	unsigned long long *d_cells = &e->d_cnt->ext_cells, *d_calls = &e->d_cnt->ext_calls;
	e->tic();
	CK(cudaEventRecord(e->ev_fork, e->stream));
	int64_t pos = 0;
	int k = 0;
	for (int c = EXT_N_CLASS - 1; c >= 0; --c) {
		const int n_c = (int)cnt[c];
		if (n_c == 0) continue;
		cudaStream_t st = e->side[k % Engine::N_SIDE];
		CK(cudaStreamWaitEvent(st, e->ev_fork, 0));
		const int qcap = c < EXT_N_CLASS - 1 ? class_cap[c] : e->max_len;
		const int wpb = warps_for(qcap);
		if (wpb && (n_c < 2048 || c >= EXT_N_CLASS - 3))
			k_ext_dp_warp<<<grid_for(n_c, wpb), 32 * wpb, ext_warp_smem_bytes(wpb, qcap), st>>>(eo, e->fm.pac, (const uint8_t *)e->d_codes.p, d_rec, d_ord + pos, n_c, qcap, d_cells, d_calls);
		else if (c < EXT_N_CLASS - 1) {
			launch_ext_dp(e, st, eo, e->fm.pac, (const uint8_t *)e->d_codes.p, d_rec, d_ord + pos, n_c, c, qcap, d_cells, d_calls, d_next + c);
		} else {
			const int64_t stride = ((int64_t)n_c + 31) & ~31ll;
			int32_t *d_eh = e->b_eh.as<int32_t>((size_t)stride * 2 * (e->max_len + 2));
			k_ext_dp_big<<<grid_for(n_c, 128), 128, 0, st>>>(eo, e->fm.pac, (const uint8_t *)e->d_codes.p, d_rec, d_ord + pos, n_c, d_eh, stride, d_cells, d_calls);
		}
		CK(cudaGetLastError());
		CK(cudaEventRecord(e->ev_join[k % Engine::N_SIDE], st));
		pos += n_c;
		++k;
	}
	for (int q = 0; q < k && q < Engine::N_SIDE; ++q) CK(cudaStreamWaitEvent(e->stream, e->ev_join[q], 0));
	const double ms = e->toc();
	Counters cc = e->read_counters();
	*cells = (int64_t)cc.ext_cells;
	return ms;
}

// b200_ksw_extend2_batch: the caller's jobs run through the same three DP kernels as the pipeline (one job per lane,
// one warp per job, general int32 path), with the target read from the caller's byte buffer and exactly one call per
// job with the caller's band.  B200_EXT_KERNEL=lane|warp|big forces one kernel (used by the parity tests to fuzz each).
void stage_extend_bytes(Engine *e, const ExtOpt &eo, int64_t n_jobs, b200_extend_job_t *jobs,
                        const uint8_t *query, int64_t qbytes, const uint8_t *target, int64_t tbytes)
{
	CK(cudaSetDevice(e->device));
	if (n_jobs <= 0) return;
	const char *force = getenv("B200_EXT_KERNEL");
	const int mode = !force ? 0 : !strcmp(force, "lane") ? 1 : !strcmp(force, "warp") ? 2 : !strcmp(force, "big") ? 3 : 0;
	const int class_cap[EXT_N_CLASS] = { 32, 64, 96, 128, 160, 256, 704, 0x7fffffff };
	std::vector<ExtJob> hj(n_jobs);
	std::vector<int32_t> order(n_jobs);
	int64_t cnt[EXT_N_CLASS] = { 0 }, pos[EXT_N_CLASS];
	std::vector<uint8_t> cls(n_jobs);
	int max_q = 0;
	for (int64_t i = 0; i < n_jobs; ++i) {
		const b200_extend_job_t &j = jobs[i];
		ExtJob &x = hj[i];
		memset(&x, 0, sizeof x);
		x.qaddr = j.q_off; x.qstep = 1; x.f0 = j.t_off; x.fstep = 1; x.comp = 2;
		x.qlen = j.qlen; x.tlen = j.tlen; x.h0 = j.h0; x.prev = -1; x.bonus = j.end_bonus; x.w0 = j.w > 0 ? j.w : 1;
		int c = 0;
		if ((long long)j.h0 + (long long)j.qlen * eo.max_sc >= 32768) c = EXT_N_CLASS - 1;
		else while (j.qlen > class_cap[c]) ++c;
		if (mode == 3) c = EXT_N_CLASS - 1;
		cls[i] = (uint8_t)c; ++cnt[c];
		max_q = std::max(max_q, j.qlen);
	}
	pos[0] = 0;
	for (int c = 1; c < EXT_N_CLASS; ++c) pos[c] = pos[c - 1] + cnt[c - 1];
	{ int64_t w[EXT_N_CLASS]; for (int c = 0; c < EXT_N_CLASS; ++c) w[c] = pos[c]; for (int64_t i = 0; i < n_jobs; ++i) order[w[cls[i]]++] = (int32_t)i; }
	e->zero_counters();
	ExtJob *dj = e->b_xjobs.as<ExtJob>(n_jobs);
	int32_t *d_ord = e->b_xord.as<int32_t>(n_jobs);
	uint8_t *dq = e->b_q.as<uint8_t>(qbytes + 16), *dt = e->b_t.as<uint8_t>(tbytes + 16);
	e->h2d(dj, hj.data(), sizeof(ExtJob) * n_jobs);
	e->h2d(d_ord, order.data(), sizeof(int32_t) * n_jobs);
	e->h2d(dq, query, qbytes);
	e->h2d(dt, target, tbytes);
	ext_set_attrs(e);
	auto warps_for = [](int qcap) { return ext_warp_smem_bytes(4, qcap) <= 200 * 1024 ? 4 : ext_warp_smem_bytes(1, qcap) <= 200 * 1024 ? 1 : 0; };
	unsigned long long *d_cells = &e->d_cnt->ext_cells, *d_calls = &e->d_cnt->ext_calls;
	int *d_next = e->b_xctr.as<int>(16);
	CK(cudaMemsetAsync(d_next, 0, 16 * sizeof(int), e->stream));
	e->tic();
	for (int c = 0; c < EXT_N_CLASS; ++c) {
		const int n = (int)cnt[c];
		if (n == 0) continue;
		const int qcap = c < EXT_N_CLASS - 1 ? class_cap[c] : max_q;
		const int wpb = warps_for(qcap);
		const bool lane_ok = c < EXT_N_CLASS - 1;
		const bool use_warp = wpb && mode != 3 && (mode == 2 || !lane_ok || (mode == 0 && n < 8192));
		if (use_warp)
			k_ext_dp_warp<<<grid_for(n, wpb), 32 * wpb, ext_warp_smem_bytes(wpb, qcap), e->stream>>>(eo, dt, dq, dj, d_ord + pos[c], n, qcap, d_cells, d_calls);
		else if (lane_ok) {
			launch_ext_dp(e, e->stream, eo, dt, dq, dj, d_ord + pos[c], n, c, qcap, d_cells, d_calls, d_next + c);
		} else {
			const int64_t stride = ((int64_t)n + 31) & ~31ll;
			int32_t *d_eh = e->b_eh.as<int32_t>((size_t)stride * 2 * (max_q + 2));
			k_ext_dp_big<<<grid_for(n, 128), 128, 0, e->stream>>>(eo, dt, dq, dj, d_ord + pos[c], n, d_eh, stride, d_cells, d_calls);
		}
		CK(cudaGetLastError());
		e->stats.n_launches += 1;
	}
	e->stats.ms_k_extend += e->toc();
	e->d2h(hj.data(), dj, sizeof(ExtJob) * n_jobs);
	Counters cc = e->read_counters();
	for (int64_t i = 0; i < n_jobs; ++i) {
		const ExtJob &x = hj[i];
		b200_extend_job_t &j = jobs[i];
		j.score = x.score; j.qle = x.qle; j.tle = x.tle; j.gtle = x.gtle; j.gscore = x.gscore; j.max_off = x.max_off;
	}
	e->stats.extend_cells += (int64_t)cc.ext_cells;
	e->stats.n_extend_jobs += n_jobs;
}

/* ------------------------------------------------------------------ local Smith-Waterman (mate rescue, seed filter) */

// query accessor with a run-time orientation (read only while a pass loads its strip of the query into registers)
struct SQAny {
	const uint8_t *p; int l, rev;
	__device__ __forceinline__ int operator()(int j) const { if (!rev) return p[j]; int c = p[l - 1 - j]; return c < 4 ? 3 - c : 4; }
};

// job sources of the Smith-Waterman kernels: the rescue/seed-filter jobs of the pipeline (query = a resident read, target =
// a window of the packed reference) and the flat byte buffers of b200_ksw_align2_batch
struct SwSrcPipeline {
	const SwJob *jobs; const int64_t *off; const uint8_t *codes; const uint8_t *pac; int64_t l_pac; SwRes *res;
	typedef STPac Target;
	__device__ __forceinline__ void load(int jx, int &qlen, int &tlen, int &xtra, SQAny &q, STPac &t) const
	{
		const SwJob j = jobs[jx];
		qlen = j.q_len; tlen = j.tlen; xtra = j.xtra;
		q.p = codes + off[j.read] + j.q_beg; q.l = j.q_len; q.rev = j.is_rev;
		t.pac = pac; t.l_pac = l_pac; t.beg = j.rb;
		prefetch_ref_window(pac, l_pac, j.rb, j.rb + j.tlen, threadIdx.x & 31, 32);     // (one warp per job; harmless when one thread runs it)
	}
	__device__ __forceinline__ void store(int jx, const SwRes &r) const { res[jx] = r; }
};
struct SwSrcBytes {
	b200_align_job_t *jobs; const uint8_t *query, *target;
	typedef STBytes Target;
	__device__ __forceinline__ void load(int jx, int &qlen, int &tlen, int &xtra, SQAny &q, STBytes &t) const
	{
		const b200_align_job_t &j = jobs[jx];
		qlen = j.qlen; tlen = j.tlen; xtra = j.xtra;
		q.p = query + j.q_off; q.l = j.qlen; q.rev = 0;
		t.p = target + j.t_off;
	}
	__device__ __forceinline__ void store(int jx, const SwRes &r) const
	{
		kswr_t &o = jobs[jx].r;
		o.score = r.score; o.te = r.te; o.qe = r.qe; o.score2 = r.score2; o.te2 = r.te2; o.tb = r.tb; o.qb = r.qb;
	}
};

// one warp per job (sw_warp_kernel.cuh); order[w] = job of warp w, B = run-list scratch (bcap entries per warp)
template <int C, class SRC>
__global__ void __launch_bounds__(128) k_sw_warp(SwOpt so, SRC src, const int32_t *__restrict__ order, int n, uint64_t *B, int64_t bcap, Counters *cnt)
{
	__shared__ uint32_t lut[10];
	sw_fill_lut(so, lut);
	__syncthreads();
	const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	long long cells = 0;
	if (w < n) {
		const int jx = order[w];
		int qlen, tlen, xtra;
		SQAny q;
		typename SRC::Target t;
		src.load(jx, qlen, tlen, xtra, q, t);
		SwRes r;
		sw_align_warp<C>(qlen, q, tlen, t, so, lut, xtra, B + (int64_t)w * bcap, &r, &cells);
		if ((threadIdx.x & 31) == 0) src.store(jx, r);
	}
	if (cells) atomicAdd(&cnt->sw_cells, (unsigned long long)cells);
}

// general path (queries wider than 256 padded columns): one job per thread, rows in global memory
template <class SRC>
__global__ void __launch_bounds__(128) k_sw_thread(SwOpt so, SRC src, const int32_t *__restrict__ order, int n, uint16_t *H, uint16_t *E, uint64_t *B,
                                                   int64_t stride, Counters *cnt)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	int64_t cells = 0;
	if (t < n) {
		const int jx = order[t];
		int qlen, tlen, xtra;
		SQAny q;
		typename SRC::Target ta;
		src.load(jx, qlen, tlen, xtra, q, ta);
		Row16 h = { H + t, stride }, ee = { E + t, stride };
		List64 bl = { B + t, stride };
		SwRes r;
		sw_align(qlen, q, tlen, ta, so, xtra, h, ee, bl, &r, &cells);
		src.store(jx, r);
	}
	warp_add(&cnt->sw_cells, cells);
}

// Launches the ksw_align2 kernels over classified jobs: d_ord lists the jobs class by class (general kernel first, then strip
// widths 2, 4, 5, 8 query columns per lane), cnt[k] jobs in class k; one launch per class on concurrent streams.
template <class SRC>
static void sw_launch_classes(Engine *e, const SwOpt &so, const SRC &src, const int32_t *d_ord, const int32_t cnt[5], int max_t, int max_q)
{
	int64_t pos[5], n = 0;
	pos[0] = 0;
	for (int k = 1; k < 5; ++k) pos[k] = pos[k - 1] + cnt[k - 1];
	n = pos[4] + cnt[4];
	const int64_t bcap = max_t / 2 + 2;
	uint64_t *B = e->b_b.as<uint64_t>((size_t)(n + 32) * bcap);
	e->tic();
	CK(cudaEventRecord(e->ev_fork, e->stream));
	int used = 0;
	for (int k = 4; k >= 0; --k) {
		if (cnt[k] == 0) continue;
		cudaStream_t st = e->side[used % Engine::N_SIDE];
		CK(cudaStreamWaitEvent(st, e->ev_fork, 0));
		const int nk = cnt[k];
		const int32_t *ord = d_ord + pos[k];
		uint64_t *Bk = B + pos[k] * bcap;
		const int grid = grid_for((int64_t)nk * 32, 128);
		if (k == 4) k_sw_warp<8, SRC><<<grid, 128, 0, st>>>(so, src, ord, nk, Bk, bcap, e->d_cnt);
		else if (k == 3) k_sw_warp<5, SRC><<<grid, 128, 0, st>>>(so, src, ord, nk, Bk, bcap, e->d_cnt);
		else if (k == 2) k_sw_warp<4, SRC><<<grid, 128, 0, st>>>(so, src, ord, nk, Bk, bcap, e->d_cnt);
		else if (k == 1) k_sw_warp<2, SRC><<<grid, 128, 0, st>>>(so, src, ord, nk, Bk, bcap, e->d_cnt);
		else {
			const int64_t stride = ((int64_t)nk + 31) & ~31ll;
			uint16_t *H = e->b_h.as<uint16_t>((size_t)stride * (max_q + 16));
			uint16_t *E = e->b_e.as<uint16_t>((size_t)stride * (max_q + 16));
			uint64_t *B0 = e->b_eh.as<uint64_t>((size_t)stride * bcap);
			k_sw_thread<SRC><<<grid_for(nk, 128), 128, 0, st>>>(so, src, ord, nk, H, E, B0, stride, e->d_cnt);
		}
		CK(cudaGetLastError());
		CK(cudaEventRecord(e->ev_join[used % Engine::N_SIDE], st));
		e->stats.n_launches += 1;
		++used;
	}
	for (int q = 0; q < used && q < Engine::N_SIDE; ++q) CK(cudaStreamWaitEvent(e->stream, e->ev_join[q], 0));
	e->stats.ms_k_sw += e->toc();
}

// caller-provided batches (stage_sw / stage_sw_bytes): lengths are on the host, so the classification is a host loop
template <class SRC, class LEN>
static void run_sw(Engine *e, const SwOpt &so, const SRC &src, int64_t n, LEN len_of /* (i, &qlen, &tlen, &xtra) */)
{
	std::vector<int32_t> order(n);
	int32_t cnt[5] = { 0, 0, 0, 0, 0 };
	int64_t pos[5];
	int max_q = 0, max_t = 0;
	std::vector<uint8_t> cls(n);
	for (int64_t i = 0; i < n; ++i) {
		int ql, tl, xt;
		len_of(i, ql, tl, xt);
		const int c = sw_warp_class(ql, xt);
		const int k = c == 0 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 5 ? 3 : 4;
		cls[i] = (uint8_t)k; ++cnt[k];
		max_t = std::max(max_t, tl);
		if (k == 0) max_q = std::max(max_q, ql);
	}
	pos[0] = 0;
	for (int k = 1; k < 5; ++k) pos[k] = pos[k - 1] + cnt[k - 1];
	{ int64_t w[5]; for (int k = 0; k < 5; ++k) w[k] = pos[k]; for (int64_t i = 0; i < n; ++i) order[w[cls[i]]++] = (int32_t)i; }
	int32_t *d_ord = e->b_xord.as<int32_t>(n);
	e->h2d(d_ord, order.data(), sizeof(int32_t) * n);
	sw_launch_classes(e, so, src, d_ord, cnt, max_t, max_q);
}

void stage_sw(Engine *e, const SwOpt &so, const std::vector<SwJob> &jobs, std::vector<SwRes> &out)
{
	CK(cudaSetDevice(e->device));
	out.resize(jobs.size());
	if (jobs.empty()) return;
	e->zero_counters();
	const int64_t n = (int64_t)jobs.size();
	SwJob *dj = e->b_jobs.as<SwJob>(n);
	SwRes *dr = e->b_res.as<SwRes>(n);
	e->h2d(dj, jobs.data(), sizeof(SwJob) * n);
	SwSrcPipeline src = { dj, (const int64_t *)e->d_off.p, (const uint8_t *)e->d_codes.p, e->fm.pac, e->fm.l_pac, dr };
	run_sw(e, so, src, n, [&](int64_t i, int &ql, int &tl, int &xt) { ql = jobs[i].q_len; tl = jobs[i].tlen; xt = jobs[i].xtra; });
	e->d2h(out.data(), dr, sizeof(SwRes) * n);
	Counters c = e->read_counters();
	e->stats.sw_cells += (int64_t)c.sw_cells;
	e->stats.n_sw_jobs += n;
}

void stage_sw_bytes(Engine *e, const SwOpt &so, int64_t n_jobs, b200_align_job_t *jobs,
                    const uint8_t *query, int64_t qbytes, const uint8_t *target, int64_t tbytes)
{
	CK(cudaSetDevice(e->device));
	if (n_jobs <= 0) return;
	e->zero_counters();
	b200_align_job_t *dj = e->b_jobs.as<b200_align_job_t>(n_jobs);
	uint8_t *dq = e->b_q.as<uint8_t>(qbytes + 16), *dt = e->b_t.as<uint8_t>(tbytes + 16);
	e->h2d(dj, jobs, sizeof(b200_align_job_t) * n_jobs);
	e->h2d(dq, query, qbytes);
	e->h2d(dt, target, tbytes);
	SwSrcBytes src = { dj, dq, dt };
	run_sw(e, so, src, n_jobs, [&](int64_t i, int &ql, int &tl, int &xt) { ql = jobs[i].qlen; tl = jobs[i].tlen; xt = jobs[i].xtra; });
	e->d2h(jobs, dj, sizeof(b200_align_job_t) * n_jobs);
	Counters c = e->read_counters();
	e->stats.sw_cells += (int64_t)c.sw_cells;
	e->stats.n_sw_jobs += n_jobs;
}

/* ------------------------------------------------------------------ CIGAR stage (banded global alignment + traceback) */

// Fast path: one region per lane.  The H/E row of a lane is a band-wide circular window (S columns, S >= 2*band+2) in
// shared memory laid out [column][lane] as {h,e} pairs, the oriented query sits in shared memory as bytes, substitution
// scores come from one PRMT over the two-register score row of the target base.
struct GlobalRowSh {
	int2 *base; int mask;
	__device__ __forceinline__ int32_t h(int j) const { return base[(j & mask) << 5].x; }
	__device__ __forceinline__ int32_t e(int j) const { return base[(j & mask) << 5].y; }
	__device__ __forceinline__ void set_h(int j, int32_t v) const { base[(j & mask) << 5].x = v; }
	__device__ __forceinline__ void set_e(int j, int32_t v) const { base[(j & mask) << 5].y = v; }
};
struct GlobalSeqsSh {
	const uint8_t *Q; int l_query;
	const uint8_t *pac; int64_t l_pac, rb, re; int rev;
	const uint32_t *lut;
	__device__ __forceinline__ int qa(int j) const { return Q[ext_qidx(j)]; }
	__device__ __forceinline__ int ta(int i) const { return fm_base(pac, l_pac, rev ? re - 1 - i : rb + i); }
	__device__ __forceinline__ uint2 trow(const GlobalOpt &, int i) const { const int t = ta(i); return make_uint2(lut[t * 2], lut[t * 2 + 1]); }
	__device__ __forceinline__ int sub(const uint2 &row, int j) const { return prmt_score(row.x, row.y, (uint32_t)qa(j) * 0x1111u + 0x8880u); }
};

__global__ void __launch_bounds__(64) k_global_lanes(GlobalOpt go, const uint8_t *__restrict__ pac, int64_t l_pac, int n,
                                                     const GlobalJob *__restrict__ jobs, const int32_t *__restrict__ order,
                                                     const int64_t *__restrict__ off, const uint8_t *__restrict__ codes, uint8_t *z,
                                                     uint32_t *cig, GlobalRes *res, int S, int qcap, Counters *cnt)
{
	extern __shared__ uint32_t smem[];
	__shared__ uint32_t lut[10];
	if (threadIdx.x < 5) {
		const int8_t *m = go.mat + threadIdx.x * 5;
		lut[threadIdx.x * 2] = (uint32_t)(uint8_t)m[0] | (uint32_t)(uint8_t)m[1] << 8 | (uint32_t)(uint8_t)m[2] << 16 | (uint32_t)(uint8_t)m[3] << 24;
		lut[threadIdx.x * 2 + 1] = (uint32_t)(uint8_t)m[4];
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int qpad = (qcap + 4) & ~3;
	const size_t per_warp_words = (size_t)S * 64 + (size_t)qpad * 8;
	int2 *rows = (int2 *)(smem + wib * per_warp_words) + lane;
	uint8_t *Q = (uint8_t *)(smem + wib * per_warp_words + (size_t)S * 64) + lane * 4;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	int64_t cells = 0;
	if (t < n) {
		const int jx = order[t];
		const GlobalJob j = jobs[jx];
		GlobalSeqsSh s;
		s.Q = Q; s.l_query = j.qe - j.qb; s.pac = pac; s.l_pac = l_pac; s.rb = j.rb; s.re = j.re; s.rev = j.rb >= l_pac; s.lut = lut;
		const uint8_t *q = codes + off[j.read] + j.qb;
		prefetch_ref_window(pac, l_pac, j.rb, j.re, 0, 1);
		for (int x = 0; x < s.l_query; ++x) Q[ext_qidx(x)] = s.rev ? q[s.l_query - 1 - x] : q[x];
		GlobalRowSh eh = { rows, S - 1 };
		global_task(go, s, j, eh, z + j.zoff, cig + j.cig_off, &res[jx], &cells, (S - 2) >> 1);
	}
	warp_add(&cnt->global_cells, cells);
}

// general path: one region per thread; H/E row interleaved over the threads of the launch in global memory
__global__ void __launch_bounds__(128) k_global_jobs(GlobalOpt go, const uint8_t *__restrict__ pac, int64_t l_pac, int n,
                                                     const GlobalJob *__restrict__ jobs, const int32_t *__restrict__ order,
                                                     const int64_t *__restrict__ off, const uint8_t *__restrict__ codes, int32_t *rows,
                                                     int64_t stride, uint8_t *z, uint32_t *cig, GlobalRes *res, Counters *cnt)
{
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	int64_t cells = 0;
	if (t < n) {
		const int jx = order[t];
		const GlobalJob j = jobs[jx];
		GlobalRow eh = { rows + t, stride };
		global_task(go, global_seqs(pac, l_pac, codes + off[j.read] + j.qb, j), j, eh, z + j.zoff, cig + j.cig_off, &res[jx], &cells);
	}
	warp_add(&cnt->global_cells, cells);
}

// Launches the CIGAR-stage kernels over classified jobs (finish_stage.h: GlobalClassTask): d_ord lists the jobs class by class
// - row windows of 32, 64, 128, 256, 512 columns in shared memory, then the general kernel - and, within a class, by band and
// target length, so that the lanes of a warp get regions of similar cost.  One launch per class on concurrent streams.
static void global_launch_classes(Engine *e, const GlobalOpt &go, const GlobalJob *dj, const int32_t *d_ord, const int32_t cnt[6], const int32_t qmax[6],
                                  const int64_t *d_off, const uint8_t *d_codes, uint8_t *z, uint32_t *cig, GlobalRes *dr,
                                  const uint8_t *pac = nullptr, int64_t l_pac = 0)
{
	if (!pac) { pac = e->fm.pac; l_pac = e->fm.l_pac; }
	static const int cls_S[5] = { 32, 64, 128, 256, 512 };
	if (!e->global_attr_set) { CK(cudaFuncSetAttribute(k_global_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); e->global_attr_set = true; }
	int64_t pos[6];
	pos[0] = 0;
	for (int k = 1; k < 6; ++k) pos[k] = pos[k - 1] + cnt[k - 1];
	e->tic();
	CK(cudaEventRecord(e->ev_fork, e->stream));
	int used = 0;
	for (int k = 5; k >= 0; --k) {
		if (cnt[k] == 0) continue;
		cudaStream_t st = e->side[used % Engine::N_SIDE];
		CK(cudaStreamWaitEvent(st, e->ev_fork, 0));
		const int nk = cnt[k];
		if (k < 5) {
			const int S = cls_S[k], qcap = qmax[k];
			const size_t per_warp = ((size_t)S * 64 + (size_t)((qcap + 4) & ~3) * 8) * 4;
			const int threads = per_warp * 2 <= 200 * 1024 ? 64 : 32;
			k_global_lanes<<<grid_for(nk, threads), threads, per_warp * (threads / 32), st>>>(go, pac, l_pac, nk, dj, d_ord + pos[k],
				d_off, d_codes, z, cig, dr, S, qcap, e->d_cnt);
		} else {
			const int64_t stride = ((int64_t)nk + 31) & ~31ll;
			int32_t *rows = e->b_grow.as<int32_t>((size_t)stride * 2 * (qmax[k] + 2));
			k_global_jobs<<<grid_for(nk, 128), 128, 0, st>>>(go, pac, l_pac, nk, dj, d_ord + pos[k], d_off, d_codes, rows, stride, z, cig, dr, e->d_cnt);
		}
		CK(cudaGetLastError());
		CK(cudaEventRecord(e->ev_join[used % Engine::N_SIDE], st));
		e->stats.n_launches += 1;
		++used;
	}
	for (int q = 0; q < used && q < Engine::N_SIDE; ++q) CK(cudaStreamWaitEvent(e->stream, e->ev_join[q], 0));
	e->stats.ms_k_global += e->toc();
}

// ksw_global2 for one caller-provided job (the C wrapper): the byte buffers stand in for read and reference
__global__ void k_global_one(GlobalOpt go, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int w, int32_t *rows, uint8_t *z, uint32_t *cig, int32_t *out)
{
	struct Seqs {
		const uint8_t *q, *t; int l_query;
		__device__ int qa(int j) const { return q[j]; }
		__device__ int ta(int i) const { return t[i]; }
		__device__ const int8_t *trow(const GlobalOpt &o, int i) const { return o.mat + t[i] * 5; }
		__device__ int sub(const int8_t *row, int j) const { return row[q[j]]; }
	} s = { q, t, qlen };
	GlobalRow eh = { rows, 1 };
	int n_cigar = 0;
	out[0] = global_dp(go, s, tlen, w, eh, z, cig, &n_cigar, nullptr);
	out[1] = n_cigar;
}

int stage_global_bytes(Engine *e, const GlobalOpt &go, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int w, std::vector<uint32_t> *cigar)
{
	CK(cudaSetDevice(e->device));
	uint8_t *dq = e->b_q.as<uint8_t>(qlen + 16), *dt = e->b_t.as<uint8_t>(tlen + 16);
	const int n_col = ((qlen < 2 * w + 1 ? qlen : 2 * w + 1) + 3) & ~3;
	int32_t *rows = e->b_grow.as<int32_t>((size_t)2 * (qlen + 2));
	uint8_t *z = e->fb[FB_GZ].as<uint8_t>((size_t)n_col * (tlen + 1) + 64);
	uint32_t *cig = e->fb[FB_CIG].as<uint32_t>((size_t)qlen + tlen + 8);
	int32_t *d_out = e->fb[FB_CTR].as<int32_t>(64);
	e->h2d(dq, query, qlen);
	e->h2d(dt, target, tlen);
	k_global_one<<<1, 1, 0, e->stream>>>(go, qlen, dq, tlen, dt, w, rows, z, cig, d_out);
	CK(cudaGetLastError());
	e->stats.n_launches += 1;
	int32_t h[2];
	e->d2h(h, d_out, sizeof h);
	e->sync();
	if (cigar) {
		cigar->resize(h[1]);
		if (h[1] > 0) { e->d2h(cigar->data(), cig, sizeof(uint32_t) * h[1]); e->sync(); }
	}
	return h[0];
}

/* ------------------------------------------------------------------ finish stages: the CUDA backend of finish_stage.h */

template <class TASK>
__global__ void __launch_bounds__(128) k_task(int64_t n, TASK task)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) task(i);
}

struct CudaBK {
	Engine *e;
	bool text_phase = false;
	template <class T> T *buf(int id, size_t n) { return e->fb[id].as<T>(n ? n : 1); }
	template <class T> T *grow(int id, size_t n, size_t keep)
	{
		DevBuf &b = e->fb[id];
		if (n * sizeof(T) <= b.cap) return (T *)b.p;
		void *q = nullptr;
		const size_t cap = n * sizeof(T) + (n * sizeof(T) >> 2) + 256;
		CK(cudaMallocAsync(&q, cap, e->stream));
		if (keep) CK(cudaMemcpyAsync(q, b.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, e->stream));
		if (b.p) CK(cudaFreeAsync(b.p, e->stream));
		b.p = q; b.cap = cap;
		return (T *)b.p;
	}
	template <class TASK> void run(int64_t n, const TASK &t)
	{
		if (n <= 0) return;
		static_assert(sizeof(TASK) <= 4000, "task does not fit the kernel parameter space");
		cudaEvent_t a, b;
		ev_pair(a, b);
		CK(cudaEventRecord(a, e->stream));
		k_task<TASK><<<grid_for(n, 128), 128, 0, e->stream>>>(n, t);
		CK(cudaGetLastError());
		CK(cudaEventRecord(b, e->stream));
		e->stats.n_launches += 1;
	}
	void scan(const int32_t *in, int64_t *out, int64_t n) { exclusive_scan(e, in, out, n); }
	void sort_pairs(uint32_t *key, int32_t *val, int64_t n)
	{
		uint32_t *key2 = e->b_xkey2.as<uint32_t>(n);
		int32_t *val2 = e->b_xact1.as<int32_t>(n);
		cub::DoubleBuffer<uint32_t> dk(key, key2);
		cub::DoubleBuffer<int32_t> dv(val, val2);
		size_t tmp = 0;
		CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)n, 0, 20, e->stream));
		void *d = e->b_cub.need(tmp);
		CK(cub::DeviceRadixSort::SortPairs(d, tmp, dk, dv, (int)n, 0, 20, e->stream));
		if (dv.Current() != val) CK(cudaMemcpyAsync(val, dv.Current(), sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, e->stream));
		if (dk.Current() != key) CK(cudaMemcpyAsync(key, dk.Current(), sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, e->stream));
		e->stats.n_launches += 2;
	}
	void zero(void *p, size_t bytes) { CK(cudaMemsetAsync(p, 0, bytes, e->stream)); }
	int64_t get64(const int64_t *p) { int64_t v = 0; e->d2h(&v, p, sizeof v); e->sync(); return v; }
	int32_t get32(const int32_t *p) { int32_t v = 0; e->d2h(&v, p, sizeof v); e->sync(); return v; }
	void upload(void *dst, const void *src, size_t bytes) { e->h2d(dst, src, bytes); e->sync(); }
	void download(void *dst, const void *src, size_t bytes) { e->d2h(dst, src, bytes); e->sync(); }
	void sw_launch(const SwOpt &so, const SwJob *jobs, SwRes *res, const int32_t *order, const int32_t cnt[5], int max_t, int max_q)
	{
		SwSrcPipeline src = { jobs, (const int64_t *)e->d_off.p, (const uint8_t *)e->d_codes.p, e->fm.pac, e->fm.l_pac, res };
		sw_launch_classes(e, so, src, order, cnt, max_t, max_q);
	}
	void global_launch(const GlobalOpt &go, const GlobalJob *jobs, const int32_t *order, const int32_t cnt[6], const int32_t qmax[6], uint8_t *z,
	                   uint32_t *cig, GlobalRes *res)
	{
		global_launch_classes(e, go, jobs, order, cnt, qmax, (const int64_t *)e->d_off.p, (const uint8_t *)e->d_codes.p, z, cig, res);
	}
	// CUDA-event pairs around every task launch: summed after the call (no synchronisation inside the sequence)
	std::vector<size_t> text_from;
	size_t ev_used = 0;
	void ev_pair(cudaEvent_t &a, cudaEvent_t &b)
	{
		if (ev_used + 2 > e->ev_pool.size()) { e->ev_pool.resize(ev_used + 2); CK(cudaEventCreate(&e->ev_pool[ev_used])); CK(cudaEventCreate(&e->ev_pool[ev_used + 1])); }
		a = e->ev_pool[ev_used]; b = e->ev_pool[ev_used + 1];
		ev_used += 2;
	}
	double task_ms()
	{
		double ms = 0;
		for (size_t i = 0; i < ev_used; i += 2) { float x; CK(cudaEventElapsedTime(&x, e->ev_pool[i], e->ev_pool[i + 1])); ms += x; }
		return ms;
	}
};

static double fin_clock_ms()
{
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct IsNewline { const char *t; __device__ bool operator()(int64_t i) const { return t[i] == '\n'; } };
__global__ void __launch_bounds__(256) k_count_newlines(const char *__restrict__ t, int64_t n, unsigned long long *out)
{
	long long c = 0;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += t[i] == '\n';
	warp_add(out, c);
}
struct FastqReadTask {
	FastqView v; ReadText *rt; int64_t *seq_at; int32_t *len, *ctr;
	B200_HD void operator()(int64_t r) const
	{
		int32_t l;
		fastq_read(v, r, &rt[r], &seq_at[r], &l);
		len[r] = l;
		FIN_ATOMIC_MAX(ctr, l);
	}
};
// one warp per read: the bases of its sequence line -> codes at the read's offset
__global__ void __launch_bounds__(256) k_fastq_encode(int n, const char *__restrict__ text, const int64_t *__restrict__ seq_at, const int64_t *__restrict__ off, uint8_t *codes)
{
	const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= n) return;
	const int lane = threadIdx.x & 31;
	const int64_t o = off[w], s = seq_at[w];
	const int l = (int)(off[w + 1] - o);
	for (int j = lane; j < l; j += 32) codes[o + j] = fq_code((uint8_t)text[s + j]);
}

void stage_upload_fastq(Engine *e, const char *fq1, int64_t len1, const char *fq2, int64_t len2, FastqInfo *info)
{
	CK(cudaSetDevice(e->device));
	const int paired = fq2 != nullptr;
	if (!paired) len2 = 0;
	char *text = e->d_text.as<char>((size_t)(len1 + len2) + 16);
	// page-locked buffers (b200_big_alloc, cudaHostRegister) go straight to the copy engine; pageable ones through the driver's staging
	e->h2d(text, fq1, (size_t)len1);
	if (paired) e->h2d(text + len1, fq2, (size_t)len2);
	CudaBK bk = { e };
	int32_t *ctr = e->fb[FB_CTR].as<int32_t>(64);
	CK(cudaMemsetAsync(ctr, 0, 64 * sizeof(int32_t), e->stream));
	int64_t *nl[2] = { nullptr, nullptr };
	int64_t n_lines[2] = { 0, 0 };
	const int64_t len[2] = { len1, len2 }, base[2] = { 0, len1 };
	for (int f = 0; f < 1 + paired; ++f) {          // count the line ends first: the list is sized exactly, whatever the bytes are
		k_count_newlines<<<148 * 8, 256, 0, e->stream>>>(text + base[f], len[f], (unsigned long long *)(ctr + 16 + 2 * f));
		CK(cudaGetLastError());
	}
	int64_t h_cnt[2];
	e->d2h(h_cnt, ctr + 16, sizeof h_cnt);
	e->sync();
	for (int f = 0; f < 1 + paired; ++f) {
		nl[f] = e->fb[f ? FB_GSEL2 : FB_GSEL].as<int64_t>((size_t)h_cnt[f] + 16);
		int64_t *d_n = (int64_t *)(ctr + 8 + 2 * f);
		cub::CountingInputIterator<int64_t> it(0);
		IsNewline pred = { text + base[f] };
		size_t tmp = 0;
		CK(cub::DeviceSelect::If(nullptr, tmp, it, nl[f], d_n, len[f], pred, e->stream));
		void *d = e->b_cub.need(tmp);
		CK(cub::DeviceSelect::If(d, tmp, it, nl[f], d_n, len[f], pred, e->stream));
		e->stats.n_launches += 2;
	}
	int64_t h_n[2];
	e->d2h(h_n, ctr + 8, sizeof h_n);
	e->sync();
	n_lines[0] = h_n[0]; n_lines[1] = h_n[1];
	const int64_t n_rec = n_lines[0] >> 2;                    // an unfinished record at the end is dropped, as the hosts' loop does
	if (paired && (n_lines[1] >> 2) != n_rec) die("the two fastq buffers hold different numbers of reads");
	const int64_t n = paired ? 2 * n_rec : n_rec;
	if (n > 0x7fffffff) die("too many reads in one chunk");
	e->n_reads = (int)n;
	ReadText *rt = e->d_rtext.as<ReadText>(n + 1);
	int64_t *seq_at = e->fb[FB_O_Z].as<int64_t>(n + 1);
	int32_t *l_seq = e->fb[FB_LEN].as<int32_t>(n + 1);
	int64_t *off = e->d_off.as<int64_t>(n + 1);
	FastqView v = { text, { nl[0], nl[1] }, { base[0], base[1] }, paired };
	bk.run(n, FastqReadTask{ v, rt, seq_at, l_seq, ctr });
	CK(cudaMemsetAsync(l_seq + n, 0, sizeof(int32_t), e->stream));
	exclusive_scan(e, l_seq, off, n + 1);
	int64_t total = 0;
	int32_t max_len = 0;
	e->d2h(&total, off + n, sizeof total);
	e->d2h(&max_len, ctr, sizeof max_len);
	e->sync();
	uint8_t *codes = e->d_codes.as<uint8_t>((size_t)total + 16);
	if (n > 0) {
		k_fastq_encode<<<grid_for(n * 32, 256), 256, 0, e->stream>>>((int)n, text, seq_at, off, codes);
		CK(cudaGetLastError());
		e->stats.n_launches += 1;
	}
	e->sync();
	e->max_len = max_len;
	e->h_off.clear();
	info->n_reads = (int)n; info->n_bases = total; info->max_len = max_len;
}

// the caller's jobs through the CIGAR-stage kernels: same classification (GlobalClassTask + radix sort), same launches, same rerun
// pass for jobs whose band outgrows the row window of their class as finish_run() uses
void stage_global_batch(Engine *e, const GlobalOpt &go, int64_t n, b200_global_job_t *jobs, const uint8_t *query, int64_t qbytes,
                        const uint8_t *target, int64_t tbytes, std::vector<uint32_t> &cigar)
{
	CK(cudaSetDevice(e->device));
	cigar.clear();
	if (n <= 0) return;
	CudaBK bk = { e };
	std::vector<GlobalJob> hj(n);
	std::vector<int64_t> off(n + 1);
	int64_t zb = 0, cb = 0;
	for (int64_t i = 0; i < n; ++i) {
		const b200_global_job_t &j = jobs[i];
		GlobalJob &g = hj[i];
		g.rb = j.t_off; g.re = j.t_off + j.tlen; g.zoff = zb; g.read = (int32_t)i; g.qb = 0; g.qe = j.qlen; g.w2 = j.w; g.truesc = B200_GLOBAL_RAW;
		g.wmax = j.w; g.cig_off = cb;
		const int n_col = ((j.qlen < 2 * j.w + 1 ? j.qlen : 2 * j.w + 1) + 3) & ~3;
		zb += ((int64_t)n_col * j.tlen + 15) & ~(int64_t)15;
		cb += (int64_t)j.qlen + j.tlen + 4;
		off[i] = j.q_off;
	}
	off[n] = qbytes;
	std::vector<uint8_t> pac((size_t)tbytes / 4 + 2, 0);
	for (int64_t l = 0; l < tbytes; ++l) pac[l >> 2] |= (uint8_t)((target[l] & 3) << ((~l & 3) << 1));
	GlobalJob *dj = bk.buf<GlobalJob>(FB_GJOBS, n);
	GlobalRes *dr = bk.buf<GlobalRes>(FB_GRES, n);
	uint8_t *z = bk.buf<uint8_t>(FB_GZ, (size_t)zb + 64);
	uint32_t *cig = bk.buf<uint32_t>(FB_CIG, (size_t)cb + 4);
	uint32_t *key = bk.buf<uint32_t>(FB_GKEY, n);
	int32_t *sel = bk.buf<int32_t>(FB_GSEL, n), *sel2 = bk.buf<int32_t>(FB_GSEL2, n);
	int32_t *ctr = bk.buf<int32_t>(FB_CTR, 64);
	int64_t *d_off = e->b_soff.as<int64_t>(n + 2);
	uint8_t *d_q = e->b_q.as<uint8_t>((size_t)qbytes + 16), *d_pac = e->b_t.as<uint8_t>(pac.size() + 16);
	e->h2d(dj, hj.data(), sizeof(GlobalJob) * n);
	e->h2d(d_off, off.data(), sizeof(int64_t) * (n + 1));
	e->h2d(d_q, query, (size_t)qbytes);
	e->h2d(d_pac, pac.data(), pac.size());
	bk.zero(ctr, 64 * sizeof(int32_t));
	bk.run(n, IotaTask{ sel });
	int64_t m = n;
	const int squeeze = getenv("B200_GLOBAL_SQUEEZE") != nullptr;
	e->zero_counters();
	for (int pass = 0; pass < 2 && m > 0; ++pass) {
		int32_t *sl = pass == 0 ? sel : sel2;
		bk.zero(ctr + 8, 16 * sizeof(int32_t));
		bk.run(m, GlobalClassTask{ go, dj, sl, key, ctr + 8, pass, squeeze });
		bk.sort_pairs(key, sl, m);
		int32_t h[12];
		bk.download(h, ctr + 8, sizeof h);
		global_launch_classes(e, go, dj, sl, h, h + 6, d_off, d_q, z, cig, dr, d_pac, tbytes);
		if (pass == 0) {
			bk.zero(ctr + 2, sizeof(int32_t));
			bk.run(n, GlobalRerunTask{ dr, sel2, ctr + 2 });
			m = bk.get32(ctr + 2);
		}
	}
	std::vector<GlobalRes> hr(n);
	std::vector<uint32_t> arena((size_t)cb);
	e->d2h(hr.data(), dr, sizeof(GlobalRes) * n);
	e->d2h(arena.data(), cig, sizeof(uint32_t) * cb);
	Counters c = e->read_counters();
	e->stats.global_cells += (int64_t)c.global_cells;
	e->stats.n_global_jobs += n;
	for (int64_t i = 0; i < n; ++i) {
		jobs[i].score = hr[i].score; jobs[i].n_cigar = hr[i].n_cigar; jobs[i].cigar_off = (int64_t)cigar.size();
		cigar.insert(cigar.end(), arena.begin() + hj[i].cig_off, arena.begin() + hj[i].cig_off + (hr[i].n_cigar > 0 ? hr[i].n_cigar : 0));
	}
}

void stage_upload_text(Engine *e, int n_reads, const ReadText *rtext, const char *text, int64_t bytes)
{
	CK(cudaSetDevice(e->device));
	if (n_reads != e->n_reads) die("stage_upload_text: does not match the uploaded reads");
	e->h2d(e->d_rtext.as<ReadText>(n_reads + 1), rtext, sizeof(ReadText) * n_reads);
	e->h2d(e->d_text.as<char>(bytes + 16), text, (size_t)bytes);
	e->sync();
}

void stage_finish(Engine *e, const FinishArgs &a)
{
	CK(cudaSetDevice(e->device));
	CudaBK bk = { e };
	FinCtx cx;
	memset(&cx, 0, sizeof cx);
	cx.opt = *a.opt;
	cx.fm = e->fm;
	cx.ctg_alt = (const uint8_t *)e->d_ctg_alt;
	cx.ctg_name_off = (const int64_t *)e->d_ctg_name_off; cx.ctg_names = (const char *)e->d_ctg_names;
	cx.ctg_anno_off = (const int64_t *)e->d_ctg_anno_off; cx.ctg_annos = (const char *)e->d_ctg_annos;
	cx.n_reads = e->n_reads; cx.pe = (a.opt->flag & MEM_F_PE) ? 1 : 0;
	cx.n_processed = a.n_processed;
	cx.off = (const int64_t *)e->d_off.p; cx.codes = (const uint8_t *)e->d_codes.p;
	cx.rtext = (const ReadText *)e->d_rtext.p; cx.text = (const char *)e->d_text.p;
	cx.rg_len = a.rg_id ? (int)strnlen(a.rg_id, 255) : 0;
	if (cx.rg_len) memcpy(cx.rg_id, a.rg_id, cx.rg_len);
	FinishIn in = { (const DReg *)e->b_xout.p, (const int64_t *)e->b_soff.p, a.pes0, e->max_len, (const double *)e->d_logtab, e->n_log, a.route };
	FinishOut fo;
	e->zero_counters();
	const double k_sw0 = e->stats.ms_k_sw, k_gl0 = e->stats.ms_k_global;
	finish_run(bk, cx, in, fo, e->stats, fin_clock_ms);
	e->sync();
	e->fin_out = fo;
	e->stats.ms_k_finish += bk.task_ms();
	(void)k_sw0; (void)k_gl0;
	Counters c = e->read_counters();
	e->stats.sw_cells += (int64_t)c.sw_cells;
	e->stats.global_cells += (int64_t)c.global_cells;
	e->stats.sam_bytes += fo.sam_bytes;
}

void stage_fetch_sam(Engine *e, const FinishArgs &a, SamChunk &out)
{
	CK(cudaSetDevice(e->device));
	const FinishOut fo = e->fin_out;
	if (getenv("B200_DEBUG")) {          // device memory at the end of a chunk: the index, its tables and every slot's buffers are allocated by now
		size_t free_b = 0, total_b = 0;
		CK(cudaMemGetInfo(&free_b, &total_b));
		fprintf(stderr, "[mem] %.1f of %.1f GB of device memory in use\n", (total_b - free_b) / 1e9, total_b / 1e9);
	}
	// the text and (on request) the per-read offsets come back in page-locked memory
	const double t0 = fin_clock_ms();
	const bool routed = fo.routed != nullptr;
	const char *d_text = routed ? fo.routed : fo.sam;
	const int64_t bytes = routed ? fo.routed_bytes : fo.sam_bytes;
	char *h_sam = a.alloc ? (char *)a.alloc((size_t)bytes + 1) : (char *)e->h_sam.need((size_t)bytes + 16);
	e->d2h(h_sam, d_text, (size_t)bytes);
	int64_t *h_off = nullptr;
	if (a.want_offsets) {
		h_off = (int64_t *)e->h_sam_off.need(sizeof(int64_t) * (e->n_reads + 1));
		e->d2h(h_off, fo.sam_off, sizeof(int64_t) * (e->n_reads + 1));
	}
	SamLine *h_lines = nullptr;
	int64_t *h_dest = nullptr;
	const int n_dest = e->fm.n_ctg + 2;
	if (a.route && !routed && fo.lines) {
		h_lines = (SamLine *)e->h_lines.need(sizeof(SamLine) * (fo.n_lines + 1));
		e->d2h(h_lines, fo.lines, sizeof(SamLine) * fo.n_lines);
	}
	if (routed) {
		h_dest = (int64_t *)e->h_dest_off.need(sizeof(int64_t) * (n_dest + 1));
		e->d2h(h_dest, fo.dest_off, sizeof(int64_t) * (n_dest + 1));
	}
	e->sync();
	h_sam[bytes] = 0;
	e->stats.ms_deliver += fin_clock_ms() - t0;
	out.sam = h_sam; out.sam_off = h_off; out.bytes = bytes; out.lines = h_lines; out.n_lines = h_lines ? fo.n_lines : 0;
	out.dest_off = h_dest; out.n_dest = n_dest;
}

void *stage_host_alloc(size_t bytes)
{
	void *p = nullptr;
	CK(cudaHostAlloc(&p, bytes, cudaHostAllocPortable));
	return p;
}
void stage_host_free(void *p) { if (p) cudaFreeHost(p); }

/* ------------------------------------------------------------------ int32 issue-rate micro-benchmark (roofline denominator)
 * Measures the int32 instruction issue rate of this GPU ("measure it with a dependent-chain-free
 * micro-benchmark on the box").  Eight independent chains per thread, fully unrolled, operands from memory.
 * MODE 0: every chain alternates max / xor - both issue on the integer ALU pipe and have no fused 3-input form, so one
 *         PTX op is one SASS instruction: this is the issue rate available to min/max/add/select, the ops of the DP.
 * MODE 1: half the chains are integer multiply-adds (FMA pipe) - the dual-pipe ceiling when adds go through IMAD. */
template <int MODE>
__global__ void __launch_bounds__(256) k_int32_peak(const int *__restrict__ in, int *out, int iters)
{
	int x0 = in[threadIdx.x], x1 = x0 ^ 1, x2 = x0 ^ 2, x3 = x0 ^ 3, x4 = x0 ^ 4, x5 = x0 ^ 5, x6 = x0 ^ 6, x7 = x0 ^ 7;
	const int a = in[256], b = in[257], one = in[258];
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < 16; ++u) {
#define B200_MX(x) asm volatile("max.s32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(x) : "r"(a), "r"(b))
#define B200_MD(x) asm volatile("mad.lo.s32 %0, %0, %1, %2;\n\tmad.lo.s32 %0, %0, %1, %3;" : "+r"(x) : "r"(one), "r"(a), "r"(b))
			B200_MX(x0); B200_MX(x1); B200_MX(x2); B200_MX(x3);
			if (MODE == 0) { B200_MX(x4); B200_MX(x5); B200_MX(x6); B200_MX(x7); }
			else { B200_MD(x4); B200_MD(x5); B200_MD(x6); B200_MD(x7); }
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

double int32_peak_gops(int device, int mode)
{
	CK(cudaSetDevice(device));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, device));
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
	int *in = nullptr, *out = nullptr;
	CK(cudaMalloc(&in, 259 * sizeof(int)));
	CK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(int)));
	int h[259];
	for (int i = 0; i < 258; ++i) h[i] = i * 7 + 1;
	h[258] = 1;
	CK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float best = 1e30f;
	for (int rep = 0; rep < 6; ++rep) {
		CK(cudaEventRecord(e0));
		if (mode == 0) k_int32_peak<0><<<blocks, threads>>>(in, out, iters);
		else k_int32_peak<1><<<blocks, threads>>>(in, out, iters);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (rep > 0 && ms < best) best = ms;
	}
	CK(cudaFree(in)); CK(cudaFree(out));
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	double ops = (double)blocks * threads * iters * 16.0 * 16.0;
	return ops / (best * 1e-3) / 1e9;
}

/* ------------------------------------------------------------------ random-sector bandwidth micro-benchmark
 * The FM-index kernels read 32-byte occ sectors at random places of a table far larger than L2.  HBM delivers much less than its
 * streaming (copy) bandwidth to such a pattern - every sector opens a DRAM page for 32 useful bytes - so the copy bandwidth of
 * MEASURED_PEAKS.json is not what bounds them.  This measures what the device delivers to the access pattern itself: independent
 * 256-bit loads (the seeding kernels' own instruction: ld.global.nc.L1::no_allocate.v8.u32) of uniformly random sectors of a
 * table of `bytes`, four in flight per thread, 2048 threads per SM. */
__global__ void __launch_bounds__(128) k_random_sectors(const uint32_t *__restrict__ tab, uint64_t n_sec, int iters, uint32_t *out)
{
	uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
	uint32_t acc = 0;
	for (int i = 0; i < iters; ++i) {
		uint32_t w[4][8];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			x = x * 6364136223846793005ull + 1442695040888963407ull;
			const uint64_t sec = __umul64hi(x, n_sec);
			const uint32_t *p = tab + (sec << 3);
			asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			    : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]), "=r"(w[u][4]), "=r"(w[u][5]), "=r"(w[u][6]), "=r"(w[u][7]) : "l"(p));
		}
#pragma unroll
		for (int u = 0; u < 4; ++u) acc ^= w[u][0] ^ w[u][1] ^ w[u][2] ^ w[u][3] ^ w[u][4] ^ w[u][5] ^ w[u][6] ^ w[u][7];
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

double random_sector_gbs(int device, size_t bytes)
{
	CK(cudaSetDevice(device));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, device));
	const int blocks = prop.multiProcessorCount * 16, threads = 128, iters = 256;
	const uint64_t n_sec = bytes / 32;
	uint32_t *tab = nullptr, *out = nullptr;
	CK(cudaMalloc(&tab, n_sec * 32));
	CK(cudaMemset(tab, 1, n_sec * 32));
	CK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(uint32_t)));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float best = 1e30f;
	for (int rep = 0; rep < 4; ++rep) {
		CK(cudaEventRecord(e0));
		k_random_sectors<<<blocks, threads>>>(tab, n_sec, iters, out);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (rep > 0 && ms < best) best = ms;
	}
	CK(cudaFree(tab)); CK(cudaFree(out));
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return (double)blocks * threads * iters * 4.0 * 32.0 / (best * 1e-3) / 1e9;
}

} // namespace b200

extern "C" double b200_hbm_random_sector_peak(int device, size_t table_bytes) { return b200::random_sector_gbs(device, table_bytes); }
extern "C" double b200_int32_peak(int device) { return b200::int32_peak_gops(device, 0); }
extern "C" double b200_int32_peak_dual_pipe(int device) { return b200::int32_peak_gops(device, 1); }
