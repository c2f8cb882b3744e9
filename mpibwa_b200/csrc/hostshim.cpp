// hostshim.cpp - the few host-side lines of the mpiBWA mains that sit directly around mem_process_seqs, exported
// through the C ABI so that test and bench harnesses (Python/ctypes, tools/b200_driver.c) drive the alignment core
// exactly like the MPI hosts do:
//   b200_fastq_parse    in-place fastq parse into bseq1_t      reference src/mainParallel.c:1257-1304
//   b200_plan_chunks    "close the chunk when bases > maxsiz"  reference src/parallel_aux.c:1532-1549 (same-size
//                       pairs: R1 bases vs K/2), :1068-1082 (trimmed pairs: R1+R2 vs K), src/mainParallel.c:2773 (SE)
//   b200_chunk_seqs / b200_collect_sam   interleave mates; concatenate seqs[i].sam in input order
//                       reference src/mainParallel.c:1271-1314 and copy_buffer_thr :103-127 (b200_align_chunk itself: capi.cpp)
#include "../../include/mpibwa_b200.h"
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <mutex>
#include <thread>
#include <algorithm>

namespace b200 { void *stage_host_alloc(size_t bytes); void stage_host_free(void *p); }      // stages.h: page-locked on the CUDA engine

// number of host threads used by the shim's own loops (parse, SAM concatenation); 0 = all hardware threads
static int g_shim_threads = 0;
static int shim_threads()
{
	int n = g_shim_threads > 0 ? g_shim_threads : (int)std::thread::hardware_concurrency();
	return n > 0 ? n : 1;
}

template <class F>
static void shim_parallel(int nt, F body)      // body(t) for t in [0, nt)
{
	std::vector<std::thread> pool;
	for (int t = 1; t < nt; ++t) pool.emplace_back(body, t);
	body(0);
	for (auto &th : pool) th.join();
}

extern "C" {

void b200_set_host_threads(int n) { g_shim_threads = n; }

// Same result as the hosts' serial loop (4 lines per record, '\n' -> NUL, name cut at the first blank and stripped of a
// trailing "/[0-9]"), done in two parallel sweeps: count the line ends per segment, then every segment fills the
// records whose lines END inside it.
int64_t b200_fastq_parse(char *buf, int64_t len, bseq1_t **out)
{
	const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(shim_threads(), len >> 20));
	std::vector<int64_t> seg(nt + 1), first_line(nt + 1, 0);
	for (int t = 0; t <= nt; ++t) seg[t] = len * t / nt;
	shim_parallel(nt, [&](int t) {
		int64_t c = 0;
		const char *p = buf + seg[t], *e = buf + seg[t + 1];
		while (p < e && (p = (const char *)memchr(p, '\n', (size_t)(e - p))) != nullptr) { ++c; ++p; }
		first_line[t + 1] = c;
	});
	for (int t = 0; t < nt; ++t) first_line[t + 1] += first_line[t];
	const int64_t n = first_line[nt] >> 2;          // an unfinished record at the end is dropped, as the serial loop does
	bseq1_t *s = (bseq1_t *)malloc((size_t)(n + 1) * sizeof(bseq1_t));
	shim_parallel(nt, [&](int t) {
		int64_t line = first_line[t];
		char *q = buf + seg[t], *e = buf + seg[t + 1];
		char *p = q;                                  // start of the line that contains q: just after the previous line end
		while (p > buf && p[-1] != '\n' && p[-1] != 0) --p;
		while (q < e && (q = (char *)memchr(q, '\n', (size_t)(e - q))) != nullptr) {
			const int64_t rec = line >> 2;
			if (rec < n) {
				*q = 0;
				bseq1_t &r = s[rec];
				switch (line & 3) {
				case 0: {
					r.id = 0; r.comment = nullptr; r.sam = nullptr;        // seq/l_seq/qual come from the other lines
					r.name = p + 1;
					char *z = p;
					while (*z && !isspace((unsigned char)*z)) ++z;
					if (z - 2 > r.name && *(z - 2) == '/' && isdigit((unsigned char)*(z - 1))) *(z - 2) = 0;
					if (*z) *z = 0;
					break;
				}
				case 1: r.seq = p; r.l_seq = (int)(q - p); break;
				case 2: break;
				default: r.qual = p; break;
				}
			}
			p = ++q; ++line;
		}
	});
	*out = s;
	return n;
}

int64_t b200_plan_chunks(int64_t n, const bseq1_t *s1, const bseq1_t *s2, int64_t K, int trimmed, int64_t **ends)
{
	std::vector<int64_t> e;
	const int64_t maxsiz = (s2 && !trimmed) ? K / 2 : K;
	int64_t bases = 0;
	for (int64_t i = 0; i < n; ++i) {
		bases += s1[i].l_seq;
		if (s2 && trimmed) bases += s2[i].l_seq;
		if (bases > maxsiz || i + 1 == n) { e.push_back(i + 1); bases = 0; }
	}
	*ends = (int64_t *)malloc((e.size() + 1) * sizeof(int64_t));
	memcpy(*ends, e.data(), e.size() * sizeof(int64_t));
	return (int64_t)e.size();
}

bseq1_t *b200_chunk_seqs(int64_t n, const bseq1_t *s1, const bseq1_t *s2)
{
	const int64_t total = s2 ? 2 * n : n;
	bseq1_t *seqs = (bseq1_t *)malloc((size_t)(total + 1) * sizeof(bseq1_t));
	for (int64_t i = 0; i < n; ++i) {
		if (s2) { seqs[2 * i] = s1[i]; seqs[2 * i + 1] = s2[i]; }
		else seqs[i] = s1[i];
	}
	return seqs;
}

int64_t b200_collect_sam(int64_t total, bseq1_t *seqs, char **sam)
{
	const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(shim_threads(), total >> 12));
	std::vector<size_t> len((size_t)total);
	std::vector<size_t> part(nt + 1, 0);
	shim_parallel(nt, [&](int t) {
		size_t sum = 0;
		for (int64_t i = total * t / nt; i < total * (t + 1) / nt; ++i) { len[i] = seqs[i].sam ? strlen(seqs[i].sam) : 0; sum += len[i]; }
		part[t + 1] = sum;
	});
	for (int t = 0; t < nt; ++t) part[t + 1] += part[t];
	const size_t sum = part[nt];
	char *buf = sam ? (char *)malloc(sum + 1) : nullptr;
	shim_parallel(nt, [&](int t) {
		char *w = buf ? buf + part[t] : nullptr;
		for (int64_t i = total * t / nt; i < total * (t + 1) / nt; ++i) {
			if (buf) { memcpy(w, seqs[i].sam, len[i]); w += len[i]; }
			free(seqs[i].sam);
			seqs[i].sam = nullptr;
		}
	});
	if (buf) { buf[sum] = 0; *sam = buf; }
	return (int64_t)sum;
}

// Large result buffers (the SAM text of a chunk: hundreds of MB) are page-locked - the device copies the text straight into them -
// and recycled instead of returned to the system: pinning and unpinning that much memory per chunk costs milliseconds and a TLB
// shoot-down on every core of a busy host.  b200_big_alloc() hands out a buffer of at least `bytes`; b200_free() recognises such a
// buffer and parks it for the next call (at most sixteen stay parked).
static struct BigPool { std::mutex mu; struct Ent { void *p; size_t cap; bool busy; }; std::vector<Ent> ents; } g_big;

void *b200_big_alloc(size_t bytes)
{
	{
		std::lock_guard<std::mutex> lk(g_big.mu);
		// best fit among the parked buffers, and never a buffer more than twice the size asked for: the pool serves two very
		// different sizes (fastq chunks, SAM chunks) and a small request must not walk away with a large buffer
		BigPool::Ent *best = nullptr;
		for (auto &e : g_big.ents) if (!e.busy && e.cap >= bytes && e.cap <= 2 * bytes + (1 << 20) && (!best || e.cap < best->cap)) best = &e;
		if (best) { best->busy = true; return best->p; }
	}
	const size_t cap = bytes + (bytes >> 3) + 4096;
	void *p = b200::stage_host_alloc(cap);
	std::lock_guard<std::mutex> lk(g_big.mu);
	g_big.ents.push_back({ p, cap, true });
	return p;
}

void b200_free(void *p)
{
	if (!p) return;
	{
		std::lock_guard<std::mutex> lk(g_big.mu);
		size_t parked = 0;
		for (auto &e : g_big.ents) parked += !e.busy;
		for (size_t k = 0; k < g_big.ents.size(); ++k)
			if (g_big.ents[k].p == p) {
				if (parked >= 16) { b200::stage_host_free(p); g_big.ents.erase(g_big.ents.begin() + k); }
				else g_big.ents[k].busy = false;
				return;
			}
	}
	free(p);
}

} // extern "C"
