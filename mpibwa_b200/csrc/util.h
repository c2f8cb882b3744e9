// util.h - small host-side building blocks whose exact behaviour is part of the output contract.
//
//  * tie_sort():  BWA's sorts are unstable and the order of equal keys decides which chain / region survives
//                 (SURVEY.md App. C-1).  tie_sort reproduces the permutation of klib's ks_introsort
//                 (reference src/ksort.h:162-214: median-of-3 quicksort on ranges > 16, comb sort when the
//                 depth budget runs out, one final insertion sort) so that ties land where the reference puts them.
//  * PosTree:     the order-5 B-tree that mem_chain() keys by chain position (reference src/kbtree.h, t = 5 for
//                 a 40-byte key and KB_DEFAULT_SIZE 512).  With duplicate positions the predecessor that is found
//                 and the in-order traversal depend on the node splits, so the tree is rebuilt node for node.
//  * mix64():     Thomas Wang's 64-bit mix used for tie-break ids (reference src/utils.h:98-109).
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <thread>
#include <atomic>
#include <utility>
#include <functional>
#include <memory>
#include <mutex>
#include <condition_variable>

namespace b200 {

static inline uint64_t mix64(uint64_t key)
{
	key += ~(key << 32);
	key ^= (key >> 22);
	key += ~(key << 13);
	key ^= (key >> 8);
	key += (key << 3);
	key ^= (key >> 15);
	key += ~(key << 27);
	key ^= (key >> 31);
	return key;
}

namespace detail {
template <class T, class Less>
static inline void insertion_pass(T *s, T *t, Less lt)
{
	for (T *i = s + 1; i < t; ++i)
		for (T *j = i; j > s && lt(*j, *(j - 1)); --j) std::swap(*j, *(j - 1));
}

template <class T, class Less>
static void comb_pass(size_t n, T *a, Less lt)
{
	const double shrink = 1.2473309501039786540366528676643;
	size_t gap = n;
	bool swapped;
	do {
		if (gap > 2) {
			gap = (size_t)(gap / shrink);
			if (gap == 9 || gap == 10) gap = 11;
		}
		swapped = false;
		for (T *i = a; i < a + n - gap; ++i) {
			T *j = i + gap;
			if (lt(*j, *i)) { std::swap(*i, *j); swapped = true; }
		}
	} while (swapped || gap > 2);
	if (gap != 1) insertion_pass(a, a + n, lt);
}
} // namespace detail

template <class T, class Less>
void tie_sort(size_t n, T *a, Less lt)
{
	if (n < 1) return;
	if (n == 2) { if (lt(a[1], a[0])) std::swap(a[0], a[1]); return; }
	struct Frame { T *lo, *hi; int depth; };
	int d;
	for (d = 2; (1ul << d) < n; ++d) {}
	std::vector<Frame> stack;
	stack.reserve(sizeof(size_t) * d + 2);
	T *s = a, *t = a + (n - 1);
	d <<= 1;
	for (;;) {
		if (s < t) {
			if (--d == 0) { detail::comb_pass((size_t)(t - s + 1), s, lt); t = s; continue; }
			T *i = s, *j = t, *k = i + ((j - i) >> 1) + 1;
			if (lt(*k, *i)) { if (lt(*k, *j)) k = j; }
			else k = lt(*j, *i) ? i : j;
			T pivot = *k;
			if (k != t) std::swap(*k, *t);
			for (;;) {
				do ++i; while (lt(*i, pivot));
				do --j; while (i <= j && lt(pivot, *j));
				if (j <= i) break;
				std::swap(*i, *j);
			}
			std::swap(*i, *t);
			if (i - s > t - i) {
				if (i - s > 16) stack.push_back({s, i - 1, d});
				s = t - i > 16 ? i + 1 : t;
			} else {
				if (t - i > 16) stack.push_back({i + 1, t, d});
				t = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (stack.empty()) { detail::insertion_pass(a, a + n, lt); return; }
			Frame f = stack.back(); stack.pop_back();
			s = f.lo; t = f.hi; d = f.depth;
		}
	}
}

template <class T, class Less>
inline void tie_sort(std::vector<T> &v, Less lt) { tie_sort(v.size(), v.data(), lt); }

// Order-5 B-tree over integer handles, ordered by an external int64 position table.
class PosTree {
public:
	static const int T = 5, MAXK = 2 * T - 1;
	explicit PosTree(const std::vector<int64_t> *pos) : pos_(pos) { nodes_.reserve(16); root_ = new_node(false); }
	int size() const { return n_keys_; }
	void clear() { nodes_.clear(); n_keys_ = 0; root_ = new_node(false); }

	// handle of the closest entry at or below `p` (the `lower` of kb_intervalp), -1 if none
	int lower(int64_t p) const
	{
		int x = root_, low = -1;
		while (x >= 0) {
			int r = 0, i = locate(nodes_[x], p, &r);
			if (i >= 0 && r == 0) return nodes_[x].key[i];
			if (i >= 0) low = nodes_[x].key[i];
			if (!nodes_[x].internal) return low;
			x = nodes_[x].child[i + 1];
		}
		return low;
	}

	void insert(int handle)
	{
		++n_keys_;
		int r = root_;
		if (nodes_[r].n == MAXK) {
			int s = new_node(true);
			nodes_[s].child[0] = r;
			root_ = s;
			split(s, 0, r);
			r = s;
		}
		insert_nonfull(r, handle);
	}

	template <class F> void in_order(F f) const { walk(root_, f); }

private:
	struct Node { bool internal; int n; int key[MAXK]; int child[MAXK + 1]; };
	const std::vector<int64_t> *pos_;
	std::vector<Node> nodes_;
	int root_ = -1, n_keys_ = 0;

	int new_node(bool internal)
	{
		Node z; z.internal = internal; z.n = 0;
		for (int i = 0; i <= MAXK; ++i) z.child[i] = -1;
		nodes_.push_back(z);
		return (int)nodes_.size() - 1;
	}
	int64_t P(int handle) const { return (*pos_)[handle]; }
	// index of the first key equal to p (r = 0) or of the last key below it (r > 0); -1 if all keys are above
	int locate(const Node &x, int64_t p, int *r) const
	{
		if (x.n == 0) return -1;
		int lo = 0, hi = x.n;
		while (lo < hi) {
			int mid = (lo + hi) >> 1;
			if (P(x.key[mid]) < p) lo = mid + 1; else hi = mid;
		}
		if (lo == x.n) { if (r) *r = 1; return x.n - 1; }
		int64_t q = P(x.key[lo]);
		int c = (p > q) - (p < q);
		if (r) *r = c;
		return c < 0 ? lo - 1 : lo;
	}
	void split(int xi, int i, int yi)
	{
		int zi = new_node(nodes_[yi].internal);
		Node &x = nodes_[xi], &y = nodes_[yi], &z = nodes_[zi];
		z.n = T - 1;
		for (int k = 0; k < T - 1; ++k) z.key[k] = y.key[T + k];
		if (y.internal) for (int k = 0; k < T; ++k) z.child[k] = y.child[T + k];
		y.n = T - 1;
		for (int k = x.n; k > i; --k) x.child[k + 1] = x.child[k];
		x.child[i + 1] = zi;
		for (int k = x.n - 1; k >= i; --k) x.key[k + 1] = x.key[k];
		x.key[i] = y.key[T - 1];
		++x.n;
	}
	void insert_nonfull(int xi, int handle)
	{
		int64_t p = P(handle);
		if (!nodes_[xi].internal) {
			Node &x = nodes_[xi];
			int i = locate(x, p, nullptr);
			for (int k = x.n - 1; k > i; --k) x.key[k + 1] = x.key[k];
			x.key[i + 1] = handle;
			++x.n;
		} else {
			int i = locate(nodes_[xi], p, nullptr) + 1;
			if (nodes_[nodes_[xi].child[i]].n == MAXK) {
				split(xi, i, nodes_[xi].child[i]);
				if (p > P(nodes_[xi].key[i])) ++i;
			}
			insert_nonfull(nodes_[xi].child[i], handle);
		}
	}
	template <class F> void walk(int xi, F &f) const
	{
		if (xi < 0) return;
		const Node &x = nodes_[xi];
		for (int i = 0; i < x.n; ++i) {
			if (x.internal) walk(x.child[i], f);
			f(x.key[i]);
		}
		if (x.internal) walk(x.child[x.n], f);
	}
};

// Dynamic parallel loop over [0,n) in blocks of `grain`: body(tid, begin, end) with tid < n_threads, at most one block per
// tid at a time.  All loops of the process - several sub-batch lanes and chunk jobs run their host stages concurrently -
// share ONE pool of n_threads - 1 workers (the calling thread takes part as tid 0 of its own loop), so the number of
// runnable host threads stays at the core count however many chunks are in flight; the threads that drive the device
// stages are then not starved by an oversubscribed run queue.
namespace detail {
struct PoolLoop {
	std::function<void(int, int64_t, int64_t)> body;
	int64_t n, grain;
	int nt;                                  // workers with index >= nt stay out (results never depend on it; tid-indexed scratch does)
	std::atomic<int64_t> next{0}, done{0};
	int64_t n_blocks;
};
class WorkerPool {
public:
	static WorkerPool &get() { static WorkerPool p; return p; }
	void run(int n_threads, int64_t n, int64_t grain, std::function<void(int, int64_t, int64_t)> body)
	{
		auto loop = std::make_shared<PoolLoop>();
		loop->body = std::move(body); loop->n = n; loop->grain = grain; loop->nt = n_threads;
		loop->n_blocks = (n + grain - 1) / grain;
		{
			std::lock_guard<std::mutex> lk(mu_);
			while ((int)workers_.size() < n_threads - 1) { const int id = (int)workers_.size() + 1; workers_.emplace_back([this, id] { work(id); }); }
			loops_.push_back(loop);
		}
		cv_.notify_all();
		drain(*loop, 0);
		std::unique_lock<std::mutex> lk(mu_);
		for (size_t k = 0; k < loops_.size(); ++k) if (loops_[k] == loop) { loops_.erase(loops_.begin() + k); break; }
		done_cv_.wait(lk, [&] { return loop->done.load() == loop->n_blocks; });
	}
private:
	std::mutex mu_;
	std::condition_variable cv_, done_cv_;
	std::vector<std::shared_ptr<PoolLoop>> loops_;      // loops that may still have unclaimed blocks
	std::vector<std::thread> workers_;
	bool stop_ = false;
	~WorkerPool()
	{
		{ std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
		cv_.notify_all();
		for (auto &t : workers_) t.join();
	}
	void drain(PoolLoop &L, int tid)
	{
		for (;;) {
			const int64_t b = L.next.fetch_add(L.grain);
			if (b >= L.n) return;
			const int64_t e = b + L.grain < L.n ? b + L.grain : L.n;
			L.body(tid, b, e);
			if (L.done.fetch_add(1) + 1 == L.n_blocks) { std::lock_guard<std::mutex> lk(mu_); done_cv_.notify_all(); }
		}
	}
	void work(int id)
	{
		std::unique_lock<std::mutex> lk(mu_);
		for (;;) {
			std::shared_ptr<PoolLoop> pick;
			for (auto &l : loops_) if (id < l->nt && l->next.load() < l->n) { pick = l; break; }
			if (!pick) { if (stop_) return; cv_.wait(lk); continue; }
			lk.unlock();
			drain(*pick, id);
			lk.lock();
		}
	}
};
} // namespace detail

template <class F>
void parallel_for(int n_threads, int64_t n, int64_t grain, F body)
{
	if (n <= 0) return;
	if (n_threads <= 1 || n <= grain) { body(0, (int64_t)0, n); return; }
	detail::WorkerPool::get().run(n_threads, n, grain, std::function<void(int, int64_t, int64_t)>(std::ref(body)));
}

} // namespace b200
