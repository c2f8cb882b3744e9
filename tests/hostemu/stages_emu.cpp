// TEST-ONLY implementation of mpibwa_b200/csrc/stages.h: runs the per-thread task bodies of the device stages in
// plain loops on the CPU so that the host orchestration (chaining, dedup, pairing, SAM text, rescue replay) can be
// checked against the oracle in a container without a GPU.  It is compiled into tests/_build/ only; the shipped
// libmpibwa_b200.so links stages_cuda.cu and has no CPU execution path.
#include "../../mpibwa_b200/csrc/stages.h"
#include "../../mpibwa_b200/csrc/util.h"
#include "../../mpibwa_b200/csrc/smem_kernel.cuh"
#include "../../mpibwa_b200/csrc/smem_sweeps.cuh"
#include "../../mpibwa_b200/csrc/finish_stage.h"
#include "emu_bodies.h"
#include <chrono>
#include <cmath>
#include <string>
#include <cstdio>
#include <memory>
#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace b200 {

class Engine {
public:
	FmView fm;
	std::vector<int64_t> ctg_off;
	std::vector<int32_t> ctg_len;
	Stats stats;
	int n_reads = 0;
	std::vector<int64_t> off;
	std::vector<uint8_t> codes;
	std::vector<uint8_t> staging;
	std::vector<char> pinned[PIN_N_SLOTS];
	std::vector<int64_t> seed_off;
	std::vector<SeedRec> seeds;
	std::vector<int32_t> l_rep;
	std::vector<uint8_t> ctg_alt;
	std::shared_ptr<std::vector<uint32_t>> occ;      // occ sectors built from the reference layout (shared by clones)
	std::shared_ptr<std::vector<uint8_t>> isa5;
	std::shared_ptr<std::vector<uint64_t>> bloom;
	std::shared_ptr<std::vector<uint8_t>> pac_padded;   // the 2-bit text with the slack the 9-byte window loads may read past its end
	std::shared_ptr<std::vector<uint8_t>> sa5;       // the whole suffix array, expanded from the samples like the upload kernel does
	std::shared_ptr<std::vector<Q4>> ktab;           // k-mer interval tables, built level by level with the routine the upload kernel runs
	// finish stages
	std::vector<DReg> xregs;
	std::vector<int64_t> xoff;
	std::vector<char> fb[FB_N];
	std::vector<ReadText> rtext;
	std::vector<char> text, sam;
	std::vector<int64_t> sam_off, ctg_name_off, ctg_anno_off;
	std::string ctg_names, ctg_annos;
	std::shared_ptr<std::vector<double>> logtab;
	int max_len = 0;
	FinishOut fin_out;
	std::vector<SamLine> lines;
	std::vector<int64_t> dest_off;
	int64_t plain_blocks = 0, sweep_blocks = 0;                        // occ blocks the sweeps touched or looked up (fm_occ_blocks counts the plain routine)
};

Engine *engine_create(const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int)
{
	Engine *e = new Engine();
	{	// the device's occ-sector layout, built with the same per-sector routine the upload kernel runs
		const uint64_t n_sec = (bwt->seq_len >> 6) + 2;
		std::vector<uint32_t> ref(bwt->bwt, bwt->bwt + bwt->bwt_size);
		ref.resize(((n_sec >> 1) + 2) * 16, 0);
		e->occ = std::make_shared<std::vector<uint32_t>>(n_sec * 8);
		for (uint64_t b = 0; b < n_sec; ++b) occ_convert_block(ref.data(), b, e->occ->data() + b * 8);
	}
	e->fm.occ = e->occ->data(); e->fm.sa = bwt->sa; e->fm.primary = bwt->primary;
	for (int i = 0; i < 5; ++i) e->fm.L2[i] = bwt->L2[i];
	e->fm.seq_len = bwt->seq_len; e->fm.sa_intv = bwt->sa_intv;
	e->pac_padded = std::make_shared<std::vector<uint8_t>>((size_t)(bns->l_pac / 4 + 1) + 16, 0);
	memcpy(e->pac_padded->data(), pac, (size_t)(bns->l_pac / 4 + 1));
	e->fm.pac = e->pac_padded->data(); e->fm.l_pac = bns->l_pac;
	e->fm.bloom = nullptr; e->fm.bloom_mask = 0; e->fm.bloom_k = 0;
	if (!(getenv("B200_BLOOM") && atoi(getenv("B200_BLOOM")) == 0) && (int64_t)bwt->seq_len > 64) {
		const int K = 19;
		const uint64_t n_words = bloom_words_for(bwt->seq_len);
		e->bloom = std::make_shared<std::vector<uint64_t>>(2 * n_words, 0);
		for (int64_t p = 0; p + K <= (int64_t)bwt->seq_len; ++p) {
			uint64_t word, bits;
			bloom_of_text(pac, bns->l_pac, p, K, n_words - 1, word, bits);
			bloom_insert(e->bloom->data(), word, bits);
		}
		e->fm.bloom = e->bloom->data(); e->fm.bloom_mask = n_words - 1; e->fm.bloom_k = K;
	}
	e->fm.sa5 = nullptr; e->fm.isa5 = nullptr;
	const int sa_full = getenv("B200_SA_FULL") ? atoi(getenv("B200_SA_FULL")) : 2;
	if (sa_full > 0) {
		e->sa5 = std::make_shared<std::vector<uint8_t>>((size_t)(bwt->seq_len + 1) * 5 + 16);
		if (sa_full > 1) e->isa5 = std::make_shared<std::vector<uint8_t>>((size_t)(bwt->seq_len + 1) * 5 + 16);
		for (uint64_t j = 0; j < (uint64_t)bwt->n_sa; ++j) sa5_expand(e->fm, j, e->sa5->data(), e->isa5 ? e->isa5->data() : nullptr);
		// every 61st row (and the rows around the sentinel's) against the reference's walk to a sampled row
		for (uint64_t k = 1; k <= bwt->seq_len; k += (k + 3 > bwt->primary && k < bwt->primary + 3) ? 1 : 61) {
			int st;
			const uint64_t want = fm_sa(e->fm, k, &st);       // (fm.sa5 is still null here: the walk)
			if (sa5_read(e->sa5->data(), k) != want) { fprintf(stderr, "[hostemu] expanded suffix array differs at row %llu\n", (unsigned long long)k); abort(); }
		}
		e->fm.sa5 = e->sa5->data();
		if (e->isa5) {
			for (uint64_t k = 1; k <= bwt->seq_len; k += 61)
				if (sa5_read(e->isa5->data(), sa5_read(e->sa5->data(), k)) != k) { fprintf(stderr, "[hostemu] inverse suffix array differs at row %llu\n", (unsigned long long)k); abort(); }
			e->fm.isa5 = e->isa5->data();
		}
	}
	{
		int kmax = ktab_default_kmax(bwt->seq_len);
		if (getenv("B200_KMER_MAX")) kmax = std::max(0, std::min(16, atoi(getenv("B200_KMER_MAX"))));
		e->ktab = std::make_shared<std::vector<Q4>>(ktab_entries(kmax) + 2);
		e->fm.ktab = (const uint32_t *)e->ktab->data(); e->fm.kmax = kmax;
		if (getenv("B200_DEBUG")) fprintf(stderr, "[hostemu] k-mer interval tables up to %d bases (%llu entries)\n", kmax, (unsigned long long)ktab_entries(kmax));
		for (int L = 1; L <= kmax; ++L)
			for (uint64_t idx = 0; idx < (uint64_t)1 << (2 * L); ++idx) (*e->ktab)[ktab_off(L) + idx] = ktab_make(e->fm, L, (uint32_t)idx);
	}
	for (int i = 0; i < bns->n_seqs; ++i) { e->ctg_off.push_back(bns->anns[i].offset); e->ctg_len.push_back(bns->anns[i].len); }
	for (int i = 0; i < bns->n_seqs; ++i) e->ctg_alt.push_back(bns->anns[i].is_alt ? 1 : 0);
	e->ctg_alt.push_back(0);
	e->fm.ctg_off = e->ctg_off.data(); e->fm.ctg_len = e->ctg_len.data(); e->fm.n_ctg = bns->n_seqs;
	e->ctg_name_off.assign(1, 0); e->ctg_anno_off.assign(1, 0);
	for (int i = 0; i < bns->n_seqs; ++i) {
		e->ctg_names += bns->anns[i].name; e->ctg_name_off.push_back((int64_t)e->ctg_names.size());
		if (bns->anns[i].anno) e->ctg_annos += bns->anns[i].anno;
		e->ctg_anno_off.push_back((int64_t)e->ctg_annos.size());
	}
	e->logtab = std::make_shared<std::vector<double>>(1 << 20);
	for (size_t i = 0; i < e->logtab->size(); ++i) (*e->logtab)[i] = log((double)i);
	memset(static_cast<b200_stats_t *>(&e->stats), 0, sizeof(b200_stats_t));
	return e;
}
Engine *engine_clone(Engine *base)
{
	Engine *e = new Engine();
	e->fm = base->fm;
	e->occ = base->occ; e->ktab = base->ktab; e->sa5 = base->sa5; e->isa5 = base->isa5; e->bloom = base->bloom; e->pac_padded = base->pac_padded;
	e->ctg_off = base->ctg_off; e->ctg_len = base->ctg_len; e->ctg_alt = base->ctg_alt;
	e->fm.ctg_off = e->ctg_off.data(); e->fm.ctg_len = e->ctg_len.data();
	e->ctg_name_off = base->ctg_name_off; e->ctg_anno_off = base->ctg_anno_off; e->ctg_names = base->ctg_names; e->ctg_annos = base->ctg_annos;
	e->logtab = base->logtab;
	memset(static_cast<b200_stats_t *>(&e->stats), 0, sizeof(b200_stats_t));
	return e;
}
void engine_destroy(Engine *e)
{
	if (getenv("B200_DEBUG") && e->plain_blocks) fprintf(stderr, "[hostemu] occ blocks: %lld by the reference's loops, %lld touched or looked up by the sweeps\n", (long long)e->plain_blocks, (long long)e->sweep_blocks);
	delete e;
}
Stats &engine_stats(Engine *e) { return e->stats; }
const char *engine_kind() { return "hostemu"; }
int engine_device_count() { return 1; }

uint8_t *stage_read_buffer(Engine *e, int64_t bytes)
{
	e->staging.resize((size_t)bytes);
	return e->staging.data();
}

void stage_upload_reads(Engine *e, int n_reads, const int64_t *off, const uint8_t *codes)
{
	e->n_reads = n_reads;
	e->off.assign(off, off + n_reads + 1);
	e->codes.assign(codes, codes + off[n_reads]);
	e->codes.resize(off[n_reads] + 8);
	e->max_len = 0;
	for (int i = 0; i < n_reads; ++i) e->max_len = std::max(e->max_len, (int)(off[i + 1] - off[i]));
}

void stage_upload_fastq(Engine *e, const char *fq1, int64_t len1, const char *fq2, int64_t len2, FastqInfo *info)
{
	const int paired = fq2 != nullptr;
	if (!paired) len2 = 0;
	e->text.assign(fq1, fq1 + len1);
	if (paired) e->text.insert(e->text.end(), fq2, fq2 + len2);
	std::vector<int64_t> nl[2];
	const int64_t len[2] = { len1, len2 }, base[2] = { 0, len1 };
	for (int f = 0; f < 1 + paired; ++f)
		for (int64_t i = 0; i < len[f]; ++i) if (e->text[base[f] + i] == '\n') nl[f].push_back(i);
	const int64_t n_rec = (int64_t)nl[0].size() >> 2;
	if (paired && ((int64_t)nl[1].size() >> 2) != n_rec) { fprintf(stderr, "[mpibwa_b200] the two fastq buffers hold different numbers of reads\n"); abort(); }
	const int64_t n = paired ? 2 * n_rec : n_rec;
	e->n_reads = (int)n;
	e->rtext.assign(n + 1, ReadText());
	e->off.assign(n + 1, 0);
	std::vector<int64_t> seq_at(n + 1);
	FastqView v = { e->text.data(), { nl[0].data(), nl[1].data() }, { base[0], base[1] }, paired };
	e->max_len = 0;
	for (int64_t r = 0; r < n; ++r) {
		int32_t l;
		fastq_read(v, r, &e->rtext[r], &seq_at[r], &l);
		e->off[r + 1] = e->off[r] + l;
		e->max_len = std::max(e->max_len, (int)l);
	}
	e->codes.assign(e->off[n] + 8, 0);
	for (int64_t r = 0; r < n; ++r)
		for (int64_t j = 0; j < e->off[r + 1] - e->off[r]; ++j) e->codes[e->off[r] + j] = fq_code((uint8_t)e->text[seq_at[r] + j]);
	info->n_reads = (int)n; info->n_bases = e->off[n]; info->max_len = e->max_len;
}

void stage_upload_text(Engine *e, int n_reads, const ReadText *rtext, const char *text, int64_t bytes)
{
	e->rtext.assign(rtext, rtext + n_reads);
	e->text.assign(text, text + bytes);
}

void stage_collect_intv(Engine *e, const SeedOpt &so, int n_reads, const int64_t *off, const uint8_t *codes,
                        std::vector<int64_t> &intv_off, std::vector<Intv> &intv)
{
	intv_off.assign(n_reads + 1, 0);
	intv.clear();
	std::vector<Intv> scratch, out;
	for (int r = 0; r < n_reads; ++r) {
		int len = (int)(off[r + 1] - off[r]);
		int n = 0;
		const int64_t blocks_before = e->stats.fm_occ_blocks;
		if (len >= so.min_seed_len) {
			scratch.resize(3 * (len + 1));
			int cap = len + 32;
			for (;;) {
				out.resize(cap);
				n = fm_collect_intv(e->fm, so, len, codes + off[r], out.data(), cap, scratch.data(), &e->stats.fm_occ_blocks);
				if (n >= 0) break;
				cap = -n * 2;
			}
		}
		{	// cross-check: the lane state machine of the CUDA seeding kernel (smem_kernel.cuh), run here on the CPU with
			// a tiny shared quota so that the spill path is exercised too, must give the same interval set
			const int quota = 1 + r % 5, cap2 = std::max(len + 32, n + 1);
			std::vector<uint32_t> sh(4 * quota);
			std::vector<Q4> spill(len + 2);
			std::vector<Intv> out2(cap2);
			SeedList L; L.sh = sh.data(); L.stride = 1; L.quota = quota; L.spill = spill.data(); L.sstride = 1;
			SeedLane ln;
			int64_t blocks = 0;
			ln.begin(so, len, codes + off[r], out2.data());
			bool need = ln.advance(e->fm, so, cap2, L);
			while (need) {
				uint64_t o0, o1, o2;
				fm_extend_sel(e->fm, ln.k0, ln.k1, ln.k2, ln.is_back, ln.c, o0, o1, o2, blocks);
				const int slow = (r & 1) ? ln.fast_step(so, L, o0, o1, o2) : 1;     // odd reads also exercise the fast path
				if (slow) {
					if (slow == 1) ln.consume(so, cap2, L, o0, o1, o2);
					need = ln.advance(e->fm, so, cap2, L);
				}
			}
			bool same = ln.n_out == n;
			if (same) {
				std::sort(out2.begin(), out2.begin() + n, [](const Intv &a, const Intv &b) { return a.info < b.info; });
				for (int i = 0; i < n && same; ++i)
					same = out2[i].x0 == out[i].x0 && out2[i].x1 == out[i].x1 && out2[i].x2 == out[i].x2 && out2[i].info == out[i].info;
			}
			if (!same && getenv("B200_EMU_DEBUG")) for (int i = 0; i < n; ++i) fprintf(stderr, "%d: v1 %llu %llu %llu %d-%d | v2 %llu %llu %llu %d-%d\n", i,
				(unsigned long long)out[i].x0, (unsigned long long)out[i].x1, (unsigned long long)out[i].x2, (int)(out[i].info >> 32), (int)(uint32_t)out[i].info,
				(unsigned long long)out2[i].x0, (unsigned long long)out2[i].x1, (unsigned long long)out2[i].x2, (int)(out2[i].info >> 32), (int)(uint32_t)out2[i].info);
			if (same && len >= so.min_seed_len) {     // the four homogeneous sweeps of smem_sweeps.cuh (throughput path of the CUDA stage)
				const int strip_cap = 3 * len + 8;
				std::vector<Q4> strip(strip_cap);
				std::vector<Intv> out3(cap2);
				int n_out = 0, n_first = 0, n_sw = 0;
				int64_t blocks3 = 0;
				const int pk_words = packed_words_for(len);
				std::vector<uint64_t> pk(2 * pk_words);
				for (int w = 0; w < pk_words; ++w) pack_read_word(codes + off[r], len, w, pk[2 * w], pk[2 * w + 1]);
				const PackedRead pr = { pk.data(), pk_words };
				for (int pass = 1; pass <= 2 && n_sw >= 0; ++pass) {
					FwdLane f;
					f.begin(so, e->fm, pass, len, codes + off[r], pr, out3.data(), strip.data(), strip_cap, n_out, pass == 1 ? 0 : n_first);
					bool nd = f.advance(e->fm, so);
					while (nd) {
						uint64_t o0, o1, o2;
						fwd_lane_fetch(e->fm, f, o0, o1, o2, blocks3);
						if (!f.step(e->fm, so, cap2, o0, o1, o2)) nd = f.advance(e->fm, so);
					}
					n_out = f.n_out; n_sw = f.over ? -1 : f.n_sweeps;
					if (pass == 1) n_first = n_out;
					if (n_sw <= 0) continue;
					BwdLane b;
					uint32_t traj[2 * BwdLane::TRAJ];
					b.begin(so, e->fm.kmax, e->fm.bloom && so.min_seed_len >= e->fm.bloom_k ? e->fm.bloom_k : 0, len, codes + off[r], pr, out3.data(), strip.data(), n_sw, n_out, traj, 1);
					nd = b.advance(so, cap2);
					while (nd) {
						uint64_t o0, o1, o2;
						bwd_lane_fetch(e->fm, b, o0, o1, o2, blocks3);
						if (!b.step(so, cap2, o0, o1, o2)) nd = b.advance(so, cap2);
					}
					n_out = b.n_out;
					if (b.over) n_sw = -1;                    // chain budget exceeded: the read goes to the general kernel instead
				}
				bool same3 = n_sw < 0 || n_out == n;      // n_sw < 0: strip overflow, the read goes to the general kernel instead
				if (same3 && n_sw >= 0) {
					std::sort(out3.begin(), out3.begin() + n, [](const Intv &a, const Intv &b) { return a.info < b.info; });
					for (int i = 0; i < n && same3; ++i)
						same3 = out3[i].x0 == out[i].x0 && out3[i].x1 == out[i].x1 && out3[i].x2 == out[i].x2 && out3[i].info == out[i].info;
				}
				if (!same3) { fprintf(stderr, "[hostemu] seeding sweeps disagree with fm_collect_intv on read %d (%d vs %d intervals, sweeps %d)\n", r, n_out, n, n_sw); abort(); }
				e->sweep_blocks += blocks3; e->plain_blocks += e->stats.fm_occ_blocks - blocks_before;
			}
			if (!same) { fprintf(stderr, "[hostemu] seeding state machine disagrees with fm_collect_intv on read %d (%d vs %d intervals)\n", r, ln.n_out, n); abort(); }
		}
		intv.insert(intv.end(), out.begin(), out.begin() + n);
		intv_off[r + 1] = (int64_t)intv.size();
	}
}

void stage_seed(Engine *e, const SeedOpt &so, SeedOut &res, bool)
{
	std::vector<int64_t> &seed_off = e->seed_off;
	std::vector<SeedRec> &seeds = e->seeds;
	std::vector<int32_t> &l_rep = e->l_rep;
	std::vector<int64_t> io;
	std::vector<Intv> iv;
	stage_collect_intv(e, so, e->n_reads, e->off.data(), e->codes.data(), io, iv);
	e->stats.n_intv = (int64_t)iv.size();
	seed_off.assign(e->n_reads + 1, 0);
	l_rep.assign(e->n_reads, 0);
	seeds.clear();
	for (int r = 0; r < e->n_reads; ++r) {
		int b = 0, en = 0, rep = 0;
		for (int64_t i = io[r]; i < io[r + 1]; ++i) {
			const Intv &p = iv[i];
			int sb = (int)(p.info >> 32), se = (int)(uint32_t)p.info;
			if (p.x2 <= (uint64_t)so.max_occ) continue;
			if (sb > en) { rep += en - b; b = sb; en = se; }
			else en = en > se ? en : se;
		}
		rep += en - b;
		l_rep[r] = rep;
		for (int64_t i = io[r]; i < io[r + 1]; ++i) {
			const Intv &p = iv[i];
			int slen = (int)(uint32_t)p.info - (int)(p.info >> 32);
			int cnt = seed_slots(p.x2, so.max_occ);
			uint64_t step = seed_step(p.x2, so.max_occ);
			for (int c = 0; c < cnt; ++c) {
				int steps;
				SeedRec s;
				s.rbeg = (int64_t)fm_sa(e->fm, p.x0 + (uint64_t)c * step, &steps);
				s.qbeg = (uint16_t)(p.info >> 32); s.len = (uint16_t)slen;
				s.rid = fm_intv2rid(e->fm, s.rbeg, s.rbeg + s.len);
				seeds.push_back(s);
				e->stats.fm_sa_steps += steps; ++e->stats.fm_sa_lookups;
			}
		}
		seed_off[r + 1] = (int64_t)seeds.size();
	}
	res.seed_off = seed_off.data(); res.seeds = seeds.data(); res.l_rep = l_rep.data(); res.n_seeds = (int64_t)seeds.size();
}

void *stage_pinned(Engine *e, int slot, size_t bytes)
{
	if (e->pinned[slot].size() < bytes) e->pinned[slot].resize(bytes + bytes / 4 + 64);
	return e->pinned[slot].data();
}

double stage_extend_replay(Engine *, const ExtOpt &, int64_t *cells, int64_t *n_jobs) { *cells = 0; *n_jobs = 0; return 0; }

// the two chaining kernels of stages_cuda.cu as loops over the reads
void stage_chain(Engine *e, const ChainOpt &co, ExtIn &in, bool)
{
	const int n = e->n_reads;
	const int64_t n_in = (int64_t)e->seeds.size();
	std::vector<int32_t> scr((size_t)CH_N_PLANES * (n_in + 1));
	std::vector<BtNode> nodes((size_t)(n_in >> 1) + 2 * (size_t)n + 4);
	ChainScratch S;
	S.scr = scr.data(); S.n_total = n_in + 1; S.nodes = nodes.data(); S.ctg_alt = e->ctg_alt.data();
	std::vector<int32_t> n_kc(n + 1, 0), n_ks(n + 1, 0);
	for (int r = 0; r < n; ++r) {
		int ns = 0;
		n_kc[r] = chain_build_filter(co, e->fm.l_pac, S, r, (int)(e->off[r + 1] - e->off[r]), e->seeds.data(), e->seed_off[r], (int)(e->seed_off[r + 1] - e->seed_off[r]), &ns);
		n_ks[r] = ns;
	}
	std::vector<int64_t> coff(n + 1, 0), soff(n + 1, 0);
	for (int r = 0; r < n; ++r) { coff[r + 1] = coff[r] + n_kc[r]; soff[r + 1] = soff[r] + n_ks[r]; }
	int32_t *h_co = (int32_t *)stage_pinned(e, PIN_CHAIN_OFF, sizeof(int32_t) * (n + 1));
	DChain *h_ch = (DChain *)stage_pinned(e, PIN_CHAINS, sizeof(DChain) * (coff[n] + 1));
	DSeed *h_se = (DSeed *)stage_pinned(e, PIN_DSEEDS, sizeof(DSeed) * (soff[n] + 1));
	int32_t *h_srt = (int32_t *)stage_pinned(e, PIN_SRT, sizeof(int32_t) * (soff[n] + 1));
	for (int r = 0; r <= n; ++r) h_co[r] = (int32_t)coff[r];
	for (int r = 0; r < n; ++r)
		if (n_kc[r]) chain_emit(co, e->fm, S, (int)(e->off[r + 1] - e->off[r]), e->l_rep[r], e->seeds.data(), e->seed_off[r], n_kc[r], coff[r], soff[r], h_ch, h_se, h_srt);
	in.n_reads = n; in.chain_off = h_co; in.chains = h_ch; in.seeds = h_se; in.srt = h_srt;
	in.n_chains = coff[n]; in.n_seeds = soff[n]; in.on_device = false;
}

void stage_extend(Engine *e, const ExtOpt &eo, const ExtIn &in)
{
	const int n = in.n_reads;
	const int32_t *chain_off = in.chain_off;
	const DChain *chains = in.chains;
	std::vector<DReg> regs;
	std::vector<int64_t> reg_off(n + 1, 0);
	std::vector<int32_t> eh;
	std::vector<DReg> tmp;
	for (int r = 0; r < n; ++r) {
		int nc = chain_off[r + 1] - chain_off[r];
		reg_off[r] = (int64_t)regs.size();
		if (nc == 0) continue;
		int l_query = (int)(e->off[r + 1] - e->off[r]);
		eh.resize(2 * (l_query + 2));
		EhStrided acc = { eh.data(), 1 };
		int calls = 0;
		int n_seeds = 0;
		for (int c = 0; c < nc; ++c) n_seeds += chains[chain_off[r] + c].n_seeds;
		tmp.assign(n_seeds + 1, DReg());
		int nr = chain2aln_read(eo, e->fm.pac, e->fm.l_pac, l_query, e->codes.data() + e->off[r], &chains[chain_off[r]], nc,
		                        in.seeds, const_cast<int32_t *>(in.srt), acc, tmp.data(), &e->stats.extend_cells, &calls);
		regs.insert(regs.end(), tmp.begin(), tmp.begin() + nr);
		e->stats.n_extend_jobs += calls;
	}
	reg_off[n] = (int64_t)regs.size();
	regs.push_back(DReg());
	e->xregs.swap(regs); e->xoff.swap(reg_off);
}

void stage_extend_download(Engine *e, ExtRegs &res) { res.regs = e->xregs.data(); res.reg_off = e->xoff.data(); }

void stage_sw(Engine *e, const SwOpt &so, const std::vector<SwJob> &jobs, std::vector<SwRes> &out)
{
	out.resize(jobs.size());
	std::vector<uint16_t> H, E;
	std::vector<uint64_t> b;
	for (size_t x = 0; x < jobs.size(); ++x) {
		const SwJob &j = jobs[x];
		int qpad = j.q_len + 16;
		H.resize(qpad); E.resize(qpad); b.resize(j.tlen / 2 + 2);
		Row16 h = { H.data(), 1 }, ee = { E.data(), 1 };
		List64 bl = { b.data(), 1 };
		STPac ta = { e->fm.pac, e->fm.l_pac, j.rb };
		const uint8_t *q = e->codes.data() + e->off[j.read] + j.q_beg;
		if (j.is_rev) { SQRevComp qa = { q, j.q_len }; sw_align(j.q_len, qa, j.tlen, ta, so, j.xtra, h, ee, bl, &out[x], &e->stats.sw_cells); }
		else { SQFwd qa = { q }; sw_align(j.q_len, qa, j.tlen, ta, so, j.xtra, h, ee, bl, &out[x], &e->stats.sw_cells); }
		++e->stats.n_sw_jobs;
	}
}

int stage_global_bytes(Engine *, const GlobalOpt &go, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int w, std::vector<uint32_t> *cigar)
{
	struct Seqs {
		const uint8_t *q, *t; int l_query;
		int qa(int j) const { return q[j]; }
		int ta(int i) const { return t[i]; }
		const int8_t *trow(const GlobalOpt &o, int i) const { return o.mat + t[i] * 5; }
		int sub(const int8_t *row, int j) const { return row[q[j]]; }
	} s = { query, target, qlen };
	std::vector<int32_t> row(2 * (size_t)(qlen + 2));
	const int n_col = ((qlen < 2 * w + 1 ? qlen : 2 * w + 1) + 3) & ~3;
	std::vector<uint8_t> z((size_t)n_col * (tlen + 1) + 64);
	std::vector<uint32_t> cig((size_t)qlen + tlen + 8);
	GlobalRow eh = { row.data(), 1 };
	int n_cigar = 0;
	const int score = global_dp(go, s, tlen, w, eh, z.data(), cig.data(), &n_cigar, nullptr);
	if (cigar) cigar->assign(cig.begin(), cig.begin() + n_cigar);
	return score;
}

void stage_global_batch(Engine *e, const GlobalOpt &go, int64_t n, b200_global_job_t *jobs, const uint8_t *query, int64_t,
                        const uint8_t *target, int64_t tbytes, std::vector<uint32_t> &cigar)
{
	cigar.clear();
	std::vector<uint8_t> pac((size_t)tbytes / 4 + 2, 0);
	for (int64_t l = 0; l < tbytes; ++l) pac[l >> 2] |= (uint8_t)((target[l] & 3) << ((~l & 3) << 1));
	for (int64_t i = 0; i < n; ++i) {
		b200_global_job_t &j = jobs[i];
		GlobalJob g;
		g.rb = j.t_off; g.re = j.t_off + j.tlen; g.zoff = 0; g.read = 0; g.qb = 0; g.qe = j.qlen; g.w2 = j.w; g.truesc = B200_GLOBAL_RAW; g.wmax = j.w; g.cig_off = 0;
		std::vector<int32_t> row(2 * (size_t)(j.qlen + 2));
		const int n_col = ((j.qlen < 2 * j.w + 1 ? j.qlen : 2 * j.w + 1) + 3) & ~3;
		std::vector<uint8_t> z((size_t)n_col * (j.tlen + 1) + 64);
		std::vector<uint32_t> cg((size_t)j.qlen + j.tlen + 8);
		GlobalRow eh = { row.data(), 1 };
		GlobalRes r;
		global_task(go, global_seqs(pac.data(), tbytes, query + j.q_off, g), g, eh, z.data(), cg.data(), &r, &e->stats.global_cells);
		j.score = r.score; j.n_cigar = r.n_cigar; j.cigar_off = (int64_t)cigar.size();
		cigar.insert(cigar.end(), cg.begin(), cg.begin() + r.n_cigar);
	}
}

// finish_stage.h over plain loops
struct HostBK {
	Engine *e;
	template <class T> T *buf(int id, size_t n) { e->fb[id].assign((n ? n : 1) * sizeof(T) + 64, (char)0x5a); return (T *)e->fb[id].data(); }   // (poisoned: nothing may rely on zeroed scratch)
	template <class T> T *grow(int id, size_t n, size_t) { e->fb[id].resize(n * sizeof(T) + 64); return (T *)e->fb[id].data(); }
	template <class TASK> void run(int64_t n, const TASK &t) { for (int64_t i = 0; i < n; ++i) t(i); }
	void scan(const int32_t *in, int64_t *out, int64_t n) { int64_t s = 0; for (int64_t i = 0; i < n; ++i) { out[i] = s; s += in[i]; } }
	void sort_pairs(uint32_t *key, int32_t *val, int64_t n)
	{
		std::vector<std::pair<uint32_t, int32_t>> v(n);
		for (int64_t i = 0; i < n; ++i) v[i] = { key[i], val[i] };
		std::stable_sort(v.begin(), v.end(), [](const std::pair<uint32_t, int32_t> &a, const std::pair<uint32_t, int32_t> &b) { return a.first < b.first; });
		for (int64_t i = 0; i < n; ++i) { key[i] = v[i].first; val[i] = v[i].second; }
	}
	void zero(void *p, size_t bytes) { memset(p, 0, bytes); }
	int64_t get64(const int64_t *p) { return *p; }
	int32_t get32(const int32_t *p) { return *p; }
	void upload(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
	void download(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
	void sw_launch(const SwOpt &so, const SwJob *jobs, SwRes *res, const int32_t *order, const int32_t cnt[5], int, int)
	{
		const int64_t n = (int64_t)cnt[0] + cnt[1] + cnt[2] + cnt[3] + cnt[4];
		std::vector<uint16_t> H, E;
		std::vector<uint64_t> b;
		for (int64_t x = 0; x < n; ++x) {
			const SwJob &j = jobs[order[x]];
			H.resize(j.q_len + 16); E.resize(j.q_len + 16); b.resize(j.tlen / 2 + 2);
			Row16 h = { H.data(), 1 }, ee = { E.data(), 1 };
			List64 bl = { b.data(), 1 };
			STPac ta = { e->fm.pac, e->fm.l_pac, j.rb };
			const uint8_t *q = e->codes.data() + e->off[j.read] + j.q_beg;
			if (j.is_rev) { SQRevComp qa = { q, j.q_len }; sw_align(j.q_len, qa, j.tlen, ta, so, j.xtra, h, ee, bl, &res[order[x]], &e->stats.sw_cells); }
			else { SQFwd qa = { q }; sw_align(j.q_len, qa, j.tlen, ta, so, j.xtra, h, ee, bl, &res[order[x]], &e->stats.sw_cells); }
		}
	}
	void global_launch(const GlobalOpt &go, const GlobalJob *jobs, const int32_t *order, const int32_t cnt[6], const int32_t *, uint8_t *z,
	                   uint32_t *cig, GlobalRes *res)
	{
		static const int cls_S[5] = { 32, 64, 128, 256, 512 };
		std::vector<int32_t> row;
		int64_t x = 0;
		for (int k = 0; k < 6; ++k)
			for (int c = 0; c < cnt[k]; ++c, ++x) {
				const GlobalJob &j = jobs[order[x]];
				row.resize(2 * (size_t)(j.qe - j.qb + 2));
				GlobalRow eh = { row.data(), 1 };
				// (the shared-memory classes of the CUDA stage hold a window of S columns: a wider retry comes back flagged for the rerun pass)
				global_task(go, global_seqs(e->fm.pac, e->fm.l_pac, e->codes.data() + e->off[j.read] + j.qb, j), j, eh, z + j.zoff, cig + j.cig_off,
				            &res[order[x]], &e->stats.global_cells, k < 5 ? (cls_S[k] - 2) >> 1 : 0x7fffffff);
			}
	}
};

static double emu_clock_ms()
{
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void stage_finish(Engine *e, const FinishArgs &a)
{
	HostBK bk = { e };
	FinCtx cx;
	memset(&cx, 0, sizeof cx);
	cx.opt = *a.opt;
	cx.fm = e->fm;
	cx.ctg_alt = e->ctg_alt.data();
	cx.ctg_name_off = e->ctg_name_off.data(); cx.ctg_names = e->ctg_names.data();
	cx.ctg_anno_off = e->ctg_anno_off.data(); cx.ctg_annos = e->ctg_annos.data();
	cx.n_reads = e->n_reads; cx.pe = (a.opt->flag & MEM_F_PE) ? 1 : 0;
	cx.n_processed = a.n_processed;
	cx.off = e->off.data(); cx.codes = e->codes.data();
	cx.rtext = e->rtext.data(); cx.text = e->text.data();
	cx.rg_len = a.rg_id ? (int)strnlen(a.rg_id, 255) : 0;
	if (cx.rg_len) memcpy(cx.rg_id, a.rg_id, cx.rg_len);
	FinishIn in = { e->xregs.data(), e->xoff.data(), a.pes0, e->max_len, e->logtab->data(), (int)e->logtab->size(), a.route };
	FinishOut fo;
	finish_run(bk, cx, in, fo, e->stats, emu_clock_ms);
	e->fin_out = fo;
	e->stats.sam_bytes += fo.sam_bytes;
}

void stage_fetch_sam(Engine *e, const FinishArgs &a, SamChunk &out)
{
	const FinishOut fo = e->fin_out;
	const bool routed = fo.routed != nullptr;
	const char *text = routed ? fo.routed : fo.sam;
	const int64_t bytes = routed ? fo.routed_bytes : fo.sam_bytes;
	char *dst;
	if (a.alloc) dst = (char *)a.alloc((size_t)bytes + 1);
	else { e->sam.resize((size_t)bytes + 16); dst = e->sam.data(); }
	memcpy(dst, text, (size_t)bytes);
	dst[bytes] = 0;
	e->sam_off.assign(fo.sam_off, fo.sam_off + e->n_reads + 1);
	const int n_dest = e->fm.n_ctg + 2;
	const bool lines = a.route && !routed && fo.lines;
	if (lines) e->lines.assign(fo.lines, fo.lines + fo.n_lines);
	if (routed) e->dest_off.assign(fo.dest_off, fo.dest_off + n_dest + 1);
	out.lines = lines ? e->lines.data() : nullptr; out.n_lines = lines ? fo.n_lines : 0;
	out.dest_off = routed ? e->dest_off.data() : nullptr; out.n_dest = n_dest;
	out.sam = dst; out.sam_off = e->sam_off.data(); out.bytes = bytes;
}

void *stage_host_alloc(size_t bytes) { return malloc(bytes); }
void stage_host_free(void *p) { free(p); }

void stage_extend_bytes(Engine *e, const ExtOpt &eo, int64_t n_jobs, b200_extend_job_t *jobs,
                        const uint8_t *query, int64_t, const uint8_t *target, int64_t)
{
	std::vector<int32_t> eh;
	for (int64_t x = 0; x < n_jobs; ++x) {
		b200_extend_job_t &j = jobs[x];
		eh.resize(2 * (j.qlen + 2));
		EhStrided acc = { eh.data(), 1 };
		QFwd qa = { query + j.q_off };
		TBytes ta = { target + j.t_off };
		ExtOut o;
		extend_core(j.qlen, qa, j.tlen, ta, eo, j.w, j.end_bonus, j.h0, acc, &o, &e->stats.extend_cells);
		j.score = o.score; j.qle = o.qle; j.tle = o.tle; j.gtle = o.gtle; j.gscore = o.gscore; j.max_off = o.max_off;
	}
}

void stage_sw_bytes(Engine *e, const SwOpt &so, int64_t n_jobs, b200_align_job_t *jobs,
                    const uint8_t *query, int64_t, const uint8_t *target, int64_t)
{
	std::vector<uint16_t> H, E;
	std::vector<uint64_t> b;
	for (int64_t x = 0; x < n_jobs; ++x) {
		b200_align_job_t &j = jobs[x];
		H.resize(j.qlen + 16); E.resize(j.qlen + 16); b.resize(j.tlen / 2 + 2);
		Row16 h = { H.data(), 1 }, ee = { E.data(), 1 };
		List64 bl = { b.data(), 1 };
		SQFwd qa = { query + j.q_off };
		STBytes ta = { target + j.t_off };
		SwRes r;
		sw_align(j.qlen, qa, j.tlen, ta, so, j.xtra, h, ee, bl, &r, &e->stats.sw_cells);
		j.r.score = r.score; j.r.te = r.te; j.r.qe = r.qe; j.r.score2 = r.score2; j.r.te2 = r.te2; j.r.tb = r.tb; j.r.qb = r.qb;
	}
}

void stage_sa(Engine *e, int64_t n, const uint64_t *k, uint64_t *sa)
{
	for (int64_t i = 0; i < n; ++i) sa[i] = fm_sa(e->fm, k[i], nullptr);
}

int stage_smem1(Engine *e, int len, const uint8_t *q, int x, uint64_t min_intv, std::vector<Intv> &mem)
{
	std::vector<Intv> a(len + 1), b(len + 1);
	mem.assign(len + 1, Intv());
	int n = 0;
	const int ret = fm_smem1(e->fm, len, q, x, min_intv, mem.data(), &n, a.data(), b.data(), nullptr);
	mem.resize(n);
	return ret;
}

void stage_fm_extend(Engine *e, const Intv &ik, Intv ok[4], int is_back)
{
	fm_extend(e->fm, ik, ok, is_back, nullptr);
}

} // namespace b200

// ---- test-only entry points (tests/test_host_pipeline.py) ----------------------------------------------------------------
// occ sectors with counts beyond 2^32 (human-sized references: 2 x 3.1 Gbp of BWT rows need 33 bits): a synthetic stretch of
// the reference layout placed at row k0 is re-blocked with occ_convert_block and every Occ() is compared with a direct count.
extern "C" int64_t b200_emu_occ_selftest(uint64_t k0, uint32_t seed)
{
	using namespace b200;
	const int n_blk = 6;                                      // reference blocks of 128 symbols
	std::vector<uint32_t> ref((n_blk + 1) * 16, 0);
	std::vector<uint8_t> sym(n_blk * 128);
	uint64_t x = seed * 2654435761u + 12345;
	for (auto &s : sym) { x = x * 6364136223846793005ull + 1442695040888963407ull; s = (uint8_t)(x >> 61 & 3); }
	k0 &= ~(uint64_t)127;
	uint64_t c[4] = { k0 / 8, k0 / 2, k0 / 8, 0 };    // one count beyond 2^32 when k0 is
	c[3] = k0 - c[0] - c[1] - c[2];
	for (int j = 0; j < n_blk; ++j) {
		for (int t = 0; t < 4; ++t) { ref[j * 16 + 2 * t] = (uint32_t)c[t]; ref[j * 16 + 2 * t + 1] = (uint32_t)(c[t] >> 32); }
		for (int i = 0; i < 128; ++i) {
			const int s = sym[j * 128 + i];
			ref[j * 16 + 8 + (i >> 4)] |= (uint32_t)s << ((~i & 15) << 1);
			++c[s];
		}
	}
	int64_t bad = 0;
	uint64_t run[4] = { k0 / 8, k0 / 2, k0 / 8, 0 };
	run[3] = k0 - run[0] - run[1] - run[2];
	for (int i = 0; i < n_blk * 128; ++i) {
		++run[sym[i]];                                        // Occ(., k) counts rows 0..k inclusive
		OccRaw r;
		occ_convert_block(ref.data(), (uint64_t)(i >> 6), r.w);
		uint64_t cnt[4];
		occ4_sector(r, k0 + i, cnt);
		for (int t = 0; t < 4; ++t) bad += cnt[t] != run[t];
		bad += sector_symbol(r, k0 + i) != sym[i];
	}
	return bad;
}

// the packed interval entries of the seeding kernels (smem_kernel.cuh SeedList, smem_sweeps.cuh SweepStrip: three 33-bit values and
// a 29-bit end position in 16 bytes) round-trip values beyond 2^32, through the shared-memory slots and through the spill strip
extern "C" int64_t b200_emu_pack_selftest(uint32_t seed)
{
	using namespace b200;
	uint64_t x = seed * 0x9e3779b97f4a7c15ull + 1;
	auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
	int64_t bad = 0;
	const int quota = 3, n = 9;
	std::vector<uint32_t> sh(4 * quota * 2);
	std::vector<Q4> spill(n);
	SeedList L; L.sh = sh.data(); L.stride = 2; L.quota = quota; L.spill = spill.data(); L.sstride = 1;
	for (int rep = 0; rep < 2000; ++rep) {
		uint64_t v[n][3]; int end[n];
		for (int k = 0; k < n; ++k) {
			for (int t = 0; t < 3; ++t) { v[k][t] = rnd() & ((1ull << 33) - 1); if (rep % 3 == 0) v[k][t] |= 1ull << 32; }
			end[k] = (int)(rnd() & 0x1fffffff);
			L.set(k, v[k][0], v[k][1], v[k][2], end[k]);
		}
		for (int k = 0; k < n; ++k) {
			uint64_t a, b, c; int e;
			L.get(k, a, b, c, e);
			bad += a != v[k][0] || b != v[k][1] || c != v[k][2] || e != end[k];
			Q4 p = SweepStrip::pack(v[k][0], v[k][1], v[k][2], end[k]);
			SeedList S2 = L; S2.quota = 0; S2.spill = &p;
			S2.get(0, a, b, c, e);
			bad += a != v[k][0] || b != v[k][1] || c != v[k][2] || e != end[k];
		}
	}
	return bad;
}

// the derived index structures of round 2 (k-mer interval tables, whole suffix array and inverse, Bloom filters), built by the
// routines the upload kernels run, against the plain FM-index routines: returns the number of disagreements
extern "C" int64_t b200_emu_index_tables_selftest(const bwaidx_t *idx, uint32_t seed, int64_t *n_checked, double *bloom_fp_rate)
{
	using namespace b200;
	Engine *e = engine_create(idx->bwt, idx->bns, idx->pac, 0);
	const FmView &fm = e->fm;
	uint64_t x = seed * 0x9e3779b97f4a7c15ull + 1;
	auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
	int64_t bad = 0, n = 0;
	auto interval_of = [&](const int *pat, int L, Intv &ik) {      // forward extensions from the first base, like a forward sweep
		fm_set_intv(fm, pat[0], ik);
		for (int t = 1; t < L; ++t) { Intv ok[4]; fm_extend(fm, ik, ok, 0, nullptr); ik = ok[3 - pat[t]]; }
	};
	// 1. tables: random patterns of every length, half of them taken from the text so that long ones occur
	for (int L = 1; L <= fm.kmax; ++L)
		for (int t = 0; t < 400; ++t) {
			int pat[32] = { 0 };
			const int64_t p = (int64_t)(rnd() % (fm.seq_len - 40));
			for (int k = 0; k < L; ++k) pat[k] = (t & 1) ? fm_base(fm.pac, fm.l_pac, p + k) : (int)(rnd() & 3);
			uint32_t w = 0;
			for (int k = 0; k < L; ++k) w = w << 2 | (uint32_t)pat[k];
			Intv ik;
			interval_of(pat, L, ik);
			int half, tb;
			const uint32_t *sec = ktab_sector(fm, L, w, half);
			OccRaw r;
			for (int k = 0; k < 8; ++k) r.w[k] = sec[k];
			uint64_t t0, t1, t2;
			ktab_unpack(r, half, t0, t1, t2, tb);
			bad += t2 != ik.x2 || t0 != ik.x0 || t1 != ik.x1;       // (absent patterns included: the reference's coordinates of the empty interval)
			// the same interval reached backwards (what the backward chains use the table for)
			Intv bk;
			fm_set_intv(fm, pat[L - 1], bk);
			for (int k = L - 2; k >= 0 && bk.x2; --k) { Intv ok[4]; fm_extend(fm, bk, ok, 1, nullptr); bk = ok[pat[k]]; }
			if (ik.x2) bad += bk.x0 != t0 || bk.x1 != t1 || bk.x2 != t2;
			else bad += bk.x2 != 0;
			++n;
		}
	// 2. suffix array and inverse
	if (fm.sa5 && fm.isa5) {
		FmView walk = fm;
		walk.sa5 = nullptr;
		for (int t = 0; t < 3000; ++t) {
			const uint64_t k = 1 + rnd() % fm.seq_len;
			const uint64_t p = fm_sa(walk, k, nullptr);
			bad += sa5_read(fm.sa5, k) != p || sa5_read(fm.isa5, p) != k;
			++n;
		}
		bad += sa5_read(fm.isa5, 0) != fm.primary || sa5_read(fm.isa5, fm.seq_len) != 0;
	} else ++bad;
	// 3. Bloom filters: no false negatives (every window of the text is "present", every repeated one "more than once"); few false positives
	if (fm.bloom) {
		const int K = fm.bloom_k;
		for (int t = 0; t < 4000; ++t) {
			const int64_t p = (int64_t)(rnd() % (fm.seq_len - K));
			uint64_t word, bits;
			bloom_of_text(fm.pac, fm.l_pac, p, K, fm.bloom_mask, word, bits);
			bad += (fm.bloom[2 * word] & bits) != bits;
			int pat[32] = { 0 };
			for (int k = 0; k < K; ++k) pat[k] = fm_base(fm.pac, fm.l_pac, p + k);
			Intv ik;
			interval_of(pat, K, ik);
			bad += ik.x2 < 1;
			if (ik.x2 >= 2) bad += (fm.bloom[2 * word + 1] & bits) != bits;
			++n;
		}
		int64_t fp = 0, absent = 0;
		for (int t = 0; t < 20000; ++t) {
			int pat[32] = { 0 };
			uint64_t v = 0;
			for (int k = 0; k < K; ++k) { pat[k] = (int)(rnd() & 3); v = v << 2 | (uint64_t)pat[k]; }
			Intv ik;
			interval_of(pat, K, ik);
			if (ik.x2) continue;
			++absent;
			const uint64_t h = bloom_mix(v), bits = bloom_bits(h);
			fp += (fm.bloom[2 * (h & fm.bloom_mask)] & bits) == bits;
		}
		*bloom_fp_rate = absent ? (double)fp / absent : 0.;
	} else ++bad;
	*n_checked = n;
	engine_destroy(e);
	return bad;
}
