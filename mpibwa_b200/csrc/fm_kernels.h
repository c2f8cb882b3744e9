// fm_kernels.h - per-thread task bodies of the FM-index stage (seeding + suffix-array look-up).
//
// These are the bodies that the CUDA kernels in stages_cuda.cu run, one read (or one seed slot) per thread.
// They are written as host/device functions so that the test-only host build (tests/hostemu) can execute the
// very same arithmetic on the CPU to exercise the host orchestration without a GPU; the shipped library only
// contains the CUDA instantiation.
//
// Semantics follow the reference line by line where the output depends on it:
//   occ4 / extend            reference src/bwt.c:169-186, 262-275
//   smem1                    reference src/bwt.c:289-351   (max_intv is always 0 on the mem path; the single-call bwt_smem1 wrapper.
//                            The seeding stage itself runs the sweeps of smem_sweeps.cuh / the lane state machine of smem_kernel.cuh.)
//   sa_lookup                reference src/bwt.c:53-59, 86-96, 107-129
//   pos2rid / intv2rid       reference src/bntseq.c:349-375
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#define B200_HDN __host__ __device__
#else
#define B200_HD inline
#define B200_HDN inline
#endif

namespace b200 {

struct Intv { uint64_t x0, x1, x2, info; };     // same 32-byte layout as bwtintv_t

// FM-index in HBM: "occ sectors".  The reference interleaves, per 128 BWT symbols, four 64-bit counts with 32 bytes of
// 2-bit symbols (64-byte blocks, reference src/bwt.h:72-78): one Occ() is two 32-byte sectors, 4-8 load instructions and a
// popcount per 16-symbol word.  The device copy is re-blocked ONCE at upload (occ_convert_block) into 32-byte sectors of
// 64 symbols so that one Occ() for all four bases is ONE 256-bit load (LDG.256, one sector) and six popcounts:
//   w0,w1,w2  low 32 bits of #C, #G, #T before the block        w3  bits 32-39 of the three counts (8 bits each)
//   w4,w5     low-bit plane of the 64 symbols (symbol i at bit 31-(i&31) of word i>>5)      w6,w7  high-bit plane
// #A is implied: 64*block - #C - #G - #T.  Same footprint as the reference layout (0.5 byte per symbol).
struct FmView {
	const uint32_t *occ;        // occ sectors, 8 words per 64 symbols
	const uint64_t *sa;         // SA samples
	uint64_t primary, L2[5], seq_len;
	int sa_intv;
	const uint8_t *pac;         // 2-bit forward strand
	int64_t l_pac;
	const int64_t *ctg_off;     // contig offsets (forward strand)
	const int32_t *ctg_len;
	int n_ctg;
	// k-mer interval tables (smem_kernel.cuh, "k-mer tables"): the bi-interval of every pattern of 1 .. kmax bases, 16 bytes each,
	// length L at entry ktab_off(L) + (pattern as a base-4 number, first base most significant); kmax == 0: none
	const uint32_t *ktab;
	int kmax;
	// the whole suffix array, five bytes per row (fm_sa below), or null: only the reference's samples
	const uint8_t *sa5;
	// ... and its inverse (the row of the suffix that starts at a text position), same packing, or null
	const uint8_t *isa5;
	// which patterns of bloom_k bases the text holds at all (word 2i) and which it holds more than once (word 2i + 1): two blocked
	// Bloom filters (one 64-bit word per pattern, four bits) interleaved, or null
	const uint64_t *bloom; uint64_t bloom_mask; int bloom_k;
};

struct SeedOpt {                // the subset of mem_opt_t the seeding stage reads
	int min_seed_len, split_len, split_width, max_occ;
	int max_mem_intv;
};

B200_HD int popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
	return __popc(x);
#else
	return __builtin_popcount(x);
#endif
}

// counts of A,C,G,T among the 16 two-bit symbols of w, packed one per byte (A in the low byte)
B200_HD uint32_t sym_counts16(uint32_t w)
{
	uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
	uint32_t c3 = popc32(hi & lo), c2 = popc32(hi & ~lo & 0x55555555u), c1 = popc32(~hi & lo & 0x55555555u);
	uint32_t c0 = 16 - c1 - c2 - c3;
	return c0 | c1 << 8 | c2 << 16 | c3 << 24;
}

// the even bits of w gathered into the low 16 bits (bit 2m -> bit m)
B200_HD uint32_t even_bits16(uint32_t w)
{
	w &= 0x55555555u;
	w = (w | w >> 1) & 0x33333333u;
	w = (w | w >> 2) & 0x0f0f0f0fu;
	w = (w | w >> 4) & 0x00ff00ffu;
	w = (w | w >> 8) & 0x0000ffffu;
	return w;
}

// occ sector `b` (64 symbols) from the reference's occ-interleaved BWT (reference src/bwt.h:72-78, src/bwt.c:169-186).
// ref_bwt must be readable (zero-padded) up to the end of the 64-byte block that holds the sector.
B200_HD void occ_convert_block(const uint32_t *ref_bwt, uint64_t b, uint32_t out[8])
{
	const uint32_t *p = ref_bwt + ((b >> 1) << 4);
	uint64_t c[4];
	for (int i = 0; i < 4; ++i) c[i] = (uint64_t)p[2 * i + 1] << 32 | p[2 * i];
	const uint32_t *sy = p + 8;
	if (b & 1) {
		uint32_t x = 0;
		for (int i = 0; i < 4; ++i) x += sym_counts16(sy[i]);
		c[0] += x & 0xff; c[1] += x >> 8 & 0xff; c[2] += x >> 16 & 0xff; c[3] += x >> 24;
		sy += 4;
	}
	out[0] = (uint32_t)c[1]; out[1] = (uint32_t)c[2]; out[2] = (uint32_t)c[3];
	out[3] = (uint32_t)(c[1] >> 32 & 0xff) | (uint32_t)(c[2] >> 32 & 0xff) << 8 | (uint32_t)(c[3] >> 32 & 0xff) << 16;
	out[4] = even_bits16(sy[0]) << 16 | even_bits16(sy[1]);
	out[5] = even_bits16(sy[2]) << 16 | even_bits16(sy[3]);
	out[6] = even_bits16(sy[0] >> 1) << 16 | even_bits16(sy[1] >> 1);
	out[7] = even_bits16(sy[2] >> 1) << 16 | even_bits16(sy[3] >> 1);
}

struct OccRaw { uint32_t w[8]; };

// one 32-byte sector (p is 32-byte aligned) of a table read at random
B200_HD OccRaw ld_sector(const uint32_t *p)
{
	OccRaw r;
#if defined(__CUDA_ARCH__)
	asm("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
#else
	for (int i = 0; i < 8; ++i) r.w[i] = p[i];
#endif
	return r;
}

B200_HD OccRaw ld_occ(const FmView &fm, uint64_t blk)
{
	OccRaw r;
#if defined(__CUDA_ARCH__)
	const uint32_t *p = fm.occ + (blk << 3);
	// one 256-bit load; L1::no_allocate: the sectors are touched once at random, letting them through L1 evicts the read
	// bases and interval lists that ARE reused (measured: seeding 21.1 -> 16.1 ms per chunk)
	asm("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
#else
	const uint32_t *p = fm.occ + (blk << 3);
	for (int i = 0; i < 8; ++i) r.w[i] = p[i];
#endif
	return r;
}

// #C, #G, #T among the first kin+1 symbols of the sector (kin in 0..63)
B200_HD void occ_sector_counts(const OccRaw &r, uint32_t kin, uint32_t &n1, uint32_t &n2, uint32_t &n3)
{
	const uint32_t ma = kin >= 31u ? 0xffffffffu : 0xffffffffu << (31u - kin);
	const uint32_t mb = kin < 32u ? 0u : 0xffffffffu << (63u - kin);
	const uint32_t la = r.w[4] & ma, lb = r.w[5] & mb, ha = r.w[6] & ma, hb = r.w[7] & mb;
	n3 = (uint32_t)popc32(la & ha) + (uint32_t)popc32(lb & hb);
	n1 = (uint32_t)popc32(la & ~ha) + (uint32_t)popc32(lb & ~hb);
	n2 = (uint32_t)popc32(~la & ha) + (uint32_t)popc32(~lb & hb);
}

// Occ(c, ka) for the four symbols from a loaded sector; ka = row index after the sentinel adjustment
B200_HD void occ4_sector(const OccRaw &r, uint64_t ka, uint64_t cnt[4])
{
	const uint32_t kin = (uint32_t)ka & 63u;
	uint32_t n1, n2, n3;
	occ_sector_counts(r, kin, n1, n2, n3);
	cnt[1] = ((uint64_t)(r.w[3] & 0xffu) << 32 | r.w[0]) + n1;
	cnt[2] = ((uint64_t)(r.w[3] >> 8 & 0xffu) << 32 | r.w[1]) + n2;
	cnt[3] = ((uint64_t)(r.w[3] >> 16 & 0xffu) << 32 | r.w[2]) + n3;
	cnt[0] = ka + 1 - cnt[1] - cnt[2] - cnt[3];
}

// Occ(c, k) for the four symbols; k == (uint64_t)-1 gives zeros.  *blk receives the index of the reference's 64-byte block
// (the unit of the algorithmic traffic count), or -1.
B200_HD void fm_occ4(const FmView &fm, uint64_t k, uint64_t cnt[4], int64_t *blk)
{
	if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; if (blk) *blk = -1; return; }
	k -= (k >= fm.primary);
	if (blk) *blk = (int64_t)(k >> 7);
	occ4_sector(ld_occ(fm, k >> 6), k, cnt);
}

// bi-directional extension of ik by each of the four bases
B200_HD void fm_extend(const FmView &fm, const Intv &ik, Intv ok[4], int is_back, int64_t *n_blocks)
{
	uint64_t tk[4], tl[4];
	uint64_t base = is_back ? ik.x0 : ik.x1;
	int64_t b0, b1;
	fm_occ4(fm, base - 1, tk, &b0);
	fm_occ4(fm, base - 1 + ik.x2, tl, &b1);
	if (n_blocks) *n_blocks += (b0 >= 0) + (b1 >= 0 && b1 != b0);
	uint64_t nb[4];
	for (int i = 0; i < 4; ++i) {
		nb[i] = fm.L2[i] + 1 + tk[i];
		ok[i].x2 = tl[i] - tk[i];
	}
	uint64_t other = is_back ? ik.x1 : ik.x0;
	uint64_t o3 = other + (base <= fm.primary && base + ik.x2 - 1 >= fm.primary);
	uint64_t o2 = o3 + ok[3].x2, o1 = o2 + ok[2].x2, o0 = o1 + ok[1].x2;
	if (is_back) {
		ok[0].x0 = nb[0]; ok[1].x0 = nb[1]; ok[2].x0 = nb[2]; ok[3].x0 = nb[3];
		ok[0].x1 = o0; ok[1].x1 = o1; ok[2].x1 = o2; ok[3].x1 = o3;
	} else {
		ok[0].x1 = nb[0]; ok[1].x1 = nb[1]; ok[2].x1 = nb[2]; ok[3].x1 = nb[3];
		ok[0].x0 = o0; ok[1].x0 = o1; ok[2].x0 = o2; ok[3].x0 = o3;
	}
}

B200_HD void fm_set_intv(const FmView &fm, int c, Intv &ik)
{
	ik.x0 = fm.L2[c] + 1;
	ik.x2 = fm.L2[c + 1] - fm.L2[c];
	ik.x1 = fm.L2[3 - c] + 1;
	ik.info = 0;
}

// All SMEMs through position x (max_intv == 0 flavour).  `a` and `b` are scratch of len+1 entries each,
// `mem` receives the result (n_mem entries, sorted by start).  Returns the next x.
B200_HDN int fm_smem1(const FmView &fm, int len, const uint8_t *q, int x, uint64_t min_intv,
                      Intv *mem, int *n_mem, Intv *a, Intv *b, int64_t *n_blocks)
{
	*n_mem = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	Intv ik, ok[4];
	Intv *prev = a, *curr = b;
	int n_prev, n_curr = 0, i;
	fm_set_intv(fm, q[x], ik);
	ik.info = x + 1;
	for (i = x + 1; i < len; ++i) {               // forward sweep
		if (q[i] < 4) {
			int c = 3 - q[i];
			fm_extend(fm, ik, ok, 0, n_blocks);
			if (ok[c].x2 != ik.x2) {
				curr[n_curr++] = ik;
				if (ok[c].x2 < min_intv) break;
			}
			ik = ok[c]; ik.info = i + 1;
		} else {
			curr[n_curr++] = ik;
			break;
		}
	}
	if (i == len) curr[n_curr++] = ik;
	for (int j = 0; j < n_curr >> 1; ++j) { Intv t = curr[j]; curr[j] = curr[n_curr - 1 - j]; curr[n_curr - 1 - j] = t; }
	int ret = (int)curr[0].info;
	{ Intv *t = curr; curr = prev; prev = t; }
	n_prev = n_curr;
	int nm = 0;
	for (i = x - 1; i >= -1; --i) {               // backward sweep
		int c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
		n_curr = 0;
		for (int j = 0; j < n_prev; ++j) {
			const Intv p = prev[j];
			if (c >= 0) fm_extend(fm, p, ok, 1, n_blocks);
			if (c < 0 || ok[c].x2 < min_intv) {
				if (n_curr == 0) {
					if (nm == 0 || (uint64_t)(i + 1) < (mem[nm - 1].info >> 32)) {
						ik = p; ik.info |= (uint64_t)(i + 1) << 32;
						mem[nm++] = ik;
					}
				}
			} else if (n_curr == 0 || ok[c].x2 != curr[n_curr - 1].x2) {
				ok[c].info = p.info;
				curr[n_curr++] = ok[c];
			}
		}
		if (n_curr == 0) break;
		{ Intv *t = curr; curr = prev; prev = t; }
		n_prev = n_curr;
	}
	for (int j = 0; j < nm >> 1; ++j) { Intv t = mem[j]; mem[j] = mem[nm - 1 - j]; mem[nm - 1 - j] = t; }
	*n_mem = nm;
	return ret;
}

// number of suffix-array look-ups mem_chain() makes for one interval (reference src/bwamem.c:278-279)
B200_HD int seed_slots(uint64_t x2, int max_occ)
{
	if (x2 <= (uint64_t)max_occ) return (int)x2;
	uint64_t step = x2 / (uint64_t)max_occ;
	uint64_t cnt = (x2 + step - 1) / step;
	return cnt < (uint64_t)max_occ ? (int)cnt : max_occ;
}
B200_HD uint64_t seed_step(uint64_t x2, int max_occ) { return x2 > (uint64_t)max_occ ? x2 / (uint64_t)max_occ : 1; }

B200_HD int sector_symbol(const OccRaw &r, uint64_t ka)
{
	const uint32_t kin = (uint32_t)ka & 63u, sh = 31u - (kin & 31u);
	const uint32_t lo = kin < 32u ? r.w[4] : r.w[5], hi = kin < 32u ? r.w[6] : r.w[7];
	return (int)((lo >> sh & 1u) | (hi >> sh & 1u) << 1);
}

B200_HD int fm_B0(const FmView &fm, uint64_t k) { return sector_symbol(ld_occ(fm, k >> 6), k); }

// Occ(c,k) for one symbol
B200_HD uint64_t fm_occ1(const FmView &fm, uint64_t k, int c)
{
	if (k == fm.seq_len) return fm.L2[c + 1] - fm.L2[c];
	if (k == (uint64_t)-1) return 0;
	k -= (k >= fm.primary);
	uint64_t cnt[4];
	occ4_sector(ld_occ(fm, k >> 6), k, cnt);
	return cnt[c];
}

// bwt_invPsi (reference src/bwt.c:53-59): the row of the suffix that starts one base earlier
B200_HD uint64_t fm_lf(const FmView &fm, uint64_t k)
{
	if (k == fm.primary) return 0;
	// for k != primary the symbol row k - (k > primary) and the row of bwt_occ's count k - (k >= primary) coincide, so one sector gives both
	const uint64_t kk = k - (k > fm.primary);
	const OccRaw r = ld_occ(fm, kk >> 6);
	const int c = sector_symbol(r, kk);
	uint64_t cnt[4];
	occ4_sector(r, kk, cnt);
	return (c == 0 ? fm.L2[0] : c == 1 ? fm.L2[1] : c == 2 ? fm.L2[2] : fm.L2[3]) + (c == 0 ? cnt[0] : c == 1 ? cnt[1] : c == 2 ? cnt[2] : cnt[3]);
}

/* The whole suffix array in HBM.  bwt_sa (reference src/bwt.c:86-96) walks bwt_invPsi from row k to a sampled row - 31 dependent
 * random sectors on average with the shipped sampling of 32 - because a host cannot afford 8 bytes per row.  A B200 can afford five
 * (31 GB for a human-sized reference): the samples are expanded ONCE when the index is uploaded (sa5_expand: from every sampled
 * row the walk of invPsi visits exactly the unsampled rows up to the next sampled one, each one base earlier in the text), and a
 * look-up is one random access.  Row 0 keeps the reference's own sa[0]. */
B200_HD uint64_t sa5_read(const uint8_t *sa5, uint64_t k)
{
	const uint64_t o = k * 5;
	const uint32_t *w = reinterpret_cast<const uint32_t *>(sa5) + (o >> 2);
	const uint64_t v = (uint64_t)w[1] << 32 | w[0];
	return v >> ((o & 3) << 3) & 0xffffffffffull;
}
B200_HD void sa5_write(uint8_t *sa5, uint64_t k, uint64_t v)
{
	uint8_t *p = sa5 + k * 5;
	p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); p[4] = (uint8_t)(v >> 32);
}
// the rows between sampled row j * sa_intv and the next sampled row on the walk (isa5: the inverse array as well, or null)
B200_HD void sa5_expand(const FmView &fm, uint64_t j, uint8_t *sa5, uint8_t *isa5)
{
	const uint64_t mask = (uint64_t)fm.sa_intv - 1;
	uint64_t row = j * (uint64_t)fm.sa_intv;
	uint64_t p = j == 0 ? fm.seq_len : fm.sa[j];          // (the reference stores -1 for the row of the empty suffix)
	sa5_write(sa5, row, p);
	if (isa5 && p <= fm.seq_len) sa5_write(isa5, p, row);
	for (int guard = 0; guard < (1 << 24); ++guard) {      // (the bound only matters for a damaged index whose walk never reaches a sample)
		if (p == 0) break;                                 // the suffix at the start of the text: the walk is over
		row = fm_lf(fm, row); --p;
		if (!(row & mask)) break;
		sa5_write(sa5, row, p);
		if (isa5 && p <= fm.seq_len) sa5_write(isa5, p, row);
	}
}

/* A seed shorter than min_seed_len is never kept (reference src/bwamem.c:130), and most entries of a backward sweep end one or two
 * bases short of it.  Whether the min_seed_len-base window that ends where an entry ends occurs in the text AT ALL is one random
 * 8-byte read of a Bloom filter over the text's bloom_k-mers (11 bits per text position, four hash bits inside one word; built
 * when the index is uploaded): "absent" is certain, so the entry can be dropped without walking it; "maybe" walks it as before.
 * The re-seeding pass (src/bwamem.c:135-147) only keeps intervals of MORE occurrences than the SMEM it splits (min_intv >= 2), so
 * it asks a second filter that holds the windows inserted more than once (bloom_insert: the word's previous value says whether all
 * of the pattern's bits were there already - by an earlier occurrence or by chance, never the other way round). */
B200_HD uint64_t bloom_mix(uint64_t v)
{
	v += 0x9e3779b97f4a7c15ull; v = (v ^ (v >> 30)) * 0xbf58476d1ce4e5b9ull; v = (v ^ (v >> 27)) * 0x94d049bb133111ebull;
	return v ^ (v >> 31);
}
B200_HD uint64_t bloom_bits(uint64_t h) { return (uint64_t)1 << (h >> 40 & 63) | (uint64_t)1 << (h >> 46 & 63) | (uint64_t)1 << (h >> 52 & 63) | (uint64_t)1 << (h >> 58); }
// the filter word and bits of the K bases that start at text position p (forward strand followed by its reverse complement)
B200_HD void bloom_of_text(const uint8_t *pac, int64_t l_pac, int64_t p, int K, uint64_t mask, uint64_t &word, uint64_t &bits);
// host-side insertion (tests/hostemu); the device does the same with atomicOr's return value
inline void bloom_insert(uint64_t *bloom, uint64_t word, uint64_t bits)
{
	const uint64_t old = bloom[2 * word];
	bloom[2 * word] = old | bits;
	if ((old & bits) == bits) bloom[2 * word + 1] |= bits;
}
inline uint64_t bloom_words_for(uint64_t seq_len)
{
	uint64_t n = 1024;
	while (n * 64 < seq_len * 11) n <<= 1;
	return n;
}

// SA[k]: from the expanded array, else by walking to a sampled row
B200_HD uint64_t fm_sa(const FmView &fm, uint64_t k, int *steps)
{
	if (fm.sa5 && k) { if (steps) *steps = 0; return sa5_read(fm.sa5, k); }
	uint64_t sa = 0, mask = (uint64_t)fm.sa_intv - 1;
	int n = 0;
	while (k & mask) { ++sa; ++n; k = fm_lf(fm, k); }
	if (steps) *steps = n;
	return sa + fm.sa[k / (uint64_t)fm.sa_intv];
}

B200_HD int fm_pos2rid(const FmView &fm, int64_t pos_f)
{
	if (pos_f >= fm.l_pac) return -1;
	int left = 0, mid = 0, right = fm.n_ctg;
	while (left < right) {
		mid = (left + right) >> 1;
		if (pos_f >= fm.ctg_off[mid]) {
			if (mid == fm.n_ctg - 1) break;
			if (pos_f < fm.ctg_off[mid + 1]) break;
			left = mid + 1;
		} else right = mid;
	}
	return mid;
}

B200_HD void bloom_of_text(const uint8_t *pac, int64_t l_pac, int64_t p, int K, uint64_t mask, uint64_t &word, uint64_t &bits)
{
	uint64_t v = 0;
	for (int t = 0; t < K; ++t) {
		const int64_t a = p + t;
		int c;
		if (a < l_pac) c = pac[a >> 2] >> ((~a & 3) << 1) & 3;
		else { const int64_t f = (l_pac << 1) - 1 - a; c = 3 - (pac[f >> 2] >> ((~f & 3) << 1) & 3); }
		v = v << 2 | (uint64_t)c;
	}
	const uint64_t h = bloom_mix(v);
	word = h & mask; bits = bloom_bits(h);
}

B200_HD int64_t fm_depos(const FmView &fm, int64_t pos, int *is_rev)
{
	return (*is_rev = (pos >= fm.l_pac)) ? (fm.l_pac << 1) - 1 - pos : pos;
}

B200_HD int fm_intv2rid(const FmView &fm, int64_t rb, int64_t re)
{
	int is_rev;
	if (rb < fm.l_pac && re > fm.l_pac) return -2;
	int rid_b = fm_pos2rid(fm, fm_depos(fm, rb, &is_rev));
	int rid_e = rb < re ? fm_pos2rid(fm, fm_depos(fm, re - 1, &is_rev)) : rid_b;
	return rid_b == rid_e ? rid_b : -1;
}

// base code (0-3) at position p of the forward+reverse-complement coordinate system [0, 2*l_pac)
B200_HD int fm_base(const uint8_t *pac, int64_t l_pac, int64_t p)
{
	if (p < l_pac) return pac[p >> 2] >> ((~p & 3) << 1) & 3;
	int64_t f = (l_pac << 1) - 1 - p;
	return 3 - (pac[f >> 2] >> ((~f & 3) << 1) & 3);
}

} // namespace b200
