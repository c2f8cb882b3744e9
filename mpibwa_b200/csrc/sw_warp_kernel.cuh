// sw_warp_kernel.cuh - batched ksw_align2 (mate rescue, reference src/bwamem_pair.c:150 -> src/ksw.c:343-365) with
// ONE WARP PER JOB.
//
// Mate rescue produces few (10^3..10^5 per chunk) but large jobs (150-250 query columns x 600-1100 target rows, two
// passes), so one job per thread (v1) leaves the chip empty and every thread on a ~10^5-cell dependent chain.  Here a
// warp runs one job as a systolic wavefront: lane L owns the C query columns [L*C, L*C+C) with their H/E state in
// registers and works on target row (t - L) at step t; the right edge of its strip (H, F), the running row maximum
// (value, column) and the target base of the row travel to lane L+1 in two 32-bit shuffles per step.  Lane 31 sees
// every completed row in order and applies the reference's sequential rules there: the run list that yields
// score2/te2, the first-row-reaching-the-best rule for te/qe, 8-bit saturation and the KSW_XSTOP early exit.
// Substitution scores come from one PRMT over a two-register row of the matrix selected by the target base.
//
// The observable semantics are those of sw_pass()/sw_align() in sw_kernels.h (the scalar restatement that the parity
// tests pin against the compiled reference): padded columns score 0 and take part in the row maximum, columns past
// the padding are masked out, H/E/F clamp at 0, 8-bit scores saturate at 255 - shift.
#pragma once
#include "sw_kernels.h"

namespace b200 {

__device__ __forceinline__ int sw_prmt(uint32_t lo, uint32_t hi, uint32_t sel)
{
	int s;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(s) : "r"(lo), "r"(hi), "r"(sel));
	return s;
}

// shared look-up table: for target code t (0..4) two words = bytes { mat[t][0..4], 0 (padding column), 0, 0 }
__device__ __forceinline__ void sw_fill_lut(const SwOpt &o, uint32_t *lut)
{
	if (threadIdx.x < 5) {
		const int8_t *m = o.mat + threadIdx.x * 5;
		lut[threadIdx.x * 2] = (uint32_t)(uint8_t)m[0] | (uint32_t)(uint8_t)m[1] << 8 | (uint32_t)(uint8_t)m[2] << 16 | (uint32_t)(uint8_t)m[3] << 24;
		lut[threadIdx.x * 2 + 1] = (uint32_t)(uint8_t)m[4];
	}
}

// one pass of ksw_u8 / ksw_i16 semantics by a full warp; results valid in every lane
template <int C, class QA, class TA>
__device__ void sw_pass_warp(int qlen, QA query, int tlen, TA target, const SwOpt &o, const uint32_t *lut, int size, int minsc,
                             int endsc, uint64_t *b, SwRes *r, long long *cells)
{
	const int lane = threadIdx.x & 31;
	const int p = size == 1 ? 16 : 8;
	const int qpad = (qlen + p - 1) / p * p;
	const int cap = size == 1 ? 255 - o.shift : 0x7fffffff;
	const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins, e_del = o.e_del, e_ins = o.e_ins;
	uint32_t qsel[C], vm[C];
	int H[C], E[C];
#pragma unroll
	for (int c = 0; c < C; ++c) {
		const int j = lane * C + c;
		const int code = j < qlen ? query(j) : 5;
		qsel[c] = (uint32_t)code * 0x1111u + 0x8880u;
		vm[c] = j < qpad ? 0xffffffffu : 0u;
		H[c] = 0; E[c] = 0;
	}
	uint32_t out0 = 0, out1 = 0;
	int diag_next = 0;
	int gmax = 0, te = -1, qe = 0, n_b = 0, stop = 0;       // meaningful in lane 31
	uint64_t cur_b = 0;
	int tb_next = (lane == 0 && tlen > 0) ? target(0) : 0;
	for (int t = 0; t < tlen + 31; ++t) {
		uint32_t in0 = __shfl_up_sync(0xffffffffu, out0, 1), in1 = __shfl_up_sync(0xffffffffu, out1, 1);
		if (lane == 0) {
			in0 = 0; in1 = (uint32_t)tb_next << 24;
			if (t + 1 < tlen) tb_next = target(t + 1);
		}
		const int row = t - lane;
		if (row >= 0 && row < tlen) {
			int diag = diag_next;
			diag_next = (int)(in0 & 0xffffu);
			int f = (int)(in0 >> 16), imax = (int)(in1 & 0xffffu), iq = (int)(in1 >> 16 & 0xffu);
			const uint32_t tb = in1 >> 24;
			const uint32_t lo = lut[tb * 2], hi = lut[tb * 2 + 1];
#pragma unroll
			for (int c = 0; c < C; ++c) {
				int h = diag + sw_prmt(lo, hi, qsel[c]);
				h = min(h, cap);
				h = max(max(h, 0), max(E[c], f));
				diag = H[c];
				H[c] = h;
				const int hm = (int)((uint32_t)h & vm[c]);
				if (hm > imax) { imax = hm; iq = lane * C + c; }
				E[c] = max(E[c] - e_del, max(h - oe_del, 0));
				f = max(f - e_ins, max(h - oe_ins, 0));
			}
			out0 = (uint32_t)H[C - 1] | (uint32_t)f << 16;
			out1 = (uint32_t)imax | (uint32_t)iq << 16 | tb << 24;
			if (lane == 31) {                                 // row `row` is complete: the reference's per-row epilogue
				if (imax >= minsc) {
					if (n_b == 0 || (int32_t)cur_b + 1 != row) {
						if (n_b > 0) b[n_b - 1] = cur_b;
						cur_b = (uint64_t)imax << 32 | (uint32_t)row;
						++n_b;
					} else if ((int)(cur_b >> 32) < imax) cur_b = (uint64_t)imax << 32 | (uint32_t)row;
				}
				if (imax > gmax) {
					gmax = imax; te = row; qe = iq;
					if (size == 1) { if (gmax + o.shift >= 255 || gmax >= endsc) stop = 1; }
					else if (gmax >= endsc) stop = 1;
				}
			}
		}
		stop = __shfl_sync(0xffffffffu, stop, 31);
		if (stop) break;
	}
	if (lane == 31 && n_b > 0) b[n_b - 1] = cur_b;
	gmax = __shfl_sync(0xffffffffu, gmax, 31);
	te = __shfl_sync(0xffffffffu, te, 31);
	qe = __shfl_sync(0xffffffffu, qe, 31);
	n_b = __shfl_sync(0xffffffffu, n_b, 31);
	__syncwarp();
	if (lane == 0 && cells) *cells += (long long)(stop ? te + 1 : tlen) * qpad;
	r->score = size == 1 ? (gmax + o.shift < 255 ? gmax : 255) : gmax;
	r->te = te; r->qe = -1; r->score2 = -1; r->te2 = -1; r->tb = -1; r->qb = -1;
	if (size == 2 || r->score != 255) {
		r->qe = qe;
		if (n_b > 0) {
			const int d = (r->score + o.max_sc - 1) / o.max_sc;
			const int low = te - d, high = te + d;
			// first entry (in list order) with the largest score outside [low, high]
			unsigned long long best = 0;
			for (int k = lane; k < n_b; k += 32) {
				const uint64_t v = b[k];
				const int e = (int32_t)v;
				if (e < low || e > high) {
					const unsigned long long key = (unsigned long long)(uint32_t)(v >> 32) << 32 | (0xffffffffu - (uint32_t)k);
					best = key > best ? key : best;
				}
			}
			for (int off = 16; off > 0; off >>= 1) {
				const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, off);
				best = other > best ? other : best;
			}
			if (best != 0) {
				const int k = (int)(0xffffffffu - (uint32_t)best);
				r->score2 = (int)(best >> 32);
				r->te2 = (int32_t)b[k];
			}
		}
	}
	__syncwarp();
}

// ksw_align2 with qry == NULL (both passes)
template <int C, class QA, class TA>
__device__ void sw_align_warp(int qlen, QA query, int tlen, TA target, const SwOpt &o, const uint32_t *lut, int xtra, uint64_t *b,
                              SwRes *r, long long *cells)
{
	const int size = (xtra & 0x10000) ? 1 : 2;
	const int minsc = (xtra & 0x40000) ? (xtra & 0xffff) : 0x10000;
	const int endsc = (xtra & 0x20000) ? (xtra & 0xffff) : 0x10000;
	sw_pass_warp<C>(qlen, query, tlen, target, o, lut, size, minsc, endsc, b, r, cells);
	if ((xtra & 0x80000) == 0 || ((xtra & 0x40000) && r->score < (xtra & 0xffff))) return;
	if (r->qe < 0) return;
	SwRes rr;
	SQFlip<QA> q2 = { query, r->qe };
	STFlip<TA> t2 = { target, r->te };
	sw_pass_warp<C>(r->qe + 1, q2, tlen, t2, o, lut, size, 0x10000, r->score & 0xffff, b, &rr, cells);
	if (r->score == rr.score) { r->tb = r->te - rr.te; r->qb = r->qe - rr.qe; }
}

} // namespace b200
