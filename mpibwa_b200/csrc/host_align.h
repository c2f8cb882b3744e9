// host_align.h - what is left of the BWA-MEM read pipeline on the host: chaining + chain filtering for batches with reads long
// enough for mem_flt_chained_seeds (>= ~730 bases; also the B200_CHAIN=check cross-check of the device chaining stage) and the
// double-precision arithmetic of mem_pestat.  Everything else runs on the device (finish_kernels.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/mpibwa_b200.h"
#include "stages.h"

namespace b200 {

extern const unsigned char kNt4[256];

struct HSeed { int64_t rbeg; int32_t qbeg, len, score; };

struct HChain {
	int64_t pos = 0;
	int rid = 0, first = -1, w = 0, kept = 0, is_alt = 0;
	float frac_rep = 0;
	std::vector<HSeed> seeds;
};

// reference src/bntseq.c:349-375, src/bntseq.h:87
int     bns_pos2rid_h(const bntseq_t *bns, int64_t pos_f);
int     bns_intv2rid_h(const bntseq_t *bns, int64_t rb, int64_t re);
static inline int64_t bns_depos_h(const bntseq_t *bns, int64_t pos, int *is_rev)
{ return (*is_rev = (pos >= bns->l_pac)) ? (bns->l_pac << 1) - 1 - pos : pos; }
// clip [beg,end) to the contig holding mid (the window part of bns_fetch_seq, reference src/bntseq.c:421-446)
void    bns_clip_window(const bntseq_t *bns, int64_t *beg, int64_t mid, int64_t *end, int *rid);
// mem_chain() minus the FM-index work: seeds arrive in look-up order.  reference src/bwamem.c:251-315
void build_chains(const mem_opt_t *opt, const bntseq_t *bns, int l_seq, const SeedRec *seeds, int64_t n_seeds,
                  int l_rep, std::vector<HChain> &chains);
// mem_chain_flt, reference src/bwamem.c:327-385
void filter_chains(const mem_opt_t *opt, std::vector<HChain> &chains);
// cal_max_gap + rmax window of mem_chain2aln, reference src/bwamem.c:621-660
void chain_window(const mem_opt_t *opt, const bntseq_t *bns, int l_query, const HChain &c, int64_t rmax[2]);

// mem_pestat's statistics from the device-gathered candidates, reference src/bwamem_pair.c:67-109 (declared in finish_stage.h)

} // namespace b200
