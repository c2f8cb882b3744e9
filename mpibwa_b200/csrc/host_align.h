// host_align.h - the host-resident part of the BWA-MEM read pipeline: chaining, chain filtering, region
// de-duplication, insert-size statistics, pairing, mapQ, CIGAR/MD and SAM text.  These steps are pointer-chasing
// and text work that the north star leaves on the host; their results must be bit-identical to the reference, so
// every function names the reference routine whose behaviour it reproduces.
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

struct Aln {                    // counterpart of mem_aln_t
	int64_t pos = -1;
	int rid = -1, flag = 0;
	uint32_t is_rev = 0, is_alt = 0, mapq = 0, NM = 0;
	std::vector<uint32_t> cigar;
	std::string md;
	const std::string *XA = nullptr;
	int score = 0, sub = 0, alt_sc = 0;
};

typedef std::vector<mem_alnreg_t> RegVec;

// Plumbing between mem_reg2aln's alignment step and the batched CIGAR stage on the device.  Before the SAM sweep the
// pipeline queues a banded global alignment for every region that mem_reg2aln may be asked about (plan_global_jobs: a
// function of the region and the read alone, so no dry run of the pairing logic is needed); the device aligns them all at
// once; the sweep then runs in LOOKUP mode, where reg2aln finds the alignment of its region among the pair's jobs by
// (read, interval, band, score).  A region that was not queued, or whose CIGAR did not fit the result record, is aligned
// on the spot by the host routine - same arithmetic, same result.
struct AlignCtx {
	enum { DIRECT = 0, LOOKUP = 2 };
	int mode = DIRECT;
	const GlobalJob *jobs = nullptr;     // jobs / results of the current pair
	const GlobalRes *res = nullptr;
	int n_jobs = 0;
	const char *seq_ptr[2] = { nullptr, nullptr };
	int read_idx[2] = { 0, 0 };
	int64_t n_host_dp = 0;               // regions the sweep had to align itself (not queued, or CIGAR too long for the record)
	std::string *sink = nullptr;         // when set, SAM records are appended here (input order within a block of pairs) instead of malloc()ed per read
};
AlignCtx &align_ctx();                   // thread-local

// reference src/bntseq.c:349-375, src/bntseq.h:87
int     bns_pos2rid_h(const bntseq_t *bns, int64_t pos_f);
int     bns_intv2rid_h(const bntseq_t *bns, int64_t rb, int64_t re);
static inline int64_t bns_depos_h(const bntseq_t *bns, int64_t pos, int *is_rev)
{ return (*is_rev = (pos >= bns->l_pac)) ? (bns->l_pac << 1) - 1 - pos : pos; }
// clip [beg,end) to the contig holding mid (the window part of bns_fetch_seq, reference src/bntseq.c:421-446)
void    bns_clip_window(const bntseq_t *bns, int64_t *beg, int64_t mid, int64_t *end, int *rid);
// bns_get_seq, reference src/bntseq.c:398-419
void    bns_get_seq_h(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, std::vector<uint8_t> &seq);

// mem_chain() minus the FM-index work: seeds arrive in look-up order.  reference src/bwamem.c:251-315
void build_chains(const mem_opt_t *opt, const bntseq_t *bns, int l_seq, const SeedRec *seeds, int64_t n_seeds,
                  int l_rep, std::vector<HChain> &chains);
// mem_chain_flt, reference src/bwamem.c:327-385
void filter_chains(const mem_opt_t *opt, std::vector<HChain> &chains);
// cal_max_gap + rmax window of mem_chain2aln, reference src/bwamem.c:621-660
void chain_window(const mem_opt_t *opt, const bntseq_t *bns, int l_query, const HChain &c, int64_t rmax[2]);

// ksw_global2, reference src/ksw.c:504-606
int  global_align(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                  int o_del, int e_del, int o_ins, int e_ins, int w, std::vector<uint32_t> *cigar);
// bwa_gen_cigar2, reference src/bwa.c:121-207.  Returns false when the reference would return NULL.
bool gen_cigar(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac,
               const uint8_t *pac, int l_query, uint8_t *query, int64_t rb, int64_t re, int *score,
               std::vector<uint32_t> *cigar, int *NM, std::string *md, const GlobalRes *pre = nullptr);

// mem_sort_dedup_patch / mem_patch_reg, reference src/bwamem.c:406-489
int  sort_dedup_patch(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, uint8_t *query, int n, mem_alnreg_t *a);
// mem_mark_primary_se, reference src/bwamem.c:493-558
int  mark_primary_se(const mem_opt_t *opt, int n, mem_alnreg_t *a, int64_t id);
// mem_reorder_primary5, reference src/bwamem.c:978-1000
void reorder_primary5(int T, RegVec &a);
// mem_approx_mapq_se, reference src/bwamem.c:952-976
int  approx_mapq_se(const mem_opt_t *opt, const mem_alnreg_t *a);
// mem_pestat, reference src/bwamem_pair.c:46-109
void pestat(const mem_opt_t *opt, int64_t l_pac, int n, const RegVec *regs, mem_pestat_t pes[4]);
// mem_infer_dir, reference src/bwamem_pair.c:23-30
int  infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist);
// mem_pair, reference src/bwamem_pair.c:182-243
int  pair_ends(const mem_opt_t *opt, const bntseq_t *bns, const mem_pestat_t pes[4], RegVec a[2], int id,
               int *sub, int *n_sub, int z[2], int n_pri[2]);
// the CIGAR-stage job mem_reg2aln would need for this region (false: no banded DP - gap-free path or degenerate interval)
bool reg_global_job(const mem_opt_t *opt, const bntseq_t *bns, const mem_alnreg_t *ar, int read, GlobalJob *job);
// mem_reg2aln, reference src/bwamem.c:1089-1159
void reg2aln(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_query, const char *query,
             const mem_alnreg_t *ar, Aln *out);
// mem_gen_alt, reference src/bwamem_extra.c:98-140.  Returns false when the reference returns NULL.
bool gen_alt(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const RegVec &a, int l_query,
             const char *query, std::vector<std::string> &XA);
// mem_aln2sam, reference src/bwamem.c:825-946
void aln2sam(const mem_opt_t *opt, const bntseq_t *bns, std::string &str, const bseq1_t *s, int n, const Aln *list,
             int which, const Aln *mate);
// mem_reg2sam, reference src/bwamem.c:1003-1049
void reg2sam(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, bseq1_t *s, RegVec &a, int extra_flag,
             const Aln *mate);
// the part of mem_sam_pe after mate rescue, reference src/bwamem_pair.c:277-393
void sam_pe_finish(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4],
                   uint64_t id, bseq1_t s[2], RegVec a[2]);

// Cycle accounting of the SAM sweep (B200_HOST_PROF=1; printed per call at verbosity >= 3): where the host threads spend
// their time.  Thread-local accumulators, flushed by host_prof_flush() at the end of a block of pairs.
enum { HP_SAM_PE = 0, HP_MARK_PRIMARY, HP_PAIR, HP_GEN_ALT, HP_REG2ALN, HP_GEN_CIGAR, HP_ALN2SAM, HP_DUP, HP_REG2SAM, HP_N };
void host_prof_flush();
void host_prof_report(const char *what);      // prints and resets

char *dup_cstr(const std::string &s);   // malloc()ed copy (the host free()s seqs[i].sam)

} // namespace b200
