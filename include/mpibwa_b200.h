/* mpibwa_b200.h - C ABI of the B200-native BWA-MEM alignment core.
 *
 * This library replaces, symbol for symbol, the part of libbwa.a that the mpiBWA hosts
 * (reference src/mainParallel.c, src/mainParallelByChromosome.c, src/parallel_aux.c, src/pidx.c) link against
 * (reference src/Makefile.am:4-15).  Struct layouts below are the x86-64 LP64 layouts of the reference structs
 * because the hosts pass them by pointer and the `.map` index image is a raw dump of them.
 * Each declaration cites the reference declaration it stands in for.
 *
 * Everything prefixed b200_ is an addition that the reference does not have (device set-up, batched kernels,
 * counters for the bench harness).  There is NO CPU fallback: every entry point that computes aborts with a
 * message when no CUDA device is usable.
 */
#ifndef MPIBWA_B200_H
#define MPIBWA_B200_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ index types */

typedef uint64_t bwtint_t;                       /* reference src/bwt.h:44 */

/* FM-index. reference src/bwt.h:46-58.  bwt[] is the occ-interleaved BWT: per 128 symbols one 64-byte block
 * = uint64 occ[4] (counts before the block) followed by 8 words of 2-bit symbols, first symbol in the top bits. */
typedef struct {
	bwtint_t primary;
	bwtint_t L2[5];
	bwtint_t seq_len;
	bwtint_t bwt_size;
	uint32_t *bwt;
	uint32_t cnt_table[256];
	int sa_intv;
	bwtint_t n_sa;
	bwtint_t *sa;
} bwt_t;

typedef struct { bwtint_t x[3], info; } bwtintv_t;            /* reference src/bwt.h:60-62 */
typedef struct { size_t n, m; bwtintv_t *a; } bwtintv_v;      /* reference src/bwt.h:64 */

typedef struct {                                              /* reference src/bntseq.h:44-51 */
	int64_t offset;
	int32_t len;
	int32_t n_ambs;
	uint32_t gi;
	int32_t is_alt;
	char *name, *anno;
} bntann1_t;

typedef struct {                                              /* reference src/bntseq.h:53-57 */
	int64_t offset;
	int32_t len;
	char amb;
} bntamb1_t;

typedef struct {                                              /* reference src/bntseq.h:59-67 */
	int64_t l_pac;
	int32_t n_seqs;
	uint32_t seed;
	bntann1_t *anns;
	int32_t n_holes;
	bntamb1_t *ambs;
	FILE *fp_pac;
} bntseq_t;

typedef struct {                                              /* reference src/bwa.h:20-28 */
	bwt_t    *bwt;
	bntseq_t *bns;
	uint8_t  *pac;
	int    is_shm;
	int64_t l_mem;
	uint8_t  *mem;
} bwaidx_t;

#define BWA_IDX_BWT 0x1                                       /* reference src/bwa.h:9-12 */
#define BWA_IDX_BNS 0x2
#define BWA_IDX_PAC 0x4
#define BWA_IDX_ALL 0x7

/* ------------------------------------------------------------------ read / option types */

typedef struct {                                              /* reference src/bwa.h:30-33 */
	int l_seq, id;
	char *name, *comment, *seq, *qual, *sam;
} bseq1_t;

#define MEM_F_PE             0x2                              /* reference src/bwamem.h:14-24 */
#define MEM_F_NOPAIRING      0x4
#define MEM_F_ALL            0x8
#define MEM_F_NO_MULTI       0x10
#define MEM_F_NO_RESCUE      0x20
#define MEM_F_REF_HDR        0x100
#define MEM_F_SOFTCLIP       0x200
#define MEM_F_SMARTPE        0x400
#define MEM_F_PRIMARY5       0x800
#define MEM_F_KEEP_SUPP_MAPQ 0x1000

typedef struct {                                              /* reference src/bwamem.h:25-57 */
	int a, b;
	int o_del, e_del;
	int o_ins, e_ins;
	int pen_unpaired;
	int pen_clip5, pen_clip3;
	int w;
	int zdrop;
	uint64_t max_mem_intv;
	int T;
	int flag;
	int min_seed_len;
	int min_chain_weight;
	int max_chain_extend;
	float split_factor;
	int split_width;
	int max_occ;
	int max_chain_gap;
	int n_threads;
	int chunk_size;
	float mask_level;
	float drop_ratio;
	float XA_drop_ratio;
	float mask_level_redun;
	float mapQ_coef_len;
	int mapQ_coef_fac;
	int max_ins;
	int max_matesw;
	int max_XA_hits, max_XA_hits_alt;
	int8_t mat[25];
} mem_opt_t;

typedef struct {                                              /* reference src/bwamem.h:59-77 */
	int64_t rb, re;
	int qb, qe;
	int rid;
	int score;
	int truesc;
	int sub;
	int alt_sc;
	int csub;
	int sub_n;
	int w;
	int seedcov;
	int secondary;
	int secondary_all;
	int seedlen0;
	int n_comp:30, is_alt:2;
	float frac_rep;
	uint64_t hash;
} mem_alnreg_t;

typedef struct { size_t n, m; mem_alnreg_t *a; } mem_alnreg_v;  /* reference src/bwamem.h:79 */

typedef struct {                                              /* reference src/bwamem.h:81-85 */
	int low, high;
	int failed;
	double avg, std;
} mem_pestat_t;

/* chain of seeds; private to the reference's bwamem.c (src/bwamem.c:168-182) but part of the mem_chain2aln
 * call surface, so it is spelled out here with the same layout. */
typedef struct {
	int64_t rbeg;
	int32_t qbeg, len;
	int score;
} mem_seed_t;

typedef struct {
	int n, m, first, rid;
	uint32_t w:29, kept:2, is_alt:1;
	float frac_rep;
	int64_t pos;
	mem_seed_t *seeds;
} mem_chain_t;

typedef struct {                                              /* reference src/ksw.h:14-19 */
	int score;
	int te, qe;
	int score2, te2;
	int tb, qb;
} kswr_t;

#define KSW_XBYTE  0x10000                                    /* reference src/ksw.h:6-9 */
#define KSW_XSTOP  0x20000
#define KSW_XSUBO  0x40000
#define KSW_XSTART 0x80000

struct _kswq_t;                                               /* opaque; never dereferenced here */
typedef struct _kswq_t kswq_t;

/* ------------------------------------------------------------------ symbols the mpiBWA hosts link against */

extern int   bwa_verbose;                                     /* reference src/bwa.c:16 */
extern char  bwa_rg_id[256];                                  /* reference src/bwa.c:17 */
extern char *bwa_pg;                                          /* reference src/bwa.c:18 */

mem_opt_t *mem_opt_init(void);                                /* reference src/bwamem.c:48  (host free()s it) */
void  bwa_fill_scmat(int a, int b, int8_t mat[25]);           /* reference src/bwa.c:109 */
char *bwa_set_rg(const char *s);                              /* reference src/bwa.c:431 */
char *bwa_insert_header(const char *s, char *hdr);            /* reference src/bwa.c:464 */
int   bwa_mem2idx(int64_t l_mem, uint8_t *mem, bwaidx_t *idx);/* reference src/bwa.c:310 (.map image -> pointers) */
int   bwa_idx2mem(bwaidx_t *idx);                             /* reference src/bwa.c:347 (pointers -> .map image) */
bwaidx_t *bwa_idx_load(const char *hint, int which);          /* reference src/bwa.c:291 (.bwt .sa .ann .amb .pac) */
void  bwa_idx_destroy(bwaidx_t *idx);                         /* reference src/bwa.c:296 */

/* THE hot path. reference src/bwamem.c:1205.  Same contract: seqs[i].seq is overwritten with codes 0-4,
 * seqs[i].sam is malloc()ed (caller frees), PE mates interleaved, pes0 NULL => infer per call.
 * The index is uploaded to the current CUDA device on first use (keyed by the bwt pointer) unless
 * b200_gpu_init() was called before.  CUDA errors print and abort() like the reference's err_fatal. */
void mem_process_seqs(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                      int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0);

/* ------------------------------------------------------------------ inner call surface (single-job wrappers)
 * Same signatures as the reference; each runs as a batch of one on the GPU, on an engine of its own (own stream, scratch and
 * resident reads: a call never disturbs a chunk job in flight) and one call at a time (they serialise on a mutex), so they may be
 * called from several threads.  ksw_extend2 requires h0 > 0 and mem_chain2aln an empty region list (both abort otherwise): the
 * hot path never calls them any other way. */

int ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore, int *max_off);          /* reference src/ksw.c:380 */
kswr_t ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
                  int o_del, int e_del, int o_ins, int e_ins, int xtra, kswq_t **qry); /* reference src/ksw.c:343 */
int ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t **cigar); /* reference src/ksw.c:504 */
void bwt_extend(const bwt_t *bwt, const bwtintv_t *ik, bwtintv_t ok[4], int is_back); /* reference src/bwt.c:262 */
int  bwt_smem1(const bwt_t *bwt, int len, const uint8_t *q, int x, int min_intv, bwtintv_v *mem,
               bwtintv_v *tmpvec[2]);                                                /* reference src/bwt.c:353 */
bwtint_t bwt_sa(const bwt_t *bwt, bwtint_t k);                                       /* reference src/bwt.c:86 */
void mem_chain2aln(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_query,
                   const uint8_t *query, const mem_chain_t *c, mem_alnreg_v *av);    /* reference src/bwamem.c:632 */

/* ------------------------------------------------------------------ B200 additions */

/* Select the CUDA device and upload the index (occ-interleaved BWT, SA samples, pac, contig table) into HBM, then build the
 * structures derived from it there (k-mer interval tables, the whole suffix array and its inverse, Bloom filters over the text's
 * 19-mers: 107 GB for a human-sized reference, under two seconds; each one is skipped when it would leave less than
 * B200_TABLE_RESERVE_GB (72) of device memory free, and B200_KMER_MAX=0 / B200_SA_FULL=0 / B200_BLOOM=0 switch them off).
 * Returns 0 on success; aborts on CUDA failure.  The three-line patch for the mpiBWA hosts calls this right
 * after map_indexes(). */
int  b200_gpu_init(const bwaidx_t *idx, int device);
/* Optional: encode the reads of the coming mem_process_seqs(…, n, seqs, …) call in place and make them resident in HBM
 * ahead of time (the call then skips its own encode + upload).  Used to time the path with device-resident inputs. */
void b200_stage_reads(const mem_opt_t *opt, const bwaidx_t *idx, int n, bseq1_t *seqs);
/* (a staged chunk occupies one of the library's four chunk slots until the call for the same `seqs` takes it over: stage at most
 *  four chunks ahead, or the next stage / begin call waits for a slot) */
void b200_gpu_release(void);
int  b200_device_count(void);

/* batched seed extension: job j reads query[q_off[j] .. +qlen[j]) and target[t_off[j] .. +tlen[j]) (codes 0-4). */
typedef struct {
	int32_t qlen, tlen;       /* in */
	int64_t q_off, t_off;     /* in: offsets into the flat code buffers */
	int32_t h0, w, end_bonus; /* in */
	int32_t score, qle, tle, gtle, gscore, max_off; /* out (reference ksw_extend2 return value and out-params) */
} b200_extend_job_t;
int b200_ksw_extend2_batch(int64_t n_jobs, b200_extend_job_t *jobs, const uint8_t *query, int64_t query_bytes,
                           const uint8_t *target, int64_t target_bytes, const int8_t mat[25],
                           int o_del, int e_del, int o_ins, int e_ins, int zdrop);

/* batched local SW (mate rescue flavour of ksw_align2). */
typedef struct {
	int32_t qlen, tlen;
	int64_t q_off, t_off;
	int32_t xtra;
	kswr_t  r;                /* out */
} b200_align_job_t;
int b200_ksw_align2_batch(int64_t n_jobs, b200_align_job_t *jobs, const uint8_t *query, int64_t query_bytes,
                          const uint8_t *target, int64_t target_bytes, const int8_t mat[25],
                          int o_del, int e_del, int o_ins, int e_ins);

/* batched banded global alignment with traceback (ksw_global2, reference src/ksw.c:504-606) through the kernels of the CIGAR stage.
 * Queries hold codes 0-4, targets codes 0-3 (the stage reads its targets from the 2-bit reference).  *cigar is one malloc()ed
 * array (release with b200_free): job j's operations are (*cigar)[cigar_off .. cigar_off + n_cigar), len << 4 | op. */
typedef struct {
	int32_t qlen, tlen;       /* in */
	int64_t q_off, t_off;     /* in: offsets into the flat code buffers */
	int32_t w;                /* in: band */
	int32_t score, n_cigar;   /* out */
	int64_t cigar_off;        /* out */
} b200_global_job_t;
int b200_ksw_global2_batch(int64_t n_jobs, b200_global_job_t *jobs, const uint8_t *query, int64_t query_bytes,
                           const uint8_t *target, int64_t target_bytes, const int8_t mat[25],
                           int o_del, int e_del, int o_ins, int e_ins, uint32_t **cigar);

/* batched seeding: for read r (codes 0-4 at seq[off[r] .. off[r+1])) the sorted interval list of
 * mem_collect_intv (reference src/bwamem.c:114-162) is returned in a malloc()ed array; *intv_off has n+1 entries. */
int b200_collect_intv_batch(const mem_opt_t *opt, int n_reads, const int64_t *off, const uint8_t *seq,
                            bwtintv_t **intv, int64_t **intv_off);

/* batched suffix-array look-up (reference src/bwt.c:86-96) */
int b200_bwt_sa_batch(int64_t n, const bwtint_t *k, bwtint_t *sa);

/* host-side lines of the mpiBWA mains around mem_process_seqs, for harnesses that stand in for the MPI host:
 * in-place fastq parse (reference src/mainParallel.c:1257-1304), the chunk rule "close when bases > maxsiz"
 * (reference src/parallel_aux.c:1532-1549, 1068-1082; src/mainParallel.c:2773) and "interleave mates, align,
 * concatenate seqs[i].sam" (reference src/mainParallel.c:1271-1314, 103-127).  Buffers come from malloc();
 * release them with b200_free(). */
int64_t b200_fastq_parse(char *buf, int64_t len, bseq1_t **seqs);
void    b200_set_host_threads(int n);   /* threads used by the parse / concatenation helpers (0 = all hardware threads) */
int64_t b200_plan_chunks(int64_t n, const bseq1_t *s1, const bseq1_t *s2, int64_t K, int trimmed, int64_t **ends);
bseq1_t *b200_chunk_seqs(int64_t n, const bseq1_t *s1, const bseq1_t *s2);      /* interleaved mates 2i, 2i+1 */
int64_t b200_collect_sam(int64_t total, bseq1_t *seqs, char **sam);            /* concatenates and frees seqs[i].sam */
int64_t b200_align_chunk(const mem_opt_t *opt, const bwaidx_t *idx, int64_t n_processed, int64_t n, bseq1_t *s1, bseq1_t *s2,
                         char **sam, int64_t *sam_len);
void b200_free(void *p);            /* releases any buffer the b200_* calls return (large SAM buffers are parked for reuse, the rest is free()d) */
void *b200_big_alloc(size_t bytes); /* a buffer from the same recycling pool (release with b200_free) */

/* counters filled by the last mem_process_seqs call on this thread's context; used by bench.py */
typedef struct {
	double ms_total;          /* wall time of the call */
	double ms_seed, ms_sa, ms_chain_host, ms_extend, ms_regs_host, ms_rescue, ms_sam_host; /* stage walls */
	double ms_k_smem, ms_k_sa, ms_k_extend, ms_k_sw, ms_k_global;  /* CUDA-event kernel times */
	int64_t n_reads, n_bases;
	int64_t n_intv, n_seeds, n_chains;
	int64_t n_extend_jobs, extend_cells;       /* cells = sum over executed rows of (end-beg), as the reference runs them */
	int64_t n_sw_jobs, sw_cells;
	int64_t n_global_jobs, global_cells;
	int64_t fm_occ_blocks, fm_sa_steps, fm_sa_lookups;  /* algorithmic FM-index traffic counters */
	int64_t n_launches;
	int64_t h2d_bytes, d2h_bytes;
	double ms_k_extend_dp;    /* CUDA-event time of the ksw_extend2 DP kernels alone (ms_k_extend = whole extension stage) */
	int64_t n_extend_rounds;
	double ms_sam_plan, ms_global;   /* inside ms_sam_host: the dry-run sweep that queues the CIGAR jobs; the device CIGAR stage (wall) */
	int64_t n_global_host;    /* regions whose CIGAR the SAM sweep computed with the host routine (not queued for the device stage) */
	double ms_k_chain;        /* CUDA-event time of the chaining kernels (ms_chain_host is the wall of the whole chaining stage, host or device) */
	/* finish stages (everything after seed extension runs on the device): ms_regs_host / ms_rescue / ms_sam_plan / ms_global /
	 * ms_sam_host above are now the WALLS of region de-duplication, mate rescue, pairing + record plan, the CIGAR stage and
	 * NM/MD + SAM text; the names are kept for the bench records of earlier rounds */
	double ms_k_finish;       /* CUDA-event time of all finish-stage kernels but ksw_align2 and ksw_global2 */
	double ms_k_samtext;      /* of which: NM/MD + SAM text kernels */
	double ms_upload, ms_deliver; /* host walls: read text staging + H2D; SAM D2H + hand-over (per-read malloc for mem_process_seqs) */
	int64_t n_patch_reads, n_rescue_rounds, n_global_rerun, n_aln_slots, sam_bytes;
} b200_stats_t;
/* kernel-isolated ksw_extend2: with B200_EXT_RECORD set in the environment the extension stage keeps every job of the call; this
 * replays them as ONE batch through the DP kernels on the primary engine and returns the time in ms (bench.py; run one chunk alone first) */
double b200_ext_replay(const mem_opt_t *opt, int64_t *cells, int64_t *n_jobs);
void b200_get_stats(b200_stats_t *out);   /* counters of the call that finished last */
void b200_get_aux_stats(b200_stats_t *out); /* running counters of the single-job wrappers and the b200_*_batch calls (their own engine) */

/* Chunk jobs - mem_process_seqs (reference src/bwamem.h:134) split into begin / end so that the host can keep two chunks
 * in flight: begin() returns at once and the chunk is aligned by a library thread; end() waits and leaves the result where
 * mem_process_seqs leaves it (seqs[i].sam).  Jobs run in submission order, B200_INFLIGHT (default 4) at a time, each in its
 * own set of device buffers: the device stages of chunk i+1 (seeding, chaining, extension) run under the host stages of chunk
 * i (rescue replay, pairing, SAM text), which a single synchronous call cannot overlap because the insert-size statistics
 * separate them.  The reference host loop (src/mainParallel.c:1271-1314) becomes: read chunk i+1; begin(i+1); end(i);
 * hand chunk i to the writer thread.  mem_process_seqs() itself is begin() + end().
 * b200_align_chunk_begin/_end is the same for b200_align_chunk (the SAM concatenation runs in the job thread). */
typedef struct b200_job b200_job_t;
b200_job_t *b200_process_seqs_begin(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                                    int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0);
void b200_process_seqs_end(b200_job_t *job, b200_stats_t *stats /* may be NULL */);
b200_job_t *b200_align_chunk_begin(const mem_opt_t *opt, const bwaidx_t *idx, int64_t n_processed, int64_t n, bseq1_t *s1, bseq1_t *s2);
int64_t b200_align_chunk_end(b200_job_t *job, char **sam, int64_t *sam_len, b200_stats_t *stats /* may be NULL */);
/* the same over an interleaved array (mates 2i, 2i+1; what mem_process_seqs takes - and what b200_stage_reads made resident):
 * the chunk's SAM text comes back as ONE buffer from b200_align_chunk_end instead of one malloc()ed string per read */
b200_job_t *b200_align_seqs_begin(const mem_opt_t *opt, const bwaidx_t *idx, int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0);
/* the same from the raw fastq bytes of the chunk (one buffer per mate file, fq2 NULL for single-end; parsed IN PLACE by the job
 * thread, so begin() returns at once and the buffers must stay untouched until end()); finish with b200_align_chunk_end */
b200_job_t *b200_align_fastq_begin(const mem_opt_t *opt, const bwaidx_t *idx, int64_t n_processed, char *fq1, int64_t len1, char *fq2, int64_t len2);
/* Per-chromosome routing (row f3; reference src/mainParallelByChromosome.c:1395-1457: the ByChr host parses RNAME and RNEXT back out
 * of every SAM line and copies the line into the buffer of its contig, of "unmapped" or - without fixmate - also of "discordant").
 * The device writes the lines, so it knows both contigs.  b200_set_routing(flags) applies to the one-buffer chunk jobs begun after it
 * (b200_align_chunk_begin / _seqs_begin / _fastq_begin); finish those jobs with b200_align_chunk_end_routed:
 *   B200_ROUTE_LINES      the text comes back in input order plus one b200_sam_line_t per line (malloc()ed; free() it)
 *   B200_ROUTE_BY_CONTIG  the text comes back GROUPED BY DESTINATION - contig 0 .. n_seqs-1, then "discordant" (index n_seqs), then
 *                         "unmapped" (n_seqs+1) - lines of a destination in input order; (*dest_off)[d] .. [d+1] is destination d's
 *                         range (n_dest + 1 = n_seqs + 3 entries, malloc()ed): what the reference builds in buffer_out_vec[]
 *   | B200_ROUTE_DISCORDANT  (with BY_CONTIG) a mapped line whose mate maps to another contig is ALSO copied into "discordant"
 *                         (the branch without fixmate, src/mainParallelByChromosome.c:1440-1444; single-end data has none) */
typedef struct { int64_t off; int32_t len, rid, mate_rid, read; } b200_sam_line_t;   /* rid / mate_rid: contig printed as RNAME / RNEXT, -1 for '*' */
#define B200_ROUTE_LINES 1
#define B200_ROUTE_BY_CONTIG 2
#define B200_ROUTE_DISCORDANT 4
void b200_set_routing(int flags);
int64_t b200_align_chunk_end_routed(b200_job_t *job, char **sam, int64_t *sam_len, b200_sam_line_t **lines, int64_t *n_lines,
                                    int64_t **dest_off, int *n_dest, b200_stats_t *stats /* may be NULL */);
/* measured int32 instruction issue rate of the device in Gop/s: integer ALU pipe only (min/max/add/logic; the DP
 * roofline denominator) and with half of the work as IMAD on the FMA pipe (dual-pipe ceiling) */
double b200_int32_peak(int device);
double b200_int32_peak_dual_pipe(int device);
/* measured bandwidth (GB/s) the device delivers to the FM-index kernels' access pattern: independent 256-bit loads of uniformly
 * random 32-byte sectors of a table of table_bytes (the roofline of the seeding / SA kernels once the index outgrows L2) */
double b200_hbm_random_sector_peak(int device, size_t table_bytes);
const char *b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
