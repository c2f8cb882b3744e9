/* TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the reference algorithms on the hot path, used as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Never linked into, imported by or
 * executed from the product library (mpibwa_b200/libmpibwa_b200.so).
 *
 * Parity pinning: the reference has no golden vectors or tests for this path, so every function here
 * is pinned against the UNMODIFIED reference compiled into oracle/_ref/libbwa_ref.so (tests/test_oracle_pinned.py
 * fuzzes each one against ksw_extend2 / ksw_align2 / ksw_global2 / bwt_extend / bwt_smem1 / bwt_seed_strategy1 / bwt_sa /
 * bns_fetch_seq of that library).  Without oracle/_ref the oracle is "parity unpinned".
 */
#ifndef B200_ORACLE_H
#define B200_ORACLE_H
#include <stdint.h>

typedef struct { int score, qle, tle, gtle, gscore, max_off; int64_t cells; } orc_ext_t;
typedef struct { int score, te, qe, score2, te2, tb, qb; int64_t cells; } orc_aln_t;
typedef struct { uint64_t x0, x1, x2, info; } orc_intv_t;
typedef struct {
	const uint32_t *bwt;           /* occ-interleaved BWT (reference src/bwt.h:72-78) */
	const uint64_t *sa;            /* sampled suffix array */
	uint64_t primary, L2[5], seq_len;
	int sa_intv;
} orc_fm_t;

/* reference src/ksw.c:380-479 */
void orc_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t mat[25],
                     int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, orc_ext_t *out);
/* reference src/ksw.c:63-365 (ksw_align2 with qry == NULL) */
void orc_ksw_align2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t mat[25],
                    int o_del, int e_del, int o_ins, int e_ins, int xtra, orc_aln_t *out);
/* reference src/ksw.c:504-606; *cigar is malloc()ed when asked for; *cells = band cells computed */
int orc_ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t mat[25],
                    int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t **cigar, int64_t *cells);
/* reference src/bwt.c:107-129,169-186 */
void orc_occ4(const orc_fm_t *fm, uint64_t k, uint64_t cnt[4]);
/* reference src/bwt.c:262-275 */
void orc_extend(const orc_fm_t *fm, const orc_intv_t *ik, orc_intv_t ok[4], int is_back);
/* reference src/bwt.c:289-356; mem must hold len+1 entries; returns next x */
int orc_smem1(const orc_fm_t *fm, int len, const uint8_t *q, int x, uint64_t min_intv, orc_intv_t *mem, int *n_mem);
/* reference src/bwt.c:358-379 */
int orc_seed_strategy1(const orc_fm_t *fm, int len, const uint8_t *q, int x, int min_len, int max_intv, orc_intv_t *mem);
/* reference src/bwamem.c:114-162; returns the number of intervals (out must hold 3*len+8) */
int orc_collect_intv(const orc_fm_t *fm, int min_seed_len, float split_factor, int split_width, int max_mem_intv,
                     int len, const uint8_t *seq, orc_intv_t *out);
/* reference src/bwt.c:53-59,86-96 */
uint64_t orc_sa(const orc_fm_t *fm, uint64_t k);
/* reference src/bntseq.c:398-419 (bns_get_seq) */
int64_t orc_get_seq(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, uint8_t *out);
#endif
