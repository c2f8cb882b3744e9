/* TEST ONLY: prints sizeof / offsetof of every ABI type the mpiBWA hosts pass by pointer.  Compiled twice by
 * tests/test_capi_symbols.py - once with -DUSE_REF against the REFERENCE's own headers (/root/reference/src, when present),
 * once against include/mpibwa_b200.h - and the two outputs must be identical. */
#include <stdio.h>
#include <stddef.h>
#include <stdint.h>
#ifdef USE_REF
#include "bwamem.h"
#include "bwa.h"
#include "bwt.h"
#include "bntseq.h"
#include "ksw.h"
#else
#include "mpibwa_b200.h"
#endif
#define S(t) printf("sizeof %s %zu\n", #t, sizeof(t))
#define O(t, f) printf("offsetof %s.%s %zu\n", #t, #f, offsetof(t, f))
int main(void)
{
	S(mem_opt_t);
	O(mem_opt_t, a); O(mem_opt_t, b); O(mem_opt_t, o_del); O(mem_opt_t, e_del); O(mem_opt_t, o_ins); O(mem_opt_t, e_ins); O(mem_opt_t, pen_unpaired);
	O(mem_opt_t, pen_clip5); O(mem_opt_t, pen_clip3); O(mem_opt_t, w); O(mem_opt_t, zdrop); O(mem_opt_t, max_mem_intv); O(mem_opt_t, T); O(mem_opt_t, flag);
	O(mem_opt_t, min_seed_len); O(mem_opt_t, min_chain_weight); O(mem_opt_t, max_chain_extend); O(mem_opt_t, split_factor); O(mem_opt_t, split_width);
	O(mem_opt_t, max_occ); O(mem_opt_t, max_chain_gap); O(mem_opt_t, n_threads); O(mem_opt_t, chunk_size); O(mem_opt_t, mask_level); O(mem_opt_t, drop_ratio);
	O(mem_opt_t, XA_drop_ratio); O(mem_opt_t, mask_level_redun); O(mem_opt_t, mapQ_coef_len); O(mem_opt_t, mapQ_coef_fac); O(mem_opt_t, max_ins);
	O(mem_opt_t, max_matesw); O(mem_opt_t, max_XA_hits); O(mem_opt_t, max_XA_hits_alt); O(mem_opt_t, mat);
	S(mem_alnreg_t);
	O(mem_alnreg_t, rb); O(mem_alnreg_t, re); O(mem_alnreg_t, qb); O(mem_alnreg_t, qe); O(mem_alnreg_t, rid); O(mem_alnreg_t, score); O(mem_alnreg_t, truesc);
	O(mem_alnreg_t, sub); O(mem_alnreg_t, alt_sc); O(mem_alnreg_t, csub); O(mem_alnreg_t, sub_n); O(mem_alnreg_t, w); O(mem_alnreg_t, seedcov);
	O(mem_alnreg_t, secondary); O(mem_alnreg_t, secondary_all); O(mem_alnreg_t, seedlen0); O(mem_alnreg_t, frac_rep); O(mem_alnreg_t, hash);
	S(mem_alnreg_v); O(mem_alnreg_v, n); O(mem_alnreg_v, m); O(mem_alnreg_v, a);
	S(mem_pestat_t); O(mem_pestat_t, low); O(mem_pestat_t, high); O(mem_pestat_t, failed); O(mem_pestat_t, avg); O(mem_pestat_t, std);
	S(bwt_t); O(bwt_t, primary); O(bwt_t, L2); O(bwt_t, seq_len); O(bwt_t, bwt_size); O(bwt_t, bwt); O(bwt_t, cnt_table); O(bwt_t, sa_intv); O(bwt_t, n_sa); O(bwt_t, sa);
	S(bwtintv_t); O(bwtintv_t, x); O(bwtintv_t, info);
	S(bwtintv_v); O(bwtintv_v, n); O(bwtintv_v, m); O(bwtintv_v, a);
	S(bntann1_t); O(bntann1_t, offset); O(bntann1_t, len); O(bntann1_t, n_ambs); O(bntann1_t, gi); O(bntann1_t, is_alt); O(bntann1_t, name); O(bntann1_t, anno);
	S(bntamb1_t); O(bntamb1_t, offset); O(bntamb1_t, len); O(bntamb1_t, amb);
	S(bntseq_t); O(bntseq_t, l_pac); O(bntseq_t, n_seqs); O(bntseq_t, seed); O(bntseq_t, anns); O(bntseq_t, n_holes); O(bntseq_t, ambs); O(bntseq_t, fp_pac);
	S(bwaidx_t); O(bwaidx_t, bwt); O(bwaidx_t, bns); O(bwaidx_t, pac); O(bwaidx_t, is_shm); O(bwaidx_t, l_mem); O(bwaidx_t, mem);
	S(bseq1_t); O(bseq1_t, l_seq); O(bseq1_t, id); O(bseq1_t, name); O(bseq1_t, comment); O(bseq1_t, seq); O(bseq1_t, qual); O(bseq1_t, sam);
	S(kswr_t); O(kswr_t, score); O(kswr_t, te); O(kswr_t, qe); O(kswr_t, score2); O(kswr_t, te2); O(kswr_t, tb); O(kswr_t, qb);
	printf("MEM_F_PE %d MEM_F_NOPAIRING %d MEM_F_ALL %d MEM_F_NO_MULTI %d MEM_F_NO_RESCUE %d MEM_F_REF_HDR %d MEM_F_SOFTCLIP %d MEM_F_SMARTPE %d MEM_F_PRIMARY5 %d MEM_F_KEEP_SUPP_MAPQ %d\n",
	       MEM_F_PE, MEM_F_NOPAIRING, MEM_F_ALL, MEM_F_NO_MULTI, MEM_F_NO_RESCUE, MEM_F_REF_HDR, MEM_F_SOFTCLIP, MEM_F_SMARTPE, MEM_F_PRIMARY5, MEM_F_KEEP_SUPP_MAPQ);
	printf("KSW_XBYTE %d KSW_XSTOP %d KSW_XSUBO %d KSW_XSTART %d BWA_IDX_ALL %d\n", KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART, BWA_IDX_ALL);
	return 0;
}
