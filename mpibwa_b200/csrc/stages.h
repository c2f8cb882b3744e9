// stages.h - interface between the host orchestration (pipeline.cpp) and the device stages.
//
// stages_cuda.cu implements it with sm_100a kernels and is the only implementation in the shipped library.
// tests/hostemu/stages_emu.cpp implements the same interface by looping the per-thread task bodies on the CPU;
// it exists so that the host orchestration can be unit-tested in a container without a GPU and is never part
// of the product.
#pragma once
#include <cstdint>
#include <vector>
#include "../../include/mpibwa_b200.h"
#include "fm_kernels.h"
#include "ext_kernels.h"
#include "sw_kernels.h"
#include "global_kernels.h"
#include "chain_kernels.h"
#include "finish_kernels.h"
#include "fastq_kernels.h"

namespace b200 {

// result of stage_seed: arrays owned by the engine (page-locked on the CUDA engine), valid until the next stage_seed call
struct SeedOut { const int64_t *seed_off; const SeedRec *seeds; const int32_t *l_rep; int64_t n_seeds; };

struct Stats : b200_stats_t {};

class Engine;               // owns the device index, streams and scratch

Engine *engine_create(const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int device);
Engine *engine_clone(Engine *base);           // own stream + scratch + resident reads, index shared with base (destroy clones first)
void    engine_destroy(Engine *e);
Stats  &engine_stats(Engine *e);
const char *engine_kind();                     // "cuda" or "hostemu"
int     engine_device_count();

// host staging area (page-locked on the CUDA engine) in which the caller may assemble the encoded reads before the upload
uint8_t *stage_read_buffer(Engine *e, int64_t bytes);
// upload the encoded reads of the batch (codes 0-4, read r at codes[off[r] .. off[r+1]))
void stage_upload_reads(Engine *e, int n_reads, const int64_t *off, const uint8_t *codes);

// seeding + SA look-up: per read the seed list in mem_chain() order (interval order x SA order), with the
// contig id already resolved (rid < 0 = bridging, to be dropped by the caller), and l_rep per read.
// With keep_on_device the seed list stays in HBM for stage_chain() and only out.n_seeds is filled in.
void stage_seed(Engine *e, const SeedOpt &so, SeedOut &out, bool keep_on_device = false);

// page-locked host scratch owned by the engine (a handful of numbered slots); a slot's contents stay valid until the
// same slot is requested again.  The pipeline assembles stage inputs in place there and reads stage outputs from there.
void *stage_pinned(Engine *e, int slot, size_t bytes);
enum { PIN_CHAIN_OFF = 0, PIN_CHAINS = 1, PIN_DSEEDS = 2, PIN_SRT = 3, PIN_REGS = 4, PIN_REG_OFF = 5, PIN_RTEXT = 6, PIN_TEXT = 7, PIN_N_SLOTS = 8 };

// extension: chains of read r are chains[chain_off[r] .. chain_off[r+1]); the regions of read r are left in HBM, compacted as
// regs[reg_off[r] .. reg_off[r+1]) in the order mem_chain2aln appends them, for stage_finish().
// (on_device: the four input arrays were left in HBM by stage_chain(); the pointers are then unused)
struct ExtIn { int n_reads; const int32_t *chain_off; const DChain *chains; int64_t n_chains; const DSeed *seeds; int64_t n_seeds; const int32_t *srt; bool on_device = false; };
void stage_extend(Engine *e, const ExtOpt &eo, const ExtIn &in);
// host copy of the regions of the last stage_extend call (the single-job mem_chain2aln wrapper; PIN_REGS / PIN_REG_OFF slots)
struct ExtRegs { const DReg *regs; const int64_t *reg_off; };
void stage_extend_download(Engine *e, ExtRegs &out);

// the chunk straight from its raw fastq bytes (fastq_kernels.h): uploaded once, parsed, interleaved and encoded on the device;
// replaces stage_upload_reads + stage_upload_text.  fq2 == null: single-end.  Aborts when the files hold different record counts.
struct FastqInfo { int n_reads; int64_t n_bases; int max_len; };
void stage_upload_fastq(Engine *e, const char *fq1, int64_t len1, const char *fq2, int64_t len2, FastqInfo *info);

// names, qualities and comments of the batch's reads (for the SAM text): rtext[r] holds offsets into text[0..bytes)
void stage_upload_text(Engine *e, int n_reads, const ReadText *rtext, const char *text, int64_t bytes);

// Everything after extension (finish_stage.h): de-duplication, insert-size statistics, mate rescue, pairing, CIGARs, SAM text.
// The text comes back NUL-terminated in page-locked host memory owned by the engine (valid until the next call on the same
// engine), or in the buffer a.alloc returns; sam_off[r] .. sam_off[r+1] are the records of read r.
struct FinishArgs {
	const mem_opt_t *opt; const mem_pestat_t *pes0; int64_t n_processed; const char *rg_id;
	bool want_offsets;              // also bring the per-read offsets back
	int route;                      // B200_ROUTE_* (finish_stage.h ROUTE_*): bring back the per-line routing table, or the text grouped by contig
	void *(*alloc)(size_t bytes);   // when set: called once with the size of the text (+1); the text is copied there instead
};
struct SamChunk { char *sam; const int64_t *sam_off; int64_t bytes; const SamLine *lines; int64_t n_lines; const int64_t *dest_off; int n_dest; };
void stage_finish(Engine *e, const FinishArgs &a);                        // the kernels; leaves the text in HBM
void stage_fetch_sam(Engine *e, const FinishArgs &a, SamChunk &out);      // brings it to the host (copy engine only)

// page-locked host memory for buffers the library hands to its caller (b200_big_alloc)
void *stage_host_alloc(size_t bytes);
void  stage_host_free(void *p);

// kernel-isolated replay of every ksw_extend2 job the last stage_extend call recorded (B200_EXT_RECORD set): one batch, DP kernels only
double stage_extend_replay(Engine *e, const ExtOpt &eo, int64_t *cells, int64_t *n_jobs);

// chaining stage (mem_chain + mem_chain_flt + flattening, chain_kernels.h) over the seeds of the last stage_seed(keep_on_device)
// call; fills `in` for stage_extend.  Valid for reads to which mem_flt_chained_seeds does not apply (see pipeline.cpp).
// With `download` the four arrays are also copied to the PIN_CHAIN_OFF.. slots and `in` points at them (checking mode).
void stage_chain(Engine *e, const ChainOpt &co, ExtIn &in, bool download = false);

// local SW batch against reference windows
void stage_sw(Engine *e, const SwOpt &so, const std::vector<SwJob> &jobs, std::vector<SwRes> &out);

// b200_ksw_global2_batch: the caller's jobs through the CIGAR-stage kernels (queries codes 0-4, targets codes 0-3); cigar receives
// the operations of all jobs back to back, jobs[i].cigar_off / n_cigar say where
void stage_global_batch(Engine *e, const GlobalOpt &go, int64_t n_jobs, b200_global_job_t *jobs, const uint8_t *query, int64_t qbytes,
                        const uint8_t *target, int64_t tbytes, std::vector<uint32_t> &cigar);
// ksw_global2 for caller-provided byte buffers (the single-job C wrapper): score and CIGAR of one banded global alignment
int stage_global_bytes(Engine *e, const GlobalOpt &go, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int w,
                       std::vector<uint32_t> *cigar);

// generic batches over caller-provided byte buffers (C-ABI b200_*_batch and the single-job wrappers)
void stage_extend_bytes(Engine *e, const ExtOpt &eo, int64_t n_jobs, b200_extend_job_t *jobs,
                        const uint8_t *query, int64_t qbytes, const uint8_t *target, int64_t tbytes);
void stage_sw_bytes(Engine *e, const SwOpt &so, int64_t n_jobs, b200_align_job_t *jobs,
                    const uint8_t *query, int64_t qbytes, const uint8_t *target, int64_t tbytes);
void stage_collect_intv(Engine *e, const SeedOpt &so, int n_reads, const int64_t *off, const uint8_t *codes,
                        std::vector<int64_t> &intv_off, std::vector<Intv> &intv);
void stage_sa(Engine *e, int64_t n, const uint64_t *k, uint64_t *sa);
// bwt_smem1 (reference src/bwt.c:289-356) for one query position: all SMEMs through x with interval size >= min_intv; returns the next x
int  stage_smem1(Engine *e, int len, const uint8_t *q, int x, uint64_t min_intv, std::vector<Intv> &mem);
void stage_fm_extend(Engine *e, const Intv &ik, Intv ok[4], int is_back);

} // namespace b200
