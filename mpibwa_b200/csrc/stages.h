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

namespace b200 {

// result of stage_seed: arrays owned by the engine (page-locked on the CUDA engine), valid until the next stage_seed call
struct SeedOut { const int64_t *seed_off; const SeedRec *seeds; const int32_t *l_rep; int64_t n_seeds; };

struct SwJob {
	int64_t rb;             // first reference position of the target window (forward+reverse coordinate)
	int32_t tlen;
	int32_t read;           // index of the query read in the current batch
	int32_t is_rev;         // use the reverse complement of the read as the query
	int32_t xtra;
	int32_t q_beg, q_len;   // sub-range of the read used as query (whole read for mate rescue)
};

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
enum { PIN_CHAIN_OFF = 0, PIN_CHAINS = 1, PIN_DSEEDS = 2, PIN_SRT = 3, PIN_REGS = 4, PIN_REG_OFF = 5, PIN_N_SLOTS = 8 };

// extension: chains of read r are chains[chain_off[r] .. chain_off[r+1]); the regions of read r come back compacted
// as regs[reg_off[r] .. reg_off[r+1]) in the order mem_chain2aln appends them (arrays in the PIN_REGS / PIN_REG_OFF slots).
// (on_device: the four arrays were left in HBM by stage_chain(); the pointers are then unused)
struct ExtIn { int n_reads; const int32_t *chain_off; const DChain *chains; int64_t n_chains; const DSeed *seeds; int64_t n_seeds; const int32_t *srt; bool on_device = false; };
struct ExtRegs { const DReg *regs; const int64_t *reg_off; };
void stage_extend(Engine *e, const ExtOpt &eo, const ExtIn &in, ExtRegs &out);

// kernel-isolated replay of every ksw_extend2 job the last stage_extend call recorded (B200_EXT_RECORD set): one batch, DP kernels only
double stage_extend_replay(Engine *e, const ExtOpt &eo, int64_t *cells, int64_t *n_jobs);

// chaining stage (mem_chain + mem_chain_flt + flattening, chain_kernels.h) over the seeds of the last stage_seed(keep_on_device)
// call; fills `in` for stage_extend.  Valid for reads to which mem_flt_chained_seeds does not apply (see pipeline.cpp).
// With `download` the four arrays are also copied to the PIN_CHAIN_OFF.. slots and `in` points at them (checking mode).
void stage_chain(Engine *e, const ChainOpt &co, ExtIn &in, bool download = false);

// local SW batch against reference windows
void stage_sw(Engine *e, const SwOpt &so, const std::vector<SwJob> &jobs, std::vector<SwRes> &out);

// CIGAR stage: banded global alignment + traceback of regions of the resident reads (jobs carry zoff/slot filled by the caller)
// Returns the results in job order, in memory owned by the engine (page-locked on the CUDA engine) that stays valid until
// the next stage_global call on the same engine.
const GlobalRes *stage_global(Engine *e, const GlobalOpt &go, const std::vector<GlobalJob> &jobs, int64_t z_bytes);

// generic batches over caller-provided byte buffers (C-ABI b200_*_batch and the single-job wrappers)
void stage_extend_bytes(Engine *e, const ExtOpt &eo, int64_t n_jobs, b200_extend_job_t *jobs,
                        const uint8_t *query, int64_t qbytes, const uint8_t *target, int64_t tbytes);
void stage_sw_bytes(Engine *e, const SwOpt &so, int64_t n_jobs, b200_align_job_t *jobs,
                    const uint8_t *query, int64_t qbytes, const uint8_t *target, int64_t tbytes);
void stage_collect_intv(Engine *e, const SeedOpt &so, int n_reads, const int64_t *off, const uint8_t *codes,
                        std::vector<int64_t> &intv_off, std::vector<Intv> &intv);
void stage_sa(Engine *e, int64_t n, const uint64_t *k, uint64_t *sa);
void stage_fm_extend(Engine *e, const Intv &ik, Intv ok[4], int is_back);

} // namespace b200
