/*
 * zmo_b200.h -- C ABI of libzmo_b200.so, the B200 (sm_100a) kernel library behind `wtzmo`.
 *
 * The reference (ruanjue/smartdenovo) has no FFI/plugin interface: its boundary is the `wtzmo`
 * process.  This ABI is the seam between the C host (smartdenovo_b200/csrc/host/, which replaces
 * overlap_wtzmo + thread.h, wtzmo.c:1251-1357) and the device kernels.  Every entry point replaces
 * one *pure* stage of the reference worker (wtzmo.c:803-1134); all cross-read state (masked reads,
 * tried pairs, per-read overlap counters) stays in the host replay.
 *
 * Conventions: plain pointers and sizes, caller-allocated host buffers, int return code
 * (0 = ok, <0 = error, see zmo_last_error()).  Calls are synchronous with respect to the host.
 * There is NO CPU fallback: every function fails with ZMO_ERR_CUDA if no sm_100-class device is
 * usable.
 *
 * Naming inside alignment results follows the reference's DP convention:
 *   "t" coordinates (tb,te) are on the query read q of the run (pb1, printed first),
 *   "q" coordinates (qb,qe) are on the candidate c (pb2) on the strand shown.
 */
#ifndef ZMO_B200_H
#define ZMO_B200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ZMO_OK            0
#define ZMO_ERR_CUDA     -1   /* CUDA runtime / no device */
#define ZMO_ERR_ARG      -2   /* bad argument */
#define ZMO_ERR_CAPACITY -3   /* caller buffer too small; required size reported through the *_needed out-parameter */
#define ZMO_ERR_STATE    -4   /* call order violated (e.g. query before index) */

typedef struct zmo_ctx zmo_ctx;

/* Numeric parameters of the path (wtzmo.c:1543-1588 defaults in brackets). */
typedef struct {
	int32_t hk, hz;            /* homopolymer compression of k-mers / z-mers (-H) [1,1] */
	int32_t ksize, zsize;      /* -k [16], -z [10] */
	int32_t ksave;             /* -S k-mer sub-sampling [4] */
	int32_t kovl;              /* -d min union length of k-mer hits for a candidate [300] */
	int32_t zcut, kvar;        /* -Z [64], -l [2] */
	int32_t kwin, kstep;       /* -y [800], kwin/2 */
	int32_t zovl, ztot;        /* -R [200], -r [300] */
	int32_t w, ew, W;          /* -w [50], -e [800], -W [3200] */
	int32_t M, X, O, E, T;     /* -M 2 -X -5 -O -3 -E -1 -T -50 */
	float   min_id;            /* -m, used by the per-window region filter (wtzmo.c:1026) */
	int32_t xvar, yvar, min_block_len, max_overhang;   /* -U dot-matrix parameters */
	float   deviation_penalty, gap_penalty;
} zmo_params_t;

/* ---- context ------------------------------------------------------------------------------- */
int  zmo_ctx_create(zmo_ctx **ctx, int device, const zmo_params_t *par);
/* A clone shares the root's read store and k-mer index (read-only) and has its own streams, scratch and batch slots,
 * so that independent batches can be in flight concurrently (one host thread per context).  Reads are uploaded and
 * the index is built through the root, while no clone is executing; clones are destroyed before the root. */
int  zmo_ctx_clone(zmo_ctx *root, zmo_ctx **clone);
void zmo_ctx_destroy(zmo_ctx *ctx);
/* -n (wtzmo.c:1648,1031-1034): re-align every stitched alignment with kswx_refine_alignment (kswx.h:483-659) inside the
 * band around its CIGAR before the record is judged.  Set on the root before cloning; clones inherit it. */
int  zmo_set_refine(zmo_ctx *ctx, int on);
const char *zmo_last_error(void);
/* number of kernels launched by this context so far (bench.py's gpu_launches claim) */
uint64_t zmo_kernel_launches(const zmo_ctx *ctx);
/* cumulative device time (ms) spent per stage, measured with CUDA events on the context stream:
 * out[0]=index, [1]=candidates, [2]=pair_windows, [3]=window_align, [4]=gap_global, [5]=end_extend,
 * [6]=dotmatrix, [7]=h2d/d2h copies, [8]=wall time of the concurrent end-extension + gap phase (the per-class kernel
 * times in [4],[5] overlap each other), [9..11] reserved */
void zmo_stage_ms(const zmo_ctx *ctx, double out[12]);
/* cumulative work counters: out[0]=DP cells end-extension, [1]=cells window extension,
 * [2]=cells gap global, [3]=z-mer match pairs, [4]=postings visited, [5]=bytes H2D, [6]=bytes D2H */
void zmo_counters(const zmo_ctx *ctx, uint64_t out[8]);

/* process-wide wall time spent growing buffers: out[0..2] = seconds, calls, bytes of device-buffer growth, out[3..5] = the same for
 * page-locked host buffers (what a cold start pays on top of the steady state) */
void zmo_alloc_stats(double out[6]);

/* page-locked host memory for result buffers (records, CIGARs): lets the D2H copies run at PCIe speed */
void *zmo_host_alloc(size_t bytes);
void  zmo_host_free(void *p);

/* ---- read store: replaces BaseBank + pbread_t (dna.h:318-410, wtzmo.c:87-90,207-215) -------- */
/* bank: the reference's packed layout, 32 bases per uint64, base i of the bank at bits
 * ((~i)&31)*2 of word i>>5.  rdoff/rdlen in bases (after -b clipping).  Reads are re-packed on the
 * device into a 128-bit aligned per-read layout. */
int zmo_reads_upload(zmo_ctx *ctx, const uint64_t *bank, uint64_t n_bases,
                     const uint64_t *rdoff, const uint32_t *rdlen, uint32_t n_reads);

/* ---- global k-mer index: replaces index_wtzmo (wtzmo.c:349-430) ------------------------------ */
typedef struct { uint64_t n_kmers, n_postings, n_filtered_high, n_indexed; uint32_t kcut, kavg; } zmo_index_stats_t;
/* index reads [beg,end).  *kcut_io < 2 => auto (5*max(20,avg depth)), written back. */
int zmo_index_build(zmo_ctx *ctx, uint32_t beg, uint32_t end, uint32_t *kcut_io, zmo_index_stats_t *stats);

/* ---- candidate events: replaces the merge in query_wtzmo (wtzmo.c:433-562) ------------------- */
typedef struct { uint32_t tkey; uint32_t ol; } zmo_event_t;   /* tkey = target_id<<1 | strand */
/* For each query read qids[i]: events (ascending tkey) of every (target,strand) whose k-mer union
 * length ol >= kovl.  ev_off[i]..ev_off[i+1] delimits read i's events (nq+1 entries). */
int zmo_candidates(zmo_ctx *ctx, const uint32_t *qids, uint32_t nq,
                   uint64_t *ev_off, zmo_event_t *events, uint64_t ev_cap, uint64_t *ev_needed);

/* ---- pair seeding: replaces wtzmo.c:845-914 (z-index, z-match, windows, chain) --------------- */
typedef struct { uint32_t qid, cid; } zmo_pair_t;
typedef struct {
	uint32_t n_zpair;          /* cache->size: number of z-mer match pairs */
	int32_t  ovl[2];           /* chain weight per strand (chaining_wtseedv) */
	uint32_t win_off[2], n_win[2];   /* kept windows of strand d: wins[win_off[d] .. +n_win[d]) (only if ovl[d] >= ztot) */
} zmo_pairseed_t;
typedef struct { int32_t beg[2], end[2]; } zmo_window_t;     /* [0] on q, [1] on c (strand coords) */
/* Results stay resident on the device in batch slot `slot` (0 or 1) for zmo_pair_align. */
int zmo_pair_windows(zmo_ctx *ctx, int slot, const zmo_pair_t *pairs, uint32_t np,
                     zmo_pairseed_t *seeds, zmo_window_t *wins, uint64_t win_cap, uint64_t *win_needed);

/* ---- pair alignment: replaces wtzmo.c:1011-1030 (fast_seeds_align + global_align_regs) ------- */
typedef struct { uint32_t pair_idx; uint32_t dir; } zmo_task_t;          /* index into the slot's pair list */
typedef struct {
	int32_t ok;                /* 0 = no window region survived (regs->size==0, wtzmo.c:1029) */
	int32_t score, tb, te, qb, qe, aln, mat, mis, ins, del;
	uint64_t cigar_off; uint32_t n_cigar;      /* ops = len<<4|op (0 M,1 I,2 D) in cigars[] */
} zmo_record_t;
int zmo_pair_align(zmo_ctx *ctx, int slot, const zmo_task_t *tasks, uint32_t nt,
                   zmo_record_t *recs, uint32_t *cigars, uint64_t cigar_cap, uint64_t *cigar_needed);
/* Same, but the CIGARs come back as the text `print_hits_wtzmo` prints (kswx_cigar2string, kswx.h:1093-1120: "<len><M|I|D>"
 * per op), formatted on the device: recs[i].cigar_off / n_cigar are the byte offset / byte length in cigar_text, sizes in bytes. */
int zmo_pair_align_text(zmo_ctx *ctx, int slot, const zmo_task_t *tasks, uint32_t nt,
                        zmo_record_t *recs, char *cigar_text, uint64_t text_cap, uint64_t *text_needed);
/* Same, records only (cigar_off = n_cigar = 0): for consumers that drop the CIGAR column -- the pipeline's `wtzmo ... -fo - | cut -f1-16`
 * (smartdenovo.pl:58) and the `.ovl` readers of wtclp / wtlay, which parse columns 1-12 / 1-13 (wtclp.c:123-152, wtlay.h:243-268). */
int zmo_pair_align_records(zmo_ctx *ctx, int slot, const zmo_task_t *tasks, uint32_t nt, zmo_record_t *recs);

/* ---- dot-matrix mode: replaces dot_matrix_align_hzmps for one pair (wtzmo.c:853-863) --------- */
typedef struct { uint32_t n_zpair; int32_t score, qb, qe, tb, te, strand; } zmo_dotres_t;
int zmo_pair_dotmatrix(zmo_ctx *ctx, const zmo_pair_t *pairs, uint32_t np, zmo_dotres_t *out);

/* ---- multi-GPU: the one exchange step of the path ---------------------------------------------------- */
/* Every GPU runs its own query shard = the reference's job `-P n -p g` (wtzmo.c:1291,1314) without communication; at the end the
 * record text of all jobs is gathered on the GPU of ctxs[0] over NVLink (NCCL: one all-gather of the sizes, one grouped send/receive
 * of the payloads) and copied to `out` in job order, i.e. `cat part*.ovl` (usage, wtzmo.c:1431-1433).  parts[g] / sizes[g]: job g's text
 * in host memory; ctxs on n DISTINCT devices; one caller thread.  NCCL is loaded at run time (libnccl.so.2); no fallback. */
int zmo_gather_records(zmo_ctx **ctxs, int n, const void *const *parts, const uint64_t *sizes,
                       void *out, uint64_t out_cap, uint64_t *total, double *device_ms);
/* optional: create the NCCL communicators of the device set ahead of time (about a second; safe to call from a helper thread while the jobs run) */
int zmo_gather_prepare(zmo_ctx **ctxs, int n);
int zmo_device_count(void);

/* ---- stand-alone DP operators (unit-testable; same kernels the pipeline uses) ---------------- */
/* One problem = kswx_extend_align_core (mode 0, kswx.h:234) or kswx_extend_align_shift_core
 * (mode 1, kswx.h:101) on slices of uploaded reads.  Element k of the DP "query" is base
 * q_start + k*q_step of read q_rid, complemented if q_comp; same for the DP "target". */
typedef struct {
	uint32_t q_rid, t_rid;
	int32_t  q_start, q_step, q_comp, qlen;
	int32_t  t_start, t_step, t_comp, tlen;
	int32_t  init_score, W;                 /* W as passed to the reference (negative = exact band) */
} zmo_dp_problem_t;
/* cells: zmo_dp_global = the band width the doubling loop ended with (hzm_aln.h:1402-1417), extensions = 0; DP cell totals are in zmo_counters().
 * Slices outside their reads, steps other than +-1 and bands <= 0 are rejected with ZMO_ERR_ARG. */
typedef struct { int32_t score, qe, te, aln, mat, mis, ins, del; uint64_t cigar_off; uint32_t n_cigar; uint64_t cells; } zmo_dp_result_t;
int zmo_dp_extend(zmo_ctx *ctx, int mode, const zmo_dp_problem_t *probs, uint32_t n,
                  zmo_dp_result_t *res, uint32_t *cigars, uint64_t cigar_cap, uint64_t *cigar_needed);
/* ksw_global2 (ksw.c:503) with band w on the same slice description (init_score, W ignored; w = band). */
int zmo_dp_global(zmo_ctx *ctx, const zmo_dp_problem_t *probs, const int32_t *w, uint32_t n,
                  zmo_dp_result_t *res, uint32_t *cigars, uint64_t cigar_cap, uint64_t *cigar_needed);

#ifdef __cplusplus
}
#endif
#endif
