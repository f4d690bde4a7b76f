/*
 * zmo_align.cu -- pair alignment pipeline on the device (replaces wtzmo.c:1017-1030:
 * fast_seeds_align_hzmo per window + global_align_regs_hzmo).
 *
 *   k_window_align  one warp per (task, window): walk the window's anchors, bridge each gap with the
 *                   fixed-band extension DP, pad, run-length-align the anchor (hzm_aln.h:1247-1302)
 *   k_plan          one thread per task: region filter (wtzmo.c:1026), left end-extension job and
 *                   gap-filling jobs, scratch bump-allocated from the device arena
 *   k_ext_* / k_glb_*   persistent DP executors (zmo_dp.cu)
 *   k_plan2         accumulate left + regions + gaps, create the right end-extension job
 *   k_finish_size / k_finish   final record and stitched CIGAR (hzm_aln.h:1345-1486)
 */
#include <cub/cub.cuh>
#include "zmo_jobs.cuh"
#include "zmo_seed_core.cuh"
#include "zmo_winalign.cuh"
#include "zmo_winbridge.cuh"
#include "zmo_stitch_kernels.cuh"
#include "zmo_refine_kernels.cuh"

int zmo_launch_ext(zmo_ctx *c, int mode, int cls, const DPJob *d_jobs, const uint32_t *d_order, uint32_t n, uint32_t *arena, uint32_t *cig, DPRes *d_res, int ctr_cells);
int zmo_launch_glb(zmo_ctx *c, bool wide, const DPJob *d_jobs, const uint32_t *d_order, uint32_t n, uint32_t *arena, uint32_t *cig, DPRes *d_res, int ctr_cells);
int zmo_launch_ext_on(zmo_ctx *c, cudaStream_t st, int wk, int mode, int cls, const DPJob *d_jobs, const uint32_t *d_order, uint32_t n, uint32_t *arena, DPSlab SB, uint32_t *cig, DPRes *d_res, int ctr_cells);
int zmo_ext_grid(const zmo_ctx *c, int cls, uint32_t n, uint32_t *n_exec);
int zmo_glb_grid(const zmo_ctx *c, bool wide, uint32_t n, uint32_t *n_exec);
int zmo_launch_glb_on(zmo_ctx *c, cudaStream_t st, int wk, bool wide, const DPJob *d_jobs, const uint32_t *d_order, uint32_t n, uint32_t *arena, DPSlab SB, uint32_t *cig, DPRes *d_res, int ctr_cells);

#define CUB_CALL(c, call_expr) do { size_t _tb = 0; void *_tp = nullptr; { auto d_temp = _tp; size_t &temp_bytes = _tb; CUDA_TRY(call_expr); } \
	if((c)->cubtmp.reserve(_tb + 256)) return ZMO_ERR_CUDA; { void *d_temp = (c)->cubtmp.p; size_t &temp_bytes = _tb; CUDA_TRY(call_expr); } (c)->launches++; } while(0)

__global__ void k_job_keys(const DPJob *jobs, uint32_t n, uint32_t *keys, uint32_t *idx){
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n){ keys[i] = jobs[i].sw32; idx[i] = i; }      /* longest scratch first (~ longest job first): see DPSlab */
}

/* per-task bookkeeping shared by the plan/finish kernels */
/* ---- -n: refinement of the stitched alignment (kswx_refine_alignment, kswx.h:483-659, called at wtzmo.c:1031-1034) ----
 * k_refine_size   rows / output capacity per task
 * k_refine_band   one thread per task: half widths, per-row band [zb, ze) from the CIGAR (kswx.h:524-601), executor class
 * k_refine_warp / k_refine_cta   persistent executors: sweep KIND 3 + walk (zmo_dpr.cuh), new record and CIGAR */
/* refine every ok record of d_recs in place; the new ops land in c->s6 (the CIGAR arena is dead after k_finish), *out_words = its size */
static int refine_records(zmo_ctx *c, SeedSlot &SL, const AlnTask *d_tasks, uint32_t nt, zmo_record_t *d_recs, const uint32_t *d_ops, const AlnPar &A, unsigned long long *out_words){
	unsigned long long *ctr = c->d_ctr.as<unsigned long long>();
	/* s4 (DP results, dead): rows | row_off | outw | out_off | scr | scr_off, nt+1 each */
	if(c->s4.reserve(((size_t)nt + 1) * 6 * 8 + 64)) return ZMO_ERR_CUDA;
	unsigned long long *d_rows = c->s4.as<unsigned long long>(), *d_roff = d_rows + nt + 1, *d_outw = d_roff + nt + 1, *d_ooff = d_outw + nt + 1, *d_scr = d_ooff + nt + 1, *d_soff = d_scr + nt + 1;
	k_refine_size<<<(nt + 127) / 128, 128, 0, c->stream>>>(d_recs, nt, d_rows, d_outw); c->launches++;
	CUDA_TRY(cudaMemsetAsync(d_rows + nt, 0, 8, c->stream)); CUDA_TRY(cudaMemsetAsync(d_outw + nt, 0, 8, c->stream));
	size_t tb = 0;
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_rows, d_roff, nt + 1, c->stream));
	if(c->s2.reserve(tb + 256)) return ZMO_ERR_CUDA;              /* scan temp: not cubtmp, which holds d_recs */
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->s2.p, tb, d_rows, d_roff, nt + 1, c->stream)); c->launches++;
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->s2.p, tb, d_outw, d_ooff, nt + 1, c->stream)); c->launches++;
	unsigned long long tot[2] = {0, 0};
	CUDA_TRY(cudaMemcpyAsync(&tot[0], d_roff + nt, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(&tot[1], d_ooff + nt, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	*out_words = tot[1];
	if(c->s1.reserve((tot[0] * 3 + 16) * 4) || c->s3.reserve(((size_t)nt * 3 + 2) * sizeof(RefJob)) || c->s6.reserve((tot[1] + 16) * 4)) return ZMO_ERR_CUDA;
	CUDA_TRY(cudaMemsetAsync(ctr + CTR_N1, 0, 24, c->stream));
	k_refine_band<<<(nt + 63) / 64, 64, 0, c->stream>>>(d_recs, nt, d_ops, d_roff, d_ooff, A.w, c->s1.as<int>(), c->s3.as<RefJob>(), ctr + CTR_N1, d_scr); c->launches++;
	CUDA_TRY(cudaMemsetAsync(d_scr + nt, 0, 8, c->stream));
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->s2.p, tb, d_scr, d_soff, nt + 1, c->stream)); c->launches++;
	unsigned long long h[3] = {0, 0, 0}, stot = 0;
	CUDA_TRY(cudaMemcpyAsync(h, ctr + CTR_N1, 24, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(&stot, d_soff + nt, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if(c->arena.reserve((stot + 64) * 4)) return ZMO_ERR_CUDA;
	StageTimer tm(c, ST_GAP);
	if(h[0]){
		CUDA_TRY(cudaMemsetAsync(ctr + CTR_WORK, 0, 8, c->stream));
		const int grid = (int)std::min<uint64_t>((h[0] + 3) / 4, (uint64_t)c->n_sm * 4);
		k_refine_warp<<<grid, 128, 0, c->stream>>>(c->s3.as<RefJob>(), (uint32_t)h[0], d_soff, SL.pairs.as<zmo_pair_t>(), d_tasks, dev_reads(c), A.P, c->s1.as<int>(), c->arena.as<uint32_t>(), c->s6.as<uint32_t>(), d_recs, ctr, CTR_WORK, CTR_CELLS_GAP); c->launches++;
	}
	if(h[1]){
		CUDA_TRY(cudaMemsetAsync(ctr + CTR_WORKK, 0, 8, c->stream));
		const int grid = (int)std::min<uint64_t>(h[1], (uint64_t)c->n_sm * 3);
		k_refine_cta<<<grid, CL3_NT, 0, c->stream>>>(c->s3.as<RefJob>() + nt, (uint32_t)h[1], d_soff, SL.pairs.as<zmo_pair_t>(), d_tasks, dev_reads(c), A.P, c->s1.as<int>(), c->arena.as<uint32_t>(), c->s6.as<uint32_t>(), d_recs, ctr, CTR_WORKK, CTR_CELLS_GAP); c->launches++;
	}
	if(h[2]){
		/* bands beyond the register executors: chunked sweep, one CTA per job (rare: an indel run of several hundred bases) */
		CUDA_TRY(cudaMemsetAsync(ctr + CTR_WORKK + 1, 0, 8, c->stream));
		const int grid = (int)std::min<uint64_t>(h[2], (uint64_t)c->n_sm * 2);
		k_refine_wide<<<grid, REFW_NT, 0, c->stream>>>(c->s3.as<RefJob>() + 2 * (size_t)nt, (uint32_t)h[2], d_soff, SL.pairs.as<zmo_pair_t>(), d_tasks, dev_reads(c), A.P, c->s1.as<int>(), c->arena.as<uint32_t>(), c->s6.as<uint32_t>(), d_recs, ctr, CTR_WORKK + 1, CTR_CELLS_GAP); c->launches++;
	}
	CUDA_TRY(cudaGetLastError());
	return 0;
}

/* ---- CIGAR text on the device (kswx_cigar2string, kswx.h:1093-1120: "%d%c" per op with len > 0, ops M/I/D) ---- */
__device__ __forceinline__ uint32_t cig_op_chars(uint32_t op){
	const uint32_t len = op >> 4;
	if(len == 0) return 0;
	return 2u + (len >= 10u) + (len >= 100u) + (len >= 1000u) + (len >= 10000u) + (len >= 100000u) + (len >= 1000000u) + (len >= 10000000u) + (len >= 100000000u);
}
/* one warp per task: text length of its stitched CIGAR */
__global__ void k_cig_textlen(const zmo_record_t *recs, uint32_t nt, const uint32_t *ops, unsigned long long *tlen){
	const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if(t >= nt) return;
	const zmo_record_t r = recs[t];
	unsigned long long n = 0;
	if(r.ok){ const uint32_t *o = ops + r.cigar_off; for(uint32_t k = lane; k < r.n_cigar; k += 32) n += cig_op_chars(o[k]); }
	#pragma unroll
	for(int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
	if(lane == 0) tlen[t] = n;
}
/* one warp per task: write the text, re-point the record at it (cigar_off / n_cigar become byte offset / byte length) */
__global__ void k_cig_text(zmo_record_t *recs, uint32_t nt, const uint32_t *ops, const unsigned long long *toff, char *text){
	const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if(t >= nt) return;
	const zmo_record_t r = recs[t];
	const unsigned long long t0 = toff[t], t1 = toff[t + 1];
	if(r.ok){
		const uint32_t *o = ops + r.cigar_off; char *dst = text + t0; unsigned long long cur = 0;
		for(uint32_t base = 0; base < r.n_cigar; base += 32){
			const uint32_t k = base + lane; const uint32_t op = k < r.n_cigar? o[k] : 0u; const uint32_t nc = cig_op_chars(op);
			uint32_t incl = nc;
			#pragma unroll
			for(int d = 1; d < 32; d <<= 1){ const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if((int)lane >= d) incl += v; }
			if(nc){
				char *p = dst + cur + incl - 1;      /* last char of this op */
				*p-- = "MIDX"[op & 3u];
				uint32_t len = op >> 4;
				do { *p-- = (char)('0' + len % 10u); len /= 10u; } while(len);
			}
			cur += __shfl_sync(0xffffffffu, incl, 31);
		}
	}
	if(lane == 0){ recs[t].cigar_off = t0; recs[t].n_cigar = (uint32_t)(t1 - t0); }
}

/* Run the job lists [first[k], n[k]) of all six executor classes CONCURRENTLY (one auxiliary stream per class, forked from
 * and joined back to the context stream), each work queue ordered longest-job-first to cut the tail. */
/* One DP phase: the six job lists run concurrently, each on its own stream.  Every list is ordered by scratch size
 * (descending) and its executors work in private slabs: executor e starts with job e, so its slab is as large as job e's
 * scratch and the arena holds the prefix sum over the first #executors jobs of every class (DPSlab) -- not the scratch
 * of every job of the wave.  Nothing in the arena is live between phases, so it may grow here. */
struct SlabWords { const uint32_t *k; uint32_t m; __host__ __device__ unsigned long long operator()(uint32_t i) const { return i < m? (unsigned long long)k[i] << 5 : 0ull; } };
static int run_dp_lists(zmo_ctx *c, const JobLists &L, const uint32_t *n, const uint32_t *first, uint32_t *cig_arena, DPRes *d_res){
	uint32_t tot = 0, cnt[6], off[6], nex[6]; size_t nexs = 0;
	for(int k = 0; k < 6; k++){
		off[k] = first? first[k] : 0; cnt[k] = n[k] - off[k]; tot += cnt[k]; nex[k] = 0;
		if(cnt[k]){ if(k < 4) zmo_ext_grid(c, k, cnt[k], &nex[k]); else zmo_glb_grid(c, k == 5, cnt[k], &nex[k]); }
		nexs += nex[k] + 2;
	}
	if(tot == 0) return 0;
	DevBuf &ob = c->s5;      /* s5 is free until k_finish: keys | idx | sorted keys | sorted idx per class, then the slab offsets */
	if(ob.reserve((size_t)tot * 16 + nexs * 8 + 512)) return ZMO_ERR_CUDA;
	uint32_t *base = ob.as<uint32_t>(); uint32_t *order[6]; size_t pos = 0;
	unsigned long long *soff[6], *sbase = (unsigned long long*)(((uintptr_t)(base + (size_t)tot * 4) + 63) & ~(uintptr_t)63), htot[6] = {0, 0, 0, 0, 0, 0};
	for(int k = 0; k < 6; k++){
		order[k] = nullptr; soff[k] = sbase; sbase += nex[k] + 2;
		if(cnt[k] == 0) continue;
		uint32_t *keys = base + pos, *idx = keys + cnt[k], *skeys = idx + cnt[k], *sidx = skeys + cnt[k]; pos += (size_t)cnt[k] * 4;
		k_job_keys<<<(cnt[k] + 255) / 256, 256, 0, c->stream>>>(L.list[k] + off[k], cnt[k], keys, idx); c->launches++;
		if(cnt[k] >= 2){
			CUB_CALL(c, cub::DeviceRadixSort::SortPairsDescending(d_temp, temp_bytes, keys, skeys, idx, sidx, (int)cnt[k], 0, 32, c->stream));
			order[k] = sidx;
		} else skeys = keys;
		SlabWords f; f.k = skeys; f.m = std::min(nex[k], cnt[k]);
		cub::TransformInputIterator<unsigned long long, SlabWords, cub::CountingInputIterator<uint32_t>> it(cub::CountingInputIterator<uint32_t>(0), f);
		CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, it, soff[k], (int)nex[k] + 1, c->stream));
		CUDA_TRY(cudaMemcpyAsync(&htot[k], soff[k] + nex[k], 8, cudaMemcpyDeviceToHost, c->stream));
	}
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	DPSlab SB[6]; unsigned long long need = 0;
	for(int k = 0; k < 6; k++){ SB[k].base = need; SB[k].off = soff[k]; need += htot[k] + 32; }
	if(c->arena.reserve((need + 64) * 4)) return ZMO_ERR_CUDA;
	uint32_t *arena = c->arena.as<uint32_t>();
	CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
	CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
	for(int k = 0; k < 6; k++){
		if(cnt[k] == 0) continue;
		CUDA_TRY(cudaStreamWaitEvent(c->aux[k], c->ev_fork, 0));
		CUDA_TRY(cudaEventRecord(c->ev_a0[k], c->aux[k]));
		int rc;
		if(k < 4) rc = zmo_launch_ext_on(c, c->aux[k], CTR_WORKK + k, 1, k, L.list[k] + off[k], order[k], cnt[k], arena, SB[k], cig_arena, d_res, CTR_CELLS_EXT);
		else rc = zmo_launch_glb_on(c, c->aux[k], CTR_WORKK + k, k == 5, L.list[k] + off[k], order[k], cnt[k], arena, SB[k], cig_arena, d_res, CTR_CELLS_GAP);
		if(rc) return rc;
		CUDA_TRY(cudaEventRecord(c->ev_a1[k], c->aux[k]));
		CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_a1[k], 0));
	}
	CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	{ float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); c->stage_ms[ST_DPWALL] += ms; }
	for(int k = 0; k < 6; k++){
		if(cnt[k] == 0) continue;
		float ms = 0; cudaEventElapsedTime(&ms, c->ev_a0[k], c->ev_a1[k]);
		c->stage_ms[k < 4? ST_EXT : ST_GAP] += ms;
	}
	return 0;
}

/* res index -> position in the concatenated job array [ext_w | ext_n | glb_w | glb_n] is the identity by construction */
static int pair_align_impl(zmo_ctx *c, int slot, const zmo_task_t *tasks, uint32_t nt, zmo_record_t *recs, uint32_t *cigars, uint64_t cigar_cap, uint64_t *cigar_needed, int out_mode);
extern "C" int zmo_pair_align(zmo_ctx *c, int slot, const zmo_task_t *tasks, uint32_t nt, zmo_record_t *recs, uint32_t *cigars, uint64_t cigar_cap, uint64_t *cigar_needed){
	return pair_align_impl(c, slot, tasks, nt, recs, cigars, cigar_cap, cigar_needed, 0);
}
extern "C" int zmo_pair_align_text(zmo_ctx *c, int slot, const zmo_task_t *tasks, uint32_t nt, zmo_record_t *recs, char *cigar_text, uint64_t text_cap, uint64_t *text_needed){
	return pair_align_impl(c, slot, tasks, nt, recs, (uint32_t*)cigar_text, text_cap, text_needed, 1);
}
/* records only: what a consumer that drops the CIGAR column needs (`wtzmo ... | cut -f1-16`, smartdenovo.pl:58); no CIGAR text is formatted or copied */
extern "C" int zmo_pair_align_records(zmo_ctx *c, int slot, const zmo_task_t *tasks, uint32_t nt, zmo_record_t *recs){
	return pair_align_impl(c, slot, tasks, nt, recs, nullptr, 0, nullptr, 2);
}
/* out_mode 0: binary ops, 1: text, 2: records only; cigar_cap / *cigar_needed count ops (binary) or bytes (text) */
static int pair_align_impl(zmo_ctx *c, int slot, const zmo_task_t *tasks, uint32_t nt, zmo_record_t *recs, uint32_t *cigars, uint64_t cigar_cap, uint64_t *cigar_needed, int out_mode){
	const bool as_text = out_mode == 1, no_cigar = out_mode == 2;
	if(!c || (nt && (!tasks || !recs))) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	if(slot < 0 || slot > 1) return zmo_set_err(ZMO_ERR_ARG, "slot must be 0 or 1");
	if(cigar_needed) *cigar_needed = 0;
	if(nt == 0) return 0;
	SeedSlot &SL = c->slot[slot];
	if(SL.np == 0) return zmo_set_err(ZMO_ERR_STATE, "zmo_pair_windows has not filled slot %d", slot);
	CUDA_TRY(cudaSetDevice(c->device));
	/* host: items and per-item cigar regions */
	std::vector<AlnTask> ht(nt); std::vector<WItem> items; std::vector<unsigned long long> icig, istep, ibound;
	unsigned long long cig_words = 0, nsteps64 = 0; int max_rows = 16;
	for(uint32_t t = 0; t < nt; t++){
		if(tasks[t].pair_idx >= SL.np || tasks[t].dir > 1) return zmo_set_err(ZMO_ERR_ARG, "task %u out of range", t);
		const zmo_pairseed_t &ps = SL.h_seeds[tasks[t].pair_idx]; const uint32_t d = tasks[t].dir;
		ht[t].pair_idx = tasks[t].pair_idx; ht[t].dir = d; ht[t].item_off = (uint32_t)items.size(); ht[t].n_item = ps.n_win[d];
		for(uint32_t k = 0; k < ps.n_win[d]; k++){
			WItem it; it.task = t; it.win = ps.win_off[d] + k; items.push_back(it);
			const int s0 = SL.h_wspan[3 * (size_t)it.win], s1 = SL.h_wspan[3 * (size_t)it.win + 1], na = SL.h_wspan[3 * (size_t)it.win + 2];
			icig.push_back(cig_words); cig_words += (unsigned long long)(s0 + s1 + 16 + 2 * na);
			istep.push_back(nsteps64); nsteps64 += (unsigned long long)na; ibound.push_back((unsigned long long)(s1 + 16));      /* rows of all bridges of a window <= its span on c */
			if(s1 + 8 > max_rows) max_rows = s1 + 8;
			if(s0 + 8 > max_rows) max_rows = s0 + 8;
		}
	}
	const uint32_t nitems = (uint32_t)items.size();
	AlnPar A; A.w = c->par.w; A.ew = c->par.ew; A.W = c->par.W; A.zovl = c->par.zovl; A.min_id = c->par.min_id;
	A.P.M = c->par.M; A.P.X = c->par.X; A.P.I = c->par.O; A.P.D = c->par.O; A.P.E = c->par.E; A.P.T = c->par.T;
	DevReads R = dev_reads(c);
	unsigned long long *ctr = c->d_ctr.as<unsigned long long>();
	/* window-align executors and their slabs */
	static const bool wb_env = [](){ const char *e = getenv("ZMO_WA_BRIDGE"); return !(e && e[0] == '0'); }();      /* ZMO_WA_BRIDGE=0: every window through k_window_align (A/B runs) */
	const bool use_wb = wb_env && c->par.w >= 1 && c->par.w <= WB_MAX_W && nitems > 0 && nsteps64 < 0xFFFFFFF0ull;
	const int wb_ring = wb_cap(c->par.w), wb_rw = wb_row_words(c->par.w), wb_acap = 2 * (int)c->par.zsize + 2;      /* an anchor is a z-mer: zsize runs, at most M + I|D each */
	/* scratch of the swept bridges, bounded from the window spans (wb_scr_words per bridge); a wave whose bound exceeds the budget is swept in several
	 * passes over item ranges, so the arena stays bounded whatever the wave size (ZMO_WB_CHUNK_MB, default 6144) */
	static const unsigned long long wb_budget = [](){ const char *e = getenv("ZMO_WB_CHUNK_MB"); const long long mb = e? atoll(e) : 6144; return (unsigned long long)(mb < 1? 1 : mb) * (1ull << 18); }();     /* words */
	std::vector<uint32_t> wb_chunk;      /* first item of every pass, then nitems */
	unsigned long long wb_scr_cap = 0;
	if(use_wb){
		unsigned long long acc = 0; wb_chunk.push_back(0);
		for(uint32_t i = 0; i < nitems; i++){
			const unsigned long long na_i = (i + 1 < nitems? istep[i + 1] : nsteps64) - istep[i];
			const unsigned long long b = ibound[i] * (unsigned long long)(wb_rw + 5) + ibound[i] + na_i * (unsigned long long)(c->par.w + 24) + 64;      /* sum of wb_scr_words: rows <= span on c, columns of a bridge <= its rows + w */
			if(acc && acc + b > wb_budget){ wb_chunk.push_back(i); if(acc > wb_scr_cap) wb_scr_cap = acc; acc = 0; }
			acc += b;
		}
		wb_chunk.push_back(nitems); if(acc > wb_scr_cap) wb_scr_cap = acc;
		wb_scr_cap += 1024;
	}
	/* with the bridge pipeline k_window_align only sees the windows that pipeline leaves out: a quarter of the executors (all of them are used if needed, just in more rounds) */
	const int wgrid = (int)std::min<uint64_t>((nitems + WA_WARPS - 1) / WA_WARPS + 1, (uint64_t)c->n_sm * (use_wb? 2 : 8));
	const int wcol = std::min(max_rows + c->par.w, 2 * c->par.w + 1);
	unsigned long long slab = (unsigned long long)max_rows * band_row_words<32, WA_C>(wcol) + max_rows + (2ull * max_rows + 2ull * c->par.w + 16) + ((unsigned long long)max_rows >> 3) + (c->par.w >> 3) + 8;
	if(2 * c->par.w + 3 > WA_CAP){ unsigned long long cap = 1; while(cap < (unsigned long long)(2 * c->par.w + 3)) cap <<= 1; slab += 3 * cap; }
	slab = (slab + 63) & ~63ull;
	const unsigned long long slabs_total = slab * (unsigned long long)wgrid * WA_WARPS;
	const uint32_t jcap = nitems + 2 * nt + 8;
	/* device buffers: s0 tasks|items|icig, s1 regs, s2 task state, s3 jobs (4 lists), s4 results, s6 cig arena, s7 out offsets */
	if(c->s0.reserve((size_t)nt * sizeof(AlnTask) + (size_t)nitems * (sizeof(WItem) + 8) + 64) || c->s1.reserve(((size_t)nitems + 1) * sizeof(DevReg)) || c->s2.reserve(((size_t)nt + 1) * sizeof(TaskState))
		|| c->s3.reserve((size_t)jcap * 6 * sizeof(DPJob)) || c->s4.reserve((size_t)jcap * 6 * sizeof(DPRes)) || c->s7.reserve(((size_t)nt + 2) * 16)) return ZMO_ERR_CUDA;
	AlnTask *d_tasks = c->s0.as<AlnTask>(); WItem *d_items = (WItem*)(d_tasks + nt); unsigned long long *d_icig = (unsigned long long*)(((uintptr_t)(d_items + nitems) + 7) & ~(uintptr_t)7);
	DevReg *d_regs = c->s1.as<DevReg>(); TaskState *d_ts = c->s2.as<TaskState>();
	DPJob *d_jobs = c->s3.as<DPJob>(); DPRes *d_res = c->s4.as<DPRes>();
	CUDA_TRY(cudaMemcpyAsync(d_tasks, ht.data(), (size_t)nt * sizeof(AlnTask), cudaMemcpyHostToDevice, c->stream));
	if(nitems){
		CUDA_TRY(cudaMemcpyAsync(d_items, items.data(), (size_t)nitems * sizeof(WItem), cudaMemcpyHostToDevice, c->stream));
		CUDA_TRY(cudaMemcpyAsync(d_icig, icig.data(), (size_t)nitems * 8, cudaMemcpyHostToDevice, c->stream));
	}
	c->counters[5] += (size_t)nt * sizeof(AlnTask) + (size_t)nitems * 16;
	unsigned long long cig_cap_words = cig_words + (unsigned long long)nt * 4096 + (1ull << 20);
	for(int attempt = 0; ; attempt++){
		if(c->arena.reserve((std::max(slabs_total, wb_scr_cap) + 64) * 4) || c->s6.reserve(cig_cap_words * 4)) return ZMO_ERR_CUDA;
		if(use_wb && (c->wb0.reserve((size_t)(nsteps64 + 1) * (sizeof(WBStep) + 4 * (size_t)(wb_acap + 1))) || c->wb1.reserve(((size_t)nitems + 1) * 8 + ((size_t)nsteps64 + 2) * 16 + 16 + (size_t)nsteps64 * 16 + (size_t)nitems * 5 + 64))) return ZMO_ERR_CUDA;
		cig_cap_words = c->s6.cap / 4;
		uint32_t *arena = c->arena.as<uint32_t>(), *cig_arena = c->s6.as<uint32_t>();
		if(nitems && use_wb){
			/* bridge-level pipeline (zmo_winbridge.cuh); the windows it leaves out go through k_window_align below */
			StageTimer tm(c, ST_WINALN);
			const uint32_t nsteps = (uint32_t)nsteps64;
			WBStep *d_steps = c->wb0.as<WBStep>(); uint32_t *d_aops = (uint32_t*)(d_steps + nsteps + 1);
			unsigned long long *d_istep = c->wb1.as<unsigned long long>(), *d_scrw = d_istep + nitems + 1, *d_scro = d_scrw + nsteps + 1;
			uint32_t *d_keys = (uint32_t*)(d_scro + nsteps + 2), *d_ord = d_keys + nsteps, *d_skeys = d_ord + nsteps, *d_sord = d_skeys + nsteps, *d_fb = d_sord + nsteps;
			uint8_t *d_iseq = (uint8_t*)(d_fb + nitems);
			CUDA_TRY(cudaMemcpyAsync(d_istep, istep.data(), (size_t)nitems * 8, cudaMemcpyHostToDevice, c->stream));
			CUDA_TRY(cudaMemsetAsync(ctr + CTR_N4, 0, 16, c->stream));            /* CTR_N4 = windows left to k_window_align, CTR_N5 = scratch bound violated */
			int per_sm = 0;      /* resident CTAs per SM with this ring: one wave of persistent CTAs, longest bridges first */
			CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wb_sweep, WB_NT, (size_t)wb_ring * 4 * WB_NT));
			for(size_t ch = 0; ch + 1 < wb_chunk.size(); ch++){
				/* one pass: items [i0, i1), their steps [st0, st1); every per-item / per-step array is addressed from its global base, the per-pass views below only
				 * shift the item-indexed ones (k_wb_prep lists the windows it leaves out by their index inside the pass) */
				const uint32_t i0 = wb_chunk[ch], i1 = wb_chunk[ch + 1], ni = i1 - i0;
				const unsigned long long st0 = istep[i0], st1 = i1 < nitems? istep[i1] : nsteps64; const uint32_t ns = (uint32_t)(st1 - st0);
				if(ch) CUDA_TRY(cudaMemsetAsync(ctr + CTR_N4, 0, 8, c->stream));
				CUDA_TRY(cudaMemsetAsync(ctr + CTR_WORK, 0, 8, c->stream));
				k_wb_prep<<<(unsigned)(((unsigned long long)ni * 32 + 127) / 128), 128, 0, c->stream>>>(d_items + i0, ni, d_tasks, SL.pairs.as<zmo_pair_t>(), SL.wins.as<DevWin>(), SL.anchors.as<DevZPair>(), R, A,
					d_istep + i0, wb_rw, d_steps, d_aops, wb_acap, d_scrw, d_keys, d_ord, d_iseq + i0, d_fb, ctr + CTR_N4); c->launches++;
				if(ns){
					CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_scrw + st0, d_scro + st0, (int)ns + 1, c->stream));      /* offsets from 0 in every pass; the input behind the last step is not used */
					const uint32_t *sk = d_keys + st0, *so = d_ord + st0;
					if(ns >= 2){ CUB_CALL(c, cub::DeviceRadixSort::SortPairsDescending(d_temp, temp_bytes, d_keys + st0, d_skeys + st0, d_ord + st0, d_sord + st0, (int)ns, 0, 32, c->stream)); sk = d_skeys + st0; so = d_sord + st0; }
					const int sgrid = (int)std::min<uint64_t>(((uint64_t)ns + WB_NT - 1) / WB_NT, (uint64_t)c->n_sm * (uint64_t)std::max(per_sm, 1));
					k_wb_sweep<<<sgrid, WB_NT, (size_t)wb_ring * 4 * WB_NT, c->stream>>>(d_steps, so, sk, ns, d_scro, wb_scr_cap, R.words, A.P, arena, wb_ring, wb_rw, ctr + CTR_WORK, ctr + CTR_N5); c->launches++;
					k_wb_ends<<<(ni + 63) / 64, 64, 0, c->stream>>>(ni, d_items + i0, SL.wins.as<DevWin>(), A, d_istep + i0, d_iseq + i0, d_steps, d_scro, arena, wb_rw, ctr + CTR_N5, ctr, CTR_CELLS_WIN); c->launches++;
					k_wb_walk<<<(ns + 127) / 128, 128, 0, c->stream>>>(d_steps, so, sk, ns, d_scro, R.words, A.P, arena, wb_rw, ctr + CTR_N5); c->launches++;
				}
				k_wb_stitch<<<(unsigned)(((unsigned long long)ni * 32 + 127) / 128), 128, 0, c->stream>>>(d_items + i0, ni, SL.wins.as<DevWin>(), A, d_istep + i0, d_iseq + i0, d_steps, d_aops, wb_acap, d_scro, arena, wb_rw, ctr + CTR_N5, cig_arena, d_icig + i0, d_regs + i0); c->launches++;
				CUDA_TRY(cudaMemsetAsync(ctr + CTR_WORK, 0, 8, c->stream));
				k_window_align<<<wgrid, 32 * WA_WARPS, 0, c->stream>>>(d_items + i0, ni, d_tasks, SL.pairs.as<zmo_pair_t>(), SL.wins.as<DevWin>(), SL.anchors.as<DevZPair>(), R, A,
					arena, slab, max_rows, cig_arena, d_icig + i0, d_regs + i0, ctr, CTR_WORK, CTR_CELLS_WIN, d_fb, ctr + CTR_N4);
				c->launches++;
			}
			CUDA_TRY(cudaGetLastError());
		} else if(nitems){
			StageTimer tm(c, ST_WINALN);
			CUDA_TRY(cudaMemsetAsync(ctr + CTR_WORK, 0, 8, c->stream));
			CUDA_TRY(cudaMemsetAsync(ctr + CTR_N4, 0, 16, c->stream));
			k_window_align<<<wgrid, 32 * WA_WARPS, 0, c->stream>>>(d_items, nitems, d_tasks, SL.pairs.as<zmo_pair_t>(), SL.wins.as<DevWin>(), SL.anchors.as<DevZPair>(), R, A,
				arena, slab, max_rows, cig_arena, d_icig, d_regs, ctr, CTR_WORK, CTR_CELLS_WIN, nullptr, nullptr);
			c->launches++;
			CUDA_TRY(cudaGetLastError());
		}
		/* plan: left extensions + gaps */
		JobLists L; L.cap = jcap;
		for(int k = 0; k < 6; k++){ L.list[k] = d_jobs + (size_t)k * jcap; L.cnt[k] = ctr + CTR_JOBS + k; L.res_base[k] = (uint32_t)k * jcap; }
		L.cig_cur = ctr + CTR_CIG; L.cig_cap = cig_cap_words; L.overflow = ctr + CTR_OVERFLOW;
		unsigned long long init_ctr[2] = {0, cig_words};
		CUDA_TRY(cudaMemsetAsync(ctr + CTR_JOBS, 0, 6 * 8, c->stream));
		CUDA_TRY(cudaMemsetAsync(ctr + CTR_OVERFLOW, 0, 8, c->stream));
		CUDA_TRY(cudaMemcpyAsync(ctr + CTR_CIG, &init_ctr[1], 8, cudaMemcpyHostToDevice, c->stream));
		k_plan<<<(nt + 63) / 64, 64, 0, c->stream>>>(d_tasks, nt, SL.pairs.as<zmo_pair_t>(), d_regs, R, A, L, d_ts); c->launches++;
		unsigned long long h[CTR_TOTAL];
		CUDA_TRY(cudaMemcpyAsync(h, ctr, CTR_TOTAL * 8, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		if(use_wb && h[CTR_N5]) return zmo_set_err(ZMO_ERR_STATE, "window alignment: bridge scratch bound violated (%llu bridges)", h[CTR_N5]);
		c->counters[7] += use_wb? h[CTR_N4] : 0;      /* windows that took the sequential path */
		if(h[CTR_OVERFLOW]){
			if(attempt >= 6) return zmo_set_err(ZMO_ERR_CAPACITY, "cigar arena overflow after %d attempts", attempt);
			cig_cap_words = std::max(cig_cap_words * 2, h[CTR_CIG] + (1ull << 20));
			continue;
		}
		uint32_t n1[6]; for(int k = 0; k < 6; k++) n1[k] = (uint32_t)h[CTR_JOBS + k];
		if(int rc = run_dp_lists(c, L, n1, nullptr, cig_arena, d_res)) return rc;      /* the window slabs are dead by now: the DP slabs reuse the arena from 0 */
		/* plan2: right extensions appended to the same extension lists */
		k_plan2<<<(nt + 63) / 64, 64, 0, c->stream>>>(d_tasks, nt, SL.pairs.as<zmo_pair_t>(), d_regs, d_res, R, A, L, d_ts); c->launches++;
		CUDA_TRY(cudaMemcpyAsync(h, ctr, CTR_TOTAL * 8, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		if(h[CTR_OVERFLOW]){
			if(attempt >= 6) return zmo_set_err(ZMO_ERR_CAPACITY, "cigar arena overflow after %d attempts", attempt);
			cig_cap_words = std::max(cig_cap_words * 2, h[CTR_CIG] + (1ull << 20));
			continue;
		}
		{
			uint32_t n2[6]; for(int k = 0; k < 6; k++) n2[k] = (uint32_t)h[CTR_JOBS + k];
			n2[4] = n1[4]; n2[5] = n1[5];      /* no new gap jobs in the second phase */
			if(int rc = run_dp_lists(c, L, n2, n1, cig_arena, d_res)) return rc;
		}
		/* final sizes, offsets, stitched CIGARs */
		unsigned long long *d_need = c->s7.as<unsigned long long>(), *d_ooff = d_need + nt + 1;
		k_finish_size<<<(nt + 127) / 128, 128, 0, c->stream>>>(nt, d_res, d_ts, d_need); c->launches++;
		CUDA_TRY(cudaMemsetAsync(d_need + nt, 0, 8, c->stream));
		CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_need, d_ooff, nt + 1, c->stream));
		unsigned long long total = 0;
		CUDA_TRY(cudaMemcpyAsync(&total, d_ooff + nt, 8, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		if(!as_text && !no_cigar){
			if(cigar_needed) *cigar_needed = total;
			if(total > cigar_cap) return zmo_set_err(ZMO_ERR_CAPACITY, "cigar buffer too small: need %llu", total);
		}
		if(c->s5.reserve((total + 16) * 4) || c->cubtmp.reserve((size_t)nt * sizeof(zmo_record_t) + 256)) return ZMO_ERR_CUDA;
		/* a job's index in [ext_w | ext_n | glb_w | glb_n] equals its result index by construction */
		zmo_record_t *d_recs = c->cubtmp.as<zmo_record_t>();
		{
			k_finish_warp<<<(unsigned)(((unsigned long long)nt * 32 + 255) / 256), 256, 0, c->stream>>>(d_tasks, nt, d_regs, d_res, d_jobs, cig_arena, A, d_ts, d_ooff, c->s5.as<uint32_t>(), d_recs);
			c->launches++;
		}
		CUDA_TRY(cudaGetLastError());
		const uint32_t *d_final_ops = c->s5.as<uint32_t>();
		if(c->refine){
			unsigned long long ow = 0;
			if(int rc = refine_records(c, SL, d_tasks, nt, d_recs, c->s5.as<uint32_t>(), A, &ow)) return rc;
			d_final_ops = c->s6.as<uint32_t>(); total = ow;
			if(!as_text && !no_cigar){
				if(cigar_needed) *cigar_needed = total;
				if(total > cigar_cap) return zmo_set_err(ZMO_ERR_CAPACITY, "cigar buffer too small: need %llu", total);
			}
		}
		if(as_text){
			/* text lengths -> offsets (d_need / d_ooff are free again) -> text in s3 (the job lists are dead after k_finish) */
			k_cig_textlen<<<(nt + 3) / 4, 128, 0, c->stream>>>(d_recs, nt, d_final_ops, d_need); c->launches++;
			CUDA_TRY(cudaMemsetAsync(d_need + nt, 0, 8, c->stream));
			{
				size_t tb = 0; unsigned long long *d_toff = d_ooff;
				CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_need, d_toff, nt + 1, c->stream));
				if(c->s2.reserve(tb + 256)) return ZMO_ERR_CUDA;                      /* scan temp must not alias d_recs (cubtmp); the task state in s2 is dead after k_finish */
				CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->s2.p, tb, d_need, d_toff, nt + 1, c->stream)); c->launches++;
			}
			unsigned long long tbytes = 0;
			CUDA_TRY(cudaMemcpyAsync(&tbytes, d_ooff + nt, 8, cudaMemcpyDeviceToHost, c->stream));
			CUDA_TRY(cudaStreamSynchronize(c->stream));
			if(cigar_needed) *cigar_needed = tbytes;
			if(tbytes > cigar_cap) return zmo_set_err(ZMO_ERR_CAPACITY, "cigar text buffer too small: need %llu bytes", tbytes);
			if(c->s3.reserve(tbytes + 64)) return ZMO_ERR_CUDA;
			k_cig_text<<<(nt + 3) / 4, 128, 0, c->stream>>>(d_recs, nt, d_final_ops, d_ooff, c->s3.as<char>()); c->launches++;
			CUDA_TRY(cudaGetLastError());
			{
				StageTimer tm(c, ST_COPY);
				CUDA_TRY(cudaMemcpyAsync(recs, d_recs, (size_t)nt * sizeof(zmo_record_t), cudaMemcpyDeviceToHost, c->stream));
				if(tbytes) CUDA_TRY(cudaMemcpyAsync(cigars, c->s3.p, tbytes, cudaMemcpyDeviceToHost, c->stream));
			}
			CUDA_TRY(cudaStreamSynchronize(c->stream));
			c->counters[6] += (size_t)nt * sizeof(zmo_record_t) + tbytes;
			return 0;
		}
		if(no_cigar) total = 0;
		{
			StageTimer tm(c, ST_COPY);
			CUDA_TRY(cudaMemcpyAsync(recs, d_recs, (size_t)nt * sizeof(zmo_record_t), cudaMemcpyDeviceToHost, c->stream));
			if(total) CUDA_TRY(cudaMemcpyAsync(cigars, d_final_ops, total * 4, cudaMemcpyDeviceToHost, c->stream));
		}
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		if(no_cigar) for(uint32_t t = 0; t < nt; t++){ recs[t].cigar_off = 0; recs[t].n_cigar = 0; }
		c->counters[6] += (size_t)nt * sizeof(zmo_record_t) + total * 4;
		return 0;
	}
}
