/*
 * zmo_dp_kernels.cuh -- the persistent DP executor kernels (extension classes 0-3, gap filling).  Kept in a header so that
 * the test-only host simulation (tests/hostsim/dp_host.cpp) runs this very source; included by zmo_dp.cu only.
 */
#pragma once
#include "zmo_jobs.cuh"

#define EXT_NT 256
#define EXT_C  7
#define EXT_CAP 2048
#define EXT_SEQW 4096
/* executor class of an extension band: host and device must agree (scratch sizing depends on it) */
#define WRP_C  7
#define WRP_CAP 256
#define WRP_SEQW 256
#define WRP_PER_CTA 4

/* CTA-per-job extension kernels (register-resident sweep): NT threads x C columns per block; classes 1/2/3 =
 * 64x7 / 128x7 / CL3_NT x CL3_C serve bands up to 435 / 883 / 1639 columns.  The class-3 kernel also takes the
 * bands beyond that (chunked sweep with rows in global memory). */
template<int NT, int C, int MODE>
__global__ void __launch_bounds__(NT, (NT == 256? 2 : (NT == 128? (C > 7? 3 : 5) : 8))) k_ext_cta(const DPJob *jobs, const uint32_t *order, uint32_t njobs, DevReads R, DPPar P,
		uint32_t *arena, DPSlab SB, uint32_t *cig_arena, DPRes *res, unsigned long long *ctr, int ctr_work, int ctr_cells){
	constexpr int SEQW = (NT == 256 || C > 7)? 4096 : 2048;
	__shared__ uint32_t s_seq[SEQW];
	__shared__ int s_red[2 * (NT / 32)];
	__shared__ long long s_redk[NT / 32];
	__shared__ int s_misc[16];
	__shared__ uint32_t s_job;
	ExecSmem<NT> X; X.carve(nullptr, 0, s_seq, SEQW, s_red, s_redk, s_misc);
	const int tid = threadIdx.x;
	uint32_t *slab = SB.off? arena + SB.base + SB.off[blockIdx.x] : nullptr;
	for(bool first = true; ; first = false){
		if(tid == 0) s_job = first? blockIdx.x : gridDim.x + (uint32_t)atomicAdd(ctr + ctr_work, 1ULL);
		__syncthreads();
		const uint32_t jn = s_job;
		__syncthreads();
		if(jn >= njobs) break;
		run_ext_job<NT, C, MODE>(jobs[order? order[jn] : jn], R, P, X, arena, slab, cig_arena, res, ctr + ctr_cells, tid);
		__syncthreads();
	}
}

/* warp-per-job extension kernel (bands up to 211 columns) */
template<int MODE>
__global__ void __launch_bounds__(32 * WRP_PER_CTA) k_ext_warp(const DPJob *jobs, const uint32_t *order, uint32_t njobs, DevReads R, DPPar P,
		uint32_t *arena, DPSlab SB, uint32_t *cig_arena, DPRes *res, unsigned long long *ctr, int ctr_work, int ctr_cells){
	__shared__ uint32_t s_seq[WRP_PER_CTA][WRP_SEQW];
	__shared__ int s_misc[WRP_PER_CTA][16];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	ExecSmem<32> X; X.carve(nullptr, 0, s_seq[warp], WRP_SEQW, nullptr, nullptr, s_misc[warp]);
	const uint32_t ex = blockIdx.x * WRP_PER_CTA + warp;
	uint32_t *slab = SB.off? arena + SB.base + SB.off[ex] : nullptr;
	for(bool first = true; ; first = false){
		uint32_t jn = ex;
		if(!first){ if(lane == 0) jn = gridDim.x * WRP_PER_CTA + (uint32_t)atomicAdd(ctr + ctr_work, 1ULL); jn = __shfl_sync(0xffffffffu, jn, 0); }
		if(jn >= njobs) break;
		run_ext_job<32, WRP_C, MODE>(jobs[order? order[jn] : jn], R, P, X, arena, slab, cig_arena, res, ctr + ctr_cells, lane);
		__syncwarp();
	}
}

__global__ void __launch_bounds__(32 * WRP_PER_CTA) k_glb_warp(const DPJob *jobs, const uint32_t *order, uint32_t njobs, DevReads R, DPPar P,
		uint32_t *arena, DPSlab SB, uint32_t *cig_arena, DPRes *res, unsigned long long *ctr, int ctr_work, int ctr_cells){
	__shared__ int s_h[WRP_PER_CTA][3 * WRP_CAP];
	__shared__ uint32_t s_seq[WRP_PER_CTA][WRP_SEQW];
	__shared__ int s_misc[WRP_PER_CTA][16];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	ExecSmem<32> X; X.carve(s_h[warp], WRP_CAP, s_seq[warp], WRP_SEQW, nullptr, nullptr, s_misc[warp]);
	const uint32_t ex = blockIdx.x * WRP_PER_CTA + warp;
	uint32_t *slab = SB.off? arena + SB.base + SB.off[ex] : nullptr;
	for(bool first = true; ; first = false){
		uint32_t jn = ex;
		if(!first){ if(lane == 0) jn = gridDim.x * WRP_PER_CTA + (uint32_t)atomicAdd(ctr + ctr_work, 1ULL); jn = __shfl_sync(0xffffffffu, jn, 0); }
		if(jn >= njobs) break;
		run_glb_job<32, WRP_C>(jobs[order? order[jn] : jn], R, P, X, arena, slab, cig_arena, res, ctr + ctr_cells, lane);
		__syncwarp();
	}
}

/* CTA-per-job global kernel for gaps whose band does not fit a warp executor comfortably */
__global__ void __launch_bounds__(EXT_NT) k_glb_cta(const DPJob *jobs, const uint32_t *order, uint32_t njobs, DevReads R, DPPar P,
		uint32_t *arena, DPSlab SB, uint32_t *cig_arena, DPRes *res, unsigned long long *ctr, int ctr_work, int ctr_cells){
	__shared__ int s_h[3 * EXT_CAP];
	__shared__ uint32_t s_seq[EXT_SEQW];
	__shared__ int s_red[2 * (EXT_NT / 32)];
	__shared__ long long s_redk[EXT_NT / 32];
	__shared__ int s_misc[16];
	__shared__ uint32_t s_job;
	ExecSmem<EXT_NT> X; X.carve(s_h, EXT_CAP, s_seq, EXT_SEQW, s_red, s_redk, s_misc);
	const int tid = threadIdx.x;
	uint32_t *slab = SB.off? arena + SB.base + SB.off[blockIdx.x] : nullptr;
	for(bool first = true; ; first = false){
		if(tid == 0) s_job = first? blockIdx.x : gridDim.x + (uint32_t)atomicAdd(ctr + ctr_work, 1ULL);
		__syncthreads();
		const uint32_t jn = s_job;
		__syncthreads();
		if(jn >= njobs) break;
		run_glb_job<EXT_NT, EXT_C>(jobs[order? order[jn] : jn], R, P, X, arena, slab, cig_arena, res, ctr + ctr_cells, tid);
		__syncthreads();
	}
}
