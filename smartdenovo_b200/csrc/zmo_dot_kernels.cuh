/*
 * zmo_dot_kernels.cuh -- the warp-per-pair dot-matrix kernel (dot_matrix_align_hzmps, hzm_aln.h:1134-1181).  Kept in a header so
 * that the test-only host simulation (tests/hostsim/seedk_host.cpp) runs this very source; included by zmo_dot.cu only.
 */
#pragma once
#include "zmo_ctx.cuh"
#include "zmo_dot_core.cuh"

/* One WARP per pair (persistent warps pulling pairs from a work counter).  Lane 0 runs the order-exact serial logic of
 * zmo_dot_core.cuh; whenever it reaches one of the sort_array emulations that dominate a pair's time it posts the array in the
 * warp's mailbox, all 32 lanes stage it into shared memory (coalesced), lane 0 sorts it there -- shared-memory latency
 * instead of a dependent global-memory round trip per comparison -- and the lanes copy it back.  The other lanes wait in a
 * helper loop; the __syncwarp()s of the two code paths pair up one to one.  Arrays that do not fit DOT_BUF words are sorted
 * in place. */
#define DOT_WARPS 8
#define DOT_BUF 640              /* 64-bit words of sort staging per warp (5 KB) */
struct DotMail { void *ptr; uint32_t n, words, cmd; };       /* cmd: 1 = staged sort, 2 = pair done */
struct WarpSort {
	DotMail *m; unsigned long long *buf;
	__device__ __forceinline__ static void stage(const DotMail *m, unsigned long long *buf, int lane, bool in){
		const uint32_t tot = m->n * m->words; unsigned long long *a = (unsigned long long*)m->ptr;
		if(in){ for(uint32_t i = lane; i < tot; i += 32) buf[i] = a[i]; }
		else { for(uint32_t i = lane; i < tot; i += 32) a[i] = buf[i]; }
	}
	template<class T, class GT> __device__ void operator()(T *a, size_t n, GT gt) const {
		static_assert(sizeof(T) % 8 == 0, "staged in 64-bit words");
		constexpr uint32_t W = sizeof(T) / 8;
		if(n < 24 || n * W > DOT_BUF || ((uintptr_t)a & 7)){ zmo_ref_sort(a, n, gt); return; }
		m->ptr = a; m->n = (uint32_t)n; m->words = W; m->cmd = 1;
		__syncwarp();
		stage(m, buf, 0, true);
		__syncwarp();
		zmo_ref_sort((T*)buf, n, gt);
		__syncwarp();
		stage(m, buf, 0, false);
		__syncwarp();
	}
};
__global__ void __launch_bounds__(32 * DOT_WARPS) k_p_dot(const unsigned long long *cache_off, const zmo_pair_t *pairs, uint32_t np, DevZPair *cache, const uint8_t *tie, uint8_t *scratch, size_t per, DevReads R, DotPar par, uint32_t zsize, uint32_t ztot, zmo_dotres_t *out, unsigned long long *work){
	__shared__ __align__(16) unsigned long long s_buf[DOT_WARPS][DOT_BUF];
	__shared__ DotMail s_mail[DOT_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned long long *buf = s_buf[warp]; DotMail *m = &s_mail[warp];
	while(1){
		uint32_t p = 0;
		if(lane == 0) p = (uint32_t)atomicAdd(work, 1ULL);
		p = __shfl_sync(0xffffffffu, p, 0);
		if(p >= np) break;
		const unsigned long long c0 = cache_off[p]; const uint32_t n = (uint32_t)(cache_off[p + 1] - c0);
		if((unsigned long long)n * zsize < ztot){
			if(lane == 0){ zmo_dotres_t o; o.n_zpair = n; o.score = 0; o.qb = o.tb = 0x7FFFFFFF; o.qe = o.te = 0; o.strand = 0; out[p] = o; }
			continue;
		}
		if(lane == 0){
			WarpSort ws; ws.m = m; ws.buf = buf;
			const DotRes r = zmo_dot_pair(cache + c0, n, (int)R.len[pairs[p].qid], (int)R.len[pairs[p].cid], par, scratch + c0 * per + (size_t)(2 * per + 64) * p, tie[p]? 2 : 1, ws);
			zmo_dotres_t o; o.n_zpair = n; o.score = r.score; o.qb = r.qb; o.qe = r.qe; o.tb = r.tb; o.te = r.te; o.strand = r.strand;
			out[p] = o;
			m->cmd = 2;
			__syncwarp();
		} else {
			while(1){
				__syncwarp();
				if(m->cmd == 2) break;
				WarpSort::stage(m, buf, lane, true);
				__syncwarp();
				__syncwarp();
				WarpSort::stage(m, buf, lane, false);
				__syncwarp();
			}
		}
		__syncwarp();
	}
}

