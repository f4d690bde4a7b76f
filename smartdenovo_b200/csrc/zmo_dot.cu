/*
 * zmo_dot.cu -- dot-matrix ("-U") mode on the device: replaces dot_matrix_align_hzmps for every
 * (query, candidate) pair of a batch (wtzmo.c:853-863).  The z-index / z-match front end is shared
 * with the SW path (seed_prepare, zmo_seed.cu); the block denoising, merging and overhang chaining
 * run one warp per pair: lane 0 executes the order-exact serial logic of zmo_dot_core.cuh, the other lanes
 * stage its sorts through shared memory.
 */
#include "zmo_seed.cuh"
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

extern "C" int zmo_pair_dotmatrix(zmo_ctx *c, const zmo_pair_t *pairs, uint32_t np, zmo_dotres_t *out){
	if(!c || (np && (!pairs || !out))) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	if(c->st->n_reads == 0) return zmo_set_err(ZMO_ERR_STATE, "no reads uploaded");
	if(np == 0) return 0;
	CUDA_TRY(cudaSetDevice(c->device));
	StageTimer tm(c, ST_DOT);
	SeedWork W;
	if(int rc = seed_prepare(c, pairs, np, 1, W, c->s5)) return rc;
	const size_t per = 4 + sizeof(DevDiag) + 4 + 4 + sizeof(DevZPairG) + sizeof(DevWin) + 16;     /* zmo_dot_scratch_bytes(n) = (n+2)*per + 64 */
	/* scratch in s6; pairs + results in s1 (the per-read z counters there are dead) */
	if(c->s6.reserve(W.T * per + (size_t)(2 * per + 64) * np + 256) || c->s1.reserve((size_t)np * (sizeof(zmo_pair_t) + sizeof(zmo_dotres_t)) + 128)) return ZMO_ERR_CUDA;
	zmo_pair_t *d_pairs = c->s1.as<zmo_pair_t>(); zmo_dotres_t *d_out = (zmo_dotres_t*)(d_pairs + np + 1);
	CUDA_TRY(cudaMemcpyAsync(d_pairs, pairs, (size_t)np * sizeof(zmo_pair_t), cudaMemcpyHostToDevice, c->stream));
	DotPar par; par.xvar = c->par.xvar; par.yvar = c->par.yvar; par.min_block_len = c->par.min_block_len; par.max_overhang = c->par.max_overhang;
	par.deviation_penalty = c->par.deviation_penalty; par.gap_penalty = c->par.gap_penalty;
	{
		unsigned long long *ctr = c->d_ctr.as<unsigned long long>();
		CUDA_TRY(cudaMemsetAsync(ctr + CTR_WORK, 0, 8, c->stream));
		const int grid = (int)std::min<uint64_t>((np + DOT_WARPS - 1) / DOT_WARPS, (uint64_t)c->n_sm * 4);
		k_p_dot<<<grid, 32 * DOT_WARPS, 0, c->stream>>>(W.cache_off, d_pairs, np, W.cache, W.tie, c->s6.as<uint8_t>(), per, dev_reads(c), par, (uint32_t)c->par.zsize, (uint32_t)c->par.ztot, d_out, ctr + CTR_WORK);
		c->launches++;
	}
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpyAsync(out, d_out, (size_t)np * sizeof(zmo_dotres_t), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	c->counters[5] += (size_t)np * sizeof(zmo_pair_t); c->counters[6] += (size_t)np * sizeof(zmo_dotres_t);
	return 0;
}
