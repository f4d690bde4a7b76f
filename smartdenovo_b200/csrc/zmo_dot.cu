/*
 * zmo_dot.cu -- dot-matrix ("-U") mode on the device: replaces dot_matrix_align_hzmps for every
 * (query, candidate) pair of a batch (wtzmo.c:853-863).  The z-index / z-match front end is shared
 * with the SW path (seed_prepare, zmo_seed.cu); the block denoising, merging and overhang chaining
 * run one warp per pair: lane 0 executes the order-exact serial logic of zmo_dot_core.cuh, the other lanes
 * stage its sorts through shared memory.
 */
#include "zmo_seed.cuh"
#include "zmo_dot_core.cuh"
#include "zmo_dot_kernels.cuh"

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
