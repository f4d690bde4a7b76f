/*
 * zmo_dp.cu -- DP kernels (persistent executors pulling jobs from a device work counter) and the
 * stand-alone DP operators of the C ABI.
 */
#include "zmo_jobs.cuh"
#include "zmo_dp_kernels.cuh"

static DPPar dp_par(const zmo_ctx *c){ DPPar P; P.M = c->par.M; P.X = c->par.X; P.I = c->par.O; P.D = c->par.O; P.E = c->par.E; P.T = c->par.T; return P; }

/* ---- launch helpers used by the API and the pipeline --------------------------------------- */
template<int NT, int C> static void launch_ext_cta(cudaStream_t st, int wk, int mode, int grid, const DPJob *d_jobs, const uint32_t *d_order, uint32_t n, DevReads R, DPPar P, uint32_t *arena, DPSlab SB, uint32_t *cig, DPRes *d_res, unsigned long long *ctr, int ctr_cells){
	if(mode == 1) k_ext_cta<NT, C, 1><<<grid, NT, 0, st>>>(d_jobs, d_order, n, R, P, arena, SB, cig, d_res, ctr, wk, ctr_cells);
	else k_ext_cta<NT, C, 0><<<grid, NT, 0, st>>>(d_jobs, d_order, n, R, P, arena, SB, cig, d_res, ctr, wk, ctr_cells);
}
/* grid (CTAs) and executor count of a DP launch: the pipeline sizes the executor slabs with the same numbers */
int zmo_ext_grid(const zmo_ctx *c, int cls, uint32_t n, uint32_t *n_exec){
	int grid;
	if(cls == 0){ grid = (int)std::min<uint64_t>((n + WRP_PER_CTA - 1) / WRP_PER_CTA, (uint64_t)c->n_sm * 8); *n_exec = (uint32_t)grid * WRP_PER_CTA; }
	else { grid = (int)std::min<uint64_t>(n, (uint64_t)c->n_sm * (cls == 1? 8 : (cls == 2? 5 : (CL3_NT == 256? 2 : 3)))); *n_exec = (uint32_t)grid; }
	return grid;
}
int zmo_glb_grid(const zmo_ctx *c, bool wide, uint32_t n, uint32_t *n_exec){
	int grid;
	if(wide){ grid = (int)std::min<uint64_t>(n, (uint64_t)c->n_sm * 4); *n_exec = (uint32_t)grid; }
	else { grid = (int)std::min<uint64_t>((n + WRP_PER_CTA - 1) / WRP_PER_CTA, (uint64_t)c->n_sm * 8); *n_exec = (uint32_t)grid * WRP_PER_CTA; }
	return grid;
}
/* cls: 0 = warp executor, 1/2/3 = CTA executors (see ext_class) */
int zmo_launch_ext_on(zmo_ctx *c, cudaStream_t st, int wk, int mode, int cls, const DPJob *d_jobs, const uint32_t *d_order, uint32_t n, uint32_t *arena, DPSlab SB, uint32_t *cig, DPRes *d_res, int ctr_cells){
	if(n == 0) return 0;
	unsigned long long *ctr = c->d_ctr.as<unsigned long long>();
	CUDA_TRY(cudaMemsetAsync(ctr + wk, 0, 8, st));
	DevReads R = dev_reads(c); DPPar P = dp_par(c);
	uint32_t nex = 0; const int grid = zmo_ext_grid(c, cls, n, &nex);
	if(cls == 0){
		if(mode == 1) k_ext_warp<1><<<grid, 32 * WRP_PER_CTA, 0, st>>>(d_jobs, d_order, n, R, P, arena, SB, cig, d_res, ctr, wk, ctr_cells);
		else k_ext_warp<0><<<grid, 32 * WRP_PER_CTA, 0, st>>>(d_jobs, d_order, n, R, P, arena, SB, cig, d_res, ctr, wk, ctr_cells);
	} else if(cls == 1) launch_ext_cta<64, 7>(st, wk, mode, grid, d_jobs, d_order, n, R, P, arena, SB, cig, d_res, ctr, ctr_cells);
	else if(cls == 2) launch_ext_cta<128, 7>(st, wk, mode, grid, d_jobs, d_order, n, R, P, arena, SB, cig, d_res, ctr, ctr_cells);
	else launch_ext_cta<CL3_NT, CL3_C>(st, wk, mode, grid, d_jobs, d_order, n, R, P, arena, SB, cig, d_res, ctr, ctr_cells);
	c->launches++;
	CUDA_TRY(cudaGetLastError());
	return 0;
}
int zmo_launch_ext(zmo_ctx *c, int mode, int cls, const DPJob *d_jobs, const uint32_t *d_order, uint32_t n, uint32_t *arena, uint32_t *cig, DPRes *d_res, int ctr_cells){
	DPSlab SB; SB.base = 0; SB.off = nullptr;
	return zmo_launch_ext_on(c, c->stream, CTR_WORK, mode, cls, d_jobs, d_order, n, arena, SB, cig, d_res, ctr_cells);
}
int zmo_launch_glb_on(zmo_ctx *c, cudaStream_t st, int wk, bool wide, const DPJob *d_jobs, const uint32_t *d_order, uint32_t n, uint32_t *arena, DPSlab SB, uint32_t *cig, DPRes *d_res, int ctr_cells){
	if(n == 0) return 0;
	unsigned long long *ctr = c->d_ctr.as<unsigned long long>();
	CUDA_TRY(cudaMemsetAsync(ctr + wk, 0, 8, st));
	DevReads R = dev_reads(c); DPPar P = dp_par(c);
	uint32_t nex = 0; const int grid = zmo_glb_grid(c, wide, n, &nex);
	if(wide) k_glb_cta<<<grid, EXT_NT, 0, st>>>(d_jobs, d_order, n, R, P, arena, SB, cig, d_res, ctr, wk, ctr_cells);
	else k_glb_warp<<<grid, 32 * WRP_PER_CTA, 0, st>>>(d_jobs, d_order, n, R, P, arena, SB, cig, d_res, ctr, wk, ctr_cells);
	c->launches++;
	CUDA_TRY(cudaGetLastError());
	return 0;
}
int zmo_launch_glb(zmo_ctx *c, bool wide, const DPJob *d_jobs, const uint32_t *d_order, uint32_t n, uint32_t *arena, uint32_t *cig, DPRes *d_res, int ctr_cells){
	DPSlab SB; SB.base = 0; SB.off = nullptr;
	return zmo_launch_glb_on(c, c->stream, CTR_WORK, wide, d_jobs, d_order, n, arena, SB, cig, d_res, ctr_cells);
}

/* ---- stand-alone operators ------------------------------------------------------------------ */
static int dp_batch(zmo_ctx *c, int kind /*0 ext mode0, 1 ext mode1, 2 global*/, const zmo_dp_problem_t *probs, const int32_t *wv, uint32_t n,
		zmo_dp_result_t *res, uint32_t *cigars, uint64_t cigar_cap, uint64_t *cigar_needed){
	if(!c || (n && (!probs || !res))) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	if(n == 0){ if(cigar_needed) *cigar_needed = 0; return 0; }
	CUDA_TRY(cudaSetDevice(c->device));
	DPPar P = dp_par(c);
	std::vector<DPJob> jl[4];       /* per executor class (global: class 3 = CTA, class 0 = warp) */
	uint64_t scratch = 0, cig = 0;
	for(uint32_t i = 0; i < n; i++){
		const zmo_dp_problem_t &p = probs[i]; DPJob J; memset(&J, 0, sizeof(J));
		if(p.q_rid >= c->st->n_reads || p.t_rid >= c->st->n_reads) return zmo_set_err(ZMO_ERR_ARG, "problem %u: read id out of range", i);
		{
			/* both slices must lie inside their reads: element k is base start + k*step */
			const long long ql = c->st->h_rdlen[p.q_rid], tl = c->st->h_rdlen[p.t_rid];
			const long long qa = p.q_start, qz = (long long)p.q_start + (long long)(p.qlen > 0? p.qlen - 1 : 0) * p.q_step;
			const long long ta = p.t_start, tz = (long long)p.t_start + (long long)(p.tlen > 0? p.tlen - 1 : 0) * p.t_step;
			if(p.qlen < 0 || p.tlen < 0 || (p.q_step != 1 && p.q_step != -1) || (p.t_step != 1 && p.t_step != -1)) return zmo_set_err(ZMO_ERR_ARG, "problem %u: negative length or step not +-1", i);
			if((p.qlen > 0 && (qa < 0 || qa >= ql || qz < 0 || qz >= ql)) || (p.tlen > 0 && (ta < 0 || ta >= tl || tz < 0 || tz >= tl))) return zmo_set_err(ZMO_ERR_ARG, "problem %u: slice outside its read", i);
			if(kind == 2 && wv[i] <= 0) return zmo_set_err(ZMO_ERR_ARG, "problem %u: band must be positive", i);
		}
		J.q_rid = p.q_rid; J.t_rid = p.t_rid; J.q_start = p.q_start; J.q_step = p.q_step; J.q_comp = p.q_comp; J.qlen = p.qlen;
		J.t_start = p.t_start; J.t_step = p.t_step; J.t_comp = p.t_comp; J.tlen = p.tlen; J.init = p.init_score; J.Wp = p.W; J.Wmax = 0;
		J.out_idx = i; J.cig_off = cig; J.cig_cap = (uint32_t)((p.qlen > 0? p.qlen : 0) + (p.tlen > 0? p.tlen : 0) + 4);
		cig += J.cig_cap;
		int cls;
		if(kind == 2){
			J.Wp = wv[i];
			int w = J.Wp, dl = abs(p.qlen - p.tlen); while(w < dl) w <<= 1;
			int bw = std::min(p.qlen, 2 * w + 1);
			cls = bw > 32 * WRP_C * 2? 3 : 0;
			J.scratch = scratch;
			scratch += cls? glb_scratch_words<EXT_NT, EXT_C>(p.qlen, p.tlen, EXT_CAP) : glb_scratch_words<32, WRP_C>(p.qlen, p.tlen, WRP_CAP);
		} else {
			int init = p.init_score < 0? 0 : p.init_score;
			BandDims d; d.W = 0; d.ql = d.tl = d.ncol = 0;
			if(p.qlen > 0 && p.tlen > 0) d = band_dims(p.qlen, p.tlen, init, p.W, P);
			cls = ext_class(d.ncol);
			J.scratch = scratch;
			scratch += ext_scratch_words_cls(d, cls);
		}
		jl[cls].push_back(J);
	}
	if(cigar_needed) *cigar_needed = cig;
	if(cig > cigar_cap) return zmo_set_err(ZMO_ERR_CAPACITY, "cigar buffer too small: need %llu", (unsigned long long)cig);
	if(c->arena.reserve((scratch + 64) * 4)) return ZMO_ERR_CUDA;
	if(c->s0.reserve(((size_t)n + 1) * sizeof(DPJob))) return ZMO_ERR_CUDA;
	if(c->s1.reserve((size_t)(n + 1) * sizeof(DPRes))) return ZMO_ERR_CUDA;
	if(c->s2.reserve((cig + 16) * 4)) return ZMO_ERR_CUDA;
	DPJob *dj = c->s0.as<DPJob>(); size_t joff[5] = {0, 0, 0, 0, 0};
	for(int k = 0; k < 4; k++){
		joff[k + 1] = joff[k] + jl[k].size();
		if(jl[k].size()) CUDA_TRY(cudaMemcpyAsync(dj + joff[k], jl[k].data(), jl[k].size() * sizeof(DPJob), cudaMemcpyHostToDevice, c->stream));
	}
	{
		StageTimer t(c, kind == 2? ST_GAP : (kind == 1? ST_EXT : ST_WINALN));
		int cc = kind == 2? CTR_CELLS_GAP : (kind == 1? CTR_CELLS_EXT : CTR_CELLS_WIN);
		for(int k = 0; k < 4; k++){
			if(kind == 2){ if(zmo_launch_glb(c, k != 0, dj + joff[k], nullptr, (uint32_t)jl[k].size(), c->arena.as<uint32_t>(), c->s2.as<uint32_t>(), c->s1.as<DPRes>(), cc)) return ZMO_ERR_CUDA; }
			else { if(zmo_launch_ext(c, kind, k, dj + joff[k], nullptr, (uint32_t)jl[k].size(), c->arena.as<uint32_t>(), c->s2.as<uint32_t>(), c->s1.as<DPRes>(), cc)) return ZMO_ERR_CUDA; }
		}
	}
	std::vector<DPRes> hr(n);
	CUDA_TRY(cudaMemcpyAsync(hr.data(), c->s1.p, (size_t)n * sizeof(DPRes), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(cigars, c->s2.p, cig * 4, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	/* CIGARs come back in walk order (end -> start); flip each to alignment order like the reference's reverse_u32list */
	std::vector<const DPJob*> all; for(int k = 0; k < 4; k++) for(auto &j : jl[k]) all.push_back(&j);
	for(const DPJob *j : all){
		const DPRes &r = hr[j->out_idx]; zmo_dp_result_t &o = res[j->out_idx];
		o.score = r.score; o.qe = r.qe; o.te = r.te; o.mat = r.mat; o.mis = r.mis; o.ins = r.ins; o.del = r.del; o.aln = r.mat + r.mis + r.ins + r.del;
		o.cigar_off = j->cig_off; o.n_cigar = (uint32_t)r.ncig; o.cells = (uint64_t)r.w_used;
		uint32_t *cg = cigars + j->cig_off;
		for(int a = 0, b = r.ncig - 1; a < b; a++, b--){ uint32_t t = cg[a]; cg[a] = cg[b]; cg[b] = t; }
	}
	return 0;
}

extern "C" int zmo_dp_extend(zmo_ctx *ctx, int mode, const zmo_dp_problem_t *probs, uint32_t n, zmo_dp_result_t *res, uint32_t *cigars, uint64_t cigar_cap, uint64_t *cigar_needed){
	if(mode != 0 && mode != 1) return zmo_set_err(ZMO_ERR_ARG, "mode must be 0 or 1");
	return dp_batch(ctx, mode, probs, nullptr, n, res, cigars, cigar_cap, cigar_needed);
}
extern "C" int zmo_dp_global(zmo_ctx *ctx, const zmo_dp_problem_t *probs, const int32_t *w, uint32_t n, zmo_dp_result_t *res, uint32_t *cigars, uint64_t cigar_cap, uint64_t *cigar_needed){
	if(n && !w) return zmo_set_err(ZMO_ERR_ARG, "null band array");
	return dp_batch(ctx, 2, probs, w, n, res, cigars, cigar_cap, cigar_needed);
}
