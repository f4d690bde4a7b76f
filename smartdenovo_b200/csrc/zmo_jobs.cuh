/*
 * zmo_jobs.cuh -- DP job descriptors and the per-executor job runners shared by the stand-alone DP
 * operators (zmo_dp_extend / zmo_dp_global) and the pair-alignment pipeline.
 */
#pragma once
#include "zmo_ctx.cuh"
#include "zmo_dp.cuh"
#include "zmo_dpr.cuh"

struct DPJob {
	uint32_t q_rid, t_rid;
	int q_start, q_step, q_comp, qlen;
	int t_start, t_step, t_comp, tlen;
	int init, Wp;                 /* extension: init score and W as the reference passes it; global: Wp = first band w */
	int Wmax;                     /* global only: -W cap for band doubling (hzm_aln.h:1411) */
	unsigned long long scratch;   /* word offset of this job's scratch in the arena (stand-alone operators; the pipeline uses executor slabs) */
	unsigned long long cig_off;   /* word offset in the cigar arena */
	uint32_t cig_cap;
	uint32_t out_idx;
	uint32_t est, sw32;           /* estimated DP cells / 256 (saturating); scratch words / 32 (rounded up): the work-queue order and the slab sizes */
};
/* executor-private scratch of a DP kernel launch.  Jobs are handed out longest-scratch-first: executor e starts with job e
 * and then pulls from the shared counter, so everything it will ever run fits the scratch of job e.  off[e] (words, relative
 * to base) is the prefix sum of those sizes; off == nullptr: per-job scratch at J.scratch (stand-alone operators). */
struct DPSlab { unsigned long long base; const unsigned long long *off; };
struct DPRes { int score, qe, te, mat, mis, ins, del, ncig, w_used, pad; };

/* scratch words needed by an extension job (host + device agree through this one function) */
template<int NT, int C> __host__ __device__ inline unsigned long long ext_scratch_words(const BandDims &d, int sm_cap){
	unsigned long long zw = (unsigned long long)d.ql * band_row_words<NT, C>(d.ncol);
	unsigned long long seq = (unsigned long long)((d.ql + 15) >> 4) + ((d.tl + 15) >> 4) + 2;
	unsigned long long hb = 0;
	int need = 2 * d.W + 3 < d.tl + 2? 2 * d.W + 3 : d.tl + 2;       /* columns that can be live in the H/E rows */
	if(need > sm_cap){ unsigned long long cap = 1; while(cap < (unsigned long long)need) cap <<= 1; hb = 3 * cap; }
	return zw + (unsigned long long)d.ql + seq + hb + 8;
}
/* executor classes of an extension band (register-resident sweeps, zmo_dpr.cuh): 0 = warp x 7 columns per lane,
 * 1/2 = CTA of 64/128 threads x 7, 3 = CTA of CL3_NT threads x CL3_C columns; bands beyond class 3's capacity run the
 * chunked shared/global-memory sweep of zmo_dp.cuh inside the class-3 kernel */
#define CL3_NT 128
#define CL3_C 13
__host__ __device__ inline int ext_class(int ncol){ return ncol <= RegCap<32, 7>::ncol? 0 : (ncol <= RegCap<64, 7>::ncol? 1 : (ncol <= RegCap<128, 7>::ncol? 2 : 3)); }
__host__ __device__ inline unsigned long long ext_scratch_words_cls(const BandDims &d, int cls){
	const unsigned long long seq = (unsigned long long)((d.ql + 15) >> 4) + ((d.tl + 15) >> 4) + 2;
	if(cls == 0) return (unsigned long long)d.ql * 32 + seq + 8;
	if(cls == 1) return (unsigned long long)d.ql * 64 + seq + 8;
	if(cls == 2) return (unsigned long long)d.ql * 128 + seq + 8;
	if(d.ncol <= RegCap<CL3_NT, CL3_C>::ncol) return (unsigned long long)d.ql * CL3_NT * RegCap<CL3_NT, CL3_C>::WPT + seq + 8;
	return ext_scratch_words<CL3_NT, 7>(d, 0);
}
/* scratch for a global job sized for the widest band the retry loop can reach (ncol <= qlen) */
template<int NT, int C> __host__ __device__ inline unsigned long long glb_scratch_words(int qlen, int tlen, int sm_cap){
	unsigned long long zw = (unsigned long long)(tlen > 0? tlen : 0) * band_row_words<NT, C>(qlen > 0? qlen : 0);
	unsigned long long seq = (unsigned long long)((qlen + 15) >> 4) + ((tlen + 15) >> 4) + 2;
	unsigned long long hb = 0;
	int need = qlen + 3;
	if(need > sm_cap){ unsigned long long cap = 1; while(cap < (unsigned long long)need) cap <<= 1; hb = 3 * cap; }
	return zw + seq + hb + 8;
}

__device__ __forceinline__ SeqView job_view(const DevReads &R, uint32_t rid, int start, int step, int comp){
	SeqView v; v.w = R.words + R.woff[rid]; v.start = start; v.step = step; v.comp = comp? 3u : 0u; return v;
}

/* executor-level shared memory carve-up: [H0|H1|Ev] cap ints each, seq words, reduction scratch */
template<int NT> struct ExecSmem {
	BandSmem B; uint32_t *seq; int seq_words; int cap;
	__device__ void carve(int *base_int, int cap_, uint32_t *seq_, int seq_words_, int *red, long long *redk, int *misc){
		cap = cap_; B.H0 = base_int; B.H1 = base_int + cap_; B.Ev = base_int + 2 * cap_; B.cap_mask = cap_ - 1;
		B.sred = red; B.sredk = redk; B.smisc = misc; seq = seq_; seq_words = seq_words_;
	}
};

template<int NT, int C, int MODE>
__device__ void run_ext_job(const DPJob &J, const DevReads &R, const DPPar &P, ExecSmem<NT> &X, uint32_t *arena, uint32_t *slab, uint32_t *cig_arena,
		DPRes *res, unsigned long long *cells, int tid){
	int init = J.init < 0? 0 : J.init;
	DPOut o; o.score = init; o.qe = o.te = o.mat = o.mis = o.ins = o.del = o.ncig = 0;
	if(J.qlen > 0 && J.tlen > 0){
		BandDims d = band_dims(J.qlen, J.tlen, init, J.Wp, P);
		uint32_t *scr = slab? slab : arena + J.scratch;      /* executor-private slab (pipeline) or per-job scratch (stand-alone operators) */
		const bool reg = d.ncol <= RegCap<NT, C>::ncol;
		const int rw = reg? NT * RegCap<NT, C>::WPT : band_row_words<NT, 7>(d.ncol);
		uint32_t *z = scr; scr += (size_t)d.ql * rw;
		int *zb = (int*)scr; if(!reg) scr += d.ql;
		const int qw = (d.ql + 15) >> 4, tw = (d.tl + 15) >> 4;
		uint32_t *qpk, *tpk;
		if(qw + tw + 2 <= X.seq_words){ qpk = X.seq; tpk = X.seq + qw; }
		else { qpk = scr; tpk = scr + qw; }
		scr += qw + tw + 2;
		stage_packed<NT>(job_view(R, J.q_rid, J.q_start, J.q_step, J.q_comp), d.ql, qpk, tid);
		stage_packed<NT>(job_view(R, J.t_rid, J.t_start, J.t_step, J.t_comp), d.tl, tpk, tid);
		ex_sync<NT>();
		if(reg) reg_extend<NT, C, MODE>(X.B, qpk, J.qlen, tpk, J.tlen, init, d, P, z, cig_arena + J.cig_off, (int)J.cig_cap, o, cells, tid);
		else {
			/* band wider than the register executor: chunked sweep with H/E rows in shared (if they fit) or global memory */
			BandSmem S = X.B;
			const int need = 2 * d.W + 3 < d.tl + 2? 2 * d.W + 3 : d.tl + 2;
			if(need > X.cap){ int cap = 1; while(cap < need) cap <<= 1; S.H0 = (int*)scr; S.H1 = S.H0 + cap; S.Ev = S.H1 + cap; S.cap_mask = cap - 1; }
			band_extend<NT, 7, MODE>(S, qpk, J.qlen, tpk, J.tlen, init, d, P, z, zb, cig_arena + J.cig_off, (int)J.cig_cap, o, cells, tid);
		}
	}
	if(tid == 0){ DPRes r; r.score = o.score; r.qe = o.qe; r.te = o.te; r.mat = o.mat; r.mis = o.mis; r.ins = o.ins; r.del = o.del; r.ncig = o.ncig; r.w_used = 0; r.pad = 0; res[J.out_idx] = r; }
}

/* gap filling with the reference's band-doubling retry (hzm_aln.h:1400-1418) */
template<int NT, int C>
__device__ void run_glb_job(const DPJob &J, const DevReads &R, const DPPar &P, ExecSmem<NT> &X, uint32_t *arena, uint32_t *slab, uint32_t *cig_arena,
		DPRes *res, unsigned long long *cells, int tid){
	const int qlen = J.qlen, tlen = J.tlen;
	uint32_t *scr = slab? slab : arena + J.scratch;
	const int rwmax = band_row_words<NT, C>(qlen > 0? qlen : 0);
	uint32_t *z = scr; scr += (size_t)(tlen > 0? tlen : 0) * rwmax;
	const int qw = (qlen + 15) >> 4, tw = (tlen + 15) >> 4;
	uint32_t *qpk, *tpk;
	if(qw + tw + 2 <= X.seq_words){ qpk = X.seq; tpk = X.seq + qw; }
	else { qpk = scr; tpk = scr + qw; }
	scr += qw + tw + 2;
	BandSmem Sg = X.B;      /* global-memory H/E rows for bands wider than the shared-memory capacity */
	if(qlen + 3 > X.cap){ int cap = 1; while(cap < qlen + 3) cap <<= 1; Sg.H0 = (int*)scr; Sg.H1 = Sg.H0 + cap; Sg.Ev = Sg.H1 + cap; Sg.cap_mask = cap - 1; }
	stage_packed<NT>(job_view(R, J.q_rid, J.q_start, J.q_step, J.q_comp), qlen, qpk, tid);
	stage_packed<NT>(job_view(R, J.t_rid, J.t_start, J.t_step, J.t_comp), tlen, tpk, tid);
	ex_sync<NT>();
	int w = J.Wp; DPOut o;
	const int dl = qlen > tlen? qlen - tlen : tlen - qlen, mxl = qlen > tlen? qlen : tlen;
	while(1){
		if(w < dl){ w <<= 1; continue; }
		{
			const int nc = qlen < 2 * w + 1? qlen : 2 * w + 1, bw = nc + 3;
			uint32_t *cg = cig_arena + J.cig_off; const int cc = (int)J.cig_cap;
			if(NT == 32 && nc <= RegCap<32, 1>::ncol) reg_global<NT, 1>(X.B, qpk, qlen, tpk, tlen, w, P, z, cg, cc, o, cells, tid);
			else if(NT == 32 && nc <= RegCap<32, 2>::ncol) reg_global<NT, 2>(X.B, qpk, qlen, tpk, tlen, w, P, z, cg, cc, o, cells, tid);
			else if(NT == 32 && nc <= RegCap<32, 4>::ncol) reg_global<NT, 4>(X.B, qpk, qlen, tpk, tlen, w, P, z, cg, cc, o, cells, tid);
			else if(nc <= RegCap<NT, C>::ncol) reg_global<NT, C>(X.B, qpk, qlen, tpk, tlen, w, P, z, cg, cc, o, cells, tid);
			else { const BandSmem &SS = bw <= X.cap? X.B : Sg; band_global<NT, C>(SS, qpk, qlen, tpk, tlen, w, P, z, cg, cc, o, cells, tid); }
		}
		if(J.Wmax > 0 && o.score < 0 && w < J.Wmax && w < mxl) w <<= 1; else break;
	}
	if(tid == 0){ DPRes r; r.score = o.score; r.qe = o.qe; r.te = o.te; r.mat = o.mat; r.mis = o.mis; r.ins = o.ins; r.del = o.del; r.ncig = o.ncig; r.w_used = w; r.pad = 0; res[J.out_idx] = r; }
}
