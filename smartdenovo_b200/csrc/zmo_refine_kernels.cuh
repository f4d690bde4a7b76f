/*
 * zmo_refine_kernels.cuh -- -n refinement (kswx_refine_alignment, kswx.h:483-659) on the device: per-row band from the stitched CIGAR
 * (k_refine_band) and the warp / CTA executors running the variable-band sweep (reg_refine, zmo_dpr.cuh).  Kept in a header so that the
 * test-only host simulation (tests/hostsim/dp_host.cpp) runs this very source; included by zmo_align.cu only.
 */
#pragma once
#include "zmo_winalign.cuh"

struct RefJob { uint32_t task; int ql, tl; unsigned long long band, scratch, out; uint32_t out_cap, pad; };
#define REFW_NT 256            /* executor of the wide-band fallback: 256 threads x 7 columns per chunk */
#define REFW_C 7
#define REF_SEQW 1280          /* staged sequence words kept in shared memory by a refine executor (<= ~10 kb per side) */
__global__ void k_refine_size(const zmo_record_t *recs, uint32_t nt, unsigned long long *rows, unsigned long long *outw){
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= nt) return;
	const zmo_record_t r = recs[t]; const long long ql = (long long)r.qe - r.qb, tl = (long long)r.te - r.tb;
	const bool v = r.ok && ql > 0 && tl > 0;
	rows[t] = v? (unsigned long long)ql + 2 : 0; outw[t] = v? (unsigned long long)(ql + tl + 4) : 0;
}
__global__ void k_refine_band(zmo_record_t *recs, uint32_t nt, const uint32_t *ops, const unsigned long long *row_off, const unsigned long long *out_off, int W,
		int *bands, RefJob *jobs, unsigned long long *njobs, unsigned long long *scr_words){
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= nt) return;
	scr_words[t] = 0;
	zmo_record_t r = recs[t];
	if(!r.ok) return;
	const int ql = r.qe - r.qb, tl = r.te - r.tb;
	if(ql <= 0 || tl <= 0){
		/* KSWX_NULL (kswx.h:506): an all-zero alignment, which the score / identity thresholds then judge */
		r.score = r.tb = r.te = r.qb = r.qe = r.aln = r.mat = r.mis = r.ins = r.del = 0; r.n_cigar = 0; r.cigar_off = 0; recs[t] = r; return;
	}
	const uint32_t *cg = ops + r.cigar_off; const uint32_t nc = r.n_cigar;
	const unsigned long long R = row_off[t];
	int *zw = bands + 3 * R, *zb = zw + (ql + 2), *ze = zb + (ql + 2);
	int qx = 0, tx = 0;
	for(uint32_t k = 0; k < nc; k++){
		const uint32_t op = cg[k] & 0xFu; const int len = (int)(cg[k] >> 4);
		if(op == 0) for(int j = 0; j < len; j++) zw[qx++] = W;
		else if(op == 1) for(int j = 0; j < len; j++) zw[qx++] = W + len;
	}
	qx = 0;
	for(uint32_t k = 0; k < nc; k++){
		const uint32_t op = cg[k] & 0xFu; const int len = (int)(cg[k] >> 4);
		if(op == 0) qx += len;
		else if(op == 1){
			for(int j = 1; j < len && j < qx; j++) zw[qx - j] += len - j;
			qx += len - 1;
			for(int j = 1; j < len && j + qx < ql; j++) zw[qx + j] += len - j;
			qx++;
		} else {
			for(int j = 1; j < len && j < qx; j++) zw[qx - j] += len - j;
			for(int j = 1; j < len && j + qx < ql; j++) zw[qx + j] += len - j;
		}
	}
	qx = 0;
	for(uint32_t k = 0; k < nc; k++){
		const uint32_t op = cg[k] & 0xFu; const int len = (int)(cg[k] >> 4);
		if(op <= 1){
			for(int j = 0; j < len; j++){
				const int b = tx - zw[qx], e = tx + 1 + zw[qx];
				zb[qx] = b < 0? 0 : b; ze[qx] = e > tl? tl : e;
				if(op == 0) tx++;
				qx++;
			}
		} else tx += len;
	}
	{ int lim = 0; for(int i = 0; i < ql; i++){ if(zb[i] < lim) zb[i] = lim; else lim = zb[i]; } }
	int wmax = 1;
	{ int lim = tl; for(int i = ql - 1; i >= 0; i--){ if(ze[i] > lim) ze[i] = lim; else lim = ze[i]; if(ze[i] - zb[i] > wmax) wmax = ze[i] - zb[i]; } }
	const int cls = wmax <= RegCap<32, 7>::ncol? 0 : (wmax <= RegCap<CL3_NT, CL3_C>::ncol? 1 : 2);
	/* class 2: band beyond the register executors (an indel run of several hundred bases): chunked sweep with its rows in the job's scratch */
	const unsigned long long zwords = cls == 2? (unsigned long long)ql * band_row_words<REFW_NT, REFW_C>(wmax) + 3ull * (unsigned long long)(tl + 2)
		: (unsigned long long)ql * (cls == 0? 32 : CL3_NT * RegCap<CL3_NT, CL3_C>::WPT);
	const unsigned long long seq = (unsigned long long)((ql + 15) >> 4) + ((tl + 15) >> 4) + 4;
	scr_words[t] = (zwords + seq + 31) & ~31ull;
	RefJob J; J.task = t; J.ql = ql; J.tl = tl; J.band = 3 * R + (unsigned long long)(ql + 2); J.scratch = (unsigned long long)wmax /* widest row: sizes the traceback of the wide-band fallback */; J.out = out_off[t]; J.out_cap = (uint32_t)(ql + tl + 4); J.pad = (uint32_t)cls;
	jobs[cls * (size_t)nt + atomicAdd(njobs + cls, 1ULL)] = J;
}
template<int NT, int C>
__device__ void run_refine_job(const RefJob &J, const unsigned long long *scr_off, const zmo_pair_t *pairs, const AlnTask *tasks, DevReads R, const DPPar &P, const int *bands,
		uint32_t *arena, uint32_t *out_ops, zmo_record_t *recs, const BandSmem &S, uint32_t *s_seq, unsigned long long *cells, int tid){
	const AlnTask T = tasks[J.task]; const zmo_pair_t pr = pairs[T.pair_idx]; zmo_record_t r = recs[J.task];
	uint32_t *scr = arena + scr_off[J.task];
	uint32_t *z = scr; scr += (size_t)J.ql * NT * RegCap<NT, C>::WPT;
	const int qw = (J.ql + 15) >> 4, tw = (J.tl + 15) >> 4;
	uint32_t *qpk, *tpk;
	if(qw + tw + 4 <= REF_SEQW){ qpk = s_seq; tpk = s_seq + qw + 1; } else { qpk = scr; tpk = scr + qw + 1; }
	stage_packed<NT>(view_pb2(R, pr.cid, T.dir, r.qb, 1), J.ql, qpk, tid);
	stage_packed<NT>(view_pb1(R, pr.qid, r.tb, 1), J.tl, tpk, tid);
	ex_sync<NT>();
	uint32_t *cig = out_ops + J.out; DPOut o;
	reg_refine<NT, C>(S, qpk, J.ql, tpk, J.tl, bands + J.band, bands + J.band + (J.ql + 2), P, z, cig, (int)J.out_cap, o, cells, tid);
	/* walk order -> alignment order (reverse_u32list, kswx.h:656) */
	for(int a = tid; a < o.ncig / 2; a += NT){ const uint32_t x = cig[a]; cig[a] = cig[o.ncig - 1 - a]; cig[o.ncig - 1 - a] = x; }
	if(tid == 0){
		r.score = o.score; r.mat = o.mat; r.mis = o.mis; r.ins = o.ins; r.del = o.del; r.aln = o.mat + o.mis + o.ins + o.del;
		r.cigar_off = J.out; r.n_cigar = (uint32_t)o.ncig;
		recs[J.task] = r;
	}
	ex_sync<NT>();
}
__global__ void __launch_bounds__(128) k_refine_warp(const RefJob *jobs, uint32_t njobs, const unsigned long long *scr_off, const zmo_pair_t *pairs, const AlnTask *tasks, DevReads R, DPPar P,
		const int *bands, uint32_t *arena, uint32_t *out_ops, zmo_record_t *recs, unsigned long long *ctr, int ctr_work, int ctr_cells){
	__shared__ uint32_t s_seq[4][REF_SEQW];
	__shared__ int s_misc[4][16];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	BandSmem S; S.H0 = S.H1 = S.Ev = nullptr; S.cap_mask = 0; S.sred = nullptr; S.sredk = nullptr; S.smisc = s_misc[warp];
	while(1){
		uint32_t jn = 0;
		if(lane == 0) jn = (uint32_t)atomicAdd(ctr + ctr_work, 1ULL);
		jn = __shfl_sync(0xffffffffu, jn, 0);
		if(jn >= njobs) break;
		run_refine_job<32, 7>(jobs[jn], scr_off, pairs, tasks, R, P, bands, arena, out_ops, recs, S, s_seq[warp], ctr + ctr_cells, lane);
		__syncwarp();
	}
}
__global__ void __launch_bounds__(CL3_NT, 3) k_refine_cta(const RefJob *jobs, uint32_t njobs, const unsigned long long *scr_off, const zmo_pair_t *pairs, const AlnTask *tasks, DevReads R, DPPar P,
		const int *bands, uint32_t *arena, uint32_t *out_ops, zmo_record_t *recs, unsigned long long *ctr, int ctr_work, int ctr_cells){
	__shared__ uint32_t s_seq[REF_SEQW];
	__shared__ int s_red[2 * (CL3_NT / 32)];
	__shared__ long long s_redk[CL3_NT / 32];
	__shared__ int s_misc[16];
	__shared__ uint32_t s_job;
	BandSmem S; S.H0 = S.H1 = S.Ev = nullptr; S.cap_mask = 0; S.sred = s_red; S.sredk = s_redk; S.smisc = s_misc;
	const int tid = threadIdx.x;
	while(1){
		if(tid == 0) s_job = (uint32_t)atomicAdd(ctr + ctr_work, 1ULL);
		__syncthreads();
		const uint32_t jn = s_job;
		__syncthreads();
		if(jn >= njobs) break;
		run_refine_job<CL3_NT, CL3_C>(jobs[jn], scr_off, pairs, tasks, R, P, bands, arena, out_ops, recs, S, s_seq, ctr + ctr_cells, tid);
		__syncthreads();
	}
}
/* wide-band fallback (class 2 of k_refine_band): one CTA per job, band_refine (zmo_dp.cuh) with H / E rows and sequences in the job's scratch */
__global__ void __launch_bounds__(REFW_NT) k_refine_wide(const RefJob *jobs, uint32_t njobs, const unsigned long long *scr_off, const zmo_pair_t *pairs, const AlnTask *tasks, DevReads R, DPPar P,
		const int *bands, uint32_t *arena, uint32_t *out_ops, zmo_record_t *recs, unsigned long long *ctr, int ctr_work, int ctr_cells){
	__shared__ int s_red[2 * (REFW_NT / 32)];
	__shared__ int s_misc[16];
	__shared__ uint32_t s_job;
	const int tid = threadIdx.x;
	while(1){
		if(tid == 0) s_job = (uint32_t)atomicAdd(ctr + ctr_work, 1ULL);
		__syncthreads();
		const uint32_t jn = s_job;
		__syncthreads();
		if(jn >= njobs) break;
		const RefJob J = jobs[jn];
		const AlnTask T = tasks[J.task]; const zmo_pair_t pr = pairs[T.pair_idx]; zmo_record_t r = recs[J.task];
		const int wmax = (int)J.scratch;
		uint32_t *scr = arena + scr_off[J.task];
		uint32_t *z = scr; scr += (size_t)J.ql * band_row_words<REFW_NT, REFW_C>(wmax);
		BandSmem S; S.H0 = (int*)scr; S.H1 = S.H0 + (J.tl + 2); S.Ev = S.H1 + (J.tl + 2); S.cap_mask = 0; S.sred = s_red; S.sredk = nullptr; S.smisc = s_misc;
		scr += 3 * (size_t)(J.tl + 2);
		const int qw = (J.ql + 15) >> 4;
		uint32_t *qpk = scr, *tpk = scr + qw + 1;
		stage_packed<REFW_NT>(view_pb2(R, pr.cid, T.dir, r.qb, 1), J.ql, qpk, tid);
		stage_packed<REFW_NT>(view_pb1(R, pr.qid, r.tb, 1), J.tl, tpk, tid);
		__syncthreads();
		uint32_t *cig = out_ops + J.out; DPOut o;
		band_refine<REFW_NT, REFW_C>(S, qpk, J.ql, tpk, J.tl, bands + J.band, bands + J.band + (J.ql + 2), wmax, P, z, cig, (int)J.out_cap, o, ctr + ctr_cells, tid);
		for(int a = tid; a < o.ncig / 2; a += REFW_NT){ const uint32_t x = cig[a]; cig[a] = cig[o.ncig - 1 - a]; cig[o.ncig - 1 - a] = x; }
		if(tid == 0){
			r.score = o.score; r.mat = o.mat; r.mis = o.mis; r.ins = o.ins; r.del = o.del; r.aln = o.mat + o.mis + o.ins + o.del;
			r.cigar_off = J.out; r.n_cigar = (uint32_t)o.ncig;
			recs[J.task] = r;
		}
		__syncthreads();
	}
}
