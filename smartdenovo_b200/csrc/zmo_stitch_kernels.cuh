/*
 * zmo_stitch_kernels.cuh -- bookkeeping of global_align_regs_hzmo (hzm_aln.h:1345-1486) around the DP jobs: k_plan (gap-fill and left
 * end-extension jobs from the kept window regions), k_plan2 (accumulate, right end-extension job), k_finish_size / k_finish (final
 * record and stitched CIGAR).  Kept in a header so that the test-only host simulation (tests/hostsim/dp_host.cpp) runs this very
 * source; included by zmo_align.cu only.
 */
#pragma once
#include "zmo_winalign.cuh"

struct TaskState { int ok; int first, last; int left_job, right_job; int gap0, ngap; int score, tb, te, qb, qe, aln, mat, mis, ins, del; unsigned long long cig_need; };
/* six job lists: extension classes 0..3 (warp, CTA 64/128/256), gap-fill warp (4) and CTA (5); a job's index in the
 * concatenated array equals its result index */
struct JobLists { DPJob *list[6]; unsigned long long *cnt[6]; uint32_t res_base[6]; uint32_t cap; unsigned long long *cig_cur, cig_cap, *overflow; };

__device__ inline int push_job(int cls, JobLists &L, DPJob &J, unsigned long long scratch_words){
	DPJob *list = L.list[cls]; unsigned long long *cnt = L.cnt[cls]; const uint32_t cap = L.cap, res_base = L.res_base[cls];
	const unsigned long long c0 = atomicAdd(L.cig_cur, (unsigned long long)J.cig_cap);
	const unsigned long long k = atomicAdd(cnt, 1ULL);
	if(c0 + J.cig_cap > L.cig_cap || k >= cap){ atomicAdd(L.overflow, 1ULL); return -1; }
	J.scratch = 0; J.sw32 = (uint32_t)((scratch_words + 31) >> 5); J.cig_off = c0; J.out_idx = res_base + (uint32_t)k;
	list[k] = J;
	return (int)J.out_idx;
}

__global__ void k_plan(const AlnTask *tasks, uint32_t nt, const zmo_pair_t *pairs, const DevReg *regs, DevReads R, AlnPar A, JobLists L, TaskState *ts){
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= nt) return;
	const AlnTask T = tasks[t]; const zmo_pair_t pr = pairs[T.pair_idx];
	TaskState S; memset(&S, 0, sizeof(S)); S.left_job = S.right_job = -1; S.first = S.last = -1; S.gap0 = -1;
	int prev = -1;
	for(uint32_t k = 0; k < T.n_item; k++){
		const DevReg &r = regs[T.item_off + k];
		if(!r.kept) continue;
		if(S.first < 0) S.first = (int)(T.item_off + k);
		if(prev >= 0){
			/* gap between regions prev and this one (hzm_aln.h:1395-1407) */
			const DevReg &r1 = regs[prev];
			DPJob J; memset(&J, 0, sizeof(J));
			int gq = r.qb - r1.qe, gt = r.tb - r1.te; if(gq < 0) gq = 0; if(gt < 0) gt = 0;
			const SeqView q = view_pb2(R, pr.cid, T.dir, r1.qe, 1), tv = view_pb1(R, pr.qid, r1.te, 1);
			J.q_rid = pr.cid; J.q_start = q.start; J.q_step = q.step; J.q_comp = q.comp? 1 : 0; J.qlen = gq;
			J.t_rid = pr.qid; J.t_start = tv.start; J.t_step = tv.step; J.t_comp = 0; J.tlen = gt;
			J.init = 0; J.Wp = A.w; J.Wmax = A.W; J.cig_cap = (uint32_t)(gq + gt + 4);
			{ const unsigned long long e = ((unsigned long long)gt * (unsigned long long)(gq < 2 * A.w + 1? gq : 2 * A.w + 1)) >> 8; J.est = e > 0xFFFFFFFFull? 0xFFFFFFFFu : (uint32_t)e; }
			int w = A.w; const int dl = gq > gt? gq - gt : gt - gq; while(w < dl) w <<= 1;
			const int bw = gq < 2 * w + 1? gq : 2 * w + 1;
			const bool wide = bw > 32 * 7 * 2;
			int id;
			if(wide) id = push_job(5, L, J, glb_scratch_words<256, 7>(gq, gt, 2048));
			else id = push_job(4, L, J, glb_scratch_words<32, 7>(gq, gt, 256));
			/* gap job ids of one task are not contiguous across classes: remember them in the region record slot */
			((DevReg*)regs)[T.item_off + k].kept = 2u + (uint32_t)(id < 0? 0 : id);
			S.ngap++;
		}
		prev = (int)(T.item_off + k); S.last = prev;
	}
	S.ok = S.first >= 0;
	if(S.ok){
		const DevReg &r0 = regs[S.first];
		if(r0.qb && r0.tb){
			DPJob J; memset(&J, 0, sizeof(J));
			const SeqView q = view_pb2(R, pr.cid, T.dir, r0.qb - 1, -1), tv = view_pb1(R, pr.qid, r0.tb - 1, -1);
			J.q_rid = pr.cid; J.q_start = q.start; J.q_step = q.step; J.q_comp = q.comp? 1 : 0; J.qlen = r0.qb;
			J.t_rid = pr.qid; J.t_start = tv.start; J.t_step = tv.step; J.t_comp = 0; J.tlen = r0.tb;
			J.init = r0.score + 100 * A.P.M; J.Wp = -A.ew; J.cig_cap = (uint32_t)(r0.qb + r0.tb + 4);
			const int init = J.init < 0? 0 : J.init;
			const BandDims d = band_dims(J.qlen, J.tlen, init, J.Wp, A.P);
			{ const unsigned long long e = ((unsigned long long)d.ql * (unsigned long long)d.ncol) >> 8; J.est = e > 0xFFFFFFFFull? 0xFFFFFFFFu : (uint32_t)e; }
			{ const int cls = ext_class(d.ncol); S.left_job = push_job(cls, L, J, ext_scratch_words_cls(d, cls)); }
		}
	}
	ts[t] = S;
}

/* accumulate left extension + regions + gaps (hzm_aln.h:1357-1450) */
__device__ inline void accumulate(const AlnTask &T, const DevReg *regs, const DPRes *res, const AlnPar &A, TaskState &S){
	const DevReg &r0 = regs[S.first];
	S.score = r0.score; S.tb = r0.tb; S.te = r0.te; S.qb = r0.qb; S.qe = r0.qe; S.aln = r0.aln; S.mat = r0.mat; S.mis = r0.mis; S.ins = r0.ins; S.del = r0.del;
	unsigned long long need = 0;
	if(S.left_job >= 0){
		const DPRes &y = res[S.left_job];
		S.score = y.score - 100 * A.P.M;
		S.aln += y.mat + y.mis + y.ins + y.del; S.mat += y.mat; S.mis += y.mis; S.ins += y.ins; S.del += y.del;
		S.qb -= y.qe; S.tb -= y.te;
		need += y.ncig;
	}
	need += r0.cig_len;
	for(uint32_t k = (uint32_t)S.first + 1 - T.item_off; k < T.n_item; k++){
		const DevReg &r = regs[T.item_off + k];
		if(r.kept < 2u) continue;
		const DPRes &g = res[r.kept - 2u];
		S.score += g.score;
		S.aln += g.mat + g.mis + g.ins + g.del; S.mat += g.mat; S.mis += g.mis; S.ins += g.ins; S.del += g.del;
		S.score += r.score; S.aln += r.aln; S.mat += r.mat; S.mis += r.mis; S.ins += r.ins; S.del += r.del;
		S.qe = r.qe; S.te = r.te;
		need += g.ncig + r.cig_len;
	}
	S.cig_need = need;
}

__global__ void k_plan2(const AlnTask *tasks, uint32_t nt, const zmo_pair_t *pairs, const DevReg *regs, const DPRes *res, DevReads R, AlnPar A, JobLists L, TaskState *ts){
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= nt) return;
	TaskState S = ts[t];
	if(!S.ok) return;
	const AlnTask T = tasks[t]; const zmo_pair_t pr = pairs[T.pair_idx];
	accumulate(T, regs, res, A, S);
	const int len1 = (int)R.len[pr.qid], len2 = (int)R.len[pr.cid];
	S.right_job = -1;
	if(S.te < len1 && S.qe < len2){
		DPJob J; memset(&J, 0, sizeof(J));
		const SeqView q = view_pb2(R, pr.cid, T.dir, S.qe, 1), tv = view_pb1(R, pr.qid, S.te, 1);
		J.q_rid = pr.cid; J.q_start = q.start; J.q_step = q.step; J.q_comp = q.comp? 1 : 0; J.qlen = len2 - S.qe;
		J.t_rid = pr.qid; J.t_start = tv.start; J.t_step = tv.step; J.t_comp = 0; J.tlen = len1 - S.te;
		J.init = S.score; J.Wp = -A.ew; J.cig_cap = (uint32_t)(J.qlen + J.tlen + 4);
		const int init = J.init < 0? 0 : J.init;
		const BandDims d = band_dims(J.qlen, J.tlen, init, J.Wp, A.P);
		{ const unsigned long long e = ((unsigned long long)d.ql * (unsigned long long)d.ncol) >> 8; J.est = e > 0xFFFFFFFFull? 0xFFFFFFFFu : (uint32_t)e; }
		{ const int cls = ext_class(d.ncol); S.right_job = push_job(cls, L, J, ext_scratch_words_cls(d, cls)); }
	}
	ts[t] = S;
}

__global__ void k_finish_size(uint32_t nt, const DPRes *res, TaskState *ts, unsigned long long *need){
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= nt) return;
	const TaskState &S = ts[t];
	unsigned long long n = 0;
	if(S.ok){ n = S.cig_need; if(S.right_job >= 0) n += res[S.right_job].ncig; }
	need[t] = n;
}

/* Final record and stitched CIGAR of a task (global_align_regs_hzmo's bookkeeping, hzm_aln.h:1345-1486), one WARP per task: [left extension] +
 * region 0 + sum([gap] + region i) + [right extension]; gap and extension CIGARs are stored in walk order.  The CIGAR segments of a task are copied by all
 * 32 lanes (coalesced) instead of one thread walking thousands of ops; the block-merge rule of kswx_push_cigars (kswx.h:46-52: only the
 * first op of an appended block may merge with the previous last op) is applied by lane 0 between the copies.  (A one-thread-per-task version
 * took 15.2 ms per quarter shard of cfg2 against 0.93 ms for this one: profiles/r02_experiments_decided.md.) */
__device__ __forceinline__ void cig_cat_warp(uint32_t *c, uint32_t &n, const uint32_t *src, uint32_t len, bool reversed, int lane){
	if(len == 0) return;                                 /* uniform: every lane sees the same arguments */
	const uint32_t first = reversed? src[len - 1] : src[0];
	const uint32_t last = n? c[n - 1] : 0xFFFFFFFFu;     /* written by the previous call, ordered by its trailing __syncwarp() */
	__syncwarp();                                        /* everybody has read c[n-1] before lane 0 rewrites it */
	uint32_t i0 = 0;
	if(n && (last & 0xFu) == (first & 0xFu)){ if(lane == 0) c[n - 1] = last + (first & 0xFFFFFFF0u); i0 = 1; }
	if(reversed) for(uint32_t i = i0 + (uint32_t)lane; i < len; i += 32) c[n + i - i0] = src[len - 1 - i];
	else for(uint32_t i = i0 + (uint32_t)lane; i < len; i += 32) c[n + i - i0] = src[i];
	n += len - i0;
	__syncwarp();
}
__global__ void __launch_bounds__(256) k_finish_warp(const AlnTask *tasks, uint32_t nt, const DevReg *regs, const DPRes *res, const DPJob *jobs_all,
		const uint32_t *cig_arena, AlnPar A, const TaskState *ts, const unsigned long long *out_off, uint32_t *out_cig, zmo_record_t *recs){
	const uint32_t t = (uint32_t)(((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); const int lane = threadIdx.x & 31;
	if(t >= nt) return;                                  /* whole warps leave together */
	TaskState S = ts[t]; zmo_record_t rec; memset(&rec, 0, sizeof(rec));
	if(!S.ok){ if(lane == 0) recs[t] = rec; return; }
	const AlnTask T = tasks[t];
	uint32_t *dst = out_cig + out_off[t]; uint32_t n = 0;
	if(S.left_job >= 0){ const DPRes &y = res[S.left_job]; const DPJob &J = jobs_all[S.left_job]; cig_cat_warp(dst, n, cig_arena + J.cig_off, (uint32_t)y.ncig, false, lane); }
	{ const DevReg &r0 = regs[S.first]; cig_cat_warp(dst, n, cig_arena + r0.cig_off, r0.cig_len, false, lane); }
	for(uint32_t k = (uint32_t)S.first + 1 - T.item_off; k < T.n_item; k++){
		const DevReg &r = regs[T.item_off + k];
		if(r.kept < 2u) continue;
		const DPRes &g = res[r.kept - 2u]; const DPJob &J = jobs_all[r.kept - 2u];
		cig_cat_warp(dst, n, cig_arena + J.cig_off, (uint32_t)g.ncig, true, lane);
		cig_cat_warp(dst, n, cig_arena + r.cig_off, r.cig_len, false, lane);
	}
	if(S.right_job >= 0){
		const DPRes &y = res[S.right_job]; const DPJob &J = jobs_all[S.right_job];
		S.score = y.score;
		S.aln += y.mat + y.mis + y.ins + y.del; S.mat += y.mat; S.mis += y.mis; S.ins += y.ins; S.del += y.del;
		S.qe += y.qe; S.te += y.te;
		cig_cat_warp(dst, n, cig_arena + J.cig_off, (uint32_t)y.ncig, true, lane);
	}
	if(lane == 0){
		rec.ok = 1; rec.score = S.score; rec.tb = S.tb; rec.te = S.te; rec.qb = S.qb; rec.qe = S.qe; rec.aln = S.aln; rec.mat = S.mat; rec.mis = S.mis; rec.ins = S.ins; rec.del = S.del;
		rec.cigar_off = out_off[t]; rec.n_cigar = n;
		recs[t] = rec;
	}
}
