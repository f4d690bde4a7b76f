/*
 * zmo_winalign.cuh -- per-window anchored alignment (fast_seeds_align_hzmo, hzm_aln.h:1247-1302): the warp-per-window executor
 * kernel and its records.  Kept in a header so that the test-only host simulation (tests/hostsim/dp_host.cpp) runs this very
 * source; included by zmo_align.cu only.
 */
#pragma once
#include "zmo_jobs.cuh"
#include "zmo_seed_core.cuh"

#define WA_C 7
#define WA_CAP 256
#define WA_SEQW 192
#define WA_WARPS 4
#define WA_ZROWS 64      /* extension problems with <= 64 rows keep their traceback in shared memory */

struct WItem { uint32_t task, win; };
struct DevReg { int score, tb, te, qb, qe, aln, mat, mis, ins, del; unsigned long long cig_off; uint32_t cig_len, kept; };
struct AlnTask { uint32_t pair_idx, dir, item_off, n_item; };
struct AlnPar { int w, ew, W, zovl; float min_id; DPPar P; };

/* views of the two reads of a task: pb1 = q forward, pb2 = c on strand dir (hzm_aln.h naming) */
__device__ __forceinline__ SeqView view_pb1(const DevReads &R, uint32_t qid, int start, int step){ SeqView v; v.w = R.words + R.woff[qid]; v.start = start; v.step = step; v.comp = 0; return v; }
__device__ __forceinline__ SeqView view_pb2(const DevReads &R, uint32_t cid, uint32_t dir, int start, int step){
	SeqView v; v.w = R.words + R.woff[cid];
	if(dir){ v.start = (int)R.len[cid] - 1 - start; v.step = -step; v.comp = 3u; } else { v.start = start; v.step = step; v.comp = 0; }
	return v;
}

__device__ __forceinline__ void cig_put(uint32_t *c, uint32_t &n, uint32_t op, uint32_t len){         /* kswx.h:39-44 */
	if(len == 0) return;
	if(n && (c[n - 1] & 0xFu) == op) c[n - 1] += len << 4; else c[n++] = (len << 4) | op;
}
__device__ __forceinline__ void cig_cat(uint32_t *c, uint32_t &n, const uint32_t *src, uint32_t len, bool reversed){   /* kswx.h:46-52 */
	if(len == 0) return;
	uint32_t i = 0; const uint32_t first = reversed? src[len - 1] : src[0];
	if(n && (c[n - 1] & 0xFu) == (first & 0xFu)){ c[n - 1] += first & 0xFFFFFFF0u; i = 1; }
	if(reversed) for(; i < len; i++) c[n++] = src[len - 1 - i];
	else for(; i < len; i++) c[n++] = src[i];
}

/* run-length alignment of one anchor pair (hz_align_hzmo, hzm_aln.h:278-314): equal base runs -> M, length difference -> I or D.  Up to cap
 * ops go to ops[] (cap + 1 slots; the reference's CIGAR list has no limit, the pipeline keeps 94 ops, far beyond any z-mer span).
 * res = {score, aln, mat, ins, del, flags (1 = run bases differ, 2 = more than cap ops), number of ops} */
#ifdef __CUDACC__
#define ZMO_NOINLINE __noinline__
#else
#define ZMO_NOINLINE __attribute__((noinline))
#endif
__device__ ZMO_NOINLINE void anchor_runlen(const SeqView a, uint32_t la, const SeqView b, uint32_t lb, const DPPar P, uint32_t *ops, uint32_t cap, int *res){
	uint32_t sa = 0, sb = 0, n2 = 0; int y_score = 0, y_aln = 0, y_mat = 0, y_ins = 0, y_del = 0, flags = 0;
	while(sa < la || sb < lb){
		const uint32_t ca = sa < la? sv_base(a, (int)sa) : 4u, cb = sb < lb? sv_base(b, (int)sb) : 5u;
		if(ca != cb){ flags |= 1; break; }
		uint32_t ea = sa + 1; while(ea < la && sv_base(a, (int)ea) == ca) ea++;
		uint32_t eb = sb + 1; while(eb < lb && sv_base(b, (int)eb) == cb) eb++;
		const uint32_t na = ea - sa, nbb = eb - sb;
		if(n2 >= cap) flags |= 2;
		if(na < nbb){ y_aln += nbb; y_mat += na; y_ins += nbb - na; y_score += na * P.M + P.I + (int)(nbb - na) * P.E; if(n2 < cap){ cig_put(ops, n2, 0, na); cig_put(ops, n2, 1, nbb - na); } }
		else if(na == nbb){ y_aln += na; y_mat += na; y_score += na * P.M; if(n2 < cap) cig_put(ops, n2, 0, na); }
		else { y_aln += na; y_mat += nbb; y_del += na - nbb; y_score += nbb * P.M + P.D + (int)(na - nbb) * P.E; if(n2 < cap){ cig_put(ops, n2, 0, nbb); cig_put(ops, n2, 2, na - nbb); } }
		sa = ea; sb = eb;
	}
	res[0] = y_score; res[1] = y_aln; res[2] = y_mat; res[3] = y_ins; res[4] = y_del; res[5] = flags; res[6] = (int)n2;
}
#define WA_ANC_OPS 8        /* ops per anchor kept in the per-warp anchor cache (an anchor with more is redone by lane 0) */
#define WA_ANC_INTS 16      /* ints per cache entry: 7 results + WA_ANC_OPS + 1 ops */

/* warp-per-window executor */
__global__ void __launch_bounds__(32 * WA_WARPS) k_window_align(const WItem *items, uint32_t nitems, const AlnTask *tasks, const zmo_pair_t *pairs,
		const DevWin *wins, const DevZPair *anchors, DevReads R, AlnPar A, uint32_t *arena, unsigned long long slab_words, int max_rows,
		uint32_t *cig_arena, const unsigned long long *item_cig_off, DevReg *regs, unsigned long long *ctr, int ctr_work, int ctr_cells,
		const uint32_t *sel, const unsigned long long *nsel){      /* sel != null: only the items sel[0 .. *nsel) (the windows k_wb_prep left to this kernel) */
	__shared__ int s_h[WA_WARPS][3 * WA_CAP];
	__shared__ uint32_t s_seq[WA_WARPS][WA_SEQW];
	__shared__ int s_misc[WA_WARPS][16];
	__shared__ uint32_t s_z[WA_WARPS][WA_ZROWS * 32];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned gw = blockIdx.x * WA_WARPS + warp;
	uint32_t *slab = arena + (unsigned long long)gw * slab_words;
	BandSmem S; S.H0 = s_h[warp]; S.H1 = S.H0 + WA_CAP; S.Ev = S.H1 + WA_CAP; S.cap_mask = WA_CAP - 1; S.sred = nullptr; S.sredk = nullptr; S.smisc = s_misc[warp];
	const DPPar P = A.P;
	while(1){
		uint32_t it = 0;
		if(lane == 0) it = (uint32_t)atomicAdd(ctr + ctr_work, 1ULL);
		it = __shfl_sync(0xffffffffu, it, 0);
		if(sel){ if(it >= *nsel) break; it = sel[it]; }
		else if(it >= nitems) break;
		const WItem I = items[it]; const AlnTask T = tasks[I.task]; const zmo_pair_t pr = pairs[T.pair_idx]; const DevWin W = wins[I.win];
		uint32_t *cig = cig_arena + item_cig_off[it]; uint32_t ncig = 0;
		int x_score = 0, x_tb = 0, x_te = 0, x_qb = 0, x_qe = 0, x_aln = 0, x_mat = 0, x_mis = 0, x_ins = 0, x_del = 0;
		/* anchor cache: the run-length alignments of the next 32 anchors of the window, one per lane, in the (otherwise idle) H/E rows of the
		 * shared-memory fallback sweep; the anchors of a window are known up front, only WHICH of them are used depends on the bridges */
		uint32_t cbase = 0; bool cvalid = false; int *const acache = s_h[warp];
		for(uint32_t ai = W.anc0; ai < W.anc1; ai++){
			const DevZPair p = anchors[ai];
			if(x_aln == 0){ x_tb = x_te = (int)p.off1; x_qb = x_qe = (int)p.off2; }
			if((int)p.off1 < x_te) continue;
			if((int)p.off2 < x_qe) continue;
			const int qlen = (int)p.off2 - x_qe, tlen = (int)p.off1 - x_te;
			const int init = x_score < 0? 0 : x_score;
			DPOut o; o.score = init; o.qe = o.te = o.mat = o.mis = o.ins = o.del = o.ncig = 0;
			uint32_t *tmpc = nullptr;
			if(qlen > 0 && tlen > 0){
				const BandDims d = band_dims(qlen, tlen, init, A.w, P);
				const int rw = band_row_words<32, WA_C>(d.ncol);      /* >= the row words of the narrower variants (32) */
				uint32_t *scr = slab;
				uint32_t *z = (d.ql <= WA_ZROWS && rw == 32)? s_z[warp] : scr; scr += (size_t)max_rows * rw;
				int *zb = (int*)scr; scr += max_rows;
				tmpc = scr; scr += 2 * (size_t)max_rows + 2 * (size_t)A.w + 16;
				const int qw = (d.ql + 15) >> 4, tw = (d.tl + 15) >> 4;
				uint32_t *qpk, *tpk;
				if(qw + tw + 2 <= WA_SEQW){ qpk = s_seq[warp]; tpk = qpk + qw; } else { qpk = scr; tpk = scr + qw; }
				scr += ((size_t)max_rows >> 3) + ((size_t)A.w >> 3) + 8;
				BandSmem S2 = S;
				if(2 * d.W + 3 > WA_CAP){ int cap = 1; while(cap < 2 * d.W + 3) cap <<= 1; S2.H0 = (int*)scr; S2.H1 = S2.H0 + cap; S2.Ev = S2.H1 + cap; S2.cap_mask = cap - 1; }
				stage_packed<32>(view_pb2(R, pr.cid, T.dir, x_qe, 1), d.ql, qpk, lane);
				stage_packed<32>(view_pb1(R, pr.qid, x_te, 1), d.tl, tpk, lane);
				__syncwarp();
				/* columns per lane chosen by band width: narrow bridges (the common case, ~50 columns) run 1-2 cells per lane
				 * instead of 7 mostly idle ones, which cuts the per-row instruction count several-fold */
				const int ccap = 2 * max_rows + 2 * A.w + 16;
				if(d.ncol <= RegCap<32, 1>::ncol) reg_extend<32, 1, 0>(S2, qpk, qlen, tpk, tlen, init, d, P, z, tmpc, ccap, o, ctr + ctr_cells, lane);
				else if(d.ncol <= RegCap<32, 2>::ncol) reg_extend<32, 2, 0>(S2, qpk, qlen, tpk, tlen, init, d, P, z, tmpc, ccap, o, ctr + ctr_cells, lane);
				else if(d.ncol <= RegCap<32, 4>::ncol) reg_extend<32, 4, 0>(S2, qpk, qlen, tpk, tlen, init, d, P, z, tmpc, ccap, o, ctr + ctr_cells, lane);
				else if(d.ncol <= RegCap<32, WA_C>::ncol) reg_extend<32, WA_C, 0>(S2, qpk, qlen, tpk, tlen, init, d, P, z, tmpc, ccap, o, ctr + ctr_cells, lane);
				else { band_extend<32, WA_C, 0>(S2, qpk, qlen, tpk, tlen, init, d, P, z, zb, tmpc, ccap, o, ctr + ctr_cells, lane); cvalid = false; }      /* may have used the rows the anchor cache lives in */
			}
			x_score = o.score;
			x_aln += o.mat + o.mis + o.ins + o.del; x_mat += o.mat; x_mis += o.mis; x_ins += o.ins; x_del += o.del;
			x_te += o.te; x_qe += o.qe;
			/* the reference merges [extension cigar + D pad + I pad] into the window cigar as ONE block
			 * (kswx_push_cigars merges only its first op with the previous last op) */
			uint32_t padD = 0, padI = 0;
			if(x_te < (int)p.off1){ padD = p.off1 - x_te; x_del += padD; x_aln += padD; x_te = p.off1; }
			if(x_qe < (int)p.off2){ padI = p.off2 - x_qe; x_ins += padI; x_aln += padI; x_qe = p.off2; }
			int ok = 1;
			if(!cvalid || ai - cbase >= 32u){
				/* refill: lane l aligns anchor ai + l */
				__syncwarp();
				cbase = ai; cvalid = true;
				const uint32_t aj = ai + (uint32_t)lane;
				if(aj < W.anc1){
					const DevZPair pj = anchors[aj];
					anchor_runlen(view_pb1(R, pr.qid, (int)pj.off1, 1), pj.len1, view_pb2(R, pr.cid, T.dir, (int)pj.off2, 1), pj.len2, P, (uint32_t*)(acache + lane * WA_ANC_INTS + 7), WA_ANC_OPS, acache + lane * WA_ANC_INTS);
				}
				__syncwarp();
			}
			if(lane == 0){
				/* block = reverse(walk-order ops) ++ D pad ++ I pad with run merging inside the block */
				const uint32_t base = ncig; uint32_t nb = ncig;   /* build the block in place after position base, merging only within the block */
				uint32_t *blk = cig + base; uint32_t bn = 0;
				for(int k = o.ncig - 1; k >= 0; k--) blk[bn++] = tmpc[k];
				if(padD){ if(bn && (blk[bn - 1] & 0xFu) == 2u) blk[bn - 1] += padD << 4; else blk[bn++] = (padD << 4) | 2u; }
				if(padI){ if(bn && (blk[bn - 1] & 0xFu) == 1u) blk[bn - 1] += padI << 4; else blk[bn++] = (padI << 4) | 1u; }
				/* now splice: merge first op of the block with the previous op if equal */
				if(bn){
					if(base && (cig[base - 1] & 0xFu) == (blk[0] & 0xFu)){
						cig[base - 1] += blk[0] & 0xFFFFFFF0u;
						for(uint32_t k = 1; k < bn; k++) cig[base + k - 1] = blk[k];
						nb = base + bn - 1;
					} else nb = base + bn;
				}
				ncig = nb;
				/* run-length alignment of the anchor itself (hzm_aln.h:278-314): from the cache, or (more than WA_ANC_OPS ops) redone here */
				const int *e = acache + (ai - cbase) * WA_ANC_INTS; int big[8]; uint32_t blk2[96]; const uint32_t *ops2 = (const uint32_t*)(e + 7);
				if(e[5] & 2){
					anchor_runlen(view_pb1(R, pr.qid, (int)p.off1, 1), p.len1, view_pb2(R, pr.cid, T.dir, (int)p.off2, 1), p.len2, P, blk2, 94u, big);
					e = big; ops2 = blk2;
				}
				if((e[5] & 1) || e[1] == 0) ok = 0;
				else {
					s_misc[warp][9] = e[0]; s_misc[warp][10] = e[1]; s_misc[warp][11] = e[2]; s_misc[warp][12] = e[3]; s_misc[warp][13] = e[4];
					cig_cat(cig, ncig, ops2, (uint32_t)e[6], false);
				}
			}
			ok = __shfl_sync(0xffffffffu, ok, 0);
			__syncwarp();
			if(!ok) break;               /* "should never happen": window truncated (hzm_aln.h:1288-1291) */
			x_score += s_misc[warp][9]; x_aln += s_misc[warp][10]; x_mat += s_misc[warp][11]; x_ins += s_misc[warp][12]; x_del += s_misc[warp][13];
			x_te += s_misc[warp][11] + s_misc[warp][13]; x_qe += s_misc[warp][11] + s_misc[warp][12];
			__syncwarp();
		}
		if(lane == 0){
			DevReg r; r.score = x_score; r.tb = x_tb; r.te = x_te; r.qb = x_qb; r.qe = x_qe; r.aln = x_aln; r.mat = x_mat; r.mis = x_mis; r.ins = x_ins; r.del = x_del;
			r.cig_off = item_cig_off[it]; r.cig_len = ncig;
			r.kept = !(x_aln * 2 < A.zovl || (float)x_mat < (float)x_aln * A.min_id);       /* wtzmo.c:1026 */
			regs[it] = r;
		}
		__syncwarp();
	}
}
