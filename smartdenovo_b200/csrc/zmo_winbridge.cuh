/*
 * zmo_winbridge.cuh -- per-window anchored alignment (fast_seeds_align_hzmo, hzm_aln.h:1247-1302) with the BRIDGE as the unit of DP work.
 *
 * A window is a chain [bridge, anchor, bridge, anchor, ...]: between two consecutive anchors the reference runs the fixed-band extension
 * kswx_extend_align_core (kswx.h:234-335) seeded with the running score of the window, so the bridges of a window look sequential.  They
 * are not, up to a shift:
 *   - WHICH anchors are used and where every bridge starts and ends depends on the anchors alone (an anchor is taken iff it starts at or
 *     after the end of the last taken one on both reads; a taken anchor ends at off + mat + del / off + mat + ins), not on any DP result;
 *   - the band half-width min(w, max_gap(init), max(ql, tl)) does not depend on init once max_gap(0) >= min(w, max(ql, tl)) (always, for the
 *     default scores and -w);
 *   - every H / E / F value of the bridge is its value for init = 0 plus init, and every traceback bit is the same, as long as the values that
 *     derive from the sequences stay above the -10,000 sentinels (guarded per bridge in k_wb_prep), because a sentinel then loses every
 *     comparison it takes part in, for any init >= 0.
 * So:  k_wb_prep   (a warp per window)  run-length alignment of every anchor of the window (hzm_aln.h:278-314, a lane per anchor), then the
 *                  chain: which anchors are taken, the geometry of the bridge in front of each;
 *      k_wb_sweep  (a lane per bridge; the bridges of ALL windows of the wave in one list, longest first, 32 at a time per warp) row sweep for
 *                  init = 0: per row the maximum, its last column and the value of the row's last column, plus the 4-bit traceback;
 *      k_wb_ends   (a lane per window)  the reference's per-row rules with the real init: best cell, end-of-sequence cell, stop at the first
 *                  row that neither improves nor is positive -> score and end cell of every bridge;
 *      k_wb_walk   (a lane per bridge)  traceback walk from the end cell (kswx.h:311-330);
 *      k_wb_stitch (a warp per window)  pads, CIGAR blocks, anchors -> DevReg + window CIGAR, exactly what k_window_align leaves.
 * The sweep is where the cells are.  All lanes of a warp are in the same phase there (an 8-column group of some row of some bridge) --
 * what the lane-per-WINDOW kernel tried earlier in this round lacked (8.5 of 32 lanes active, profiles/r02_rejected_lane_per_window.md).
 * Windows the argument does not cover (band depends on init, int16 range, an anchor with more CIGAR ops than 2 per z-mer base) are listed by
 * k_wb_prep and go through k_window_align (zmo_winalign.cuh) unchanged.
 */
#pragma once
#include "zmo_winalign.cuh"

#ifndef ZMO_DYN_SMEM
#define ZMO_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

#define WB_NT 64            /* lanes per CTA of the sweep */
#define WB_KEY_SH 10        /* row arg-max key = h * 1024 + (column - band start + 1) */
#define WB_KEY_MIN (-0x40000000)
#define WB_MAX_W 91         /* ring of wb_cap(w) packed cells x WB_NT lanes within 48 KB of shared memory */

/* one per anchor of a window, in anchor order */
struct WBStep {
	unsigned long long qoff, coff;      /* word offsets of the two reads */
	int32_t qlen, tlen;                 /* bridge in front of the anchor: rows (on c), columns (on q); 0 x 0 = none */
	uint32_t x_te, x_qe;                /* where the bridge starts: on q (pb1), on c's strand (pb2) */
	uint32_t flags, W, clen, dir;       /* flags: 1 taken, 2 window ends after this bridge, 4 bridge has cells, 8 anchor's run bases differ, 16 anchor needs the sequential path */
	int32_t a_score; uint32_t a_nops; uint16_t a_aln, a_mat, a_ins, a_del;      /* run-length alignment of the anchor */
	int32_t o_score, end_i, end_j;      /* k_wb_ends */
	int32_t o_mat, o_mis, o_ins, o_del; uint32_t o_ncig;      /* k_wb_walk */
	/* k_wb_sweep: outcome of the row rules relative to init, valid when no row stops the sweep, i.e. s_low + init > 0 (kswx.h:281-305):
	 * best cell (value s_best + init), end-of-sequence cell (value s_g + init), lowest maximum of a row that did not improve, cells swept */
	int32_t s_best, s_bi, s_bj, s_g, s_gi, s_gj, s_low; uint32_t s_cells;
	uint32_t pad_[2];
};

/* ring slots per lane: the 2w+1 columns of a row + the column entering next + 7 so that the dead cells an 8-aligned group touches on either side of the
 * band (written, not predicated) never alias a live column */
__host__ __device__ __forceinline__ int wb_cap(int w){ return (2 * w + 8 + 7) & ~7; }
__host__ __device__ __forceinline__ int wb_row_words(int w){ return (((2 * w + 1) >> 3) + 2 + 3) & ~3; }   /* traceback words per row (multiple of 4: the row stats behind them are int4) */
/* scratch of one bridge: [traceback ql x rw | row stats ql x int4 | walk ops ql + tl + 4], a multiple of 4 words */
__host__ __device__ __forceinline__ unsigned long long wb_scr_words(int ql, int tl, int rw){ return ((unsigned long long)ql * (unsigned long long)(rw + 4 + 1) + (unsigned long long)tl + 4ull + 3ull) & ~3ull; }

__device__ __forceinline__ uint32_t wb_b1(const uint32_t *w, int p){ return (__ldg(w + (p >> 4)) >> (((~p) & 15) << 1)) & 3u; }
__device__ __forceinline__ uint32_t wb_b2(const uint32_t *w, int clen, uint32_t dir, int p){
	if(dir){ const int r = clen - 1 - p; return ((__ldg(w + (r >> 4)) >> (((~r) & 15) << 1)) & 3u) ^ 3u; }
	return wb_b1(w, p);
}

/* ---- 1. a warp per window: anchors, then the chain -------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) k_wb_prep(const WItem *items, uint32_t nitems, const AlnTask *tasks, const zmo_pair_t *pairs, const DevWin *wins, const DevZPair *anchors,
		DevReads R, AlnPar A, const unsigned long long *item_step_off, int rw, WBStep *steps, uint32_t *aops, int acap, unsigned long long *scr_words, uint32_t *keys, uint32_t *order,
		uint8_t *item_seq, uint32_t *fb_items, unsigned long long *fb_count){
	const uint32_t it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
	if(it >= nitems) return;
	const WItem I = items[it]; const AlnTask T = tasks[I.task]; const zmo_pair_t pr = pairs[T.pair_idx]; const DevWin Wn = wins[I.win];
	const DPPar P = A.P;
	const unsigned long long s0 = item_step_off[it]; const uint32_t na = Wn.anc1 - Wn.anc0;
	for(uint32_t a = (uint32_t)lane; a < na; a += 32u){
		const DevZPair p = anchors[Wn.anc0 + a];
		WBStep *S = steps + s0 + a; int res[8]; uint32_t fl = 0;
		anchor_runlen(view_pb1(R, pr.qid, (int)p.off1, 1), p.len1, view_pb2(R, pr.cid, T.dir, (int)p.off2, 1), p.len2, P, aops + (s0 + a) * (unsigned long long)(acap + 1), (uint32_t)acap, res);
		if((res[5] & 1) || res[1] == 0) fl |= 8u;
		else if((res[5] & 2) || res[1] > 65535) fl |= 16u;
		S->a_score = res[0]; S->a_nops = (uint32_t)res[6]; S->a_aln = (uint16_t)res[1]; S->a_mat = (uint16_t)res[2]; S->a_ins = (uint16_t)res[3]; S->a_del = (uint16_t)res[4];
		S->flags = fl;
	}
	__syncwarp();
	if(lane) return;
	const unsigned long long qoff = R.woff[pr.qid], coff = R.woff[pr.cid]; const uint32_t clen = R.len[pr.cid];
	const int pen = (P.X < P.E? -P.X : -P.E) + 1;
	int x_te = 0, x_qe = 0; bool started = false, seq = false, ended = false;
	for(uint32_t a = 0; a < na && !seq; a++){
		WBStep *S = steps + s0 + a; const unsigned long long si = s0 + a;
		uint32_t fl = 0; unsigned long long sw = 0; uint32_t key = 0;
		if(!ended){
			const DevZPair p = anchors[Wn.anc0 + a];
			if(!started){ x_te = (int)p.off1; x_qe = (int)p.off2; }
			if((int)p.off1 >= x_te && (int)p.off2 >= x_qe){
				const uint32_t afl = S->flags;
				const int qlen = (int)p.off2 - x_qe, tlen = (int)p.off1 - x_te;
				fl = 1u; S->x_te = (uint32_t)x_te; S->x_qe = (uint32_t)x_qe; S->qlen = qlen; S->tlen = tlen; S->qoff = qoff; S->coff = coff; S->clen = clen; S->dir = T.dir; S->W = 0;
				if(qlen > 0 && tlen > 0){
					/* the band must not depend on the running score: max_gap(init = 0) >= min(w, longer side) (kswx.h:244-250) */
					const int mx = (qlen < tlen? qlen : tlen) * P.M + (-P.T);
					int max_gap = (mx + (P.I > P.D? P.I : P.D)) / (-P.E) + 1; if(max_gap < 1) max_gap = 1;
					const BandDims d = band_dims(qlen, tlen, 0, A.w, P);
					/* int16 cells and the sentinel argument: values from the sequences stay in (-9,800, 30,000), sentinel chains above -32,000 */
					const long long lo = (long long)(P.I < P.D? P.I : P.D) * 3 + (long long)P.E * (d.W + 3) + (long long)P.X * (d.ql + 1);
					const int mxl = qlen > tlen? qlen : tlen;      /* W(init) = min(w, max_gap(init), mxl) with max_gap(init) >= max_gap(0) */
					if(max_gap < (A.w < mxl? A.w : mxl) || A.w < 1 || lo < -9800 || (long long)d.ql * P.M > 30000 || 10200 + (long long)d.ql * pen > 32000 || wb_row_words(d.W) > rw) seq = true;
					else {
						fl |= 4u; S->W = (uint32_t)d.W; sw = wb_scr_words(d.ql, d.tl, rw);
						const unsigned long long itn = (unsigned long long)d.ql * (unsigned long long)(((d.ncol + 7) >> 3) + 1);
						key = itn > 0xFFFFFFFFull? 0xFFFFFFFFu : (uint32_t)itn;
					}
				}
				if(afl & 16u) seq = true;
				if(afl & 8u){ fl |= 2u; ended = true; }
				else { x_te = (int)p.off1 + S->a_mat + S->a_del; x_qe = (int)p.off2 + S->a_mat + S->a_ins; started = true; }
			}
		}
		S->flags = fl; scr_words[si] = sw; keys[si] = key; order[si] = (uint32_t)si;
	}
	item_seq[it] = seq? 1 : 0;
	if(seq){
		/* sequential path: none of the window's bridges is swept here */
		for(uint32_t a = 0; a < na; a++){ const unsigned long long si = s0 + a; steps[si].flags = 0; scr_words[si] = 0; keys[si] = 0; order[si] = (uint32_t)si; }
		fb_items[atomicAdd(fb_count, 1ULL)] = it;
	}
}

/* ---- 2. a lane per bridge: row sweep for init = 0 --------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(WB_NT) k_wb_sweep(WBStep *steps, const uint32_t *order, const uint32_t *skeys, uint32_t nsteps, const unsigned long long *scr_off,
		unsigned long long scr_cap, const uint32_t *words, DPPar P, uint32_t *arena, int cap, int rw, unsigned long long *work, unsigned long long *overflow){
	ZMO_DYN_SMEM(wb_raw);
	uint4 *const ring4 = (uint4*)wb_raw; uint32_t *const ring = (uint32_t*)wb_raw;     /* slot s of lane t: word ((s >> 2) * WB_NT + t) * 4 + (s & 3) = H (low 16 bits) | E (high 16 bits) */
	const int t = threadIdx.x, lane = t & 31;
#define WB_SLOT(s) ring[(((((s) >> 2) * WB_NT) + t) << 2) + ((s) & 3)]
	const int IE = P.I + P.E, DE = P.D + P.E, E = P.E, cM = P.M, cX = P.X;
	const uint32_t negpk = ((uint32_t)(uint16_t)(int16_t)ZMO_NEG) | ((uint32_t)(uint16_t)(int16_t)ZMO_NEG << 16);
	for(;;){
		unsigned long long base = 0;
		if(lane == 0) base = atomicAdd(work, 32ULL);
		base = __shfl_sync(0xffffffffu, base, 0);
		if(base >= nsteps) break;
		const unsigned long long k = base + (unsigned)lane;
		bool busy = k < nsteps && skeys[k] != 0;
		if(!__any_sync(0xffffffffu, busy)) break;        /* longest first: nothing with cells beyond this point */
		const uint32_t *qwp = words, *cwp = words; int clen = 0; uint32_t dir = 0;
		int x_te = 0, x_qe = 0, ql = 0, tl = 0, W = 0, i = 0, jb = 0, je = 0, j8 = 0, slot8 = 0, se = 0, hd = 0, f = 0, key = 0, hl = 0, qnext = 0, qvalid = 0;
		unsigned long long qv = 0; uint32_t rmask = 0; uint32_t *zrow = arena; int4 *stat = nullptr;
		int tlen = 0, qlen = 0, sb = 0, sbi = -1, sbj = -1, sg = WB_KEY_MIN, sgi = -1, sgj = -1, slow = 0x7FFFFFFF; uint32_t scells = 0; WBStep *Sw = steps;
#define WB_ROW_BEGIN() do { \
			jb = i > W? i - W : 0; je = i + W + 1 < tl? i + W + 1 : tl; \
			j8 = jb & ~7; slot8 = j8 % cap; \
			if(jb == 0) hd = i == 0? 0 : P.I + E * i; \
			else { int s_ = slot8 + (jb & 7) - 1; if(s_ < 0) s_ += cap; hd = (int)(short)(WB_SLOT(s_) & 0xFFFFu); } \
			if(i > 0 && i + W < tl) WB_SLOT(se) = negpk;      /* column i + W enters the band: nothing above it */ \
			se = se + 1 == cap? 0 : se + 1; \
			f = ZMO_NEG; key = WB_KEY_MIN; \
			rmask = 0x5555u * wb_b2(cwp, clen, dir, x_qe + i); \
			{ const int tp = x_te + j8, wi = tp >> 4, sh = (tp & 15) << 1; \
			  qv = ((((unsigned long long)__ldg(qwp + wi)) << 32) | __ldg(qwp + wi + 1)) << sh; qvalid = 64 - sh; qnext = wi + 2; } \
		} while(0)
		if(busy){
			const uint32_t si = order[k];
			WBStep *S = steps + si; Sw = S; qlen = S->qlen; tlen = S->tlen;
			const unsigned long long so = scr_off[si];
			qwp = words + S->qoff; cwp = words + S->coff; clen = (int)S->clen; dir = S->dir; x_te = (int)S->x_te; x_qe = (int)S->x_qe; W = (int)S->W;
			const BandDims d = band_dims(S->qlen, S->tlen, 0, -W, P);       /* the clamps of kswx.h:251-258 with the half-width k_wb_prep decided */
			ql = d.ql; tl = d.tl;
			if(so + wb_scr_words(ql, tl, rw) > scr_cap){ atomicAdd(overflow, 1ULL); busy = false; }
			else {
				zrow = arena + so; stat = (int4*)(zrow + (size_t)ql * rw);
				const int je0 = tl < W + 1? tl : W + 1;
				for(int j = 0; j < je0; j++) WB_SLOT(j) = ((uint32_t)(uint16_t)(int16_t)(P.D + E * (j + 1))) | ((uint32_t)(uint16_t)(int16_t)ZMO_NEG << 16);     /* row -1 (kswx.h:263-270), init = 0 */
				i = 0; se = W;
				WB_ROW_BEGIN();
			}
		}
		while(busy){
			/* one 8-aligned column group of row i; every cell predicated, one code path for all lanes */
			const uint32_t x = (uint32_t)(qv >> 48) ^ rmask;       /* 2-bit field c (from the top) is 0 iff column j8 + c matches the row base */
			qv <<= 16; qvalid -= 16;
			if(qvalid < 16){ qv |= ((unsigned long long)__ldg(qwp + qnext)) << (32 - qvalid); qnext++; qvalid += 32; }
			uint4 *const sp = ring4 + (slot8 >> 2) * WB_NT + t;
			uint4 v0 = sp[0], v1 = sp[WB_NT];
			uint32_t wv[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
			uint32_t zw = 0;
			/* cells left of the band (first group of a row) must leave hd / f alone, cells right of it (last group) the row's last value; the ring slots
			 * and traceback nibbles of both are dead (those columns never come back into the band / the walk never reads them), so they are just written */
			const int jr = j8 - jb + 1, lo = jb - j8, hi = je - j8;
			#pragma unroll
			for(int c = 0; c < 8; c++){
				const bool ge = c >= lo, lt = c < hi;
				const uint32_t w = wv[c];
				const int hup = (int)(short)(w & 0xFFFFu), e0 = (int)w >> 16;
				const int m = hd + (((x >> (14 - 2 * c)) & 3u)? cX : cM);
				uint32_t dd = m >= e0? 0u : 1u;
				int h = m >= e0? m : e0;
				dd = h < f? 2u : dd; h = h < f? f : h;
				const int kk = h * (1 << WB_KEY_SH) + (jr + c);
				const int t1 = m + IE, e1 = e0 + E; dd |= e1 > t1? 4u : 0u; const int e = e1 > t1? e1 : t1;
				const int t2 = m + DE, f1 = f + E; dd |= f1 > t2? 8u : 0u; const int fn = f1 > t2? f1 : t2;
				const int kc = (ge && lt)? kk : WB_KEY_MIN;
				key = kc > key? kc : key;
				zw |= dd << (4 * c);
				wv[c] = ((uint32_t)h & 0xFFFFu) | ((uint32_t)e << 16);
				hd = ge? hup : hd; f = ge? fn : f;
			}
			v0.x = wv[0]; v0.y = wv[1]; v0.z = wv[2]; v0.w = wv[3]; v1.x = wv[4]; v1.y = wv[5]; v1.z = wv[6]; v1.w = wv[7];
			sp[0] = v0; sp[WB_NT] = v1;
			zrow[(size_t)i * rw + ((j8 >> 3) - (jb >> 3))] = zw;
			j8 += 8; slot8 += 8; if(slot8 == cap) slot8 = 0;
			if(j8 >= je){
				/* row statistics for k_wb_ends: maximum (no floor at 0), its LAST column (kswx.h:284-285), value of the row's last column */
				const int Mx = key >> WB_KEY_SH, arg = jb + (key & ((1 << WB_KEY_SH) - 1)) - 1;
				{ const int sl = (je - 1) % cap; hl = (int)(short)(WB_SLOT(sl) & 0xFFFFu); }      /* H of the row's last column */
				stat[i] = make_int4(Mx, arg, hl, 0);
				/* the row rules relative to init, assuming no row stops the sweep (k_wb_ends checks that with s_low and replays the rows otherwise) */
				scells += (uint32_t)(je - jb);
				if(je == tlen && hl > sg){ sg = hl; sgi = i; sgj = je - 1; }
				if(i + 1 == qlen && Mx > sg){ sg = Mx; sgi = i; sgj = arg; }
				if(Mx > sb){ sb = Mx; sbi = i; sbj = arg; } else if(Mx < slow) slow = Mx;
				i++;
				if(i >= ql){
					busy = false;
					Sw->s_best = sb; Sw->s_bi = sbi; Sw->s_bj = sbj; Sw->s_g = sg; Sw->s_gi = sgi; Sw->s_gj = sgj; Sw->s_low = slow; Sw->s_cells = scells;
				} else WB_ROW_BEGIN();
			}
		}
		__syncwarp();
	}
#undef WB_ROW_BEGIN
#undef WB_SLOT
}

/* ---- 3. a lane per window: score and end cell of every bridge with the real init (kswx.h:281-305) ---------------------------------------------- */
__global__ void k_wb_ends(uint32_t nitems, const WItem *items, const DevWin *wins, AlnPar A, const unsigned long long *item_step_off, const uint8_t *item_seq,
		WBStep *steps, const unsigned long long *scr_off, const uint32_t *arena, int rw, const unsigned long long *overflow, unsigned long long *ctr, int ctr_cells){
	const uint32_t it = blockIdx.x * blockDim.x + threadIdx.x;
	if(it >= nitems || item_seq[it] || *overflow) return;
	const DevWin Wn = wins[items[it].win]; const DPPar P = A.P;
	const unsigned long long s0 = item_step_off[it]; const uint32_t na = Wn.anc1 - Wn.anc0;
	int x_score = 0; unsigned long long cells = 0;
	for(uint32_t a = 0; a < na; a++){
		WBStep *S = steps + s0 + a; const uint32_t fl = S->flags;
		if(!(fl & 1u)) continue;
		const int init = x_score < 0? 0 : x_score;
		int o_score = init, ii = -1, jj = -1;
		if(fl & 4u){
			const int W = (int)S->W, qlen = S->qlen, tlen = S->tlen; const BandDims d = band_dims(qlen, tlen, 0, -W, P); const int ql = d.ql, tl = d.tl;
			const int4 *stat = (const int4*)(arena + scr_off[s0 + a] + (size_t)ql * rw);
			int best = init, bi = -1, bj = -1, gbest = 0, gi = -1, gj = -1;
			if(S->s_low == 0x7FFFFFFF || S->s_low + init > 0){
				/* no row stops the sweep: what k_wb_sweep found, shifted by init (an end-of-sequence cell counts only above 0) */
				best = S->s_best + init; bi = S->s_bi; bj = S->s_bj;
				if(S->s_gi >= 0 && S->s_g + init > 0){ gbest = S->s_g + init; gi = S->s_gi; gj = S->s_gj; }
				cells += S->s_cells;
			} else
			for(int i = 0; i < ql; i++){
				const int4 st = stat[i];
				const int Mx = st.x + init, rowmax = Mx >= 0? Mx : 0, rowarg = Mx >= 0? st.y : -1, hlast = st.z + init;
				const int jb = i > W? i - W : 0, je = i + W + 1 < tl? i + W + 1 : tl;
				cells += (unsigned long long)(je - jb);
				if(je == tlen && gbest < hlast){ gbest = hlast; gi = i; gj = je - 1; }
				if(i + 1 == qlen && gbest < rowmax){ gbest = rowmax; gi = i; gj = rowarg; }
				if(rowmax > best){ best = rowmax; bi = i; bj = rowarg; } else if(rowmax <= 0) break;
			}
			if(gbest > 0 && gbest >= best + P.T){ o_score = gbest; ii = gi; jj = gj; } else { o_score = best; ii = bi; jj = bj; }
		}
		S->o_score = o_score; S->end_i = ii; S->end_j = jj;
		x_score = o_score;
		if(fl & 2u) break;
		x_score += S->a_score;
	}
	if(cells) atomicAdd(ctr + ctr_cells, cells);
}

/* ---- 4. a lane per bridge: traceback walk (kswx.h:311-330), ops in WALK order ------------------------------------------------------------------ */
__global__ void k_wb_walk(WBStep *steps, const uint32_t *order, const uint32_t *skeys, uint32_t nsteps, const unsigned long long *scr_off, const uint32_t *words, DPPar P, uint32_t *arena, int rw, const unsigned long long *overflow){
	/* in the order of the sweep (longest first): the lanes of a warp walk bridges of similar length, and the bridges without cells (sort key 0) are a
	 * contiguous tail that returns at once */
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if(k >= nsteps || *overflow || skeys[k] == 0) return;
	const uint32_t si = order[k];
	WBStep *S = steps + si;
	if((S->flags & 5u) != 5u) return;
	const int W = (int)S->W; const BandDims d = band_dims(S->qlen, S->tlen, 0, -W, P); const int ql = d.ql;
	const uint32_t *z = arena + scr_off[si]; uint32_t *ops = arena + scr_off[si] + (size_t)ql * (rw + 4);
	const uint32_t *qwp = words + S->qoff, *cwp = words + S->coff; const int clen = (int)S->clen; const uint32_t dir = S->dir; const int x_te = (int)S->x_te, x_qe = (int)S->x_qe;
	int ii = S->end_i, jj = S->end_j, sw = 0, mat = 0, mis = 0, ins = 0, del = 0; uint32_t n = 0, cur_op = 0xFu, cur_len = 0;
	while(ii >= 0 && jj >= 0){
		const uint32_t nib = (z[(size_t)ii * rw + ((jj >> 3) - ((ii > W? ii - W : 0) >> 3))] >> ((jj & 7) << 2)) & 0xFu;
		if(sw == 0) sw = (int)(nib & 3u); else if(sw == 1) sw = (nib & 4u)? 1 : 0; else sw = (nib & 8u)? 2 : 0;
		if(sw == 0){ if(wb_b2(cwp, clen, dir, x_qe + ii) == wb_b1(qwp, x_te + jj)) mat++; else mis++; ii--; jj--; }
		else if(sw == 1){ ii--; ins++; }
		else { jj--; del++; }
		if((uint32_t)sw == cur_op) cur_len++;
		else { if(cur_len) ops[n++] = (cur_len << 4) | cur_op; cur_op = (uint32_t)sw; cur_len = 1; }
	}
	if(ii >= 0){ ins += ii + 1; if(cur_op == 1u) cur_len += (uint32_t)(ii + 1); else { if(cur_len) ops[n++] = (cur_len << 4) | cur_op; cur_op = 1u; cur_len = (uint32_t)(ii + 1); } }
	if(jj >= 0){ del += jj + 1; if(cur_op == 2u) cur_len += (uint32_t)(jj + 1); else { if(cur_len) ops[n++] = (cur_len << 4) | cur_op; cur_op = 2u; cur_len = (uint32_t)(jj + 1); } }
	if(cur_len) ops[n++] = (cur_len << 4) | cur_op;
	S->o_mat = mat; S->o_mis = mis; S->o_ins = ins; S->o_del = del; S->o_ncig = n;
}

/* ---- 5. a warp per window: pads, CIGAR blocks, anchors (hzm_aln.h:1264-1297) ------------------------------------------------------------------- */
/* One block joins the window CIGAR the way kswx_push_cigars does it (kswx.h:46-52): only its FIRST op may merge with the op in front of it.  The block is
 * bulk[0 .. nb) (read backwards when rev) followed by the scalars tail[0 .. nt); lastv mirrors cig[ncig - 1] so that nothing is read back from memory.
 * All arguments are warp-uniform; the lanes share the copy. */
__device__ __forceinline__ void wb_push_block(uint32_t *cig, uint32_t &ncig, uint32_t &lastv, const uint32_t *bulk, uint32_t nb, bool rev, const uint32_t *tail, uint32_t nt, int lane){
	const uint32_t bn = nb + nt;
	if(bn == 0) return;      /* warp-uniform */
	const uint32_t f0 = nb? __ldg(bulk + (rev? nb - 1 : 0)) : tail[0];
	const uint32_t mg = (ncig && (lastv & 0xFu) == (f0 & 0xFu))? 1u : 0u;
	if(mg){ lastv += f0 & 0xFFFFFFF0u; if(lane == 0) cig[ncig - 1] = lastv; }
	for(uint32_t k = (uint32_t)lane + mg; k < nb; k += 32u) cig[ncig + k - mg] = __ldg(bulk + (rev? nb - 1 - k : k));
	if(lane == 0) for(uint32_t k = nb > mg? 0u : mg - nb; k < nt; k++) cig[ncig + nb + k - mg] = tail[k];
	if(bn > mg) lastv = nt? tail[nt - 1] : __ldg(bulk + (rev? 0 : nb - 1));
	ncig += bn - mg;
	__syncwarp();      /* the next block may rewrite this block's last op from another lane */
}
__global__ void __launch_bounds__(128) k_wb_stitch(const WItem *items, uint32_t nitems, const DevWin *wins, AlnPar A, const unsigned long long *item_step_off, const uint8_t *item_seq,
		const WBStep *steps, const uint32_t *aops, int acap, const unsigned long long *scr_off, const uint32_t *arena, int rw, const unsigned long long *overflow,
		uint32_t *cig_arena, const unsigned long long *item_cig_off, DevReg *regs){
	const uint32_t it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
	if(it >= nitems || item_seq[it] || *overflow) return;
	const DevWin Wn = wins[items[it].win]; const DPPar P = A.P;
	uint32_t *cig = cig_arena + item_cig_off[it]; uint32_t ncig = 0, lastv = 0;
	int x_score = 0, x_tb = 0, x_te = 0, x_qb = 0, x_qe = 0, x_aln = 0, x_mat = 0, x_mis = 0, x_ins = 0, x_del = 0;
	const unsigned long long s0 = item_step_off[it]; const uint32_t na = Wn.anc1 - Wn.anc0;
	for(uint32_t a = 0; a < na; a++){
		const WBStep *S = steps + s0 + a; const uint32_t fl = S->flags;
		if(!(fl & 1u)) continue;
		const int a_off1 = (int)S->x_te + S->tlen, a_off2 = (int)S->x_qe + S->qlen;          /* the anchor's start */
		if(x_aln == 0){ x_tb = x_te = (int)S->x_te; x_qb = x_qe = (int)S->x_qe; }               /* = the anchor's start: no bridge in front of the first anchor */
		x_score = S->o_score;
		const uint32_t *ops = arena; uint32_t n = 0;
		if(fl & 4u){
			const int W = (int)S->W; const BandDims d = band_dims(S->qlen, S->tlen, 0, -W, P);
			ops = arena + scr_off[s0 + a] + (size_t)d.ql * (rw + 4); n = S->o_ncig;
			x_aln += S->o_mat + S->o_mis + S->o_ins + S->o_del; x_mat += S->o_mat; x_mis += S->o_mis; x_ins += S->o_ins; x_del += S->o_del;
			x_te += S->end_j + 1; x_qe += S->end_i + 1;
		}
		/* pads count in del / ins / aln, not in the score; block = [bridge ops in alignment order (the walk wrote them backwards) + D pad + I pad], pads merging
		 * with the op in front of them inside the block */
		uint32_t tail[3], nt = 0;
		if(n) tail[nt++] = __ldg(ops);                   /* the block's last bridge op = first op of the walk */
		if(x_te < a_off1){ const uint32_t pd = (uint32_t)(a_off1 - x_te); x_del += (int)pd; x_aln += (int)pd; x_te = a_off1; if(nt && (tail[nt - 1] & 0xFu) == 2u) tail[nt - 1] += pd << 4; else tail[nt++] = (pd << 4) | 2u; }
		if(x_qe < a_off2){ const uint32_t pi = (uint32_t)(a_off2 - x_qe); x_ins += (int)pi; x_aln += (int)pi; x_qe = a_off2; if(nt && (tail[nt - 1] & 0xFu) == 1u) tail[nt - 1] += pi << 4; else tail[nt++] = (pi << 4) | 1u; }
		wb_push_block(cig, ncig, lastv, ops + 1, n? n - 1 : 0u, true, tail, nt, lane);
		if(fl & 2u) break;        /* "should never happen": the anchor's run bases differ, the window is truncated here (hzm_aln.h:1288-1291) */
		wb_push_block(cig, ncig, lastv, aops + (s0 + a) * (unsigned long long)(acap + 1), S->a_nops, false, tail, 0u, lane);
		x_score += S->a_score; x_aln += S->a_aln; x_mat += S->a_mat; x_ins += S->a_ins; x_del += S->a_del;
		x_te += S->a_mat + S->a_del; x_qe += S->a_mat + S->a_ins;
	}
	if(lane) return;
	DevReg r; r.score = x_score; r.tb = x_tb; r.te = x_te; r.qb = x_qb; r.qe = x_qe; r.aln = x_aln; r.mat = x_mat; r.mis = x_mis; r.ins = x_ins; r.del = x_del;
	r.cig_off = item_cig_off[it]; r.cig_len = ncig;
	r.kept = !(x_aln * 2 < A.zovl || (float)x_mat < (float)x_aln * A.min_id);       /* wtzmo.c:1026 */
	regs[it] = r;
}
