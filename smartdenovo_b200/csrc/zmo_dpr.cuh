/*
 * zmo_dpr.cuh -- register-resident form of the banded affine-gap sweeps of zmo_dp.cuh (sm_100a, integer DP).
 *
 * Same recurrences, sentinels, tie rules and traceback bits as band_extend / band_global (which remain the fallback
 * for bands wider than an executor's capacity), but the H and E rows never leave the register file:
 *
 *   - columns are grouped in blocks of C; block b belongs to thread b mod NT for the whole problem ("block-cyclic"),
 *     so while a block is inside the moving band its H/E values of the previous row are simply the thread's own
 *     registers.  The only value that crosses threads per row is H(i-1, first column - 1): one shuffle (one
 *     shared-memory word per warp across warps).
 *   - a band of ncol <= (NT-2)*C+1 columns touches at most NT-1 blocks, so a thread holds one block at a time; when
 *     the band start passes a block, its thread moves on to block b+NT at the right edge and resets its registers
 *     to the reference's out-of-band sentinel (what the smem version obtained from its "in previous band?" tests).
 *   - cells of the edge blocks outside [jb, je) are masked: they feed nothing into the F scan and leave sentinels.
 *   - the F chain is the same max-plus prefix scan, taken in band order, i.e. rotated by the thread that owns the
 *     first block (segmented shuffle scan in the one warp that holds both the head and the tail of the band).
 *   - traceback: 4 bits per cell, one (C <= 8) or two words per thread per row at z[(row*NT + thread)*WPT]; no
 *     per-row band start is needed to address it.  The walk is done by one warp with a 32-row look-ahead window
 *     (each lane prefetches the words around the diagonal for one row), so it costs one memory latency per 32 rows
 *     instead of one per step.
 *
 * Restated behaviour: kswx_extend_align_core (kswx.h:234-335), kswx_extend_align_shift_core (kswx.h:101-232),
 * ksw_global2 (ksw.c:503-586).
 */
#pragma once
#include "zmo_dp.cuh"

template<int NT, int C> struct RegCap { static constexpr int ncol = (NT - 2) * C + 1; static constexpr int WPT = (C + 7) / 8; };

/* 32-row look-ahead window over the traceback for the walking warp */
template<int NT, int C> struct ZWin {
	static constexpr int WPT = (C + 7) / 8;
	int iw, jw; uint32_t pw[3][WPT];
	__device__ __forceinline__ static int gblk(int jg){ return jg >= 0? jg / C : 0; }
	__device__ __forceinline__ void fill(const uint32_t *z, int ii, int jj, int lane){
		iw = ii; jw = jj;
		const int r = ii - lane, bg = gblk(jj - lane);
		#pragma unroll
		for(int s = 0; s < 3; s++){
			const int b = bg - 1 + s;
			#pragma unroll
			for(int wd = 0; wd < WPT; wd++) pw[s][wd] = (r >= 0 && b >= 0)? z[((size_t)r * NT + (b & (NT - 1))) * WPT + wd] : 0u;
		}
	}
	/* traceback nibble of cell (ii, jj); uniform arguments, all 32 lanes call */
	__device__ __forceinline__ uint32_t get(const uint32_t *z, int ii, int jj, int lane){
		int L = iw - ii;
		const int bcur = jj / C;
		int s = bcur - gblk(jw - L) + 1;
		if((unsigned)L >= 32u || (unsigned)s > 2u){ fill(z, ii, jj, lane); L = 0; s = 1; }
		const int k = jj - bcur * C;
		uint32_t mine;
		if(WPT == 1) mine = s == 0? pw[0][0] : (s == 1? pw[1][0] : pw[2][0]);
		else { const int wd = k >> 3; mine = s == 0? (wd? pw[0][WPT - 1] : pw[0][0]) : (s == 1? (wd? pw[1][WPT - 1] : pw[1][0]) : (wd? pw[2][WPT - 1] : pw[2][0])); }
		const uint32_t w = __shfl_sync(0xffffffffu, mine, L);
		return (w >> ((k & 7) << 2)) & 0xFu;
	}
};

/* walk from (ii, jj) to the origin (kswx.h:207-230 / ksw.c:569-583); opE / opF = CIGAR op of a vertical / horizontal
 * step.  Ops are emitted in WALK order by lane 0; every lane returns the same counts.  cnt[0..3] = mat, mis, op1, op2 */
template<int NT, int C>
__device__ __forceinline__ int reg_walk(const uint32_t *z, const uint32_t *rowpk, const uint32_t *colpk, int ii, int jj, uint32_t opE, uint32_t opF,
		uint32_t *cig, int cig_cap, int lane, int cnt[4]){
	static_assert(C <= 16, "two traceback words per thread at most");
	ZWin<NT, C> win; win.iw = -0x40000000; win.jw = 0;
	#pragma unroll
	for(int s = 0; s < 3; s++){
		#pragma unroll
		for(int wd = 0; wd < ZWin<NT, C>::WPT; wd++) win.pw[s][wd] = 0;
	}
	int st = 0, n = 0, mat = 0, mis = 0, c1 = 0, c2 = 0;
	uint32_t cur_op = 0xF, cur_len = 0;
	while(ii >= 0 && jj >= 0){
		const uint32_t nib = win.get(z, ii, jj, lane);
		uint32_t op;
		if(st == 0) st = nib & 3u; else if(st == 1) st = (nib & 4u)? 1 : 0; else st = (nib & 8u)? 2 : 0;
		if(st == 0){ if(pk_base(rowpk, ii) == pk_base(colpk, jj)) mat++; else mis++; ii--; jj--; op = 0; }
		else if(st == 1){ ii--; op = opE; }
		else { jj--; op = opF; }
		if(op == 1u) c1++; else if(op == 2u) c2++;
		if(op == cur_op) cur_len++;
		else { if(cur_len){ if(lane == 0 && n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = op; cur_len = 1; }
	}
	if(ii >= 0){
		if(opE == 1u) c1 += ii + 1; else c2 += ii + 1;
		if(cur_op == opE) cur_len += ii + 1; else { if(cur_len){ if(lane == 0 && n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = opE; cur_len = ii + 1; }
	}
	if(jj >= 0){
		if(opF == 1u) c1 += jj + 1; else c2 += jj + 1;
		if(cur_op == opF) cur_len += jj + 1; else { if(cur_len){ if(lane == 0 && n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = opF; cur_len = jj + 1; }
	}
	if(cur_len){ if(lane == 0 && n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; }
	cnt[0] = mat; cnt[1] = mis; cnt[2] = c1; cnt[3] = c2;
	return n;
}

/*
 * The sweep.  KIND 0: extension, fixed band (kswx_extend_align_core); 1: extension, band follows the row arg-max
 * (kswx_extend_align_shift_core); 2: global (ksw_global2; rows = target, columns = query, init = 0); 3: global inside the
 * per-row band [zb[i], ze[i]) of kswx_refine_alignment (kswx.h:602-633: both arrays monotone, H(-1,-1) = 0, every other
 * out-of-band neighbour -10000; the band may jump by whole blocks between rows, so the value handed over from the left
 * neighbour is only taken if that thread really held the adjacent block in the previous row).
 * nrow rows are swept over tl columns with half band W; full_rows / full_cols are the unclamped lengths the end-point
 * rules test against.  Requires min(tl, 2W+1) <= RegCap<NT,C>::ncol.  S.sred holds 2*NW ints (NW scan totals | NW
 * edge values), S.sredk NW keys, S.smisc >= 16 ints.  Returns through o_*: extension = chosen end point (0-based),
 * global = score.
 */
template<int NT, int C, int KIND>
__device__ __forceinline__ void reg_sweep(const BandSmem &S, const uint32_t *rowpk, const uint32_t *colpk, int nrow, int tl, int full_rows, int full_cols,
		int W, int init, const DPPar &P, uint32_t *z, int &o_score, int &o_i, int &o_j, unsigned long long &o_cells, int tid,
		const int *zb = nullptr, const int *ze = nullptr){
	static_assert(KIND != 1 || C >= 2, "the shifting band may advance two columns per row");
	static_assert((NT & (NT - 1)) == 0 && NT >= 32, "NT must be a power of two");
	constexpr int NW = NT / 32, WPT = (C + 7) / 8;
	constexpr int NEGV = KIND == 2? ZMO_GNEG : ZMO_NEG;
	static_assert(KIND >= 0 && KIND <= 3, "unknown sweep kind");
	constexpr unsigned FULL = 0xffffffffu;
	const int lane = tid & 31, warp = tid >> 5;
	const int IE = P.I + P.E, DE = P.D + P.E, E = P.E, CE = C * P.E;
	int H[C], Ev[C];
	int myblk = tid;
	unsigned long long xw = 0;       /* bases of my block's columns, 2 bits each from bit 63 down */
	auto load_cols = [&](int j0) -> unsigned long long {
		if(j0 >= tl) return 0ull;
		const int w0 = j0 >> 4;
		return (((unsigned long long)colpk[w0] << 32) | colpk[w0 + 1]) << ((j0 & 15) << 1);
	};
	{
		/* "row -1": H(-1, j) = boundary value read by cell (0, j+1) (kswx.h:140-146 / ksw.c:527-531); nothing right of row 0's band is ever read */
		const int je0 = tl < W + 1? tl : W + 1;
		#pragma unroll
		for(int k = 0; k < C; k++){ const int j = tid * C + k; H[k] = (KIND != 3 && j < je0)? init + P.D + E * (j + 1) : NEGV; Ev[k] = NEGV; }
		xw = load_cols(tid * C);
		if(NW > 1){ if(lane == 31){ S.sred[NW + warp] = H[C - 1]; if(KIND == 3) S.sredk[warp] = tid; } __syncthreads(); }
	}
	int best = init, bi = -1, bj = -1, gbest = 0, gi = -1, gj = -1;
	int c = 0;
	unsigned long long cells = 0;
	uint32_t rwd = 0;                /* sixteen row bases, reloaded every 16 rows (keeps the shared-memory load off the per-row critical path) */
	uint32_t *zp = z + (size_t)tid * WPT;
	for(int i = 0; i < nrow; i++, zp += NT * WPT){
		int jb, je;
		if((i & 15) == 0) rwd = rowpk[i >> 4];
		if(KIND == 3){ jb = zb[i]; je = ze[i]; }
		else if(KIND == 1){ jb = c - W; je = c + W + 1; } else { jb = i - W; je = i + W + 1; }
		if(jb < 0) jb = 0;
		if(je > tl) je = tl;
		cells += (unsigned long long)(je > jb? je - jb : 0);
		const int b0 = jb / C;
		const int r = (tid - b0) & (NT - 1);
		const int blk = b0 + r, j0 = blk * C;
		/* H(i-1, j0-1) from the thread that owns the block on my left (before anybody moves to a new block) */
		int left;
		if(NW == 1) left = __shfl_sync(FULL, H[C - 1], (lane + 31) & 31);
		else { left = __shfl_up_sync(FULL, H[C - 1], 1); if(lane == 0) left = S.sred[NW + ((warp + NW - 1) & (NW - 1))]; }
		if(KIND == 3){
			int lblk;
			if(NW == 1) lblk = __shfl_sync(FULL, myblk, (lane + 31) & 31);
			else { lblk = __shfl_up_sync(FULL, myblk, 1); if(lane == 0) lblk = (int)S.sredk[(warp + NW - 1) & (NW - 1)]; }
			if(lblk != blk - 1) left = NEGV;
		}
		if(blk != myblk){
			myblk = blk;
			#pragma unroll
			for(int k = 0; k < C; k++){ H[k] = NEGV; Ev[k] = NEGV; }
			xw = load_cols(j0);
		}
		if(j0 == 0) left = KIND == 3? (i == 0? 0 : NEGV) : (i == 0? init : init + P.I + E * i);
		const bool act = j0 < je;
		const int lo = jb - j0, hi = je - j0;
		const unsigned long long x = xw ^ (0x5555555555555555ull * (unsigned long long)((rwd >> (((~i) & 15) << 1)) & 3u));
		/* m = diagonal move; g[k] = best F value reaching cell k from gaps opened INSIDE this thread's block (no dependence on
		 * what enters from the left), so that after the scan every cell's F is max(entry + k*E, g[k]) with no serial chain */
		int m[C], g[C]; int b = ZMO_BIGNEG;
		if(act){
			int hd = left;
			#pragma unroll
			for(int k = 0; k < C; k++){
				const int mm = hd + (((x >> (62 - 2 * k)) & 3ull)? P.X : P.M);
				m[k] = (k >= lo && k < hi)? mm : ZMO_BIGNEG;
				hd = H[k];
				g[k] = b;
				const int t2 = m[k] + DE; b += E; if(b < t2) b = t2;
			}
		} else {
			#pragma unroll
			for(int k = 0; k < C; k++){ m[k] = ZMO_BIGNEG; g[k] = ZMO_BIGNEG; }
		}
		/* exclusive max-plus prefix over the threads in band order (rank r) */
		const int t0 = b0 & (NT - 1), w0 = t0 >> 5, l0 = t0 & 31;
		const int seg = (warp == w0 && lane >= l0)? l0 : 0;
		int incl = b - r * CE;
		#pragma unroll
		for(int d = 1; d < 32; d <<= 1){ const int o = __shfl_up_sync(FULL, incl, d); if(lane - d >= seg && o > incl) incl = o; }
		int excl = __shfl_up_sync(FULL, incl, 1);
		if(lane == seg) excl = ZMO_BIGNEG;
		if(NW == 1){ const int ht = __shfl_sync(FULL, incl, 31); if(lane < l0 && ht > excl) excl = ht; }
		else {
			if(lane == 31) S.sred[warp] = incl;
			__syncthreads();
			const int nq = warp == w0? (lane < l0? NW : 0) : ((warp - w0) & (NW - 1));
			#pragma unroll
			for(int q = 0; q < NW; q++){ if(q < nq){ const int o = S.sred[(w0 + q) & (NW - 1)]; if(o > excl) excl = o; } }
		}
		int f = NEGV + E * (r * C - (jb - b0 * C));
		if(r > 0){ const int o = (r - 1) * CE + excl; if(o > f) f = o; }
		int lmax = 0, larg = -1;
		if(act){
			uint32_t zw[WPT];
			#pragma unroll
			for(int wd = 0; wd < WPT; wd++) zw[wd] = 0;
			#pragma unroll
			for(int k = 0; k < C; k++){
				const int j = j0 + k; const int mm = m[k]; int ee = Ev[k]; int h; uint32_t d;
				const bool inb = k >= lo && k < hi;
				int fk = f + k * E; if(fk < g[k]) fk = g[k];
				if(mm >= ee){ d = 0; h = mm; } else { d = 1; h = ee; }
				if(h < fk){ d = 2; h = fk; }
				const int hn = inb? h : NEGV;
				if(KIND == 1){ if(hn > lmax){ lmax = hn; larg = j; } }
				else if(KIND == 0){ if(hn >= lmax){ lmax = hn; larg = j; } }
				const int t1 = mm + IE; ee += E; if(ee > t1) d |= 4u; else ee = t1;
				if(fk + E > mm + DE) d |= 8u;
				H[k] = hn; Ev[k] = inb? ee : NEGV;
				zw[k >> 3] |= d << ((k & 7) << 2);
			}
			#pragma unroll
			for(int wd = 0; wd < WPT; wd++) zp[wd] = zw[wd];
			/* H of the row's last column, read by the end-point rules below / by reg_global after the last row */
			const int kq = je - 1 - j0;
			if((unsigned)kq < (unsigned)C){
				int hv = H[0];
				#pragma unroll
				for(int k = 1; k < C; k++) if(kq == k) hv = H[k];
				S.smisc[0] = hv;
			}
		}
		if(KIND == 2 || KIND == 3){
			if(NW > 1){ if(lane == 31){ S.sred[NW + warp] = H[C - 1]; if(KIND == 3) S.sredk[warp] = myblk; } __syncthreads(); } else __syncwarp();
			continue;
		}
		/* row arg-max: max h over the row, then the first (shifting band) / last (fixed band) column reaching it */
		int rowmax = __reduce_max_sync(FULL, lmax), rowarg;
		if(KIND == 1){ const int cand = (lmax == rowmax && larg >= 0)? larg : 0x7FFFFFFF; rowarg = __reduce_min_sync(FULL, cand); }
		else { const int cand = (lmax == rowmax)? larg : -1; rowarg = __reduce_max_sync(FULL, cand); }
		if(NW > 1){
			if(lane == 0) S.sredk[warp] = ((long long)rowmax << 32) | (unsigned)rowarg;
			if(lane == 31) S.sred[NW + warp] = H[C - 1];
			__syncthreads();
			rowmax = (int)(S.sredk[0] >> 32); rowarg = (int)(unsigned)(S.sredk[0] & 0xffffffffu);
			#pragma unroll
			for(int w2 = 1; w2 < NW; w2++){
				const long long kk = S.sredk[w2];
				const int hm = (int)(kk >> 32), ha = (int)(unsigned)(kk & 0xffffffffu);
				if(KIND == 1){ if(hm > rowmax || (hm == rowmax && ha < rowarg)){ rowmax = hm; rowarg = ha; } }
				else { if(hm > rowmax || (hm == rowmax && ha > rowarg)){ rowmax = hm; rowarg = ha; } }
			}
		} else __syncwarp();
		if(KIND == 1 && rowarg == 0x7FFFFFFF) rowarg = -1;
		const int hlast = (je > jb)? S.smisc[0] : (jb == 0? init + P.I + E * (i + 1) : ZMO_NEG);
		if(je == full_cols && gbest < hlast){ gbest = hlast; gi = i; gj = je - 1; }
		if(i + 1 == full_rows && gbest < rowmax){ gbest = rowmax; gi = i; gj = rowarg; }
		if(rowmax > best){ best = rowmax; bi = i; bj = rowarg; }
		else if(rowmax <= 0) break;
		if(KIND == 1){ c++; if(c < rowarg) c++; else if(c > rowarg) c--; }
	}
	if(KIND != 2){
		if(gbest > 0 && gbest >= best + P.T){ o_score = gbest; o_i = gi; o_j = gj; }
		else { o_score = best; o_i = bi; o_j = bj; }
	}
	o_cells = cells;
}

/* drop-in for band_extend (same contract; z needs ql * NT * WPT words, no band-start array) */
template<int NT, int C, int MODE>
__device__ void reg_extend(const BandSmem &S, const uint32_t *rowpk, int qlen, const uint32_t *colpk, int tlen,
		int init, const BandDims &bd, const DPPar &P, uint32_t *z, uint32_t *cig, int cig_cap,
		DPOut &out, unsigned long long *cells_acc, int tid){
	unsigned long long cells = 0;
	reg_sweep<NT, C, MODE>(S, rowpk, colpk, bd.ql, bd.tl, qlen, tlen, bd.W, init, P, z, out.score, out.qe, out.te, cells, tid);
	ex_sync<NT>();
	if(tid < 32){
		int cnt[4];
		const int n = reg_walk<NT, C>(z, rowpk, colpk, out.qe, out.te, 1u, 2u, cig, cig_cap, tid, cnt);
		if(tid == 0){
			S.smisc[4] = cnt[0]; S.smisc[5] = cnt[1]; S.smisc[6] = cnt[2]; S.smisc[7] = cnt[3]; S.smisc[8] = n;
			if(cells_acc) atomicAdd(cells_acc, cells);
		}
	}
	ex_sync<NT>();
	out.mat = S.smisc[4]; out.mis = S.smisc[5]; out.ins = S.smisc[6]; out.del = S.smisc[7]; out.ncig = S.smisc[8];
	out.qe++; out.te++;
	ex_sync<NT>();
}

/* drop-in for band_global (z needs tlen * NT * WPT words) */
template<int NT, int C>
__device__ void reg_global(const BandSmem &S, const uint32_t *colpk /*query*/, int qlen, const uint32_t *rowpk /*target*/, int tlen,
		int w, const DPPar &P, uint32_t *z, uint32_t *cig, int cig_cap, DPOut &out, unsigned long long *cells_acc, int tid){
	const int o_del = -P.I, e_del = -P.E, o_ins = -P.D, e_ins = -P.E;   /* hzm_aln.h:1407 argument order */
	unsigned long long cells = 0; int d0 = 0, d1 = 0, d2 = 0, score;
	reg_sweep<NT, C, 2>(S, rowpk, colpk, tlen, qlen, tlen, qlen, w, 0, P, z, d0, d1, d2, cells, tid);
	/* score = H(tlen-1, qlen-1) as left in eh[qlen].h (ksw.c:567), including the degenerate shapes */
	if(tlen == 0) score = qlen == 0? 0 : (qlen <= w? -(o_ins + e_ins * qlen) : ZMO_GNEG);
	else {
		const int i = tlen - 1, end = i + w + 1 < qlen? i + w + 1 : qlen, beg = i > w? i - w : 0;
		if(end == qlen){
			if(end > beg) score = S.smisc[0];
			else score = beg == 0? -(o_del + e_del * (i + 1)) : ZMO_GNEG;
		} else score = ZMO_GNEG;
	}
	out.score = score;
	ex_sync<NT>();
	if(tid < 32){
		int cnt[4];
		const int ii = tlen - 1, kk = (ii + w + 1 < qlen? ii + w + 1 : qlen) - 1;
		const int n = reg_walk<NT, C>(z, rowpk, colpk, ii, kk, 2u, 1u, cig, cig_cap, tid, cnt);
		if(tid == 0){
			S.smisc[4] = cnt[0]; S.smisc[5] = cnt[1]; S.smisc[6] = cnt[2]; S.smisc[7] = cnt[3]; S.smisc[8] = n;
			if(cells_acc) atomicAdd(cells_acc, cells);
		}
	}
	ex_sync<NT>();
	out.mat = S.smisc[4]; out.mis = S.smisc[5]; out.ins = S.smisc[6]; out.del = S.smisc[7]; out.ncig = S.smisc[8];
	out.qe = qlen; out.te = tlen;
	ex_sync<NT>();
}


/* kswx_refine_alignment's DP + walk (kswx.h:602-655) for a band that fits the executor: max_i (ze[i] - zb[i]) <= RegCap<NT,C>::ncol.
 * rowpk = query (c on its strand from qb), colpk = target (q from tb); z needs ql * NT * WPT words.  cig receives the ops in
 * WALK order.  out.score = H(ql-1, tl-1). */
template<int NT, int C>
__device__ void reg_refine(const BandSmem &S, const uint32_t *rowpk, int ql, const uint32_t *colpk, int tl, const int *zb, const int *ze,
		const DPPar &P, uint32_t *z, uint32_t *cig, int cig_cap, DPOut &out, unsigned long long *cells_acc, int tid){
	unsigned long long cells = 0; int d0 = 0, d1 = 0, d2 = 0;
	reg_sweep<NT, C, 3>(S, rowpk, colpk, ql, tl, ql, tl, 0, 0, P, z, d0, d1, d2, cells, tid, zb, ze);
	ex_sync<NT>();
	out.score = (ze[ql - 1] == tl && ze[ql - 1] > zb[ql - 1])? S.smisc[0] : ZMO_NEG;
	ex_sync<NT>();
	if(tid < 32){
		int cnt[4];
		const int n = reg_walk<NT, C>(z, rowpk, colpk, ql - 1, tl - 1, 1u, 2u, cig, cig_cap, tid, cnt);
		if(tid == 0){
			S.smisc[4] = cnt[0]; S.smisc[5] = cnt[1]; S.smisc[6] = cnt[2]; S.smisc[7] = cnt[3]; S.smisc[8] = n;
			if(cells_acc) atomicAdd(cells_acc, cells);
		}
	}
	ex_sync<NT>();
	out.mat = S.smisc[4]; out.mis = S.smisc[5]; out.ins = S.smisc[6]; out.del = S.smisc[7]; out.ncig = S.smisc[8];
	out.qe = ql; out.te = tl;
	ex_sync<NT>();
}
