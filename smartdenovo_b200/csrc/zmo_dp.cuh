/*
 * zmo_dp.cuh -- banded affine-gap DP device routines for sm_100a (integer DP, no tensor cores).
 *
 * One "executor" (a warp, or a CTA of NT threads) sweeps the band row by row.  Each thread owns C
 * consecutive columns of the current row; the horizontal-gap chain F, which in the reference is a
 * serial dependency along the row (kswx.h:174-177, ksw.c:559-562), is opened from the diagonal move
 * m = H(i-1,j-1)+s only, so F(j) = E*j + max_{k<j}(m(k)+D-E*k) is an exclusive max-plus prefix
 * scan: thread-local scan, warp shuffle scan, one shared-memory hop across warps.  Integer max/add
 * are exact and associative, so every h/e/f value and every traceback bit equals the serial result.
 * Row maximum + arg-max (first column for the shifting band, last column for the fixed band) is a
 * 64-bit key max-reduction.  H/E rows live in shared memory indexed by absolute column (masked),
 * out-of-band neighbours read as the reference's sentinels.
 * Traceback is 4 bits per cell (2-bit H source, E-extended, F-extended; the reference uses 6 of 8
 * bits, kswx.h:163-178), one 32-bit word per thread per row, walked by one lane.
 *
 * Restated behaviour: kswx_extend_align_core (kswx.h:234-335), kswx_extend_align_shift_core
 * (kswx.h:101-232), ksw_global2 (ksw.c:503-586).
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define ZMO_NEG     (-10000)          /* kswx.h:145,153 */
#define ZMO_GNEG    (-0x40000000)     /* ksw.c:489 */
#define ZMO_BIGNEG  (-1600000000)     /* "no value" for scan padding, below every reachable score */

struct SeqView { const uint32_t *w; int start, step; uint32_t comp; };
/* logical element k of a view over a device read (16 bases per uint32, MSB first) */
__device__ __forceinline__ uint32_t sv_base(const SeqView &s, int k){
	int p = s.start + k * s.step;
	return ((__ldg(s.w + (p >> 4)) >> (((~p) & 15) << 1)) & 3u) ^ s.comp;
}
__device__ __forceinline__ uint32_t pk_base(const uint32_t *pk, int k){ return (pk[k >> 4] >> (((~k) & 15) << 1)) & 3u; }

template<int NT> __device__ __forceinline__ void ex_sync(){ if(NT == 32) __syncwarp(); else __syncthreads(); }

/* materialise n logical bases of a view as packed words (logical orientation): every output word is cut out of
 * two source words with a funnel shift; backward views reverse the sixteen 2-bit fields, complemented views
 * invert them.  Fields past n are garbage and never read.  (Reads carry spare words after their last base.) */
template<int NT> __device__ __forceinline__ void stage_packed(const SeqView &s, int n, uint32_t *dst, int tid){
	const int nw = (n + 15) >> 4;
	for(int wi = tid; wi < nw; wi += NT){
		const int t0 = wi << 4; uint32_t v;
		if(s.step == 1){
			const int a = s.start + t0, i0 = a >> 4, sh = (a & 15) << 1;
			const uint32_t hi = __ldg(s.w + i0), lo = __ldg(s.w + i0 + 1);
			v = sh? (hi << sh) | (lo >> (32 - sh)) : hi;
		} else {
			const int a = s.start - t0 - 15, i0 = a >> 4, sh = (a & 15) << 1;      /* a may be negative in the last word */
			const uint32_t hi = i0 >= 0? __ldg(s.w + i0) : 0u, lo = i0 + 1 >= 0? __ldg(s.w + i0 + 1) : 0u;
			uint32_t u = sh? (hi << sh) | (lo >> (32 - sh)) : hi;
			u = __brev(u);
			v = ((u >> 1) & 0x55555555u) | ((u & 0x55555555u) << 1);
		}
		dst[wi] = s.comp? ~v : v;
	}
}

struct DPPar { int M, X, I, D, E, T; };
struct DPOut { int score, qe, te, mat, mis, ins, del, ncig; };

/* executor-private shared memory */
struct BandSmem {
	int *H0, *H1, *Ev;        /* capacity cap_mask+1 ints each */
	int cap_mask;
	int *sred;                /* 2*NW scan totals */
	long long *sredk;         /* NW arg-max keys */
	int *smisc;               /* [0] h of last column, [1..2] F carry between column chunks, [4..] broadcast */
};

/* effective dimensions after the band/length clamps (kswx.h:115-129,244-258) */
struct BandDims { int W, ql, tl, ncol; };
__host__ __device__ __forceinline__ BandDims band_dims(int qlen, int tlen, int init, int Wp, const DPPar &P){
	BandDims d; int W = Wp;
	if(W > 0){
		int mx = (qlen < tlen? qlen : tlen) * P.M + init + (-P.T);
		int max_gap = (mx + (P.I > P.D? P.I : P.D)) / (-P.E) + 1;
		if(max_gap < 1) max_gap = 1;
		if(W > max_gap) W = max_gap;
	} else W = -W;
	{ int mxl = qlen > tlen? qlen : tlen; if(W > mxl) W = mxl; }
	d.ql = qlen; d.tl = tlen;
	if(qlen < tlen){ if(qlen + W < tlen) d.tl = qlen + W; }
	else { if(tlen + W < qlen) d.ql = tlen + W; }
	d.W = W; d.ncol = d.tl < 2 * W + 1? d.tl : 2 * W + 1;
	return d;
}

/* number of 32-bit traceback words per row for a band of ncol columns */
template<int NT, int C> __host__ __device__ __forceinline__ int band_row_words(int ncol){ return ((ncol + NT * C - 1) / (NT * C)) * NT; }

/*
 * Forward sweep + traceback of one extension problem.  rowpk/colpk: packed logical query/target of
 * at least ql/tl bases.  z: >= ql*band_row_words(ncol) words, zb: >= ql ints (row band starts).
 * cig: receives run-length ops in WALK order (alignment end -> start), i.e. what the reference holds
 * before its final reverse_u32list.  All threads return the same DPOut.
 */
template<int NT, int C, int MODE>
__device__ void band_extend(const BandSmem &S, const uint32_t *rowpk, int qlen, const uint32_t *colpk, int tlen,
		int init, const BandDims &bd, const DPPar &P, uint32_t *z, int *zb, uint32_t *cig, int cig_cap,
		DPOut &out, unsigned long long *cells_acc, int tid){
	constexpr int NW = NT / 32;
	constexpr int PC = NT * C;
	const int lane = tid & 31, warp = tid >> 5;
	const int W = bd.W, ql = bd.ql, tl = bd.tl;
	const int rw = band_row_words<NT, C>(bd.ncol);
	const int mask = S.cap_mask;
	const int IE = P.I + P.E, DE = P.D + P.E, E = P.E, CE = C * P.E;
	int best = init, bi = -1, bj = -1, gbest = 0, gi = -1, gj = -1;
	int c = 0, pjb = 0, pje = tl, i;
	int *Hp = S.H0, *Hc = S.H1;
	unsigned long long cells = 0;
	for(i = 0; i < ql; i++){
		int jb, je;
		if(MODE == 1){ jb = c - W; if(jb < 0) jb = 0; je = c + W + 1; if(je > tl) je = tl; }
		else { jb = i - W; if(jb < 0) jb = 0; je = i + W + 1; if(je > tl) je = tl; }
		const uint32_t qb = pk_base(rowpk, i);
		int lmax = 0, larg = -1;
		int chunk = 0;
		if(MODE == 1 && tid == 0) zb[i] = jb;          /* the fixed band start is recomputed by the walker */
		cells += (unsigned long long)(je > jb? je - jb : 0);
		for(int cb = jb; cb < je; cb += PC, chunk++){
			const int j0 = cb + tid * C;
			int m[C], e[C];
			/* the thread's C column bases in one 64-bit window (2 shared loads), XORed with the row base: a 2-bit field is 0 on a match */
			unsigned long long xw = 0; const int xs = 62 - ((j0 & 15) << 1);
			if(j0 < je){ const int w0 = j0 >> 4; xw = (((unsigned long long)colpk[w0] << 32) | colpk[w0 + 1]) ^ (0x5555555555555555ull * qb); }
			#pragma unroll
			for(int k = 0; k < C; k++){
				const int j = j0 + k;
				if(j < je){
					int hd, ee;
					if(i == 0){ hd = j == 0? init : init + P.D + E * j; ee = ZMO_NEG; }
					else {
						if(j == 0) hd = init + P.I + E * i;
						else hd = (j - 1 >= pjb && j - 1 < pje)? Hp[(j - 1) & mask] : ZMO_NEG;
						ee = (j >= pjb && j < pje)? S.Ev[j & mask] : ZMO_NEG;
					}
					m[k] = hd + (((xw >> (xs - 2 * k)) & 3ull)? P.X : P.M);
					e[k] = ee;
				} else { m[k] = ZMO_BIGNEG; e[k] = ZMO_BIGNEG; }
			}
			/* thread summary of the F chain: value leaving the thread if nothing entered it */
			int b = ZMO_BIGNEG;
			#pragma unroll
			for(int k = 0; k < C; k++){ int t2 = m[k] + DE; b = b + E; if(b < t2) b = t2; }
			int v = b - tid * CE;
			int incl = v;
			#pragma unroll
			for(int d = 1; d < 32; d <<= 1){ int o = __shfl_up_sync(0xffffffffu, incl, d); if(lane >= d && o > incl) incl = o; }
			int excl = __shfl_up_sync(0xffffffffu, incl, 1);
			if(lane == 0) excl = ZMO_BIGNEG;
			if(NW > 1){
				int *sr = S.sred + (chunk & 1) * NW;
				if(lane == 31) sr[warp] = incl;
				__syncthreads();
				for(int w2 = 0; w2 < warp; w2++){ int o = sr[w2]; if(o > excl) excl = o; }
			}
			if(NW == 1) __syncwarp();
			int fcarry = (cb == jb)? ZMO_NEG : S.smisc[1 + (chunk & 1)];
			int f = fcarry + tid * CE;
			if(tid > 0){ int o = (tid - 1) * CE + excl; if(o > f) f = o; }
			uint32_t zw = 0;
			#pragma unroll
			for(int k = 0; k < C; k++){
				const int j = j0 + k;
				if(j < je){
					const int mm = m[k]; int ee = e[k]; int h; uint32_t d;
					if(mm >= ee){ d = 0; h = mm; } else { d = 1; h = ee; }
					if(h < f){ d = 2; h = f; }
					Hc[j & mask] = h;
					if(MODE == 1){ if(h > lmax){ lmax = h; larg = j; } }
					else { if(h >= lmax){ lmax = h; larg = j; } }
					int t1 = mm + IE; ee += E; if(ee > t1) d |= 4u; else ee = t1;
					S.Ev[j & mask] = ee;
					int t2 = mm + DE; f += E; if(f > t2) d |= 8u; else f = t2;
					zw |= d << (k << 2);
					if(j == je - 1) S.smisc[0] = h;
				} else { f += E; }
			}
			z[(size_t)i * rw + chunk * NT + tid] = zw;
			if(tid == NT - 1) S.smisc[1 + ((chunk + 1) & 1)] = f;
		}
		/* row arg-max: max h over the row, then the first (shifting band) / last (fixed band) column reaching it */
		int rowmax = __reduce_max_sync(0xffffffffu, lmax), rowarg;
		if(MODE == 1){ const int cand = (lmax == rowmax && larg >= 0)? larg : 0x7FFFFFFF; rowarg = __reduce_min_sync(0xffffffffu, cand); }
		else { const int cand = (lmax == rowmax)? larg : -1; rowarg = __reduce_max_sync(0xffffffffu, cand); }
		if(NW > 1){
			if(lane == 0) S.sredk[warp] = ((long long)rowmax << 32) | (unsigned)rowarg;
			__syncthreads();
			rowmax = (int)(S.sredk[0] >> 32); rowarg = (int)(unsigned)(S.sredk[0] & 0xffffffffu);
			#pragma unroll
			for(int w2 = 1; w2 < NW; w2++){
				const int hm = (int)(S.sredk[w2] >> 32), ha = (int)(unsigned)(S.sredk[w2] & 0xffffffffu);
				if(MODE == 1){ if(hm > rowmax || (hm == rowmax && ha < rowarg)){ rowmax = hm; rowarg = ha; } }
				else { if(hm > rowmax || (hm == rowmax && ha > rowarg)){ rowmax = hm; rowarg = ha; } }
			}
		} else __syncwarp();
		if(MODE == 1 && rowarg == 0x7FFFFFFF) rowarg = -1;
		const int hlast = (je > jb)? S.smisc[0] : (jb == 0? init + P.I + E * (i + 1) : ZMO_NEG);
		if(je == tlen && gbest < hlast){ gbest = hlast; gi = i; gj = je - 1; }
		if(i + 1 == qlen && gbest < rowmax){ gbest = rowmax; gi = i; gj = rowarg; }
		{ int *sw = Hp; Hp = Hc; Hc = sw; }
		pjb = jb; pje = je;
		if(rowmax > best){ best = rowmax; bi = i; bj = rowarg; }
		else if(rowmax <= 0) break;
		if(MODE == 1){ c++; if(c < rowarg) c++; else if(c > rowarg) c--; }
	}
	if(gbest > 0 && gbest >= best + P.T){ out.score = gbest; out.qe = gi; out.te = gj; }
	else { out.score = best; out.qe = bi; out.te = bj; }
	ex_sync<NT>();
	/* traceback by one lane (kswx.h:207-230) */
	if(tid == 0){
		int ii = out.qe, jj = out.te, st = 0, mat = 0, mis = 0, ins = 0, del = 0, n = 0;
		uint32_t cur_op = 0xF, cur_len = 0;
		while(ii >= 0 && jj >= 0){
			const int rel = jj - (MODE == 1? zb[ii] : (ii > W? ii - W : 0));
			const int ch = rel / PC, r2 = rel - ch * PC;
			const uint32_t nib = (z[(size_t)ii * rw + ch * NT + r2 / C] >> ((r2 % C) << 2)) & 0xFu;
			if(st == 0) st = nib & 3u; else if(st == 1) st = (nib & 4u)? 1 : 0; else st = (nib & 8u)? 2 : 0;
			if(st == 0){ if(pk_base(rowpk, ii) == pk_base(colpk, jj)) mat++; else mis++; ii--; jj--; }
			else if(st == 1){ ii--; ins++; }
			else { jj--; del++; }
			if((uint32_t)st == cur_op) cur_len++;
			else { if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = st; cur_len = 1; }
		}
		if(ii >= 0){
			ins += ii + 1;
			if(cur_op == 1u) cur_len += ii + 1; else { if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = 1; cur_len = ii + 1; }
		}
		if(jj >= 0){
			del += jj + 1;
			if(cur_op == 2u) cur_len += jj + 1; else { if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = 2; cur_len = jj + 1; }
		}
		if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; }
		S.smisc[4] = mat; S.smisc[5] = mis; S.smisc[6] = ins; S.smisc[7] = del; S.smisc[8] = n;
		if(cells_acc) atomicAdd(cells_acc, cells);
	}
	ex_sync<NT>();
	out.mat = S.smisc[4]; out.mis = S.smisc[5]; out.ins = S.smisc[6]; out.del = S.smisc[7]; out.ncig = S.smisc[8];
	out.qe++; out.te++;
	ex_sync<NT>();
}

/*
 * Banded global alignment (ksw.c:503-586): rows = target (the q-read slice), columns = query (the
 * c slice), band |i-j| <= w.  CIGAR op 1 consumes a column base, op 2 a row base.  Emits ops in
 * WALK order (end -> start).  mat/mis are counted during the walk (the reference counts them from
 * the CIGAR afterwards, hzm_aln.h:1423-1432).  z: >= tlen*band_row_words(ncol) words.
 */
template<int NT, int C>
__device__ void band_global(const BandSmem &S, const uint32_t *colpk /*query*/, int qlen, const uint32_t *rowpk /*target*/, int tlen,
		int w, const DPPar &P, uint32_t *z, uint32_t *cig, int cig_cap, DPOut &out, unsigned long long *cells_acc, int tid){
	constexpr int NW = NT / 32;
	constexpr int PC = NT * C;
	const int lane = tid & 31, warp = tid >> 5;
	const int o_del = -P.I, e_del = -P.E, o_ins = -P.D, e_ins = -P.E;   /* hzm_aln.h:1407 argument order */
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	const int ncol = qlen < 2 * w + 1? qlen : 2 * w + 1;
	const int rw = band_row_words<NT, C>(ncol);
	const int mask = S.cap_mask;
	const int CE = -C * e_ins;
	int *Hp = S.H0, *Hc = S.H1;
	unsigned long long cells = 0;
	int score;
	for(int i = 0; i < tlen; i++){
		const int beg = i > w? i - w : 0, end = i + w + 1 < qlen? i + w + 1 : qlen;
		const int pbeg = (i - 1) > w? i - 1 - w : 0, pend = i + w < qlen? i + w : qlen;   /* previous row band */
		const uint32_t tb = pk_base(rowpk, i);
		int chunk = 0;
		cells += (unsigned long long)(end > beg? end - beg : 0);
		for(int cb = beg; cb < end; cb += PC, chunk++){
			const int j0 = cb + tid * C;
			int m[C], e[C];
			unsigned long long xw = 0; const int xs = 62 - ((j0 & 15) << 1);
			if(j0 < end){ const int w0 = j0 >> 4; xw = (((unsigned long long)colpk[w0] << 32) | colpk[w0 + 1]) ^ (0x5555555555555555ull * tb); }
			#pragma unroll
			for(int k = 0; k < C; k++){
				const int j = j0 + k;
				if(j < end){
					int hd, ee;
					if(i == 0){ hd = j == 0? 0 : -(o_ins + e_ins * j); ee = ZMO_GNEG; }
					else {
						if(j == 0) hd = -(o_del + e_del * i);
						else hd = (j - 1 >= pbeg && j - 1 < pend)? Hp[(j - 1) & mask] : ZMO_GNEG;
						ee = (j >= pbeg && j < pend)? S.Ev[j & mask] : ZMO_GNEG;
					}
					m[k] = hd + (((xw >> (xs - 2 * k)) & 3ull)? P.X : P.M);
					e[k] = ee;
				} else { m[k] = ZMO_BIGNEG; e[k] = ZMO_BIGNEG; }
			}
			int b = ZMO_BIGNEG;
			#pragma unroll
			for(int k = 0; k < C; k++){ int t2 = m[k] - oe_ins; b = b - e_ins; if(b < t2) b = t2; }
			int v = b - tid * CE;
			int incl = v;
			#pragma unroll
			for(int d = 1; d < 32; d <<= 1){ int o = __shfl_up_sync(0xffffffffu, incl, d); if(lane >= d && o > incl) incl = o; }
			int excl = __shfl_up_sync(0xffffffffu, incl, 1);
			if(lane == 0) excl = ZMO_BIGNEG;
			if(NW > 1){
				int *sr = S.sred + (chunk & 1) * NW;
				if(lane == 31) sr[warp] = incl;
				__syncthreads();
				for(int w2 = 0; w2 < warp; w2++){ int o = sr[w2]; if(o > excl) excl = o; }
			}
			if(NW == 1) __syncwarp();
			int fcarry = (cb == beg)? ZMO_GNEG : S.smisc[1 + (chunk & 1)];
			int f = fcarry + tid * CE;
			if(tid > 0){ int o = (tid - 1) * CE + excl; if(o > f) f = o; }
			uint32_t zw = 0;
			#pragma unroll
			for(int k = 0; k < C; k++){
				const int j = j0 + k;
				if(j < end){
					const int mm = m[k]; int ee = e[k]; int h; uint32_t d;
					d = mm >= ee? 0 : 1; h = mm >= ee? mm : ee;
					if(!(h >= f)){ d = 2; h = f; }
					Hc[j & mask] = h;
					int t1 = mm - oe_del; ee -= e_del; if(ee > t1) d |= 4u; else ee = t1;
					S.Ev[j & mask] = ee;
					int t2 = mm - oe_ins; f -= e_ins; if(f > t2) d |= 8u; else f = t2;
					zw |= d << (k << 2);
					if(j == end - 1) S.smisc[0] = h;
				} else { f -= e_ins; }
			}
			z[(size_t)i * rw + chunk * NT + tid] = zw;
			if(tid == NT - 1) S.smisc[1 + ((chunk + 1) & 1)] = f;
		}
		ex_sync<NT>();
		{ int *sw = Hp; Hp = Hc; Hc = sw; }
	}
	/* score = H(tlen-1, qlen-1) as left in eh[qlen].h (ksw.c:567), including the degenerate shapes */
	if(tlen == 0) score = qlen == 0? 0 : (qlen <= w? -(o_ins + e_ins * qlen) : ZMO_GNEG);
	else {
		const int i = tlen - 1, end = i + w + 1 < qlen? i + w + 1 : qlen, beg = i > w? i - w : 0;
		if(end == qlen){
			if(end > beg) score = S.smisc[0];
			else score = beg == 0? -(o_del + e_del * (i + 1)) : ZMO_GNEG;
		} else score = ZMO_GNEG;      /* last cell outside the band: eh[qlen] keeps its initial -inf (w >= |dlen| is guaranteed by the caller) */
	}
	out.score = score;
	ex_sync<NT>();
	if(tid == 0){
		int ii = tlen - 1, kk = (ii + w + 1 < qlen? ii + w + 1 : qlen) - 1, st = 0, mat = 0, mis = 0, ins = 0, del = 0, n = 0;
		uint32_t cur_op = 0xF, cur_len = 0;
		while(ii >= 0 && kk >= 0){
			const int rel = kk - (ii > w? ii - w : 0);
			const int ch = rel / PC, r2 = rel - ch * PC;
			const uint32_t nib = (z[(size_t)ii * rw + ch * NT + r2 / C] >> ((r2 % C) << 2)) & 0xFu;
			uint32_t op;
			if(st == 0) st = nib & 3u; else if(st == 1) st = (nib & 4u)? 1 : 0; else st = (nib & 8u)? 2 : 0;
			if(st == 0){ if(pk_base(rowpk, ii) == pk_base(colpk, kk)) mat++; else mis++; ii--; kk--; op = 0; }
			else if(st == 1){ ii--; del++; op = 2; }
			else { kk--; ins++; op = 1; }
			if(op == cur_op) cur_len++;
			else { if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = op; cur_len = 1; }
		}
		if(ii >= 0){
			del += ii + 1;
			if(cur_op == 2u) cur_len += ii + 1; else { if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = 2; cur_len = ii + 1; }
		}
		if(kk >= 0){
			ins += kk + 1;
			if(cur_op == 1u) cur_len += kk + 1; else { if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = 1; cur_len = kk + 1; }
		}
		if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; }
		S.smisc[4] = mat; S.smisc[5] = mis; S.smisc[6] = ins; S.smisc[7] = del; S.smisc[8] = n;
		if(cells_acc) atomicAdd(cells_acc, cells);
	}
	ex_sync<NT>();
	out.mat = S.smisc[4]; out.mis = S.smisc[5]; out.ins = S.smisc[6]; out.del = S.smisc[7]; out.ncig = S.smisc[8];
	out.qe = qlen; out.te = tlen;
	ex_sync<NT>();
}

/*
 * kswx_refine_alignment's sweep + walk (kswx.h:602-655) for bands wider than the register executors: global alignment of the
 * query (rows, rowpk, ql) against the target (columns, colpk, tl) inside the per-row band [zb[i], ze[i]) (both arrays monotone).
 * H(-1,-1) = 0, every other neighbour outside the previous row's band reads as -10000 (the reference's rolling rows are only ever
 * written inside the bands, and the bands never move left).  Columns are swept in chunks of NT*C with the H / E rows in global
 * memory (H0, H1, Ev: tl + 2 ints each, indexed by absolute column); the F chain is the same max-plus scan as in band_extend with a
 * carry between chunks.  z: ql * band_row_words<NT,C>(wmax) words, one nibble per cell at column offset j - zb[i].  The walk runs
 * on thread 0 (this path serves the rare alignment with an indel run of several hundred bases).  Ops are emitted in walk order.
 */
template<int NT, int C>
__device__ void band_refine(const BandSmem &S, const uint32_t *rowpk, int ql, const uint32_t *colpk, int tl, const int *zb, const int *ze, int wmax,
		const DPPar &P, uint32_t *z, uint32_t *cig, int cig_cap, DPOut &out, unsigned long long *cells_acc, int tid){
	constexpr int NW = NT / 32;
	constexpr int PC = NT * C;
	const int lane = tid & 31, warp = tid >> 5;
	const int IE = P.I + P.E, DE = P.D + P.E, E = P.E, CE = C * P.E;
	const int rw = band_row_words<NT, C>(wmax);
	int *Hp = S.H0, *Hc = S.H1;
	unsigned long long cells = 0;
	int pbeg = 0, pend = 0;
	for(int i = 0; i < ql; i++){
		const int beg = zb[i], end = ze[i];
		const uint32_t qb = pk_base(rowpk, i);
		int chunk = 0;
		cells += (unsigned long long)(end > beg? end - beg : 0);
		for(int cb = beg; cb < end; cb += PC, chunk++){
			const int j0 = cb + tid * C;
			int m[C], e[C];
			unsigned long long xw = 0; const int xs = 62 - ((j0 & 15) << 1);
			if(j0 < end){ const int w0 = j0 >> 4; xw = (((unsigned long long)colpk[w0] << 32) | colpk[w0 + 1]) ^ (0x5555555555555555ull * qb); }
			#pragma unroll
			for(int k = 0; k < C; k++){
				const int j = j0 + k;
				if(j < end){
					int hd, ee;
					if(i == 0){ hd = j == 0? 0 : ZMO_NEG; ee = ZMO_NEG; }
					else {
						hd = (j - 1 >= pbeg && j - 1 < pend)? Hp[j - 1] : ZMO_NEG;
						ee = (j >= pbeg && j < pend)? S.Ev[j] : ZMO_NEG;
					}
					m[k] = hd + (((xw >> (xs - 2 * k)) & 3ull)? P.X : P.M);
					e[k] = ee;
				} else { m[k] = ZMO_BIGNEG; e[k] = ZMO_BIGNEG; }
			}
			int b = ZMO_BIGNEG;
			#pragma unroll
			for(int k = 0; k < C; k++){ const int t2 = m[k] + DE; b += E; if(b < t2) b = t2; }
			int incl = b - tid * CE;
			#pragma unroll
			for(int d = 1; d < 32; d <<= 1){ const int o = __shfl_up_sync(0xffffffffu, incl, d); if(lane >= d && o > incl) incl = o; }
			int excl = __shfl_up_sync(0xffffffffu, incl, 1);
			if(lane == 0) excl = ZMO_BIGNEG;
			if(NW > 1){
				int *sr = S.sred + (chunk & 1) * NW;
				if(lane == 31) sr[warp] = incl;
				__syncthreads();
				for(int w2 = 0; w2 < warp; w2++){ const int o = sr[w2]; if(o > excl) excl = o; }
			}
			if(NW == 1) __syncwarp();
			const int fcarry = (cb == beg)? ZMO_NEG : S.smisc[1 + (chunk & 1)];
			int f = fcarry + tid * CE;
			if(tid > 0){ const int o = (tid - 1) * CE + excl; if(o > f) f = o; }
			uint32_t zw = 0;
			#pragma unroll
			for(int k = 0; k < C; k++){
				const int j = j0 + k;
				if(j < end){
					const int mm = m[k]; int ee = e[k]; int h; uint32_t d;
					if(mm >= ee){ d = 0; h = mm; } else { d = 1; h = ee; }
					if(h < f){ d = 2; h = f; }
					Hc[j] = h;
					const int t1 = mm + IE; ee += E; if(ee > t1) d |= 4u; else ee = t1;
					S.Ev[j] = ee;
					const int t2 = mm + DE; f += E; if(f > t2) d |= 8u; else f = t2;
					zw |= d << (k << 2);
					if(j == end - 1) S.smisc[0] = h;
				} else f += E;
			}
			z[(size_t)i * rw + chunk * NT + tid] = zw;
			if(tid == NT - 1) S.smisc[1 + ((chunk + 1) & 1)] = f;
		}
		ex_sync<NT>();
		{ int *sw = Hp; Hp = Hc; Hc = sw; }
		pbeg = beg; pend = end;
	}
	out.score = (ql > 0 && ze[ql - 1] == tl && ze[ql - 1] > zb[ql - 1])? S.smisc[0] : ZMO_NEG;
	ex_sync<NT>();
	if(tid == 0){
		int ii = ql - 1, jj = tl - 1, st = 0, mat = 0, mis = 0, ins = 0, del = 0, n = 0;
		uint32_t cur_op = 0xF, cur_len = 0;
		while(ii >= 0 && jj >= 0){
			if(jj < zb[ii] || jj >= ze[ii]) break;        /* the reference reads outside its traceback here (undefined); stop instead */
			const int rel = jj - zb[ii];
			const int ch = rel / PC, r2 = rel - ch * PC;
			const uint32_t nib = (z[(size_t)ii * rw + ch * NT + r2 / C] >> ((r2 % C) << 2)) & 0xFu;
			uint32_t op;
			if(st == 0) st = nib & 3u; else if(st == 1) st = (nib & 4u)? 1 : 0; else st = (nib & 8u)? 2 : 0;
			if(st == 0){ if(pk_base(rowpk, ii) == pk_base(colpk, jj)) mat++; else mis++; ii--; jj--; op = 0; }
			else if(st == 1){ ii--; ins++; op = 1; }
			else { jj--; del++; op = 2; }
			if(op == cur_op) cur_len++;
			else { if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = op; cur_len = 1; }
		}
		if(ii >= 0){
			ins += ii + 1;
			if(cur_op == 1u) cur_len += ii + 1; else { if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = 1; cur_len = ii + 1; }
		}
		if(jj >= 0){
			del += jj + 1;
			if(cur_op == 2u) cur_len += jj + 1; else { if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; } cur_op = 2; cur_len = jj + 1; }
		}
		if(cur_len){ if(n < cig_cap) cig[n] = (cur_len << 4) | cur_op; n++; }
		S.smisc[4] = mat; S.smisc[5] = mis; S.smisc[6] = ins; S.smisc[7] = del; S.smisc[8] = n;
		if(cells_acc) atomicAdd(cells_acc, cells);
	}
	ex_sync<NT>();
	out.mat = S.smisc[4]; out.mis = S.smisc[5]; out.ins = S.smisc[6]; out.del = S.smisc[7]; out.ncig = S.smisc[8];
	out.qe = ql; out.te = tl;
	ex_sync<NT>();
}
