/*
 * zmo_winlane.cuh -- per-window anchored alignment (fast_seeds_align_hzmo, hzm_aln.h:1247-1302), ONE LANE PER WINDOW.
 *
 * Why not a warp per window (zmo_winalign.cuh, kept as the fallback for very wide -w): the bridges between the anchors of a
 * window are tiny -- median 28 x 28 cells, 53% of them at most 32 x 32, 16% empty (measured with the oracle on cfg2-like reads) --
 * so a warp-wide row sweep spends its instructions on the per-row scan / reduction / hand-over and on single-lane phases (walk,
 * anchor alignment): 165 thread-instructions per cell, 6.9 warp-instructions per cell (ncu, round 1).  Here every lane runs the
 * plain serial recurrence of kswx_extend_align_core (kswx.h:234-335) on its own window: no scan, no reduction, no shuffle.
 *
 *   - H/E of the previous row live in a per-lane ring in shared memory, indexed by absolute column mod `cap` (cap = a multiple of 8
 *     >= 2W + 2); slot s of lane t is sm[s * WL_NT + t], so the lanes of a warp never conflict whatever columns they are at.
 *   - cells are processed in 8-aligned column groups (one 32-bit traceback word, 4 bits per cell: 2-bit H source, E extended,
 *     F extended, as in zmo_dpr.cuh); the eight column bases come from one funnel shift of the packed read.
 *   - SIMT divergence is managed, not avoided: the main loop is flat -- every iteration each lane that is sweeping handles one
 *     8-cell group of whatever row / bridge / window it is at -- and the serial in-between phases (traceback walk, D/I padding, CIGAR
 *     block, run-length alignment of the anchor, fetching the next anchor or the next window from the work counter) run as ONE
 *     "advance" step that the warp enters only when at least WL_EPI_MIN lanes wait for it (or nobody can sweep), so its cost is
 *     shared between those lanes instead of being paid once per lane.
 *   - the walk reads the traceback through a 4-row look-ahead FIFO (the rows above, at the word the diagonal predicts), so it pays
 *     about one memory latency per four rows.
 *
 * Output contract = k_window_align's: DevReg per (task, window) item and the window's CIGAR ops in its region of the CIGAR arena.
 * Restated behaviour: hzm_aln.h:1247-1302 (anchor walk, pads), kswx.h:234-335 (fixed-band extension), hzm_aln.h:278-314 (run-length
 * anchor alignment), kswx.h:39-52 (CIGAR run merging), wtzmo.c:1026 (region filter).
 */
#pragma once
#include "zmo_winalign.cuh"

#ifndef ZMO_DYN_SMEM
#define ZMO_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

#define WL_NT 64            /* lanes (= windows in flight) per CTA */
#define WL_EPI_MIN 8        /* lanes that must wait before the warp takes an "advance" step (launch parameter epi_min; ZMO_WL_EPI overrides it for tuning) */
#define WL_KEY_SH 10
#define WL_KEY (1 << WL_KEY_SH)       /* row arg-max key = h * WL_KEY + (column - band start + 1): bands up to 1,022 columns */
enum { WL_EPI = 0, WL_CELLS = 1, WL_DONE = 2 };

/* ring capacity (slots per lane) and traceback words per row for half band w */
__host__ __device__ __forceinline__ int wl_cap(int w){ return (2 * w + 2 + 7) & ~7; }
__host__ __device__ __forceinline__ int wl_row_words(int w){ return ((2 * w + 1) >> 3) + 2; }

__device__ __forceinline__ uint32_t wl_b1(const uint32_t *w, int p){ return (__ldg(w + (p >> 4)) >> (((~p) & 15) << 1)) & 3u; }
/* base p of c on the strand of the task (reverse complement for dir = 1, view_pb2) */
__device__ __forceinline__ uint32_t wl_b2(const uint32_t *w, int clen, uint32_t dir, int p){
	if(dir){ const int r = clen - 1 - p; return ((__ldg(w + (r >> 4)) >> (((~r) & 15) << 1)) & 3u) ^ 3u; }
	return wl_b1(w, p);
}

__global__ void __launch_bounds__(WL_NT) k_wa_lane(const WItem *items, uint32_t nitems, const AlnTask *tasks, const zmo_pair_t *pairs,
		const DevWin *wins, const DevZPair *anchors, DevReads R, AlnPar A, uint32_t *arena, unsigned long long slab_words, int cap, int rw, int epi_min,
		uint32_t *cig_arena, const unsigned long long *item_cig_off, DevReg *regs, unsigned long long *ctr, int ctr_work, int ctr_cells){
	ZMO_DYN_SMEM(wl_raw);
	int2 *const sm = (int2*)wl_raw + threadIdx.x;         /* slot s of this lane: sm[s * WL_NT] */
	uint32_t *const z = arena + ((unsigned long long)blockIdx.x * WL_NT + threadIdx.x) * slab_words;      /* traceback: row i at z[i * rw] */
	constexpr unsigned FULL = 0xffffffffu;
	const DPPar P = A.P; const int IE = P.I + P.E, DE = P.D + P.E, E = P.E, cM = P.M, cX = P.X;
	/* window state */
	uint32_t it = 0, ncig = 0, ai = 0, anc1 = 0, dir = 0; uint32_t *cig = nullptr; const uint32_t *qwp = nullptr, *cwp = nullptr; int clen = 0;
	int x_score = 0, x_tb = 0, x_te = 0, x_qb = 0, x_qe = 0, x_aln = 0, x_mat = 0, x_mis = 0, x_ins = 0, x_del = 0;
	uint32_t a_off1 = 0, a_off2 = 0, a_len1 = 0, a_len2 = 0;
	/* bridge state */
	int qlen = 0, tlen = 0, ql = 0, tl = 0, W = 0, init = 0, i = 0, jb = 0, je = 0, j8 = 0, slot8 = 0, se = 0;
	int hd = 0, f = 0, key = 0, hl = 0, best = 0, bi = -1, bj = -1, gbest = 0, gi = -1, gj = -1;
	uint32_t rmask = 0; uint32_t *zp = z;
	int st = WL_EPI; bool open = false, have = false;
	unsigned long long cells = 0;

#define WL_ROW_BEGIN() do { \
		jb = i > W? i - W : 0; je = i + W + 1 < tl? i + W + 1 : tl; \
		j8 = jb & ~7; slot8 = j8 % cap; \
		if(jb == 0) hd = i == 0? init : init + P.I + E * i; \
		else { int s_ = slot8 + (jb & 7) - 1; if(s_ < 0) s_ += cap; hd = sm[s_ * WL_NT].x; } \
		if(i > 0 && i + W < tl) sm[se * WL_NT] = make_int2(ZMO_NEG, ZMO_NEG);      /* column i+W enters the band: nothing above it */ \
		se = se + 1 == cap? 0 : se + 1; \
		f = ZMO_NEG; key = 0; \
		rmask = 0x5555u * wl_b2(cwp, clen, dir, x_qe + i); \
		zp = z + (size_t)i * rw; \
	} while(0)

	for(;;){
		if(st == WL_CELLS){
			/* one 8-aligned column group of row i */
			const int tp = x_te + j8, wi = tp >> 4;
			const unsigned long long v = ((((unsigned long long)__ldg(qwp + wi)) << 32) | __ldg(qwp + wi + 1)) << ((tp & 15) << 1);
			const uint32_t x = (uint32_t)(v >> 48) ^ rmask;       /* field k (bits 15-2k, 14-2k) is 0 iff column j8+k matches the row base */
			uint32_t zw = 0;
			int2 *const sp = sm + slot8 * WL_NT;
#define WL_CELL(k) do { \
				const int2 w = sp[(k) * WL_NT];                     /* H(i-1, j), E(i-1, j) */ \
				const int m = hd + (((x >> (14 - 2 * (k))) & 3u)? cX : cM); \
				int e = w.y, h; uint32_t d; \
				if(m >= e){ d = 0; h = m; } else { d = 1; h = e; } \
				if(h < f){ d = 2; h = f; } \
				{ const int kk = h * WL_KEY + (jr + (k)); if(kk > key) key = kk; }      /* h >= rowmax: the last column wins (kswx.h:284-285) */ \
				{ const int t1 = m + IE; e += E; if(e > t1) d |= 4u; else e = t1; } \
				{ const int t2 = m + DE; f += E; if(f > t2) d |= 8u; else f = t2; } \
				zw |= d << (4 * (k)); \
				sp[(k) * WL_NT] = make_int2(h, e); \
				hd = w.x; hl = h; \
			} while(0)
			const int jr = j8 - jb + 1;
			if(j8 >= jb && j8 + 8 <= je){
				/* interior group: no edge tests */
				WL_CELL(0); WL_CELL(1); WL_CELL(2); WL_CELL(3); WL_CELL(4); WL_CELL(5); WL_CELL(6); WL_CELL(7);
			} else {
				#pragma unroll
				for(int k = 0; k < 8; k++){ const int j = j8 + k; if(j >= jb && j < je) WL_CELL(k); }
			}
#undef WL_CELL
			zp[(j8 >> 3) - (jb >> 3)] = zw;
			j8 += 8; slot8 += 8; if(slot8 == cap) slot8 = 0;
			if(j8 >= je){
				/* end of row i (kswx.h:288-303) */
				const int rowmax = key >> WL_KEY_SH, rowarg = jb + (key & (WL_KEY - 1)) - 1;      /* key >= 0 */
				bool stop = false;
				cells += (unsigned long long)(je - jb);
				if(je == tlen && gbest < hl){ gbest = hl; gi = i; gj = je - 1; }
				if(i + 1 == qlen && gbest < rowmax){ gbest = rowmax; gi = i; gj = rowarg; }
				if(rowmax > best){ best = rowmax; bi = i; bj = rowarg; } else if(rowmax <= 0) stop = true;
				i++;
				if(stop || i >= ql) st = WL_EPI; else WL_ROW_BEGIN();
			}
		}
		const unsigned m_epi = __ballot_sync(FULL, st == WL_EPI), m_cells = __ballot_sync(FULL, st == WL_CELLS);
		if((m_epi | m_cells) == 0u) break;
		if(st == WL_EPI && (__popc(m_epi) >= epi_min || m_cells == 0u)){
			/* ---- advance: finish the open bridge, then anchors / windows until the next bridge that has cells ---- */
			int o_score = 0, o_qe = 0, o_te = 0, o_mat = 0, o_mis = 0, o_ins = 0, o_del = 0; uint32_t bn = 0;
			int ph = open? 2 : (have? 1 : 0);
			for(;;){
				if(ph == 0){
					/* next window from the work counter */
					it = (uint32_t)atomicAdd(ctr + ctr_work, 1ULL);
					if(it >= nitems){ st = WL_DONE; have = false; break; }
					const WItem I = items[it]; const AlnTask T = tasks[I.task]; const zmo_pair_t pr = pairs[T.pair_idx]; const DevWin Wn = wins[I.win];
					qwp = R.words + R.woff[pr.qid]; cwp = R.words + R.woff[pr.cid]; clen = (int)R.len[pr.cid]; dir = T.dir;
					cig = cig_arena + item_cig_off[it]; ncig = 0; ai = Wn.anc0; anc1 = Wn.anc1;
					x_score = x_tb = x_te = x_qb = x_qe = x_aln = x_mat = x_mis = x_ins = x_del = 0;
					have = true; ph = 1;
				}
				if(ph == 1){
					/* next anchor (hzm_aln.h:1256-1262) */
					if(ai >= anc1){ ph = 4; }
					else {
						const DevZPair p = anchors[ai];
						if(x_aln == 0){ x_tb = x_te = (int)p.off1; x_qb = x_qe = (int)p.off2; }
						if((int)p.off1 < x_te || (int)p.off2 < x_qe){ ai++; continue; }
						a_off1 = p.off1; a_off2 = p.off2; a_len1 = p.len1; a_len2 = p.len2;
						qlen = (int)p.off2 - x_qe; tlen = (int)p.off1 - x_te; init = x_score < 0? 0 : x_score;
						o_score = init; o_qe = o_te = o_mat = o_mis = o_ins = o_del = 0; bn = 0;
						if(qlen > 0 && tlen > 0){
							const BandDims d = band_dims(qlen, tlen, init, A.w, P);
							W = d.W; ql = d.ql; tl = d.tl;
							{ const int je0 = tl < W + 1? tl : W + 1; for(int j = 0; j < je0; j++) sm[j * WL_NT] = make_int2(init + P.D + E * (j + 1), ZMO_NEG); }     /* row -1 (kswx.h:140-146) */
							best = init; bi = bj = -1; gbest = 0; gi = gj = -1; i = 0; se = W;
							WL_ROW_BEGIN();
							open = true; st = WL_CELLS;
							break;
						}
						ph = 3;
					}
				}
				if(ph == 2){
					/* bridge swept: end point (kswx.h:305-310) and traceback walk (kswx.h:311-330), ops in WALK order at cig + ncig */
					int ii, jj;
					if(gbest > 0 && gbest >= best + P.T){ o_score = gbest; ii = gi; jj = gj; } else { o_score = best; ii = bi; jj = bj; }
					o_qe = ii + 1; o_te = jj + 1;
					o_mat = o_mis = o_ins = o_del = 0; bn = 0;
					uint32_t *blk = cig + ncig; uint32_t cur_op = 0xFu, cur_len = 0; int sw = 0;
					/* look-ahead FIFO: p0 = word of row ii, p1..p3 = the rows above at the word the diagonal predicts */
					uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0; int k0 = -1, k1 = -1, k2 = -1, k3 = -1;
					auto WL_WIDX = [&](int r_, int c_) -> int { const int k_ = (c_ >> 3) - ((r_ > W? r_ - W : 0) >> 3); return k_ < 0? 0 : (k_ >= rw? rw - 1 : k_); };
					if(ii >= 0 && jj >= 0){
						k0 = WL_WIDX(ii, jj); p0 = z[(size_t)ii * rw + k0];
						if(ii >= 1){ const int c_ = jj >= 1? jj - 1 : 0; k1 = WL_WIDX(ii - 1, c_); p1 = z[(size_t)(ii - 1) * rw + k1]; }
						if(ii >= 2){ const int c_ = jj >= 2? jj - 2 : 0; k2 = WL_WIDX(ii - 2, c_); p2 = z[(size_t)(ii - 2) * rw + k2]; }
						if(ii >= 3){ const int c_ = jj >= 3? jj - 3 : 0; k3 = WL_WIDX(ii - 3, c_); p3 = z[(size_t)(ii - 3) * rw + k3]; }
					}
					while(ii >= 0 && jj >= 0){
						const int kn = WL_WIDX(ii, jj);
						if(kn != k0){ k0 = kn; p0 = z[(size_t)ii * rw + kn]; }
						const uint32_t nib = (p0 >> ((jj & 7) << 2)) & 0xFu;
						uint32_t op; bool up = false;
						if(sw == 0) sw = (int)(nib & 3u); else if(sw == 1) sw = (nib & 4u)? 1 : 0; else sw = (nib & 8u)? 2 : 0;
						if(sw == 0){ if(wl_b2(cwp, clen, dir, x_qe + ii) == wl_b1(qwp, x_te + jj)) o_mat++; else o_mis++; ii--; jj--; op = 0; up = true; }
						else if(sw == 1){ ii--; o_ins++; op = 1u; up = true; }
						else { jj--; o_del++; op = 2u; }
						if(op == cur_op) cur_len++;
						else { if(cur_len) blk[bn++] = (cur_len << 4) | cur_op; cur_op = op; cur_len = 1; }
						if(up){
							p0 = p1; k0 = k1; p1 = p2; k1 = k2; p2 = p3; k2 = k3; k3 = -1;
							if(ii >= 3 && jj >= 0){ const int c_ = jj >= 3? jj - 3 : 0; k3 = WL_WIDX(ii - 3, c_); p3 = z[(size_t)(ii - 3) * rw + k3]; }
						}
					}
					if(ii >= 0){ o_ins += ii + 1; if(cur_op == 1u) cur_len += (uint32_t)(ii + 1); else { if(cur_len) blk[bn++] = (cur_len << 4) | cur_op; cur_op = 1u; cur_len = (uint32_t)(ii + 1); } }
					if(jj >= 0){ o_del += jj + 1; if(cur_op == 2u) cur_len += (uint32_t)(jj + 1); else { if(cur_len) blk[bn++] = (cur_len << 4) | cur_op; cur_op = 2u; cur_len = (uint32_t)(jj + 1); } }
					if(cur_len) blk[bn++] = (cur_len << 4) | cur_op;
					for(uint32_t a = 0, b = bn; a + 1 < b; a++){ b--; const uint32_t t_ = blk[a]; blk[a] = blk[b]; blk[b] = t_; }       /* alignment order */
					open = false; ph = 3;
				}
				if(ph == 3){
					/* bridge result -> window (hzm_aln.h:1264-1286): pads count in del / ins / aln, not in the score */
					x_score = o_score;
					x_aln += o_mat + o_mis + o_ins + o_del; x_mat += o_mat; x_mis += o_mis; x_ins += o_ins; x_del += o_del;
					x_te += o_te; x_qe += o_qe;
					uint32_t *blk = cig + ncig;
					if(x_te < (int)a_off1){ const uint32_t pd = a_off1 - (uint32_t)x_te; x_del += (int)pd; x_aln += (int)pd; x_te = (int)a_off1; if(bn && (blk[bn - 1] & 0xFu) == 2u) blk[bn - 1] += pd << 4; else blk[bn++] = (pd << 4) | 2u; }
					if(x_qe < (int)a_off2){ const uint32_t pi = a_off2 - (uint32_t)x_qe; x_ins += (int)pi; x_aln += (int)pi; x_qe = (int)a_off2; if(bn && (blk[bn - 1] & 0xFu) == 1u) blk[bn - 1] += pi << 4; else blk[bn++] = (pi << 4) | 1u; }
					/* the block joins the window CIGAR with run merging of its FIRST op only (kswx_push_cigars, kswx.h:46-52) */
					if(bn){
						if(ncig && (cig[ncig - 1] & 0xFu) == (blk[0] & 0xFu)){
							cig[ncig - 1] += blk[0] & 0xFFFFFFF0u;
							for(uint32_t k = 1; k < bn; k++) cig[ncig + k - 1] = blk[k];
							ncig += bn - 1;
						} else ncig += bn;
					}
					/* run-length alignment of the anchor itself (hz_align_hzmo, hzm_aln.h:278-314); its ops are staged one slot above the CIGAR end */
					{
						uint32_t *b2 = cig + ncig + 1; uint32_t n2 = 0, sa = 0, sb = 0; bool bad = false;
						int y_score = 0, y_aln = 0, y_mat = 0, y_ins = 0, y_del = 0;
						while(sa < a_len1 || sb < a_len2){
							const uint32_t ca = sa < a_len1? wl_b1(qwp, (int)(a_off1 + sa)) : 4u, cb = sb < a_len2? wl_b2(cwp, clen, dir, (int)(a_off2 + sb)) : 5u;
							if(ca != cb){ bad = true; break; }
							uint32_t ea = sa + 1; while(ea < a_len1 && wl_b1(qwp, (int)(a_off1 + ea)) == ca) ea++;
							uint32_t eb = sb + 1; while(eb < a_len2 && wl_b2(cwp, clen, dir, (int)(a_off2 + eb)) == cb) eb++;
							const uint32_t na = ea - sa, nb = eb - sb;
							if(na < nb){ y_aln += (int)nb; y_mat += (int)na; y_ins += (int)(nb - na); y_score += (int)na * P.M + P.I + (int)(nb - na) * P.E; if(n2 < 94u){ cig_put(b2, n2, 0, na); cig_put(b2, n2, 1, nb - na); } }
							else if(na == nb){ y_aln += (int)na; y_mat += (int)na; y_score += (int)na * P.M; if(n2 < 94u) cig_put(b2, n2, 0, na); }
							else { y_aln += (int)na; y_mat += (int)nb; y_del += (int)(na - nb); y_score += (int)nb * P.M + P.D + (int)(na - nb) * P.E; if(n2 < 94u){ cig_put(b2, n2, 0, nb); cig_put(b2, n2, 2, na - nb); } }
							sa = ea; sb = eb;
						}
						if(bad || y_aln == 0){ ph = 4; }        /* "should never happen": the window is truncated here (hzm_aln.h:1288-1291) */
						else {
							cig_cat(cig, ncig, b2, n2, false);
							x_score += y_score; x_aln += y_aln; x_mat += y_mat; x_ins += y_ins; x_del += y_del;
							x_te += y_mat + y_del; x_qe += y_mat + y_ins;
							ai++; ph = 1; continue;
						}
					}
				}
				if(ph == 4){
					/* window done: record + region filter (wtzmo.c:1026) */
					DevReg r; r.score = x_score; r.tb = x_tb; r.te = x_te; r.qb = x_qb; r.qe = x_qe; r.aln = x_aln; r.mat = x_mat; r.mis = x_mis; r.ins = x_ins; r.del = x_del;
					r.cig_off = item_cig_off[it]; r.cig_len = ncig;
					r.kept = !(x_aln * 2 < A.zovl || (float)x_mat < (float)x_aln * A.min_id);
					regs[it] = r;
					have = false; ph = 0;
				}
			}
		}
	}
#undef WL_ROW_BEGIN
	if(cells) atomicAdd(ctr + ctr_cells, cells);
}
