/*
 * zmo_dot_core.cuh -- dot-matrix ("-U") pair aligner as plain host+device functions; runs one
 * thread per pair inside k_p_dot (zmo_dot.cu) and, for the CPU-only tests, inside tests/hostsim.
 *
 * Restated behaviour: denoising_hzmps (hzm_aln.h:721-889), fast_merge_wtseedv (:933-1054),
 * chaining_overhang_wtseedv (:1056-1132), dot_matrix_align_hzmps (:1134-1181), including the quirks
 * listed in SURVEY 8a-23 (a diagonal bucket never contains the strand's last diagonal, a diagonal's
 * members are read as `cnt` consecutive entries of the mixed-strand list, `len` restarts from the
 * previous element, group ids merge to the minimum root, 30-bit node weights).
 */
#pragma once
#include "zmo_seed_core.cuh"

struct DevDiag { int32_t offset; uint32_t off, cnt; };
struct DevZPairG { DevZPair p; uint32_t gid; };
struct DotPar { int xvar, yvar, min_block_len, max_overhang; float deviation_penalty, gap_penalty; };
struct DotRes { int score, qb, qe, tb, te, strand; };
/* scratch of one pair with n matches */
struct DotScratch { uint32_t *gid; DevDiag *diags; uint32_t *block; uint32_t *grps; DevZPairG *dst; DevWin *regs; int *nodes; };
ZMO_HD size_t zmo_dot_scratch_bytes(uint32_t n){ return (size_t)(n + 2) * (4 + sizeof(DevDiag) + 4 + 4 + sizeof(DevZPairG) + sizeof(DevWin) + 16) + 64; }
ZMO_HD DotScratch zmo_dot_scratch_carve(uint8_t *base, uint32_t n){
	DotScratch S; uint8_t *p = base; const size_t m = (size_t)n + 2;
	S.dst = (DevZPairG*)p; p += m * sizeof(DevZPairG);
	S.regs = (DevWin*)p; p += m * sizeof(DevWin);
	S.diags = (DevDiag*)p; p += m * sizeof(DevDiag);
	S.nodes = (int*)p; p += m * 16;
	S.gid = (uint32_t*)p; p += m * 4;
	S.block = (uint32_t*)p; p += m * 4;
	S.grps = (uint32_t*)p;
	return S;
}

struct GtZPairDiag { ZMO_HDM bool operator()(const DevZPair &a, const DevZPair &b) const {
	return ((((int64_t)a.off1 - (int64_t)a.off2) << 32) | (int64_t)a.off1) > ((((int64_t)b.off1 - (int64_t)b.off2) << 32) | (int64_t)b.off1); } };
struct GtIdxOff1 { const DevZPair *rs; ZMO_HDM bool operator()(uint32_t a, uint32_t b) const { return rs[a].off1 > rs[b].off1; } };
struct DotKey { uint64_t key; uint32_t idx, pad; };
struct GtDotKey { ZMO_HDM bool operator()(const DotKey &a, const DotKey &b) const { return a.key > b.key; } };
struct GtZPairGidOff1 { ZMO_HDM bool operator()(const DevZPairG &a, const DevZPairG &b) const { return (a.gid > b.gid)? true : ((a.gid < b.gid)? false : (a.p.off1 > b.p.off1)); } };
struct GtWinDiag { ZMO_HDM bool operator()(const DevWin &a, const DevWin &b) const {
	return ((((int64_t)(a.beg[0] - a.beg[1])) << 32) | (int64_t)a.beg[0]) > ((((int64_t)(b.beg[0] - b.beg[1])) << 32) | (int64_t)b.beg[0]); } };
struct GtIdxWBeg0 { const DevWin *w; ZMO_HDM bool operator()(uint32_t a, uint32_t b) const { return w[a].beg[0] > w[b].beg[0]; } };
struct GtWinGrpBeg0 { ZMO_HDM bool operator()(const DevWin &a, const DevWin &b) const { return (a.pb2 > b.pb2)? true : ((a.pb2 < b.pb2)? false : (a.beg[0] > b.beg[0])); } };
struct GtWinClosed { ZMO_HDM bool operator()(const DevWin &a, const DevWin &b) const { return a.closed > b.closed; } };
struct GtWinBeg0 { ZMO_HDM bool operator()(const DevWin &a, const DevWin &b) const { return a.beg[0] > b.beg[0]; } };

/* how the exact sort_array emulation is run: SerialSort = in place by the calling thread (host tests); the device kernel
 * substitutes a sorter that stages the array in shared memory with the help of the other lanes of the warp */
struct SerialSort { template<class T, class GT> ZMO_HDM void operator()(T *a, size_t n, GT gt) const { zmo_ref_sort(a, n, gt); } };

ZMO_HDN void zmo_tidy_groups(uint32_t *g, uint32_t n){          /* hzm_aln.h:836-846 / 1016-1026 */
	for(uint32_t i = 1; i < n; i++){
		if(g[i] < i) continue;
		for(uint32_t j = i + 1; j < n; j++){
			if(g[j] != i) continue;
			for(uint32_t k = j + 1; k < n; k++) if(g[k] == j) g[k] = i;
		}
	}
}

/* hzm_aln.h:721-889 for one strand; rs is the (diagonal,off1)-sorted match list, gid[] its group ids
 * (zero on entry for this strand's entries).  Returns the number of blocks written to S.regs. */
template<class SORT> ZMO_HDN uint32_t zmo_denoise_strand(const DevZPair *rs, uint32_t n, int dir, const DotPar &par, DotScratch &S, const SORT &sorter){
	const int xvar = par.xvar, yvar = par.yvar, min_len = par.min_block_len;
	uint32_t i, j, k, doff, dcnt, gid, ndiag = 0, ngrp = 0, nblock, ndst = 0, nreg = 0; int lst_offset, end_offset, len;
	/* packed sort keys of a diagonal bucket live in the (still unused) block region; those of the group sort in the node region */
	uint64_t *blk = (uint64_t*)(((uintptr_t)S.regs + 7) & ~(uintptr_t)7); DotKey *gk = (DotKey*)(((uintptr_t)S.nodes + 7) & ~(uintptr_t)7);
	for(i = 0; i < n; i++){
		if(rs[i].dir1 ^ rs[i].dir2 ^ dir) continue;
		const int dg = (int)rs[i].off1 - (int)rs[i].off2;
		if(ndiag && S.diags[ndiag - 1].offset == dg) S.diags[ndiag - 1].cnt++;
		else { S.diags[ndiag].offset = dg; S.diags[ndiag].off = i; S.diags[ndiag].cnt = 1; ndiag++; }
	}
	doff = 0; end_offset = -0x7FFFFFFF;
	S.grps[ngrp++] = 0;
	while(doff < n && ndiag){
		lst_offset = S.diags[doff].offset; dcnt = 0;
		while(1){
			if(S.diags[dcnt + doff].offset > lst_offset + yvar) break;
			if(dcnt + doff + 1 >= ndiag) break;
			dcnt++;
		}
		if(dcnt == 0) break;
		if(S.diags[doff + dcnt].offset == end_offset){ doff += dcnt; continue; }
		end_offset = S.diags[doff + dcnt].offset;
		nblock = 0;
		for(i = 0; i < dcnt; i++){
			const DevDiag d = S.diags[i + doff];
			for(j = 0; j < d.cnt; j++){
				if(d.off + j >= n) break;
				const DevZPair &q = rs[d.off + j];
				if(q.dir1 ^ q.dir2 ^ dir) continue;
				blk[nblock++] = ((uint64_t)q.off1 << 32) | (d.off + j);      /* sort key | index: the comparator reads no other memory */
			}
		}
		sorter(blk, (size_t)nblock, GtHi32());
		if(nblock){
			int p0_off1 = (int)(blk[0] >> 32), p0_len1 = (int)rs[(uint32_t)blk[0]].len1;
			len = p0_len1; j = 0;
			for(i = 1; i <= nblock; i++){
				const int p_off1 = (i == nblock)? 0x7FFFFFFF : (int)(blk[i] >> 32), p_len1 = (i == nblock)? 0 : (int)rs[(uint32_t)blk[i]].len1;
				if(p_off1 <= p0_off1 + p0_len1 || p_off1 <= p0_off1 + p0_len1 + xvar){
					len += (int)((uint32_t)p_off1 + (uint32_t)p_len1) - (p0_off1 + p0_len1);
				} else {
					if(len >= min_len){
						gid = 0;
						for(k = j; k < i; k++){ const uint32_t g = S.gid[(uint32_t)blk[k]]; if(g){ if(gid == 0) gid = S.grps[g]; else if(gid > S.grps[g]) gid = S.grps[g]; } }
						if(gid == 0){ gid = ngrp; S.grps[ngrp++] = gid; }
						else { for(k = j; k < i; k++){ const uint32_t g = S.gid[(uint32_t)blk[k]]; if(g) S.grps[g] = gid; } }
						for(; j < i; j++) S.gid[(uint32_t)blk[j]] = gid;
					}
					j = i;
					len = p0_len1;
				}
				p0_off1 = p_off1; p0_len1 = p_len1;
			}
		}
		for(i = doff; i < doff + dcnt; i++) if(S.diags[i].offset > lst_offset + yvar / 2) break;
		doff = i;
	}
	zmo_tidy_groups(S.grps, ngrp);
	for(i = 0; i < n; i++){
		if(rs[i].dir1 ^ rs[i].dir2 ^ dir) continue;
		if(S.gid[i] == 0) continue;
		S.gid[i] = S.grps[S.gid[i]];
		gk[ndst].key = ((uint64_t)S.gid[i] << 32) | rs[i].off1; gk[ndst].idx = i; gk[ndst].pad = 0; ndst++;
	}
	/* (gid, off1) as one 64-bit key: the same comparison outcomes as the two-level comparator, hence the same permutation */
	sorter(gk, (size_t)ndst, GtDotKey());
	for(i = 0; i < ndst; i++){ S.dst[i].p = rs[gk[i].idx]; S.dst[i].gid = (uint32_t)(gk[i].key >> 32); }
	j = 0;
	for(i = 1; i <= ndst; i++){
		DevWin s; uint32_t lst = 0;
		if(i < ndst && S.dst[i].gid == S.dst[j].gid) continue;
		s.pb2 = 0; s.closed = 0; s.dir = (uint8_t)dir; s.pad = 0; s.anc0 = j; s.anc1 = i;
		s.beg[0] = s.beg[1] = 0x7FFFFFFF; s.end[0] = s.end[1] = 0; s.ovl = 0;
		for(k = j; k < i; k++){
			const DevZPair &p = S.dst[k].p;
			if((int)p.off1 < s.beg[0]) s.beg[0] = p.off1;
			if((int)(p.off1 + p.len1) > s.end[0]) s.end[0] = p.off1 + p.len1;
			if((int)p.off2 < s.beg[1]) s.beg[1] = p.off2;
			if((int)(p.off2 + p.len2) > s.end[1]) s.end[1] = p.off2 + p.len2;
			s.ovl = (s.ovl + ((p.off1 > lst)? (uint32_t)p.len1 : p.off1 + p.len1 - lst)) & ZMO_WIN_OVL_MASK;
			lst = p.off1 + p.len1;
		}
		if(s.end[0] - s.beg[0] >= min_len) S.regs[nreg++] = s;
		j = i;
	}
	return nreg;
}

/* hzm_aln.h:933-1054; returns the number of blocks left */
ZMO_HDN uint32_t zmo_merge_blocks(DevWin *regs, uint32_t n, int xvar, int yvar, DotScratch &S){
	uint32_t i, j, k, doff, dcnt, gid, ngrp = 0, nblock; int lst_offset, end_offset;
	zmo_ref_sort(regs, (size_t)n, GtWinDiag());
	for(i = 0; i < n; i++){ S.diags[i].offset = regs[i].beg[0] - regs[i].beg[1]; S.diags[i].off = i; S.diags[i].cnt = 1; }
	doff = 0; end_offset = -0x7FFFFFFF;
	S.grps[ngrp++] = 0;
	while(doff < n){
		lst_offset = S.diags[doff].offset; dcnt = 0;
		while(1){
			if(S.diags[dcnt + doff].offset > lst_offset + yvar) break;
			if(dcnt + doff + 1 >= n) break;
			dcnt++;
		}
		if(dcnt == 0) break;
		if(S.diags[doff + dcnt].offset == end_offset){ doff += dcnt; continue; }
		end_offset = S.diags[doff + dcnt].offset;
		nblock = 0;
		for(i = 0; i < dcnt; i++) S.block[nblock++] = S.diags[i + doff].off;
		{ GtIdxWBeg0 g; g.w = regs; zmo_ref_sort(S.block, (size_t)nblock, g); }
		{
			int s0_end0 = regs[S.block[0]].end[0];
			j = 0;
			for(i = 1; i <= nblock; i++){
				const int s_beg0 = (i == nblock)? 0x7FFFFFFF : regs[S.block[i]].beg[0];
				const int s_end0 = (i == nblock)? 0 : regs[S.block[i]].end[0];
				if(s_beg0 <= s0_end0 + xvar) continue;
				gid = 0;
				for(k = j; k < i; k++){ const uint32_t g = regs[S.block[k]].pb2; if(g){ if(gid == 0) gid = S.grps[g]; else S.grps[g] = gid; } }
				if(gid == 0){ gid = ngrp; S.grps[ngrp++] = gid; }
				for(; j < i; j++) regs[S.block[j]].pb2 = gid;
				j = i; s0_end0 = s_end0;
			}
		}
		for(i = doff; i < doff + dcnt; i++) if(S.diags[i].offset > lst_offset + yvar / 2) break;
		doff = i;
	}
	zmo_tidy_groups(S.grps, ngrp);
	for(i = 0; i < n; i++) if(regs[i].pb2) regs[i].pb2 = S.grps[regs[i].pb2];
	zmo_ref_sort(regs, (size_t)n, GtWinGrpBeg0());
	for(j = 0; j < n; j++) if(regs[j].pb2) break;
	for(i = j + 1; i <= n; i++){
		if(i < n && regs[i].pb2 == regs[j].pb2) continue;
		DevWin &s0 = regs[j];
		for(k = j + 1; k < i; k++){
			DevWin &s = regs[k];
			s.closed = 1;
			if(s.beg[0] < s0.beg[0]) s0.beg[0] = s.beg[0];
			if(s.end[0] > s0.end[0]) s0.end[0] = s.end[0];
			if(s.beg[1] < s0.beg[1]) s0.beg[1] = s.beg[1];
			if(s.end[1] > s0.end[1]) s0.end[1] = s.end[1];
			s0.ovl = (s0.ovl + s.ovl) & ZMO_WIN_OVL_MASK;
		}
		j = i;
	}
	zmo_ref_sort(regs, (size_t)n, GtWinClosed());
	for(i = 0; i < n; i++) if(regs[i].closed) break;
	return i;
}

ZMO_HD int zmo_sx30(int v){ return (int)((uint32_t)v << 2) >> 2; }
/* hzm_aln.h:1056-1132; nodes: 4 ints per block (weight, head, tail, bt) */
ZMO_HDN int zmo_chain_blocks(int len1, int len2, DevWin *r, uint32_t n, int tail_margin, int max_overhang, float band_penalty, float gap_penalty, int *nodes){
	uint32_t i, j; int mw = -1000000, bt = -1, band, gap, weight, W, score;
	zmo_ref_sort(r, (size_t)n, GtWinBeg0());
	for(i = 0; i < n; i++){
		nodes[4 * i] = 0; nodes[4 * i + 1] = 0; nodes[4 * i + 2] = 0; nodes[4 * i + 3] = -1;
		if(r[i].beg[0] <= tail_margin || r[i].beg[1] <= tail_margin) nodes[4 * i + 1] = 1;
		if(r[i].end[0] + tail_margin > len1 || r[i].end[1] + tail_margin > len2) nodes[4 * i + 2] = 1;
	}
	for(i = 0; i < n; i++){
		r[i].closed = 1;
		nodes[4 * i] = zmo_sx30(nodes[4 * i] + (int)r[i].ovl);
		weight = nodes[4 * i] * ((nodes[4 * i + 1] + 3) * (nodes[4 * i + 2] + 3)) / 16;
		if(weight > mw){ mw = weight; bt = (int)i; }
		W = (int)((float)nodes[4 * i] / gap_penalty);
		for(j = i + 1; j < n; j++){
			if(r[j].beg[0] + max_overhang < r[i].end[0]) continue;
			if(r[j].beg[1] + max_overhang < r[i].end[1]) continue;
			if(r[j].beg[0] - r[i].end[0] > W) break;
			const int d0 = r[j].beg[0] - r[i].end[0], d1 = r[j].beg[1] - r[i].end[1];
			band = d0 < d1? d1 - d0 : d0 - d1;
			gap = d0 > d1? d0 : d1;
			if(gap < 0) gap = -gap;
#ifdef __CUDA_ARCH__
			score = (int)__fadd_rn(__fmul_rn((float)band, band_penalty), __fmul_rn((float)gap, gap_penalty));
#else
			{ volatile float a = (float)band * band_penalty, b2 = (float)gap * gap_penalty; score = (int)(a + b2); }
#endif
			score = nodes[4 * i] - score;
			if(nodes[4 * j] <= score){ nodes[4 * j] = zmo_sx30(score); nodes[4 * j + 3] = (int)i; nodes[4 * j + 1] = nodes[4 * i + 1]; }
		}
	}
	mw = 0;
	while(bt >= 0){ r[bt].closed = 0; mw += (int)r[bt].ovl; bt = nodes[4 * bt + 3]; }
	return mw;
}

/* hzm_aln.h:1134-1181 for one pair.  presorted: 0 = cache is in the reference emission order (sort it here),
 * 1 = already sorted by (diagonal, off1) with no tied keys, 2 = sorted but with tied keys: rebuild the emission order
 * and run the exact sort so that the tie permutation equals the reference's */
template<class SORT> ZMO_HDN DotRes zmo_dot_pair(DevZPair *cache, uint32_t n, int alen, int blen, const DotPar &par, uint8_t *scratch, int presorted, const SORT &sorter){
	DotScratch S = zmo_dot_scratch_carve(scratch, n);
	DotRes best[2]; int weight[2];
	if(presorted == 2){ GtZPairEmit g; g.clen = (uint32_t)blen; sorter(cache, (size_t)n, g); }
	if(presorted != 1) sorter(cache, (size_t)n, GtZPairDiag());
	for(uint32_t i = 0; i < n; i++) S.gid[i] = 0;
	for(int d = 0; d < 2; d++){
		uint32_t nreg = zmo_denoise_strand(cache, n, d, par, S, sorter);
		nreg = zmo_merge_blocks(S.regs, nreg, par.xvar, 2 * par.yvar, S);
		weight[d] = zmo_chain_blocks(alen, blen, S.regs, nreg, par.xvar, par.max_overhang, par.deviation_penalty, par.gap_penalty, S.nodes);
		DotRes r; r.score = weight[d]; r.qb = r.tb = 0x7FFFFFFF; r.qe = r.te = 0; r.strand = d;
		for(uint32_t i = 0; i < nreg; i++){
			const DevWin &s = S.regs[i];
			if(s.closed) continue;
			if(r.qb > s.beg[1]) r.qb = s.beg[1];
			if(r.tb > s.beg[0]) r.tb = s.beg[0];
			if(r.qe < s.end[1]) r.qe = s.end[1];
			if(r.te < s.end[0]) r.te = s.end[0];
		}
		best[d] = r;
	}
	return best[weight[0] < weight[1]];
}
ZMO_HDN DotRes zmo_dot_pair(DevZPair *cache, uint32_t n, int alen, int blen, const DotPar &par, uint8_t *scratch, int presorted){
	return zmo_dot_pair(cache, n, alen, blen, par, scratch, presorted, SerialSort());
}
