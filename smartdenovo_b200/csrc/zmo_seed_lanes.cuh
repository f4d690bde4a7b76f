/*
 * zmo_seed_lanes.cuh -- window finding + chaining of the pair-seeding stage with up to 32 PAIRS per warp (merge_paired_kmers_window ..
 * chaining_wtseedv, hzm_aln.h:316-713).  Drop-in for k_p_seed (zmo_seed_kernels.cuh: same arguments, same outputs).
 *
 * k_p_seed gives a pair to a warp and runs the sliding two-cursor scan over the pair's match list (hzm_aln.h:580-656) on all 32 lanes
 * redundantly: 39% of its instructions (profiles/r02_ncu_final.md) are that scalar scan.  Here every lane scans the list of its OWN pair
 * (ls_run: the same scan as a resumable state machine, entries read with one 16-byte load each) until it needs a span search
 * (potential_paired_kmers_windows, hzm_aln.h:410-578) or is done; the warp then serves the pending searches one after the other with all
 * 32 lanes (zmo_windows_in_span_w, unchanged, scratch in shared memory) and hands the result back to the lane that asked.  The scan only
 * needs the NUMBER of windows a search found and the new lower bound, so nothing else crosses between the two levels.  Chaining runs one pair
 * per lane, the copy-out of the kept windows and anchors is cooperative again.  A strand whose search does not fit the shared-memory
 * scratch is redone by its lane with the serial code of zmo_seed_core.cuh on the pair's global scratch, like in k_p_seed.
 */
#pragma once
#include "zmo_seed_kernels.cuh"

enum { LS_FETCH = 0, LS_WAIT, LS_RESUMED, LS_AFTER, LS_DONE };
struct LScan { uint32_t i, j, ol, lst, wlst, p0_off1, p0_len1, p_off1, p_len1, ret, nw; int st, ovf; };

/* both cursors stream through the list one 16-byte entry at a time: ask for the cache line two ahead whenever a line is entered, so that the dependent
 * load of the next entry finds it in L1 instead of paying an L2 round trip per element */
#ifdef __CUDA_ARCH__
#define LS_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" :: "l"(p))
#else
#define LS_PREFETCH(p) ((void)0)
#endif
__device__ __forceinline__ void ls_entry(const DevZPair *rs, uint32_t e, uint32_t &key, uint32_t &len){
	if((e & 7u) == 0) LS_PREFETCH(rs + e + 16);
	const uint4 v = *(const uint4*)(rs + e);                 /* off1 | off2 | len1, len2 | dir1, dir2 */
	key = v.x | (((v.w ^ (v.w >> 8)) & 1u) << 31); len = v.z & 0xFFFFu;
}
/* trailing cursor one entry forward (hzm_aln.h:603-611,630-638): the covered length loses what only the old entry covered */
__device__ __forceinline__ void ls_advance(const DevZPair *rs, LScan &s){
	uint32_t key, len; ls_entry(rs, ++s.j, key, len);
	const uint32_t p1_off1 = key & 0x7FFFFFFFu;
	const uint32_t a = s.p0_off1 > p1_off1? s.p0_off1 : p1_off1;
	const uint32_t t = (s.p0_off1 + s.p0_len1) < (p1_off1 + len)? (s.p0_off1 + s.p0_len1) : (p1_off1 + len);
	s.ol = s.ol + (a < t? t - a : 0u) - s.p0_len1;
	s.p0_off1 = p1_off1; s.p0_len1 = len;
}
__device__ __forceinline__ void ls_init(const DevZPair *rs, uint32_t n, uint32_t dbit, LScan &s){
	uint32_t key = 0, len = 0, j;
	s.i = s.j = s.ol = s.lst = s.wlst = s.ret = s.nw = 0; s.ovf = 0; s.p0_off1 = s.p0_len1 = s.p_off1 = s.p_len1 = 0; s.st = LS_DONE;
	for(j = 0; j < n; j++){ ls_entry(rs, j, key, len); if(!((key ^ dbit) >> 31)) break; }
	if(j == n) return;
	s.p0_off1 = key & 0x7FFFFFFFu; s.p0_len1 = len; s.i = s.j = j; s.st = LS_FETCH;
}
/* the scan of zmo_pair_windows_strand_w (zmo_seed_warp.cuh) for one lane: runs until a span search is due (LS_WAIT: search [j, i) with bound wlst)
 * or the list is exhausted (LS_DONE); re-entered with st = LS_RESUMED and nw / ovf / wlst set by the search */
__device__ void ls_run(const DevZPair *rs, uint32_t n, uint32_t dbit, uint32_t kwin, uint32_t kstep, uint32_t zovl, LScan &s){
	for(;;){
		if(s.st == LS_FETCH){
			if(s.i < n){
				uint32_t key, len; ls_entry(rs, s.i, key, len);
				if((key ^ dbit) >> 31){ s.i++; continue; }
				s.p_off1 = key & 0x7FFFFFFFu; s.p_len1 = len;
			} else { s.p_off1 = 0x1FFFFFu; s.p_len1 = 0x3FFu; }
			if(s.p_off1 > s.p0_off1 + kwin){
				if(s.ol >= zovl){ s.st = LS_WAIT; return; }
				s.st = LS_AFTER;
			} else {
				if(s.p_off1 >= s.lst){ s.ol += s.p_len1; s.lst = s.p_off1 + s.p_len1; }
				else if((int)(s.p_off1 + s.p_len1) > (int)s.lst){ s.ol += s.p_off1 + s.p_len1 - s.lst; s.lst = s.p_off1 + s.p_len1; }
				if(++s.i > n){ s.st = LS_DONE; return; }
				continue;
			}
		}
		if(s.st == LS_RESUMED){
			if(s.nw){ s.ret += s.nw; s.p0_off1 = s.p_off1; s.p0_len1 = s.p_len1; s.ol = s.p_len1; s.lst = s.p_off1 + s.p_len1; s.j = s.i; }
			else if(s.i < n){
				const uint32_t nxt = s.p0_off1 + kstep;
				while(s.p0_off1 < nxt && s.j < s.i) ls_advance(rs, s);
			}
			if(s.ovf){ s.st = LS_DONE; return; }
			s.st = LS_AFTER;
		}
		/* LS_AFTER */
		if(s.p_off1 == 0x1FFFFFu){ s.st = LS_DONE; return; }
		while(s.p_off1 > s.p0_off1 + kwin) ls_advance(rs, s);
		if(++s.i > n){ s.st = LS_DONE; return; }
		s.st = LS_FETCH;
	}
}

__global__ void __launch_bounds__(32 * PS_WARPS) k_p_seed_lanes(const unsigned long long *cache_off, uint32_t np, DevZPair *cache, const uint8_t *tie, const uint32_t *pc, DevReads R,
		uint8_t *scratch, size_t per, uint32_t F, SeedPar par, SeedOut O, zmo_pairseed_t *seeds, unsigned long long *work, uint32_t G){      /* G = pairs per warp, 1..32: the lanes below G scan */
	ZMO_DYN_SMEM(ps_raw);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	PSSmem &M = ((PSSmem*)ps_raw)[warp];
	constexpr unsigned FULL = 0xffffffffu;
	WinScratch WS; WS.ts = M.ts; WS.ak = M.ak; WS.as = (int32_t*)M.ak; WS.wb = M.wb; WS.we = M.we; WS.wo = M.wo; WS.capt = PS_MAXT; WS.capw = PS_MAXW;
	while(1){
		uint32_t base = 0;
		if(lane == 0) base = (uint32_t)atomicAdd(work, (unsigned long long)G);
		base = __shfl_sync(FULL, base, 0);
		if(base >= np) break;
		const uint32_t p = base + (uint32_t)lane; const bool valid = (uint32_t)lane < G && p < np;
		unsigned long long c0 = 0; uint32_t n = 0;
		if(valid){ c0 = cache_off[p]; n = (uint32_t)(cache_off[p + 1] - c0); }
		zmo_pairseed_t S; S.n_zpair = n; S.ovl[0] = S.ovl[1] = 0; S.win_off[0] = S.win_off[1] = 0; S.n_win[0] = S.n_win[1] = 0;
		const bool act = valid && (unsigned long long)n * par.zsize >= par.ztot;
		DevZPair *rs = cache + c0;
		/* the list arrives sorted by (off1,off2); only pairs with tied keys need the reference's exact permutation */
		if(act && tie[p]){ GtZPairEmit g; g.clen = R.len[pc[p]]; zmo_ref_sort(rs, (size_t)n, g); zmo_ref_sort(rs, (size_t)n, GtZPairOff12()); }
		__syncwarp();
		uint8_t *scr = scratch + c0 * per + (size_t)64 * p;
		bool dead = !act;
		for(int d = 0; d < 2; d++){
			PairScratch P = zmo_pair_scratch_carve(scr, n, F);
			uint32_t o_nwin = 0, o_nanc = 0; int o_ovf = 0;
			const uint32_t dbit = (uint32_t)d << 31;
			LScan s; s.st = LS_DONE; s.ret = 0; s.i = s.j = s.wlst = 0;
			if(!dead) ls_init(rs, n, dbit, s);
			for(;;){
				if(s.st != LS_DONE) ls_run(rs, n, dbit, par.kwin, par.kstep, par.zovl, s);
				uint32_t m = __ballot_sync(FULL, s.st == LS_WAIT);
				if(!m) break;
				while(m){
					const int b = __ffs(m) - 1; m &= m - 1;
					const unsigned long long c0b = __shfl_sync(FULL, c0, b);
					const uint32_t nb = __shfl_sync(FULL, n, b), pb = __shfl_sync(FULL, p, b), jb = __shfl_sync(FULL, s.j, b), ib = __shfl_sync(FULL, s.i, b), wl = __shfl_sync(FULL, s.wlst, b);
					const PairScratch Pb = zmo_pair_scratch_carve(scratch + c0b * per + (size_t)64 * pb, nb, F);
					WinOut O2; O2.wins = Pb.w2; O2.nwin = __shfl_sync(FULL, o_nwin, b); O2.capwin = Pb.capw2; O2.anc = Pb.a2; O2.nanc = __shfl_sync(FULL, o_nanc, b); O2.capanc = Pb.cap;
					O2.overflow = 0; O2.stage = M.stage; O2.capstage = PS_STAGE; O2.capwin_ovf = Pb.w2_ovf;
					const uint32_t nw = zmo_windows_in_span_w(cache + c0b, d, jb, ib, (int)wl, O2, WS, par, lane);
					__syncwarp();
					uint32_t wl2 = wl;
					for(uint32_t a = 0; a < nw; a++){ const int e0 = O2.wins[O2.nwin + a - nw].end[0] + 20; if((int)wl2 < e0) wl2 = (uint32_t)e0; }
					if(lane == b){ s.nw = nw; s.wlst = wl2; s.ovf = O2.overflow; s.st = LS_RESUMED; o_nwin = O2.nwin; o_nanc = O2.nanc; o_ovf = O2.overflow; }
					__syncwarp();
				}
			}
			/* chain (hzm_aln.h:658-713): one pair per lane, nodes in the pair's global scratch; a strand that did not fit the shared-memory scratch
			 * is redone from the start by the serial path */
			uint32_t nwin = o_nwin; int ovf = o_ovf, ovl = 0;
			if(!dead){
				if(ovf == 2) ovl = zmo_pair_seed_strand(rs, n, d, par, P, &nwin, &ovf);
				else if(s.ret && !ovf) ovl = zmo_chain_windows(P.w2, nwin, par.W, (int*)P.ws.ts);
				if(ovf){ atomicAdd(O.overflow, 1ULL); dead = true; }
				else S.ovl[d] = ovl;
			}
			__syncwarp();
			/* copy-out of the kept windows and their anchors, one pair after the other, all lanes */
			uint32_t em = __ballot_sync(FULL, !dead && (uint32_t)ovl >= par.ztot);
			while(em){
				const int b = __ffs(em) - 1; em &= em - 1;
				const unsigned long long c0b = __shfl_sync(FULL, c0, b);
				const uint32_t nb = __shfl_sync(FULL, n, b), pb = __shfl_sync(FULL, p, b), nwb = __shfl_sync(FULL, nwin, b);
				const PairScratch Pb = zmo_pair_scratch_carve(scratch + c0b * per + (size_t)64 * pb, nb, F);
				const DevWin *W2 = Pb.w2;
				uint32_t kw = 0, ka = 0;
				for(uint32_t j = lane; j < nwb; j += 32) if(!W2[j].closed){ kw++; ka += W2[j].anc1 - W2[j].anc0; }
				kw = __reduce_add_sync(FULL, kw); ka = __reduce_add_sync(FULL, ka);
				unsigned long long w0 = 0, a0 = 0;
				if(lane == 0){ w0 = atomicAdd(O.cur_wins, (unsigned long long)kw); a0 = atomicAdd(O.cur_anc, (unsigned long long)ka); }
				w0 = __shfl_sync(FULL, w0, 0); a0 = __shfl_sync(FULL, a0, 0);
				if(w0 + kw > O.cap_wins || a0 + ka > O.cap_anc){ if(lane == 0) atomicAdd(O.overflow, 1ULL); if(lane == b) dead = true; continue; }
				unsigned long long wi = w0, ai = a0;
				for(uint32_t j = 0; j < nwb; j++){
					DevWin w = W2[j];
					if(w.closed) continue;
					const uint32_t na = w.anc1 - w.anc0;
					for(uint32_t k = lane; k < na; k += 32) O.anc[ai + k] = Pb.a2[w.anc0 + k];
					w.anc0 = (uint32_t)ai; w.anc1 = (uint32_t)(ai + na); ai += na;
					if(lane == 0) O.wins[wi] = w;
					wi++;
				}
				if(lane == b){ S.win_off[d] = (uint32_t)w0; S.n_win[d] = kw; }
			}
			__syncwarp();
		}
		if(valid) seeds[p] = S;
		__syncwarp();
	}
}
