/*
 * zmo_seed_kernels.cuh -- the warp-per-pair window finding + chaining kernel of the pair-seeding stage (merge_paired_kmers_window ..
 * chaining_wtseedv, hzm_aln.h:316-713).  Kept in a header so that the test-only host simulation (tests/hostsim/seedk_host.cpp) runs
 * this very source; included by zmo_seed.cu only.
 */
#pragma once
#include "zmo_ctx.cuh"
#include "zmo_seed_core.cuh"
#include "zmo_seed_warp.cuh"

/* dynamic shared memory of a kernel (the host simulation substitutes a per-block buffer) */
#ifndef ZMO_DYN_SMEM
#define ZMO_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

struct SeedOut { DevWin *wins; DevZPair *anc; unsigned long long cap_wins, cap_anc; unsigned long long *cur_wins, *cur_anc, *overflow; };
/* Window finding + chaining, one WARP per pair (zmo_seed_warp.cuh): the sliding scan runs on all lanes over register
 * chunks of the match list, span searches are warp-cooperative with their sort keys / anchors staged in shared memory, the
 * chain runs on lane 0, the lanes copy the kept windows/anchors out.  A span, window or strand that does not fit the
 * shared-memory scratch is redone by lane 0 with the serial code of zmo_seed_core.cuh on global scratch. */
#define PS_WARPS 22
#define PS_MAXT 512       /* strand entries of one sliding span */
#define PS_MAXW 96       /* sub-windows of one span */
#define PS_STAGE 192      /* anchors of one window */
/* per-warp scratch in shared memory (~10 KB, 22 warps per SM at 80 registers: the scalar parts are latency bound, so occupancy
 * is what buys throughput); larger spans / windows fall back to the global-memory scratch; the strand's windows live in the
 * pair's global scratch; the read-only match list is read through register chunks (zmo_seed_warp.cuh) */
struct PSSmem { uint64_t ts[PS_MAXT]; uint64_t ak[256]; DevZPair stage[PS_STAGE]; uint32_t wb[PS_MAXW], we[PS_MAXW], wo[PS_MAXW]; int bc[8]; };     /* as (median scratch) aliases ak: it is dead before the anchors are gathered */
__global__ void __launch_bounds__(32 * PS_WARPS) k_p_seed(const unsigned long long *cache_off, uint32_t np, DevZPair *cache, const uint8_t *tie, const uint32_t *pc, DevReads R,
		uint8_t *scratch, size_t per, uint32_t F, SeedPar par, SeedOut O, zmo_pairseed_t *seeds, unsigned long long *work){
	ZMO_DYN_SMEM(ps_raw);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	PSSmem &M = ((PSSmem*)ps_raw)[warp];
	while(1){
		uint32_t p = 0;
		if(lane == 0) p = (uint32_t)atomicAdd(work, 1ULL);
		p = __shfl_sync(0xffffffffu, p, 0);
		if(p >= np) break;
		const unsigned long long c0 = cache_off[p]; const uint32_t n = (uint32_t)(cache_off[p + 1] - c0);
		zmo_pairseed_t S; S.n_zpair = n; S.ovl[0] = S.ovl[1] = 0; S.win_off[0] = S.win_off[1] = 0; S.n_win[0] = S.n_win[1] = 0;
		if((unsigned long long)n * par.zsize >= par.ztot){
			DevZPair *rs = cache + c0;
			/* the list arrives sorted by (off1,off2); only pairs with tied keys need the reference's exact permutation */
			if(tie[p] && lane == 0){ GtZPairEmit g; g.clen = R.len[pc[p]]; zmo_ref_sort(rs, (size_t)n, g); zmo_ref_sort(rs, (size_t)n, GtZPairOff12()); }
			__syncwarp();
			uint8_t *scr = scratch + c0 * per + (size_t)64 * p;
			for(int d = 0; d < 2; d++){
				PairScratch P = zmo_pair_scratch_carve(scr, n, F);
				uint32_t nwin = 0; int ovf = 0, ovl = 0;
				{
					/* cooperative path: scratch in shared memory, all lanes */
					PairScratch Q = P; Q.ws.ts = M.ts; Q.ws.ak = M.ak; Q.ws.as = (int32_t*)M.ak; Q.ws.wb = M.wb; Q.ws.we = M.we; Q.ws.wo = M.wo; Q.ws.capt = PS_MAXT; Q.ws.capw = PS_MAXW;
					Q.stage = M.stage; Q.capstage = PS_STAGE;      /* windows go to the pair's global scratch (P.w2) */
					ovl = zmo_pair_seed_strand_w(rs, n, d, par, Q, &nwin, &ovf, lane);
				}
				const bool fast = ovf != 2;
				if(!fast){
					/* a span / window / strand exceeded the shared-memory scratch: serial exact path on global scratch */
					if(lane == 0){ ovl = zmo_pair_seed_strand(rs, n, d, par, P, &nwin, &ovf); M.bc[0] = ovl; M.bc[1] = (int)nwin; M.bc[2] = ovf; }
					__syncwarp();
					ovl = M.bc[0]; nwin = (uint32_t)M.bc[1]; ovf = M.bc[2];
					__syncwarp();
				}
				const DevWin *W2 = P.w2;
				if(ovf){ if(lane == 0) atomicAdd(O.overflow, 1ULL); break; }
				S.ovl[d] = ovl;
				if((uint32_t)ovl >= par.ztot){
					uint32_t kw = 0, ka = 0;
					for(uint32_t j = 0; j < nwin; j++) if(!W2[j].closed){ kw++; ka += W2[j].anc1 - W2[j].anc0; }
					unsigned long long w0 = 0, a0 = 0;
					if(lane == 0){ w0 = atomicAdd(O.cur_wins, (unsigned long long)kw); a0 = atomicAdd(O.cur_anc, (unsigned long long)ka); }
					w0 = __shfl_sync(0xffffffffu, w0, 0); a0 = __shfl_sync(0xffffffffu, a0, 0);
					if(w0 + kw > O.cap_wins || a0 + ka > O.cap_anc){ if(lane == 0) atomicAdd(O.overflow, 1ULL); break; }
					unsigned long long wi = w0, ai = a0;
					for(uint32_t j = 0; j < nwin; j++){
						DevWin w = W2[j];
						if(w.closed) continue;
						const uint32_t na = w.anc1 - w.anc0;
						for(uint32_t k = lane; k < na; k += 32) O.anc[ai + k] = P.a2[w.anc0 + k];
						w.anc0 = (uint32_t)ai; w.anc1 = (uint32_t)(ai + na); ai += na;
						if(lane == 0) O.wins[wi] = w;
						wi++;
					}
					S.win_off[d] = (uint32_t)w0; S.n_win[d] = kw;
				}
				__syncwarp();
			}
		}
		if(lane == 0) seeds[p] = S;
		__syncwarp();
	}
}
