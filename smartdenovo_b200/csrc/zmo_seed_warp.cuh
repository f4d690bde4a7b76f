/*
 * zmo_seed_warp.cuh -- warp-cooperative form of the per-pair window search (device only).
 *
 * Same results as the serial functions of zmo_seed_core.cuh (which remain the exact fallback and the code the CPU tests
 * exercise), but the expensive inner steps of potential_paired_kmers_windows (hzm_aln.h:410-578) are done by all 32
 * lanes: ordered compaction of the span's strand entries into packed sort keys, a bitonic sort of the keys, ordered
 * gathering of the anchors, their key sort, and the union-length / bounding-box measurements (a map over adjacent
 * elements + reductions).  The reference's sorts are unstable but only ties can expose that: the parallel sort orders
 * by the full unique key, and whenever two adjacent keys tie on the compared field the keys are rebuilt in input order
 * and lane 0 runs the exact sort_array emulation instead.  Control flow is uniform: every lane executes the scalar
 * bookkeeping redundantly on the same values.
 */
#pragma once
#include "zmo_seed_core.cuh"

/* ascending bitonic sort of n2 (power of two, >= 32 ... also works below) 64-bit keys in shared memory by one warp */
__device__ __forceinline__ void zmo_warp_bitonic_u64(uint64_t *a, uint32_t n2, int lane){
	for(uint32_t k = 2; k <= n2; k <<= 1){
		for(uint32_t j = k >> 1; j > 0; j >>= 1){
			for(uint32_t i = lane; i < n2; i += 32){
				const uint32_t l = i ^ j;
				if(l > i){
					const uint64_t x = a[i], y = a[l];
					const bool up = (i & k) == 0;
					if((x > y) == up){ a[i] = y; a[l] = x; }
				}
			}
			__syncwarp();
		}
	}
}
__device__ __forceinline__ uint32_t zmo_pow2_ge(uint32_t n){ uint32_t p = 1; while(p < n) p <<= 1; return p; }

/* k-th smallest (0-based) of n int32 values in shared memory, |value| < 2^24 (diagonals off1 - off2 of reads shorter than 2^24), by all
 * 32 lanes: radix select, one vote per bit.  calculate_median_value (hzm_aln.h:316-343) is a quickselect that returns element size/2 of
 * the sorted order -- a VALUE, so any selection algorithm gives the same result (its in-place permutation of the scratch array is never
 * looked at again). */
__device__ __forceinline__ int32_t zmo_warp_kth_i32(const int32_t *a, uint32_t n, uint32_t k, int lane){
	uint32_t prefix = 0, mask = 0;
	#pragma unroll 1
	for(int bit = 24; bit >= 0; bit--){
		const uint32_t m2 = mask | (1u << bit);
		uint32_t c = 0;
		for(uint32_t i = lane; i < n; i += 32) c += (((uint32_t)(a[i] + (1 << 24)) & m2) == prefix)? 1u : 0u;
		c = __reduce_add_sync(0xffffffffu, c);
		if(k >= c){ k -= c; prefix |= 1u << bit; }
		mask = m2;
	}
	return (int32_t)prefix - (1 << 24);
}

/* warp-cooperative windows_in_span; S arrays must live in shared memory with capt a power of two >= the span's strand
 * entries and S.ak holding at least pow2(capstage) keys.  Returns like the serial version; O.overflow == 2 asks the
 * caller to redo the strand with the serial global-memory path. */
__device__ uint32_t zmo_windows_in_span_w(const DevZPair *rs, int dir, uint32_t beg, uint32_t end, int bound, WinOut &O, const WinScratch &S, const SeedPar &par, int lane){
	const uint32_t zsize = par.zsize, kwin = par.kwin, zovl = par.zovl;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint32_t n = 0, n2 = 0, ret = 0;
	/* leading entries of the other strand or before `bound` are skipped (hzm_aln.h:424-429) */
	{
		uint32_t first = end;
		for(uint32_t base = beg; base < end; base += 32){
			const uint32_t idx = base + lane; bool stop = false;
			if(idx < end){ const DevZPair p = rs[idx]; stop = !((p.dir1 ^ p.dir2 ^ dir) || (int)p.off1 < bound); }
			const uint32_t m = __ballot_sync(0xffffffffu, stop);
			if(m){ first = base + (uint32_t)__ffs(m) - 1; break; }
		}
		beg = first;
	}
	if(end - beg >= (1u << 24)){ O.overflow = 2; return 0; }
	/* ordered compaction of the strand's entries into sort keys (off2 | len2 | index), in list order */
	for(int pass = 0; pass < 2; pass++){
		n = 0;
		for(uint32_t base = beg; base < end; base += 32){
			const uint32_t idx = base + lane; bool ok = false; uint64_t key = 0;
			if(idx < end){ const DevZPair p = rs[idx]; ok = !(p.dir1 ^ p.dir2 ^ dir); key = ((uint64_t)p.off2 << 40) | ((uint64_t)p.len2 << 24) | (idx - beg); }
			const uint32_t m = __ballot_sync(0xffffffffu, ok);
			if(pass && ok) S.ts[n + __popc(m & lt_mask)] = key;
			n += __popc(m);
		}
		if(pass == 0){
			if(n * zsize < zovl) return 0;
			if(n > S.capt){ O.overflow = 2; return 0; }
		}
	}
	__syncwarp();
	{
		const uint32_t np2 = zmo_pow2_ge(n);
		for(uint32_t i = n + lane; i < np2; i += 32) S.ts[i] = ~0ull;
		__syncwarp();
		zmo_warp_bitonic_u64(S.ts, np2, lane);
		bool tie = false;
		for(uint32_t i = lane + 1; i < n; i += 32) if((S.ts[i] >> 40) == (S.ts[i - 1] >> 40)) tie = true;
		if(__any_sync(0xffffffffu, tie)){
			/* equal off2 inside the span: the reference's unstable sort decides the order -> rebuild the input order, exact sort */
			__syncwarp();
			uint32_t c = 0;
			for(uint32_t base = beg; base < end; base += 32){
				const uint32_t idx = base + lane; bool ok = false; uint64_t key = 0;
				if(idx < end){ const DevZPair p = rs[idx]; ok = !(p.dir1 ^ p.dir2 ^ dir); key = ((uint64_t)p.off2 << 40) | ((uint64_t)p.len2 << 24) | (idx - beg); }
				const uint32_t m = __ballot_sync(0xffffffffu, ok);
				if(ok) S.ts[c + __popc(m & lt_mask)] = key;
				c += __popc(m);
			}
			__syncwarp();
			if(lane == 0) zmo_ref_sort(S.ts, (size_t)n, GtTsKey());
			__syncwarp();
		}
	}
	/* sub-windows along c (hzm_aln.h:450-481).  The reference slides two cursors over the off2-sorted entries and keeps a running covered length
	 *   ol(i) = sum_{k<=i} inc(k) - sum_{k<J(i)} dec(k)   (uint32, wraps like the reference's)
	 * with inc(k) = len_k or end_k - end_{k-1} (one look at the predecessor), dec(k) = len_k - overlap(k, k+1) (one look at the successor), and the
	 * trailing cursor J(i) = min(n-1, max_{i'<=i} #{j : off_j + kwin < end_i'}).  All three are scans / searches every lane can take part in; only the
	 * merge rule of neighbouring sub-windows is order dependent and stays on lane 0, reading ol(i) and J(i) from shared memory. */
	const bool par_scan = (size_t)n * 4 <= (size_t)zmo_pow2_ge(O.capstage) * 8 && (size_t)n * 6 <= (size_t)O.capstage * sizeof(DevZPair) && O.stage != nullptr;
	if(par_scan){
		uint32_t *dsum = (uint32_t*)S.ak, *olv = (uint32_t*)O.stage; uint16_t *jv = (uint16_t*)(olv + n);
		uint32_t carry = 0;
		for(uint32_t base = 0; base < n; base += 32){          /* dsum[k] = sum_{k'<k} dec(k') */
			const uint32_t k = base + lane; uint32_t v = 0;
			if(k + 1 < n){ const uint64_t k0 = S.ts[k], k1 = S.ts[k + 1]; const uint32_t s_ = ZMO_TS_OFF2(k1), t_ = ZMO_TS_OFF2(k0) + ZMO_TS_LEN2(k0); v = ZMO_TS_LEN2(k0) - (s_ < t_? t_ - s_ : 0u); }
			uint32_t inc = v;
			#pragma unroll
			for(int d = 1; d < 32; d <<= 1){ const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if(lane >= d) inc += o; }
			if(k < n) dsum[k] = carry + inc - v;
			carry += __shfl_sync(0xffffffffu, inc, 31);
		}
		__syncwarp();
		uint32_t csum = 0, cmax = 0;
		for(uint32_t base = 0; base < n; base += 32){
			const uint32_t i = base + lane; uint32_t v = 0, T = 0;
			if(i < n){
				const uint64_t pk = S.ts[i]; const uint32_t off = ZMO_TS_OFF2(pk), len = ZMO_TS_LEN2(pk), lst = i? ZMO_TS_OFF2(S.ts[i - 1]) + ZMO_TS_LEN2(S.ts[i - 1]) : 0u;
				v = (off > lst)? len : off + len - lst;
				uint32_t lo = 0, hi = n;                           /* T = #{j : off_j + kwin < off + len} (off_j ascending) */
				while(lo < hi){ const uint32_t mid = (lo + hi) >> 1; if(ZMO_TS_OFF2(S.ts[mid]) + kwin < off + len) lo = mid + 1; else hi = mid; }
				T = lo;
			}
			uint32_t inc = v, mx = T;
			#pragma unroll
			for(int d = 1; d < 32; d <<= 1){
				const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d), m_ = __shfl_up_sync(0xffffffffu, mx, d);
				if(lane >= d){ inc += o; if(m_ > mx) mx = m_; }
			}
			if(cmax > mx) mx = cmax;
			if(i < n){ const uint32_t J = mx < n - 1? mx : n - 1; olv[i] = csum + inc - dsum[J]; jv[i] = (uint16_t)J; }
			csum += __shfl_sync(0xffffffffu, inc, 31); cmax = __shfl_sync(0xffffffffu, mx, 31);
		}
		__syncwarp();
		if(lane == 0){
			int ovf = 0;
			for(uint32_t i = 0; i < n; i++){
				const uint32_t ol = olv[i];
				if(ol >= zovl){
					const uint32_t j = jv[i], p_off2 = ZMO_TS_OFF2(S.ts[i]);
					if(n2 && ( p_off2 <= ZMO_TS_OFF2(S.ts[S.we[n2-1]]) + kwin / 3 || ZMO_TS_OFF2(S.ts[j]) <= ZMO_TS_OFF2(S.ts[S.wb[n2-1]]) + kwin / 3 )){
						if(ol > S.wo[n2-1]){ S.wb[n2-1] = j; S.we[n2-1] = i; S.wo[n2-1] = ol; }
					} else { if(n2 >= S.capw){ ovf = 1; break; } S.wb[n2] = j; S.we[n2] = i; S.wo[n2] = ol; n2++; }
				}
			}
			if(ovf) n2 = 0xFFFFFFFFu;
		}
	} else { O.overflow = 2; return 0; }      /* scratch too small for the scan arrays: the strand is redone by the serial path */
	n2 = __shfl_sync(0xffffffffu, n2, 0);
	__syncwarp();
	if(n2 == 0xFFFFFFFFu){ O.overflow = 2; return 0; }
	for(uint32_t i = 0; i < n2; i++){
		const uint32_t size = O.nanc, wb = S.wb[i], we = S.we[i], offn = we - wb + 1;
		int32_t offset;
		for(uint32_t j = wb + lane; j <= we; j += 32){ const DevZPair p = rs[beg + ZMO_TS_IDX(S.ts[j])]; S.as[j - wb] = (int32_t)p.off1 - (int32_t)p.off2; }
		__syncwarp();
		offset = zmo_warp_kth_i32(S.as, offn, offn / 2, lane);
		/* anchors within +-50 of the median diagonal, gathered in span order */
		uint32_t na = 0;
		for(int pass = 0; pass < 2; pass++){
			na = 0;
			for(uint32_t base = wb; base <= we; base += 32){
				const uint32_t j = base + lane; bool ok = false; DevZPair p;
				if(j <= we){ p = rs[beg + ZMO_TS_IDX(S.ts[j])]; const int32_t off = (int32_t)p.off1 - (int32_t)p.off2; ok = !(off < offset - ZMO_KWIN_MAX_OFFSET_DEV || off > offset + ZMO_KWIN_MAX_OFFSET_DEV); }
				const uint32_t m = __ballot_sync(0xffffffffu, ok);
				if(pass && ok){ const uint32_t k = na + __popc(m & lt_mask); O.stage[k] = p; S.ak[k] = ((uint64_t)p.off1 << 32) | k; }
				na += __popc(m);
			}
			if(pass == 0){
				if(na == 0) break;
				if(size + na > O.capanc){ O.overflow = 1; return ret; }
				if(na > O.capstage){ O.overflow = 2; return ret; }
			}
		}
		if(na == 0) continue;
		__syncwarp();
		{
			const uint32_t np2 = zmo_pow2_ge(na);
			for(uint32_t k = na + lane; k < np2; k += 32) S.ak[k] = ~0ull;
			__syncwarp();
			zmo_warp_bitonic_u64(S.ak, np2, lane);
			bool tie = false;
			for(uint32_t k = lane + 1; k < na; k += 32) if((S.ak[k] >> 32) == (S.ak[k - 1] >> 32)) tie = true;
			if(__any_sync(0xffffffffu, tie)){
				__syncwarp();
				for(uint32_t k = lane; k < na; k += 32) S.ak[k] = ((uint64_t)O.stage[k].off1 << 32) | k;
				__syncwarp();
				if(lane == 0) zmo_ref_sort(S.ak, (size_t)na, GtHi32());
				__syncwarp();
			}
		}
		/* union length along q + bounding box in sorted order: each element only needs its predecessor */
		uint32_t ol = 0; int b0 = 0x7FFFFFFF, b1 = 0x7FFFFFFF, e0 = 0, e1 = 0;
		for(uint32_t k = lane; k < na; k += 32){
			const DevZPair p = O.stage[(uint32_t)S.ak[k]];
			uint32_t lst = 0;
			if(k){ const DevZPair q = O.stage[(uint32_t)S.ak[k - 1]]; lst = q.off1 + q.len1; }
			ol += (p.off1 > lst)? (uint32_t)p.len1 : p.off1 + p.len1 - lst;
			if((int)p.off1 < b0) b0 = p.off1;
			if((int)(p.off1 + p.len1) > e0) e0 = p.off1 + p.len1;
			if((int)p.off2 < b1) b1 = p.off2;
			if((int)(p.off2 + p.len2) > e1) e1 = p.off2 + p.len2;
		}
		ol = __reduce_add_sync(0xffffffffu, ol);
		b0 = __reduce_min_sync(0xffffffffu, b0); b1 = __reduce_min_sync(0xffffffffu, b1);
		e0 = __reduce_max_sync(0xffffffffu, e0); e1 = __reduce_max_sync(0xffffffffu, e1);
		if(ol * 2 < zovl) continue;
		if(ret){
			const DevWin w = O.wins[O.nwin - 1];
			if(e1 <= (int)(w.end[1] + kwin / 3) && ol <= w.ovl) continue;
		}
		if(O.nwin >= O.capwin){ O.overflow = O.capwin_ovf; return ret; }
		for(uint32_t k = lane; k < na; k += 32) O.anc[size + k] = O.stage[(uint32_t)S.ak[k]];
		O.nanc = size + na;
		ret++;
		if(lane == 0){
			DevWin W0; W0.closed = 0; W0.dir = (uint8_t)dir; W0.pad = 0; W0.pb2 = 0; W0.anc0 = size; W0.anc1 = O.nanc;
			W0.beg[0] = b0; W0.beg[1] = b1; W0.end[0] = e0; W0.end[1] = e1; W0.ovl = ol & ZMO_WIN_OVL_MASK;
			O.wins[O.nwin] = W0;
		}
		O.nwin++;
		__syncwarp();
	}
	return ret;
}

/* Register-resident window over the pair's match list for the scalar scan below: every lane keeps one entry of the
 * current 32-entry chunk as (off1 | strand-xor << 31, len1); the scan reads entry idx by shuffle from lane idx & 31, so
 * a chunk costs ONE coalesced 16-byte load per lane instead of one dependent global-memory round trip per element.
 * Two chunks are kept, one per cursor of the two-pointer scan (all arguments are warp-uniform). */
struct ZChunk { uint32_t base, key, len; };
__device__ __forceinline__ void zmo_chunk_get(const DevZPair *rs, uint32_t n, ZChunk &C, uint32_t idx, int lane, uint32_t &key, uint32_t &len){
	if(idx - C.base >= 32u){
		C.base = idx & ~31u;
		const uint32_t e = C.base + lane;
		if(e < n){ const uint4 v = *(const uint4*)(rs + e); C.key = v.x | (((v.w ^ (v.w >> 8)) & 1u) << 31); C.len = v.z & 0xFFFFu; }
	}
	key = __shfl_sync(0xffffffffu, C.key, idx & 31); len = __shfl_sync(0xffffffffu, C.len, idx & 31);
}

/* hzm_aln.h:580-656 with every lane running the (cheap) scalar scan redundantly and the span searches done cooperatively */
__device__ uint32_t zmo_pair_windows_strand_w(const DevZPair *rs, uint32_t n, int dir, WinOut &O, const WinScratch &S, const SeedPar &par, int lane){
	const uint32_t kwin = par.kwin, kstep = par.kstep, zovl = par.zovl, dbit = (uint32_t)dir << 31;
	uint32_t i, j, a, nw, ol = 0, ol2, lst = 0, wlst = 0, s, t, ret = 0;
	uint32_t p0_off1, p0_len1, p_off1, p_len1, key, len;
	ZChunk Ci, Cj; Cj.base = 0xFFFFFF00u; Cj.key = Cj.len = 0;
	for(j = 0; j < n; j++){ zmo_chunk_get(rs, n, Cj, j, lane, key, len); if(!((key ^ dbit) >> 31)) break; }
	if(j == n) return 0;
	p0_off1 = key & 0x7FFFFFFFu; p0_len1 = len;
	Ci = Cj;
	for(i = j; i <= n; i++){
		if(i < n){
			zmo_chunk_get(rs, n, Ci, i, lane, key, len);
			if((key ^ dbit) >> 31) continue;
			p_off1 = key & 0x7FFFFFFFu; p_len1 = len;
		} else { p_off1 = 0x1FFFFFu; p_len1 = 0x3FFu; }
		if(p_off1 > p0_off1 + kwin){
			if(ol >= zovl){
				if((nw = zmo_windows_in_span_w(rs, dir, j, i, (int)wlst, O, S, par, lane))){
					for(a = 0; a < nw; a++){ const int e0 = O.wins[O.nwin + a - nw].end[0] + 20; if((int)wlst < e0) wlst = e0; }
					ret += nw;
					p0_off1 = p_off1; p0_len1 = p_len1;
					ol = p_len1; lst = p_off1 + p_len1; j = i; Cj = Ci;
				} else if(i < n){
					const uint32_t nxt = p0_off1 + kstep;
					while(p0_off1 < nxt && j < i){
						zmo_chunk_get(rs, n, Cj, ++j, lane, key, len);
						const uint32_t p1_off1 = key & 0x7FFFFFFFu;
						s = p0_off1 > p1_off1? p0_off1 : p1_off1;
						t = (p0_off1 + p0_len1) < (p1_off1 + len)? (p0_off1 + p0_len1) : (p1_off1 + len);
						ol2 = s < t? t - s : 0;
						ol = ol + ol2 - p0_len1;
						p0_off1 = p1_off1; p0_len1 = len;
					}
				}
				if(O.overflow) return ret;
			}
			if(p_off1 == 0x1FFFFFu) break;
			while(p_off1 > p0_off1 + kwin){
				zmo_chunk_get(rs, n, Cj, ++j, lane, key, len);
				const uint32_t p1_off1 = key & 0x7FFFFFFFu;
				s = p0_off1 > p1_off1? p0_off1 : p1_off1;
				t = (p0_off1 + p0_len1) < (p1_off1 + len)? (p0_off1 + p0_len1) : (p1_off1 + len);
				ol2 = s < t? t - s : 0;
				ol = ol + ol2 - p0_len1;
				p0_off1 = p1_off1; p0_len1 = len;
			}
		} else {
			if(p_off1 >= lst) ol += p_len1;
			else if((int)(p_off1 + p_len1) > (int)lst) ol += p_off1 + p_len1 - lst;
			else continue;
			lst = p_off1 + p_len1;
		}
	}
	return ret;
}

/* strand driver: windows (cooperative) + chain (lane 0); all lanes return the same values */
__device__ int zmo_pair_seed_strand_w(const DevZPair *rs, uint32_t n, int dir, const SeedPar &par, PairScratch &P, uint32_t *nwin, int *overflow, int lane){
	WinOut O; O.wins = P.w2; O.nwin = 0; O.capwin = P.capw2; O.anc = P.a2; O.nanc = 0; O.capanc = P.cap; O.overflow = 0; O.stage = P.stage; O.capstage = P.capstage; O.capwin_ovf = P.w2_ovf;
	int ovl = 0;
	const uint32_t got = zmo_pair_windows_strand_w(rs, n, dir, O, P.ws, par, lane);
	__syncwarp();
	if(got && !O.overflow && O.nwin > P.ws.capt) O.overflow = 2;      /* chain nodes (2 ints per window) live in ts */
	if(got && !O.overflow){
		if(lane == 0) ovl = zmo_chain_windows(P.w2, O.nwin, par.W, (int*)P.ws.ts);
		ovl = __shfl_sync(0xffffffffu, ovl, 0);
		__syncwarp();
	}
	*nwin = O.nwin; *overflow = O.overflow;
	return ovl;
}
