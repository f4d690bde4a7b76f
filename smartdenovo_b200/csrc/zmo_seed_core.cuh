/*
 * zmo_seed_core.cuh -- per-read / per-pair seeding logic as plain host+device functions.
 *
 * These are the irregular, order-sensitive parts of the path (k-/z-mer scanning, z-mer matching,
 * the sort_array-exact sort, window finding, median diagonal, window chaining).  In the product they
 * run inside CUDA kernels, one thread per read or per pair, on per-thread slices of a device arena
 * (zmo_seed.cu).  They are written without CUDA intrinsics so that the SAME source can also be
 * compiled for the host by the test-only harness tests/hostsim/ and compared with the oracle on a
 * CPU-only box; the product library never calls them on the host.
 *
 * Restated behaviour: wtzmo.c:246-271 (k-mer scan), hzm_aln.h:70-224 (z-index, z-match),
 * sort.h:104-155 (sort_array), hzm_aln.h:316-343 (median), :410-578 (windows in a span),
 * :580-656 (sliding window), :658-713 (window chain).
 */
#pragma once
#include <stdint.h>
#include <stddef.h>
#ifdef __CUDACC__
#define ZMO_HD __host__ __device__ __forceinline__
#define ZMO_HDN __host__ __device__ inline
#define ZMO_HDM __host__ __device__ __forceinline__
#else
#define ZMO_HD static inline
#define ZMO_HDN static
#define ZMO_HDM inline
#endif

struct DevZPair { uint32_t off1, off2; uint16_t len1, len2; uint8_t dir1, dir2; uint16_t pad; };
struct DevWin { int32_t beg[2], end[2]; uint32_t ovl, anc0, anc1; uint8_t dir, closed; uint16_t pad; uint32_t pb2; };
struct DevSlot { uint32_t mer, off, cnt; };
struct DevZSeed { uint32_t off; uint16_t len; uint8_t dir, pad; };

ZMO_HD uint32_t rd_base(const uint32_t *w, uint32_t p){ return (w[p >> 4] >> (((~p) & 15u) << 1)) & 3u; }

ZMO_HD uint64_t zmo_kmer_revcomp(uint64_t x, int k){          /* dna.h:85-97 */
	x = ~x;
	x = ((x & 0x3333333333333333ULL) << 2) | ((x >> 2) & 0x3333333333333333ULL);
	x = ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL);
	x = ((x & 0x00FF00FF00FF00FFULL) << 8) | ((x >> 8) & 0x00FF00FF00FF00FFULL);
	x = ((x & 0x0000FFFF0000FFFFULL) << 16) | ((x >> 16) & 0x0000FFFF0000FFFFULL);
	x = (x << 32) | (x >> 32);
	return x >> (64 - 2 * k);
}
ZMO_HD uint32_t zmo_jenkins32(uint32_t key){                  /* hashset.h:452-462 */
	key += (key << 12); key ^= (key >> 22); key += (key << 4); key ^= (key >> 9);
	key += (key << 10); key ^= (key >> 2); key += (key << 7); key ^= (key >> 12);
	return key;
}
ZMO_HD bool zmo_kmer_sampled(uint64_t mer, uint32_t ksave){ return zmo_jenkins32((uint32_t)mer) % (1024u * ksave) < 1024u; }   /* wtzmo.c:270-271 */

/* Scan the homopolymer-compressed canonical k-mers of a read (wtzmo.c:246-271 / hzm_aln.h:83-99).
 * f(mer, dir, off, len) is called for every non-palindromic k-mer; off = position of the k-mer's
 * first kept base, len = span up to and including the current kept base (capped at 0xFFFF). */
template<class F> ZMO_HD void zmo_scan_kmers(const uint32_t *words, uint32_t len, int k, int hp, F f){
	const uint64_t kmask = 0xFFFFFFFFFFFFFFFFULL >> ((32 - k) << 1);
	uint64_t kmer = 0; uint32_t ring[32]; uint32_t kept = 0, b = 4;
	for(uint32_t j = 0; j < len; j++){
		const uint32_t c = rd_base(words, j);
		if(hp && c == b) continue;
		b = c; ring[kept & 31] = j; kept++;
		kmer = ((kmer << 2) | b) & kmask;
		if(kept < (uint32_t)k) continue;
		const uint64_t krev = zmo_kmer_revcomp(kmer, k);
		if(krev == kmer) continue;
		const uint32_t off = ring[(kept - k) & 31];
		const uint32_t ln = (j + 1 - off > 0xFFFFu)? 0xFFFFu : j + 1 - off;
		f(krev > kmer? kmer : krev, (uint32_t)(krev > kmer? 0 : 1), off, ln);
	}
}

/* Chunk-parallel form of zmo_scan_kmers: emits exactly the k-mers whose current (last) kept base lies in
 * [s, e), so that disjoint chunks of one read can be scanned by independent threads.  The state at s
 * is recovered by walking back until k-1 kept bases are collected (a base is kept iff it differs from
 * its predecessor, which is a local rule). */
template<class F> ZMO_HD void zmo_scan_kmers_chunk(const uint32_t *words, uint32_t len, int k, int hp, uint32_t s, uint32_t e, F f){
	const uint64_t kmask = 0xFFFFFFFFFFFFFFFFULL >> ((32 - k) << 1);
	uint64_t kmer = 0; uint32_t ring[32]; uint32_t kept = 0, b = 4, p = s; int need = k - 1;
	if(e > len) e = len;
	if(s >= e) return;
	while(p > 0 && need > 0){
		p--;
		if(!hp || p == 0 || rd_base(words, p) != rd_base(words, p - 1)) need--;
	}
	/* p is a kept position (or 0), so starting with "no previous base" reproduces the global scan */
	for(uint32_t j = p; j < e; j++){
		const uint32_t c = rd_base(words, j);
		if(hp && c == b) continue;
		b = c; ring[kept & 31] = j; kept++;
		kmer = ((kmer << 2) | b) & kmask;
		if(j < s || kept < (uint32_t)k) continue;
		const uint64_t krev = zmo_kmer_revcomp(kmer, k);
		if(krev == kmer) continue;
		const uint32_t off = ring[(kept - k) & 31];
		const uint32_t ln = (j + 1 - off > 0xFFFFu)? 0xFFFFu : j + 1 - off;
		f(krev > kmer? kmer : krev, (uint32_t)(krev > kmer? 0 : 1), off, ln);
	}
}

/* ---- sort_array-exact sort (sort.h:104-155): median-of-3 quicksort, <=5-element partitions left to
 * a final bubble pass; the resulting permutation of equal keys is part of the contract. */
template<class T, class GT> ZMO_HDN void zmo_ref_sort(T *rs, size_t n, GT gt){
	size_t stack[64][2], x, s, e, i, j, m; T p, t;
	if(n < 2) return;
	stack[0][0] = 0; stack[0][1] = n - 1; x = 1;
	while(x){
		x--; s = stack[x][0]; e = stack[x][1];
		m = s + (e - s) / 2;
		if(gt(rs[s], rs[m])){ t = rs[s]; rs[s] = rs[m]; rs[m] = t; }
		if(gt(rs[m], rs[e])){
			t = rs[e]; rs[e] = rs[m]; rs[m] = t;
			if(gt(rs[s], rs[m])){ t = rs[s]; rs[s] = rs[m]; rs[m] = t; }
		}
		p = rs[m];
		i = s + 1; j = e - 1;
		while(1){
			while(gt(p, rs[i])) i++;
			while(gt(rs[j], p)) j--;
			if(i < j){ t = rs[i]; rs[i] = rs[j]; rs[j] = t; i++; j--; }
			else break;
		}
		if(i == j){ i++; j--; }
		if(j - s > e - i){
			if(s + 4 < j){ stack[x][0] = s; stack[x][1] = j; x++; }
			if(i + 4 < e){ stack[x][0] = i; stack[x][1] = e; x++; }
		} else {
			if(i + 4 < e){ stack[x][0] = i; stack[x][1] = e; x++; }
			if(s + 4 < j){ stack[x][0] = s; stack[x][1] = j; x++; }
		}
	}
	for(i = 0; i < n; i++){
		bool sw = false;
		for(j = n - 1; j > i; j--){
			if(gt(rs[j - 1], rs[j])){ t = rs[j - 1]; rs[j - 1] = rs[j]; rs[j] = t; sw = true; }
		}
		if(!sw) break;
	}
}

struct GtZPairOff12 { ZMO_HDM bool operator()(const DevZPair &a, const DevZPair &b) const { return (((int64_t)a.off1 << 32) | a.off2) > (((int64_t)b.off1 << 32) | b.off2); } };
/* reference emission order of the match list (c position major, then q occurrence): used to rebuild the input of
 * the exact sort for lists with tied keys.  c position = off2 on the same strand, clen-off2-len2 on the other. */
struct GtZPairEmit { uint32_t clen; ZMO_HDM uint64_t k(const DevZPair &a) const { const uint32_t cp = (a.dir1 ^ a.dir2)? clen - a.off2 - a.len2 : a.off2; return ((uint64_t)cp << 32) | a.off1; }
	ZMO_HDM bool operator()(const DevZPair &a, const DevZPair &b) const { return k(a) > k(b); } };
struct GtZPairOff1 { ZMO_HDM bool operator()(const DevZPair &a, const DevZPair &b) const { return a.off1 > b.off1; } };
struct GtIdxOff2 { const DevZPair *rs; ZMO_HDM bool operator()(uint32_t a, uint32_t b) const { return rs[a].off2 > rs[b].off2; } };

/* ---- z-mer matching of candidate c against the z-index of q (hzm_aln.h:173-224) ----------------
 * slots: distinct indexed z-mers of q (mer ascending) -> occurrences zs[off..off+cnt) in off order.
 * kcnts: one uint8 per slot (zeroed by the caller).  out == NULL counts only.  Returns #pairs. */
ZMO_HDN uint32_t zmo_zmatch(const uint32_t *cwords, uint32_t clen, const DevSlot *slots, uint32_t nslot, const DevZSeed *zs,
		uint8_t *kcnts, int zsize, int hz, uint32_t zcut, uint32_t kvar, DevZPair *out){
	uint32_t n = 0;
	zmo_scan_kmers(cwords, clen, zsize, hz, [&](uint64_t mer64, uint32_t dir, uint32_t off, uint32_t ln){
		const uint32_t mer = (uint32_t)mer64;
		uint32_t lo = 0, hi = nslot;
		while(lo < hi){ uint32_t mid = (lo + hi) >> 1; if(slots[mid].mer < mer) lo = mid + 1; else hi = mid; }
		if(lo >= nslot || slots[lo].mer != mer) return;
		if(kcnts[lo] >= zcut) return;
		kcnts[lo]++;
		const DevSlot s = slots[lo];
		for(uint32_t k = 0; k < s.cnt; k++){
			const DevZSeed p1 = zs[s.off + k];
			const uint32_t dl = p1.len > ln? p1.len - ln : ln - p1.len;
			if(dl > kvar) continue;
			if(out){
				DevZPair z; z.dir1 = p1.dir; z.off1 = p1.off; z.dir2 = (uint8_t)dir; z.len1 = p1.len; z.len2 = (uint16_t)ln; z.pad = 0;
				z.off2 = (p1.dir ^ dir)? clen - (off + ln) : off;
				out[n] = z;
			}
			n++;
		}
	});
	return n;
}

/* ---- median diagonal (hzm_aln.h:316-343) */
ZMO_HDN int32_t zmo_median_select(int32_t *rs, int32_t size){
	int32_t i, j, key, mid, beg = 0, end = size - 1, tmp;
	if(size == 0) return 0;
	while(beg < end){
		mid = beg + (end - beg) / 2;
		if(rs[beg] > rs[mid]){ tmp = rs[beg]; rs[beg] = rs[mid]; rs[mid] = tmp; }
		if(rs[mid] > rs[end]){
			tmp = rs[end]; rs[end] = rs[mid]; rs[mid] = tmp;
			if(rs[beg] > rs[mid]){ tmp = rs[beg]; rs[beg] = rs[mid]; rs[mid] = tmp; }
		}
		key = rs[mid]; i = beg + 1; j = end - 1;
		while(1){
			while(key > rs[i]) i++;
			while(rs[j] > key) j--;
			if(i < j){ tmp = rs[i]; rs[i] = rs[j]; rs[j] = tmp; i++; j--; } else break;
		}
		if(i == j){ i++; j--; }
		if(i <= size / 2) beg = i; else end = j;
	}
	return rs[size / 2];
}

struct SeedPar { uint32_t zsize, kwin, kstep, zovl, ztot; int W; };
struct WinScratch { uint64_t *ts, *ak; int32_t *as; uint32_t *wb, *we, *wo; uint32_t capt, capw; };     /* ts/ak/as hold capt entries, wb/we/wo capw */
/* stage: optional fast scratch (shared memory) where a window's anchors are gathered, sorted and measured before the
 * accepted ones are appended to anc; capwin_ovf: overflow code when wins is full (1 = grow arenas, 2 = retry in global) */
struct WinOut { DevWin *wins; uint32_t nwin, capwin; DevZPair *anc; uint32_t nanc, capanc; int overflow; DevZPair *stage; uint32_t capstage; int capwin_ovf; };

#define ZMO_KWIN_MAX_OFFSET_DEV 50
#define ZMO_WIN_OVL_MASK 0x1FFFFFFFu

/* hzm_aln.h:410-578 (fast-chaining branch) */
/* sort keys: the comparators of the reference look at one field only (off2 at hzm_aln.h:449, off1 at :519), so the
 * sorts run on packed 64-bit (key | payload) words instead of 16-byte structs reached through an index: same comparison
 * outcomes, hence the same sort_array permutation, with a fraction of the memory traffic. */
struct GtTsKey { ZMO_HDM bool operator()(uint64_t a, uint64_t b) const { return (a >> 40) > (b >> 40); } };     /* off2:24 | len2:16 | idx:24 */
struct GtHi32 { ZMO_HDM bool operator()(uint64_t a, uint64_t b) const { return (a >> 32) > (b >> 32); } };       /* key:32 | payload:32 */
#define ZMO_TS_OFF2(k) ((uint32_t)((k) >> 40))
#define ZMO_TS_LEN2(k) ((uint32_t)(((k) >> 24) & 0xFFFFu))
#define ZMO_TS_IDX(k)  ((uint32_t)((k) & 0xFFFFFFu))

/* hzm_aln.h:410-578 (fast-chaining branch) */
ZMO_HDN uint32_t zmo_windows_in_span(const DevZPair *rs, int dir, uint32_t beg, uint32_t end, int bound, WinOut &O, const WinScratch &S, const SeedPar &par){
	const uint32_t zsize = par.zsize, kwin = par.kwin, zovl = par.zovl;
	uint32_t i, j, n = 0, n2 = 0, ol, ol2, lst, ret = 0;
	while(beg < end){
		const DevZPair &p = rs[beg];
		if((p.dir1 ^ p.dir2 ^ dir) || (int)p.off1 < bound) beg++; else break;
	}
	for(i = beg; i < end; i++) if(!(rs[i].dir1 ^ rs[i].dir2 ^ dir)) n++;
	if(n * zsize < zovl) return 0;
	if(n > S.capt || end - beg >= (1u << 24)){ O.overflow = 2; return 0; }
	n = 0;
	for(i = beg; i < end; i++){ const DevZPair &p = rs[i]; if(!(p.dir1 ^ p.dir2 ^ dir)) S.ts[n++] = ((uint64_t)p.off2 << 40) | ((uint64_t)p.len2 << 24) | (i - beg); }
	zmo_ref_sort(S.ts, (size_t)n, GtTsKey());
	ol = 0; lst = 0;
	for(i = j = 0; i < n; i++){
		const uint64_t pk = S.ts[i]; const uint32_t p_off2 = ZMO_TS_OFF2(pk), p_len2 = ZMO_TS_LEN2(pk);
		while(p_off2 + p_len2 > ZMO_TS_OFF2(S.ts[j]) + kwin && j + 1 < n){
			const uint64_t k0 = S.ts[j++], k1 = S.ts[j];
			const uint32_t s = ZMO_TS_OFF2(k1), t = ZMO_TS_OFF2(k0) + ZMO_TS_LEN2(k0);
			ol2 = s < t? t - s : 0;
			ol = ol + ol2 - ZMO_TS_LEN2(k0);
		}
		ol += (p_off2 > lst)? p_len2 : p_off2 + p_len2 - lst;
		lst = p_off2 + p_len2;
		if(ol >= zovl){
			if(n2 && ( p_off2 <= ZMO_TS_OFF2(S.ts[S.we[n2-1]]) + kwin / 3 || ZMO_TS_OFF2(S.ts[j]) <= ZMO_TS_OFF2(S.ts[S.wb[n2-1]]) + kwin / 3 )){
				if(ol > S.wo[n2-1]){ S.wb[n2-1] = j; S.we[n2-1] = i; S.wo[n2-1] = ol; }
			} else { if(n2 >= S.capw){ O.overflow = 2; return 0; } S.wb[n2] = j; S.we[n2] = i; S.wo[n2] = ol; n2++; }
		}
	}
	for(i = 0; i < n2; i++){
		const uint32_t size = O.nanc; int32_t offset, offn = 0; DevWin W0;
		for(j = S.wb[i]; j <= S.we[i]; j++){ const DevZPair &p = rs[beg + ZMO_TS_IDX(S.ts[j])]; S.as[offn++] = (int32_t)p.off1 - (int32_t)p.off2; }
		offset = zmo_median_select(S.as, offn);
		/* anchors within +-50 of the median diagonal */
		uint32_t na = 0;
		for(j = S.wb[i]; j <= S.we[i]; j++){
			const DevZPair &p = rs[beg + ZMO_TS_IDX(S.ts[j])]; const int32_t off = (int32_t)p.off1 - (int32_t)p.off2;
			if(off < offset - ZMO_KWIN_MAX_OFFSET_DEV || off > offset + ZMO_KWIN_MAX_OFFSET_DEV) continue;
			na++;
		}
		if(na == 0) continue;
		if(size + na > O.capanc){ O.overflow = 1; return ret; }
		/* staged path: gather into fast memory, sort (off1 | slot) keys, measure through the keys, copy accepted windows out in
		 * sorted order; otherwise gather straight into the arena and sort the structs there */
		const bool staged = O.stage && na <= O.capstage;
		DevZPair *A = staged? O.stage : O.anc + size;
		na = 0;
		for(j = S.wb[i]; j <= S.we[i]; j++){
			const DevZPair &p = rs[beg + ZMO_TS_IDX(S.ts[j])]; const int32_t off = (int32_t)p.off1 - (int32_t)p.off2;
			if(off < offset - ZMO_KWIN_MAX_OFFSET_DEV || off > offset + ZMO_KWIN_MAX_OFFSET_DEV) continue;
			if(staged) S.ak[na] = ((uint64_t)p.off1 << 32) | na;
			A[na++] = p;
		}
		if(staged) zmo_ref_sort(S.ak, (size_t)na, GtHi32()); else zmo_ref_sort(A, (size_t)na, GtZPairOff1());
		W0.closed = 0; W0.dir = (uint8_t)dir; W0.pad = 0; W0.pb2 = 0; W0.anc0 = size; W0.beg[0] = W0.beg[1] = 0x7FFFFFFF; W0.end[0] = W0.end[1] = 0;
		ol = lst = 0;
		for(j = 0; j < na; j++){
			const DevZPair &p = staged? A[(uint32_t)S.ak[j]] : A[j];
			ol += (p.off1 > lst)? (uint32_t)p.len1 : p.off1 + p.len1 - lst;
			lst = p.off1 + p.len1;
			if((int)p.off1 < W0.beg[0]) W0.beg[0] = p.off1;
			if((int)(p.off1 + p.len1) > W0.end[0]) W0.end[0] = p.off1 + p.len1;
			if((int)p.off2 < W0.beg[1]) W0.beg[1] = p.off2;
			if((int)(p.off2 + p.len2) > W0.end[1]) W0.end[1] = p.off2 + p.len2;
		}
		if(ol * 2 < zovl) continue;
		if(ret){
			const DevWin &w = O.wins[O.nwin - 1];
			if(W0.end[1] <= (int)(w.end[1] + kwin / 3) && ol <= w.ovl) continue;
		}
		if(O.nwin >= O.capwin){ O.overflow = O.capwin_ovf; return ret; }
		if(staged) for(j = 0; j < na; j++) O.anc[size + j] = A[(uint32_t)S.ak[j]];
		O.nanc = size + na;
		ret++;
		W0.ovl = ol & ZMO_WIN_OVL_MASK; W0.anc1 = O.nanc;
		O.wins[O.nwin++] = W0;
	}
	return ret;
}

/* hzm_aln.h:580-656 */
ZMO_HDN uint32_t zmo_pair_windows_strand(const DevZPair *rs, uint32_t n, int dir, WinOut &O, const WinScratch &S, const SeedPar &par){
	const uint32_t kwin = par.kwin, kstep = par.kstep, zovl = par.zovl;
	uint32_t i, j, a, nw, ol = 0, ol2, lst = 0, wlst = 0, s, t, ret = 0;
	uint32_t p0_off1, p0_len1, p_off1, p_len1;
	for(j = 0; j < n; j++) if(!(rs[j].dir1 ^ rs[j].dir2 ^ dir)) break;
	if(j == n) return 0;
	p0_off1 = rs[j].off1; p0_len1 = rs[j].len1;
	for(i = j; i <= n; i++){
		if(i < n){
			if(rs[i].dir1 ^ rs[i].dir2 ^ dir) continue;
			p_off1 = rs[i].off1; p_len1 = rs[i].len1;
		} else { p_off1 = 0x1FFFFFu; p_len1 = 0x3FFu; }
		if(p_off1 > p0_off1 + kwin){
			if(ol >= zovl){
				if((nw = zmo_windows_in_span(rs, dir, j, i, (int)wlst, O, S, par))){
					for(a = 0; a < nw; a++){ const int e0 = O.wins[O.nwin + a - nw].end[0] + 20; if((int)wlst < e0) wlst = e0; }
					ret += nw;
					p0_off1 = p_off1; p0_len1 = p_len1;
					ol = p_len1; lst = p_off1 + p_len1; j = i;
				} else if(i < n){
					const uint32_t nxt = p0_off1 + kstep;
					while(p0_off1 < nxt && j < i){
						const DevZPair &p1 = rs[++j];
						s = p0_off1 > p1.off1? p0_off1 : p1.off1;
						t = (p0_off1 + p0_len1) < ((uint32_t)p1.off1 + p1.len1)? (p0_off1 + p0_len1) : ((uint32_t)p1.off1 + p1.len1);
						ol2 = s < t? t - s : 0;
						ol = ol + ol2 - p0_len1;
						p0_off1 = p1.off1; p0_len1 = p1.len1;
					}
				}
				if(O.overflow) return ret;
			}
			if(p_off1 == 0x1FFFFFu) break;
			while(p_off1 > p0_off1 + kwin){
				const DevZPair &p1 = rs[++j];
				s = p0_off1 > p1.off1? p0_off1 : p1.off1;
				t = (p0_off1 + p0_len1) < ((uint32_t)p1.off1 + p1.len1)? (p0_off1 + p0_len1) : ((uint32_t)p1.off1 + p1.len1);
				ol2 = s < t? t - s : 0;
				ol = ol + ol2 - p0_len1;
				p0_off1 = p1.off1; p0_len1 = p1.len1;
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

/* hzm_aln.h:658-713; nodes: 2 ints per window of scratch */
ZMO_HDN int zmo_chain_windows(DevWin *w, uint32_t n, int W, int *nodes){
	uint32_t i, j; int mw = -1000000, bt = -1, band;
	const float band_penalty = 0.05f;
	for(i = 0; i < n; i++){ nodes[2 * i] = 0; nodes[2 * i + 1] = -1; }
	for(i = 0; i < n; i++){
		w[i].closed = 1;
		nodes[2 * i] += (int)w[i].ovl;
		if(nodes[2 * i] > mw){ mw = nodes[2 * i]; bt = (int)i; }
		for(j = i + 1; j < n; j++){
			if(w[j].beg[1] < w[i].end[1]) continue;
			if(w[j].beg[0] < w[i].end[0]) continue;
			if(w[j].beg[0] - w[i].end[0] > W && w[j].beg[1] - w[i].end[1] > W) break;
			const int d0 = w[j].beg[0] - w[i].end[0], d1 = w[j].beg[1] - w[i].end[1];
			band = d0 < d1? d1 - d0 : d0 - d1;
			if(band > W) continue;
			band = (int)((float)band * band_penalty);
			if(nodes[2 * j] < nodes[2 * i] - band){ nodes[2 * j] = nodes[2 * i] - band; nodes[2 * j + 1] = (int)i; }
		}
	}
	mw = 0;
	while(bt >= 0){ w[bt].closed = 0; mw += w[bt].end[0] - w[bt].beg[0]; bt = nodes[2 * bt + 1]; }
	return mw;
}

/* ---- per-pair, per-strand seeding (wtzmo.c:888-912 for one strand): windows + chain ------------
 * Scratch layout for a pair with n z-mer matches (bytes): see zmo_pair_scratch_bytes().  After the
 * call the strand's windows (with closed flags set by the chain) are in S.w2[0..nwin) and their
 * anchors in S.a2; the caller copies the kept ones out. */
struct PairScratch { WinScratch ws; DevWin *w2; DevZPair *a2; uint32_t cap, capw2; DevZPair *stage; uint32_t capstage; int w2_ovf; };
/* F = capacity factor for windows/anchors (an anchor can belong to several overlapping sub-windows,
 * hzm_aln.h:483-514, so the anchor list of a strand may exceed the match count) */
ZMO_HD size_t zmo_pair_scratch_per(uint32_t F){ return 8 + 8 + 4 * 4 + (size_t)F * (sizeof(DevWin) + sizeof(DevZPair)); }
ZMO_HD size_t zmo_pair_scratch_bytes(uint32_t n, uint32_t F){ return (size_t)n * zmo_pair_scratch_per(F) + 64; }
ZMO_HD PairScratch zmo_pair_scratch_carve(uint8_t *base, uint32_t n, uint32_t F){
	PairScratch P; uint8_t *p = base;
	P.a2 = (DevZPair*)p; p += (size_t)n * F * sizeof(DevZPair);
	P.w2 = (DevWin*)p; p += (size_t)n * F * sizeof(DevWin);
	P.ws.ts = (uint64_t*)p; p += (size_t)n * 8;
	P.ws.ak = (uint64_t*)p; p += (size_t)n * 8;
	P.ws.as = (int32_t*)p; p += (size_t)n * 4;
	P.ws.wb = (uint32_t*)p; p += (size_t)n * 4;
	P.ws.we = (uint32_t*)p; p += (size_t)n * 4;
	P.ws.wo = (uint32_t*)p;
	P.ws.capt = n; P.ws.capw = n;
	P.cap = n * F; P.capw2 = n * F; P.stage = nullptr; P.capstage = 0; P.w2_ovf = 1;
	return P;
}
/* returns the chain weight (0 when the strand has no window); nwin/nanc = all windows found */
ZMO_HDN int zmo_pair_seed_strand(const DevZPair *cache, uint32_t n, int dir, const SeedPar &par, PairScratch &P, uint32_t *nwin, int *overflow){
	WinOut O; O.wins = P.w2; O.nwin = 0; O.capwin = P.capw2; O.anc = P.a2; O.nanc = 0; O.capanc = P.cap; O.overflow = 0; O.stage = P.stage; O.capstage = P.capstage; O.capwin_ovf = P.w2_ovf;
	int ovl = 0;
	if(zmo_pair_windows_strand(cache, n, dir, O, P.ws, par) && !O.overflow){
		/* ts/as (8 bytes per match) are free again: reuse as chain nodes (8 bytes per window) */
		ovl = zmo_chain_windows(P.w2, O.nwin, par.W, (int*)P.ws.ts);
	}
	*nwin = O.nwin; *overflow = O.overflow;
	return ovl;
}
