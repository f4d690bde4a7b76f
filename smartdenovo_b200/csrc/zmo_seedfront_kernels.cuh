/*
 * zmo_seedfront_kernels.cuh -- front end of the pair-seeding stage, fully parallel: z-index of the batch's query reads
 * (index_single_read_seeds, hzm_aln.h:70-115), z-mer hits of every (pair, 128-base chunk of c) with the per-query slot filter,
 * rank cap + expansion into matches (query_single_read_seeds, hzm_aln.h:173-224), unpacking of the sorted match lists.  Kept in
 * a header so that the test-only host simulation (tests/hostsim/seedk_host.cpp) runs this very source; included by zmo_seed.cu only
 * (which supplies the CUB sorts and scans between the kernels: seed_prepare).
 */
#pragma once
#include "zmo_ctx.cuh"
#include "zmo_seed_core.cuh"

#define ZSCAN_CH 128
/* z-mers of the batch's query reads, one thread per (read, 128-base chunk): chunk c of read u emits the z-mers whose last kept base lies in the
 * chunk (zmo_scan_kmers_chunk), in offset order; the chunk table choff (nuq + 1 entries, built on the host from the read lengths) maps chunks to
 * reads.  PASS 0 counts per chunk, PASS 1 writes key = u << 32 | mer, value = off << 17 | len << 1 | dir at the chunk's offset. */
template<int PASS>
__global__ void k_z_scan(DevReads R, const uint32_t *uq, uint32_t nuq, const unsigned long long *choff, unsigned long long NC, int zsize, int hz,
		unsigned long long *cnt_or_off, unsigned long long *keys, unsigned long long *vals){
	unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(c >= NC) return;
	uint32_t lo = 0, hi = nuq;
	while(lo + 1 < hi){ uint32_t mid = (lo + hi) >> 1; if(choff[mid] <= c) lo = mid; else hi = mid; }
	const uint32_t u = lo, rid = uq[u], s0 = (uint32_t)(c - choff[u]) * ZSCAN_CH;
	unsigned long long n = PASS? cnt_or_off[c] : 0;
	zmo_scan_kmers_chunk(R.words + R.woff[rid], R.len[rid], zsize, hz, s0, s0 + ZSCAN_CH, [&](uint64_t mer, uint32_t dir, uint32_t off, uint32_t ln){
		if(PASS){ keys[n] = ((unsigned long long)u << 32) | (uint32_t)mer; vals[n] = ((unsigned long long)off << 17) | ((unsigned long long)ln << 1) | dir; }
		n++;
	});
	if(!PASS) cnt_or_off[c] = n;
}
/* first z-mer of every read = offset of its first chunk (zoff[nuq] = total) */
__global__ void k_z_readoff(const unsigned long long *choff, const unsigned long long *coff, uint32_t nuq, unsigned long long *zoff){
	uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
	if(u <= nuq) zoff[u] = coff[choff[u]];
}
/* per sorted z-seed: unpack payload; run heads count their run and flag it as a slot if shorter than zcut */
__global__ void k_z_heads(const unsigned long long *keys, const unsigned long long *vals, unsigned long long n, uint32_t zcut, uint32_t *flag, uint32_t *runlen, DevZSeed *zs){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	const unsigned long long v = vals[i]; DevZSeed s; s.off = (uint32_t)(v >> 17); s.len = (uint16_t)((v >> 1) & 0xFFFFu); s.dir = (uint8_t)(v & 1u); s.pad = 0;
	zs[i] = s;
	const unsigned long long key = keys[i];
	if(i && keys[i - 1] == key){ flag[i] = 0; return; }
	uint32_t c = 1;
	for(unsigned long long j = i + 1; j < n && keys[j] == key && c <= zcut; j++) c++;
	flag[i] = c < zcut; runlen[i] = c;
}
/* Per-query membership filter over the slot z-mers: ZF_BITS bits per query read, one hashed bit per slot.  A c z-mer whose
 * bit is clear cannot hit a slot (>= 97% of all look-ups at ~8,000 slots per read), so k_hit pays one L2-resident load
 * instead of a 13-step binary search for it; the reference's 4^z bit vector (hzm_aln.h:107-114,152) plays the same role. */
#define ZF_LOG 18
#define ZF_WORDS (1u << (ZF_LOG - 5))
__device__ __forceinline__ uint32_t zf_hash(uint32_t mer){ return (mer * 2654435761u) >> (32 - ZF_LOG); }
__global__ void k_z_slots(const unsigned long long *keys, const uint32_t *flag, const uint32_t *pos, const uint32_t *runlen, unsigned long long n, const unsigned long long *zoff, DevSlot *slots, uint32_t *filt){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n || !flag[i]) return;
	const uint32_t u = (uint32_t)(keys[i] >> 32);
	DevSlot s; s.mer = (uint32_t)keys[i]; s.off = (uint32_t)(i - zoff[u]); s.cnt = runlen[i];
	slots[pos[i]] = s;
	{ const uint32_t h = zf_hash(s.mer); atomicOr(filt + (size_t)u * ZF_WORDS + (h >> 5), 1u << (h & 31)); }
}
__global__ void k_z_ranges(const unsigned long long *zoff, const uint32_t *pos, uint32_t nuq, unsigned long long Z, uint32_t NS, uint32_t *slot_beg){
	uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
	if(u > nuq) return;
	slot_beg[u] = (u < nuq && zoff[u] < Z)? pos[zoff[u]] : NS;
}
struct ZIdxView { const DevSlot *slots; const uint32_t *slot_beg; const DevZSeed *zs; const unsigned long long *zoff; const uint32_t *filt; };
#define SEED_CH 128
/* chunk table of the candidate reads: one thread per pair */
__global__ void k_c_nchunks(DevReads R, const uint32_t *pc, uint32_t np, unsigned long long *nch){
	uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if(p < np) nch[p] = (R.len[pc[p]] + SEED_CH - 1) / SEED_CH;
}
/* one thread per (pair, 128-base chunk of c): z-mers of the chunk that hit an indexed slot of q (hzm_aln.h:189-207).
 * key = pair<<32 | slot (slot index inside q's list), val = off<<17 | len<<1 | dir of the c z-mer.  Threads write in
 * (pair, chunk, position) order, so a stable sort by key keeps the hits of one slot in c-position order. */
template<int PASS>
__global__ void k_hit(DevReads R, ZIdxView Z, const uint32_t *pq, const uint32_t *pc, uint32_t np, const unsigned long long *choff, unsigned long long NC,
		int zsize, int hz, unsigned long long *cnt_or_off, unsigned long long *hkey, unsigned long long *hval){
	unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(c >= NC) return;
	uint32_t lo = 0, hi = np;
	while(lo + 1 < hi){ uint32_t mid = (lo + hi) >> 1; if(choff[mid] <= c) lo = mid; else hi = mid; }
	const uint32_t p = lo, u = pq[p], cid = pc[p], s = (uint32_t)(c - choff[p]) * SEED_CH;
	const uint32_t sb = Z.slot_beg[u], ns = Z.slot_beg[u + 1] - sb; const DevSlot *slots = Z.slots + sb;
	const uint32_t *fl = Z.filt + (size_t)u * ZF_WORDS;
	unsigned long long n = PASS? cnt_or_off[c] : 0;
	zmo_scan_kmers_chunk(R.words + R.woff[cid], R.len[cid], zsize, hz, s, s + SEED_CH, [&](uint64_t mer64, uint32_t dir, uint32_t off, uint32_t ln){
		const uint32_t mer = (uint32_t)mer64;
		{ const uint32_t h = zf_hash(mer); if(!((__ldg(fl + (h >> 5)) >> (h & 31)) & 1u)) return; }
		uint32_t a = 0, b = ns;
		while(a < b){ uint32_t mid = (a + b) >> 1; if(slots[mid].mer < mer) a = mid + 1; else b = mid; }
		if(a >= ns || slots[a].mer != mer) return;
		if(PASS){ hkey[n] = ((unsigned long long)p << 32) | a; hval[n] = ((unsigned long long)off << 17) | ((unsigned long long)ln << 1) | dir; }
		n++;
	});
	if(!PASS) cnt_or_off[c] = n;
}
/* one thread per sorted hit: rank among the hits of the same (pair, slot) = number of earlier c positions that hit
 * the slot; the reference's uint8 per-slot counter admits the first Z positions (hzm_aln.h:208-211; with Z > 255 the
 * counter wraps and never blocks).  Surviving hits expand into one match per q occurrence whose span length differs by
 * <= kvar (hzm_aln.h:212-220).  MODE 0: key = pair<<48 | off1<<24 | off2 (SW path, process_hzmps order);
 * MODE 1: key = pair<<49 | (off1-off2+2^24)<<24 | off1 (dot-matrix path, denoising_hzmps order).  val = len1<<18|len2<<2|dir1<<1|dir2. */
template<int PASS, int MODE>
__global__ void k_expand(DevReads R, ZIdxView Z, const uint32_t *pq, const uint32_t *pc, const unsigned long long *hkey, const unsigned long long *hval, unsigned long long NH,
		uint32_t zcut, uint32_t kvar, unsigned long long *cnt_or_off, unsigned long long *zkey, unsigned long long *zval){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= NH) return;
	const unsigned long long key = hkey[i];
	unsigned long long n = PASS? cnt_or_off[i] : 0;
	uint32_t rank = 0;
	if(zcut <= 255u){ for(unsigned long long j = i; j > 0 && hkey[j - 1] == key && rank < zcut; j--) rank++; }
	if(zcut > 255u || rank < zcut){
		const uint32_t p = (uint32_t)(key >> 32), u = pq[p]; const DevSlot s = Z.slots[Z.slot_beg[u] + (uint32_t)key];
		const DevZSeed *zs = Z.zs + Z.zoff[u] + s.off;
		const unsigned long long v = hval[i]; const uint32_t coff = (uint32_t)(v >> 17), ln = (uint32_t)((v >> 1) & 0xFFFFu), dir = (uint32_t)(v & 1u);
		const uint32_t clen = R.len[pc[p]];
		for(uint32_t k = 0; k < s.cnt; k++){
			const DevZSeed p1 = zs[k];
			const uint32_t dl = p1.len > ln? p1.len - ln : ln - p1.len;
			if(dl > kvar) continue;
			if(PASS){
				const uint32_t off2 = (p1.dir ^ dir)? clen - (coff + ln) : coff;
				if(MODE == 0) zkey[n] = ((unsigned long long)p << 48) | ((unsigned long long)p1.off << 24) | off2;
				else zkey[n] = ((unsigned long long)p << 49) | ((unsigned long long)(p1.off + 0x1000000u - off2) << 24) | p1.off;
				zval[n] = ((unsigned long long)p1.len << 18) | ((unsigned long long)ln << 2) | ((unsigned long long)p1.dir << 1) | dir;
			}
			n++;
		}
	}
	if(!PASS) cnt_or_off[i] = n;
}
/* sorted (key,val) -> DevZPair list; equal adjacent keys (same q occurrence and same c coordinate on the two strands)
 * are the only ties of the reference's unstable sorts: flag the pair so that k_p_seed / k_p_dot re-creates the reference
 * emission order and runs the exact sort_array emulation for it */
template<int MODE>
__global__ void k_unpack(const unsigned long long *zkey, const unsigned long long *zval, unsigned long long T, DevZPair *cache, uint8_t *tie){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= T) return;
	const unsigned long long key = zkey[i], v = zval[i]; DevZPair z;
	if(MODE == 0){ z.off1 = (uint32_t)((key >> 24) & 0xFFFFFFu); z.off2 = (uint32_t)(key & 0xFFFFFFu); }
	else { z.off1 = (uint32_t)(key & 0xFFFFFFu); z.off2 = z.off1 + 0x1000000u - (uint32_t)((key >> 24) & 0x1FFFFFFu); }
	z.len1 = (uint16_t)(v >> 18); z.len2 = (uint16_t)((v >> 2) & 0xFFFFu); z.dir1 = (uint8_t)((v >> 1) & 1u); z.dir2 = (uint8_t)(v & 1u); z.pad = 0;
	cache[i] = z;
	if(i && zkey[i - 1] == key) tie[(uint32_t)(key >> (MODE == 0? 48 : 49))] = 1;
}
template<int MODE>
__global__ void k_pair_offsets(const unsigned long long *zkey, unsigned long long T, uint32_t np, unsigned long long *cache_off){
	uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if(p > np) return;
	unsigned long long lo = 0, hi = T;
	while(lo < hi){ unsigned long long mid = (lo + hi) >> 1; if((uint32_t)(zkey[mid] >> (MODE == 0? 48 : 49)) < p) lo = mid + 1; else hi = mid; }
	cache_off[p] = lo;
}
