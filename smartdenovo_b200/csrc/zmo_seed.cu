/*
 * zmo_seed.cu -- pair seeding on the device: z-index of every query read of the batch, z-mer
 * matching of every (query, candidate) pair, and (one thread per pair, order-exact serial logic from
 * zmo_seed_core.cuh) the (off1,off2) sort, window finding and window chaining of both strands.
 * Replaces wtzmo.c:845-914 minus the windeps side effect, which stays in the host replay.
 *
 * z-index layout: all z-mers of the batch's distinct query reads are emitted as
 * key = qlocal<<32 | mer, payload = (off,len,dir), stable-radix-sorted by key (emission is in offset
 * order, so each mer run stays offset-sorted like hzm_aln.h:101).  Runs shorter than -Z become slots
 * (the reference's bit-vector + rank, hzm_aln.h:107-114,152, is replaced by a binary search over the
 * read's sorted slot list).
 */
#include <cub/cub.cuh>
#include <unordered_map>
#include "zmo_ctx.cuh"
#include "zmo_seed_core.cuh"
#include "zmo_seed.cuh"
#include "zmo_seed_warp.cuh"
#include "zmo_seed_kernels.cuh"

#define CUB_CALL(c, call_expr) do { size_t _tb = 0; void *_tp = nullptr; { auto d_temp = _tp; size_t &temp_bytes = _tb; CUDA_TRY(call_expr); } \
	if((c)->cubtmp.reserve(_tb + 256)) return ZMO_ERR_CUDA; { void *d_temp = (c)->cubtmp.p; size_t &temp_bytes = _tb; CUDA_TRY(call_expr); } (c)->launches++; } while(0)

template<int PASS>
__global__ void k_z_scan(DevReads R, const uint32_t *uq, uint32_t nuq, int zsize, int hz, unsigned long long *cnt_or_off, unsigned long long *keys, unsigned long long *vals){
	uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
	if(u >= nuq) return;
	const uint32_t rid = uq[u];
	unsigned long long n = PASS? cnt_or_off[u] : 0;
	zmo_scan_kmers(R.words + R.woff[rid], R.len[rid], zsize, hz, [&](uint64_t mer, uint32_t dir, uint32_t off, uint32_t ln){
		if(PASS){ keys[n] = ((unsigned long long)u << 32) | (uint32_t)mer; vals[n] = ((unsigned long long)off << 17) | ((unsigned long long)ln << 1) | dir; }
		n++;
	});
	if(!PASS) cnt_or_off[u] = n;
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

/* steps shared by the SW and dot-matrix paths: z-index of the batch's query reads + match lists */
int seed_prepare(zmo_ctx *c, const zmo_pair_t *pairs, uint32_t np, int mode, SeedWork &W, DevBuf &cache_buf){
	DevReads R = dev_reads(c);
	std::vector<uint32_t> uq, pq(np), pc(np); std::unordered_map<uint32_t, uint32_t> qmap;
	for(uint32_t i = 0; i < np; i++){
		if(pairs[i].qid >= c->st->n_reads || pairs[i].cid >= c->st->n_reads) return zmo_set_err(ZMO_ERR_ARG, "pair %u: read id out of range", i);
		auto it = qmap.find(pairs[i].qid);
		if(it == qmap.end()){ it = qmap.emplace(pairs[i].qid, (uint32_t)uq.size()).first; uq.push_back(pairs[i].qid); }
		pq[i] = it->second; pc[i] = pairs[i].cid;
	}
	const uint32_t nuq = (uint32_t)uq.size();
	W.np = np; W.nuq = nuq;
	/* layout of small per-batch arrays in s0: uq | pq | pc | slot_beg ; s1: zcnt/zoff ; */
	if(c->s0.reserve(((size_t)nuq + 2 + 2 * (size_t)np + nuq + 4) * 4) || c->s1.reserve(((size_t)nuq + 2) * 16)) return ZMO_ERR_CUDA;
	uint32_t *d_uq = c->s0.as<uint32_t>(), *d_pq = d_uq + nuq + 1, *d_pc = d_pq + np, *d_slot_beg = d_pc + np;
	unsigned long long *d_zcnt = c->s1.as<unsigned long long>(), *d_zoff = d_zcnt + nuq + 1;
	CUDA_TRY(cudaMemcpyAsync(d_uq, uq.data(), (size_t)nuq * 4, cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(d_pq, pq.data(), (size_t)np * 4, cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(d_pc, pc.data(), (size_t)np * 4, cudaMemcpyHostToDevice, c->stream));
	c->counters[5] += ((size_t)nuq + 2 * (size_t)np) * 4;
	const int bs = 32;
	k_z_scan<0><<<(nuq + bs - 1) / bs, bs, 0, c->stream>>>(R, d_uq, nuq, c->par.zsize, c->par.hz, d_zcnt, nullptr, nullptr); c->launches++;
	CUDA_TRY(cudaMemsetAsync(d_zcnt + nuq, 0, 8, c->stream));
	CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_zcnt, d_zoff, nuq + 1, c->stream));
	unsigned long long Z = 0;
	CUDA_TRY(cudaMemcpyAsync(&Z, d_zoff + nuq, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if(Z >= 0xFFFFFFF0ull) return zmo_set_err(ZMO_ERR_CAPACITY, "pair batch too large (%llu z-mers)", Z);
	/* s2 keys_in, s3 keys_out, s4 vals_in, s5 vals_out, s6 flag|pos|runlen, s7 zs|slots */
	const unsigned long long Zp = Z + 4;
	if(c->s2.reserve(Zp * 8) || c->s3.reserve(Zp * 8) || c->s4.reserve(Zp * 8) || c->s5.reserve(Zp * 8) || c->s6.reserve(Zp * 12) || c->s7.reserve(Zp * (sizeof(DevZSeed) + sizeof(DevSlot)))) return ZMO_ERR_CUDA;
	unsigned long long *k_in = c->s2.as<unsigned long long>(), *k_out = c->s3.as<unsigned long long>(), *v_in = c->s4.as<unsigned long long>(), *v_out = c->s5.as<unsigned long long>();
	uint32_t *d_flag = c->s6.as<uint32_t>(), *d_pos = d_flag + Zp, *d_run = d_pos + Zp;
	DevZSeed *d_zs = c->s7.as<DevZSeed>(); DevSlot *d_slots = (DevSlot*)(d_zs + Zp);
	uint32_t NS = 0;
	if(c->zfilt.reserve((size_t)(nuq + 1) * ZF_WORDS * 4)) return ZMO_ERR_CUDA;
	CUDA_TRY(cudaMemsetAsync(c->zfilt.p, 0, (size_t)(nuq + 1) * ZF_WORDS * 4, c->stream));
	if(Z){
		k_z_scan<1><<<(nuq + bs - 1) / bs, bs, 0, c->stream>>>(R, d_uq, nuq, c->par.zsize, c->par.hz, d_zoff, k_in, v_in); c->launches++;
		int qbits = 1; while((1ull << qbits) < nuq) qbits++;
		CUB_CALL(c, cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, k_in, k_out, v_in, v_out, (uint64_t)Z, 0, 32 + qbits, c->stream));
		k_z_heads<<<(unsigned)((Z + 255) / 256), 256, 0, c->stream>>>(k_out, v_out, Z, (uint32_t)c->par.zcut, d_flag, d_run, d_zs); c->launches++;
		CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_flag, d_pos, (uint64_t)Z, c->stream));
		uint32_t lp = 0, lf = 0;
		CUDA_TRY(cudaMemcpyAsync(&lp, d_pos + (Z - 1), 4, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaMemcpyAsync(&lf, d_flag + (Z - 1), 4, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		NS = lp + lf;
		k_z_slots<<<(unsigned)((Z + 255) / 256), 256, 0, c->stream>>>(k_out, d_flag, d_pos, d_run, Z, d_zoff, d_slots, c->zfilt.as<uint32_t>()); c->launches++;
	}
	k_z_ranges<<<(nuq + 1 + 127) / 128, 128, 0, c->stream>>>(d_zoff, d_pos, nuq, Z, NS, d_slot_beg); c->launches++;
	/* ---- match lists, fully parallel: hits per (pair, c-chunk) -> stable sort by (pair, slot) -> rank cap + expansion
	 * -> sort by the key of the consumer's first sort -> DevZPair lists.  s2: pair chunk table | tie flags,
	 * s3/s4: keys, s6/arena: values + per-item counters (s7 keeps the z-index). */
	if(np > (mode? 32768u : 65536u)) return zmo_set_err(ZMO_ERR_ARG, "at most %u pairs per call", mode? 32768u : 65536u);
	if(c->st->max_rdlen >= (1u << 24)) return zmo_set_err(ZMO_ERR_ARG, "reads of 2^24 bases or more are not supported (reference limit rdlen:24, wtzmo.c:88)");
	if(c->s2.reserve(((size_t)np + 2) * 24 + np + 64)) return ZMO_ERR_CUDA;
	unsigned long long *d_pnch = c->s2.as<unsigned long long>(), *d_pchoff = d_pnch + np + 1, *d_coff = d_pchoff + np + 1; uint8_t *d_tie = (uint8_t*)(d_coff + np + 2);
	ZIdxView ZV; ZV.slots = d_slots; ZV.slot_beg = d_slot_beg; ZV.zs = d_zs; ZV.zoff = d_zoff; ZV.filt = c->zfilt.as<uint32_t>();
	k_c_nchunks<<<(np + 127) / 128, 128, 0, c->stream>>>(R, d_pc, np, d_pnch); c->launches++;
	CUDA_TRY(cudaMemsetAsync(d_pnch + np, 0, 8, c->stream));
	CUDA_TRY(cudaMemsetAsync(d_tie, 0, np, c->stream));
	CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_pnch, d_pchoff, np + 1, c->stream));
	unsigned long long NC = 0;
	CUDA_TRY(cudaMemcpyAsync(&NC, d_pchoff + np, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if(c->s3.reserve((NC + 2) * 16)) return ZMO_ERR_CUDA;
	unsigned long long *d_ccnt = c->s3.as<unsigned long long>(), *d_choff = d_ccnt + NC + 1;
	k_hit<0><<<(unsigned)((NC + 127) / 128), 128, 0, c->stream>>>(R, ZV, d_pq, d_pc, np, d_pchoff, NC, c->par.zsize, c->par.hz, d_ccnt, nullptr, nullptr); c->launches++;
	CUDA_TRY(cudaMemsetAsync(d_ccnt + NC, 0, 8, c->stream));
	CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_ccnt, d_choff, (uint64_t)NC + 1, c->stream));
	unsigned long long NH = 0;
	CUDA_TRY(cudaMemcpyAsync(&NH, d_choff + NC, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	unsigned long long T = 0;
	if(NH >= 0xFFFFFFF0ull) return zmo_set_err(ZMO_ERR_CAPACITY, "pair batch too large (%llu z-mer hits)", NH);
	if(NH){
		/* hit arrays in the DP arena (idle during seeding): hkey_in | hkey_out | hval_in | hval_out | counts | offsets */
		if(c->arena.reserve((NH + 2) * 48 + 256)) return ZMO_ERR_CUDA;
		unsigned long long *hk_in = c->arena.as<unsigned long long>(), *hk_out = hk_in + NH + 1, *hv_in = hk_out + NH + 1, *hv_out = hv_in + NH + 1, *d_hcnt = hv_out + NH + 1, *d_hoff = d_hcnt + NH + 1;
		k_hit<1><<<(unsigned)((NC + 127) / 128), 128, 0, c->stream>>>(R, ZV, d_pq, d_pc, np, d_pchoff, NC, c->par.zsize, c->par.hz, d_choff, hk_in, hv_in); c->launches++;
		int pbits = 1; while((1ull << pbits) < np) pbits++;
		CUB_CALL(c, cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, hk_in, hk_out, hv_in, hv_out, (uint64_t)NH, 0, 32 + pbits, c->stream));
		/* hk_in / hv_in are free now: expansion counters live there */
		d_hcnt = hk_in; d_hoff = hv_in;
		if(mode == 0) k_expand<0, 0><<<(unsigned)((NH + 127) / 128), 128, 0, c->stream>>>(R, ZV, d_pq, d_pc, hk_out, hv_out, NH, (uint32_t)c->par.zcut, (uint32_t)c->par.kvar, d_hcnt, nullptr, nullptr);
		else k_expand<0, 1><<<(unsigned)((NH + 127) / 128), 128, 0, c->stream>>>(R, ZV, d_pq, d_pc, hk_out, hv_out, NH, (uint32_t)c->par.zcut, (uint32_t)c->par.kvar, d_hcnt, nullptr, nullptr);
		c->launches++;
		CUDA_TRY(cudaMemsetAsync(d_hcnt + NH, 0, 8, c->stream));
		CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_hcnt, d_hoff, (uint64_t)NH + 1, c->stream));
		CUDA_TRY(cudaMemcpyAsync(&T, d_hoff + NH, 8, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		if(T >= 0xFFFFFFF0ull) return zmo_set_err(ZMO_ERR_CAPACITY, "pair batch too large (%llu z-mer matches)", T);
		if(T){
			/* match keys/values: s3 (chunk tables dead) zk_in | zk_out ; s4 zv_in | zv_out */
			if(c->s3.reserve((T + 2) * 16) || c->s4.reserve((T + 2) * 16) || cache_buf.reserve((T + 4) * sizeof(DevZPair))) return ZMO_ERR_CUDA;
			unsigned long long *zk_in = c->s3.as<unsigned long long>(), *zk_out = zk_in + T + 1, *zv_in = c->s4.as<unsigned long long>(), *zv_out = zv_in + T + 1;
			if(mode == 0) k_expand<1, 0><<<(unsigned)((NH + 127) / 128), 128, 0, c->stream>>>(R, ZV, d_pq, d_pc, hk_out, hv_out, NH, (uint32_t)c->par.zcut, (uint32_t)c->par.kvar, d_hoff, zk_in, zv_in);
			else k_expand<1, 1><<<(unsigned)((NH + 127) / 128), 128, 0, c->stream>>>(R, ZV, d_pq, d_pc, hk_out, hv_out, NH, (uint32_t)c->par.zcut, (uint32_t)c->par.kvar, d_hoff, zk_in, zv_in);
			c->launches++;
			CUB_CALL(c, cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, zk_in, zk_out, zv_in, zv_out, (uint64_t)T, 0, (mode? 49 : 48) + pbits, c->stream));
			if(mode == 0){
				k_unpack<0><<<(unsigned)((T + 255) / 256), 256, 0, c->stream>>>(zk_out, zv_out, T, cache_buf.as<DevZPair>(), d_tie);
				k_pair_offsets<0><<<(np + 1 + 127) / 128, 128, 0, c->stream>>>(zk_out, T, np, d_coff);
			} else {
				k_unpack<1><<<(unsigned)((T + 255) / 256), 256, 0, c->stream>>>(zk_out, zv_out, T, cache_buf.as<DevZPair>(), d_tie);
				k_pair_offsets<1><<<(np + 1 + 127) / 128, 128, 0, c->stream>>>(zk_out, T, np, d_coff);
			}
			c->launches += 2;
		}
	}
	if(T == 0){ if(cache_buf.reserve(64)) return ZMO_ERR_CUDA; CUDA_TRY(cudaMemsetAsync(d_coff, 0, ((size_t)np + 1) * 8, c->stream)); }
	CUDA_TRY(cudaGetLastError());
	W.tie = d_tie; W.pc = d_pc;
	W.T = T; W.cache_off = d_coff; W.cache = cache_buf.as<DevZPair>();
	c->counters[3] += T;
	return 0;
}

extern "C" int zmo_pair_windows(zmo_ctx *c, int slot, const zmo_pair_t *pairs, uint32_t np, zmo_pairseed_t *seeds, zmo_window_t *wins, uint64_t win_cap, uint64_t *win_needed){
	if(!c || (np && (!pairs || !seeds))) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	if(slot < 0 || slot > 1) return zmo_set_err(ZMO_ERR_ARG, "slot must be 0 or 1");
	if(c->st->n_reads == 0) return zmo_set_err(ZMO_ERR_STATE, "no reads uploaded");
	if(win_needed) *win_needed = 0;
	SeedSlot &SL = c->slot[slot];
	SL.np = 0; SL.n_wins = SL.n_anchors = 0;
	if(np == 0) return 0;
	CUDA_TRY(cudaSetDevice(c->device));
	StageTimer tm(c, ST_SEED);
	SeedPar par; par.zsize = c->par.zsize; par.kwin = c->par.kwin; par.kstep = c->par.kstep; par.zovl = c->par.zovl; par.ztot = c->par.ztot; par.W = c->par.W;
	unsigned long long *ctr = c->d_ctr.as<unsigned long long>();
	unsigned long long nw = 0, na = 0;
	for(int attempt = 0; attempt < 4; attempt++){
		SeedWork W; const uint32_t F = 2u << (2 * attempt); const size_t per = zmo_pair_scratch_per(F);     /* 2, 8, 32, 128 */
		if(int rc = seed_prepare(c, pairs, np, 0, W, c->s5)) return rc;      /* match lists live in s5 */
		const unsigned long long T = W.T;
		const unsigned long long cap_w = (attempt? 2 * T * F : T / 2) + 64, cap_a = 2 * T * F + 64;
		if(c->s6.reserve(T * per + (size_t)64 * np + 256) || SL.wins.reserve(cap_w * sizeof(DevWin)) || SL.anchors.reserve(cap_a * sizeof(DevZPair)) || SL.seeds.reserve((size_t)np * sizeof(zmo_pairseed_t)) || SL.pairs.reserve((size_t)np * sizeof(zmo_pair_t))) return ZMO_ERR_CUDA;
		CUDA_TRY(cudaMemsetAsync(ctr + CTR_N1, 0, 24, c->stream));
		SeedOut O; O.wins = SL.wins.as<DevWin>(); O.anc = SL.anchors.as<DevZPair>(); O.cap_wins = cap_w; O.cap_anc = cap_a; O.cur_wins = ctr + CTR_N1; O.cur_anc = ctr + CTR_N2; O.overflow = ctr + CTR_N3;
		CUDA_TRY(cudaMemsetAsync(ctr + CTR_WORK, 0, 8, c->stream));
		{
			static bool attr_set = false;
			if(!attr_set){ CUDA_TRY(cudaFuncSetAttribute(k_p_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PS_WARPS * sizeof(PSSmem)))); attr_set = true; }
			const int grid = (int)std::min<uint64_t>((np + PS_WARPS - 1) / PS_WARPS, (uint64_t)c->n_sm);      /* one CTA of PS_WARPS warps per SM (shared-memory bound) */
			k_p_seed<<<grid, 32 * PS_WARPS, PS_WARPS * sizeof(PSSmem), c->stream>>>(W.cache_off, np, W.cache, W.tie, W.pc, dev_reads(c), c->s6.as<uint8_t>(), per, F, par, O, SL.seeds.as<zmo_pairseed_t>(), ctr + CTR_WORK); c->launches++;
		}
		CUDA_TRY(cudaGetLastError());
		unsigned long long h[3];
		CUDA_TRY(cudaMemcpyAsync(h, ctr + CTR_N1, 24, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		if(h[2] == 0){ nw = h[0]; na = h[1]; break; }
		if(attempt == 3) return zmo_set_err(ZMO_ERR_CAPACITY, "window arena overflow");
	}
	if(win_needed) *win_needed = nw;
	if(nw > win_cap) return zmo_set_err(ZMO_ERR_CAPACITY, "window buffer too small: need %llu", nw);
	CUDA_TRY(cudaMemcpyAsync(SL.pairs.p, pairs, (size_t)np * sizeof(zmo_pair_t), cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(seeds, SL.seeds.p, (size_t)np * sizeof(zmo_pairseed_t), cudaMemcpyDeviceToHost, c->stream));
	std::vector<DevWin> hw(nw);
	if(nw) CUDA_TRY(cudaMemcpyAsync(hw.data(), SL.wins.p, nw * sizeof(DevWin), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	SL.h_wspan.resize(nw * 3);
	for(unsigned long long i = 0; i < nw; i++){
		wins[i].beg[0] = hw[i].beg[0]; wins[i].beg[1] = hw[i].beg[1]; wins[i].end[0] = hw[i].end[0]; wins[i].end[1] = hw[i].end[1];
		SL.h_wspan[3 * i] = hw[i].end[0] - hw[i].beg[0]; SL.h_wspan[3 * i + 1] = hw[i].end[1] - hw[i].beg[1]; SL.h_wspan[3 * i + 2] = (int32_t)(hw[i].anc1 - hw[i].anc0);
	}
	SL.np = np; SL.n_wins = nw; SL.n_anchors = na;
	SL.h_seeds.assign(seeds, seeds + np);
	c->counters[5] += (size_t)np * sizeof(zmo_pair_t);
	c->counters[6] += (size_t)np * sizeof(zmo_pairseed_t) + nw * sizeof(DevWin);
	return 0;
}
