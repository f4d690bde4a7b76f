/*
 * zmo_index_kernels.cuh -- kernels of the global k-mer index (index_wtzmo, wtzmo.c:227-430) and of the candidate query
 * (query_wtzmo up to the ordered (target, strand, ol) event stream, wtzmo.c:433-562).  Kept in a header so that the test-only host
 * simulation (tests/hostsim/index_host.cpp) runs this very source; included by zmo_index.cu only, which supplies the CUB sorts,
 * scans and reductions between the kernels (zmo_index_build, zmo_candidates).
 */
#pragma once
#include "zmo_ctx.cuh"
#include "zmo_seed_core.cuh"

__global__ void k_idx_count(DevReads R, uint32_t beg, uint32_t end, int ksize, int hk, uint32_t ksave, unsigned long long *cnt){
	uint32_t rid = beg + blockIdx.x * blockDim.x + threadIdx.x;
	if(rid >= end) return;
	unsigned long long n = 0;
	zmo_scan_kmers(R.words + R.woff[rid], R.len[rid], ksize, hk, [&](uint64_t mer, uint32_t, uint32_t, uint32_t){ if(zmo_kmer_sampled(mer, ksave)) n++; });
	cnt[rid - beg] = n;
}
__global__ void k_idx_fill(DevReads R, uint32_t beg, uint32_t end, int ksize, int hk, uint32_t ksave, const unsigned long long *off, unsigned long long *keys, uint32_t *vals){
	uint32_t rid = beg + blockIdx.x * blockDim.x + threadIdx.x;
	if(rid >= end) return;
	unsigned long long p = off[rid - beg];
	zmo_scan_kmers(R.words + R.woff[rid], R.len[rid], ksize, hk, [&](uint64_t mer, uint32_t dir, uint32_t, uint32_t){
		if(zmo_kmer_sampled(mer, ksave)){ keys[p] = mer; vals[p] = (rid << 1) | dir; p++; }
	});
}
__global__ void k_idx_heads(const unsigned long long *keys, unsigned long long n, uint32_t *flag){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]);
}
__global__ void k_idx_runs(const unsigned long long *keys, const uint32_t *flag, const uint32_t *pos, unsigned long long n, unsigned long long *mer, unsigned long long *run_start){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n && flag[i]){ mer[pos[i]] = keys[i]; run_start[pos[i]] = i; }
}
__global__ void k_idx_counts(const unsigned long long *run_start, unsigned long long ne, unsigned long long n, uint32_t *cnt){
	unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(r < ne){ unsigned long long c = (r + 1 < ne? run_start[r + 1] : n) - run_start[r]; cnt[r] = c > 0xFFFFFFFFull? 0xFFFFFFFFu : (uint32_t)c; }
}
struct SatCount { __host__ __device__ unsigned long long operator()(uint32_t c) const { return c > 0xFFFFu? 0xFFFFull : (unsigned long long)c; } };
__global__ void k_idx_flags(const uint32_t *cnt, uint64_t n, uint32_t K, uint8_t *flt, unsigned long long *kept, unsigned long long *stats){
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	uint32_t c = cnt[i];
	bool high = c > 0xFFFFu || c > K, f = high || c <= 1;
	flt[i] = f; kept[i] = f? 0 : c;
	if(high) atomicAdd(stats + 0, 1ULL);
	if(!f) atomicAdd(stats + 1, 1ULL);
}
__global__ void k_idx_gather(const unsigned long long *run_start, const unsigned long long *kept_off, const uint8_t *flt, const uint32_t *cnt, uint64_t n, const uint32_t *vals, uint32_t *post){
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n || flt[i]) return;
	const unsigned long long s = run_start[i], d = kept_off[i]; const uint32_t c = cnt[i];
	for(uint32_t k = 0; k < c; k++) post[d + k] = vals[s + k];
}

struct IdxView { const unsigned long long *mer; const unsigned long long *off; const uint8_t *flt; const uint32_t *post; unsigned long long n; };
__device__ __forceinline__ long long idx_find(const IdxView &I, unsigned long long mer){
	unsigned long long lo = 0, hi = I.n;
	while(lo < hi){ unsigned long long mid = (lo + hi) >> 1; if(I.mer[mid] < mer) lo = mid + 1; else hi = mid; }
	return (lo < I.n && I.mer[lo] == mer)? (long long)lo : -1;
}
/* Candidate query, parallel form: (1) chunk-parallel hp-k-mer scan of the query reads (one thread per
 * 128-base chunk) emits the sampled k-mers in (query, position) order; (2) one thread per k-mer looks it up
 * in the index and counts the postings that survive the self / length filters (wtzmo.c:488-489,509-510);
 * (3) one thread per k-mer writes its tuples key = qlocal<<32 | tkey, val = off<<16 | len at the scanned
 * offset, so tuples of one query stay in ascending query-offset order for the stable sort. */
#define SCAN_CH 128
__global__ void k_q_nchunks(DevReads R, const uint32_t *qids, uint32_t nq, unsigned long long *nch){
	uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if(q < nq) nch[q] = (R.len[qids[q]] + SCAN_CH - 1) / SCAN_CH;
}
template<int PASS>
__global__ void k_qk_scan(DevReads R, const uint32_t *qids, uint32_t nq, const unsigned long long *choff, unsigned long long NC, int ksize, int hk, uint32_t ksave,
		unsigned long long *cnt_or_off, unsigned long long *km_mer, unsigned long long *km_info){
	unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(c >= NC) return;
	uint32_t lo = 0, hi = nq;
	while(lo + 1 < hi){ uint32_t mid = (lo + hi) >> 1; if(choff[mid] <= c) lo = mid; else hi = mid; }
	const uint32_t q = lo, qid = qids[q], s = (uint32_t)(c - choff[q]) * SCAN_CH;
	unsigned long long n = PASS? cnt_or_off[c] : 0;
	zmo_scan_kmers_chunk(R.words + R.woff[qid], R.len[qid], ksize, hk, s, s + SCAN_CH, [&](uint64_t mer, uint32_t, uint32_t off, uint32_t ln){
		if(!zmo_kmer_sampled(mer, ksave)) return;
		if(PASS){ km_mer[n] = mer; km_info[n] = ((unsigned long long)q << 48) | ((unsigned long long)off << 16) | ln; }
		n++;
	});
	if(!PASS) cnt_or_off[c] = n;
}
#define ENT_NONE 0xFFFFFFFFu
__global__ void k_qk_lookup(IdxView I, DevReads R, const uint32_t *qids, const unsigned long long *km_mer, const unsigned long long *km_info, unsigned long long NK, uint32_t *ent, unsigned long long *cnt){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= NK) return;
	const long long e = idx_find(I, km_mer[i]);
	unsigned long long n = 0;
	if(e >= 0 && !I.flt[e]){
		const uint32_t qid = qids[(uint32_t)(km_info[i] >> 48)], up = (uint32_t)((double)R.len[qid] * 1.2);
		const unsigned long long b0 = I.off[e], b1 = I.off[e + 1];
		for(unsigned long long b = b0; b < b1; b++){
			const uint32_t tid = I.post[b] >> 1;
			if(tid == qid || R.len[tid] > up) continue;
			n++;
		}
		ent[i] = (uint32_t)e;
	} else ent[i] = ENT_NONE;
	cnt[i] = n;
}
__global__ void k_qk_expand(IdxView I, DevReads R, const uint32_t *qids, const unsigned long long *km_info, const uint32_t *ent, const unsigned long long *toff, unsigned long long NK,
		unsigned long long *keys, unsigned long long *vals){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= NK || ent[i] == ENT_NONE) return;
	const unsigned long long info = km_info[i]; const uint32_t q = (uint32_t)(info >> 48), qid = qids[q], up = (uint32_t)((double)R.len[qid] * 1.2);
	const unsigned long long val = info & 0xFFFFFFFFFFFFull;       /* off<<16 | len */
	unsigned long long n = toff[i]; const unsigned long long b0 = I.off[ent[i]], b1 = I.off[ent[i] + 1];
	for(unsigned long long b = b0; b < b1; b++){
		const uint32_t tk = I.post[b], tid = tk >> 1;
		if(tid == qid || R.len[tid] > up) continue;
		keys[n] = ((unsigned long long)q << 32) | tk; vals[n] = val; n++;
	}
}
/* one thread per sorted tuple; group heads accumulate the union length of their group */
__global__ void k_cand_union(const unsigned long long *keys, const unsigned long long *vals, unsigned long long n, uint32_t kovl, uint32_t *flag, uint32_t *ol_out){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	const unsigned long long key = keys[i];
	if(i && keys[i - 1] == key){ flag[i] = 0; return; }
	uint32_t ol = 0, lst = 0;
	for(unsigned long long j = i; j < n && keys[j] == key; j++){
		const uint32_t off = (uint32_t)(vals[j] >> 16), ln = (uint32_t)(vals[j] & 0xFFFFu);
		if(off >= lst) ol += ln; else ol += off + ln - lst;
		lst = off + ln;
	}
	flag[i] = ol >= kovl; ol_out[i] = ol;
}
__global__ void k_cand_emit(const unsigned long long *keys, const uint32_t *flag, const uint32_t *pos, const uint32_t *ol, unsigned long long n, zmo_event_t *ev, uint32_t *ev_q){
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n || !flag[i]) return;
	zmo_event_t e; e.tkey = (uint32_t)keys[i]; e.ol = ol[i];
	ev[pos[i]] = e; ev_q[pos[i]] = (uint32_t)(keys[i] >> 32);
}
__global__ void k_cand_offsets(const uint32_t *ev_q, uint32_t nev, uint32_t nq, unsigned long long *ev_off){
	uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if(q > nq) return;
	uint32_t lo = 0, hi = nev;     /* first event with query index >= q */
	while(lo < hi){ uint32_t mid = (lo + hi) >> 1; if(ev_q[mid] < q) lo = mid + 1; else hi = mid; }
	ev_off[q] = lo;
}
