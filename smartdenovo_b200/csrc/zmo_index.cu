/*
 * zmo_index.cu -- global k-mer index build and candidate-event query on the device.
 *
 * Index (replaces index_wtzmo / midx / msrt, wtzmo.c:227-430): scan the homopolymer-compressed
 * canonical k-mers of reads [beg,end), keep the sampled ones (Jenkins hash of the low 32 bits,
 * wtzmo.c:270-271), radix-sort (k-mer) with the posting (rd_id<<1|dir) as payload, run-length encode
 * into distinct k-mers with 16-bit-saturating counts (wtzmo.c:276), derive K (wtzmo.c:380-393), and
 * keep the postings of k-mers with 1 < count <= K (wtzmo.c:401-406).  The reference's 1024 sharded
 * hash tables are replaced by one sorted array of distinct k-mers (binary search); only posting-list
 * CONTENT is observable, not table layout.
 *
 * Candidate events (replaces the heap merge of query_wtzmo, wtzmo.c:433-562): gather
 * (query, target<<1|strand, query offset, span) tuples of a batch of query reads, stable radix sort by
 * (query, target<<1|strand) -- tuples are emitted in ascending query offset, so each group stays
 * offset-ordered -- then per group ol = sum_i(off_i >= end_{i-1} ? len_i : end_i - end_{i-1}) with
 * uint32 wrap (wtzmo.c:558-560).  Groups with ol >= kovl are returned in ascending key order; the
 * host replays the top-ncand heap quirks (wtzmo.c:521-571).
 *
 * Device-wide sort/scan/run-length primitives come from CUB (shipped with the CUDA toolkit).
 */
#include <cub/cub.cuh>
#include "zmo_ctx.cuh"
#include "zmo_seed_core.cuh"
#include "zmo_index_kernels.cuh"

/* ---------------------------------------------------------------- index build */
#define CUB_CALL(c, call_expr) do { size_t _tb = 0; void *_tp = nullptr; { auto d_temp = _tp; size_t &temp_bytes = _tb; CUDA_TRY(call_expr); } \
	if((c)->cubtmp.reserve(_tb + 256)) return ZMO_ERR_CUDA; { void *d_temp = (c)->cubtmp.p; size_t &temp_bytes = _tb; CUDA_TRY(call_expr); } (c)->launches++; } while(0)

extern "C" int zmo_index_build(zmo_ctx *c, uint32_t beg, uint32_t end, uint32_t *kcut_io, zmo_index_stats_t *stats){
	if(!c || !kcut_io) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	if(c->is_clone) return zmo_set_err(ZMO_ERR_STATE, "the index is built through the root context");
	if(c->st->n_reads == 0) return zmo_set_err(ZMO_ERR_STATE, "no reads uploaded");
	if(end > c->st->n_reads) end = c->st->n_reads;     /* the reference reads past the table here when n_rd % n_idx != 0 (wtzmo.c:1283) */
	if(beg >= end) return zmo_set_err(ZMO_ERR_ARG, "empty read range");
	CUDA_TRY(cudaSetDevice(c->device));
	StageTimer tm(c, ST_INDEX);
	const uint32_t nr = end - beg; const int bs = 64; DevReads R = dev_reads(c);
	DevBuf *const tmpbuf[9] = { &c->s0, &c->s1, &c->s2, &c->s3, &c->s4, &c->s5, &c->s6, &c->s7, &c->cubtmp }; size_t cap0[9];
	for(int k = 0; k < 9; k++) cap0[k] = tmpbuf[k]->cap;
	if(c->s0.reserve(((size_t)nr + 1) * 8) || c->s1.reserve(((size_t)nr + 1) * 8)) return ZMO_ERR_CUDA;
	unsigned long long *d_cnt = c->s0.as<unsigned long long>(), *d_off = c->s1.as<unsigned long long>();
	k_idx_count<<<(nr + bs - 1) / bs, bs, 0, c->stream>>>(R, beg, end, c->par.ksize, c->par.hk, (uint32_t)c->par.ksave, d_cnt); c->launches++;
	CUDA_TRY(cudaMemsetAsync(d_cnt + nr, 0, 8, c->stream));
	CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_cnt, d_off, nr + 1, c->stream));
	unsigned long long N = 0;
	CUDA_TRY(cudaMemcpyAsync(&N, d_off + nr, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if(N == 0){ c->st->n_ent = 0; c->st->n_post = 0; c->st->have_index = true; if(*kcut_io < 2) *kcut_io = 100; c->st->kcut = *kcut_io; if(stats) memset(stats, 0, sizeof(*stats)); return 0; }
	if(c->s2.reserve(N * 8) || c->s3.reserve(N * 8) || c->s4.reserve(N * 4) || c->s5.reserve(N * 4)) return ZMO_ERR_CUDA;
	unsigned long long *k_in = c->s2.as<unsigned long long>(), *k_out = c->s3.as<unsigned long long>();
	uint32_t *v_in = c->s4.as<uint32_t>(), *v_out = c->s5.as<uint32_t>();
	k_idx_fill<<<(nr + bs - 1) / bs, bs, 0, c->stream>>>(R, beg, end, c->par.ksize, c->par.hk, (uint32_t)c->par.ksave, d_off, k_in, v_in); c->launches++;
	CUB_CALL(c, cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, k_in, k_out, v_in, v_out, (uint64_t)N, 0, 2 * c->par.ksize, c->stream));
	/* run-length encode (hand-rolled, 64-bit safe): head flags -> exclusive scan -> scatter of
	 * distinct k-mers (ix_mer) and run starts; counts = difference of consecutive run starts */
	if(N >= 0xFFFFFFF0ull) return zmo_set_err(ZMO_ERR_CAPACITY, "index partition too large (%llu sampled k-mers); split it with -G", N);
	if(c->s0.reserve((N + 2) * 8) || c->s6.reserve(N * 4 + 16) || c->s7.reserve(N * 4 + 16)) return ZMO_ERR_CUDA;
	unsigned long long *ctr = c->d_ctr.as<unsigned long long>();
	uint32_t *d_hflag = c->s6.as<uint32_t>(), *d_hpos = c->s7.as<uint32_t>();
	unsigned long long *run_start = c->s0.as<unsigned long long>();
	k_idx_heads<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>(k_out, N, d_hflag); c->launches++;
	CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_hflag, d_hpos, (uint64_t)N, c->stream));
	uint32_t lp = 0, lf = 0;
	CUDA_TRY(cudaMemcpyAsync(&lp, d_hpos + (N - 1), 4, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(&lf, d_hflag + (N - 1), 4, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	const unsigned long long ne = (unsigned long long)lp + lf;
	if(c->st->ix_mer.reserve((ne + 2) * 8)) return ZMO_ERR_CUDA;      /* distinct k-mers only (cfg4: 7 M of 2.3 G sampled) */
	k_idx_runs<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>(k_out, d_hflag, d_hpos, N, c->st->ix_mer.as<unsigned long long>(), run_start); c->launches++;
	uint32_t *d_rc = v_in;       /* v_in is free after the sort */
	k_idx_counts<<<(unsigned)((ne + 255) / 256), 256, 0, c->stream>>>(run_start, ne, N, d_rc); c->launches++;
	/* K (wtzmo.c:380-393): ktot = sum of saturated counts over all distinct k-mers */
	uint32_t K = *kcut_io, kavg = 0;
	{
		cub::TransformInputIterator<unsigned long long, SatCount, const uint32_t*> it(d_rc, SatCount());
		CUB_CALL(c, cub::DeviceReduce::Sum(d_temp, temp_bytes, it, ctr + CTR_N2, (uint64_t)ne, c->stream));
		unsigned long long ktot = 0;
		CUDA_TRY(cudaMemcpyAsync(&ktot, ctr + CTR_N2, 8, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		kavg = (uint32_t)(ktot / (ne + 1));
		if(K < 2){ uint32_t ka = kavg < 20? 20 : kavg; K = ka * 5; }
	}
	*kcut_io = K; c->st->kcut = K;
	/* filter flags, kept offsets, postings */
	if(c->s1.reserve((ne + 1) * 8) || c->st->ix_off.reserve((ne + 2) * 8) || c->st->ix_flt.reserve(ne + 8)) return ZMO_ERR_CUDA;
	unsigned long long *kept = c->s1.as<unsigned long long>();
	CUDA_TRY(cudaMemsetAsync(ctr + CTR_N3, 0, 16, c->stream));
	k_idx_flags<<<(unsigned)((ne + 255) / 256), 256, 0, c->stream>>>(d_rc, ne, K, c->st->ix_flt.as<uint8_t>(), kept, ctr + CTR_N3); c->launches++;
	CUDA_TRY(cudaMemsetAsync(kept + ne, 0, 8, c->stream));
	CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, kept, c->st->ix_off.as<unsigned long long>(), (uint64_t)ne + 1, c->stream));
	unsigned long long np = 0, st2[2] = {0, 0};
	CUDA_TRY(cudaMemcpyAsync(&np, c->st->ix_off.as<unsigned long long>() + ne, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(st2, ctr + CTR_N3, 16, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if(c->st->ix_post.reserve((np + 4) * 4)) return ZMO_ERR_CUDA;
	k_idx_gather<<<(unsigned)((ne + 255) / 256), 256, 0, c->stream>>>(run_start, c->st->ix_off.as<unsigned long long>(), c->st->ix_flt.as<uint8_t>(), d_rc, ne, v_out, c->st->ix_post.as<uint32_t>()); c->launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	c->st->n_ent = ne; c->st->n_post = np; c->st->have_index = true;
	/* the sort buffers of a large partition (cfg4: 120 GB for 2.3 G sampled k-mers) are not needed by the batches: give them back, the
	 * per-batch stages re-grow what they use (a few GB); buffers that were already that large (the batches' own scratch) and small builds
	 * keep theirs, so that rebuilding the index in a warm session costs no allocation */
	for(int k = 0; k < 9; k++) if(tmpbuf[k]->cap > cap0[k] && tmpbuf[k]->cap > (2ull << 30)) tmpbuf[k]->release();      /* only what THIS build grew */
	if(stats){ stats->n_kmers = ne; stats->n_postings = np; stats->n_filtered_high = st2[0]; stats->n_indexed = st2[1]; stats->kcut = K; stats->kavg = kavg; }
	return 0;
}

/* ---------------------------------------------------------------- candidate events */
extern "C" int zmo_candidates(zmo_ctx *c, const uint32_t *qids, uint32_t nq, uint64_t *ev_off, zmo_event_t *events, uint64_t ev_cap, uint64_t *ev_needed){
	if(!c || !qids || !ev_off) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	if(!c->st->have_index) return zmo_set_err(ZMO_ERR_STATE, "zmo_index_build has not been called");
	if(ev_needed) *ev_needed = 0;
	if(nq == 0){ ev_off[0] = 0; return 0; }
	for(uint32_t i = 0; i < nq; i++) if(qids[i] >= c->st->n_reads) return zmo_set_err(ZMO_ERR_ARG, "query id out of range");
	CUDA_TRY(cudaSetDevice(c->device));
	StageTimer tm(c, ST_CAND);
	DevReads R = dev_reads(c);
	IdxView I; I.mer = c->st->ix_mer.as<unsigned long long>(); I.off = c->st->ix_off.as<unsigned long long>(); I.flt = c->st->ix_flt.as<uint8_t>(); I.post = c->st->ix_post.as<uint32_t>(); I.n = c->st->n_ent;
	if(nq > 65535) return zmo_set_err(ZMO_ERR_ARG, "at most 65535 query reads per zmo_candidates call");
	if(c->s0.reserve((size_t)nq * 4) || c->s1.reserve(((size_t)nq + 2) * 8) || c->s2.reserve(((size_t)nq + 2) * 8)) return ZMO_ERR_CUDA;
	uint32_t *d_q = c->s0.as<uint32_t>(); unsigned long long *d_nch = c->s1.as<unsigned long long>(), *d_off = c->s2.as<unsigned long long>();
	CUDA_TRY(cudaMemcpyAsync(d_q, qids, (size_t)nq * 4, cudaMemcpyHostToDevice, c->stream));
	c->counters[5] += (uint64_t)nq * 4;
	/* chunk table */
	k_q_nchunks<<<(nq + 127) / 128, 128, 0, c->stream>>>(R, d_q, nq, d_nch); c->launches++;
	CUDA_TRY(cudaMemsetAsync(d_nch + nq, 0, 8, c->stream));
	CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_nch, d_off, nq + 1, c->stream));
	unsigned long long NC = 0;
	CUDA_TRY(cudaMemcpyAsync(&NC, d_off + nq, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	/* sampled k-mers of the queries: count per chunk, scan, fill (s3: chunk counts | chunk offsets) */
	if(c->s3.reserve((NC + 2) * 16)) return ZMO_ERR_CUDA;
	unsigned long long *d_ccnt = c->s3.as<unsigned long long>(), *d_coff = d_ccnt + NC + 1;
	k_qk_scan<0><<<(unsigned)((NC + 127) / 128), 128, 0, c->stream>>>(R, d_q, nq, d_off, NC, c->par.ksize, c->par.hk, (uint32_t)c->par.ksave, d_ccnt, nullptr, nullptr); c->launches++;
	CUDA_TRY(cudaMemsetAsync(d_ccnt + NC, 0, 8, c->stream));
	CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_ccnt, d_coff, (uint64_t)NC + 1, c->stream));
	unsigned long long NK = 0;
	CUDA_TRY(cudaMemcpyAsync(&NK, d_coff + NC, 8, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	unsigned long long N = 0;
	unsigned long long *d_kmer = nullptr, *d_kinfo = nullptr, *d_kcnt = nullptr, *d_koff = nullptr; uint32_t *d_ent = nullptr;
	if(NK){
		/* s4: km_mer | km_info ; s5: per-k-mer counts | offsets ; s6: entries */
		if(c->s4.reserve((NK + 2) * 16) || c->s5.reserve((NK + 2) * 16) || c->s6.reserve((NK + 2) * 4)) return ZMO_ERR_CUDA;
		d_kmer = c->s4.as<unsigned long long>(); d_kinfo = d_kmer + NK + 1; d_kcnt = c->s5.as<unsigned long long>(); d_koff = d_kcnt + NK + 1; d_ent = c->s6.as<uint32_t>();
		k_qk_scan<1><<<(unsigned)((NC + 127) / 128), 128, 0, c->stream>>>(R, d_q, nq, d_off, NC, c->par.ksize, c->par.hk, (uint32_t)c->par.ksave, d_coff, d_kmer, d_kinfo); c->launches++;
		k_qk_lookup<<<(unsigned)((NK + 127) / 128), 128, 0, c->stream>>>(I, R, d_q, d_kmer, d_kinfo, NK, d_ent, d_kcnt); c->launches++;
		CUDA_TRY(cudaMemsetAsync(d_kcnt + NK, 0, 8, c->stream));
		CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_kcnt, d_koff, (uint64_t)NK + 1, c->stream));
		CUDA_TRY(cudaMemcpyAsync(&N, d_koff + NK, 8, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
	}
	uint32_t nev = 0;
	if(N){
		if(N >= 0xFFFFFFFFull) return zmo_set_err(ZMO_ERR_CAPACITY, "candidate batch too large (%llu tuples); use a smaller query batch", N);
		/* tuple arrays: keys_in in s3 (the chunk tables there are dead); k_out | v_in | v_out | flag,ol,pos carved from
		 * the DP arena, which is idle during the candidate stage */
		if(c->s3.reserve(N * 8 + 64)) return ZMO_ERR_CUDA;
		unsigned long long *k_in = c->s3.as<unsigned long long>();
		DevBuf &bk = c->arena;
		if(bk.reserve(N * (8 + 8 + 8 + 12) + 256)) return ZMO_ERR_CUDA;
		unsigned long long *k_out = bk.as<unsigned long long>(), *v_in = k_out + N, *v_out = v_in + N;
		uint32_t *d_flag = (uint32_t*)(v_out + N), *d_ol = d_flag + N, *d_pos = d_ol + N;
		k_qk_expand<<<(unsigned)((NK + 127) / 128), 128, 0, c->stream>>>(I, R, d_q, d_kinfo, d_ent, d_koff, NK, k_in, v_in); c->launches++;
		int qbits = 1; while((1ull << qbits) < nq) qbits++;
		CUB_CALL(c, cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, k_in, k_out, v_in, v_out, (uint64_t)N, 0, 32 + qbits, c->stream));
		k_cand_union<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>(k_out, v_out, N, (uint32_t)c->par.kovl, d_flag, d_ol); c->launches++;
		CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_flag, d_pos, (uint64_t)N, c->stream));
		uint32_t lastpos = 0, lastflag = 0;
		CUDA_TRY(cudaMemcpyAsync(&lastpos, d_pos + (N - 1), 4, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaMemcpyAsync(&lastflag, d_flag + (N - 1), 4, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		nev = lastpos + lastflag;
		/* events + their query index reuse k_in / v_in (free after the sort) */
		zmo_event_t *d_ev = (zmo_event_t*)k_in; uint32_t *d_evq = (uint32_t*)v_in;
		k_cand_emit<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>(k_out, d_flag, d_pos, d_ol, N, d_ev, d_evq); c->launches++;
		k_cand_offsets<<<(nq + 1 + 127) / 128, 128, 0, c->stream>>>(d_evq, nev, nq, d_off); c->launches++;
		CUDA_TRY(cudaGetLastError());
		if(ev_needed) *ev_needed = nev;
		if(nev > ev_cap){ CUDA_TRY(cudaStreamSynchronize(c->stream)); return zmo_set_err(ZMO_ERR_CAPACITY, "event buffer too small: need %u", nev); }
		if(nev && !events) return zmo_set_err(ZMO_ERR_ARG, "null event buffer");
		CUDA_TRY(cudaMemcpyAsync(ev_off, d_off, ((size_t)nq + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
		if(nev) CUDA_TRY(cudaMemcpyAsync(events, d_ev, (size_t)nev * sizeof(zmo_event_t), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		c->counters[6] += ((size_t)nq + 1) * 8 + (size_t)nev * sizeof(zmo_event_t);
		c->counters[4] += N;
	} else {
		for(uint32_t i = 0; i <= nq; i++) ev_off[i] = 0;
	}
	return 0;
}
