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
#include "zmo_seed_lanes.cuh"
#include "zmo_seedfront_kernels.cuh"

#define CUB_CALL(c, call_expr) do { size_t _tb = 0; void *_tp = nullptr; { auto d_temp = _tp; size_t &temp_bytes = _tb; CUDA_TRY(call_expr); } \
	if((c)->cubtmp.reserve(_tb + 256)) return ZMO_ERR_CUDA; { void *d_temp = (c)->cubtmp.p; size_t &temp_bytes = _tb; CUDA_TRY(call_expr); } (c)->launches++; } while(0)

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
	unsigned long long *d_zoff = c->s1.as<unsigned long long>() + nuq + 1;
	CUDA_TRY(cudaMemcpyAsync(d_uq, uq.data(), (size_t)nuq * 4, cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(d_pq, pq.data(), (size_t)np * 4, cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(d_pc, pc.data(), (size_t)np * 4, cudaMemcpyHostToDevice, c->stream));
	c->counters[5] += ((size_t)nuq + 2 * (size_t)np) * 4;
	/* z-mer scan, chunk-parallel: chunk table from the host copy of the read lengths (no device round trip), per-chunk counts and offsets
	 * behind the slot filters in zfilt */
	std::vector<unsigned long long> h_choff(nuq + 1);
	unsigned long long NCz = 0;
	for(uint32_t u = 0; u < nuq; u++){ h_choff[u] = NCz; NCz += (c->st->h_rdlen[uq[u]] + ZSCAN_CH - 1) / ZSCAN_CH; }
	h_choff[nuq] = NCz;
	const size_t filt_bytes = (size_t)(nuq + 1) * ZF_WORDS * 4;
	if(c->zfilt.reserve(filt_bytes + (NCz + 2) * 16 + ((size_t)nuq + 2) * 8 + 64)) return ZMO_ERR_CUDA;
	unsigned long long *d_zccnt = (unsigned long long*)((uint8_t*)c->zfilt.p + filt_bytes), *d_zccoff = d_zccnt + NCz + 1, *d_zchoff = d_zccoff + NCz + 1;
	CUDA_TRY(cudaMemcpyAsync(d_zchoff, h_choff.data(), ((size_t)nuq + 1) * 8, cudaMemcpyHostToDevice, c->stream));
	c->counters[5] += ((size_t)nuq + 1) * 8;
	unsigned long long Z = 0;
	if(NCz){
		k_z_scan<0><<<(unsigned)((NCz + 127) / 128), 128, 0, c->stream>>>(R, d_uq, nuq, d_zchoff, NCz, c->par.zsize, c->par.hz, d_zccnt, nullptr, nullptr); c->launches++;
		CUDA_TRY(cudaMemsetAsync(d_zccnt + NCz, 0, 8, c->stream));
		CUB_CALL(c, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_zccnt, d_zccoff, (uint64_t)NCz + 1, c->stream));
		k_z_readoff<<<(nuq + 1 + 127) / 128, 128, 0, c->stream>>>(d_zchoff, d_zccoff, nuq, d_zoff); c->launches++;
		CUDA_TRY(cudaMemcpyAsync(&Z, d_zccoff + NCz, 8, cudaMemcpyDeviceToHost, c->stream));
	} else CUDA_TRY(cudaMemsetAsync(d_zoff, 0, ((size_t)nuq + 1) * 8, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if(Z >= 0xFFFFFFF0ull) return zmo_set_err(ZMO_ERR_CAPACITY, "pair batch too large (%llu z-mers)", Z);
	/* s2 keys_in, s3 keys_out, s4 vals_in, s5 vals_out, s6 flag|pos|runlen, s7 zs|slots */
	const unsigned long long Zp = Z + 4;
	if(c->s2.reserve(Zp * 8) || c->s3.reserve(Zp * 8) || c->s4.reserve(Zp * 8) || c->s5.reserve(Zp * 8) || c->s6.reserve(Zp * 12) || c->s7.reserve(Zp * (sizeof(DevZSeed) + sizeof(DevSlot)))) return ZMO_ERR_CUDA;
	unsigned long long *k_in = c->s2.as<unsigned long long>(), *k_out = c->s3.as<unsigned long long>(), *v_in = c->s4.as<unsigned long long>(), *v_out = c->s5.as<unsigned long long>();
	uint32_t *d_flag = c->s6.as<uint32_t>(), *d_pos = d_flag + Zp, *d_run = d_pos + Zp;
	DevZSeed *d_zs = c->s7.as<DevZSeed>(); DevSlot *d_slots = (DevSlot*)(d_zs + Zp);
	uint32_t NS = 0;
	CUDA_TRY(cudaMemsetAsync(c->zfilt.p, 0, filt_bytes, c->stream));
	if(Z){
		k_z_scan<1><<<(unsigned)((NCz + 127) / 128), 128, 0, c->stream>>>(R, d_uq, nuq, d_zchoff, NCz, c->par.zsize, c->par.hz, d_zccoff, k_in, v_in); c->launches++;
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
			if(c->s3.reserve((T + 2) * 16) || c->s4.reserve((T + 2) * 16) || cache_buf.reserve((T + 36) * sizeof(DevZPair))) return ZMO_ERR_CUDA;
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

/* function attributes are per device: called by ctx_init for every context, with that context's device current */
int zmo_seed_init_device(void){
	CUDA_TRY(cudaFuncSetAttribute(k_p_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PS_WARPS * sizeof(PSSmem))));
	CUDA_TRY(cudaFuncSetAttribute(k_p_seed_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PS_WARPS * sizeof(PSSmem))));
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
			/* ZMO_SEED_LANES = pairs per warp of k_p_seed_lanes (0: k_p_seed, one pair per warp; default: as many as keep every warp of the grid busy, at most 8) */
			static const int lanes_env = [](){ const char *e = getenv("ZMO_SEED_LANES"); return e? atoi(e) : -1; }();
			uint32_t G = 0;
			if(lanes_env > 0) G = (uint32_t)std::min(lanes_env, 32);      /* default: k_p_seed (measured: the lane scan saves 44% of the instructions but not time, profiles/r02_experiments_decided.md) */
			/* experiment knobs: warps per CTA (<= PS_WARPS) and CTAs per SM of the seeding kernel: a smaller footprint lets the kernels of the other contexts in flight share the SM */
			static const int warps_env = [](){ const char *e = getenv("ZMO_SEED_WARPS"); const int v = e? atoi(e) : PS_WARPS; return v < 1? 1 : (v > PS_WARPS? PS_WARPS : v); }();
			static const int ctas_env = [](){ const char *e = getenv("ZMO_SEED_CTAS"); const int v = e? atoi(e) : 1; return v < 1? 1 : v; }();
			const uint64_t per_cta = (uint64_t)warps_env * (G? G : 1);
			const int grid = (int)std::min<uint64_t>(((uint64_t)np + per_cta - 1) / per_cta, (uint64_t)c->n_sm * ctas_env);
			if(G == 0) k_p_seed<<<grid, 32 * warps_env, warps_env * sizeof(PSSmem), c->stream>>>(W.cache_off, np, W.cache, W.tie, W.pc, dev_reads(c), c->s6.as<uint8_t>(), per, F, par, O, SL.seeds.as<zmo_pairseed_t>(), ctr + CTR_WORK);
			else k_p_seed_lanes<<<grid, 32 * warps_env, warps_env * sizeof(PSSmem), c->stream>>>(W.cache_off, np, W.cache, W.tie, W.pc, dev_reads(c), c->s6.as<uint8_t>(), per, F, par, O, SL.seeds.as<zmo_pairseed_t>(), ctr + CTR_WORK, G);
			c->launches++;
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
