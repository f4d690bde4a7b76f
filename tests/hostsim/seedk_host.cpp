/*
 * TEST-ONLY host simulation of the pair-seeding KERNEL k_p_seed and the dot-matrix KERNEL k_p_dot (smartdenovo_b200/csrc/
 * zmo_seed_kernels.cuh, zmo_dot_kernels.cuh with the
 * warp-cooperative span searches of zmo_seed_warp.cuh and the serial core of zmo_seed_core.cuh underneath), compiled for the
 * host against tests/hostsim/emu/cuda_runtime.h: one warp per pair runs as 32 cooperative fibers.  The front end (z-index,
 * z-match, sort by (off1, off2)) is done here on the host the way tests/hostsim/seed_host.cpp does it.
 * Never linked into libzmo_b200.so or wtzmo.
 *
 * build: g++ -O1 -std=c++17 -Itests/hostsim/emu -fPIC -shared tests/hostsim/seedk_host.cpp
 */
#include "cuda_runtime.h"
#include <string>
#include <numeric>
#include <unordered_map>
thread_local std::string g_zmo_err;
int zmo_set_err(int code, const char *, ...){ return code; }
namespace emu { Block *g_blk = nullptr; }
#include "../../smartdenovo_b200/csrc/zmo_seed_kernels.cuh"
#include "../../smartdenovo_b200/csrc/zmo_seed_lanes.cuh"
/* which pair-seeding kernel the simulations launch: 0 = k_p_seed (one pair per warp), G > 0 = k_p_seed_lanes with G pairs per warp */
static int g_seed_lanes = 0;
extern "C" void simk_set_seed_lanes(int g){ g_seed_lanes = g; }
#include "../../smartdenovo_b200/csrc/zmo_seedfront_kernels.cuh"
#include "../../smartdenovo_b200/csrc/zmo_dot_kernels.cuh"

static std::vector<uint32_t> pack(const uint8_t *s, int n){
	std::vector<uint32_t> w((n + 15) / 16 + 4, 0);
	for(int i = 0; i < n; i++) w[i >> 4] |= (uint32_t)(s[i] & 3) << (((~i) & 15) << 1);
	return w;
}

/* match list of (q = pb1, c = pb2) in the reference's emission order (c position order) */
static std::vector<DevZPair> match_list(const uint8_t *pb1, int alen, const uint8_t *pb2, int blen, int zsize, int hz, int zcut, int kvar){
	std::vector<uint32_t> qw = pack(pb1, alen), cw = pack(pb2, blen);
	struct ZE { uint32_t mer; DevZSeed s; };
	std::vector<ZE> ze;
	zmo_scan_kmers(qw.data(), (uint32_t)alen, zsize, hz, [&](uint64_t mer, uint32_t dir, uint32_t off, uint32_t ln){ ZE e; e.mer = (uint32_t)mer; e.s.off = off; e.s.len = (uint16_t)ln; e.s.dir = (uint8_t)dir; e.s.pad = 0; ze.push_back(e); });
	std::stable_sort(ze.begin(), ze.end(), [](const ZE &a, const ZE &b){ return a.mer < b.mer; });
	std::vector<DevZSeed> zs(ze.size()); std::vector<DevSlot> slots;
	for(size_t i = 0; i < ze.size(); i++) zs[i] = ze[i].s;
	for(size_t i = 0, j; i < ze.size(); i = j){
		for(j = i + 1; j < ze.size() && ze[j].mer == ze[i].mer; j++);
		if(j - i < (size_t)zcut){ DevSlot s; s.mer = ze[i].mer; s.off = (uint32_t)i; s.cnt = (uint32_t)(j - i); slots.push_back(s); }
	}
	std::vector<uint8_t> kc(slots.size() + 1, 0);
	uint32_t n = zmo_zmatch(cw.data(), (uint32_t)blen, slots.data(), (uint32_t)slots.size(), zs.data(), kc.data(), zsize, hz, (uint32_t)zcut, (uint32_t)kvar, nullptr);
	std::vector<DevZPair> cache(n + 1);
	std::fill(kc.begin(), kc.end(), 0);
	zmo_zmatch(cw.data(), (uint32_t)blen, slots.data(), (uint32_t)slots.size(), zs.data(), kc.data(), zsize, hz, (uint32_t)zcut, (uint32_t)kvar, cache.data());
	cache.resize(n);
	return cache;
}

/*
 * Same contract as seed_host.cpp's sim_pair_windows / the oracle's orc_pair_windows, but the window finding and chaining run
 * through k_p_seed.  The pair is queued `copies` times (each with its own match-list copy and scratch) on a grid of 2 CTAs so
 * that several warps and the work counter are exercised; all copies must agree.  force_tie = 1 marks the pair as having tied
 * sort keys even if it has none (the kernel then rebuilds the emission order and runs the exact sort emulation: same result).
 * F = capacity factor of the pair scratch (the product starts at 2 and retries with 8, 32, 128 on overflow); returns -1 on
 * overflow, -2 if copies disagree.
 */
extern "C" int simk_pair_windows(const uint8_t *pb1, int alen, const uint8_t *pb2, int blen, int zsize, int hz, int zcut, int kvar,
		int kwin, int kstep, int zovl, int ztot, int W, int copies, int force_tie, int F,
		int *n_hzmp, int *ovl, int *win_out, int win_cap, int *anc_out, int anc_cap, int *n_anc_out){
	std::vector<DevZPair> em = match_list(pb1, alen, pb2, blen, zsize, hz, zcut, kvar);
	const uint32_t n = (uint32_t)em.size();
	/* what the device front end delivers: sorted by (off1, off2); the order inside a run of equal keys is an artefact of the
	 * radix pipeline, so make it adversarial (reversed emission order) */
	std::vector<DevZPair> srt = em;
	auto key = [](const DevZPair &z){ return ((uint64_t)z.off1 << 32) | z.off2; };
	std::stable_sort(srt.begin(), srt.end(), [&](const DevZPair &a, const DevZPair &b){ return key(a) < key(b); });
	bool tie = false;
	for(size_t i = 0, j; i < srt.size(); i = j){
		for(j = i + 1; j < srt.size() && key(srt[j]) == key(srt[i]); j++);
		if(j - i > 1){ tie = true; std::reverse(srt.begin() + i, srt.begin() + j); }
	}
	if(copies < 1) copies = 1;
	const uint32_t np = (uint32_t)copies;
	const size_t per = zmo_pair_scratch_per((uint32_t)F);
	std::vector<unsigned long long> coff(np + 1); std::vector<DevZPair> cache((size_t)n * np + 1);
	std::vector<uint8_t> tieflag(np, (uint8_t)((tie || force_tie)? 1 : 0)); std::vector<uint32_t> pc(np, 1u);
	for(uint32_t p = 0; p < np; p++){ coff[p] = (unsigned long long)p * n; std::copy(srt.begin(), srt.end(), cache.begin() + (size_t)p * n); }
	coff[np] = (unsigned long long)np * n;
	std::vector<uint8_t> scratch((size_t)n * np * per + (size_t)64 * np + 256, 0xEE);
	const unsigned long long T = (unsigned long long)n * np;
	const unsigned long long cap_w = 2 * T * F + 64, cap_a = 2 * T * F + 64;
	std::vector<DevWin> wins(cap_w); std::vector<DevZPair> anc(cap_a); std::vector<zmo_pairseed_t> seeds(np);
	unsigned long long ctr[4] = {0, 0, 0, 0};      /* [0] work, [1] windows, [2] anchors, [3] overflow */
	SeedOut O; O.wins = wins.data(); O.anc = anc.data(); O.cap_wins = cap_w; O.cap_anc = cap_a; O.cur_wins = ctr + 1; O.cur_anc = ctr + 2; O.overflow = ctr + 3;
	SeedPar par; par.zsize = zsize; par.kwin = kwin; par.kstep = kstep; par.zovl = zovl; par.ztot = ztot; par.W = W;
	uint64_t woff[2] = {0, 0}; uint32_t len[2] = {(uint32_t)alen, (uint32_t)blen};
	DevReads R; R.words = nullptr; R.woff = woff; R.len = len; R.n = 2;
	const unsigned long long *dco = coff.data(); DevZPair *dc = cache.data(); const uint8_t *dt = tieflag.data(); const uint32_t *dpc = pc.data();
	uint8_t *ds = scratch.data(); zmo_pairseed_t *dsd = seeds.data(); unsigned long long *work = ctr;
	const uint32_t FF = (uint32_t)F;
	{ const int gl = g_seed_lanes; emu::launch(np > 1? 2u : 1u, 32 * PS_WARPS, [=](){ if(gl) k_p_seed_lanes(dco, np, dc, dt, dpc, R, ds, per, FF, par, O, dsd, work, (uint32_t)gl); else k_p_seed(dco, np, dc, dt, dpc, R, ds, per, FF, par, O, dsd, work); }, PS_WARPS * sizeof(PSSmem)); }
	if(ctr[3]) return -1;
	*n_hzmp = (int)seeds[0].n_zpair; ovl[0] = seeds[0].ovl[0]; ovl[1] = seeds[0].ovl[1];
	auto emit = [&](const zmo_pairseed_t &S, std::vector<int> &wv, std::vector<int> &av){
		for(int d = 0; d < 2; d++){
			for(uint32_t j = 0; j < S.n_win[d]; j++){
				const DevWin &w = wins[S.win_off[d] + j];
				const int o[7] = {d, w.beg[0], w.end[0], w.beg[1], w.end[1], (int)w.ovl, (int)(w.anc1 - w.anc0)};
				wv.insert(wv.end(), o, o + 7);
				for(uint32_t k = w.anc0; k < w.anc1; k++){ const DevZPair &p = anc[k]; const int a[6] = {(int)p.off1, (int)p.off2, p.len1, p.len2, p.dir1, p.dir2}; av.insert(av.end(), a, a + 6); }
			}
		}
	};
	std::vector<int> w0, a0; emit(seeds[0], w0, a0);
	for(uint32_t p = 1; p < np; p++){
		std::vector<int> w1, a1; emit(seeds[p], w1, a1);
		if(w1 != w0 || a1 != a0 || seeds[p].n_zpair != seeds[0].n_zpair || seeds[p].ovl[0] != seeds[0].ovl[0] || seeds[p].ovl[1] != seeds[0].ovl[1]) return -2;
	}
	/* the oracle reports the windows of every strand that has windows; the kernel only exports strands whose chain weight reaches
	 * ztot (the others are never aligned, wtzmo.c:896-914): the caller filters the expectation accordingly */
	const int nw = (int)(w0.size() / 7), na = (int)(a0.size() / 6);
	for(int i = 0; i < nw && i < win_cap; i++) std::copy(w0.begin() + 7 * i, w0.begin() + 7 * i + 7, win_out + 7 * i);
	for(int i = 0; i < na && i < anc_cap; i++) std::copy(a0.begin() + 6 * i, a0.begin() + 6 * i + 6, anc_out + 6 * i);
	*n_anc_out = na;
	return nw;
}

/*
 * Dot-matrix mode (-U) for one pair through k_p_dot: same contract as seed_host.cpp's sim_pair_dotmatrix / the oracle's
 * orc_pair_dotmatrix.  The match list is delivered the way the device front end does in this mode: sorted by
 * (off1 - off2, off1), runs of equal keys in adversarial order and flagged as ties.  copies / force_tie as above.
 */
extern "C" int simk_pair_dotmatrix(const uint8_t *pb1, int alen, const uint8_t *pb2, int blen, int zsize, int hz, int zcut, int kvar,
		int xvar, int yvar, int min_block_len, int max_overhang, float dev_pen, float gap_pen, int ztot, int copies, int force_tie, int *out){
	std::vector<DevZPair> srt = match_list(pb1, alen, pb2, blen, zsize, hz, zcut, kvar);
	const uint32_t n = (uint32_t)srt.size();
	auto key = [](const DevZPair &z){ return (int64_t)((((int64_t)z.off1 - (int64_t)z.off2) << 32) | (int64_t)z.off1); };
	std::stable_sort(srt.begin(), srt.end(), [&](const DevZPair &a, const DevZPair &b){ return key(a) < key(b); });
	bool tie = false;
	for(size_t i = 0, j; i < srt.size(); i = j){
		for(j = i + 1; j < srt.size() && key(srt[j]) == key(srt[i]); j++);
		if(j - i > 1){ tie = true; std::reverse(srt.begin() + i, srt.begin() + j); }
	}
	if(copies < 1) copies = 1;
	const uint32_t np = (uint32_t)copies;
	const size_t per = 4 + sizeof(DevDiag) + 4 + 4 + sizeof(DevZPairG) + sizeof(DevWin) + 16;      /* zmo_dot.cu: zmo_dot_scratch_bytes(n) = (n+2)*per + 64 */
	std::vector<unsigned long long> coff(np + 1); std::vector<DevZPair> cache((size_t)n * np + 1);
	std::vector<uint8_t> tieflag(np, (uint8_t)((tie || force_tie)? 1 : 0)); std::vector<zmo_pair_t> pairs(np);
	for(uint32_t p = 0; p < np; p++){ coff[p] = (unsigned long long)p * n; std::copy(srt.begin(), srt.end(), cache.begin() + (size_t)p * n); pairs[p].qid = 0; pairs[p].cid = 1; }
	coff[np] = (unsigned long long)np * n;
	std::vector<uint8_t> scratch((size_t)n * np * per + (size_t)(2 * per + 64) * np + 256, 0xEE);
	std::vector<zmo_dotres_t> res(np);
	unsigned long long work = 0;
	uint64_t woff[2] = {0, 0}; uint32_t len[2] = {(uint32_t)alen, (uint32_t)blen};
	DevReads R; R.words = nullptr; R.woff = woff; R.len = len; R.n = 2;
	DotPar par; par.xvar = xvar; par.yvar = yvar; par.min_block_len = min_block_len; par.max_overhang = max_overhang; par.deviation_penalty = dev_pen; par.gap_penalty = gap_pen;
	const unsigned long long *dco = coff.data(); const zmo_pair_t *dp = pairs.data(); DevZPair *dc = cache.data(); const uint8_t *dt = tieflag.data();
	uint8_t *ds = scratch.data(); zmo_dotres_t *dr = res.data(); unsigned long long *dw = &work;
	emu::launch(np > 1? 2u : 1u, 32 * DOT_WARPS, [=](){ k_p_dot(dco, dp, np, dc, dt, ds, per, R, par, (uint32_t)zsize, (uint32_t)ztot, dr, dw); });
	for(uint32_t p = 1; p < np; p++) if(memcmp(&res[p], &res[0], sizeof(zmo_dotres_t))) return -2;
	out[0] = res[0].score; out[1] = res[0].qb; out[2] = res[0].qe; out[3] = res[0].tb; out[4] = res[0].te; out[5] = res[0].strand;
	return (int)res[0].n_zpair;
}

/* ---- the whole seeding stage of a batch of pairs: seed_prepare (zmo_seed.cu) restated with std:: sorts and scans around the
 * product's front-end kernels, then k_p_seed ---- */
template<class T> static void excl_scan(const T *in, T *out, size_t n){ T acc = 0; for(size_t i = 0; i < n; i++){ const T v = in[i]; out[i] = acc; acc += v; } }
static void sort_pairs(std::vector<unsigned long long> &k, std::vector<unsigned long long> &v, size_t n){      /* cub::DeviceRadixSort::SortPairs is stable */
	std::vector<size_t> ix(n); std::iota(ix.begin(), ix.end(), (size_t)0);
	std::stable_sort(ix.begin(), ix.end(), [&](size_t a, size_t b){ return k[a] < k[b]; });
	std::vector<unsigned long long> k2(n), v2(n);
	for(size_t i = 0; i < n; i++){ k2[i] = k[ix[i]]; v2[i] = v[ix[i]]; }
	std::copy(k2.begin(), k2.end(), k.begin()); std::copy(v2.begin(), v2.end(), v.begin());
}

/*
 * reads: nreads sequences (0..3 codes) back to back with lens[]; pairs: np x {qid, cid}.  mode 0: SW path (lists sorted by (off1, off2)),
 * k_p_seed runs and the windows / anchors of pair `which` are returned like simk_pair_windows.  Returns -1 on scratch overflow.
 * tie_out[p] = tie flag of every pair, nz_out[p] = its match count.
 */
extern "C" int simk_batch_windows(const uint8_t *seqs, const int *lens, int nreads, const int *pairs, int np_, int zsize, int hz, int zcut, int kvar,
		int kwin, int kstep, int zovl, int ztot, int W, int F, int which, int *tie_out, int *nz_out,
		int *n_hzmp, int *ovl, int *win_out, int win_cap, int *anc_out, int anc_cap, int *n_anc_out){
	const uint32_t np = (uint32_t)np_;
	/* device read store */
	std::vector<uint32_t> words; std::vector<uint64_t> woff(nreads); std::vector<uint32_t> rlen(nreads);
	{ size_t o = 0; for(int r = 0; r < nreads; r++){ std::vector<uint32_t> w = pack(seqs + o, lens[r]); while(w.size() & 3) w.push_back(0); woff[r] = words.size(); rlen[r] = (uint32_t)lens[r]; words.insert(words.end(), w.begin(), w.end()); o += (size_t)lens[r]; } }
	DevReads R; R.words = words.data(); R.woff = woff.data(); R.len = rlen.data(); R.n = (uint32_t)nreads;
	std::vector<uint32_t> uq, pq(np), pc(np); std::unordered_map<uint32_t, uint32_t> qmap;
	for(uint32_t i = 0; i < np; i++){
		const uint32_t q = (uint32_t)pairs[2 * i], c = (uint32_t)pairs[2 * i + 1];
		auto it = qmap.find(q);
		if(it == qmap.end()){ it = qmap.emplace(q, (uint32_t)uq.size()).first; uq.push_back(q); }
		pq[i] = it->second; pc[i] = c;
	}
	const uint32_t nuq = (uint32_t)uq.size();
	std::vector<unsigned long long> zcnt(nuq + 1, 0), zoff(nuq + 2, 0);
	const uint32_t *d_uq = uq.data(), *d_pq = pq.data(), *d_pc = pc.data();
	unsigned long long *d_zcnt = zcnt.data(), *d_zoff = zoff.data();
	const int bs = 32;
	/* chunk table from the read lengths, per-chunk counts -> offsets, read offsets (what seed_prepare does with a CUB scan) */
	std::vector<unsigned long long> zchoff(nuq + 1, 0);
	unsigned long long NCz = 0;
	for(uint32_t u = 0; u < nuq; u++){ zchoff[u] = NCz; NCz += (R.len[d_uq[u]] + ZSCAN_CH - 1) / ZSCAN_CH; }
	zchoff[nuq] = NCz;
	std::vector<unsigned long long> zccnt(NCz + 2, 0), zccoff(NCz + 2, 0);
	const unsigned long long *d_zchoff = zchoff.data(); unsigned long long *d_zccnt = zccnt.data(), *d_zccoff = zccoff.data();
	(void)bs; (void)d_zcnt;
	emu::launch((unsigned)((NCz + 127) / 128), 128, [=](){ k_z_scan<0>(R, d_uq, nuq, d_zchoff, NCz, zsize, hz, d_zccnt, nullptr, nullptr); });
	zccnt[NCz] = 0; excl_scan(zccnt.data(), zccoff.data(), (size_t)NCz + 1);
	emu::launch((nuq + 1 + 127) / 128, 128, [=](){ k_z_readoff(d_zchoff, d_zccoff, nuq, d_zoff); });
	const unsigned long long Z = zoff[nuq], Zp = Z + 4;
	std::vector<unsigned long long> zk(Zp), zv(Zp); std::vector<uint32_t> flag(Zp, 0), pos(Zp, 0), run(Zp, 0); std::vector<DevZSeed> zs(Zp); std::vector<DevSlot> slots(Zp);
	std::vector<uint32_t> filt((size_t)(nuq + 1) * ZF_WORDS, 0u), slot_beg(nuq + 2, 0);
	uint32_t NS = 0;
	unsigned long long *d_zk = zk.data(), *d_zv = zv.data(); uint32_t *d_flag = flag.data(), *d_pos = pos.data(), *d_run = run.data(); DevZSeed *d_zs = zs.data(); DevSlot *d_slots = slots.data();
	uint32_t *d_filt = filt.data(), *d_sb = slot_beg.data();
	if(Z){
		emu::launch((unsigned)((NCz + 127) / 128), 128, [=](){ k_z_scan<1>(R, d_uq, nuq, d_zchoff, NCz, zsize, hz, d_zccoff, d_zk, d_zv); });
		sort_pairs(zk, zv, (size_t)Z);
		emu::launch((unsigned)((Z + 255) / 256), 256, [=](){ k_z_heads(d_zk, d_zv, Z, (uint32_t)zcut, d_flag, d_run, d_zs); });
		excl_scan(flag.data(), pos.data(), (size_t)Z);
		NS = pos[Z - 1] + flag[Z - 1];
		emu::launch((unsigned)((Z + 255) / 256), 256, [=](){ k_z_slots(d_zk, d_flag, d_pos, d_run, Z, d_zoff, d_slots, d_filt); });
	}
	{ const uint32_t ns = NS; emu::launch((nuq + 1 + 127) / 128, 128, [=](){ k_z_ranges(d_zoff, d_pos, nuq, Z, ns, d_sb); }); }
	ZIdxView ZV; ZV.slots = d_slots; ZV.slot_beg = d_sb; ZV.zs = d_zs; ZV.zoff = d_zoff; ZV.filt = d_filt;
	std::vector<unsigned long long> pnch(np + 1, 0), pchoff(np + 2, 0), coff(np + 2, 0); std::vector<uint8_t> tie(np + 1, 0);
	unsigned long long *d_pnch = pnch.data(), *d_pchoff = pchoff.data(), *d_coff = coff.data(); uint8_t *d_tie = tie.data();
	emu::launch((np + 127) / 128, 128, [=](){ k_c_nchunks(R, d_pc, np, d_pnch); });
	pnch[np] = 0; excl_scan(pnch.data(), pchoff.data(), (size_t)np + 1);
	const unsigned long long NC = pchoff[np];
	std::vector<unsigned long long> ccnt(NC + 1, 0), choff(NC + 2, 0);
	unsigned long long *d_ccnt = ccnt.data(), *d_choff = choff.data();
	emu::launch((unsigned)((NC + 127) / 128), 128, [=](){ k_hit<0>(R, ZV, d_pq, d_pc, np, d_pchoff, NC, zsize, hz, d_ccnt, nullptr, nullptr); });
	ccnt[NC] = 0; excl_scan(ccnt.data(), choff.data(), (size_t)NC + 1);
	const unsigned long long NH = choff[NC];
	unsigned long long T = 0; std::vector<DevZPair> cache(4);
	if(NH){
		std::vector<unsigned long long> hk(NH + 1), hv(NH + 1), hcnt(NH + 1, 0), hoff(NH + 2, 0);
		unsigned long long *d_hk = hk.data(), *d_hv = hv.data(), *d_hcnt = hcnt.data(), *d_hoff = hoff.data();
		emu::launch((unsigned)((NC + 127) / 128), 128, [=](){ k_hit<1>(R, ZV, d_pq, d_pc, np, d_pchoff, NC, zsize, hz, d_choff, d_hk, d_hv); });
		sort_pairs(hk, hv, (size_t)NH);
		emu::launch((unsigned)((NH + 127) / 128), 128, [=](){ k_expand<0, 0>(R, ZV, d_pq, d_pc, d_hk, d_hv, NH, (uint32_t)zcut, (uint32_t)kvar, d_hcnt, nullptr, nullptr); });
		hcnt[NH] = 0; excl_scan(hcnt.data(), hoff.data(), (size_t)NH + 1);
		T = hoff[NH];
		if(T){
			std::vector<unsigned long long> mk(T + 1), mv(T + 1); cache.assign(T + 4, DevZPair());
			unsigned long long *d_mk = mk.data(), *d_mv = mv.data(); DevZPair *d_cache = cache.data();
			emu::launch((unsigned)((NH + 127) / 128), 128, [=](){ k_expand<1, 0>(R, ZV, d_pq, d_pc, d_hk, d_hv, NH, (uint32_t)zcut, (uint32_t)kvar, d_hoff, d_mk, d_mv); });
			sort_pairs(mk, mv, (size_t)T);
			const unsigned long long TT = T;
			emu::launch((unsigned)((T + 255) / 256), 256, [=](){ k_unpack<0>(d_mk, d_mv, TT, d_cache, d_tie); });
			emu::launch((np + 1 + 127) / 128, 128, [=](){ k_pair_offsets<0>(d_mk, TT, np, d_coff); });
		}
	}
	for(uint32_t p = 0; p < np; p++){ tie_out[p] = tie[p]; nz_out[p] = (int)(coff[p + 1] - coff[p]); }
	/* windows + chain */
	const size_t per = zmo_pair_scratch_per((uint32_t)F);
	std::vector<uint8_t> scratch((size_t)T * per + (size_t)64 * np + 256, 0xEE);
	const unsigned long long cap_w = 2 * T * F + 64, cap_a = 2 * T * F + 64;
	std::vector<DevWin> wins(cap_w); std::vector<DevZPair> anc(cap_a); std::vector<zmo_pairseed_t> seeds(np);
	unsigned long long ctr[4] = {0, 0, 0, 0};
	SeedOut O; O.wins = wins.data(); O.anc = anc.data(); O.cap_wins = cap_w; O.cap_anc = cap_a; O.cur_wins = ctr + 1; O.cur_anc = ctr + 2; O.overflow = ctr + 3;
	SeedPar par; par.zsize = zsize; par.kwin = kwin; par.kstep = kstep; par.zovl = zovl; par.ztot = ztot; par.W = W;
	DevZPair *dc = cache.data(); uint8_t *ds = scratch.data(); zmo_pairseed_t *dsd = seeds.data(); unsigned long long *work = ctr; const uint32_t FF = (uint32_t)F;
	{ const int gl = g_seed_lanes; emu::launch(std::min<uint32_t>((np + PS_WARPS - 1) / PS_WARPS, 2u), 32 * PS_WARPS, [=](){ if(gl) k_p_seed_lanes(d_coff, np, dc, d_tie, d_pc, R, ds, per, FF, par, O, dsd, work, (uint32_t)gl); else k_p_seed(d_coff, np, dc, d_tie, d_pc, R, ds, per, FF, par, O, dsd, work); }, PS_WARPS * sizeof(PSSmem)); }
	if(ctr[3]) return -1;
	const zmo_pairseed_t &S = seeds[which];
	*n_hzmp = (int)S.n_zpair; ovl[0] = S.ovl[0]; ovl[1] = S.ovl[1];
	int nw = 0, na = 0;
	for(int d = 0; d < 2; d++){
		for(uint32_t j = 0; j < S.n_win[d]; j++){
			const DevWin &w = wins[S.win_off[d] + j];
			if(nw < win_cap){ int *o = win_out + 7 * nw; o[0] = d; o[1] = w.beg[0]; o[2] = w.end[0]; o[3] = w.beg[1]; o[4] = w.end[1]; o[5] = (int)w.ovl; o[6] = (int)(w.anc1 - w.anc0); }
			nw++;
			for(uint32_t k = w.anc0; k < w.anc1; k++){
				const DevZPair &p = anc[k];
				if(na < anc_cap){ int *o = anc_out + 6 * na; o[0] = (int)p.off1; o[1] = (int)p.off2; o[2] = p.len1; o[3] = p.len2; o[4] = p.dir1; o[5] = p.dir2; }
				na++;
			}
		}
	}
	*n_anc_out = na;
	return nw;
}
