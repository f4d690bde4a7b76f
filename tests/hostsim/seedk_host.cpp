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
thread_local std::string g_zmo_err;
int zmo_set_err(int code, const char *, ...){ return code; }
namespace emu { Block *g_blk = nullptr; }
#include "../../smartdenovo_b200/csrc/zmo_seed_kernels.cuh"
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
	emu::launch(np > 1? 2u : 1u, 32 * PS_WARPS, [=](){ k_p_seed(dco, np, dc, dt, dpc, R, ds, per, FF, par, O, dsd, work); }, PS_WARPS * sizeof(PSSmem));
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
