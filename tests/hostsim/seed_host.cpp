/*
 * TEST-ONLY host build of smartdenovo_b200/csrc/zmo_seed_core.cuh (the per-pair seeding logic that
 * the product runs inside CUDA kernels, one thread per pair).  Lets the CPU-only test-suite compare
 * that exact source against the oracle.  Never linked into libzmo_b200.so or wtzmo.
 */
#include <vector>
#include <algorithm>
#include <cstring>
#include "../../smartdenovo_b200/csrc/zmo_seed_core.cuh"

static std::vector<uint32_t> pack(const uint8_t *s, int n){
	std::vector<uint32_t> w((n + 15) / 16 + 4, 0);
	for(int i = 0; i < n; i++) w[i >> 4] |= (uint32_t)(s[i] & 3) << (((~i) & 15) << 1);
	return w;
}

extern "C" int sim_pair_windows(const uint8_t *pb1, int alen, const uint8_t *pb2, int blen, int zsize, int hz, int zcut, int kvar,
		int kwin, int kstep, int zovl, int ztot, int W, int *n_hzmp, int *ovl, int *win_out, int win_cap, int *anc_out, int anc_cap, int *n_anc_out){
	std::vector<uint32_t> qw = pack(pb1, alen), cw = pack(pb2, blen);
	/* z-index of q: scan, stable sort by mer (emission is in off order), slots with cnt < zcut */
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
	*n_hzmp = (int)n; ovl[0] = ovl[1] = 0;
	int nw = 0, na = 0;
	if(n * (uint32_t)zsize >= (uint32_t)ztot){
		SeedPar par; par.zsize = zsize; par.kwin = kwin; par.kstep = kstep; par.zovl = zovl; par.ztot = ztot; par.W = W;
		zmo_ref_sort(cache.data(), (size_t)n, GtZPairOff12());
		std::vector<uint8_t> scr(zmo_pair_scratch_bytes(n, 8));
		for(int d = 0; d < 2; d++){
			PairScratch P = zmo_pair_scratch_carve(scr.data(), n, 8); uint32_t nwin = 0; int ovf = 0;
			ovl[d] = zmo_pair_seed_strand(cache.data(), n, d, par, P, &nwin, &ovf);
			if(ovf) return -1;
			for(uint32_t j = 0; j < nwin; j++){
				const DevWin &w = P.w2[j];
				if(w.closed) continue;
				if(nw < win_cap){ int *o = win_out + 7 * nw; o[0] = d; o[1] = w.beg[0]; o[2] = w.end[0]; o[3] = w.beg[1]; o[4] = w.end[1]; o[5] = (int)w.ovl; o[6] = (int)(w.anc1 - w.anc0); }
				nw++;
				for(uint32_t k = w.anc0; k < w.anc1; k++){
					const DevZPair &p = P.a2[k];
					if(na < anc_cap){ int *o = anc_out + 6 * na; o[0] = p.off1; o[1] = p.off2; o[2] = p.len1; o[3] = p.len2; o[4] = p.dir1; o[5] = p.dir2; }
					na++;
				}
			}
		}
	}
	*n_anc_out = na;
	return nw;
}

#include "../../smartdenovo_b200/csrc/zmo_dot_core.cuh"
extern "C" int sim_pair_dotmatrix(const uint8_t *pb1, int alen, const uint8_t *pb2, int blen, int zsize, int hz, int zcut, int kvar,
		int xvar, int yvar, int min_block_len, int max_overhang, float dev_pen, float gap_pen, int *out){
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
	DotPar par; par.xvar = xvar; par.yvar = yvar; par.min_block_len = min_block_len; par.max_overhang = max_overhang; par.deviation_penalty = dev_pen; par.gap_penalty = gap_pen;
	std::vector<uint8_t> scr(zmo_dot_scratch_bytes(n));
	DotRes r = zmo_dot_pair(cache.data(), n, alen, blen, par, scr.data(), 0);
	out[0] = r.score; out[1] = r.qb; out[2] = r.qe; out[3] = r.tb; out[4] = r.te; out[5] = r.strand;
	return (int)n;
}

/* whole-read scan vs chunk-parallel scan: returns 1 if the concatenation of chunk emissions equals the full scan */
extern "C" int sim_scan_chunks_equal(const uint8_t *seq, int len, int k, int hp, int chunk){
	std::vector<uint32_t> w = pack(seq, len);
	struct E { uint64_t mer; uint32_t dir, off, ln; bool operator==(const E &o) const { return mer == o.mer && dir == o.dir && off == o.off && ln == o.ln; } };
	std::vector<E> a, b;
	zmo_scan_kmers(w.data(), (uint32_t)len, k, hp, [&](uint64_t mer, uint32_t dir, uint32_t off, uint32_t ln){ a.push_back(E{mer, dir, off, ln}); });
	for(int s = 0; s < len; s += chunk)
		zmo_scan_kmers_chunk(w.data(), (uint32_t)len, k, hp, (uint32_t)s, (uint32_t)(s + chunk), [&](uint64_t mer, uint32_t dir, uint32_t off, uint32_t ln){ b.push_back(E{mer, dir, off, ln}); });
	return a.size() == b.size() && std::equal(a.begin(), a.end(), b.begin());
}
