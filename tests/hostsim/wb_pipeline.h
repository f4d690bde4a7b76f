/*
 * TEST-ONLY: host restatement of the bridge-level part of pair_align_impl (smartdenovo_b200/csrc/zmo_align.cu) for the simulated kernels of
 * zmo_winbridge.cuh: per-item step offsets and scratch bounds, passes over item ranges when the bound exceeds the budget, exclusive scan and
 * descending sort of the bridge list (CUB in the product, std:: here), the kernel sequence k_wb_prep -> k_wb_sweep -> k_wb_ends -> k_wb_walk ->
 * k_wb_stitch, and k_window_align for the windows the pipeline leaves out.  Shared by dp_host.cpp and align_host.cpp; never linked into the product.
 */
#pragma once
#include <vector>
#include <algorithm>

/* win3[3 * w] = {q span, c span, anchors} of window w (SeedSlot::h_wspan).  regs / cig_arena / icig as in pair_align_impl.  budget_words: scratch of one
 * pass.  Returns the number of windows that went through k_window_align, -4 if a scratch bound was violated. */
static inline long sim_wb_pipeline(const WItem *items, uint32_t nitems, const AlnTask *tasks, const zmo_pair_t *pairs, const DevWin *wins, const DevZPair *anchors, const int *win3,
		DevReads R, AlnPar A, int acap, const unsigned long long *icig, uint32_t *cig_arena, DevReg *regs, unsigned long long budget_words, unsigned long long *cells_out){
	const int w = A.w, wb_ring = wb_cap(w), wb_rw = wb_row_words(w);
	std::vector<unsigned long long> istep(nitems + 1), ibound(nitems + 1);
	unsigned long long nsteps64 = 0; int max_rows = 16;
	for(uint32_t i = 0; i < nitems; i++){
		const int s0 = win3[3 * items[i].win], s1 = win3[3 * items[i].win + 1], na = win3[3 * items[i].win + 2];
		istep[i] = nsteps64; nsteps64 += (unsigned long long)na; ibound[i] = (unsigned long long)(s1 + 16);
		if(s1 + 8 > max_rows) max_rows = s1 + 8;
		if(s0 + 8 > max_rows) max_rows = s0 + 8;
	}
	istep[nitems] = nsteps64;
	std::vector<uint32_t> chunk; unsigned long long scr_cap = 0, acc = 0; chunk.push_back(0);
	for(uint32_t i = 0; i < nitems; i++){
		const unsigned long long na_i = istep[i + 1] - istep[i];
		const unsigned long long b = ibound[i] * (unsigned long long)(wb_rw + 5) + ibound[i] + na_i * (unsigned long long)(w + 24) + 64;
		if(acc && acc + b > budget_words){ chunk.push_back(i); scr_cap = std::max(scr_cap, acc); acc = 0; }
		acc += b;
	}
	chunk.push_back(nitems); scr_cap = std::max(scr_cap, acc) + 1024;
	const int wgrid = 2;
	const int wcol = std::min(max_rows + w, 2 * w + 1);
	unsigned long long slab = (unsigned long long)max_rows * band_row_words<32, WA_C>(wcol) + max_rows + (2ull * max_rows + 2ull * w + 16) + ((unsigned long long)max_rows >> 3) + (w >> 3) + 8;
	if(2 * w + 3 > WA_CAP){ unsigned long long cap = 1; while(cap < (unsigned long long)(2 * w + 3)) cap <<= 1; slab += 3 * cap; }
	slab = (slab + 63) & ~63ull;
	const uint32_t nsteps = (uint32_t)nsteps64;
	std::vector<uint32_t> arena(std::max(slab * (unsigned long long)wgrid * WA_WARPS, scr_cap) + 64, 0xDEADBEEFu);
	std::vector<WBStep> steps(nsteps + 1); memset(steps.data(), 0xEE, steps.size() * sizeof(WBStep));
	std::vector<unsigned long long> scrw(nsteps + 2, 0xEEEEEEEEull), scro(nsteps + 2, 0);
	std::vector<uint32_t> aopsv((size_t)(nsteps + 1) * (size_t)(acap + 1), 0xEEEEEEEEu), keys(nsteps + 1), ord(nsteps + 1), skeys(nsteps + 1), sord(nsteps + 1), fb(nitems + 1);
	std::vector<uint8_t> iseq(nitems + 1);
	unsigned long long ctr[8] = {0, 0, 0, 0, 0, 0, 0, 0};       /* 0 work, 1 cells, 2 windows left out (per pass), 3 scratch overflow */
	uint32_t *ar = arena.data(), *dao = aopsv.data(); const unsigned long long *dis = istep.data(); unsigned long long *cp = ctr;
	WBStep *ds = steps.data(); unsigned long long *dsw = scrw.data(), *dso = scro.data(); uint32_t *dk = keys.data(), *dord = ord.data(), *dfb = fb.data(); uint8_t *dq = iseq.data();
	const uint32_t *wd = R.words; long n_fb = 0;
	for(size_t ch = 0; ch + 1 < chunk.size(); ch++){
		const uint32_t i0 = chunk[ch], i1 = chunk[ch + 1], ni = i1 - i0; const unsigned long long st0 = istep[i0], st1 = istep[i1]; const uint32_t ns = (uint32_t)(st1 - st0);
		ctr[0] = 0; ctr[2] = 0;
		emu::launch((unsigned)(((unsigned long long)ni * 32 + 127) / 128), 128, [=](){ k_wb_prep(items + i0, ni, tasks, pairs, wins, anchors, R, A, dis + i0, wb_rw, ds, dao, acap, dsw, dk, dord, dq + i0, dfb, cp + 2); });
		if(ns){
			scro[st0] = 0; for(unsigned long long k = st0; k < st1; k++) scro[k + 1] = scro[k] + scrw[k];
			std::vector<uint32_t> idx(ns); for(uint32_t k = 0; k < ns; k++) idx[k] = k;
			std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b){ return keys[st0 + a] > keys[st0 + b]; });
			for(uint32_t k = 0; k < ns; k++){ skeys[st0 + k] = keys[st0 + idx[k]]; sord[st0 + k] = ord[st0 + idx[k]]; }
			const uint32_t *dsk = skeys.data() + st0, *dsor = sord.data() + st0;
			emu::launch(2, WB_NT, [=](){ k_wb_sweep(ds, dsor, dsk, ns, dso, scr_cap, wd, A.P, ar, wb_ring, wb_rw, cp, cp + 3); }, (size_t)wb_ring * 4 * WB_NT);
			emu::launch((ni + 63) / 64, 64, [=](){ k_wb_ends(ni, items + i0, wins, A, dis + i0, dq + i0, ds, dso, ar, wb_rw, cp + 3, cp, 1); });
			emu::launch((ns + 127) / 128, 128, [=](){ k_wb_walk(ds, dsor, dsk, ns, dso, wd, A.P, ar, wb_rw, cp + 3); });
		}
		emu::launch((unsigned)(((unsigned long long)ni * 32 + 127) / 128), 128, [=](){ k_wb_stitch(items + i0, ni, wins, A, dis + i0, dq + i0, ds, dao, acap, dso, ar, wb_rw, cp + 3, cig_arena, icig + i0, regs + i0); });
		if(ctr[3]) return -4;
		ctr[0] = 0;
		emu::launch((unsigned)wgrid, 32 * WA_WARPS, [=](){ k_window_align(items + i0, ni, tasks, pairs, wins, anchors, R, A, ar, slab, max_rows, cig_arena, icig + i0, regs + i0, cp, 0, 1, dfb, cp + 2); });
		n_fb += (long)ctr[2];
	}
	if(cells_out) *cells_out = ctr[1];
	return n_fb;
}
