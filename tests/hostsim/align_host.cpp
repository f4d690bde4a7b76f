/*
 * TEST-ONLY host simulation of the whole pair-alignment stage (zmo_pair_align) for one (q, c, strand) task: window alignment (the bridge-level
 * pipeline of zmo_winbridge.cuh through tests/hostsim/wb_pipeline.h, k_window_align for -w beyond its ring) ->
 * k_plan -> DP executor kernels (four extension classes, two gap classes, run from sorted job lists with executor slabs) -> k_plan2 ->
 * right-extension jobs -> k_finish_size -> k_finish_warp -> optional -n refinement kernels.  The kernels are the
 * product's own source (smartdenovo_b200/csrc/zmo_*_kernels.cuh, zmo_winalign.cuh) compiled against tests/hostsim/emu/cuda_runtime.h;
 * only the host orchestration between them (buffer sizing, radix sort of the job keys, prefix sums: pair_align_impl / run_dp_lists
 * in zmo_align.cu, which call CUB and the CUDA runtime) is restated here with std:: algorithms.  Never linked into the product.
 *
 * build: g++ -O1 -std=c++17 -Itests/hostsim/emu -fPIC -shared tests/hostsim/align_host.cpp
 */
#include "cuda_runtime.h"
#include <string>
#include <numeric>
thread_local std::string g_zmo_err;
int zmo_set_err(int code, const char *, ...){ return code; }
namespace emu { Block *g_blk = nullptr; }
#include "../../smartdenovo_b200/csrc/zmo_dp_kernels.cuh"
#include "../../smartdenovo_b200/csrc/zmo_winalign.cuh"
#include "../../smartdenovo_b200/csrc/zmo_winbridge.cuh"
#include "wb_pipeline.h"
#include "../../smartdenovo_b200/csrc/zmo_stitch_kernels.cuh"
#include "../../smartdenovo_b200/csrc/zmo_refine_kernels.cuh"

struct SimReads {
	std::vector<uint32_t> words; uint64_t woff[2]; uint32_t len[2];
	SimReads(const uint8_t *a, int na, const uint8_t *b, int nb){
		const uint8_t *s[2] = {a, b}; const int n[2] = {na, nb};
		for(int r = 0; r < 2; r++){
			woff[r] = words.size(); len[r] = (uint32_t)n[r];
			size_t nw = ((size_t)(n[r] + 15) / 16 + 4 + 3) & ~(size_t)3;
			words.resize(words.size() + nw, 0u);
			for(int i = 0; i < n[r]; i++) words[woff[r] + (i >> 4)] |= (uint32_t)(s[r][i] & 3) << (((~i) & 15) << 1);
		}
	}
	DevReads dev() const { DevReads R; R.words = words.data(); R.woff = woff; R.len = len; R.n = 2; return R; }
};

/* run_dp_lists (zmo_align.cu) for the jobs [first[k], n[k]) of every class: longest-scratch-first order, executor slabs = prefix sum
 * over the first #executors jobs, kernels launched on deliberately small grids so that executors pull several jobs each */
static void run_lists(const JobLists &L, const uint32_t *n, const uint32_t *first, const DevReads &R, const DPPar &P, uint32_t *cig_arena, DPRes *d_res, unsigned long long *ctr){
	for(int k = 0; k < 6; k++){
		const uint32_t off = first? first[k] : 0, cnt = n[k] - off;
		if(cnt == 0) continue;
		const DPJob *jobs = L.list[k] + off;
		std::vector<uint32_t> order(cnt); std::iota(order.begin(), order.end(), 0u);
		std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b){ return jobs[a].sw32 > jobs[b].sw32; });      /* SortPairsDescending is stable */
		const bool warp_cls = (k == 0 || k == 4);
		const unsigned grid = warp_cls? 1u : std::min<uint32_t>(cnt, 2u);
		const uint32_t nex = warp_cls? grid * WRP_PER_CTA : grid;
		std::vector<unsigned long long> soff(nex + 2, 0);
		for(uint32_t e = 0; e < nex + 1; e++) soff[e + 1 <= nex + 1? e : e] = 0;
		{ unsigned long long acc = 0; for(uint32_t e = 0; e <= nex; e++){ soff[e] = acc; acc += e < std::min(nex, cnt)? (unsigned long long)jobs[order[e]].sw32 << 5 : 0ull; } }
		std::vector<uint32_t> arena(soff[nex] + 96, 0xDEADBEEFu);
		DPSlab SB; SB.base = 0; SB.off = soff.data();
		ctr[0] = 0;
		const uint32_t *ord = cnt >= 2? order.data() : nullptr; uint32_t *ar = arena.data();
		if(k == 0) emu::launch(grid, 32 * WRP_PER_CTA, [=](){ k_ext_warp<1>(jobs, ord, cnt, R, P, ar, SB, cig_arena, d_res, ctr, 0, 1); });
		else if(k == 1) emu::launch(grid, 64, [=](){ k_ext_cta<64, 7, 1>(jobs, ord, cnt, R, P, ar, SB, cig_arena, d_res, ctr, 0, 1); });
		else if(k == 2) emu::launch(grid, 128, [=](){ k_ext_cta<128, 7, 1>(jobs, ord, cnt, R, P, ar, SB, cig_arena, d_res, ctr, 0, 1); });
		else if(k == 3) emu::launch(grid, CL3_NT, [=](){ k_ext_cta<CL3_NT, CL3_C, 1>(jobs, ord, cnt, R, P, ar, SB, cig_arena, d_res, ctr, 0, 1); });
		else if(k == 4) emu::launch(grid, 32 * WRP_PER_CTA, [=](){ k_glb_warp(jobs, ord, cnt, R, P, ar, SB, cig_arena, d_res, ctr, 0, 2); });
		else emu::launch(grid, EXT_NT, [=](){ k_glb_cta(jobs, ord, cnt, R, P, ar, SB, cig_arena, d_res, ctr, 0, 2); });
	}
}

/*
 * q = pb1 forward, c FORWARD with strand dir; win = n_win x {q span, c span, n_anchors}, anc = the windows' anchors back to back
 * (6 ints each).  out = {score, tb, te, qb, qe, aln, mat, mis, ins, del}; stats = {jobs per class x 6, windows kept}.  Returns the number
 * of CIGAR ops, -1 if no window region survived (record not ok), -2 on a capacity overflow.
 */
extern "C" int sim_pair_align(const uint8_t *q, int qlen, const uint8_t *c, int clen, int dir, const int *win, int n_win, const int *anc,
		int w, int ew, int Wcap, int zovl, float min_id, int M, int X, int O, int E, int T, int refine, int finish_warp,
		int *out, uint32_t *cigar_out, int cigar_cap, int *stats){
	SimReads rd(q, qlen, c, clen); DevReads R = rd.dev();
	AlnPar A; A.w = w; A.ew = ew; A.W = Wcap; A.zovl = zovl; A.min_id = min_id; A.P.M = M; A.P.X = X; A.P.I = O; A.P.D = O; A.P.E = E; A.P.T = T;
	const uint32_t nt = 1, nitems = (uint32_t)n_win;
	std::vector<WItem> items(nitems + 1); std::vector<DevWin> wins(nitems + 1); std::vector<DevZPair> an; std::vector<unsigned long long> icig(nitems + 1);
	AlnTask task; task.pair_idx = 0; task.dir = (uint32_t)dir; task.item_off = 0; task.n_item = nitems;
	zmo_pair_t pair; pair.qid = 0; pair.cid = 1;
	unsigned long long cig_words = 0; int max_rows = 16, a0 = 0;
	for(uint32_t i = 0; i < nitems; i++){
		const int s0 = win[3 * i], s1 = win[3 * i + 1], na = win[3 * i + 2];
		items[i].task = 0; items[i].win = i;
		DevWin Wd; memset(&Wd, 0, sizeof(Wd)); Wd.anc0 = (uint32_t)a0; Wd.anc1 = (uint32_t)(a0 + na); Wd.dir = (uint8_t)dir; wins[i] = Wd;
		for(int k = 0; k < na; k++){
			const int *o = anc + 6 * (a0 + k); DevZPair p; memset(&p, 0, sizeof(p));
			p.off1 = (uint32_t)o[0]; p.off2 = (uint32_t)o[1]; p.len1 = (uint16_t)o[2]; p.len2 = (uint16_t)o[3]; p.dir1 = (uint8_t)o[4]; p.dir2 = (uint8_t)o[5]; an.push_back(p);
		}
		a0 += na;
		icig[i] = cig_words; cig_words += (unsigned long long)(s0 + s1 + 16 + 2 * na);
		if(s1 + 8 > max_rows) max_rows = s1 + 8;
		if(s0 + 8 > max_rows) max_rows = s0 + 8;
	}
	an.push_back(DevZPair());
	const int wgrid = (int)((nitems + WA_WARPS - 1) / WA_WARPS + 1);
	const int wcol = std::min(max_rows + w, 2 * w + 1);
	unsigned long long slab = (unsigned long long)max_rows * band_row_words<32, WA_C>(wcol) + max_rows + (2ull * max_rows + 2ull * w + 16) + ((unsigned long long)max_rows >> 3) + (w >> 3) + 8;
	if(2 * w + 3 > WA_CAP){ unsigned long long cap = 1; while(cap < (unsigned long long)(2 * w + 3)) cap <<= 1; slab += 3 * cap; }
	slab = (slab + 63) & ~63ull;
	const uint32_t jcap = nitems + 2 * nt + 8;
	const unsigned long long cig_cap_words = cig_words + 4ull * (unsigned long long)(qlen + clen) + 4096;      /* generous: every job's cigar_cap fits */
	std::vector<uint32_t> arena(slab * (unsigned long long)wgrid * WA_WARPS + 64, 0xDEADBEEFu), cig_arena(cig_cap_words + 64, 0u);
	std::vector<DevReg> regs(nitems + 1); std::vector<TaskState> ts(nt + 1); std::vector<DPJob> jobs((size_t)jcap * 6); std::vector<DPRes> res((size_t)jcap * 6);
	unsigned long long ctr[32]; memset(ctr, 0, sizeof(ctr));      /* [0] work, [1] cells ext/win, [2] cells gap, [8] cig cursor, [9] overflow, [10..15] jobs per class */
	const WItem *di = items.data(); const AlnTask *dt = &task; const zmo_pair_t *dp = &pair; const DevWin *dw = wins.data(); const DevZPair *da = an.data();
	uint32_t *ar = arena.data(), *cgp = cig_arena.data(); const unsigned long long *dic = icig.data(); DevReg *dr = regs.data(); unsigned long long *cp = ctr;
	TaskState *dts = ts.data(); DPJob *dj = jobs.data(); DPRes *dres = res.data();
	if(nitems && w >= 1 && w <= WB_MAX_W){
		/* as the product: bridge-level pipeline (zmo_winbridge.cuh), k_window_align for the windows it leaves out */
		unsigned long long cells = 0;
		if(sim_wb_pipeline(di, nitems, dt, dp, dw, da, win, R, A, 2 * 10 + 2, dic, cgp, dr, 1ull << 40, &cells) < 0) return -2;
		ctr[1] += cells;
	} else if(nitems) emu::launch((unsigned)wgrid, 32 * WA_WARPS, [=](){ k_window_align(di, nitems, dt, dp, dw, da, R, A, ar, slab, max_rows, cgp, dic, dr, cp, 0, 1, nullptr, nullptr); });
	stats[6] = 0; for(uint32_t i = 0; i < nitems; i++) stats[6] += regs[i].kept? 1 : 0;
	JobLists L; L.cap = jcap;
	for(int k = 0; k < 6; k++){ L.list[k] = dj + (size_t)k * jcap; L.cnt[k] = ctr + 10 + k; L.res_base[k] = (uint32_t)k * jcap; }
	L.cig_cur = ctr + 8; L.cig_cap = cig_cap_words; L.overflow = ctr + 9;
	ctr[8] = cig_words;
	emu::launch((nt + 63) / 64, 64, [=](){ k_plan(dt, nt, dp, dr, R, A, L, dts); });
	if(ctr[9]) return -2;
	uint32_t n1[6]; for(int k = 0; k < 6; k++) n1[k] = (uint32_t)ctr[10 + k];
	run_lists(L, n1, nullptr, R, A.P, cgp, dres, ctr);
	emu::launch((nt + 63) / 64, 64, [=](){ k_plan2(dt, nt, dp, dr, dres, R, A, L, dts); });
	if(ctr[9]) return -2;
	uint32_t n2[6]; for(int k = 0; k < 6; k++) n2[k] = (uint32_t)ctr[10 + k];
	n2[4] = n1[4]; n2[5] = n1[5];
	run_lists(L, n2, n1, R, A.P, cgp, dres, ctr);
	for(int k = 0; k < 6; k++) stats[k] = (int)n2[k];
	unsigned long long need[2] = {0, 0}, ooff[2] = {0, 0};
	unsigned long long *dn = need;
	emu::launch((nt + 127) / 128, 128, [=](){ k_finish_size(nt, dres, dts, dn); });
	ooff[1] = need[0];
	std::vector<uint32_t> final_ops(need[0] + 16, 0u); zmo_record_t rec; memset(&rec, 0xCC, sizeof(rec));
	uint32_t *fo = final_ops.data(); zmo_record_t *drec = &rec; const unsigned long long *doo = ooff;
	(void)finish_warp;
	emu::launch((unsigned)(((unsigned long long)nt * 32 + 255) / 256), 256, [=](){ k_finish_warp(dt, nt, dr, dres, dj, cgp, A, dts, doo, fo, drec); });
	if(!rec.ok) return -1;
	const uint32_t *ops = fo;
	std::vector<uint32_t> ref_ops;
	if(refine){
		/* refine_records (zmo_align.cu) for one task */
		unsigned long long rows = 0, outw = 0; unsigned long long *prow = &rows, *pout = &outw;
		emu::launch(1, 128, [=](){ k_refine_size(drec, nt, prow, pout); });
		std::vector<int> bands(rows * 3 + 16, 0x7EEEEEEE); std::vector<RefJob> rj(3 * nt + 2);
		unsigned long long rc[4] = {0, 0, 0, 0}, row_off[2] = {0, rows}, out_off[2] = {0, outw}, scr_words[2] = {0, 0};
		int *db = bands.data(); RefJob *drj = rj.data(); unsigned long long *prc = rc, *dsw = scr_words; const unsigned long long *dro = row_off, *doo2 = out_off;
		emu::launch(1, 64, [=](){ k_refine_band(drec, nt, ops, dro, doo2, A.w, db, drj, prc, dsw); });
		unsigned long long scr_off[2] = {0, scr_words[0]}; const unsigned long long *dso = scr_off;
		std::vector<uint32_t> rarena(scr_words[0] + 64, 0xDEADBEEFu); ref_ops.assign(outw + 16, 0u);
		uint32_t *ra = rarena.data(), *ro = ref_ops.data();
		const DPPar P = A.P;
		ctr[0] = 0; ctr[3] = 0;
		if(rc[0]) emu::launch(1, 128, [=](){ k_refine_warp(drj, (uint32_t)prc[0], dso, dp, dt, R, P, db, ra, ro, drec, cp, 0, 2); });
		if(rc[1]) emu::launch(1, CL3_NT, [=](){ k_refine_cta(drj + nt, (uint32_t)prc[1], dso, dp, dt, R, P, db, ra, ro, drec, cp, 3, 2); });
		ctr[4] = 0;
		if(rc[2]) emu::launch(1, REFW_NT, [=](){ k_refine_wide(drj + 2 * nt, (uint32_t)prc[2], dso, dp, dt, R, P, db, ra, ro, drec, cp, 4, 2); });
		ops = ro;
	}
	out[0] = rec.score; out[1] = rec.tb; out[2] = rec.te; out[3] = rec.qb; out[4] = rec.qe; out[5] = rec.aln; out[6] = rec.mat; out[7] = rec.mis; out[8] = rec.ins; out[9] = rec.del;
	for(uint32_t k = 0; k < rec.n_cigar && (int)k < cigar_cap; k++) cigar_out[k] = ops[rec.cigar_off + k];
	return (int)rec.n_cigar;
}
