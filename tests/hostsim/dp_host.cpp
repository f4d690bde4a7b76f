/*
 * TEST-ONLY host simulation of the DP executor kernels (smartdenovo_b200/csrc/zmo_dp_kernels.cuh and zmo_winalign.cuh with
 * zmo_jobs.cuh, zmo_dpr.cuh, zmo_dp.cuh underneath): the kernels are compiled for the host against tests/hostsim/emu/cuda_runtime.h and
 * every thread block runs as cooperative fibers, so the CPU-only test-suite executes the source the sm_100a kernels
 * are built from -- register-resident sweeps, warp shuffles, reductions, barriers, traceback walk -- and compares it with
 * the oracle.  Never linked into libzmo_b200.so or wtzmo.
 *
 * build: g++ -O1 -std=c++17 -Itests/hostsim/emu -fPIC -shared tests/hostsim/dp_host.cpp
 */
#include "cuda_runtime.h"
#include <string>
thread_local std::string g_zmo_err;
int zmo_set_err(int code, const char *, ...){ return code; }
namespace emu { Block *g_blk = nullptr; }
#include "../../smartdenovo_b200/csrc/zmo_dp_kernels.cuh"
#include "../../smartdenovo_b200/csrc/zmo_winalign.cuh"
#include "../../smartdenovo_b200/csrc/zmo_winbridge.cuh"
#include "wb_pipeline.h"
#include "../../smartdenovo_b200/csrc/zmo_stitch_kernels.cuh"
#include "../../smartdenovo_b200/csrc/zmo_refine_kernels.cuh"

/* two reads in the device layout: 16 bases per uint32, MSB first, per read 16-byte aligned, spare words behind */
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

static void fill_out(const DPRes &r, std::vector<uint32_t> &cg, int *out10, uint32_t *cig, int cig_cap){
	out10[0] = r.score; out10[1] = 0; out10[2] = r.te; out10[3] = 0; out10[4] = r.qe; out10[5] = r.mat + r.mis + r.ins + r.del;
	out10[6] = r.mat; out10[7] = r.mis; out10[8] = r.ins; out10[9] = r.del;
	/* kernels emit the CIGAR in walk order (end -> start); flip it like zmo_dp.cu's dp_batch does */
	for(int k = 0; k < r.ncig && k < cig_cap; k++) cig[k] = cg[r.ncig - 1 - k];
}

/*
 * One extension problem (q = rows, t = columns; both given as 0..3 codes in logical order) through the kernel of executor
 * class `cls` (-1: the class ext_class() picks, as the product does; a larger class than needed is legal, a smaller one is
 * not).  copies > 1 queues the same job several times on a 2-CTA grid so that the work-counter loop of the persistent
 * executors runs as well.  Returns the number of CIGAR ops; out10 mirrors ref_shim's layout (conftest.call_ext).
 */
extern "C" int sim_dp_extend(int mode, int cls, int copies, const uint8_t *q, int qlen, const uint8_t *t, int tlen, int init, int Wp,
		int M, int X, int O, int E, int T, int *out10, uint32_t *cig, int cig_cap){
	DPPar P; P.M = M; P.X = X; P.I = O; P.D = O; P.E = E; P.T = T;
	SimReads rd(q, qlen, t, tlen); DevReads R = rd.dev();
	BandDims d; d.W = 0; d.ql = d.tl = d.ncol = 0;
	const int init0 = init < 0? 0 : init;
	if(qlen > 0 && tlen > 0) d = band_dims(qlen, tlen, init0, Wp, P);
	const int need = ext_class(d.ncol);
	if(cls < 0) cls = need;
	if(cls < need) return -1;
	if(copies < 1) copies = 1;
	std::vector<DPJob> jobs(copies); std::vector<DPRes> res(copies); unsigned long long scratch = 0, cigw = 0;
	for(int k = 0; k < copies; k++){
		DPJob J; memset(&J, 0, sizeof(J));
		J.q_rid = 0; J.t_rid = 1; J.q_start = 0; J.q_step = 1; J.q_comp = 0; J.qlen = qlen; J.t_start = 0; J.t_step = 1; J.t_comp = 0; J.tlen = tlen;
		J.init = init; J.Wp = Wp; J.out_idx = (uint32_t)k; J.cig_off = cigw; J.cig_cap = (uint32_t)(qlen + tlen + 4); cigw += J.cig_cap;
		J.scratch = scratch;
		/* the widest layout any class may use for this band, so that forcing a larger class stays inside the job's scratch */
		unsigned long long sw = 0; for(int c2 = need; c2 <= 3; c2++) sw = std::max(sw, ext_scratch_words_cls(d, c2));
		scratch += sw + 64;
		jobs[k] = J;
	}
	std::vector<uint32_t> arena(scratch + 64, 0xDEADBEEFu), cg(cigw + 16, 0u);
	unsigned long long ctr[4] = {0, 0, 0, 0};
	DPSlab SB; SB.base = 0; SB.off = nullptr;
	const uint32_t n = (uint32_t)copies;
	const DPJob *dj = jobs.data(); DPRes *dr = res.data(); uint32_t *ar = arena.data(), *cgp = cg.data(); unsigned long long *cp = ctr;
	const unsigned grid = copies > 1? 2u : 1u;
#define EXT_LAUNCH(KERNEL, NT) emu::launch(grid, NT, [=](){ KERNEL(dj, nullptr, n, R, P, ar, SB, cgp, dr, cp, 0, 1); })
	if(cls == 0){ if(mode) EXT_LAUNCH((k_ext_warp<1>), 32 * WRP_PER_CTA); else EXT_LAUNCH((k_ext_warp<0>), 32 * WRP_PER_CTA); }
	else if(cls == 1){ if(mode) EXT_LAUNCH((k_ext_cta<64, 7, 1>), 64); else EXT_LAUNCH((k_ext_cta<64, 7, 0>), 64); }
	else if(cls == 2){ if(mode) EXT_LAUNCH((k_ext_cta<128, 7, 1>), 128); else EXT_LAUNCH((k_ext_cta<128, 7, 0>), 128); }
	else { if(mode) EXT_LAUNCH((k_ext_cta<CL3_NT, CL3_C, 1>), CL3_NT); else EXT_LAUNCH((k_ext_cta<CL3_NT, CL3_C, 0>), CL3_NT); }
#undef EXT_LAUNCH
	for(int k = 1; k < copies; k++){
		if(memcmp(&res[k], &res[0], sizeof(DPRes))) return -2;
		if(memcmp(cg.data() + jobs[k].cig_off, cg.data(), sizeof(uint32_t) * (size_t)std::max(res[0].ncig, 0))) return -3;
	}
	fill_out(res[0], cg, out10, cig, cig_cap);
	out10[1] = (int)(ctr[1] / (unsigned long long)copies);      /* DP cells of one job, as the kernels count them */
	return res[0].ncig;
}

/* ksw_global2 with first band w (the kernels run the reference's `w < |qlen - tlen|` doubling; wmax > 0 also the score < 0 retry):
 * wide = 0: k_glb_warp, 1: k_glb_cta.  Returns the number of CIGAR ops, *score and *w_used. */
extern "C" int sim_dp_global(int wide, const uint8_t *q, int qlen, const uint8_t *t, int tlen, int M, int X, int O, int E, int w, int wmax,
		int *score, int *w_used, int *cnt4, uint32_t *cig, int cig_cap){
	DPPar P; P.M = M; P.X = X; P.I = O; P.D = O; P.E = E; P.T = 0;
	SimReads rd(q, qlen, t, tlen); DevReads R = rd.dev();
	DPJob J; memset(&J, 0, sizeof(J));
	J.q_rid = 0; J.t_rid = 1; J.q_step = 1; J.t_step = 1; J.qlen = qlen; J.tlen = tlen; J.Wp = w; J.Wmax = wmax; J.cig_cap = (uint32_t)(qlen + tlen + 4);
	const unsigned long long sw = wide? glb_scratch_words<EXT_NT, EXT_C>(qlen, tlen, EXT_CAP) : glb_scratch_words<32, WRP_C>(qlen, tlen, WRP_CAP);
	std::vector<uint32_t> arena(sw + 64, 0xDEADBEEFu), cg(J.cig_cap + 16, 0u);
	DPRes res; memset(&res, 0, sizeof(res));
	unsigned long long ctr[4] = {0, 0, 0, 0};
	DPSlab SB; SB.base = 0; SB.off = nullptr;
	const DPJob *dj = &J; DPRes *dr = &res; uint32_t *ar = arena.data(), *cgp = cg.data(); unsigned long long *cp = ctr;
	if(wide) emu::launch(1, EXT_NT, [=](){ k_glb_cta(dj, nullptr, 1u, R, P, ar, SB, cgp, dr, cp, 0, 1); });
	else emu::launch(1, 32 * WRP_PER_CTA, [=](){ k_glb_warp(dj, nullptr, 1u, R, P, ar, SB, cgp, dr, cp, 0, 1); });
	*score = res.score; *w_used = res.w_used;
	cnt4[0] = res.mat; cnt4[1] = res.mis; cnt4[2] = res.ins; cnt4[3] = res.del;
	for(int k = 0; k < res.ncig && k < cig_cap; k++) cig[k] = cg[res.ncig - 1 - k];
	return res.ncig;
}

/*
 * n_win windows of one (q, c, strand) task through k_window_align: q = pb1 forward, c given FORWARD (the kernel addresses the
 * reverse complement itself for dir = 1, view_pb2).  win = n_win x {q span, c span, n_anchors} (what the pipeline keeps in
 * SeedSlot::h_wspan), anc = the windows' anchors back to back, 6 ints each {off1, off2, len1, len2, dir1, dir2}.  Slab and CIGAR
 * region sizes follow pair_align_impl (zmo_align.cu).  out = n_win x 11 {score, tb, te, qb, qe, aln, mat, mis, ins, del, kept};
 * cig_out = the windows' CIGARs back to back, cig_n[i] ops each.
 */
extern "C" int sim_window_align(const uint8_t *q, int qlen, const uint8_t *c, int clen, int dir, const int *win, int n_win, const int *anc,
		int w, int M, int X, int O, int E, int T, int zovl, float min_id, int *out, uint32_t *cig_out, int cig_cap, int *cig_n){
	SimReads rd(q, qlen, c, clen); DevReads R = rd.dev();
	AlnPar A; A.w = w; A.ew = 800; A.W = 3200; A.zovl = zovl; A.min_id = min_id; A.P.M = M; A.P.X = X; A.P.I = O; A.P.D = O; A.P.E = E; A.P.T = T;
	std::vector<WItem> items(n_win); std::vector<DevWin> wins(n_win); std::vector<DevZPair> an; std::vector<unsigned long long> icig(n_win);
	AlnTask task; task.pair_idx = 0; task.dir = (uint32_t)dir; task.item_off = 0; task.n_item = (uint32_t)n_win;
	zmo_pair_t pair; pair.qid = 0; pair.cid = 1;
	unsigned long long cig_words = 0; int max_rows = 16, a0 = 0;
	for(int i = 0; i < n_win; i++){
		const int s0 = win[3 * i], s1 = win[3 * i + 1], na = win[3 * i + 2];
		items[i].task = 0; items[i].win = (uint32_t)i;
		DevWin W; memset(&W, 0, sizeof(W)); W.anc0 = (uint32_t)a0; W.anc1 = (uint32_t)(a0 + na); W.dir = (uint8_t)dir; wins[i] = W;
		for(int k = 0; k < na; k++){
			const int *o = anc + 6 * (a0 + k); DevZPair p; memset(&p, 0, sizeof(p));
			p.off1 = (uint32_t)o[0]; p.off2 = (uint32_t)o[1]; p.len1 = (uint16_t)o[2]; p.len2 = (uint16_t)o[3]; p.dir1 = (uint8_t)o[4]; p.dir2 = (uint8_t)o[5]; an.push_back(p);
		}
		a0 += na;
		icig[i] = cig_words; cig_words += (unsigned long long)(s0 + s1 + 16 + 2 * na);
		if(s1 + 8 > max_rows) max_rows = s1 + 8;
		if(s0 + 8 > max_rows) max_rows = s0 + 8;
	}
	const int wgrid = (n_win + WA_WARPS - 1) / WA_WARPS + 1;
	const int wcol = std::min(max_rows + w, 2 * w + 1);
	unsigned long long slab = (unsigned long long)max_rows * band_row_words<32, WA_C>(wcol) + max_rows + (2ull * max_rows + 2ull * w + 16) + ((unsigned long long)max_rows >> 3) + (w >> 3) + 8;
	if(2 * w + 3 > WA_CAP){ unsigned long long cap = 1; while(cap < (unsigned long long)(2 * w + 3)) cap <<= 1; slab += 3 * cap; }
	slab = (slab + 63) & ~63ull;
	std::vector<uint32_t> arena(slab * (unsigned long long)wgrid * WA_WARPS + 64, 0xDEADBEEFu), cg(cig_words + 64, 0u);
	std::vector<DevReg> regs(n_win);
	unsigned long long ctr[4] = {0, 0, 0, 0};
	const WItem *di = items.data(); const AlnTask *dt = &task; const zmo_pair_t *dp = &pair; const DevWin *dw = wins.data(); const DevZPair *da = an.data();
	uint32_t *ar = arena.data(), *cgp = cg.data(); const unsigned long long *dic = icig.data(); DevReg *dr = regs.data(); unsigned long long *cp = ctr;
	const uint32_t nitems = (uint32_t)n_win;
	emu::launch((unsigned)wgrid, 32 * WA_WARPS, [=](){ k_window_align(di, nitems, dt, dp, dw, da, R, A, ar, slab, max_rows, cgp, dic, dr, cp, 0, 1, nullptr, nullptr); });
	int total = 0;
	for(int i = 0; i < n_win; i++){
		const DevReg &r = regs[i]; int *o = out + 11 * i;
		o[0] = r.score; o[1] = r.tb; o[2] = r.te; o[3] = r.qb; o[4] = r.qe; o[5] = r.aln; o[6] = r.mat; o[7] = r.mis; o[8] = r.ins; o[9] = r.del; o[10] = (int)r.kept;
		cig_n[i] = (int)r.cig_len;
		if(r.cig_off != icig[i] || r.cig_len > (unsigned long long)(win[3 * i] + win[3 * i + 1] + 16 + 2 * win[3 * i + 2])) return -1;     /* CIGAR region overrun */
		for(uint32_t k = 0; k < r.cig_len && total < cig_cap; k++) cig_out[total++] = cg[r.cig_off + k];
	}
	return total;
}

/*
 * The same windows through the bridge-level pipeline (zmo_winbridge.cuh: k_wb_prep -> scan + sort (CUB in the product, std here) -> k_wb_sweep ->
 * k_wb_ends -> k_wb_walk -> k_wb_stitch, then k_window_align on the windows the pipeline left out), sized like pair_align_impl (zmo_align.cu).
 * copies > 1 repeats the window list so that the sweep runs several rounds of 32 bridges per warp; copies = 0 runs one window per pass.  *n_fallback =
 * windows left to k_window_align.  The orchestration between the kernels is tests/hostsim/wb_pipeline.h.
 */
extern "C" int sim_window_align_bridge(const uint8_t *q, int qlen, const uint8_t *c, int clen, int dir, const int *win, int n_win0, const int *anc,
		int w, int M, int X, int O, int E, int T, int zovl, float min_id, int copies, int acap, int *out, uint32_t *cig_out, int cig_cap, int *cig_n, int *n_fallback){
	SimReads rd(q, qlen, c, clen); DevReads R = rd.dev();
	AlnPar A; A.w = w; A.ew = 800; A.W = 3200; A.zovl = zovl; A.min_id = min_id; A.P.M = M; A.P.X = X; A.P.I = O; A.P.D = O; A.P.E = E; A.P.T = T;
	if(w < 1 || w > WB_MAX_W) return -2;
	unsigned long long budget = 1ull << 40;      /* copies = 0: one window per pass (the product sweeps a wave whose scratch bound exceeds its budget in several passes) */
	if(copies <= 0){ budget = 1; copies = 1; }
	const int n_win = n_win0 * copies;
	std::vector<WItem> items(n_win); std::vector<DevWin> wins(n_win0); std::vector<DevZPair> an; std::vector<unsigned long long> icig(n_win);
	AlnTask task; task.pair_idx = 0; task.dir = (uint32_t)dir; task.item_off = 0; task.n_item = (uint32_t)n_win;
	zmo_pair_t pair; pair.qid = 0; pair.cid = 1;
	unsigned long long cig_words = 0; int a0 = 0;
	for(int i = 0; i < n_win0; i++){
		const int na = win[3 * i + 2];
		DevWin W; memset(&W, 0, sizeof(W)); W.anc0 = (uint32_t)a0; W.anc1 = (uint32_t)(a0 + na); W.dir = (uint8_t)dir; wins[i] = W;
		for(int k = 0; k < na; k++){
			const int *o = anc + 6 * (a0 + k); DevZPair p; memset(&p, 0, sizeof(p));
			p.off1 = (uint32_t)o[0]; p.off2 = (uint32_t)o[1]; p.len1 = (uint16_t)o[2]; p.len2 = (uint16_t)o[3]; p.dir1 = (uint8_t)o[4]; p.dir2 = (uint8_t)o[5]; an.push_back(p);
		}
		a0 += na;
	}
	an.push_back(DevZPair());
	for(int i = 0; i < n_win; i++){
		const int s0 = win[3 * (i % n_win0)], s1 = win[3 * (i % n_win0) + 1], na = win[3 * (i % n_win0) + 2];
		items[i].task = 0; items[i].win = (uint32_t)(i % n_win0);
		icig[i] = cig_words; cig_words += (unsigned long long)(s0 + s1 + 16 + 2 * na);
	}
	std::vector<uint32_t> cg(cig_words + 64, 0u);
	std::vector<DevReg> regs(n_win);
	const long nfb = sim_wb_pipeline(items.data(), (uint32_t)n_win, &task, &pair, wins.data(), an.data(), win, R, A, acap, icig.data(), cg.data(), regs.data(), budget, nullptr);
	if(nfb < 0) return (int)nfb;
	*n_fallback = (int)nfb;
	int total = 0;
	for(int i = 0; i < n_win; i++){
		const DevReg &r = regs[i]; const int b = i % n_win0;
		if(r.cig_off != icig[i] || r.cig_len > (unsigned long long)(win[3 * b] + win[3 * b + 1] + 16 + 2 * win[3 * b + 2])) return -1;     /* CIGAR region overrun */
		if(i >= n_win0){
			/* a repeated window must reproduce the first copy */
			const DevReg &r0 = regs[b];
			if(r.score != r0.score || r.tb != r0.tb || r.te != r0.te || r.qb != r0.qb || r.qe != r0.qe || r.aln != r0.aln || r.mat != r0.mat || r.mis != r0.mis || r.ins != r0.ins || r.del != r0.del || r.cig_len != r0.cig_len || r.kept != r0.kept) return -5;
			for(uint32_t k = 0; k < r.cig_len; k++) if(cg[r.cig_off + k] != cg[r0.cig_off + k]) return -6;
			continue;
		}
		int *o = out + 11 * i;
		o[0] = r.score; o[1] = r.tb; o[2] = r.te; o[3] = r.qb; o[4] = r.qe; o[5] = r.aln; o[6] = r.mat; o[7] = r.mis; o[8] = r.ins; o[9] = r.del; o[10] = (int)r.kept;
		cig_n[i] = (int)r.cig_len;
		for(uint32_t k = 0; k < r.cig_len && total < cig_cap; k++) cig_out[total++] = cg[r.cig_off + k];
	}
	return total;
}

/*
 * -n refinement of one stitched alignment through k_refine_size / k_refine_band / k_refine_warp|k_refine_cta: q = pb1 forward, c given
 * FORWARD with strand dir, the alignment by its start (tb on q, qb on c's strand), ends and CIGAR (alignment order).  The exclusive
 * scans the product does with CUB between the kernels are trivial for one task.  out = score, tb, te, qb, qe, aln, mat, mis, ins, del;
 * *cls_out = executor class (0 warp, 1 CTA, 2 wide-band fallback); returns the number of new CIGAR ops, -3 if the record came back not ok.
 */
extern "C" int sim_refine(const uint8_t *q, int qlen, const uint8_t *c, int clen, int dir, int tb, int te, int qb, int qe, const uint32_t *cigar_in, int n_in,
		int W, int M, int X, int O, int E, int *out, uint32_t *cigar_out, int cigar_cap, int *cls_out){
	SimReads rd(q, qlen, c, clen); DevReads R = rd.dev();
	DPPar P; P.M = M; P.X = X; P.I = O; P.D = O; P.E = E; P.T = -50;
	zmo_record_t rec; memset(&rec, 0, sizeof(rec));
	rec.ok = 1; rec.tb = tb; rec.te = te; rec.qb = qb; rec.qe = qe; rec.cigar_off = 0; rec.n_cigar = (uint32_t)n_in;
	AlnTask task; task.pair_idx = 0; task.dir = (uint32_t)dir; task.item_off = 0; task.n_item = 0;
	zmo_pair_t pair; pair.qid = 0; pair.cid = 1;
	std::vector<uint32_t> ops(cigar_in, cigar_in + n_in); ops.push_back(0);
	unsigned long long rows = 0, outw = 0;
	zmo_record_t *dr = &rec; unsigned long long *prow = &rows, *pout = &outw;
	emu::launch(1, 64, [=](){ k_refine_size(dr, 1u, prow, pout); });
	std::vector<int> bands(rows * 3 + 16, 0x7EEEEEEE); std::vector<RefJob> jobs(3 * 1 + 2);
	unsigned long long ctr[8] = {0, 0, 0, 0, 0, 0, 0, 0};      /* [0] work warp, [1] work cta, [2] cells, [3] work wide, [4..6] njobs per class */
	unsigned long long row_off[2] = {0, rows}, out_off[2] = {0, outw}, scr_words[2] = {0, 0};
	const uint32_t *dops = ops.data(); int *db = bands.data(); RefJob *dj = jobs.data(); unsigned long long *cp = ctr;
	const unsigned long long *dro = row_off, *doo = out_off; unsigned long long *dsw = scr_words;
	emu::launch(1, 64, [=](){ k_refine_band(dr, 1u, dops, dro, doo, W, db, dj, cp + 4, dsw); });
	unsigned long long scr_off[2] = {0, scr_words[0]};
	std::vector<uint32_t> arena(scr_words[0] + 64, 0xDEADBEEFu), out_ops(outw + 16, 0u);
	const unsigned long long *dso = scr_off; uint32_t *ar = arena.data(), *oo = out_ops.data(); const AlnTask *dt = &task; const zmo_pair_t *dp = &pair;
	*cls_out = ctr[6]? 2 : (ctr[5]? 1 : 0);
	if(ctr[4]) emu::launch(1, 128, [=](){ k_refine_warp(dj, (uint32_t)cp[4], dso, dp, dt, R, P, db, ar, oo, dr, cp, 0, 2); });
	if(ctr[5]) emu::launch(1, CL3_NT, [=](){ k_refine_cta(dj + 1, (uint32_t)cp[5], dso, dp, dt, R, P, db, ar, oo, dr, cp, 1, 2); });
	if(ctr[6]) emu::launch(1, REFW_NT, [=](){ k_refine_wide(dj + 2, (uint32_t)cp[6], dso, dp, dt, R, P, db, ar, oo, dr, cp, 3, 2); });
	if(!rec.ok) return -3;
	out[0] = rec.score; out[1] = rec.tb; out[2] = rec.te; out[3] = rec.qb; out[4] = rec.qe; out[5] = rec.aln; out[6] = rec.mat; out[7] = rec.mis; out[8] = rec.ins; out[9] = rec.del;
	for(uint32_t k = 0; k < rec.n_cigar && (int)k < cigar_cap; k++) cigar_out[k] = out_ops[rec.cigar_off + k];
	return (int)rec.n_cigar;
}

/*
 * k_finish_warp on caller-built stitch inputs.  Flat arrays: per task item_off, n_item and
 * ts = {ok, first, left_job, right_job, score, tb, te, qb, qe, aln, mat, mis, ins, del}; per region kept, cig_off, cig_len; per job
 * cig_off and jv = {score, qe, te, mat, mis, ins, del, ncig}.  recs_out = nt x {ok, score, tb, te, qb, qe, aln, mat, mis, ins, del, n_cigar}.
 */
extern "C" int sim_finish(int warp, int nt, const int *item_off, const int *n_item, const int *tsv, int nreg, const int *reg_kept, const int *reg_cig_off,
		const int *reg_cig_len, int njob, const int *job_cig_off, const int *jv, const uint32_t *cig_arena, const int *out_off, uint32_t *out_cig, int *recs_out){
	std::vector<AlnTask> tasks(nt); std::vector<TaskState> ts(nt); std::vector<DevReg> regs(nreg + 1); std::vector<DPRes> res(njob + 1); std::vector<DPJob> jobs(njob + 1);
	std::vector<unsigned long long> ooff(nt + 1); std::vector<zmo_record_t> recs(nt);
	for(int t = 0; t < nt; t++){
		tasks[t].pair_idx = 0; tasks[t].dir = 0; tasks[t].item_off = (uint32_t)item_off[t]; tasks[t].n_item = (uint32_t)n_item[t];
		const int *v = tsv + 14 * t; TaskState S; memset(&S, 0, sizeof(S));
		S.ok = v[0]; S.first = v[1]; S.last = -1; S.left_job = v[2]; S.right_job = v[3]; S.score = v[4]; S.tb = v[5]; S.te = v[6]; S.qb = v[7]; S.qe = v[8];
		S.aln = v[9]; S.mat = v[10]; S.mis = v[11]; S.ins = v[12]; S.del = v[13]; ts[t] = S;
		ooff[t] = (unsigned long long)out_off[t];
		memset(&recs[t], 0xCC, sizeof(zmo_record_t));
	}
	for(int i = 0; i < nreg; i++){ DevReg r; memset(&r, 0, sizeof(r)); r.kept = (uint32_t)reg_kept[i]; r.cig_off = (unsigned long long)reg_cig_off[i]; r.cig_len = (uint32_t)reg_cig_len[i]; regs[i] = r; }
	for(int j = 0; j < njob; j++){
		DPJob J; memset(&J, 0, sizeof(J)); J.cig_off = (unsigned long long)job_cig_off[j]; jobs[j] = J;
		const int *v = jv + 8 * j; DPRes r; memset(&r, 0, sizeof(r)); r.score = v[0]; r.qe = v[1]; r.te = v[2]; r.mat = v[3]; r.mis = v[4]; r.ins = v[5]; r.del = v[6]; r.ncig = v[7]; res[j] = r;
	}
	AlnPar A; memset(&A, 0, sizeof(A));
	const AlnTask *dt = tasks.data(); const DevReg *dr = regs.data(); const DPRes *ds = res.data(); const DPJob *dj = jobs.data(); const TaskState *dts = ts.data();
	const unsigned long long *doo = ooff.data(); zmo_record_t *drec = recs.data(); const uint32_t n = (uint32_t)nt;
	(void)warp;
	emu::launch((unsigned)(((unsigned long long)n * 32 + 255) / 256), 256, [=](){ k_finish_warp(dt, n, dr, ds, dj, cig_arena, A, dts, doo, out_cig, drec); });
	for(int t = 0; t < nt; t++){
		const zmo_record_t &r = recs[t]; int *o = recs_out + 12 * t;
		o[0] = r.ok; o[1] = r.score; o[2] = r.tb; o[3] = r.te; o[4] = r.qb; o[5] = r.qe; o[6] = r.aln; o[7] = r.mat; o[8] = r.mis; o[9] = r.ins; o[10] = r.del; o[11] = (int)r.n_cigar;
		if(r.ok && r.cigar_off != ooff[t]) return -1;
	}
	return 0;
}
