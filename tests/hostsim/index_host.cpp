/*
 * TEST-ONLY host simulation of the k-mer index build and the candidate query (zmo_index_build + zmo_candidates): the product's
 * kernels (smartdenovo_b200/csrc/zmo_index_kernels.cuh) compiled against tests/hostsim/emu/cuda_runtime.h, with std:: sorts, scans and
 * reductions where zmo_index.cu calls CUB.  Never linked into the product.
 *
 * build: g++ -O1 -std=c++17 -Itests/hostsim/emu -fPIC -shared tests/hostsim/index_host.cpp
 */
#include "cuda_runtime.h"
#include <string>
#include <numeric>
thread_local std::string g_zmo_err;
int zmo_set_err(int code, const char *, ...){ return code; }
namespace emu { Block *g_blk = nullptr; }
#include "../../smartdenovo_b200/csrc/zmo_index_kernels.cuh"

template<class T> static void excl_scan(const T *in, T *out, size_t n){ T acc = 0; for(size_t i = 0; i < n; i++){ const T v = in[i]; out[i] = acc; acc += v; } }
template<class V> static void sort_pairs(std::vector<unsigned long long> &k, std::vector<V> &v, size_t n){       /* cub::DeviceRadixSort::SortPairs is stable */
	std::vector<size_t> ix(n); std::iota(ix.begin(), ix.end(), (size_t)0);
	std::stable_sort(ix.begin(), ix.end(), [&](size_t a, size_t b){ return k[a] < k[b]; });
	std::vector<unsigned long long> k2(n); std::vector<V> v2(n);
	for(size_t i = 0; i < n; i++){ k2[i] = k[ix[i]]; v2[i] = v[ix[i]]; }
	std::copy(k2.begin(), k2.end(), k.begin()); std::copy(v2.begin(), v2.end(), v.begin());
}

/*
 * reads: nreads sequences of 0..3 codes back to back; index over [beg, end); candidate events of the nq query reads qids[] in one
 * zmo_candidates-shaped call.  *kcut_io < 2: automatic K (written back).  idx_stats = {distinct k-mers, postings kept}.
 * ev_off (nq + 1) delimits each query's events in ev_out (tkey, ol pairs).  Returns the number of events.
 */
extern "C" int sim_index_candidates(const uint8_t *seqs, const int *lens, int nreads, int beg_, int end_, const int *qids_, int nq_, int ksize, int hk, int ksave, int kovl,
		uint32_t *kcut_io, unsigned long long *idx_stats, unsigned long long *ev_off, uint32_t *ev_out, int ev_cap){
	std::vector<uint32_t> words; std::vector<uint64_t> woff(nreads); std::vector<uint32_t> rlen(nreads);
	{ size_t o = 0; for(int r = 0; r < nreads; r++){ const int n = lens[r]; std::vector<uint32_t> w((n + 15) / 16 + 4, 0); for(int i = 0; i < n; i++) w[i >> 4] |= (uint32_t)(seqs[o + i] & 3) << (((~i) & 15) << 1);
		while(w.size() & 3) w.push_back(0); woff[r] = words.size(); rlen[r] = (uint32_t)n; words.insert(words.end(), w.begin(), w.end()); o += (size_t)n; } }
	DevReads R; R.words = words.data(); R.woff = woff.data(); R.len = rlen.data(); R.n = (uint32_t)nreads;
	const uint32_t beg = (uint32_t)beg_, end = std::min<uint32_t>((uint32_t)end_, (uint32_t)nreads), nr = end - beg; const int bs = 64;
	/* ---- zmo_index_build ---- */
	std::vector<unsigned long long> cnt(nr + 1, 0), off(nr + 2, 0);
	unsigned long long *d_cnt = cnt.data(), *d_off = off.data();
	emu::launch((nr + bs - 1) / bs, bs, [=](){ k_idx_count(R, beg, end, ksize, hk, (uint32_t)ksave, d_cnt); });
	cnt[nr] = 0; excl_scan(cnt.data(), off.data(), (size_t)nr + 1);
	const unsigned long long N = off[nr];
	std::vector<unsigned long long> ix_mer(N + 1), ix_off(2), run_start(N + 2); std::vector<uint8_t> ix_flt(8); std::vector<uint32_t> ix_post(4);
	unsigned long long ne = 0, npost = 0; uint32_t K = *kcut_io;
	if(N == 0){ if(K < 2) K = 100; }
	else {
		std::vector<unsigned long long> keys(N); std::vector<uint32_t> vals(N), hflag(N + 4), hpos(N + 4), rc(N + 4);
		unsigned long long *d_k = keys.data(); uint32_t *d_v = vals.data(), *d_hf = hflag.data(), *d_hp = hpos.data(), *d_rc = rc.data();
		emu::launch((nr + bs - 1) / bs, bs, [=](){ k_idx_fill(R, beg, end, ksize, hk, (uint32_t)ksave, d_off, d_k, d_v); });
		sort_pairs(keys, vals, (size_t)N);
		emu::launch((unsigned)((N + 255) / 256), 256, [=](){ k_idx_heads(d_k, N, d_hf); });
		excl_scan(hflag.data(), hpos.data(), (size_t)N);
		ne = (unsigned long long)hpos[N - 1] + hflag[N - 1];
		unsigned long long *d_mer = ix_mer.data(), *d_rs = run_start.data(); const unsigned long long NE = ne;
		emu::launch((unsigned)((N + 255) / 256), 256, [=](){ k_idx_runs(d_k, d_hf, d_hp, N, d_mer, d_rs); });
		emu::launch((unsigned)((ne + 255) / 256), 256, [=](){ k_idx_counts(d_rs, NE, N, d_rc); });
		unsigned long long ktot = 0; SatCount sat; for(unsigned long long i = 0; i < ne; i++) ktot += sat(rc[i]);
		const uint32_t kavg = (uint32_t)(ktot / (ne + 1));
		if(K < 2){ const uint32_t ka = kavg < 20? 20 : kavg; K = ka * 5; }
		std::vector<unsigned long long> kept(ne + 1, 0); ix_off.assign(ne + 2, 0); ix_flt.assign(ne + 8, 0);
		unsigned long long st2[2] = {0, 0}; unsigned long long *d_kept = kept.data(), *d_st = st2; uint8_t *d_flt = ix_flt.data(); const uint32_t KK = K;
		emu::launch((unsigned)((ne + 255) / 256), 256, [=](){ k_idx_flags(d_rc, NE, KK, d_flt, d_kept, d_st); });
		kept[ne] = 0; excl_scan(kept.data(), ix_off.data(), (size_t)ne + 1);
		npost = ix_off[ne];
		ix_post.assign(npost + 4, 0);
		const unsigned long long *d_ioff = ix_off.data(); uint32_t *d_post = ix_post.data();
		emu::launch((unsigned)((ne + 255) / 256), 256, [=](){ k_idx_gather(d_rs, d_ioff, d_flt, d_rc, NE, d_v, d_post); });
	}
	*kcut_io = K; idx_stats[0] = ne; idx_stats[1] = npost;
	/* ---- zmo_candidates ---- */
	const uint32_t nq = (uint32_t)nq_;
	std::vector<uint32_t> q(qids_, qids_ + nq);
	IdxView I; I.mer = ix_mer.data(); I.off = ix_off.data(); I.flt = ix_flt.data(); I.post = ix_post.data(); I.n = ne;
	std::vector<unsigned long long> nch(nq + 1, 0), qoff(nq + 2, 0);
	const uint32_t *d_q = q.data(); unsigned long long *d_nch = nch.data(), *d_qoff = qoff.data();
	emu::launch((nq + 127) / 128, 128, [=](){ k_q_nchunks(R, d_q, nq, d_nch); });
	nch[nq] = 0; excl_scan(nch.data(), qoff.data(), (size_t)nq + 1);
	const unsigned long long NC = qoff[nq];
	std::vector<unsigned long long> ccnt(NC + 1, 0), coff(NC + 2, 0);
	unsigned long long *d_ccnt = ccnt.data(), *d_coff = coff.data();
	emu::launch((unsigned)((NC + 127) / 128), 128, [=](){ k_qk_scan<0>(R, d_q, nq, d_qoff, NC, ksize, hk, (uint32_t)ksave, d_ccnt, nullptr, nullptr); });
	ccnt[NC] = 0; excl_scan(ccnt.data(), coff.data(), (size_t)NC + 1);
	const unsigned long long NK = coff[NC];
	unsigned long long NT = 0; uint32_t nev = 0;
	std::vector<unsigned long long> kmer(NK + 1), kinfo(NK + 1), kcnt(NK + 1, 0), koff(NK + 2, 0); std::vector<uint32_t> ent(NK + 2);
	unsigned long long *d_kmer = kmer.data(), *d_kinfo = kinfo.data(), *d_kcnt = kcnt.data(), *d_koff = koff.data(); uint32_t *d_ent = ent.data();
	if(NK){
		emu::launch((unsigned)((NC + 127) / 128), 128, [=](){ k_qk_scan<1>(R, d_q, nq, d_qoff, NC, ksize, hk, (uint32_t)ksave, d_coff, d_kmer, d_kinfo); });
		emu::launch((unsigned)((NK + 127) / 128), 128, [=](){ k_qk_lookup(I, R, d_q, d_kmer, d_kinfo, NK, d_ent, d_kcnt); });
		kcnt[NK] = 0; excl_scan(kcnt.data(), koff.data(), (size_t)NK + 1);
		NT = koff[NK];
	}
	for(uint32_t i = 0; i <= nq; i++) ev_off[i] = 0;
	if(NT){
		std::vector<unsigned long long> tk(NT), tv(NT); std::vector<uint32_t> flag(NT + 4), ol(NT + 4), pos(NT + 4);
		unsigned long long *d_tk = tk.data(), *d_tv = tv.data(); uint32_t *d_flag = flag.data(), *d_ol = ol.data(), *d_pos = pos.data();
		emu::launch((unsigned)((NK + 127) / 128), 128, [=](){ k_qk_expand(I, R, d_q, d_kinfo, d_ent, d_koff, NK, d_tk, d_tv); });
		sort_pairs(tk, tv, (size_t)NT);
		const unsigned long long NN = NT;
		emu::launch((unsigned)((NT + 255) / 256), 256, [=](){ k_cand_union(d_tk, d_tv, NN, (uint32_t)kovl, d_flag, d_ol); });
		excl_scan(flag.data(), pos.data(), (size_t)NT);
		nev = pos[NT - 1] + flag[NT - 1];
		std::vector<zmo_event_t> ev(nev + 1); std::vector<uint32_t> evq(nev + 1);
		zmo_event_t *d_ev = ev.data(); uint32_t *d_evq = evq.data(); const uint32_t NEV = nev;
		emu::launch((unsigned)((NT + 255) / 256), 256, [=](){ k_cand_emit(d_tk, d_flag, d_pos, d_ol, NN, d_ev, d_evq); });
		emu::launch((nq + 1 + 127) / 128, 128, [=](){ k_cand_offsets(d_evq, NEV, nq, ev_off); });
		for(uint32_t i = 0; i < nev && (int)i < ev_cap; i++){ ev_out[2 * i] = ev[i].tkey; ev_out[2 * i + 1] = ev[i].ol; }
	}
	return (int)nev;
}
