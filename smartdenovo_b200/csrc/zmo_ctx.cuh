/*
 * zmo_ctx.cuh -- context, device buffers and small host/device helpers shared by the .cu files.
 */
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <string>
#include <cuda_runtime.h>
#include "../../include/zmo_b200.h"

extern thread_local std::string g_zmo_err;
int zmo_set_err(int code, const char *fmt, ...);

#define CUDA_TRY(expr) do { cudaError_t _e = (expr); if(_e != cudaSuccess){ return zmo_set_err(ZMO_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); } } while(0)

/* wall time spent growing device / pinned buffers (process-wide; zmo_alloc_stats): what a cold start pays */
#include <chrono>
#include <atomic>
extern std::atomic<unsigned long long> g_zmo_alloc_ns[2], g_zmo_alloc_calls[2], g_zmo_alloc_bytes[2];
struct AllocTimer { int k; size_t b; std::chrono::steady_clock::time_point t0; AllocTimer(int k_, size_t b_) : k(k_), b(b_), t0(std::chrono::steady_clock::now()) {}
	~AllocTimer(){ g_zmo_alloc_ns[k] += (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); g_zmo_alloc_calls[k]++; g_zmo_alloc_bytes[k] += b; } };

/* growable device buffer */
struct DevBuf {
	void *p = nullptr; size_t cap = 0;
	int reserve(size_t bytes){
		if(bytes <= cap) return 0;
		size_t want = cap? cap : (1u << 20);
		while(want < bytes) want = want + want / 2 + (1u << 20);
		AllocTimer at(0, want);
		if(p) cudaFree(p);
		p = nullptr; cap = 0;
		cudaError_t e = cudaMalloc(&p, want);
		if(e != cudaSuccess){ p = nullptr; return zmo_set_err(ZMO_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
		cap = want; return 0;
	}
	void release(){ if(p) cudaFree(p); p = nullptr; cap = 0; }
	template<class T> T* as() const { return (T*)p; }
};

/* pinned host staging buffer */
struct PinBuf {
	void *p = nullptr; size_t cap = 0;
	int reserve(size_t bytes){
		if(bytes <= cap) return 0;
		size_t want = cap? cap : (1u << 20);
		while(want < bytes) want = want + want / 2 + (1u << 20);
		AllocTimer at(1, want);
		if(p) cudaFreeHost(p);
		p = nullptr; cap = 0;
		cudaError_t e = cudaMallocHost(&p, want);
		if(e != cudaSuccess){ p = nullptr; return zmo_set_err(ZMO_ERR_CUDA, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e)); }
		cap = want; return 0;
	}
	void release(){ if(p) cudaFreeHost(p); p = nullptr; cap = 0; }
	template<class T> T* as() const { return (T*)p; }
};

enum { ST_INDEX = 0, ST_CAND, ST_SEED, ST_WINALN, ST_GAP, ST_EXT, ST_DOT, ST_COPY, ST_DPWALL, ST_N = 12 };

/* one batch slot of the pair-seed stage (results stay on the device for zmo_pair_align) */
struct SeedSlot {
	uint32_t np = 0;
	DevBuf pairs;       /* zmo_pair_t[np] */
	DevBuf seeds;       /* zmo_pairseed_t[np] (device copy) */
	DevBuf wins;        /* DevWin[] kept windows, both strands */
	DevBuf anchors;     /* DevZPair[] anchors of kept windows */
	uint64_t n_wins = 0, n_anchors = 0;
	std::vector<zmo_pairseed_t> h_seeds;
	std::vector<int32_t> h_wspan;       /* per kept window: q span, c span, #anchors (host copy for job sizing) */
};

/* read store + k-mer index: built by the root context, shared read-only with its clones (zmo_ctx_clone) */
struct ZStore {
	uint32_t n_reads = 0; uint64_t n_bases = 0;
	DevBuf rd_words;     /* uint32 packed, per read 16-byte aligned */
	DevBuf rd_woff;      /* uint64 word offset per read */
	DevBuf rd_len;       /* uint32 */
	std::vector<uint32_t> h_rdlen; std::vector<uint64_t> h_woff;
	uint32_t max_rdlen = 0;
	/* k-mer index */
	bool have_index = false;
	uint64_t n_ent = 0, n_post = 0; uint32_t kcut = 0;
	DevBuf ix_mer;       /* uint64 distinct sampled k-mers, ascending */
	DevBuf ix_off;       /* uint64 posting offset (n_ent+1) */
	DevBuf ix_flt;       /* uint8 filtered flag */
	DevBuf ix_post;      /* uint32 postings (rd_id<<1|dir) */
};

struct zmo_ctx {
	int device = 0, n_sm = 0;
	zmo_params_t par;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	cudaStream_t aux[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};     /* concurrent DP executor classes */
	cudaEvent_t ev_fork = nullptr, ev_a0[6] = {nullptr}, ev_a1[6] = {nullptr};
	double stage_ms[ST_N] = {0};
	uint64_t launches = 0;
	uint64_t counters[8] = {0};
	ZStore *st = nullptr;  /* read store + k-mer index: own (root context) or the root's (clone) */
	ZStore own;
	bool is_clone = false;
	bool refine = false;      /* -n: kswx_refine_alignment after the stitch (zmo_set_refine) */
	/* scratch */
	DevBuf s0, s1, s2, s3, s4, s5, s6, s7, cubtmp;
	DevBuf zfilt;        /* per-query slot filter of the batch z-index (zmo_seed.cu) */
	DevBuf wb0, wb1;     /* bridge-level window alignment (zmo_winbridge.cuh): step records; offsets / sort keys / lists */
	DevBuf arena;        /* bump-allocated DP scratch (traceback, staged sequences) */
	DevBuf d_ctr;        /* device counters / cursors (uint64[64]) */
	PinBuf h0, h1, h2;
	SeedSlot slot[2];
};

struct StageTimer {
	zmo_ctx *c; int st;
	StageTimer(zmo_ctx *c_, int st_) : c(c_), st(st_) { cudaEventRecord(c->ev0, c->stream); }
	~StageTimer(){ float ms = 0; cudaEventRecord(c->ev1, c->stream); cudaEventSynchronize(c->ev1); cudaEventElapsedTime(&ms, c->ev0, c->ev1); c->stage_ms[st] += ms; }
};

/* device-side read accessors */
struct DevReads { const uint32_t *words; const uint64_t *woff; const uint32_t *len; uint32_t n; };
static inline DevReads dev_reads(const zmo_ctx *c){ DevReads r; r.words = c->st->rd_words.as<uint32_t>(); r.woff = c->st->rd_woff.as<uint64_t>(); r.len = c->st->rd_len.as<uint32_t>(); r.n = c->st->n_reads; return r; }

/* indices into the device counter block d_ctr (uint64 each) */
enum { CTR_CELLS_EXT = 0, CTR_CELLS_WIN, CTR_CELLS_GAP, CTR_ZPAIRS, CTR_POSTINGS, CTR_ARENA = 8, CTR_WORK = 9, CTR_OVERFLOW = 10, CTR_N1 = 11, CTR_N2 = 12, CTR_N3 = 13, CTR_N4 = 14, CTR_N5 = 15, CTR_CIG = 16, CTR_JOBS = 17 /* ..22 */, CTR_WORKK = 23 /* ..28 */, CTR_TOTAL = 32 };
