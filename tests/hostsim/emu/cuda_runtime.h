/*
 * TEST-ONLY stand-in for <cuda_runtime.h>: lets g++ compile the product's device headers (zmo_dp.cuh, zmo_dpr.cuh,
 * zmo_jobs.cuh ...) for the HOST and run a thread block as cooperative fibers, so that the CPU-only test-suite can
 * execute the very source the sm_100a kernels are built from (warp shuffles, reductions, barriers included) and
 * compare it with the oracle.  Put this directory first on the include path.  Never part of the product build.
 *
 * Model: emu::launch(grid, block, fn) runs the blocks one after another; the threads of a block are ucontext fibers on
 * one OS thread, scheduled round-robin.  A fiber runs until it reaches a collective (__shfl*_sync, __reduce_*_sync,
 * __ballot_sync, __syncwarp, __syncthreads), where it deposits its operand and yields until the warp / block is
 * complete.  Deterministic and race-free by construction, so it checks the ARITHMETIC and the collective structure of a
 * kernel, not its memory-ordering assumptions (a missing __syncwarp() goes unnoticed here; that is what the -m gpu
 * tests and compute-sanitizer are for).  `__shared__` becomes `static`: one block at a time, shared by its fibers.
 */
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ucontext.h>
#include <vector>
#include <functional>
#include <algorithm>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

/* ---- the few runtime types the shared headers mention (never called in the host simulation) ---- */
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
static inline const char *cudaGetErrorString(cudaError_t){ return "host simulation: no CUDA runtime"; }
static inline cudaError_t cudaMalloc(void **p, size_t n){ *p = malloc(n); return *p? cudaSuccess : cudaErrorEmu; }
static inline cudaError_t cudaFree(void *p){ free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n){ *p = malloc(n); return *p? cudaSuccess : cudaErrorEmu; }
static inline cudaError_t cudaFreeHost(void *p){ free(p); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t){ return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t){ return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t){ *ms = 0; return cudaSuccess; }

struct uint3_emu { unsigned x, y, z; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y){ uint2 v = {x, y}; return v; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w){ uint4 v = {x, y, z, w}; return v; }
static inline int2 make_int2(int x, int y){ int2 v = {x, y}; return v; }
static inline int4 make_int4(int x, int y, int z, int w){ int4 v = {x, y, z, w}; return v; }

namespace emu {

struct Warp { int expected = 0, count = 0; unsigned gen = 0; unsigned long long buf[2][32]; unsigned nth = 0; };
struct Fiber {
	ucontext_t ctx; void *stack = nullptr; bool done = false;
	uint3_emu tid, bid, bdim, gdim; int lane = 0, warp = 0; unsigned ncoll = 0;
};
struct Block {
	std::vector<Fiber> f; std::vector<Warp> w; ucontext_t sched; int cur = 0;
	int bar_expected = 0, bar_count = 0; unsigned bar_gen = 0;
	std::function<void()> body;
	std::vector<unsigned long long> dyn;      /* dynamic shared memory of the block (ZMO_DYN_SMEM) */
};
extern Block *g_blk;       /* defined by the one translation unit that includes this header */
static inline Fiber &cur(){ return g_blk->f[g_blk->cur]; }
static inline uint8_t *dyn_smem(){ return (uint8_t*)g_blk->dyn.data(); }
static inline void yield(){ Block *b = g_blk; swapcontext(&b->f[b->cur].ctx, &b->sched); }

/* warp-level rendezvous of all 32 lanes (every collective in the product uses the full mask) */
static inline void warp_wait(Warp &w){
	const unsigned g = w.gen;
	if(++w.count == w.expected){ w.count = 0; w.gen++; }
	else while(w.gen == g) yield();
}
/* deposit v, meet, return the generation's buffer (valid until this lane's collective after next) */
static inline const unsigned long long *exchange(unsigned long long v){
	Fiber &f = cur(); Warp &w = g_blk->w[f.warp];
	unsigned long long *b = w.buf[f.ncoll & 1]; f.ncoll++;
	b[f.lane] = v;
	warp_wait(w);
	return b;
}
static inline void block_wait(){
	Block *b = g_blk; const unsigned g = b->bar_gen;
	if(++b->bar_count == b->bar_expected){ b->bar_count = 0; b->bar_gen++; }
	else while(b->bar_gen == g) yield();
}
static void trampoline(){
	Block *b = g_blk;
	b->body();
	Fiber &f = b->f[b->cur];
	f.done = true;
	/* an exited thread no longer takes part in barriers (CUDA counts it as arrived) */
	b->bar_expected--;
	if(b->bar_expected > 0 && b->bar_count == b->bar_expected){ b->bar_count = 0; b->bar_gen++; }
	Warp &w = b->w[f.warp];
	w.expected--;
	if(w.expected > 0 && w.count == w.expected){ w.count = 0; w.gen++; }
	swapcontext(&f.ctx, &b->sched);
}

/* run `grid` blocks of `block` threads; kernel() is the __global__ function call with its arguments bound */
static inline void launch(unsigned grid, unsigned block, const std::function<void()> &kernel, size_t dyn_smem_bytes = 0, size_t stack_bytes = 512u << 10){
	Block B; B.body = kernel; B.dyn.assign(dyn_smem_bytes / 8 + 2, 0xA5A5A5A5A5A5A5A5ull);
	B.f.resize(block); B.w.resize((block + 31) / 32);
	for(unsigned t = 0; t < block; t++) B.f[t].stack = malloc(stack_bytes);
	for(unsigned bx = 0; bx < grid; bx++){
		g_blk = &B;
		B.bar_expected = (int)block; B.bar_count = 0; B.bar_gen = 0;
		for(size_t wi = 0; wi < B.w.size(); wi++){ B.w[wi].expected = (int)std::min<unsigned>(32u, block - 32u * (unsigned)wi); B.w[wi].count = 0; B.w[wi].gen = 0; }
		for(unsigned t = 0; t < block; t++){
			Fiber &f = B.f[t];
			f.done = false; f.ncoll = 0; f.lane = (int)(t & 31); f.warp = (int)(t >> 5);
			f.tid = {t, 0, 0}; f.bid = {bx, 0, 0}; f.bdim = {block, 1, 1}; f.gdim = {grid, 1, 1};
			getcontext(&f.ctx);
			f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = stack_bytes; f.ctx.uc_link = nullptr;
			makecontext(&f.ctx, (void (*)())trampoline, 0);
		}
		unsigned left = block; unsigned long long spins = 0;
		while(left){
			bool any = false;
			for(unsigned t = 0; t < block; t++){
				if(B.f[t].done) continue;
				B.cur = (int)t;
				swapcontext(&B.sched, &B.f[t].ctx);
				any = true;
				if(B.f[t].done) left--;
			}
			if(!any) break;
			if(++spins > (1ull << 40)){ fprintf(stderr, "emu: block %u does not terminate (dead-lock at a collective?)\n", bx); abort(); }
		}
	}
	for(unsigned t = 0; t < block; t++) free(B.f[t].stack);
	g_blk = nullptr;
}

}  // namespace emu

#define ZMO_DYN_SMEM(name) uint8_t *name = emu::dyn_smem()
#define threadIdx (emu::cur().tid)
#define blockIdx  (emu::cur().bid)
#define blockDim  (emu::cur().bdim)
#define gridDim   (emu::cur().gdim)

/* ---- warp collectives (full mask only: anything else aborts, the product never uses partial masks) ---- */
static inline void emu_full(unsigned mask){ if(mask != 0xffffffffu){ fprintf(stderr, "emu: partial-mask collective (0x%x) is not modelled\n", mask); abort(); } }
template<class T> static inline unsigned long long emu_bits(T v){ unsigned long long b = 0; static_assert(sizeof(T) <= 8, "operand too wide"); memcpy(&b, &v, sizeof(T)); return b; }
template<class T> static inline T emu_val(unsigned long long b){ T v; memcpy(&v, &b, sizeof(T)); return v; }
template<class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32){
	emu_full(mask); const int lane = emu::cur().lane; const unsigned long long *b = emu::exchange(emu_bits(v));
	const int s = (lane & ~(width - 1)) | (src & (width - 1));
	return emu_val<T>(b[s]);
}
template<class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32){
	emu_full(mask); const int lane = emu::cur().lane; const unsigned long long *b = emu::exchange(emu_bits(v));
	const int s = lane - (int)d;
	return (s >= (lane & ~(width - 1)))? emu_val<T>(b[s]) : v;
}
template<class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32){
	emu_full(mask); const int lane = emu::cur().lane; const unsigned long long *b = emu::exchange(emu_bits(v));
	const int s = lane + (int)d;
	return (s <= (lane | (width - 1)))? emu_val<T>(b[s]) : v;
}
template<class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32){
	emu_full(mask); const int lane = emu::cur().lane; const unsigned long long *b = emu::exchange(emu_bits(v));
	(void)width; return emu_val<T>(b[lane ^ x]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred){
	emu_full(mask); const unsigned long long *b = emu::exchange(pred? 1ull : 0ull);
	unsigned r = 0; for(int l = 0; l < 32; l++) if(b[l]) r |= 1u << l; return r;
}
static inline int __any_sync(unsigned mask, int pred){ return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred){ return __ballot_sync(mask, pred) == 0xffffffffu; }
static inline unsigned __activemask(){ return 0xffffffffu; }
static inline int __reduce_max_sync(unsigned mask, int v){ emu_full(mask); const unsigned long long *b = emu::exchange(emu_bits(v)); int r = emu_val<int>(b[0]); for(int l = 1; l < 32; l++) r = std::max(r, emu_val<int>(b[l])); return r; }
static inline int __reduce_min_sync(unsigned mask, int v){ emu_full(mask); const unsigned long long *b = emu::exchange(emu_bits(v)); int r = emu_val<int>(b[0]); for(int l = 1; l < 32; l++) r = std::min(r, emu_val<int>(b[l])); return r; }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v){ emu_full(mask); const unsigned long long *b = emu::exchange(emu_bits(v)); unsigned r = emu_val<unsigned>(b[0]); for(int l = 1; l < 32; l++) r = std::max(r, emu_val<unsigned>(b[l])); return r; }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v){ emu_full(mask); const unsigned long long *b = emu::exchange(emu_bits(v)); unsigned r = emu_val<unsigned>(b[0]); for(int l = 1; l < 32; l++) r = std::min(r, emu_val<unsigned>(b[l])); return r; }
static inline int __reduce_add_sync(unsigned mask, int v){ emu_full(mask); const unsigned long long *b = emu::exchange(emu_bits(v)); int r = 0; for(int l = 0; l < 32; l++) r += emu_val<int>(b[l]); return r; }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v){ emu_full(mask); const unsigned long long *b = emu::exchange(emu_bits(v)); unsigned r = 0; for(int l = 0; l < 32; l++) r += emu_val<unsigned>(b[l]); return r; }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v){ emu_full(mask); const unsigned long long *b = emu::exchange(emu_bits(v)); unsigned r = 0; for(int l = 0; l < 32; l++) r |= emu_val<unsigned>(b[l]); return r; }
static inline void __syncwarp(unsigned mask = 0xffffffffu){ emu_full(mask); emu::warp_wait(emu::g_blk->w[emu::cur().warp]); }
static inline void __syncthreads(){ emu::block_wait(); }
static inline void __threadfence(){}
static inline void __threadfence_block(){}

/* ---- scalar intrinsics ---- */
template<class T> static inline T __ldg(const T *p){ return *p; }
static inline int __popc(unsigned v){ return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v){ return __builtin_popcountll(v); }
static inline int __clz(int v){ return v? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v){ return v? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v){ return __builtin_ffs(v); }
static inline int __ffsll(long long v){ return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v){ unsigned r = 0; for(int i = 0; i < 32; i++) if(v & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned long long __brevll(unsigned long long v){ unsigned long long r = 0; for(int i = 0; i < 64; i++) if(v & (1ull << i)) r |= 1ull << (63 - i); return r; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh){ sh &= 31; return sh? (hi << sh) | (lo >> (32 - sh)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh){ sh &= 31; return sh? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s){
	const unsigned long long v = ((unsigned long long)y << 32) | x; unsigned r = 0;
	for(int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
	return r;
}
using std::min;
using std::max;
static inline int min(int a, unsigned b){ return (long long)a < (long long)b? a : (int)b; }

/* ---- atomics (fibers never pre-empt each other between collectives: plain read-modify-write) ---- */
template<class T> static inline T atomicAdd(T *p, T v){ T o = *p; *p = o + v; return o; }
template<class T> static inline T atomicMax(T *p, T v){ T o = *p; if(v > o) *p = v; return o; }
template<class T> static inline T atomicMin(T *p, T v){ T o = *p; if(v < o) *p = v; return o; }
template<class T> static inline T atomicOr(T *p, T v){ T o = *p; *p = o | v; return o; }
template<class T> static inline T atomicExch(T *p, T v){ T o = *p; *p = v; return o; }
template<class T> static inline T atomicCAS(T *p, T c, T v){ T o = *p; if(o == c) *p = v; return o; }
