/*
 * zmo_api.cu -- context management, error reporting, read upload/re-pack.
 */
#include <stdarg.h>
#include "zmo_ctx.cuh"

thread_local std::string g_zmo_err;
std::atomic<unsigned long long> g_zmo_alloc_ns[2], g_zmo_alloc_calls[2], g_zmo_alloc_bytes[2];
/* out[0..2] = seconds, calls, bytes of device-buffer growth (cudaFree + cudaMalloc); out[3..5] = the same for page-locked host buffers */
extern "C" void zmo_alloc_stats(double out[6]){ for(int k = 0; k < 2; k++){ out[3 * k] = 1e-9 * (double)g_zmo_alloc_ns[k].load(); out[3 * k + 1] = (double)g_zmo_alloc_calls[k].load(); out[3 * k + 2] = (double)g_zmo_alloc_bytes[k].load(); } }
int zmo_set_err(int code, const char *fmt, ...){
	char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
	g_zmo_err = buf; return code;
}
extern "C" const char *zmo_last_error(void){ return g_zmo_err.c_str(); }

int zmo_seed_init_device(void);     /* zmo_seed.cu */

static int ctx_init(zmo_ctx *c, int device, const zmo_params_t *par){
	CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	if(prop.major < 10) return zmo_set_err(ZMO_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
	c->device = device; c->n_sm = prop.multiProcessorCount; c->par = *par;
	CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaEventCreate(&c->ev0)); CUDA_TRY(cudaEventCreate(&c->ev1)); CUDA_TRY(cudaEventCreate(&c->ev_fork));
	for(int k = 0; k < 6; k++){ CUDA_TRY(cudaStreamCreateWithFlags(&c->aux[k], cudaStreamNonBlocking)); CUDA_TRY(cudaEventCreate(&c->ev_a0[k])); CUDA_TRY(cudaEventCreate(&c->ev_a1[k])); }
	if(c->d_ctr.reserve(CTR_TOTAL * 8)) return ZMO_ERR_CUDA;
	CUDA_TRY(cudaMemsetAsync(c->d_ctr.p, 0, CTR_TOTAL * 8, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if(int rc = zmo_seed_init_device()) return rc;
	return ZMO_OK;
}

extern "C" int zmo_ctx_create(zmo_ctx **out, int device, const zmo_params_t *par){
	if(!out || !par) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if(e != cudaSuccess || ndev == 0) return zmo_set_err(ZMO_ERR_CUDA, "no CUDA device available (%s); libzmo_b200 has no CPU fallback", e == cudaSuccess? "0 devices" : cudaGetErrorString(e));
	if(device < 0 || device >= ndev) return zmo_set_err(ZMO_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
	CUDA_TRY(cudaSetDevice(device));
	if(par->ksize < 5 || par->ksize > 32 || par->zsize < 5 || par->zsize > 16 || par->ksave < 1 || par->E >= 0)
		return zmo_set_err(ZMO_ERR_ARG, "parameter out of range (k 5..32, z 5..16, S>=1, E<0)");
	zmo_ctx *c = new zmo_ctx();
	c->st = &c->own;
	if(int rc = ctx_init(c, device, par)){ zmo_ctx_destroy(c); return rc; }
	*out = c;
	return ZMO_OK;
}

extern "C" int zmo_ctx_clone(zmo_ctx *root, zmo_ctx **out){
	if(!root || !out) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	*out = nullptr;
	if(root->is_clone) return zmo_set_err(ZMO_ERR_ARG, "clone of a clone");
	CUDA_TRY(cudaSetDevice(root->device));
	zmo_ctx *c = new zmo_ctx();
	c->st = root->st; c->is_clone = true; c->refine = root->refine;
	if(int rc = ctx_init(c, root->device, &root->par)){ zmo_ctx_destroy(c); return rc; }
	*out = c;
	return ZMO_OK;
}

extern "C" int zmo_set_refine(zmo_ctx *c, int on){ if(!c) return zmo_set_err(ZMO_ERR_ARG, "null context"); c->refine = on != 0; return ZMO_OK; }

extern "C" void zmo_ctx_destroy(zmo_ctx *c){
	if(!c) return;
	cudaSetDevice(c->device);
	if(c->stream) cudaStreamSynchronize(c->stream);
	if(!c->is_clone){ DevBuf *sb[] = { &c->own.rd_words, &c->own.rd_woff, &c->own.rd_len, &c->own.ix_mer, &c->own.ix_off, &c->own.ix_flt, &c->own.ix_post }; for(DevBuf *b : sb) b->release(); }
	DevBuf *bufs[] = { &c->s0, &c->s1, &c->s2, &c->s3, &c->s4, &c->s5, &c->s6, &c->s7, &c->cubtmp, &c->arena, &c->d_ctr, &c->zfilt, &c->wb0, &c->wb1 };
	for(DevBuf *b : bufs) b->release();
	for(int s = 0; s < 2; s++){ c->slot[s].pairs.release(); c->slot[s].seeds.release(); c->slot[s].wins.release(); c->slot[s].anchors.release(); }
	c->h0.release(); c->h1.release(); c->h2.release();
	for(int k = 0; k < 6; k++){ if(c->aux[k]) cudaStreamDestroy(c->aux[k]); if(c->ev_a0[k]) cudaEventDestroy(c->ev_a0[k]); if(c->ev_a1[k]) cudaEventDestroy(c->ev_a1[k]); }
	if(c->ev0) cudaEventDestroy(c->ev0); if(c->ev1) cudaEventDestroy(c->ev1); if(c->ev_fork) cudaEventDestroy(c->ev_fork); if(c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

extern "C" void *zmo_host_alloc(size_t bytes){ void *p = nullptr; AllocTimer at(1, bytes); if(cudaHostAlloc(&p, bytes? bytes : 1, cudaHostAllocDefault) != cudaSuccess){ zmo_set_err(ZMO_ERR_CUDA, "cudaHostAlloc(%zu) failed", bytes); return nullptr; } return p; }
extern "C" void zmo_host_free(void *p){ if(p) cudaFreeHost(p); }

extern "C" uint64_t zmo_kernel_launches(const zmo_ctx *c){ return c? c->launches : 0; }
extern "C" void zmo_stage_ms(const zmo_ctx *c, double out[12]){ for(int i = 0; i < 12; i++) out[i] = c? c->stage_ms[i] : 0; }
extern "C" void zmo_counters(const zmo_ctx *c, uint64_t out[8]){
	for(int i = 0; i < 8; i++) out[i] = 0;
	if(!c) return;
	unsigned long long h[8];
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	if(cudaMemcpy(h, c->d_ctr.p, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess){ for(int i = 0; i < 3; i++) out[i] = h[i]; }
	out[3] = c->counters[3]; out[4] = c->counters[4]; out[5] = c->counters[5]; out[6] = c->counters[6];
}

/* re-pack the reference BaseBank layout (dna.h:78,263: 32 bases per uint64, MSB first, reads
 * concatenated at arbitrary base offsets) into per-read word-aligned uint32 words (16 bases, MSB
 * first).  One thread per output word. */
__global__ void k_repack(const unsigned long long *bank, const unsigned long long *rdoff, const uint32_t *rdlen, const unsigned long long *woff, uint32_t n_reads, uint32_t *out, unsigned long long total_words){
	unsigned long long gw = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if(gw >= total_words) return;
	/* binary search the read owning output word gw */
	uint32_t lo = 0, hi = n_reads;
	while(lo + 1 < hi){ uint32_t mid = (lo + hi) >> 1; if(woff[mid] <= gw) lo = mid; else hi = mid; }
	const uint32_t rid = lo; const unsigned long long w = gw - woff[rid];
	const uint32_t len = rdlen[rid]; const unsigned long long base0 = w * 16;
	uint32_t v = 0;
	if(base0 < len){
		int m = (int)((len - base0) < 16? (len - base0) : 16);
		for(int b = 0; b < m; b++){
			unsigned long long off = rdoff[rid] + base0 + b;
			uint32_t x = (uint32_t)((bank[off >> 5] >> (((~off) & 31ULL) << 1)) & 3ULL);
			v |= x << ((15 - b) << 1);
		}
	}
	out[gw] = v;
}

extern "C" int zmo_reads_upload(zmo_ctx *c, const uint64_t *bank, uint64_t n_bases, const uint64_t *rdoff, const uint32_t *rdlen, uint32_t n_reads){
	if(!c || !bank || !rdoff || !rdlen || n_reads == 0) return zmo_set_err(ZMO_ERR_ARG, "null/empty argument");
	if(c->is_clone) return zmo_set_err(ZMO_ERR_STATE, "reads are uploaded through the root context");
	CUDA_TRY(cudaSetDevice(c->device));
	c->st->n_reads = n_reads; c->st->n_bases = n_bases; c->st->have_index = false;
	c->st->h_rdlen.assign(rdlen, rdlen + n_reads); c->st->h_woff.resize((size_t)n_reads + 1);
	uint64_t tw = 0; c->st->max_rdlen = 0;
	for(uint32_t i = 0; i < n_reads; i++){
		if(rdoff[i] + rdlen[i] > n_bases) return zmo_set_err(ZMO_ERR_ARG, "read %u exceeds the bank", i);
		c->st->h_woff[i] = tw; tw += (((uint64_t)rdlen[i] + 15) / 16 + 3) & ~3ULL; tw += 4;   /* 16-byte aligned, one spare quad */
		if(rdlen[i] > c->st->max_rdlen) c->st->max_rdlen = rdlen[i];
	}
	c->st->h_woff[n_reads] = tw;
	const uint64_t bank_words = (n_bases + 31) / 32 + 1;
	if(c->s0.reserve(bank_words * 8) || c->s1.reserve((size_t)n_reads * 8) || c->st->rd_words.reserve(tw * 4 + 64) || c->st->rd_woff.reserve(((size_t)n_reads + 1) * 8) || c->st->rd_len.reserve((size_t)n_reads * 4)) return ZMO_ERR_CUDA;
	{
		StageTimer t(c, ST_COPY);
		CUDA_TRY(cudaMemcpyAsync(c->s0.p, bank, ((n_bases + 31) / 32) * 8, cudaMemcpyHostToDevice, c->stream));
		CUDA_TRY(cudaMemcpyAsync(c->s1.p, rdoff, (size_t)n_reads * 8, cudaMemcpyHostToDevice, c->stream));
		CUDA_TRY(cudaMemcpyAsync(c->st->rd_len.p, rdlen, (size_t)n_reads * 4, cudaMemcpyHostToDevice, c->stream));
		CUDA_TRY(cudaMemcpyAsync(c->st->rd_woff.p, c->st->h_woff.data(), ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, c->stream));
		const int bs = 256; const uint64_t nb = (tw + bs - 1) / bs;
		k_repack<<<(unsigned)nb, bs, 0, c->stream>>>(c->s0.as<unsigned long long>(), c->s1.as<unsigned long long>(), c->st->rd_len.as<uint32_t>(), c->st->rd_woff.as<unsigned long long>(), n_reads, c->st->rd_words.as<uint32_t>(), tw);
		c->launches++;
		CUDA_TRY(cudaGetLastError());
	}
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	c->counters[5] += ((n_bases + 31) / 32) * 8 + (uint64_t)n_reads * 12;
	return ZMO_OK;
}
