/*
 * zmo_nccl.cu -- the one exchange step of the multi-GPU overlap path: every GPU runs its own query shard
 * (the reference's `-P n -p g` job, wtzmo.c:1291,1314) with no communication, and at the end the variable-length record
 * text of all GPUs is gathered on the root GPU over NVLink with NCCL -- one all-gather of the sizes, then one grouped
 * send / receive of the payloads (gather to root: only the root receives) -- which is `cat part*.ovl` in job order
 * (usage, wtzmo.c:1431-1433).  One process drives all GPUs (one host thread per GPU, wtzmo_main.c: run_multi).
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2, RTLD_LOCAL): the library has no link-time dependency on it, a
 * single-GPU run never loads it, and a host process that already carries its own NCCL (PyTorch) is not disturbed.
 * No fallback: if NCCL cannot be loaded or a call fails, the gather fails with ZMO_ERR_CUDA.
 */
#include <dlfcn.h>
#include "zmo_ctx.cuh"

typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
enum { zncclUint8 = 1, zncclUint64 = 5 };      /* ncclDataType_t values of nccl.h (ncclUint8, ncclUint64) */
struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*GetVersion)(int *) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load(){
	if(g_nccl.h) return 0;
	const char *names[] = {getenv("ZMO_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	void *h = nullptr;
	for(const char *n : names){ if(n && *n && (h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break; }
	if(!h) return zmo_set_err(ZMO_ERR_CUDA, "NCCL is required for the multi-GPU record gather and could not be loaded (%s)", dlerror());
#define ZSYM(f) do { *(void**)&g_nccl.f = dlsym(h, "nccl" #f); if(!g_nccl.f){ dlclose(h); return zmo_set_err(ZMO_ERR_CUDA, "libnccl lacks nccl" #f); } } while(0)
	ZSYM(CommInitAll); ZSYM(CommDestroy); ZSYM(GroupStart); ZSYM(GroupEnd); ZSYM(AllGather); ZSYM(Send); ZSYM(Recv); ZSYM(GetErrorString); ZSYM(GetVersion);
#undef ZSYM
	g_nccl.h = h;
	return 0;
}
#define NCCL_TRY(expr) do { ncclResult_t _r = (expr); if(_r != 0) return zmo_set_err(ZMO_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, g_nccl.GetErrorString(_r)); } while(0)

/* communicators of the process, created once per device set (ncclCommInitAll costs about a second: zmo_gather_prepare lets the host do it on a
 * helper thread while the jobs are still computing) */
#include <mutex>
static std::mutex g_comm_mu;
static std::vector<int> g_comm_devs;
static std::vector<ncclComm_t> g_comms;
static int nccl_comms(const std::vector<int> &devs, std::vector<ncclComm_t> &out){
	std::lock_guard<std::mutex> lk(g_comm_mu);
	if(g_comm_devs != devs){
		for(ncclComm_t c : g_comms) g_nccl.CommDestroy(c);
		g_comms.assign(devs.size(), nullptr); g_comm_devs.clear();
		NCCL_TRY(g_nccl.CommInitAll(g_comms.data(), (int)devs.size(), devs.data()));
		g_comm_devs = devs;
	}
	out = g_comms;
	return 0;
}
extern "C" int zmo_gather_prepare(zmo_ctx **ctxs, int n){
	if(!ctxs || n < 1) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	if(n == 1) return 0;
	if(int rc = nccl_load()) return rc;
	std::vector<int> devs(n); for(int g = 0; g < n; g++){ if(!ctxs[g]) return zmo_set_err(ZMO_ERR_ARG, "null context"); devs[g] = ctxs[g]->device; }
	for(int a = 0; a < n; a++) for(int b = a + 1; b < n; b++) if(devs[a] == devs[b]) return zmo_set_err(ZMO_ERR_ARG, "contexts %d and %d share device %d: one job per GPU", a, b, devs[a]);
	std::vector<ncclComm_t> comm;
	return nccl_comms(devs, comm);
}

/* Gather n byte strings (part g = parts[g], sizes[g] bytes, produced by the job of ctxs[g]'s GPU and sitting in host memory) on the
 * GPU of ctxs[0] and copy them, in order, to out (capacity out_cap).  Every part travels host -> its own GPU -> NVLink -> root GPU ->
 * host, so that the exchange is the device-to-device gather the path defines; *total = sum of sizes.  ctxs must sit on n DISTINCT
 * devices.  Called once, by one host thread, after all jobs have finished. */
extern "C" int zmo_gather_records(zmo_ctx **ctxs, int n, const void *const *parts, const uint64_t *sizes, void *out, uint64_t out_cap, uint64_t *total, double *ms_out){
	if(!ctxs || n < 1 || !parts || !sizes || !total) return zmo_set_err(ZMO_ERR_ARG, "null argument");
	uint64_t tot = 0, mx = 0;
	for(int g = 0; g < n; g++){ if(!ctxs[g]) return zmo_set_err(ZMO_ERR_ARG, "null context"); tot += sizes[g]; if(sizes[g] > mx) mx = sizes[g]; }
	*total = tot;
	if(tot > out_cap || (tot && !out)) return zmo_set_err(ZMO_ERR_CAPACITY, "output buffer too small: need %llu bytes", (unsigned long long)tot);
	for(int a = 0; a < n; a++) for(int b = a + 1; b < n; b++) if(ctxs[a]->device == ctxs[b]->device) return zmo_set_err(ZMO_ERR_ARG, "contexts %d and %d share device %d: one job per GPU", a, b, ctxs[a]->device);
	if(n == 1){ if(tot) memcpy(out, parts[0], tot); if(ms_out) *ms_out = 0; return 0; }
	if(int rc = nccl_load()) return rc;
	std::vector<int> devs(n); for(int g = 0; g < n; g++) devs[g] = ctxs[g]->device;
	std::vector<ncclComm_t> comm;
	if(int rc0 = nccl_comms(devs, comm)) return rc0;
	int rc = 0;
	std::vector<void*> dsend(n, nullptr); std::vector<uint64_t*> dsz(n, nullptr); void *droot = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	do {
		/* sizes: one all-gather (every rank learns every size; the root needs them to lay out its receive buffer) */
		for(int g = 0; g < n && !rc; g++){
			if(cudaSetDevice(devs[g]) != cudaSuccess || cudaMalloc(&dsend[g], sizes[g] + 16) != cudaSuccess || cudaMalloc((void**)&dsz[g], ((size_t)n + 1) * 8) != cudaSuccess){ rc = zmo_set_err(ZMO_ERR_CUDA, "cudaMalloc for the gather failed on device %d", devs[g]); break; }
			if(cudaMemcpyAsync(dsz[g] + n, &sizes[g], 8, cudaMemcpyHostToDevice, ctxs[g]->stream) != cudaSuccess) rc = zmo_set_err(ZMO_ERR_CUDA, "H2D failed");
			if(sizes[g] && cudaMemcpyAsync(dsend[g], parts[g], sizes[g], cudaMemcpyHostToDevice, ctxs[g]->stream) != cudaSuccess) rc = zmo_set_err(ZMO_ERR_CUDA, "H2D failed");      /* pageable source: staged by the driver */
			ctxs[g]->counters[5] += sizes[g];
		}
		if(rc) break;
		cudaSetDevice(devs[0]);
		if(cudaMalloc(&droot, tot + 16) != cudaSuccess){ rc = zmo_set_err(ZMO_ERR_CUDA, "cudaMalloc(%llu) for the gathered records failed", (unsigned long long)tot); break; }
		cudaEventCreate(&e0); cudaEventCreate(&e1);
		cudaEventRecord(e0, ctxs[0]->stream);
		if(g_nccl.GroupStart() != 0){ rc = zmo_set_err(ZMO_ERR_CUDA, "ncclGroupStart failed"); break; }
		for(int g = 0; g < n; g++){ ncclResult_t r = g_nccl.AllGather(dsz[g] + n, dsz[g], 1, zncclUint64, comm[g], ctxs[g]->stream); if(r != 0){ rc = zmo_set_err(ZMO_ERR_CUDA, "ncclAllGather failed: %s", g_nccl.GetErrorString(r)); break; } }
		if(g_nccl.GroupEnd() != 0 && !rc) rc = zmo_set_err(ZMO_ERR_CUDA, "ncclGroupEnd failed");
		if(rc) break;
		/* payloads: gather to root = one grouped round of sends (ranks 1..n-1) and receives (root); the root's own part is a device copy */
		if(g_nccl.GroupStart() != 0){ rc = zmo_set_err(ZMO_ERR_CUDA, "ncclGroupStart failed"); break; }
		{
			uint64_t off = sizes[0];
			for(int g = 1; g < n && !rc; g++){
				if(sizes[g]){
					ncclResult_t r = g_nccl.Recv((char*)droot + off, sizes[g], zncclUint8, g, comm[0], ctxs[0]->stream);
					if(r == 0) r = g_nccl.Send(dsend[g], sizes[g], zncclUint8, 0, comm[g], ctxs[g]->stream);
					if(r != 0) rc = zmo_set_err(ZMO_ERR_CUDA, "ncclSend/Recv failed: %s", g_nccl.GetErrorString(r));
				}
				off += sizes[g];
			}
		}
		if(g_nccl.GroupEnd() != 0 && !rc) rc = zmo_set_err(ZMO_ERR_CUDA, "ncclGroupEnd failed");
		if(rc) break;
		cudaSetDevice(devs[0]);
		if(sizes[0]) cudaMemcpyAsync(droot, dsend[0], sizes[0], cudaMemcpyDeviceToDevice, ctxs[0]->stream);
		cudaEventRecord(e1, ctxs[0]->stream);
		if(tot && cudaMemcpyAsync(out, droot, tot, cudaMemcpyDeviceToHost, ctxs[0]->stream) != cudaSuccess){ rc = zmo_set_err(ZMO_ERR_CUDA, "D2H of the gathered records failed"); break; }
		for(int g = 0; g < n; g++){ cudaSetDevice(devs[g]); if(cudaStreamSynchronize(ctxs[g]->stream) != cudaSuccess){ rc = zmo_set_err(ZMO_ERR_CUDA, "gather failed on device %d: %s", devs[g], cudaGetErrorString(cudaGetLastError())); break; } }
		if(rc) break;
		/* the root checks the sizes it was told against the ones it laid its buffer out with */
		{
			std::vector<uint64_t> got(n); cudaSetDevice(devs[0]);
			cudaMemcpy(got.data(), dsz[0], (size_t)n * 8, cudaMemcpyDeviceToHost);
			for(int g = 0; g < n; g++) if(got[g] != sizes[g]){ rc = zmo_set_err(ZMO_ERR_CUDA, "size all-gather mismatch for rank %d", g); break; }
		}
		ctxs[0]->counters[6] += tot;
		if(ms_out){ float ms = 0; cudaEventElapsedTime(&ms, e0, e1); *ms_out = ms; }
	} while(0);
	for(int g = 0; g < n; g++){ cudaSetDevice(devs[g]); if(dsend[g]) cudaFree(dsend[g]); if(dsz[g]) cudaFree(dsz[g]); }
	cudaSetDevice(devs[0]);
	if(droot) cudaFree(droot);
	if(e0) cudaEventDestroy(e0);
	if(e1) cudaEventDestroy(e1);
	return rc;
}

/* number of usable sm_100-class devices (the host's ZMO_GPUS=all) */
extern "C" int zmo_device_count(void){
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}
