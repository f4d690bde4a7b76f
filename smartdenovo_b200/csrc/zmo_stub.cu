#include "zmo_ctx.cuh"
extern "C" int zmo_index_build(zmo_ctx *, uint32_t, uint32_t, uint32_t *, zmo_index_stats_t *){ return zmo_set_err(ZMO_ERR_STATE, "not implemented"); }
extern "C" int zmo_candidates(zmo_ctx *, const uint32_t *, uint32_t, uint64_t *, zmo_event_t *, uint64_t, uint64_t *){ return zmo_set_err(ZMO_ERR_STATE, "not implemented"); }
extern "C" int zmo_pair_windows(zmo_ctx *, int, const zmo_pair_t *, uint32_t, zmo_pairseed_t *, zmo_window_t *, uint64_t, uint64_t *){ return zmo_set_err(ZMO_ERR_STATE, "not implemented"); }
extern "C" int zmo_pair_align(zmo_ctx *, int, const zmo_task_t *, uint32_t, zmo_record_t *, uint32_t *, uint64_t, uint64_t *){ return zmo_set_err(ZMO_ERR_STATE, "not implemented"); }
extern "C" int zmo_pair_dotmatrix(zmo_ctx *, const zmo_pair_t *, uint32_t, zmo_dotres_t *){ return zmo_set_err(ZMO_ERR_STATE, "not implemented"); }
