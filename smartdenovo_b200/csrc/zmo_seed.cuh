/* zmo_seed.cuh -- shared between zmo_seed.cu (SW seeding) and zmo_dot.cu (dot-matrix mode) */
#pragma once
#include "zmo_ctx.cuh"
#include "zmo_seed_core.cuh"
/* z-index of the batch's query reads + per-pair z-mer match lists (in emission order, unsorted).
 * cache_off (np+1 entries, device) delimits each pair's list inside cache. */
struct SeedWork { uint32_t np, nuq; unsigned long long T; unsigned long long *cache_off; DevZPair *cache; };
int seed_prepare(zmo_ctx *c, const zmo_pair_t *pairs, uint32_t np, SeedWork &W, DevBuf &cache_buf);
