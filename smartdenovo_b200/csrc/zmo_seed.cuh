/* zmo_seed.cuh -- shared between zmo_seed.cu (SW seeding) and zmo_dot.cu (dot-matrix mode) */
#pragma once
#include "zmo_ctx.cuh"
#include "zmo_seed_core.cuh"
/* z-index of the batch's query reads + per-pair z-mer match lists (in emission order, unsorted).
 * cache_off (np+1 entries, device) delimits each pair's list inside cache. */
struct SeedWork { uint32_t np, nuq; unsigned long long T; unsigned long long *cache_off; DevZPair *cache; const uint8_t *tie; const uint32_t *pc; };
/* mode 0: lists sorted by (off1,off2) (SW path); mode 1: sorted by (off1-off2, off1) (dot-matrix path).
 * tie[p] != 0 marks pairs whose list has equal keys (see k_unpack). */
int seed_prepare(zmo_ctx *c, const zmo_pair_t *pairs, uint32_t np, int mode, SeedWork &W, DevBuf &cache_buf);
