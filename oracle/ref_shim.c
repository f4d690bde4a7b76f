/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin C-ABI shim that #includes the UNMODIFIED reference headers where they lie
 * (-I/root/reference at build time, see oracle/Makefile) and re-exports the reference's
 * static-inline stage functions with plain-pointer signatures, so tests can run known-answer
 * comparisons against the real reference code (not a restatement).  No reference source is
 * copied here; this file only contains call-through wrappers.
 *
 * Wrapped reference entry points (file:line in /root/reference):
 *   kswx_extend_align_core         kswx.h:234
 *   kswx_extend_align_shift_core   kswx.h:101
 *   ksw_global2                    ksw.c:503
 *   hz_align_hzmo                  hzm_aln.h:278
 *   index_single_read_seeds        hzm_aln.h:70
 *   query_single_read_seeds        hzm_aln.h:173
 *   process_hzmps                  hzm_aln.h:1184
 *   merge_paired_kmers_window      hzm_aln.h:580
 *   chaining_wtseedv               hzm_aln.h:658
 *   fast_seeds_align_hzmo          hzm_aln.h:1247
 *   global_align_regs_hzmo         hzm_aln.h:1345
 *   dot_matrix_align_hzmps         hzm_aln.h:1134
 *   sort_array                     sort.h:104
 */
#include "list.h"
#include "hashset.h"
#include "dna.h"
#include "kswx.h"
#include "hzm_aln.h"
#include "bitvec.h"
#include <stdint.h>
#include <string.h>

/* out[10] = score,tb,te,qb,qe,aln,mat,mis,ins,del ; cigar_out receives <=cigar_cap ops, returns n ops */
static int export_x(kswx_t x, int *out){
	out[0]=x.score; out[1]=x.tb; out[2]=x.te; out[3]=x.qb; out[4]=x.qe;
	out[5]=x.aln; out[6]=x.mat; out[7]=x.mis; out[8]=x.ins; out[9]=x.del;
	return 0;
}

int ref_extend_core(int qlen, uint8_t *q, int tlen, uint8_t *t, int strand, int init, int W,
		int M, int X, int I, int D, int E, int T, int *out, uint32_t *cigar_out, int cigar_cap){
	u8list *mem = init_u8list(1024);
	u32list *cg = init_u32list(64);
	kswx_t x = kswx_extend_align_core(qlen, q, tlen, t, strand, init, W, M, X, I, D, E, T, mem, cg);
	int n = (int)cg->size, i;
	export_x(x, out);
	for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg->buffer[i];
	free_u8list(mem); free_u32list(cg);
	return n;
}

int ref_extend_shift_core(int qlen, uint8_t *q, int tlen, uint8_t *t, int strand, int init, int W,
		int M, int X, int I, int D, int E, int T, int *out, uint32_t *cigar_out, int cigar_cap){
	u8list *mem = init_u8list(1024);
	u32list *cg = init_u32list(64);
	kswx_t x = kswx_extend_align_shift_core(qlen, q, tlen, t, strand, init, W, M, X, I, D, E, T, mem, cg);
	int n = (int)cg->size, i;
	export_x(x, out);
	for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg->buffer[i];
	free_u8list(mem); free_u32list(cg);
	return n;
}

/* returns n_cigar; *score_out = score */
int ref_global2(int qlen, uint8_t *q, int tlen, uint8_t *t, int M, int X, int o_del, int e_del,
		int o_ins, int e_ins, int w, int *score_out, uint32_t *cigar_out, int cigar_cap){
	int8_t mat[16]; int i, n = 0; uint32_t *cg = NULL;
	for(i=0;i<16;i++) mat[i] = ((i%4)==(i/4))? M : X;
	*score_out = ksw_global2(qlen, q, tlen, t, 4, mat, o_del, e_del, o_ins, e_ins, w, &n, &cg);
	for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg[i];
	free(cg);
	return n;
}

int ref_hz_align(uint8_t *pb1, uint32_t len1, uint8_t *pb2, uint32_t len2, int M, int I, int D, int E,
		int *out, uint32_t *cigar_out, int cigar_cap){
	u32list *cg = init_u32list(64);
	kswx_t x = hz_align_hzmo(pb1, len1, pb2, len2, M, I, D, E, cg);
	int n = (int)cg->size, i;
	export_x(x, out);
	for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg->buffer[i];
	free_u32list(cg);
	return n;
}

/* sort_array on uint64 keys with "a > b" / low-32-bit-desc comparators, to pin tie permutations */
void ref_sort_u64_asc(uint64_t *a_, size_t n){ sort_array(a_, n, uint64_t, a > b); }
void ref_sort_u64_lo32_desc(uint64_t *a_, size_t n){ sort_array(a_, n, uint64_t, (b & 0xFFFFFFFFU) > (a & 0xFFFFFFFFU)); }
/* sort by high 32 bits only (payload in low bits exposes the permutation) */
void ref_sort_u64_hi32_asc(uint64_t *a_, size_t n){ sort_array(a_, n, uint64_t, (a >> 32) > (b >> 32)); }
void ref_sort_u64_hi32_desc(uint64_t *a_, size_t n){ sort_array(a_, n, uint64_t, (b >> 32) > (a >> 32)); }

/*
 * Pair stage (wtzmo.c:845-914 for ONE candidate): z-index of pb1, z-matches of pb2, windows and
 * chain weight per strand.  Outputs:
 *   n_hzmp            number of z-mer match pairs (cache->size)
 *   ovl[2]            chaining_wtseedv result per strand (0 if merge returned 0)
 *   win_out           per kept window: dir,beg0,end0,beg1,end1,ovl,n_anchors  (7 ints), strand 0 first
 *   anc_out           per anchor: off1,off2,len1,len2,dir1,dir2 (6 ints) in window order
 * returns number of windows written; *n_anc_out anchors.
 */
int ref_pair_windows(uint8_t *pb1, int alen, uint8_t *pb2, int blen, int zsize, int hz, int zcut, int kvar,
		int kwin, int kstep, int zovl, int ztot, int W,
		int *n_hzmp, int *ovl, int *win_out, int win_cap, int *anc_out, int anc_cap, int *n_anc_out){
	hzmhv *zhash = init_hzmhv(1023);
	hzmv *zseeds = init_hzmv(64);
	u32list *hzoff = init_u32list(64);
	u1v *kcnts = init_u1v(1024);
	hzmpv *cache = init_hzmpv(64), *anchors2 = init_hzmpv(64);
	wtseedv *windows2 = init_wtseedv(64);
	u8list *mem_cache[2]; mem_cache[0] = init_u8list(64); mem_cache[1] = init_u8list(64);
	BitVec *zbits = init_bitvec(0xFFFFFFFFFFFFFFFFLLU >> ((32 - zsize) << 1));
	int dir, nw = 0, na = 0; uint32_t j, k;
	HZM_FAST_WINDOW_KMER_CHAINING = 1;
	index_single_read_seeds(pb1, alen, zsize, hz, zcut, zhash, zbits, zseeds, hzoff);
	query_single_read_seeds(pb2, blen, zsize, hz, zcut, kvar, zhash, zbits, zseeds, hzoff, kcnts, cache);
	*n_hzmp = (int)cache->size;
	ovl[0] = ovl[1] = 0;
	if(cache->size * zsize >= (uint32_t)ztot){
		process_hzmps(cache);
		for(dir=0;dir<2;dir++){
			clear_wtseedv(windows2); clear_hzmpv(anchors2);
			if(merge_paired_kmers_window(cache, dir, windows2, anchors2, mem_cache, zsize, kwin, kstep, zovl) == 0) continue;
			ovl[dir] = chaining_wtseedv(0, 1, dir, windows2, 0, windows2->size, mem_cache[0], W);
			for(j=0;j<windows2->size;j++){
				wt_seed_t *zp = ref_wtseedv(windows2, j);
				if(zp->closed) continue;
				if(nw < win_cap){
					int *o = win_out + 7 * nw;
					o[0]=dir; o[1]=zp->beg[0]; o[2]=zp->end[0]; o[3]=zp->beg[1]; o[4]=zp->end[1]; o[5]=zp->ovl; o[6]=zp->anchors[1]-zp->anchors[0];
				}
				nw ++;
				for(k=zp->anchors[0];k<zp->anchors[1];k++){
					hzmp_t *p = ref_hzmpv(anchors2, k);
					if(na < anc_cap){
						int *o = anc_out + 6 * na;
						o[0]=p->off1; o[1]=p->off2; o[2]=p->len1; o[3]=p->len2; o[4]=p->dir1; o[5]=p->dir2;
					}
					na ++;
				}
			}
		}
	}
	*n_anc_out = na;
	free_bitvec(zbits);
	free_hzmhv(zhash); free_hzmv(zseeds); free_u32list(hzoff); free_u1v(kcnts);
	free_hzmpv(cache); free_hzmpv(anchors2); free_wtseedv(windows2);
	free_u8list(mem_cache[0]); free_u8list(mem_cache[1]);
	return nw;
}

/* dot-matrix stage for one pair (wtzmo.c:853-863): out = score,qb,qe,tb,te,strand ; returns n_hzmp */
int ref_pair_dotmatrix(uint8_t *pb1, int alen, uint8_t *pb2, int blen, int zsize, int hz, int zcut, int kvar,
		int xvar, int yvar, int min_block_len, int max_overhang, float dev_pen, float gap_pen, int *out){
	hzmhv *zhash = init_hzmhv(1023);
	hzmv *zseeds = init_hzmv(64);
	u32list *hzoff = init_u32list(64);
	u1v *kcnts = init_u1v(1024);
	hzmpv *cache = init_hzmpv(64), *dst[2];
	wtseedv *wins[2];
	diagv *diags = init_diagv(64);
	u4v *block = init_u4v(64), *grps = init_u4v(64);
	u8list *mem = init_u8list(64);
	BitVec *zbits = init_bitvec(0xFFFFFFFFFFFFFFFFLLU >> ((32 - zsize) << 1));
	kswr_t r; int n;
	dst[0] = init_hzmpv(64); dst[1] = init_hzmpv(64);
	wins[0] = init_wtseedv(64); wins[1] = init_wtseedv(64);
	index_single_read_seeds(pb1, alen, zsize, hz, zcut, zhash, zbits, zseeds, hzoff);
	query_single_read_seeds(pb2, blen, zsize, hz, zcut, kvar, zhash, zbits, zseeds, hzoff, kcnts, cache);
	n = (int)cache->size;
	r = dot_matrix_align_hzmps(cache, dst, wins, diags, block, grps, mem, alen, blen, xvar, yvar, min_block_len, max_overhang, dev_pen, gap_pen);
	out[0]=r.score; out[1]=r.qb; out[2]=r.qe; out[3]=r.tb; out[4]=r.te; out[5]=r.score2;
	free_bitvec(zbits);
	free_hzmhv(zhash); free_hzmv(zseeds); free_u32list(hzoff); free_u1v(kcnts);
	free_hzmpv(cache); free_hzmpv(dst[0]); free_hzmpv(dst[1]); free_wtseedv(wins[0]); free_wtseedv(wins[1]);
	free_diagv(diags); free_u4v(block); free_u4v(grps); free_u8list(mem);
	return n;
}
