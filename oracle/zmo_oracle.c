/*
 * zmo_oracle.c -- CPU ORACLE for the wtzmo all-vs-all overlap path.  TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch, single-threaded restatement of the algorithm of ruanjue/smartdenovo `wtzmo`
 * (deterministic `-t 1` semantics).  It is NOT part of the product: only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline leg may build or run it.  Parity status: PINNED --
 * tests/test_oracle_vs_ref.py compares (a) every DP routine and the pair-window stage against the
 * real reference functions through oracle/_ref/libzmo_ref.so (ref_shim.c) and (b) complete .ovl /
 * .contained / -9 outputs byte-for-byte against the unmodified reference binary oracle/_ref/wtzmo.
 *
 * Structure differs from the reference on purpose: every numeric stage is a pure function
 * (index -> candidate events -> pair windows -> pair alignment) and all mutable cross-read state
 * (masked reads, tried pairs, per-read overlap counters) lives in one sequential replay loop.
 * Each function cites the reference file:line whose behaviour it restates.
 *
 * Build: see oracle/Makefile.  `zmo_oracle` takes the wtzmo command line.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <getopt.h>
#include <math.h>
#include <sys/stat.h>

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64; typedef int32_t i32; typedef int64_t i64;

/* ------------------------------------------------------------------ growable array */
#define VEC(T) struct { T *a; size_t n, m; }
#define vec_init(v) ((v).a = NULL, (v).n = (v).m = 0)
#define vec_free(v) (free((v).a), (v).a = NULL, (v).n = (v).m = 0)
#define vec_reserve(v, need) do { size_t _nd = (need); if(_nd > (v).m){ size_t _m = (v).m? (v).m : 16; while(_m < _nd) _m <<= 1; (v).a = realloc((v).a, _m * sizeof(*(v).a)); (v).m = _m; } } while(0)
#define vec_push(v, x) do { vec_reserve(v, (v).n + 1); (v).a[(v).n++] = (x); } while(0)
#define vec_clear(v) ((v).n = 0)
typedef VEC(u32) u32v; typedef VEC(u64) u64v; typedef VEC(u8) u8v; typedef VEC(i32) i32v;

#define imin(a,b) ((a) < (b)? (a) : (b))
#define imax(a,b) ((a) > (b)? (a) : (b))
#define idiff(a,b) ((a) > (b)? (a) - (b) : (b) - (a))

/* ------------------------------------------------------------------ sort_array emulation (sort.h:104-155)
 * Median-of-3 quicksort with an explicit stack that leaves partitions of <=5 elements to a final
 * bubble pass.  Deterministic but not stable: the permutation of equal keys is part of the contract
 * at several call sites, so the exact sequence of swaps is reproduced. */
typedef int (*gt_fn)(const void *a, const void *b, void *ctx);
#define SORT_MAX_ES 64
static void ref_sort(void *base, size_t n, size_t es, gt_fn gt, void *ctx){
	u8 *rs = (u8*)base, piv[SORT_MAX_ES], tmp[SORT_MAX_ES];
	size_t stack[64][2], x = 0, s, e, i, j, m;
#define EL(k) (rs + (k) * es)
#define SWAP(p, q) do { memcpy(tmp, EL(p), es); memcpy(EL(p), EL(q), es); memcpy(EL(q), tmp, es); } while(0)
	if(n < 2) return;
	stack[0][0] = 0; stack[0][1] = n - 1; x = 1;
	while(x){
		x --; s = stack[x][0]; e = stack[x][1];
		m = s + (e - s) / 2;
		if(gt(EL(s), EL(m), ctx) > 0) SWAP(s, m);
		if(gt(EL(m), EL(e), ctx) > 0){
			SWAP(e, m);
			if(gt(EL(s), EL(m), ctx) > 0) SWAP(s, m);
		}
		memcpy(piv, EL(m), es);
		i = s + 1; j = e - 1;
		while(1){
			while(gt(piv, EL(i), ctx) > 0) i ++;
			while(gt(EL(j), piv, ctx) > 0) j --;
			if(i < j){ SWAP(i, j); i ++; j --; }
			else break;
		}
		if(i == j){ i ++; j --; }
		if(j - s > e - i){
			if(s + 4 < j){ stack[x][0] = s; stack[x][1] = j; x ++; }
			if(i + 4 < e){ stack[x][0] = i; stack[x][1] = e; x ++; }
		} else {
			if(i + 4 < e){ stack[x][0] = i; stack[x][1] = e; x ++; }
			if(s + 4 < j){ stack[x][0] = s; stack[x][1] = j; x ++; }
		}
	}
	for(i=0;i<n;i++){
		int swapped = 0;
		for(j=n-1;j>i;j--){
			if(gt(EL(j - 1), EL(j), ctx) > 0){ SWAP(j - 1, j); swapped = 1; }
		}
		if(!swapped) break;
	}
#undef EL
#undef SWAP
}

/* ------------------------------------------------------------------ 2-bit read store (dna.h:78,263,397-471) */
static inline u32 bank_get(const u64 *bits, u64 off){ return (bits[off >> 5] >> (((~off) & 31) << 1)) & 3; }
static inline void bank_put(u64 *bits, u64 off, u64 b){ if((off & 31) == 0) bits[off >> 5] = 0; bits[off >> 5] |= b << (((~off) & 31) << 1); }

/* reverse complement of a right-aligned k-mer (dna.h:85-97) */
static inline u64 kmer_revcomp(u64 x, int k){
	x = ~x;
	x = ((x & 0x3333333333333333ULL) << 2) | ((x >> 2) & 0x3333333333333333ULL);
	x = ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL);
	x = __builtin_bswap64(x);
	return x >> (64 - 2 * k);
}

/* hashset.h:452-462; used as the k-mer sub-sampling function on the LOW 32 bits (wtzmo.c:35) */
static inline u32 jenkins32(u32 key){
	key += (key << 12); key ^= (key >> 22); key += (key << 4); key ^= (key >> 9);
	key += (key << 10); key ^= (key >> 2); key += (key << 7); key ^= (key >> 12);
	return key;
}

/* ------------------------------------------------------------------ parameters (wtzmo.c:1543-1588) */
typedef struct {
	int ncpu, n_job, i_job, do_align, min_rdlen, overwrite, skip_contained, write_contained, refine, debug;
	int hk, hz, ksize, zsize, kwin, kstep, kovl, ksave, n_idx, ztot, zovl, kcut, zcut, kvar;
	float wnorm, wrep;
	int ncand, nbest;
	int w, ew, W, M, X, O, E, T, min_score;
	float min_id;
	int dot_matrix, xvar, yvar, min_block_len, max_overhang;
	float deviation_penalty, gap_penalty;
	u32 max_unalign_in_contained, max_unalign_in_dovetail;
} zparams_t;

static void zparams_default(zparams_t *p){
	memset(p, 0, sizeof(*p));
	p->ncpu = 1; p->n_job = 1; p->i_job = 0; p->do_align = 1; p->skip_contained = 1; p->write_contained = 1;
	p->hk = 1; p->hz = 1; p->ksize = 16; p->zsize = 10; p->kwin = 800; p->kovl = 300; p->ksave = 4; p->n_idx = 1;
	p->wnorm = 20; p->wrep = 100; p->ncand = 500; p->nbest = 100; p->ztot = 300; p->zovl = 200; p->kcut = 0; p->zcut = 64; p->kvar = 2;
	p->w = 50; p->ew = 800; p->W = 3200; p->M = 2; p->X = -5; p->O = -3; p->E = -1; p->T = -50; p->min_score = 200; p->min_id = 0.5;
	p->dot_matrix = 0; p->xvar = 128; p->yvar = 64; p->min_block_len = 160; p->max_overhang = 256;
	p->deviation_penalty = 1.0; p->gap_penalty = 0.05;
	p->max_unalign_in_contained = 0; p->max_unalign_in_dovetail = 200; /* wtzmo.c:174-175, not settable */
}

/* ------------------------------------------------------------------ alignment result + CIGAR helpers (kswx.h:30-52) */
typedef struct { int score, tb, te, qb, qe, aln, mat, mis, ins, del; } aln_t;
static const aln_t ALN_NULL = {0,0,0,0,0,0,0,0,0,0};

static inline void cig_push(u32v *c, u32 op, u32 len){            /* kswx.h:39-44 */
	if(len == 0) return;
	if(c->n && (c->a[c->n-1] & 0xF) == op) c->a[c->n-1] += len << 4;
	else vec_push(*c, (len << 4) | op);
}
static inline void cig_append(u32v *c, const u32 *src, size_t n){   /* kswx.h:46-52 */
	size_t i = 0;
	if(n == 0) return;
	if(c->n && (c->a[c->n-1] & 0xF) == (src[0] & 0xF)){ c->a[c->n-1] += src[0] & 0xFFFFFFF0U; i = 1; }
	for(;i<n;i++) vec_push(*c, src[i]);
}
static inline void cig_reverse(u32v *c){ size_t i; for(i=0;i<c->n/2;i++){ u32 t = c->a[i]; c->a[i] = c->a[c->n-1-i]; c->a[c->n-1-i] = t; } }

/* ------------------------------------------------------------------ banded affine extension DP
 * Restates kswx_extend_align_core (kswx.h:234-335, mode 0: fixed band |i-j|<=W, arg-max keeps the
 * LAST column) and kswx_extend_align_shift_core (kswx.h:101-232, mode 1: band centre follows the
 * row arg-max, which keeps the FIRST column).  Sequences are one base per byte; element k of the
 * query is q[k*strand] (strand=-1 walks backwards from the pointer).
 *
 * Model: rows i = query, columns j = target, H(-1,-1)=init,
 * H(-1,j)=init+D+E(j+1), H(i,-1)=init+I+E(i+1) while the band touches column 0; every other
 * out-of-band H/E neighbour and the row-initial F read as -10000.  e/f are opened from m (the
 * diagonal move), not from h.  Traceback byte: bits0-1 source of H (0 M,1 E,2 F), bit2 E extended,
 * bit5 F extended. */
#define NEG_SENT (-10000)
static aln_t banded_extend(int mode, int qlen, const u8 *q, int tlen, const u8 *t, int strand, int init, int W,
		int M, int X, int I, int D, int E, int T, u32v *cig){
	aln_t x = ALN_NULL;
	int ql, tl, ncol, i, j, c, jb, je, prev_jb, prev_je;
	int best, bi, bj, gbest, gi, gj;
	int *Hp, *Hc, *Ev; u8 *z; int *zb;
	if(cig) vec_clear(*cig);
	if(init < 0) init = 0;
	if(qlen <= 0 || tlen <= 0){ x.score = init; return x; }
	if(W > 0){
		int mx = imin(qlen, tlen) * M + init + (-T);
		int max_gap = (mx + imax(I, D)) / (-E) + 1;
		if(max_gap < 1) max_gap = 1;
		if(W > max_gap) W = max_gap;
	} else W = -W;
	W = imin(W, imax(qlen, tlen));
	ql = qlen; tl = tlen;
	if(qlen < tlen){ if(qlen + W < tlen) tl = qlen + W; }
	else { if(tlen + W < qlen) ql = tlen + W; }
	ncol = imin(tl, 2 * W + 1);
	/* Hp[j+1] = H(i-1, j) for j in [-1, tl) under the out-of-band model; Ev[j] = E(i, j) */
	Hp = malloc((tl + 2) * sizeof(int)); Hc = malloc((tl + 2) * sizeof(int)); Ev = malloc((tl + 2) * sizeof(int));
	z = malloc((size_t)ql * ncol); zb = malloc((ql + 1) * sizeof(int));
	best = init; bi = bj = -1; gbest = 0; gi = gj = -1;
	prev_jb = 0; prev_je = tl; /* row -1 is defined on all columns */
	Hp[0] = init; for(j=0;j<tl;j++){ Hp[j+1] = init + D + E * (j + 1); Ev[j] = NEG_SENT; }
	for(i=0,c=0;i<ql;i++){
		int rowmax = 0, rowarg = -1, f = NEG_SENT, hleft;
		if(mode == 1){ jb = imax(0, c - W); je = imin(tl, c + W + 1); }
		else { jb = imax(0, i - W); je = imin(tl, i + W + 1); }
		zb[i] = jb;
		hleft = jb == 0? init + I + E * (i + 1) : NEG_SENT;   /* H(i, jb-1) */
		for(j=jb;j<je;j++){
			/* diagonal predecessor H(i-1, j-1) */
			int hd, e, m, h, tt; u8 d;
			if(j == 0) hd = (i == 0)? init : init + I + E * i;
			else if(j - 1 >= prev_jb && j - 1 < prev_je) hd = Hp[j];
			else hd = NEG_SENT;
			e = (j >= prev_jb && j < prev_je)? Ev[j] : NEG_SENT;
			if(i == 0) e = NEG_SENT;
			m = hd + ((q[(long)i * strand] == t[(long)j * strand])? M : X);
			if(m >= e){ d = 0; h = m; } else { d = 1; h = e; }
			if(h < f){ d = 2; h = f; }
			Hc[j+1] = h;
			if(mode == 1){ if(h > rowmax){ rowmax = h; rowarg = j; } }
			else { if(h >= rowmax){ rowmax = h; rowarg = j; } }
			tt = m + I + E; e += E; if(e > tt) d |= 1 << 2; else e = tt; Ev[j] = e;
			tt = m + D + E; f += E; if(f > tt) d |= 2 << 4; else f = tt;
			z[(size_t)i * ncol + (j - jb)] = d;
			hleft = h;
		}
		if(je == tlen && gbest < hleft){ gbest = hleft; gi = i; gj = je - 1; }
		if(i + 1 == qlen && gbest < rowmax){ gbest = rowmax; gi = i; gj = rowarg; }
		{ int *sw = Hp; Hp = Hc; Hc = sw; }
		prev_jb = jb; prev_je = je;
		if(rowmax > best){ best = rowmax; bi = i; bj = rowarg; }
		else if(rowmax <= 0) break;
		if(mode == 1){ c ++; if(c < rowarg) c ++; else if(c > rowarg) c --; }
	}
	if(gbest > 0 && gbest >= best + T){ x.score = gbest; x.qe = gi; x.te = gj; }
	else { x.score = best; x.qe = bi; x.te = bj; }
	{
		int st = 0;
		i = x.qe; j = x.te;
		while(i >= 0 && j >= 0){
			st = (z[(size_t)i * ncol + (j - zb[i])] >> (st << 1)) & 3;
			if(st == 0){ if(q[(long)i * strand] == t[(long)j * strand]) x.mat ++; else x.mis ++; i --; j --; }
			else if(st == 1){ i --; x.ins ++; }
			else { j --; x.del ++; }
			if(cig) cig_push(cig, st, 1);
		}
		if(i >= 0){ x.ins += i + 1; if(cig) cig_push(cig, 1, i + 1); }
		if(j >= 0){ x.del += j + 1; if(cig) cig_push(cig, 2, j + 1); }
		if(cig) cig_reverse(cig);
	}
	x.aln = x.mat + x.mis + x.ins + x.del;
	x.qe ++; x.te ++;
	free(Hp); free(Hc); free(Ev); free(z); free(zb);
	return x;
}

/* ------------------------------------------------------------------ banded global DP (ksw.c:503-586)
 * Rows i = target, columns j = query, band |i-j| <= w, MINUS_INF = -0x40000000 kept literally.
 * CIGAR op 1 consumes a query base, op 2 a target base.  Returns the score of cell (tlen-1,qlen-1). */
#define G_NEG (-0x40000000)
static int banded_global(int qlen, const u8 *q, int tlen, const u8 *t, int M, int X, int o_del, int e_del, int o_ins, int e_ins, int w, u32v *cig){
	int i, j, k, ncol = imin(qlen, 2 * w + 1), score;
	int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int *Hd = malloc((qlen + 2) * sizeof(int)), *Ev = malloc((qlen + 2) * sizeof(int));
	u8 *z = malloc((size_t)imax(ncol, 1) * imax(tlen, 1));
	vec_clear(*cig);
	/* Hd[j] = H(i-1, j-1) */
	Hd[0] = 0; Ev[0] = G_NEG;
	for(j=1;j<=qlen && j<=w;j++){ Hd[j] = -(o_ins + e_ins * j); Ev[j] = G_NEG; }
	for(;j<=qlen;j++){ Hd[j] = G_NEG; Ev[j] = G_NEG; }
	for(i=0;i<tlen;i++){
		int f = G_NEG, beg = i > w? i - w : 0, end = i + w + 1 < qlen? i + w + 1 : qlen;
		int h1 = beg == 0? -(o_del + e_del * (i + 1)) : G_NEG;
		for(j=beg;j<end;j++){
			int m = Hd[j], e = Ev[j], h, tt; u8 d;
			Hd[j] = h1;
			m += (q[j] == t[i])? M : X;
			d = m >= e? 0 : 1; h = m >= e? m : e;
			d = h >= f? d : 2; h = h >= f? h : f;
			h1 = h;
			tt = m - oe_del; e -= e_del; if(e > tt) d |= 1 << 2; else e = tt; Ev[j] = e;
			tt = m - oe_ins; f -= e_ins; if(f > tt) d |= 2 << 4; else f = tt;
			z[(size_t)i * ncol + (j - beg)] = d;
		}
		Hd[end] = h1; Ev[end] = G_NEG;
	}
	score = Hd[qlen];
	{
		int which = 0;
		i = tlen - 1; k = (i + w + 1 < qlen? i + w + 1 : qlen) - 1;
		while(i >= 0 && k >= 0){
			which = (z[(size_t)i * ncol + (k - (i > w? i - w : 0))] >> (which << 1)) & 3;
			if(which == 0){ cig_push(cig, 0, 1); i --; k --; }
			else if(which == 1){ cig_push(cig, 2, 1); i --; }
			else { cig_push(cig, 1, 1); k --; }
		}
		if(i >= 0) cig_push(cig, 2, i + 1);
		if(k >= 0) cig_push(cig, 1, k + 1);
		cig_reverse(cig);
	}
	free(Hd); free(Ev); free(z);
	return score;
}

/* ------------------------------------------------------------------ run-length anchor alignment (hzm_aln.h:278-314)
 * The two slices are the same homopolymer-compressed z-mer; equal runs -> M, length difference -> I or D.
 * Returns ALN_NULL (aln==0) if the run bases differ. */
static aln_t runlen_align(const u8 *a, u32 la, const u8 *b, u32 lb, int M, int I, int D, int E, u32v *cig){
	aln_t x = ALN_NULL; u32 sa = 0, sb = 0;
	while(sa < la || sb < lb){
		u32 ea, eb, na, nb;
		if(a[sa] != b[sb]) return ALN_NULL;
		ea = sa + 1; while(ea < la && a[ea] == a[sa]) ea ++;
		eb = sb + 1; while(eb < lb && b[eb] == b[sb]) eb ++;
		na = ea - sa; nb = eb - sb;
		if(na < nb){ x.aln += nb; x.mat += na; x.ins += nb - na; x.score += na * M + I + (nb - na) * E; cig_push(cig, 0, na); cig_push(cig, 1, nb - na); }
		else if(na == nb){ x.aln += na; x.mat += na; x.score += na * M; cig_push(cig, 0, na); }
		else { x.aln += na; x.mat += nb; x.del += na - nb; x.score += nb * M + D + (na - nb) * E; cig_push(cig, 0, nb); cig_push(cig, 2, na - nb); }
		sa = ea; sb = eb;
	}
	x.te = x.mat + x.del; x.qe = x.mat + x.ins;
	return x;
}

/* ------------------------------------------------------------------ read set + FASTA/FASTQ reader
 * Behaviour of file_reader.c:296-424 as used by wtzmo.c:1691-1703: type guessed from the first
 * non-empty, non-'#' line; FASTA name = header up to first blank; multi-line sequences are
 * concatenated; FASTQ = 4-line records; "*.gz" through `gzip -dc`; several files are chained. */
typedef struct { u64 off; u32 len; char *name; } read_t;
typedef struct {
	u64 *bits; u64 nbases, cap_words;
	VEC(read_t) reads;
	u32 n_rd, n_qr;
} readset_t;

static void rs_add_read(readset_t *rs, const char *name, int name_len, const char *seq, u32 len){
	read_t r; u32 i;
	u64 need = (rs->nbases + len + 31) / 32 + 2;
	if(need > rs->cap_words){ u64 m = rs->cap_words? rs->cap_words : 1024; while(m < need) m <<= 1; rs->bits = realloc(rs->bits, m * 8); memset(rs->bits + rs->cap_words, 0, (m - rs->cap_words) * 8); rs->cap_words = m; }
	r.off = rs->nbases; r.len = len; r.name = malloc(name_len + 1); memcpy(r.name, name, name_len); r.name[name_len] = 0;
	for(i=0;i<len;i++){
		u64 c;
		switch(seq[i]){ case 'A': case 'a': c = 0; break; case 'C': case 'c': c = 1; break; case 'G': case 'g': c = 2; break; case 'T': case 't': c = 3; break;
			default: c = lrand48() & 3; }                    /* dna.h:405: unseeded lrand48 on non-ACGT */
		bank_put(rs->bits, rs->nbases, c); rs->nbases ++;
	}
	vec_push(rs->reads, r);
}

typedef struct { char **files; int nfiles, fidx; FILE *fp; int is_proc; char *line; size_t cap; ssize_t n; int have_line; int type; } seqreader_t;

static int sr_open_next(seqreader_t *sr){
	while(sr->fidx < sr->nfiles){
		const char *fn = sr->files[sr->fidx ++]; size_t l = strlen(fn);
		if(!strcmp(fn, "-")){ sr->fp = stdin; sr->is_proc = 0; return 1; }
		if(l > 3 && !strcmp(fn + l - 3, ".gz")){ char *cmd = malloc(l + 20); sprintf(cmd, "gzip -dc %s", fn); sr->fp = popen(cmd, "r"); free(cmd); sr->is_proc = 1; if(sr->fp) return 1; continue; }
		sr->fp = fopen(fn, "r"); sr->is_proc = 0;
		if(sr->fp) return 1;
		fprintf(stderr, " -- Cannot open %s --\n", fn); exit(1);
	}
	return 0;
}
static int sr_getline(seqreader_t *sr){
	if(sr->have_line){ sr->have_line = 0; return 1; }
	while(1){
		if(sr->fp == NULL && !sr_open_next(sr)) return 0;
		sr->n = getline(&sr->line, &sr->cap, sr->fp);
		if(sr->n >= 0){
			while(sr->n && (sr->line[sr->n-1] == '\n')) sr->line[--sr->n] = 0;
			return 1;
		}
		if(sr->is_proc) pclose(sr->fp); else if(sr->fp != stdin) fclose(sr->fp);
		sr->fp = NULL;
	}
}
/* returns 1 and fills name/seq (growable) or 0 at end */
static int sr_next(seqreader_t *sr, u8v *name, u8v *seq){
	size_t i;
	if(sr->type == 0){
		while(sr_getline(sr)){
			if(sr->n == 0 || sr->line[0] == '#') continue;
			sr->type = sr->line[0] == '>'? 1 : (sr->line[0] == '@'? 2 : 3);
			sr->have_line = 1; break;
		}
		if(sr->type == 0) return 0;
	}
	vec_clear(*name); vec_clear(*seq);
	if(sr->type == 1){
		int flag = 0;
		while(sr_getline(sr)){
			if(sr->n && sr->line[0] == '>'){
				if(flag){ sr->have_line = 1; break; }
				flag = 1;
				for(i=1;i<(size_t)sr->n;i++){ char ch = sr->line[i]; if(ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n') break; }
				vec_reserve(*name, i); memcpy(name->a, sr->line + 1, i - 1); name->n = i - 1;
			} else if(flag){
				vec_reserve(*seq, seq->n + sr->n + 1); memcpy(seq->a + seq->n, sr->line, sr->n); seq->n += sr->n;
			}
		}
		return flag != 0;
	} else if(sr->type == 2){
		int flag = 0;
		while(flag != 4 && sr_getline(sr)){
			switch(flag){
				case 0: if(sr->line[0] != '@') break; flag = 1;
					for(i=1;i<(size_t)sr->n;i++){ char ch = sr->line[i]; if(ch == ' ' || ch == '\t' || ch == '\n') break; }
					vec_reserve(*name, i); memcpy(name->a, sr->line + 1, i - 1); name->n = i - 1; break;
				case 1: flag = 2; vec_reserve(*seq, sr->n + 1); memcpy(seq->a, sr->line, sr->n); seq->n = sr->n; break;
				case 2: if(sr->line[0] != '+') break; flag = 3; break;
				case 3: flag = 4; break;
			}
		}
		return flag == 4;
	}
	return 0;
}

static int gt_read_len_desc(const void *a, const void *b, void *ctx){ (void)ctx; return ((const read_t*)b)->len > ((const read_t*)a)->len; }

static void rs_load(readset_t *rs, char **files, int nfiles, int min_rdlen, int as_query){
	seqreader_t sr; u8v name, seq;
	memset(&sr, 0, sizeof(sr)); sr.files = files; sr.nfiles = nfiles;
	vec_init(name); vec_init(seq);
	while(sr_next(&sr, &name, &seq)){
		if((int)seq.n < min_rdlen) continue;
		rs_add_read(rs, (char*)name.a, (int)name.n, (char*)seq.a, (u32)seq.n);
		if(as_query) rs->n_qr ++; else rs->n_rd ++;
	}
	free(sr.line); vec_free(name); vec_free(seq);
}

static u32 rs_find(const readset_t *rs, u32 n, const char *name){   /* linear/bsearch-free name lookup via a tiny hash built lazily */
	static u32 *tab = NULL; static u32 tabsz = 0; static const readset_t *owner = NULL;
	u32 i, h;
	if(owner != rs){
		free(tab); tabsz = 16; while(tabsz < 2 * n + 1) tabsz <<= 1; tab = malloc(tabsz * 4); memset(tab, 0xFF, tabsz * 4); owner = rs;
		for(i=0;i<n;i++){ const char *s = rs->reads.a[i].name; h = 2166136261u; while(*s){ h = (h ^ (u8)*s++) * 16777619u; } h &= tabsz - 1; while(tab[h] != 0xFFFFFFFFU) h = (h + 1) & (tabsz - 1); tab[h] = i; }
	}
	{ const char *s = name; h = 2166136261u; while(*s){ h = (h ^ (u8)*s++) * 16777619u; } h &= tabsz - 1; }
	while(tab[h] != 0xFFFFFFFFU){ if(!strcmp(rs->reads.a[tab[h]].name, name)) return tab[h]; h = (h + 1) & (tabsz - 1); }
	return 0xFFFFFFFFU;
}

/* unpack a read to one base per byte, forward (dna.h:451) or reverse-complement (dna.h:466) */
static void rs_unpack(const readset_t *rs, u32 id, int rev, u8 *dst){
	const read_t *r = &rs->reads.a[id]; u32 i;
	if(!rev) for(i=0;i<r->len;i++) dst[i] = bank_get(rs->bits, r->off + i);
	else for(i=0;i<r->len;i++) dst[i] = (~bank_get(rs->bits, r->off + r->len - 1 - i)) & 3;
}

/* ------------------------------------------------------------------ homopolymer-compressed canonical k-mer scan
 * Shared by index build and query (wtzmo.c:246-271, 295-309, 457-481) and by the z-mer stages
 * (hzm_aln.h:83-99, 189-205).  Emits (canonical mer, dir, off, len) for every position after the
 * first k kept bases, skipping palindromes.  off = uncompressed position of the k-mer's first kept
 * base, len = span up to and including the current kept base, capped at 0xFFFF. */
typedef struct { u64 mer; u32 off, len; u8 dir; } kmer_hit_t;
typedef VEC(kmer_hit_t) kmerv;
static void scan_kmers(const u8 *seq, u32 len, int k, int hp, kmerv *out){
	u64 kmer = 0, kmask = 0xFFFFFFFFFFFFFFFFULL >> ((32 - k) << 1), krev;
	u32 j, kept = 0; u8 b = 4;
	u32 *hzoff = malloc((len + 1) * sizeof(u32));
	vec_clear(*out);
	for(j=0;j<len;j++){
		u8 c = seq[j];
		kmer_hit_t h;
		if(hp && c == b) continue;
		b = c; hzoff[kept ++] = j;
		kmer = ((kmer << 2) | b) & kmask;
		if(kept < (u32)k) continue;
		krev = kmer_revcomp(kmer, k);
		if(krev == kmer) continue;
		h.dir = krev > kmer? 0 : 1; h.mer = krev > kmer? kmer : krev;
		h.off = hzoff[kept - k]; h.len = (j + 1 - h.off > 0xFFFF)? 0xFFFF : j + 1 - h.off;
		vec_push(*out, h);
	}
	free(hzoff);
}

/* ------------------------------------------------------------------ global k-mer index (wtzmo.c:227-430)
 * Observable content only: for each sampled canonical k-mer its posting list of (rd_id<<1|dir)
 * sorted ascending, and the filter flag (count<=1 or count>K).  Sampling: keep iff
 * jenkins32(low32(mer)) % (1024*S) < 1024 (wtzmo.c:270-271). */
typedef struct { u64 mer; u64 off; u32 cnt; u8 flt; } kent_t;
typedef struct { kent_t *ents; size_t n_ent; u32 *post; u64 n_post; u32 K; } kindex_t;

static int cmp_u64pair(const void *a, const void *b){
	const u64 *x = a, *y = b;
	if(x[0] != y[0]) return x[0] < y[0]? -1 : 1;
	if(x[1] != y[1]) return x[1] < y[1]? -1 : 1;
	return 0;
}
static inline int kmer_sampled(u64 mer, int ksave){ return jenkins32((u32)mer) % (1024u * (u32)ksave) < 1024u; }

static void kindex_free(kindex_t *ix){ free(ix->ents); free(ix->post); memset(ix, 0, sizeof(*ix)); }

/* kcut_io: in = user -K (0/1 = auto); the auto value is computed on the first call only (wtzmo.c:380-393) */
static void kindex_build(kindex_t *ix, const readset_t *rs, u32 beg, u32 end, const zparams_t *par, u32 *kcut_io){
	VEC(u64) pairs; kmerv km; u8v buf; u32 id; size_t i, j, ne = 0; u64 ktot = 0, off = 0;
	vec_init(pairs); vec_init(km); vec_init(buf);
	if(end > rs->n_rd) end = rs->n_rd; /* reference reads out of bounds when n_rd % n_idx != 0 (wtzmo.c:1283) */
	for(id=beg;id<end;id++){
		u32 len = rs->reads.a[id].len;
		vec_reserve(buf, len + 1); rs_unpack(rs, id, 0, buf.a);
		scan_kmers(buf.a, len, par->ksize, par->hk, &km);
		for(i=0;i<km.n;i++){
			if(!kmer_sampled(km.a[i].mer, par->ksave)) continue;
			vec_push(pairs, km.a[i].mer); vec_push(pairs, ((u64)id << 1) | km.a[i].dir);
		}
	}
	qsort(pairs.a, pairs.n / 2, 16, cmp_u64pair);
	for(i=0;i<pairs.n/2;i=j){ for(j=i+1;j<pairs.n/2&&pairs.a[2*j]==pairs.a[2*i];j++); ne ++; }
	ix->ents = malloc((ne + 1) * sizeof(kent_t)); ix->n_ent = ne; ne = 0;
	for(i=0;i<pairs.n/2;i=j){
		for(j=i+1;j<pairs.n/2&&pairs.a[2*j]==pairs.a[2*i];j++);
		ix->ents[ne].mer = pairs.a[2*i]; ix->ents[ne].off = i; ix->ents[ne].cnt = (u32)(j - i); ix->ents[ne].flt = 0;
		ktot += (j - i > 0xFFFF)? 0xFFFF : (j - i);          /* 16-bit saturating counter, wtzmo.c:276 */
		ne ++;
	}
	if(*kcut_io < 2){
		u32 kavg = (u32)(ktot / (ne + 1));
		if(kavg < 20) kavg = 20;
		*kcut_io = kavg * 5;
	}
	ix->K = *kcut_io;
	/* keep postings only for k-mers with 1 < count <= K (wtzmo.c:401-406) */
	ix->post = malloc((pairs.n / 2 + 1) * sizeof(u32));
	for(i=0;i<ne;i++){
		kent_t *e = &ix->ents[i]; u32 c = e->cnt > 0xFFFF? 0xFFFF : e->cnt;
		if(c > ix->K || c <= 1 || e->cnt > 0xFFFF){ e->flt = 1; e->cnt = 0; e->off = off; continue; }
		for(j=0;j<e->cnt;j++) ix->post[off + j] = (u32)pairs.a[2 * (e->off + j) + 1];
		e->off = off; off += e->cnt;
	}
	ix->n_post = off;
	vec_free(pairs); vec_free(km); vec_free(buf);
}
static const kent_t* kindex_find(const kindex_t *ix, u64 mer){
	size_t lo = 0, hi = ix->n_ent;
	while(lo < hi){ size_t mid = (lo + hi) / 2; if(ix->ents[mid].mer < mer) lo = mid + 1; else hi = mid; }
	return (lo < ix->n_ent && ix->ents[lo].mer == mer)? &ix->ents[lo] : NULL;
}

/* ------------------------------------------------------------------ candidate events (wtzmo.c:433-562)
 * For query read q: every (target,strand) that shares a non-filtered sampled k-mer, in ascending
 * (target<<1|strand) order, with ol = "union length" of the query spans visited in ascending query
 * offset: ol += off>=lst ? len : off+len-lst (uint32 wrap kept), lst = off+len.  Self hits and
 * targets longer than (uint32)(1.2*len(q)) are skipped. */
typedef struct { u32 tkey, ol, cnt; } cand_event_t;
typedef VEC(cand_event_t) eventv;
static int cmp_u64(const void *a, const void *b){ u64 x = *(const u64*)a, y = *(const u64*)b; return x < y? -1 : (x > y); }

static void candidate_events(const readset_t *rs, const kindex_t *ix, u32 qid, const zparams_t *par, eventv *ev){
	u32 qlen = rs->reads.a[qid].len, up = (u32)(qlen * 1.2), i; size_t a, b;
	u8v buf; kmerv km; u64v keys;
	vec_init(buf); vec_init(km); vec_init(keys); vec_clear(*ev);
	vec_reserve(buf, qlen + 1); rs_unpack(rs, qid, 0, buf.a);
	scan_kmers(buf.a, qlen, par->ksize, par->hk, &km);
	for(i=0;i<km.n;i++){
		const kent_t *e; u32 c;
		if(!kmer_sampled(km.a[i].mer, par->ksave)) continue;
		e = kindex_find(ix, km.a[i].mer);
		if(e == NULL || e->flt) continue;
		for(c=0;c<e->cnt;c++){
			u32 tk = ix->post[e->off + c], tid = tk >> 1;
			if(tid == qid) continue;
			if(rs->reads.a[tid].len > up) continue;
			vec_push(keys, ((u64)tk << 32) | i);      /* k-mer ordinal i is monotone in query offset */
		}
	}
	qsort(keys.a, keys.n, 8, cmp_u64);
	for(a=0;a<keys.n;a=b){
		cand_event_t e; u32 lst = 0;
		e.tkey = (u32)(keys.a[a] >> 32); e.ol = 0; e.cnt = 0;
		for(b=a;b<keys.n&&(u32)(keys.a[b]>>32)==e.tkey;b++){
			const kmer_hit_t *h = &km.a[(u32)keys.a[b]];
			if(h->off >= lst) e.ol += h->len; else e.ol += h->off + h->len - lst;
			lst = h->off + h->len; e.cnt ++;
		}
		vec_push(*ev, e);
	}
	vec_free(buf); vec_free(km); vec_free(keys);
}

/* ------------------------------------------------------------------ heap macros behaviour (list.h:78-144) on u64 keyed by low 32 bits */
static inline int cand_cmp(u64 a, u64 b){ u32 x = (u32)a, y = (u32)b; return x > y? 1 : (x < y? -1 : 0); }
static void cheap_push(u64v *h, u64 v){
	size_t i = h->n, j;
	vec_push(*h, v);
	while(i){ j = (i - 1) >> 1; if(cand_cmp(h->a[i], h->a[j]) >= 0) break; { u64 t = h->a[i]; h->a[i] = h->a[j]; h->a[j] = t; } i = j; }
}
static void cheap_replace0(u64v *h, u64 v){
	size_t idx = 0, sw;
	h->a[0] = v;
	while((idx << 1) + 1 < h->n){
		sw = idx;
		if(cand_cmp(h->a[sw], h->a[(idx << 1) + 1]) > 0) sw = (idx << 1) + 1;
		if((idx << 1) + 2 < h->n && cand_cmp(h->a[sw], h->a[(idx << 1) + 2]) > 0) sw = (idx << 1) + 2;
		if(sw == idx) break;
		{ u64 t = h->a[idx]; h->a[idx] = h->a[sw]; h->a[sw] = t; }
		idx = sw;
	}
}

/* Turn the ordered event stream of one index partition into the candidate array (wtzmo.c:494-571).
 * `cands` may already hold entries carried over from earlier index partitions (-G). All quirks kept:
 * the heap-full test looks at the CURRENT ol but inserts the PREVIOUS pending candidate; the final
 * flush compares against ol==0; an empty stream pushes the sentinel 0xFFFFFFFF00000000. */
static void candidates_from_events(const eventv *ev, const zparams_t *par, u64v *cands){
	u64 x1 = 0xFFFFFFFF00000000ULL, x2; size_t i; u32 ol = 0;
	for(i=0;i<ev->n;i++){
		ol = ev->a[i].ol;
		if(ol >= (u32)par->kovl){
			x2 = ((u64)(ev->a[i].tkey >> 1) << 32) | ol;
			if((x1 >> 32) == (x2 >> 32)){ x1 = (u32)x1 > (u32)x2? x1 : x2; }
			else if(x1 == 0xFFFFFFFF00000000ULL){ x1 = x2; }
			else {
				if(cands->n >= (size_t)par->ncand){ if((u32)cands->a[0] < ol) cheap_replace0(cands, x1); }
				else cheap_push(cands, x1);
				x1 = x2;
			}
		}
	}
	ol = 0; /* the loop leaves ol reset to 0 after the last group (wtzmo.c:550) */
	if(cands->n >= (size_t)par->ncand){ if((u32)cands->a[0] < ol) cheap_replace0(cands, x1); }
	else cheap_push(cands, x1);
}

/* ------------------------------------------------------------------ z-mer stage (hzm_aln.h:70-224)
 * z-index of read q: all canonical hp-z-mers sorted by (mer,off); distinct mers with 0<count<Z get a
 * slot (slot id = rank among indexed mers).  Matching read c: positions of c in order; a position
 * whose mer is indexed and whose slot has been hit < Z times (uint8 counter) emits one pair per
 * occurrence in q whose span lengths differ by <= kvar.  Opposite-strand pairs store
 * off2 = len_c - (off+len), i.e. coordinates on revcomp(c). */
typedef struct { u32 off1, off2; u16 len1, len2; u8 dir1, dir2; u32 gid; } zpair_t;
typedef VEC(zpair_t) zpairv;
typedef struct { u32 mer, off, cnt; } zslot_t;
typedef struct { kmerv seeds; VEC(zslot_t) slots; } zindex_t;

static int cmp_zseed(const void *a, const void *b){
	const kmer_hit_t *x = a, *y = b;
	if(x->mer != y->mer) return x->mer < y->mer? -1 : 1;
	return x->off < y->off? -1 : (x->off > y->off);
}
static void zindex_build(zindex_t *zi, const u8 *seq, u32 len, const zparams_t *par){
	size_t i, j;
	scan_kmers(seq, len, par->zsize, par->hz, &zi->seeds);
	qsort(zi->seeds.a, zi->seeds.n, sizeof(kmer_hit_t), cmp_zseed);
	vec_clear(zi->slots);
	for(i=0;i<zi->seeds.n;i=j){
		zslot_t s;
		for(j=i+1;j<zi->seeds.n&&zi->seeds.a[j].mer==zi->seeds.a[i].mer;j++);
		s.mer = (u32)zi->seeds.a[i].mer; s.off = (u32)i; s.cnt = (u32)(j - i);
		if(s.cnt < (u32)par->zcut) vec_push(zi->slots, s);
	}
}
static void zmatch(const zindex_t *zi, const u8 *cseq, u32 clen, const zparams_t *par, zpairv *out){
	kmerv ck; size_t i; u32 k; u8 *kcnts = calloc(zi->slots.n + 1, 1);
	vec_init(ck); vec_clear(*out);
	scan_kmers(cseq, clen, par->zsize, par->hz, &ck);
	for(i=0;i<ck.n;i++){
		const kmer_hit_t *p2 = &ck.a[i]; size_t lo = 0, hi = zi->slots.n; const zslot_t *s;
		while(lo < hi){ size_t mid = (lo + hi) / 2; if(zi->slots.a[mid].mer < (u32)p2->mer) lo = mid + 1; else hi = mid; }
		if(lo >= zi->slots.n || zi->slots.a[lo].mer != (u32)p2->mer) continue;
		s = &zi->slots.a[lo];
		if(kcnts[lo] >= (u32)par->zcut) continue;
		kcnts[lo] ++;
		for(k=0;k<s->cnt;k++){
			const kmer_hit_t *p1 = &zi->seeds.a[s->off + k]; zpair_t z;
			if(idiff(p1->len, p2->len) > (u32)par->kvar) continue;
			z.dir1 = p1->dir; z.off1 = p1->off; z.dir2 = p2->dir; z.len1 = (u16)p1->len; z.len2 = (u16)p2->len; z.gid = 0;
			z.off2 = (p1->dir ^ p2->dir)? clen - (p2->off + p2->len) : p2->off;
			vec_push(*out, z);
		}
	}
	free(kcnts); vec_free(ck);
}
static int gt_zpair_off12(const void *a, const void *b, void *ctx){   /* process_hzmps, hzm_aln.h:1184-1186 */
	const zpair_t *x = a, *y = b; (void)ctx;
	return (((i64)x->off1 << 32) | x->off2) > (((i64)y->off1 << 32) | y->off2);
}

/* ------------------------------------------------------------------ seed windows (hzm_aln.h:316-656) */
typedef struct { u32 pb2; u32 ovl; u8 dir, closed; int beg[2], end[2]; u32 anc[2]; } win_t;
typedef VEC(win_t) winv;
#define WIN_OVL_MASK 0x1FFFFFFFU   /* wt_seed_t.ovl is a 29-bit field (hzm_aln.h:64) */

/* in-place quickselect returning the element of rank size/2 (upper median), hzm_aln.h:316-343 */
static i32 median_select(i32 *rs, i32 size){
	i32 i, j, key, mid, beg = 0, end = size - 1, tmp;
	if(size == 0) return 0;
	while(beg < end){
		mid = beg + (end - beg) / 2;
		if(rs[beg] > rs[mid]){ tmp = rs[beg]; rs[beg] = rs[mid]; rs[mid] = tmp; }
		if(rs[mid] > rs[end]){
			tmp = rs[end]; rs[end] = rs[mid]; rs[mid] = tmp;
			if(rs[beg] > rs[mid]){ tmp = rs[beg]; rs[beg] = rs[mid]; rs[mid] = tmp; }
		}
		key = rs[mid]; i = beg + 1; j = end - 1;
		while(1){
			while(key > rs[i]) i ++;
			while(rs[j] > key) j --;
			if(i < j){ tmp = rs[i]; rs[i] = rs[j]; rs[j] = tmp; i ++; j --; } else break;
		}
		if(i == j){ i ++; j --; }
		if(i <= size / 2) beg = i; else end = j;
	}
	return rs[size / 2];
}

static int gt_idx_off2(const void *a, const void *b, void *ctx){ const zpair_t *rs = ctx; return rs[*(const u32*)a].off2 > rs[*(const u32*)b].off2; }
static int gt_zpair_off1(const void *a, const void *b, void *ctx){ (void)ctx; return ((const zpair_t*)a)->off1 > ((const zpair_t*)b)->off1; }

#define KWIN_MAX_OFFSET_DEV 50
/* hzm_aln.h:410-578 (fast-chaining branch, always on for wtzmo: wtzmo.c:1540).  Looks at the
 * strand-`dir` matches in rs[beg,end): sub-windows along c with >= zovl covered bases, median
 * diagonal, anchors within +-50 of it sorted by off1.  Appends windows/anchors, returns #windows. */
static u32 windows_in_span(const zpair_t *rs, int dir, u32 beg, u32 end, int bound, winv *wins, zpairv *anchors, const zparams_t *par){
	typedef struct { u32 b, e, ovl; } wreg_t;
	u32 zsize = par->zsize, kwin = par->kwin, zovl = par->zovl;
	u32 i, j, n = 0, n2 = 0, ol, ol2, lst, ret = 0; u32 *ts; i32 *as; wreg_t *ws;
	while(beg < end){
		const zpair_t *p = &rs[beg];
		if((p->dir1 ^ p->dir2 ^ dir) || (int)p->off1 < bound) beg ++; else break;
	}
	for(i=beg;i<end;i++) if(!(rs[i].dir1 ^ rs[i].dir2 ^ dir)) n ++;
	if(n * zsize < zovl) return 0;
	ts = malloc((n + 1) * sizeof(u32)); as = malloc((n + 1) * sizeof(i32)); ws = malloc((n + 1) * sizeof(wreg_t));
	n = 0;
	for(i=beg;i<end;i++) if(!(rs[i].dir1 ^ rs[i].dir2 ^ dir)) ts[n++] = i;
	ref_sort(ts, n, sizeof(u32), gt_idx_off2, (void*)rs);
	ol = 0; lst = 0;
	for(i=j=0;i<n;i++){
		const zpair_t *p = &rs[ts[i]];
		while((u32)p->off2 + p->len2 > rs[ts[j]].off2 + kwin && j + 1 < n){
			const zpair_t *p0 = &rs[ts[j++]], *p1 = &rs[ts[j]];
			u32 s = p1->off2, t = p0->off2 + p0->len2;
			ol2 = s < t? t - s : 0;
			ol = ol + ol2 - p0->len2;
		}
		ol += (p->off2 > lst)? p->len2 : p->off2 + p->len2 - lst;
		lst = p->off2 + p->len2;
		if(ol >= zovl){
			if(n2 && ( rs[ts[i]].off2 <= rs[ts[ws[n2-1].e]].off2 + kwin / 3 || rs[ts[j]].off2 <= rs[ts[ws[n2-1].b]].off2 + kwin / 3 )){
				if(ol > ws[n2-1].ovl){ ws[n2-1].b = j; ws[n2-1].e = i; ws[n2-1].ovl = ol; }
			} else { ws[n2].b = j; ws[n2].e = i; ws[n2].ovl = ol; n2 ++; }
		}
	}
	for(i=0;i<n2;i++){
		size_t size = anchors->n; i32 offset, offn = 0; win_t *w; win_t W0;
		for(j=ws[i].b;j<=ws[i].e;j++) as[offn++] = (i32)rs[ts[j]].off1 - (i32)rs[ts[j]].off2;
		offset = median_select(as, offn);
		ol = lst = 0;
		for(j=ws[i].b;j<=ws[i].e;j++){
			const zpair_t *p = &rs[ts[j]]; i32 off = (i32)p->off1 - (i32)p->off2;
			if(off < offset - KWIN_MAX_OFFSET_DEV || off > offset + KWIN_MAX_OFFSET_DEV) continue;
			vec_push(*anchors, *p);
			ol += (p->off2 > lst)? p->len2 : p->off2 + p->len2 - lst;
			lst = p->off2 + p->len2;
		}
		if(anchors->n == size) continue;
		ref_sort(anchors->a + size, anchors->n - size, sizeof(zpair_t), gt_zpair_off1, NULL);
		memset(&W0, 0, sizeof(W0));
		W0.closed = 0; W0.dir = dir; W0.anc[0] = (u32)size; W0.beg[0] = W0.beg[1] = 0x7FFFFFFF; W0.end[0] = W0.end[1] = 0;
		ol = lst = 0;
		for(j=(u32)size;j<anchors->n;j++){
			const zpair_t *p = &anchors->a[j];
			ol += (p->off1 > lst)? p->len1 : p->off1 + p->len1 - lst;
			lst = p->off1 + p->len1;
			if((int)p->off1 < W0.beg[0]) W0.beg[0] = p->off1;
			if((int)(p->off1 + p->len1) > W0.end[0]) W0.end[0] = p->off1 + p->len1;
			if((int)p->off2 < W0.beg[1]) W0.beg[1] = p->off2;
			if((int)(p->off2 + p->len2) > W0.end[1]) W0.end[1] = p->off2 + p->len2;
		}
		if(ol * 2 < zovl){ anchors->n = size; continue; }
		if(ret){
			w = &wins->a[wins->n - 1];
			if(W0.end[1] <= (int)(w->end[1] + kwin / 3) && ol <= w->ovl){ anchors->n = size; continue; }
		}
		ret ++;
		W0.ovl = ol & WIN_OVL_MASK; W0.anc[1] = (u32)anchors->n;
		vec_push(*wins, W0);
	}
	free(ts); free(as); free(ws);
	return ret;
}

/* hzm_aln.h:580-656: slide a kwin-wide window along q over the (off1,off2)-sorted match list; where
 * the running covered length reaches zovl, look for windows; advance by kstep otherwise.  The
 * running length is uint32, is decremented with entries of BOTH strands when the start advances and
 * incremented only for strand `dir`; the element that closes a window is not added. */
static u32 pair_windows_strand(const zpair_t *rs, u32 n, int dir, winv *wins, zpairv *anchors, const zparams_t *par){
	u32 kwin = par->kwin, kstep = par->kstep, zovl = par->zovl;
	u32 i, j, a, nw, ol = 0, ol2, lst = 0, wlst = 0, s, t, ret = 0;
	u32 p0_off1, p0_len1, p_off1, p_len1;
	for(j=0;j<n;j++) if(!(rs[j].dir1 ^ rs[j].dir2 ^ dir)) break;
	if(j == n) return 0;
	p0_off1 = rs[j].off1; p0_len1 = rs[j].len1;
	for(i=j;i<=n;i++){
		if(i < n){
			if(rs[i].dir1 ^ rs[i].dir2 ^ dir) continue;
			p_off1 = rs[i].off1; p_len1 = rs[i].len1;
		} else { p_off1 = 0x1FFFFFU; p_len1 = 0x3FFU; }
		if(p_off1 > p0_off1 + kwin){
			if(ol >= zovl){
				if((nw = windows_in_span(rs, dir, j, i, (int)wlst, wins, anchors, par))){
					for(a=0;a<nw;a++){ int e0 = wins->a[wins->n + a - nw].end[0] + 20; if((int)wlst < e0) wlst = e0; }
					ret += nw;
					p0_off1 = p_off1; p0_len1 = p_len1;
					ol = p_len1; lst = p_off1 + p_len1; j = i;
				} else if(i < n){
					u32 nxt = p0_off1 + kstep;
					while(p0_off1 < nxt && j < i){
						const zpair_t *p1 = &rs[++j];
						s = imax(p0_off1, p1->off1); t = imin(p0_off1 + p0_len1, (u32)p1->off1 + p1->len1);
						ol2 = s < t? t - s : 0;
						ol = ol + ol2 - p0_len1;
						p0_off1 = p1->off1; p0_len1 = p1->len1;
					}
				}
			}
			if(p_off1 == 0x1FFFFFU) break;   /* sentinel test is by value (hzm_aln.h:637) */
			while(p_off1 > p0_off1 + kwin){
				const zpair_t *p1 = &rs[++j];
				s = imax(p0_off1, p1->off1); t = imin(p0_off1 + p0_len1, (u32)p1->off1 + p1->len1);
				ol2 = s < t? t - s : 0;
				ol = ol + ol2 - p0_len1;
				p0_off1 = p1->off1; p0_len1 = p1->len1;
			}
		} else {
			if(p_off1 >= lst) ol += p_len1;
			else if((int)(p_off1 + p_len1) > (int)lst) ol += p_off1 + p_len1 - lst;
			else continue;
			lst = p_off1 + p_len1;
		}
	}
	return ret;
}

/* hzm_aln.h:658-713: O(n^2) chain over the windows of one strand; returns the summed q-span of the
 * best chain and leaves closed=0 exactly on its members. */
static int chain_windows(win_t *w, u32 n, int W){
	typedef struct { int weight, bt; } node_t;
	node_t *nodes = malloc((n + 1) * sizeof(node_t)); u32 i, j; int mw = -1000000, bt = -1, band;
	float band_penalty = 0.05;
	for(i=0;i<n;i++){ nodes[i].weight = 0; nodes[i].bt = -1; }
	for(i=0;i<n;i++){
		w[i].closed = 1;
		nodes[i].weight += w[i].ovl;
		if(nodes[i].weight > mw){ mw = nodes[i].weight; bt = i; }
		for(j=i+1;j<n;j++){
			if(w[j].beg[1] < w[i].end[1]) continue;
			if(w[j].beg[0] < w[i].end[0]) continue;
			if(w[j].beg[0] - w[i].end[0] > W && w[j].beg[1] - w[i].end[1] > W) break;
			band = idiff(w[j].beg[0] - w[i].end[0], w[j].beg[1] - w[i].end[1]);
			if(band > W) continue;
			band = band * band_penalty;
			if(nodes[j].weight < nodes[i].weight - band){ nodes[j].weight = nodes[i].weight - band; nodes[j].bt = i; }
		}
	}
	mw = 0;
	while(bt >= 0){ w[bt].closed = 0; mw += w[bt].end[0] - w[bt].beg[0]; bt = nodes[bt].bt; }
	free(nodes);
	return mw;
}

/* Pure per-pair seeding result (wtzmo.c:849-914 without the windeps side effect) */
typedef struct {
	u32 n_zpair;            /* cache->size */
	int ovl[2];             /* chain weight per strand (0 when no windows) */
	winv wins[2];           /* kept (closed==0) windows per strand, anchors index into anc[strand] */
	zpairv anc[2];
} pair_seed_t;
static void pair_seed_init(pair_seed_t *ps){ memset(ps, 0, sizeof(*ps)); }
static void pair_seed_free(pair_seed_t *ps){ int d; for(d=0;d<2;d++){ vec_free(ps->wins[d]); vec_free(ps->anc[d]); } }

static void pair_seed_compute(pair_seed_t *ps, zpairv *cache, const zparams_t *par){
	int dir; u32 j;
	ps->n_zpair = (u32)cache->n; ps->ovl[0] = ps->ovl[1] = 0;
	for(dir=0;dir<2;dir++){ vec_clear(ps->wins[dir]); vec_clear(ps->anc[dir]); }
	if(cache->n * par->zsize < (u32)par->ztot) return;
	ref_sort(cache->a, cache->n, sizeof(zpair_t), gt_zpair_off12, NULL);
	for(dir=0;dir<2;dir++){
		winv w2; zpairv a2; vec_init(w2); vec_init(a2);
		if(pair_windows_strand(cache->a, (u32)cache->n, dir, &w2, &a2, par)){
			ps->ovl[dir] = chain_windows(w2.a, (u32)w2.n, par->W);
			if((u32)ps->ovl[dir] >= (u32)par->ztot){
				for(j=0;j<w2.n;j++){
					win_t w = w2.a[j]; u32 k, na;
					if(w.closed) continue;
					na = w.anc[1] - w.anc[0];
					for(k=0;k<na;k++) vec_push(ps->anc[dir], a2.a[w.anc[0] + k]);
					w.anc[1] = (u32)ps->anc[dir].n; w.anc[0] = w.anc[1] - na;
					vec_push(ps->wins[dir], w);
				}
			}
		}
		vec_free(w2); vec_free(a2);
	}
}

/* ------------------------------------------------------------------ per-window anchored alignment (hzm_aln.h:1247-1302)
 * Walk the window's anchors in off1 order; skip anchors starting before the current end on either
 * read; bridge to the next anchor with the fixed-band extension seeded with the running score; if
 * the extension stopped short, pad with D then I (counted in del/ins/aln, NOT in score); then align
 * the anchor itself run-length-wise.  pb1 = q (target of the DP), pb2 = c (query of the DP). */
typedef struct { aln_t x; u32 cig_off, cig_len; } alnreg_t;
static aln_t window_align(const u8 *pb1, const u8 *pb2, const win_t *w, const zpair_t *anchors, u32v *cigar, const zparams_t *par){
	aln_t x = ALN_NULL, y; u32 i; u32v tmp; vec_init(tmp);
	for(i=w->anc[0];i<w->anc[1];i++){
		const zpair_t *p = &anchors[i];
		if(x.aln == 0){ x.tb = x.te = p->off1; x.qb = x.qe = p->off2; }
		if((int)p->off1 < x.te) continue;
		if((int)p->off2 < x.qe) continue;
		y = banded_extend(0, p->off2 - x.qe, pb2 + x.qe, p->off1 - x.te, pb1 + x.te, 1, x.score, par->w, par->M, par->X, par->O, par->O, par->E, par->T, &tmp);
		x.score = y.score;
		x.aln += y.aln; x.mat += y.mat; x.mis += y.mis; x.ins += y.ins; x.del += y.del;
		x.te += y.te; x.qe += y.qe;
		if(x.te < (int)p->off1){ x.del += p->off1 - x.te; x.aln += p->off1 - x.te; cig_push(&tmp, 2, p->off1 - x.te); x.te = p->off1; }
		if(x.qe < (int)p->off2){ x.ins += p->off2 - x.qe; x.aln += p->off2 - x.qe; cig_push(&tmp, 1, p->off2 - x.qe); x.qe = p->off2; }
		cig_append(cigar, tmp.a, tmp.n);
		vec_clear(tmp);
		y = runlen_align(pb1 + p->off1, p->len1, pb2 + p->off2, p->len2, par->M, par->O, par->O, par->E, &tmp);
		if(y.aln == 0) break;       /* "should never happen": window truncated here */
		x.score += y.score;
		x.aln += y.aln; x.mat += y.mat; x.mis += y.mis; x.ins += y.ins; x.del += y.del;
		x.te += y.te; x.qe += y.qe;
		cig_append(cigar, tmp.a, tmp.n);
	}
	vec_free(tmp);
	return x;
}

/* hzm_aln.h:1345-1486: [left end extension] + reg0 + sum([banded-global gap] + reg_i) + [right end
 * extension].  End extensions use the shifting-band DP with exact band ew (W passed negative); the
 * left one runs backwards (strand -1) with init = score + 100*M and subtracts it afterwards. */
static aln_t stitch_regs(int len1, int len2, const alnreg_t *regs, u32 nreg, const u8 *pb1, const u8 *pb2, const u32 *cig_cache, u32v *cigar, const zparams_t *par){
	aln_t x = ALN_NULL, y; u32 i; u32v tmp; int w, max_gap, init_score = 100 * par->M, score;
	int M = par->M, X = par->X, I = par->O, D = par->O, E = par->E, T = par->T, ew = par->ew;
	const int esti[2] = {0, len1};
	vec_init(tmp); vec_clear(*cigar);
	if(nreg == 0) return x;
	x = regs[0].x;
	if(x.qb && x.tb){
		w = ew;
		max_gap = ((imin(x.qb, x.tb) * M + x.score + init_score + (-T)) + (I < D? D : I)) / (-E) + 1;
		if(max_gap < w) max_gap = w;
		while(1){
			y = banded_extend(1, x.qb, pb2 + x.qb - 1, x.tb, pb1 + x.tb - 1, -1, x.score + init_score, -w, M, X, I, D, E, T, &tmp);
			if(y.qe == x.qb || y.te == x.tb) break;
			if(x.tb - y.te <= esti[0]) break;
			if(w >= ew || w >= max_gap) break;
			w <<= 1;
		}
		x.score = y.score - init_score;
		x.aln += y.aln; x.mat += y.mat; x.mis += y.mis; x.ins += y.ins; x.del += y.del;
		x.qb -= y.qe; x.tb -= y.te;
		cig_reverse(&tmp);
		cig_append(cigar, tmp.a, tmp.n);
	}
	cig_append(cigar, cig_cache + regs[0].cig_off, regs[0].cig_len);
	for(i=1;i<nreg;i++){
		const alnreg_t *r1 = &regs[i-1], *r2 = &regs[i];
		const u8 *q = pb2 + r1->x.qe, *t = pb1 + r1->x.te;
		int gq = r2->x.qb - r1->x.qe, gt = r2->x.tb - r1->x.te, x1 = 0, x2 = 0; size_t k; int jj;
		w = par->w;
		while(1){
			if(w < idiff(gq, gt)){ w <<= 1; continue; }
			score = banded_global(gq, q, gt, t, M, X, -I, -E, -D, -E, w, &tmp);
			if(score < 0 && w < par->W && w < imax(gq, gt)) w <<= 1; else break;
		}
		x.score += score; x.qe = r2->x.qb; x.te = r2->x.tb;
		for(k=0;k<tmp.n;k++){
			int op = tmp.a[k] & 0xF, len = tmp.a[k] >> 4;
			x.aln += len;
			switch(op){
				case 0: for(jj=0;jj<len;jj++){ if(q[x1 + jj] == t[x2 + jj]) x.mat ++; else x.mis ++; } x1 += len; x2 += len; break;
				case 1: x1 += len; x.ins += len; break;
				case 2: x2 += len; x.del += len; break;
			}
		}
		cig_append(cigar, tmp.a, tmp.n);
		x.score += r2->x.score; x.aln += r2->x.aln; x.mat += r2->x.mat; x.mis += r2->x.mis; x.ins += r2->x.ins; x.del += r2->x.del;
		x.qe = r2->x.qe; x.te = r2->x.te;
		cig_append(cigar, cig_cache + r2->cig_off, r2->cig_len);
	}
	if(x.te < len1 && x.qe < len2){
		w = ew;
		max_gap = ((imin(len2 - x.qe, len1 - x.te) * M + x.score + (-T)) + (I < D? D : I)) / (-E) + 1;
		if(max_gap < w) max_gap = w;
		while(1){
			y = banded_extend(1, len2 - x.qe, pb2 + x.qe, len1 - x.te, pb1 + x.te, 1, x.score, -w, M, X, I, D, E, T, &tmp);
			if(y.qe == len2 - x.qe || y.te == len1 - x.te) break;
			if(x.te + y.te >= esti[1]) break;
			if(w >= ew || w >= max_gap) break;
			w <<= 1;
		}
		x.score = y.score;
		x.aln += y.aln; x.mat += y.mat; x.mis += y.mis; x.ins += y.ins; x.del += y.del;
		x.qe += y.qe; x.te += y.te;
		cig_append(cigar, tmp.a, tmp.n);
	}
	vec_free(tmp);
	return x;
}

/* ------------------------------------------------------------------ -n: refinement of the stitched alignment
 * Restates kswx_refine_alignment (kswx.h:483-659) as wtzmo calls it (wtzmo.c:1031-1034): a global affine re-alignment of
 * query[qb, qe) (= c on its strand) against target[tb, te) (= q) inside a per-row band around the path of the input CIGAR.
 *   half width of row r: W, or W + L on the rows of an insertion run of length L; rows within L-1 of an indel run of length
 *   L get (L - distance) more (a deletion run sits between two rows); a deletion's own "+= len" is overwritten and has no
 *   effect (kswx.h:529-543).  Band of row r around the path column tx: [tx - zw, tx + 1 + zw) clipped to [0, tl), then the
 *   starts are made non-decreasing from the top and the ends non-increasing from the bottom (kswx.h:590-601).
 *   DP: H(-1,-1) = 0, every other out-of-band H / E and the row-initial F read as -10000, e/f opened from the diagonal move,
 *   same precedence and traceback flags as the extension DP; score = H(ql-1, tl-1); the walk starts there in state M. */
static aln_t refine_alignment(const u8 *query, int qb, const u8 *target, int tb, int W, const zparams_t *par, u32v *cigar){
	const int M = par->M, X = par->X, I = par->O, D = par->O, E = par->E;     /* wtzmo passes O for both gap opens */
	aln_t y = ALN_NULL; u32v in; size_t k; int ql = 0, tl = 0, qx, tx, i, j, wmax = 1;
	int *zw, *zb, *ze, *hrow, *erow; u8 *z;
	vec_init(in); vec_reserve(in, cigar->n + 1); for(k=0;k<cigar->n;k++) in.a[k] = cigar->a[k]; in.n = cigar->n;
	cigar->n = 0;
	for(k=0;k<in.n;k++){ const u32 op = in.a[k] & 0xF, len = in.a[k] >> 4; if(op == 0){ ql += len; tl += len; } else if(op == 1) ql += len; else tl += len; }
	if(ql == 0 || tl == 0){ vec_free(in); return ALN_NULL; }
	zw = calloc((size_t)ql + 2, sizeof(int)); zb = calloc((size_t)ql + 2, sizeof(int)); ze = calloc((size_t)ql + 2, sizeof(int));
	for(qx=0,k=0;k<in.n;k++){
		const u32 op = in.a[k] & 0xF; const int len = (int)(in.a[k] >> 4);
		if(op == 0) for(j=0;j<len;j++) zw[qx++] = W;
		else if(op == 1) for(j=0;j<len;j++) zw[qx++] = W + len;
	}
	for(qx=0,k=0;k<in.n;k++){
		const u32 op = in.a[k] & 0xF; const int len = (int)(in.a[k] >> 4);
		if(op == 0) qx += len;
		else if(op == 1){
			for(j=1;j<len&&j<qx;j++) zw[qx-j] += len - j;
			qx += len - 1;
			for(j=1;j<len&&j+qx<ql;j++) zw[qx+j] += len - j;
			qx ++;
		} else {
			for(j=1;j<len&&j<qx;j++) zw[qx-j] += len - j;
			for(j=1;j<len&&j+qx<ql;j++) zw[qx+j] += len - j;
		}
	}
	for(qx=tx=0,k=0;k<in.n;k++){
		const u32 op = in.a[k] & 0xF; const int len = (int)(in.a[k] >> 4);
		if(op == 0 || op == 1){
			for(j=0;j<len;j++){
				int b = tx - zw[qx], e = tx + 1 + zw[qx];
				zb[qx] = b < 0? 0 : b; ze[qx] = e > tl? tl : e;
				if(op == 0) tx ++;
				qx ++;
			}
		} else tx += len;
	}
	{ int lim = 0; for(i=0;i<ql;i++){ if(zb[i] < lim) zb[i] = lim; else lim = zb[i]; } }
	{ int lim = tl; for(i=ql-1;i>=0;i--){ if(ze[i] > lim) ze[i] = lim; else lim = ze[i]; } }
	for(i=0;i<ql;i++) if(ze[i] - zb[i] > wmax) wmax = ze[i] - zb[i];
	hrow = malloc(((size_t)tl + 2) * sizeof(int)); erow = malloc(((size_t)tl + 2) * sizeof(int));
	z = calloc((size_t)ql * wmax, 1);
	hrow[0] = 0; for(j=1;j<=tl;j++) hrow[j] = NEG_SENT;
	for(j=0;j<=tl;j++) erow[j] = NEG_SENT;
	for(i=0;i<ql;i++){
		const u8 qc = query[qb + i]; int hleft = NEG_SENT, f = NEG_SENT; u8 *zi = z + (size_t)i * wmax;
		for(j=zb[i];j<ze[i];j++){
			const int m = hrow[j] + (qc == target[tb + j]? M : X); int e = erow[j], h, t; u8 d;
			hrow[j] = hleft;
			if(m >= e){ d = 0; h = m; } else { d = 1; h = e; }
			if(h < f){ d = 2; h = f; }
			hleft = h;
			t = m + I + E; e += E; if(e > t) d |= 1 << 2; else e = t;
			erow[j] = e;
			t = m + D + E; f += E; if(f > t) d |= 2 << 4; else f = t;
			zi[j - zb[i]] = d;
		}
		hrow[j] = hleft; erow[j] = NEG_SENT;
	}
	y.qb = qb; y.qe = qb + ql; y.tb = tb; y.te = tb + tl; y.score = hrow[tl];
	{
		u8 d = 0; i = ql - 1; j = tl - 1;
		while(i >= 0 && j >= 0){
			if(j < zb[i] || j >= ze[i]){ fprintf(stderr, "zmo_oracle: refine walk left the band (row %d col %d): undefined in the reference\n", i, j); exit(5); }
			d = (z[(size_t)i * wmax + (j - zb[i])] >> (d << 1)) & 3;
			if(d == 0){ if(query[qb + i] == target[tb + j]) y.mat ++; else y.mis ++; i --; j --; }
			else if(d == 1){ i --; y.ins ++; }
			else { j --; y.del ++; }
			cig_push(cigar, d, 1);
		}
		if(i >= 0){ y.ins += i + 1; cig_push(cigar, 1, (u32)(i + 1)); }
		if(j >= 0){ y.del += j + 1; cig_push(cigar, 2, (u32)(j + 1)); }
		cig_reverse(cigar);
	}
	y.aln = y.mat + y.mis + y.ins + y.del;
	free(zw); free(zb); free(ze); free(hrow); free(erow); free(z); vec_free(in);
	return y;
}

/* Pure per-pair alignment (wtzmo.c:1017-1030): windows of the chosen strand -> regions -> stitched
 * alignment.  Returns 0 if no region survived the per-window filter (wtzmo.c:1026,1029). */
static int pair_align(const u8 *pb1, int alen, const u8 *pb2, int blen, const win_t *wins, u32 nwin, const zpair_t *anchors, const zparams_t *par, aln_t *out, u32v *cigar){
	VEC(alnreg_t) regs; u32v cache; u32 j; int ok;
	vec_init(regs); vec_init(cache);
	for(j=0;j<nwin;j++){
		alnreg_t r;
		if(wins[j].closed) continue;
		r.cig_off = (u32)cache.n;
		r.x = window_align(pb1, pb2, &wins[j], anchors, &cache, par);
		r.cig_len = (u32)cache.n - r.cig_off;
		vec_push(cache, 0x0F);  /* separator keeps two windows from fusing (wtzmo.c:1025) */
		if(r.x.aln * 2 < (int)par->zovl || r.x.mat < r.x.aln * par->min_id) continue;
		vec_push(regs, r);
	}
	ok = regs.n != 0;
	if(ok) *out = stitch_regs(alen, blen, regs.a, (u32)regs.n, pb1, pb2, cache.a, cigar, par);
	if(ok && par->refine) *out = refine_alignment(pb2, out->qb, pb1, out->tb, par->w, par, cigar);     /* wtzmo.c:1031-1034 */
	vec_free(regs); vec_free(cache);
	return ok;
}

/* ------------------------------------------------------------------ u64 hash set (closed pairs) */
typedef struct { u64 *tab; size_t cap, n; } u64set_t;
static void u64set_init(u64set_t *s){ s->cap = 1024; s->n = 0; s->tab = malloc(s->cap * 8); memset(s->tab, 0xFF, s->cap * 8); }
static inline size_t u64set_slot(const u64set_t *s, u64 k){ u64 h = k * 0x9E3779B97F4A7C15ULL; size_t i = (h >> 20) & (s->cap - 1); while(s->tab[i] != ~0ULL && s->tab[i] != k) i = (i + 1) & (s->cap - 1); return i; }
static int u64set_has(const u64set_t *s, u64 k){ return s->tab[u64set_slot(s, k)] == k; }
static void u64set_add(u64set_t *s, u64 k){
	size_t i = u64set_slot(s, k);
	if(s->tab[i] == k) return;
	s->tab[i] = k; s->n ++;
	if(s->n * 2 > s->cap){
		u64 *old = s->tab; size_t oc = s->cap, j;
		s->cap <<= 1; s->tab = malloc(s->cap * 8); memset(s->tab, 0xFF, s->cap * 8);
		for(j=0;j<oc;j++) if(old[j] != ~0ULL) s->tab[u64set_slot(s, old[j])] = old[j];
		free(old);
	}
}
static inline u64 pair_key(u32 a, u32 b){ return a < b? (((u64)a << 33) | ((u64)b << 1)) : (((u64)b << 33) | ((u64)a << 1)); }  /* wtzmo.c:84-85 */

/* ------------------------------------------------------------------ records */
typedef struct { u32 pb1, pb2; u8 dir2; int qb, qe, tb, te, score, mat, mis, ins, del, aln; char *cigar; } hit_t;
typedef VEC(hit_t) hitv;

static char* cigar_to_string(const u32 *c, size_t n){     /* kswx.h:1093-1120 */
	size_t i, m = 0; char *s = malloc(n * 12 + 1);
	for(i=0;i<n;i++){
		u32 op = c[i] & 0xF, len = c[i] >> 4;
		if(len == 0) continue;
		if(op > 2){ fprintf(stderr, " -- CIGAR only support M(0),I(1),D(2) cigar, but met ?(%d) --\n", op); exit(1); }
		m += sprintf(s + m, "%u%c", len, "MIDX"[op]);
	}
	s[m] = 0;
	return s;
}

/* ------------------------------------------------------------------ whole-run state + replay (wtzmo.c:803-1134, 1170-1357) */
typedef struct { u32 pb2; u32 ovl; u8 dir, closed; u32 cand_idx; } seed_t;
typedef struct {
	readset_t rs; zparams_t par; kindex_t ix;
	u8 *masked; u32 *rdcovs; u64set_t closed; u32 avg_rdlen; u32 kcut;
	u64v *rdhits;              /* per-read candidate carry-over, only with -G > 1 */
	u64 n_records, aln_cols, n_pairs, n_zpairs, n_seeded;
	/* ZMO_ORACLE_LAG=n (analysis aid): what a speculative batch pipeline that learns of a mask n query reads late would have to seed */
	u32 *mask_k, cur_k; int lag; u64 lag_reads, lag_pairs, n_reads_done;
} zmo_t;

static int gt_cand_ol_desc(const void *a, const void *b, void *ctx){ (void)ctx; return (u32)(*(const u64*)b) > (u32)(*(const u64*)a); }
static int gt_seed_ovl_desc(const void *a, const void *b, void *ctx){ (void)ctx; return ((const seed_t*)b)->ovl > ((const seed_t*)a)->ovl; }

typedef struct { hitv hits; u32v masks; u64v closed; VEC(seed_t) seeds; u32 rd_id; } readout_t;

static void flush_read(zmo_t *z, readout_t *ro, FILE *out){
	size_t i; const readset_t *rs = &z->rs;
	if(!z->par.do_align){   /* -N: seed lines only (wtzmo.c:1176-1181) */
		for(i=0;i<ro->seeds.n;i++){
			seed_t *s = &ro->seeds.a[i];
			if(s->closed) continue;
			fprintf(out, "# %s\t%c\t%d\t%s\t%c\t%d\t%d\n", rs->reads.a[ro->rd_id].name, '+', rs->reads.a[ro->rd_id].len, rs->reads.a[s->pb2].name, "+-"[s->dir], rs->reads.a[s->pb2].len, s->ovl);
		}
	}
	/* with -N the reference prints hits only in step with the seed list (wtzmo.c:1176-1186): no seeds (dot-matrix mode) = no output, no rdcovs */
	for(i=0;z->par.do_align&&i<ro->hits.n;i++){
		hit_t *h = &ro->hits.a[i]; u32 x1, x2; int l1 = rs->reads.a[h->pb1].len, l2 = rs->reads.a[h->pb2].len;
		if(h->aln == 0) h->aln = 1;
		x1 = imin(h->tb, h->qb); x2 = imin(l1 - h->te, l2 - h->qe);
		if(x1 + x2 <= z->par.max_unalign_in_dovetail){ z->rdcovs[h->pb1] ++; z->rdcovs[h->pb2] ++; }
		fprintf(out, "%s\t%c\t%d\t%d\t%d", rs->reads.a[h->pb1].name, '+', l1, h->tb, h->te);
		fprintf(out, "\t%s\t%c\t%d\t%d\t%d", rs->reads.a[h->pb2].name, "+-"[h->dir2], l2, h->qb, h->qe);
		fprintf(out, "\t%d\t%0.3f\t%d\t%d\t%d\t%d", h->score, 1.0 * h->mat / h->aln, h->mat, h->mis, h->ins, h->del);
		if(h->cigar){ fprintf(out, "\t%s\n", h->cigar); free(h->cigar); h->cigar = NULL; } else fprintf(out, "\t0M\n");
		z->n_records ++;
		z->aln_cols += z->par.dot_matrix? (u64)h->aln : (u64)(h->mat + h->mis + h->ins + h->del);
	}
	vec_clear(ro->hits); vec_clear(ro->seeds);
	if(z->par.skip_contained) for(i=0;i<ro->masks.n;i++){ if(z->mask_k && !z->masked[ro->masks.a[i]]) z->mask_k[ro->masks.a[i]] = z->cur_k; z->masked[ro->masks.a[i]] = 1; }
	vec_clear(ro->masks);
	for(i=0;i<ro->closed.n;i++) u64set_add(&z->closed, ro->closed.a[i]);
	vec_clear(ro->closed);
}

static void masks_put(u32v *m, u32 id){ size_t i; for(i=0;i<m->n;i++) if(m->a[i] == id) return; vec_push(*m, id); }

typedef struct { int score, qb, qe, tb, te, strand; } dotres_t;
static dotres_t dot_matrix_pair(zpairv *cache, int alen, int blen, const zparams_t *par);

/* candidate list of a read for the current index partition (wtzmo.c:810-822) */
static void read_candidates(zmo_t *z, u32 pbid, u64v *cands){
	eventv ev; size_t i;
	vec_init(ev);
	candidate_events(&z->rs, &z->ix, pbid, &z->par, &ev);
	candidates_from_events(&ev, &z->par, cands);
	for(i=0;i<cands->n;i++) if(u64set_has(&z->closed, pair_key(pbid, (u32)(cands->a[i] >> 32)))) cands->a[i] &= 0xFFFFFFFF00000000ULL;
	ref_sort(cands->a, cands->n, 8, gt_cand_ol_desc, NULL);
	while(cands->n && (u32)cands->a[cands->n - 1] == 0) cands->n --;
	vec_free(ev);
}

static void process_read(zmo_t *z, u32 pbid, u32 bcov, readout_t *ro){
	const zparams_t *par = &z->par; const readset_t *rs = &z->rs;
	u32 alen = rs->reads.a[pbid].len, nbest, i, j, k, ncand;
	u64v cands_local, *cands; zindex_t zi; zpairv cache; u8 *pb1, *pb2; u32 maxlen = 0;
	VEC(pair_seed_t) pseeds; u16 *windeps; float *weights;
	ro->rd_id = pbid;
	nbest = (u32)(((size_t)par->nbest) * alen / z->avg_rdlen);
	if(nbest < (u32)par->nbest) nbest = par->nbest;
	if(bcov >= nbest) return;
	if(z->rdhits) cands = &z->rdhits[pbid]; else { vec_init(cands_local); cands = &cands_local; }
	read_candidates(z, pbid, cands);
	for(i=0;i<rs->reads.n;i++) if(rs->reads.a[i].len > maxlen) maxlen = rs->reads.a[i].len;
	pb1 = malloc(alen + 1); pb2 = malloc(maxlen + 1);
	rs_unpack(rs, pbid, 0, pb1);
	memset(&zi, 0, sizeof(zi)); vec_init(cache); vec_init(pseeds);
	zindex_build(&zi, pb1, alen, par);
	windeps = calloc(alen + 1, sizeof(u16)); weights = malloc((alen + 1) * sizeof(float));
	for(i=0;i<cands->n;i++){
		u32 id2 = (u32)(cands->a[i] >> 32), blen = rs->reads.a[id2].len; pair_seed_t ps; int dir;
		rs_unpack(rs, id2, 0, pb2);
		zmatch(&zi, pb2, blen, par, &cache);
		z->n_zpairs += cache.n; z->n_seeded ++;
		if(cache.n * par->zsize < (u32)par->ztot) continue;
		if(par->dot_matrix){
			dotres_t r; u32 ol;
			vec_push(ro->closed, pair_key(id2, pbid));
			r = dot_matrix_pair(&cache, alen, blen, par);
			ol = imax(r.qe - r.qb, r.te - r.tb);
			if(r.score >= par->min_score && r.score >= (int)(par->min_id * ol)){
				hit_t h; memset(&h, 0, sizeof(h));
				h.pb1 = pbid; h.pb2 = id2; h.dir2 = r.strand; h.score = r.score; h.tb = r.tb; h.te = r.te; h.qb = r.qb; h.qe = r.qe;
				h.mat = r.score; h.aln = ol; h.cigar = NULL;
				vec_push(ro->hits, h);
			}
			continue;
		}
		pair_seed_init(&ps);
		pair_seed_compute(&ps, &cache, par);
		for(dir=0;dir<2;dir++) for(j=0;j<ps.wins[dir].n;j++){
			win_t *w = &ps.wins[dir].a[j];
			for(k=w->beg[0];(int)k<w->end[0];k++) windeps[k] ++;
		}
		dir = ((u32)ps.ovl[0] & WIN_OVL_MASK) < ((u32)ps.ovl[1] & WIN_OVL_MASK);
		if(((u32)ps.ovl[dir] & WIN_OVL_MASK) >= (u32)par->ztot){
			seed_t s; s.pb2 = id2; s.dir = dir; s.ovl = (u32)ps.ovl[dir] & WIN_OVL_MASK; s.closed = 0; s.cand_idx = (u32)pseeds.n;
			vec_push(ro->seeds, s); vec_push(pseeds, ps);
		} else pair_seed_free(&ps);
	}
	if(!z->rdhits) vec_free(cands_local);
	if(!par->dot_matrix){
		/* repeat weighting (wtzmo.c:933-980); float/double expression shapes kept */
		for(i=0;i<alen;i++)
			weights[i] = (windeps[i] <= par->wnorm)? 1.0 : ((windeps[i] >= par->wrep)? 0.0 : par->wnorm / (float)windeps[i]);
		for(i=0;i<alen;i++) weights[i] = weights[i] * (0.3 + 0.7 * (idiff(((int)i), (int)alen / 2) / ((int)alen / 2.0)));
		for(i=0;i<ro->seeds.n;i++){
			seed_t *s = &ro->seeds.a[i]; pair_seed_t *ps = &pseeds.a[s->cand_idx]; int blen = rs->reads.a[s->pb2].len; u32 ol = 0; double avg;
			for(j=0;j<ps->wins[s->dir].n;j++){
				win_t *w = &ps->wins[s->dir].a[j];
				avg = (w->end[0] - w->beg[0]) * weights[(w->beg[0] + w->end[0]) / 2];
				avg = avg * (0.3 + 0.7 * (idiff(((int)((w->beg[1] + w->end[1]) / 2)), blen / 2) / (blen / 2.0)));
				ol += avg;
			}
			s->ovl = ol & WIN_OVL_MASK;
			if(ol * par->wrep < par->ztot * par->wnorm) s->closed = 1;
		}
		ref_sort(ro->seeds.a, ro->seeds.n, sizeof(seed_t), gt_seed_ovl_desc, NULL);
		if(par->do_align){
			u32v cigar; vec_init(cigar);
			ncand = par->ncand;
			for(i=0;i<ro->seeds.n&&i<ncand;i++){
				seed_t *s = &ro->seeds.a[i]; pair_seed_t *ps = &pseeds.a[s->cand_idx]; int blen = rs->reads.a[s->pb2].len; aln_t x; hit_t h; u32 x1, x2, x3, x4; int l1 = alen, l2 = blen;
				if(s->closed){ ncand ++; continue; }
				vec_push(ro->closed, pair_key(s->pb2, pbid));
				z->n_pairs ++;
				rs_unpack(rs, s->pb2, s->dir, pb2);
				if(!pair_align(pb1, alen, pb2, blen, ps->wins[s->dir].a, (u32)ps->wins[s->dir].n, ps->anc[s->dir].a, par, &x, &cigar)){ s->closed = 1; ncand ++; continue; }
				if(x.score < par->min_score || x.mat < x.aln * par->min_id) continue;
				memset(&h, 0, sizeof(h));
				h.pb1 = pbid; h.pb2 = s->pb2; h.dir2 = s->dir; h.score = x.score; h.tb = x.tb; h.te = x.te; h.qb = x.qb; h.qe = x.qe;
				h.mat = x.mat; h.mis = x.mis; h.ins = x.ins; h.del = x.del; h.aln = x.aln; h.cigar = cigar_to_string(cigar.a, cigar.n);
				vec_push(ro->hits, h);
				x1 = imin(h.tb, h.qb); x2 = imin(l1 - h.te, l2 - h.qe);
				if(x1 + x2 <= par->max_unalign_in_dovetail){
					if(par->skip_contained){
						x3 = ((h.tb == 0 && h.qb) || (h.te == l1 && h.qe < l2));
						x4 = ((h.qb == 0 && h.tb) || (h.qe == l2 && h.te < l1));
						x1 = l2 + h.qb - h.qe; x2 = l1 + h.tb - h.te;
						if(x1 <= par->max_unalign_in_contained && x3 == 0){
							if(x2 <= par->max_unalign_in_contained && x4 == 0){
								if(l1 > l2){ masks_put(&ro->masks, h.pb2); }
								else if(l1 < l2){ masks_put(&ro->masks, h.pb1); break; }
								else if(h.pb2 > h.pb1){ masks_put(&ro->masks, h.pb2); continue; }
								else { masks_put(&ro->masks, h.pb1); break; }
							} else { masks_put(&ro->masks, h.pb2); continue; }
							ncand ++;
						} else if(x2 <= par->max_unalign_in_contained && x4 == 0){ masks_put(&ro->masks, h.pb1); break; }
					}
					bcov ++;
					if(bcov >= nbest) break;
				}
			}
			vec_free(cigar);
		}
	}
	for(i=0;i<pseeds.n;i++) pair_seed_free(&pseeds.a[i]);
	vec_free(pseeds); vec_free(cache); vec_free(zi.seeds); vec_free(zi.slots);
	free(pb1); free(pb2); free(windeps); free(weights);
}

static void run_overlap(zmo_t *z, FILE *out){
	const zparams_t *par = &z->par; readset_t *rs = &z->rs; u32 j, beg, end, pbbeg = 0, pbend = 0, i_idx; u64 tot = 0; readout_t ro;
	memset(&ro, 0, sizeof(ro)); ro.rd_id = 0xFFFFFFFFU;
	if(rs->n_qr == 0 && rs->n_rd){ for(j=0;j<rs->n_rd;j++) tot += rs->reads.a[j].len; z->avg_rdlen = (u32)(tot / rs->n_rd); }
	else if(rs->n_qr){ for(j=0;j<rs->n_qr;j++) tot += rs->reads.a[j + rs->n_rd].len; z->avg_rdlen = (u32)(tot / rs->n_qr); }
	else z->avg_rdlen = 10000;
	if(par->n_idx > 1){ z->rdhits = calloc(rs->n_rd + rs->n_qr, sizeof(u64v)); }
	z->kcut = par->kcut;
	for(i_idx=0;i_idx<(u32)par->n_idx;i_idx++){
		pbbeg = pbend; pbend = pbbeg + (rs->n_rd + par->n_idx - 1) / par->n_idx;
		kindex_free(&z->ix);
		kindex_build(&z->ix, rs, pbbeg, pbend, par, &z->kcut);
		fprintf(stderr, "[oracle] index %u/%u: %zu k-mers, %llu postings, K=%u\n", i_idx + 1, par->n_idx, z->ix.n_ent, (unsigned long long)z->ix.n_post, z->ix.K);
		if(i_idx + 1 >= (u32)par->n_idx) break;
		for(j=0;j<rs->n_rd;j++){       /* just_query passes (wtzmo.c:1289-1301): bcov is 0, no nbest exit */
			if((j % par->n_job) != (u32)par->i_job) continue;
			if(z->masked[j]) continue;
			read_candidates(z, j, &z->rdhits[j]);
		}
	}
	if(rs->n_qr == 0){ beg = 0; end = rs->n_rd; } else { beg = rs->n_rd; end = beg + rs->n_qr; }
	if(getenv("ZMO_ORACLE_LAG")){ z->lag = atoi(getenv("ZMO_ORACLE_LAG")); z->mask_k = calloc(rs->n_rd + rs->n_qr + 1, sizeof(u32)); }
	for(j=beg;j<end;j++){
		if((j % par->n_job) != (u32)par->i_job) continue;
		z->cur_k ++;
		if(z->mask_k && (!z->masked[j] || z->mask_k[j] + (u32)z->lag > z->cur_k)){      /* not known to be masked n reads ago: a lagging pipeline seeds it */
			u64v cl; vec_init(cl); if(z->masked[j]) read_candidates(z, j, &cl);
			z->lag_reads ++; if(z->masked[j]) z->lag_pairs += cl.n;
			vec_free(cl);
		}
		if(z->masked[j]) continue;            /* checked BEFORE the previous read's masks are merged (wtzmo.c:1315 vs 1322) */
		z->n_reads_done ++;
		flush_read(z, &ro, out);
		process_read(z, j, z->rdcovs[j], &ro);
	}
	flush_read(z, &ro, out);
	vec_free(ro.hits); vec_free(ro.masks); vec_free(ro.closed); vec_free(ro.seeds);
}

/* ------------------------------------------------------------------ dot-matrix mode (-U), hzm_aln.h:721-1181 */
typedef struct { int offset; u32 off, cnt; } diag_t;
typedef VEC(diag_t) diagv;

static int gt_zpair_diag(const void *a, const void *b, void *ctx){
	const zpair_t *x = a, *y = b; (void)ctx;
	return ((((i64)x->off1 - (i64)x->off2) << 32) | (i64)x->off1) > ((((i64)y->off1 - (i64)y->off2) << 32) | (i64)y->off1);
}
static int gt_idx_off1(const void *a, const void *b, void *ctx){ const zpair_t *rs = ctx; return rs[*(const u32*)a].off1 > rs[*(const u32*)b].off1; }
static int gt_zpair_gid_off1(const void *a, const void *b, void *ctx){
	const zpair_t *x = a, *y = b; (void)ctx;
	return (x->gid > y->gid)? 1 : ((x->gid < y->gid)? 0 : (x->off1 > y->off1));
}
/* group-id map clean-up shared by both passes (hzm_aln.h:836-846, 1016-1026), cubic on purpose */
static void tidy_groups(u32v *grps){
	size_t i, j, k;
	for(i=1;i<grps->n;i++){
		if(grps->a[i] < i) continue;
		for(j=i+1;j<grps->n;j++){
			if(grps->a[j] != i) continue;
			for(k=j+1;k<grps->n;k++) if(grps->a[k] == j) grps->a[k] = (u32)i;
		}
	}
}

/* hzm_aln.h:721-889: per strand, bucket diagonals (yvar high, advancing yvar/2), link co-linear runs
 * within xvar along q, union group ids, emit blocks spanning >= min_len on q.  Quirks kept: a bucket
 * never contains the strand's last diagonal; a diagonal's members are read as `cnt` CONSECUTIVE
 * entries of the mixed-strand list; after a run is closed, len restarts from the previous element. */
static void denoise_strand(zpair_t *rs, u32 n, int dir, int xvar, int yvar, int min_len, zpairv *dst, winv *regs){
	diagv diags; u32v block, grps; u32 i, j, k, doff, dcnt, gid; int have = 0, lst_offset, end_offset, len;
	vec_init(diags); vec_init(block); vec_init(grps);
	vec_clear(*dst); vec_clear(*regs);
	for(i=0;i<n;i++){
		int dg;
		if(rs[i].dir1 ^ rs[i].dir2 ^ dir) continue;
		dg = (int)rs[i].off1 - (int)rs[i].off2;
		if(have && diags.a[diags.n-1].offset == dg) diags.a[diags.n-1].cnt ++;
		else { diag_t d; d.offset = dg; d.off = i; d.cnt = 1; vec_push(diags, d); have = 1; }
	}
	doff = 0; end_offset = -0x7FFFFFFF;
	vec_push(grps, 0);
	while(doff < n && diags.n){
		lst_offset = diags.a[doff].offset; dcnt = 0;
		while(1){
			if(diags.a[dcnt + doff].offset > lst_offset + yvar) break;
			if(dcnt + doff + 1 >= diags.n) break;
			dcnt ++;
		}
		if(dcnt == 0) break;
		if(diags.a[doff + dcnt].offset == end_offset){ doff += dcnt; continue; }
		end_offset = diags.a[doff + dcnt].offset;
		vec_clear(block);
		for(i=0;i<dcnt;i++){
			diag_t *d = &diags.a[i + doff];
			for(j=0;j<d->cnt;j++){
				if(d->off + j >= n) break;   /* reference would read past the list here (cannot: cnt entries exist at or after off) */
				if(rs[d->off + j].dir1 ^ rs[d->off + j].dir2 ^ dir) continue;
				vec_push(block, d->off + j);
			}
		}
		ref_sort(block.a, block.n, sizeof(u32), gt_idx_off1, rs);
		if(block.n){
			int p0_off1 = rs[block.a[0]].off1, p0_len1 = rs[block.a[0]].len1;
			len = p0_len1; j = 0;
			for(i=1;i<=block.n;i++){
				int p_off1 = (i == block.n)? 0x7FFFFFFF : (int)rs[block.a[i]].off1, p_len1 = (i == block.n)? 0 : (int)rs[block.a[i]].len1;
				if(p_off1 <= p0_off1 + p0_len1 || p_off1 <= p0_off1 + p0_len1 + xvar){
					len += (int)((u32)p_off1 + (u32)p_len1) - (p0_off1 + p0_len1);
				} else {
					if(len >= min_len){
						gid = 0;
						for(k=j;k<i;k++){ u32 g = rs[block.a[k]].gid; if(g){ if(gid == 0) gid = grps.a[g]; else if(gid > grps.a[g]) gid = grps.a[g]; } }
						if(gid == 0){ gid = (u32)grps.n; vec_push(grps, gid); }
						else { for(k=j;k<i;k++){ u32 g = rs[block.a[k]].gid; if(g) grps.a[g] = gid; } }
						for(;j<i;j++) rs[block.a[j]].gid = gid;
					}
					j = i;
					len = p0_len1;
				}
				p0_off1 = p_off1; p0_len1 = p_len1;
			}
		}
		for(i=doff;i<doff+dcnt;i++) if(diags.a[i].offset > lst_offset + yvar / 2) break;
		doff = i;
	}
	tidy_groups(&grps);
	for(i=0;i<n;i++){
		if(rs[i].dir1 ^ rs[i].dir2 ^ dir) continue;
		if(rs[i].gid == 0) continue;
		rs[i].gid = grps.a[rs[i].gid];
		vec_push(*dst, rs[i]);
	}
	ref_sort(dst->a, dst->n, sizeof(zpair_t), gt_zpair_gid_off1, NULL);
	j = 0;
	for(i=1;i<=dst->n;i++){
		win_t s; u32 lst = 0;
		if(i < dst->n && dst->a[i].gid == dst->a[j].gid) continue;
		memset(&s, 0, sizeof(s));
		s.pb2 = 0; s.closed = 0; s.dir = dir; s.anc[0] = j; s.anc[1] = i;
		s.beg[0] = s.beg[1] = 0x7FFFFFFF; s.end[0] = s.end[1] = 0; s.ovl = 0;
		for(k=j;k<i;k++){
			const zpair_t *p = &dst->a[k];
			if((int)p->off1 < s.beg[0]) s.beg[0] = p->off1;
			if((int)(p->off1 + p->len1) > s.end[0]) s.end[0] = p->off1 + p->len1;
			if((int)p->off2 < s.beg[1]) s.beg[1] = p->off2;
			if((int)(p->off2 + p->len2) > s.end[1]) s.end[1] = p->off2 + p->len2;
			s.ovl = (s.ovl + ((p->off1 > lst)? p->len1 : p->off1 + p->len1 - lst)) & WIN_OVL_MASK;
			lst = p->off1 + p->len1;
		}
		if(s.end[0] - s.beg[0] >= min_len) vec_push(*regs, s);
		j = i;
	}
	vec_free(diags); vec_free(block); vec_free(grps);
}

static int gt_win_diag(const void *a, const void *b, void *ctx){
	const win_t *x = a, *y = b; (void)ctx;
	return ((((i64)(x->beg[0] - x->beg[1])) << 32) | (i64)x->beg[0]) > ((((i64)(y->beg[0] - y->beg[1])) << 32) | (i64)y->beg[0]);
}
static int gt_idx_wbeg0(const void *a, const void *b, void *ctx){ const win_t *w = ctx; return w[*(const u32*)a].beg[0] > w[*(const u32*)b].beg[0]; }
static int gt_win_grp_beg0(const void *a, const void *b, void *ctx){
	const win_t *x = a, *y = b; (void)ctx;
	return (x->pb2 > y->pb2)? 1 : ((x->pb2 < y->pb2)? 0 : (x->beg[0] > y->beg[0]));
}
static int gt_win_closed(const void *a, const void *b, void *ctx){ (void)ctx; return ((const win_t*)a)->closed > ((const win_t*)b)->closed; }
static int gt_win_beg0(const void *a, const void *b, void *ctx){ (void)ctx; return ((const win_t*)a)->beg[0] > ((const win_t*)b)->beg[0]; }

/* hzm_aln.h:933-1054: merge blocks lying on nearby diagonals whose q-start is within xvar of the
 * FIRST block's end of the current group; pb2 doubles as group id; every block is its own diagonal
 * entry (d is reset inside the loop). */
static void merge_blocks(winv *regs, int xvar, int yvar){
	diagv diags; u32v block, grps; u32 i, j, k, doff, dcnt, gid; int lst_offset, end_offset; u32 n = (u32)regs->n;
	vec_init(diags); vec_init(block); vec_init(grps);
	ref_sort(regs->a, regs->n, sizeof(win_t), gt_win_diag, NULL);
	for(i=0;i<n;i++){ diag_t d; d.offset = regs->a[i].beg[0] - regs->a[i].beg[1]; d.off = i; d.cnt = 1; vec_push(diags, d); }
	doff = 0; end_offset = -0x7FFFFFFF;
	vec_push(grps, 0);
	while(doff < n){
		lst_offset = diags.a[doff].offset; dcnt = 0;
		while(1){
			if(diags.a[dcnt + doff].offset > lst_offset + yvar) break;
			if(dcnt + doff + 1 >= diags.n) break;
			dcnt ++;
		}
		if(dcnt == 0) break;
		if(diags.a[doff + dcnt].offset == end_offset){ doff += dcnt; continue; }
		end_offset = diags.a[doff + dcnt].offset;
		vec_clear(block);
		for(i=0;i<dcnt;i++) vec_push(block, diags.a[i + doff].off);
		ref_sort(block.a, block.n, sizeof(u32), gt_idx_wbeg0, regs->a);
		{
			int s0_end0 = regs->a[block.a[0]].end[0];
			j = 0;
			for(i=1;i<=block.n;i++){
				int s_beg0 = (i == block.n)? 0x7FFFFFFF : regs->a[block.a[i]].beg[0];
				int s_end0 = (i == block.n)? 0 : regs->a[block.a[i]].end[0];
				if(s_beg0 <= s0_end0 + xvar) continue;
				gid = 0;
				for(k=j;k<i;k++){ u32 g = regs->a[block.a[k]].pb2; if(g){ if(gid == 0) gid = grps.a[g]; else grps.a[g] = gid; } }
				if(gid == 0){ gid = (u32)grps.n; vec_push(grps, gid); }
				for(;j<i;j++) regs->a[block.a[j]].pb2 = gid;
				j = i; s0_end0 = s_end0;
			}
		}
		for(i=doff;i<doff+dcnt;i++) if(diags.a[i].offset > lst_offset + yvar / 2) break;
		doff = i;
	}
	tidy_groups(&grps);
	for(i=0;i<n;i++) if(regs->a[i].pb2) regs->a[i].pb2 = grps.a[regs->a[i].pb2];
	ref_sort(regs->a, regs->n, sizeof(win_t), gt_win_grp_beg0, NULL);
	for(j=0;j<n;j++) if(regs->a[j].pb2) break;
	for(i=j+1;i<=n;i++){
		win_t *s0;
		if(i < n && regs->a[i].pb2 == regs->a[j].pb2) continue;
		s0 = &regs->a[j];
		for(k=j+1;k<i;k++){
			win_t *s = &regs->a[k];
			s->closed = 1;
			if(s->beg[0] < s0->beg[0]) s0->beg[0] = s->beg[0];
			if(s->end[0] > s0->end[0]) s0->end[0] = s->end[0];
			if(s->beg[1] < s0->beg[1]) s0->beg[1] = s->beg[1];
			if(s->end[1] > s0->end[1]) s0->end[1] = s->end[1];
			s0->ovl = (s0->ovl + s->ovl) & WIN_OVL_MASK;
		}
		j = i;
	}
	ref_sort(regs->a, regs->n, sizeof(win_t), gt_win_closed, NULL);
	for(i=0;i<n;i++) if(regs->a[i].closed) break;
	regs->n = i;
	vec_free(diags); vec_free(block); vec_free(grps);
}

static inline int sx30(int v){ return (int)((u32)v << 2) >> 2; }  /* node_t.weight is a 30-bit signed field (hzm_aln.h:1057) */
/* hzm_aln.h:1056-1132: chain blocks allowing max_overhang overlap, float penalties on diagonal
 * deviation and gap; head/tail flags favour chains that reach the read ends.  Returns sum of ovl. */
static int chain_blocks(int len1, int len2, winv *regs, int tail_margin, int max_overhang, float band_penalty, float gap_penalty){
	typedef struct { int weight; u8 head, tail; int bt; } node_t;
	u32 n = (u32)regs->n, i, j; node_t *nodes = malloc((n + 1) * sizeof(node_t)); int mw = -1000000, bt = -1, band, gap, weight, W, score;
	win_t *r = regs->a;
	ref_sort(r, n, sizeof(win_t), gt_win_beg0, NULL);
	for(i=0;i<n;i++){
		nodes[i].bt = -1; nodes[i].weight = 0; nodes[i].head = nodes[i].tail = 0;
		if(r[i].beg[0] <= tail_margin || r[i].beg[1] <= tail_margin) nodes[i].head = 1;
		if(r[i].end[0] + tail_margin > len1 || r[i].end[1] + tail_margin > len2) nodes[i].tail = 1;
	}
	for(i=0;i<n;i++){
		r[i].closed = 1;
		nodes[i].weight = sx30(nodes[i].weight + (int)r[i].ovl);
		weight = nodes[i].weight * ((nodes[i].head + 3) * (nodes[i].tail + 3)) / 16;
		if(weight > mw){ mw = weight; bt = i; }
		W = nodes[i].weight / gap_penalty;
		for(j=i+1;j<n;j++){
			if(r[j].beg[0] + max_overhang < r[i].end[0]) continue;
			if(r[j].beg[1] + max_overhang < r[i].end[1]) continue;
			if(r[j].beg[0] - r[i].end[0] > W) break;
			band = idiff(r[j].beg[0] - r[i].end[0], r[j].beg[1] - r[i].end[1]);
			gap = imax(r[j].beg[0] - r[i].end[0], r[j].beg[1] - r[i].end[1]);
			if(gap < 0) gap = -gap;
			score = band * band_penalty + gap * gap_penalty;
			score = nodes[i].weight - score;
			if(nodes[j].weight <= score){ nodes[j].weight = sx30(score); nodes[j].bt = i; nodes[j].head = nodes[i].head; }
		}
	}
	mw = 0;
	while(bt >= 0){ r[bt].closed = 0; mw += r[bt].ovl; bt = nodes[bt].bt; }
	free(nodes);
	return mw;
}

static dotres_t dot_matrix_pair(zpairv *cache, int alen, int blen, const zparams_t *par){
	zpairv dst[2]; winv regs[2]; int weight[2], d; u32 i; dotres_t r;
	for(d=0;d<2;d++){ vec_init(dst[d]); vec_init(regs[d]); }
	ref_sort(cache->a, cache->n, sizeof(zpair_t), gt_zpair_diag, NULL);
	for(d=0;d<2;d++) denoise_strand(cache->a, (u32)cache->n, d, par->xvar, par->yvar, par->min_block_len, &dst[d], &regs[d]);
	for(d=0;d<2;d++) merge_blocks(&regs[d], par->xvar, 2 * par->yvar);
	for(d=0;d<2;d++) weight[d] = chain_blocks(alen, blen, &regs[d], par->xvar, par->max_overhang, par->deviation_penalty, par->gap_penalty);
	d = (weight[0] < weight[1]);
	r.score = weight[d]; r.qb = r.tb = 0x7FFFFFFF; r.qe = r.te = 0; r.strand = d;
	for(i=0;i<regs[d].n;i++){
		win_t *s = &regs[d].a[i];
		if(s->closed) continue;
		if(r.qb > s->beg[1]) r.qb = s->beg[1];
		if(r.tb > s->beg[0]) r.tb = s->beg[0];
		if(r.qe < s->end[1]) r.qe = s->end[1];
		if(r.te < s->end[0]) r.te = s->end[0];
	}
	for(d=0;d<2;d++){ vec_free(dst[d]); vec_free(regs[d]); }
	return r;
}

/* ------------------------------------------------------------------ command line (wtzmo.c:1512-1812) */
static int file_exists(const char *f){ struct stat st; return stat(f, &st) == 0; }
static int usage(void){ printf("zmo_oracle: CPU oracle restating `wtzmo -t 1`; same options as wtzmo (see INTEGRATION.md)\n"); return 1; }

#ifndef ZMO_ORACLE_LIB
int main(int argc, char **argv){
	zmo_t Z, *z = &Z; zparams_t *par = &Z.par; int c; float optval;
	char *output = NULL, *pairoutf = NULL; FILE *out;
	VEC(char*) pbs, flts, ovls, obts, tbas; u32 i;
	memset(z, 0, sizeof(*z)); zparams_default(par);
	vec_init(pbs); vec_init(flts); vec_init(ovls); vec_init(obts); vec_init(tbas);
	while((c = getopt(argc, argv, "ht:P:p:Ni:b:J:I:o:9:S:fCH:k:G:z:Z:U:y:d:r:q:l:K:A:B:r:R:L:F:W:w:e:M:X:O:E:T:s:m:nv")) != -1){
		switch(c){
			case 'h': return usage();
			case 't': par->ncpu = atoi(optarg); break;
			case 'P': par->n_job = atoi(optarg); break;
			case 'p': par->i_job = atoi(optarg); break;
			case 'N': par->do_align = 0; break;
			case 'i': vec_push(pbs, optarg); break;
			case 'b': vec_push(obts, optarg); break;
			case 'J': par->min_rdlen = atoi(optarg); break;
			case 'I': vec_push(tbas, optarg); break;
			case 'o': output = optarg; break;
			case '9': pairoutf = optarg; break;
			case 'S': par->ksave = atoi(optarg); break;
			case 'f': par->overwrite = 1; break;
			case 'C': par->write_contained = 0; break;   /* -C never reaches wt->skip_contained (wtzmo.c:168,1609,1781): it only suppresses the .contained file */
			case 'H': par->hk = atoi(optarg); par->hz = (par->hk >> 1) & 1; par->hk &= 1; break;
			case 'k': par->ksize = atoi(optarg); break;
			case 'K': par->kcut = atoi(optarg); break;
			case 'z': par->zsize = atoi(optarg); break;
			case 'Z': par->zcut = atoi(optarg); break;
			case 'U': optval = atof(optarg);
				if(optval < 0){ par->dot_matrix = 5; break; }
				switch(par->dot_matrix){
					case 0: par->xvar = optval; break;
					case 1: par->yvar = optval; break;
					case 2: par->min_block_len = optval; break;
					case 3: par->deviation_penalty = optval; break;
					case 4: par->gap_penalty = optval; break;
					default: par->dot_matrix = 5;
				}
				par->dot_matrix ++;
				break;
			case 'y': par->kwin = atoi(optarg); break;
			case 'l': par->kvar = atoi(optarg); break;
			case 'd': par->kovl = atof(optarg); break;
			case 'G': par->n_idx = atoi(optarg); break;
			case 'r': par->ztot = atof(optarg); break;
			case 'R': par->zovl = atof(optarg); break;
			case 'q': par->wrep = atoi(optarg); break;
			case 'A': par->ncand = atoi(optarg); break;
			case 'B': par->nbest = atoi(optarg); break;
			case 'w': par->w = atoi(optarg); break;
			case 'e': par->ew = atoi(optarg); break;
			case 'W': par->W = atoi(optarg); break;
			case 'M': par->M = atoi(optarg); break;
			case 'X': par->X = atoi(optarg); break;
			case 'O': par->O = atoi(optarg); break;
			case 'E': par->E = atoi(optarg); break;
			case 'T': par->T = atoi(optarg); break;
			case 'L': vec_push(ovls, optarg); break;
			case 'F': vec_push(flts, optarg); break;
			case 's': par->min_score = atoi(optarg); break;
			case 'm': par->min_id = atof(optarg); break;
			case 'n': par->refine = 1; break;
			case 'v': par->debug ++; break;
			default: return usage();
		}
	}
	if(output == NULL) return usage();
	if(!par->overwrite && strcmp(output, "-") && file_exists(output)){ fprintf(stderr, "File exists! '%s'\n\n", output); return usage(); }
	if(pbs.n == 0) return usage();
	if(par->ksize > 32 || par->ksize < 5) return usage();
	if(par->zsize > 16 || par->zsize < 5) return usage();
	if(par->ksave < 1) return usage();
	par->max_overhang = 2 * par->xvar;
	par->kstep = par->kwin / 2;
	rs_load(&z->rs, pbs.a, (int)pbs.n, par->min_rdlen, 0);
	ref_sort(z->rs.reads.a, z->rs.reads.n, sizeof(read_t), gt_read_len_desc, NULL);      /* wtzmo.c:1708 */
	if(tbas.n) rs_load(&z->rs, tbas.a, (int)tbas.n, par->min_rdlen, 1);
	z->masked = calloc(z->rs.n_rd + z->rs.n_qr + 1, 1);
	z->rdcovs = calloc(z->rs.n_rd + z->rs.n_qr + 1, sizeof(u32));
	u64set_init(&z->closed);
	{	/* side inputs: tab tables / name lists, '#' lines skipped (wtzmo.c:1732-1773) */
		char *line = NULL; size_t cap = 0; FILE *fp;
		for(i=0;i<obts.n;i++){
			if((fp = fopen(obts.a[i], "r")) == NULL) exit(1);
			while(getline(&line, &cap, fp) >= 0){
				char *nm, *a, *b, *sv; u32 id; int coff, clen;
				if(line[0] == '#') continue;
				nm = strtok_r(line, "\t\n", &sv); a = strtok_r(NULL, "\t\n", &sv); b = strtok_r(NULL, "\t\n", &sv);
				if(!nm || !a || !b) continue;
				if((id = rs_find(&z->rs, z->rs.n_rd, nm)) == 0xFFFFFFFFU) continue;
				coff = atoi(a); clen = atoi(b);
				if(coff < 0 || coff + clen > (int)z->rs.reads.a[id].len) continue;
				z->rs.reads.a[id].off += coff; z->rs.reads.a[id].len = clen;
			}
			fclose(fp);
		}
		for(i=0;i<flts.n;i++){
			if((fp = fopen(flts.a[i], "r")) == NULL) exit(1);
			while(getline(&line, &cap, fp) >= 0){
				u32 id; size_t l = strlen(line);
				while(l && line[l-1] == '\n') line[--l] = 0;
				if(line[0] == '#') continue;
				if((id = rs_find(&z->rs, z->rs.n_rd, line)) == 0xFFFFFFFFU) continue;
				z->masked[id] = 1;
			}
			fclose(fp);
		}
		for(i=0;i<ovls.n;i++){
			if((fp = fopen(ovls.a[i], "r")) == NULL) exit(1);
			while(getline(&line, &cap, fp) >= 0){
				char *a, *b, *sv; u32 p1, p2;
				if(line[0] == '#') continue;
				a = strtok_r(line, "\t\n", &sv); b = strtok_r(NULL, "\t\n", &sv);
				if(!a || !b) continue;
				if((p1 = rs_find(&z->rs, z->rs.n_rd, a)) == 0xFFFFFFFFU) continue;
				if((p2 = rs_find(&z->rs, z->rs.n_rd, b)) == 0xFFFFFFFFU) continue;
				u64set_add(&z->closed, pair_key(p1, p2));
			}
			fclose(fp);
		}
		free(line);
	}
	out = strcmp(output, "-")? fopen(output, "w") : stdout;
	run_overlap(z, out);
	if(strcmp(output, "-")) fclose(out);
	if(par->write_contained && strcmp(output, "-")){
		char *maskf = malloc(strlen(output) + 16); FILE *mf;
		sprintf(maskf, "%s.contained", output); mf = fopen(maskf, "w");
		for(i=0;i<z->rs.n_rd;i++) if(z->masked[i]) fprintf(mf, "%s\n", z->rs.reads.a[i].name);
		fclose(mf); free(maskf);
	}
	if(pairoutf){
		FILE *pf = fopen(pairoutf, "w"); size_t k;
		for(k=0;k<z->closed.cap;k++){
			u64 v = z->closed.tab[k];
			if(v == ~0ULL) continue;
			fprintf(pf, "%s\t%s\n", z->rs.reads.a[v >> 33].name, z->rs.reads.a[(v & 0xFFFFFFFFU) >> 1].name);
		}
		fclose(pf);
	}
	if(z->mask_k) fprintf(stderr, "[oracle] lag=%d: reads processed=%llu, reads a lagging pipeline seeds=%llu (+%llu pairs of reads that were already masked)\n", z->lag, (unsigned long long)z->n_reads_done, (unsigned long long)z->lag_reads, (unsigned long long)z->lag_pairs);
	fprintf(stderr, "[oracle] records=%llu aligned_cols=%llu pairs_seeded=%llu pairs_aligned=%llu zpairs=%llu\n", (unsigned long long)z->n_records, (unsigned long long)z->aln_cols, (unsigned long long)z->n_seeded, (unsigned long long)z->n_pairs, (unsigned long long)z->n_zpairs);
	return 0;
}
#endif

/* ------------------------------------------------------------------ library exports for the tests (same shapes as ref_shim.c) */
#ifdef ZMO_ORACLE_LIB
static void orc_export(aln_t x, int *out){ out[0]=x.score; out[1]=x.tb; out[2]=x.te; out[3]=x.qb; out[4]=x.qe; out[5]=x.aln; out[6]=x.mat; out[7]=x.mis; out[8]=x.ins; out[9]=x.del; }
static zparams_t orc_par(int zsize, int hz, int zcut, int kvar, int kwin, int kstep, int zovl, int ztot, int W){
	zparams_t p; zparams_default(&p); p.zsize = zsize; p.hz = hz; p.zcut = zcut; p.kvar = kvar; p.kwin = kwin; p.kstep = kstep; p.zovl = zovl; p.ztot = ztot; p.W = W; return p;
}
int orc_extend(int mode, int qlen, u8 *q, int tlen, u8 *t, int strand, int init, int W, int M, int X, int I, int D, int E, int T, int *out, u32 *cigar_out, int cigar_cap){
	u32v cg; int n, i; aln_t x; vec_init(cg);
	x = banded_extend(mode, qlen, q, tlen, t, strand, init, W, M, X, I, D, E, T, &cg);
	orc_export(x, out); n = (int)cg.n;
	for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg.a[i];
	vec_free(cg); return n;
}
int orc_global2(int qlen, u8 *q, int tlen, u8 *t, int M, int X, int o_del, int e_del, int o_ins, int e_ins, int w, int *score_out, u32 *cigar_out, int cigar_cap){
	u32v cg; int n, i; vec_init(cg);
	*score_out = banded_global(qlen, q, tlen, t, M, X, o_del, e_del, o_ins, e_ins, w, &cg);
	n = (int)cg.n; for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg.a[i];
	vec_free(cg); return n;
}
int orc_hz_align(u8 *pb1, u32 len1, u8 *pb2, u32 len2, int M, int I, int D, int E, int *out, u32 *cigar_out, int cigar_cap){
	u32v cg; int n, i; aln_t x; vec_init(cg);
	x = runlen_align(pb1, len1, pb2, len2, M, I, D, E, &cg);
	orc_export(x, out); n = (int)cg.n; for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg.a[i];
	vec_free(cg); return n;
}
static int gt_u64_asc(const void *a, const void *b, void *c){ (void)c; return *(const u64*)a > *(const u64*)b; }
static int gt_u64_hi_asc(const void *a, const void *b, void *c){ (void)c; return (*(const u64*)a >> 32) > (*(const u64*)b >> 32); }
static int gt_u64_hi_desc(const void *a, const void *b, void *c){ (void)c; return (*(const u64*)b >> 32) > (*(const u64*)a >> 32); }
void orc_sort_u64_asc(u64 *a, size_t n){ ref_sort(a, n, 8, gt_u64_asc, NULL); }
void orc_sort_u64_lo32_desc(u64 *a, size_t n){ ref_sort(a, n, 8, gt_cand_ol_desc, NULL); }
void orc_sort_u64_hi32_asc(u64 *a, size_t n){ ref_sort(a, n, 8, gt_u64_hi_asc, NULL); }
void orc_sort_u64_hi32_desc(u64 *a, size_t n){ ref_sort(a, n, 8, gt_u64_hi_desc, NULL); }

int orc_pair_windows(u8 *pb1, int alen, u8 *pb2, int blen, int zsize, int hz, int zcut, int kvar, int kwin, int kstep, int zovl, int ztot, int W,
		int *n_hzmp, int *ovl, int *win_out, int win_cap, int *anc_out, int anc_cap, int *n_anc_out){
	zparams_t par = orc_par(zsize, hz, zcut, kvar, kwin, kstep, zovl, ztot, W);
	zindex_t zi; zpairv cache; pair_seed_t ps; int d, nw = 0, na = 0; size_t j; u32 k;
	memset(&zi, 0, sizeof(zi)); vec_init(cache); pair_seed_init(&ps);
	zindex_build(&zi, pb1, alen, &par);
	zmatch(&zi, pb2, blen, &par, &cache);
	/* the shim reports windows of every strand with windows, regardless of ztot; mirror that */
	*n_hzmp = (int)cache.n; ovl[0] = ovl[1] = 0;
	if(cache.n * par.zsize >= (u32)par.ztot){
		ref_sort(cache.a, cache.n, sizeof(zpair_t), gt_zpair_off12, NULL);
		for(d=0;d<2;d++){
			winv w2; zpairv a2; vec_init(w2); vec_init(a2);
			if(pair_windows_strand(cache.a, (u32)cache.n, d, &w2, &a2, &par)){
				ovl[d] = chain_windows(w2.a, (u32)w2.n, par.W);
				for(j=0;j<w2.n;j++){
					win_t *w = &w2.a[j];
					if(w->closed) continue;
					if(nw < win_cap){ int *o = win_out + 7 * nw; o[0]=d; o[1]=w->beg[0]; o[2]=w->end[0]; o[3]=w->beg[1]; o[4]=w->end[1]; o[5]=w->ovl; o[6]=w->anc[1]-w->anc[0]; }
					nw ++;
					for(k=w->anc[0];k<w->anc[1];k++){
						zpair_t *p = &a2.a[k];
						if(na < anc_cap){ int *o = anc_out + 6 * na; o[0]=p->off1; o[1]=p->off2; o[2]=p->len1; o[3]=p->len2; o[4]=p->dir1; o[5]=p->dir2; }
						na ++;
					}
				}
			}
			vec_free(w2); vec_free(a2);
		}
	}
	*n_anc_out = na;
	vec_free(cache); vec_free(zi.seeds); vec_free(zi.slots); pair_seed_free(&ps);
	return nw;
}
int orc_pair_dotmatrix(u8 *pb1, int alen, u8 *pb2, int blen, int zsize, int hz, int zcut, int kvar, int xvar, int yvar, int min_block_len, int max_overhang, float dev_pen, float gap_pen, int *out){
	zparams_t par = orc_par(zsize, hz, zcut, kvar, 800, 400, 200, 300, 3200);
	zindex_t zi; zpairv cache; dotres_t r; int n;
	par.xvar = xvar; par.yvar = yvar; par.min_block_len = min_block_len; par.max_overhang = max_overhang; par.deviation_penalty = dev_pen; par.gap_penalty = gap_pen;
	memset(&zi, 0, sizeof(zi)); vec_init(cache);
	zindex_build(&zi, pb1, alen, &par);
	zmatch(&zi, pb2, blen, &par, &cache);
	n = (int)cache.n;
	r = dot_matrix_pair(&cache, alen, blen, &par);
	out[0]=r.score; out[1]=r.qb; out[2]=r.qe; out[3]=r.tb; out[4]=r.te; out[5]=r.strand;
	vec_free(cache); vec_free(zi.seeds); vec_free(zi.slots);
	return n;
}
/* one window of orc_pair_windows' output through the per-window anchored alignment (fast_seeds_align_hzmo, hzm_aln.h:1247-1302):
 * pb2 is c on the window's strand (reverse-complemented by the caller for strand 1), anc = n_anc x {off1, off2, len1, len2, dir1, dir2}.
 * out = score, tb, te, qb, qe, aln, mat, mis, ins, del; returns the number of CIGAR ops. */
int orc_window_align(u8 *pb1, u8 *pb2, const int *anc, int n_anc, int w, int M, int X, int O, int E, int T, int *out, u32 *cigar_out, int cigar_cap){
	zparams_t par = orc_par(10, 1, 64, 2, 800, 400, 200, 300, 3200);
	win_t win; zpair_t *a = malloc(sizeof(zpair_t) * (size_t)(n_anc > 0? n_anc : 1)); u32v cg; aln_t x; int i, n;
	par.w = w; par.M = M; par.X = X; par.O = O; par.E = E; par.T = T;
	memset(&win, 0, sizeof(win)); win.anc[0] = 0; win.anc[1] = (u32)n_anc;
	for(i=0;i<n_anc;i++){ zpair_t *p = &a[i]; const int *o = anc + 6 * i; memset(p, 0, sizeof(*p)); p->off1 = o[0]; p->off2 = o[1]; p->len1 = o[2]; p->len2 = o[3]; p->dir1 = o[4]; p->dir2 = o[5]; }
	vec_init(cg);
	x = window_align(pb1, pb2, &win, a, &cg, &par);
	out[0]=x.score; out[1]=x.tb; out[2]=x.te; out[3]=x.qb; out[4]=x.qe; out[5]=x.aln; out[6]=x.mat; out[7]=x.mis; out[8]=x.ins; out[9]=x.del;
	n = (int)cg.n; for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg.a[i];
	vec_free(cg); free(a); return n;
}
/* -n: kswx_refine_alignment (kswx.h:483-659) of an alignment given by its start (qb on `query` = c on the strand shown, tb on `target` = q)
 * and CIGAR inside half band W (wtzmo passes -w).  out as in orc_window_align (tb/te/qb/qe relative to the start, as the restatement
 * returns them); returns the number of new CIGAR ops. */
int orc_refine(u8 *query, int qb, u8 *target, int tb, int W, int M, int X, int O, int E, const u32 *cigar_in, int n_in, int *out, u32 *cigar_out, int cigar_cap){
	zparams_t par = orc_par(10, 1, 64, 2, 800, 400, 200, 300, 3200);
	u32v cg; aln_t x; int i, n;
	par.M = M; par.X = X; par.O = O; par.E = E;
	vec_init(cg); vec_reserve(cg, (size_t)(n_in > 0? n_in : 0) + 1);
	for(i=0;i<n_in;i++) cg.a[i] = cigar_in[i];
	cg.n = (size_t)(n_in > 0? n_in : 0);
	x = refine_alignment(query, qb, target, tb, W, &par, &cg);
	orc_export(x, out);
	n = (int)cg.n; for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg.a[i];
	vec_free(cg); return n;
}
/* global k-mer index over reads [beg, end) of a read set given as 0..3 codes back to back (index_wtzmo, wtzmo.c:349-430) and the candidate
 * event stream of query read qid against it (query_wtzmo, wtzmo.c:433-562): events with union length >= kovl in ascending
 * (target<<1|strand) order as tkey/ol pairs.  *kcut_io < 2: automatic K, written back.  idx_stats = {distinct k-mers, postings kept}. */
int orc_candidates(const u8 *seqs, const int *lens, int nreads, int beg, int end, int qid, int ksize, int hk, int ksave, int kovl, u32 *kcut_io,
		u64 *idx_stats, u32 *ev_out, int ev_cap){
	static const char L[4] = {'A', 'C', 'G', 'T'};
	zparams_t par = orc_par(10, 1, 64, 2, 800, 400, 200, 300, 3200);
	readset_t rs; kindex_t ix; eventv ev; size_t o = 0, i; int r, n = 0; char *tmp;
	memset(&rs, 0, sizeof(rs)); memset(&ix, 0, sizeof(ix)); vec_init(ev);
	par.ksize = ksize; par.hk = hk; par.ksave = ksave; par.kovl = kovl;
	for(r=0;r<nreads;r++){
		char name[32]; int nl = sprintf(name, "s%d", r);
		tmp = malloc((size_t)lens[r] + 1);
		for(i=0;i<(size_t)lens[r];i++) tmp[i] = L[seqs[o + i] & 3];
		rs_add_read(&rs, name, nl, tmp, (u32)lens[r]); free(tmp); o += (size_t)lens[r];
	}
	rs.n_rd = (u32)nreads;
	kindex_build(&ix, &rs, (u32)beg, (u32)end, &par, kcut_io);
	idx_stats[0] = ix.n_ent; idx_stats[1] = ix.n_post;
	candidate_events(&rs, &ix, (u32)qid, &par, &ev);
	for(i=0;i<ev.n;i++) if(ev.a[i].ol >= (u32)kovl){ if(n < ev_cap){ ev_out[2 * n] = ev.a[i].tkey; ev_out[2 * n + 1] = ev.a[i].ol; } n ++; }
	vec_free(ev); kindex_free(&ix);
	for(r=0;r<nreads;r++) free(rs.reads.a[r].name);
	vec_free(rs.reads); free(rs.bits);
	return n;
}
/* pure per-pair alignment of one strand (wtzmo.c:1017-1034): windows -> per-window regions -> region filter -> stitched alignment
 * (left extension, gaps, right extension) -> optional -n refinement.  pb2 = c on the strand of the windows; win = n_win x
 * {n_anchors}, anc as in orc_window_align (windows' anchors back to back).  Returns -1 if no region survived, else the number of
 * CIGAR ops. */
int orc_pair_align(u8 *pb1, int alen, u8 *pb2, int blen, const int *win, int n_win, const int *anc, int w, int ew, int W, int zovl, float min_id,
		int M, int X, int O, int E, int T, int refine, int *out, u32 *cigar_out, int cigar_cap){
	zparams_t par = orc_par(10, 1, 64, 2, 800, 400, zovl, 300, W);
	win_t *wins = calloc((size_t)(n_win > 0? n_win : 1), sizeof(win_t)); zpair_t *a; u32v cg; aln_t x = ALN_NULL; int i, n, na = 0, ok;
	par.w = w; par.ew = ew; par.min_id = min_id; par.M = M; par.X = X; par.O = O; par.E = E; par.T = T; par.refine = refine;
	for(i=0;i<n_win;i++){ wins[i].anc[0] = (u32)na; na += win[i]; wins[i].anc[1] = (u32)na; }
	a = calloc((size_t)(na > 0? na : 1), sizeof(zpair_t));
	for(i=0;i<na;i++){ zpair_t *p = &a[i]; const int *o = anc + 6 * i; p->off1 = o[0]; p->off2 = o[1]; p->len1 = o[2]; p->len2 = o[3]; p->dir1 = o[4]; p->dir2 = o[5]; }
	vec_init(cg);
	ok = pair_align(pb1, alen, pb2, blen, wins, (u32)n_win, a, &par, &x, &cg);
	orc_export(x, out);
	n = (int)cg.n; for(i=0;i<n&&i<cigar_cap;i++) cigar_out[i] = cg.a[i];
	vec_free(cg); free(a); free(wins);
	return ok? n : -1;
}
#endif
