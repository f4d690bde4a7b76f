/*
 * wtzmo_main.c -- B200 build of SMARTdenovo's `wtzmo` overlapper: same command line, same
 * .ovl / .dmo.ovl / .contained / -9 outputs (wtzmo.c:1422-1812), so smartdenovo.pl, run_zmo.sh and
 * run_dmo.sh work unchanged.
 *
 * The host keeps what is inherently sequential in the reference -- option parsing, FASTA/FASTQ
 * loading into the 2-bit bank, the length sort, and the read-by-read STATE REPLAY (masked reads, tried
 * pairs, per-read overlap counters, candidate heap quirks, repeat weighting, nbest/containment
 * breaks, wtzmo.c:803-1134,1170-1357) -- and calls libzmo_b200.so (include/zmo_b200.h) for every
 * numeric stage.  The pthread worker pool of thread.h is replaced by batch dispatch onto the GPU:
 * reads are processed in id order in speculative batches; the device computes the pure per-read /
 * per-pair results for a superset of what the state machine will consume, and the replay then walks
 * the reads exactly in `-t 1` order, so the output is byte-identical to `wtzmo -t 1`.
 * There is no CPU implementation of the numeric stages in this program: without a usable GPU it
 * stops with an error.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <getopt.h>
#include <math.h>
#include <time.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <pthread.h>
#include "../../../include/zmo_b200.h"


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
static inline void bank_or(u64 *bits, u64 off, u64 b){ bits[off >> 5] |= b << (((~off) & 31) << 1); }
static inline int b_is_header(const char *buf, size_t p){ return buf[p] == '>' && (p == 0 || buf[p - 1] == '\n'); }
static inline void bank_put(u64 *bits, u64 off, u64 b){ if((off & 31) == 0) bits[off >> 5] = 0; bits[off >> 5] |= b << (((~off) & 31) << 1); }



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
	{
		/* sequential append in the BaseBank layout (dna.h:78,263: base i in word i>>5, MSB first): one shift-or per base
		 * into an accumulator, one store per 32 bases; non-ACGT -> lrand48() & 3 in file order like dna.h:405 */
		static u8 tab[256]; static int tab_ok = 0;
		u64 w = rs->nbases >> 5, acc; int fill = (int)(rs->nbases & 31);
		if(!tab_ok){ memset(tab, 4, 256); tab['A'] = tab['a'] = 0; tab['C'] = tab['c'] = 1; tab['G'] = tab['g'] = 2; tab['T'] = tab['t'] = 3; tab_ok = 1; }
		acc = fill? rs->bits[w] >> (64 - 2 * fill) : 0;
		for(i=0;i<len;i++){
			u64 c = tab[(u8)seq[i]];
			if(c > 3) c = lrand48() & 3;
			acc = (acc << 2) | c;
			if(++fill == 32){ rs->bits[w++] = acc; acc = 0; fill = 0; }
		}
		if(fill) rs->bits[w] = acc << (64 - 2 * fill);
		rs->nbases += len;
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
		if(sr->fp){ setvbuf(sr->fp, NULL, _IOFBF, 4u << 20); return 1; }
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

/* ------------------------------------------------------------------ parallel FASTA loader (the step right before the path)
 * Plain FASTA files are mapped and cut into one piece per thread at header lines; every thread parses and 2-bit-packs its records into a bank
 * of its own, the banks are then shifted into the global BaseBank in parallel.  Same result as the serial reader below, including the one
 * order-dependent detail: non-ACGT bases become lrand48() & 3 in FILE order (dna.h:405) -- the threads only note their positions, the
 * values are drawn afterwards in one sequential pass.  FASTQ, gzip and stdin inputs take the serial reader. */
#include <sys/mman.h>
#include <fcntl.h>
#include <unistd.h>
static u8 g_base_tab[256];
typedef struct {
	const char *buf; size_t beg, end, fend; int min_rdlen;
	u64 *bits; u64 nbases, cap_words; VEC(read_t) reads; u64v bad; pthread_t th;
} pchunk_t;
static void* pchunk_run(void *arg){
	pchunk_t *c = arg; const char *b = c->buf; size_t p = c->beg; const u8 *tab = g_base_tab;
	while(p < c->end){
		size_t nb, q, ls, len = 0; read_t r; u64 acc; int fill;
		/* header: name = up to the first blank (file_reader.c:319-329) */
		for(nb = p + 1; nb < c->fend; nb++){ char ch = b[nb]; if(ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n') break; }
		r.name = malloc(nb - p); memcpy(r.name, b + p + 1, nb - p - 1); r.name[nb - p - 1] = 0;
		{ const char *e = memchr(b + p, '\n', c->fend - p); q = e? (size_t)(e - b) + 1 : c->fend; }
		/* sequence lines up to the next header at a line start: first their total length, then the packing */
		for(ls = q; ls < c->fend && b[ls] != '>'; ){ const char *e = memchr(b + ls, '\n', c->fend - ls); size_t le = e? (size_t)(e - b) : c->fend; len += le - ls; ls = e? le + 1 : c->fend; }
		p = ls;
		if((long long)len < (long long)c->min_rdlen || len > 0xFFFFFFFFULL){ free(r.name); continue; }
		{ u64 need = (c->nbases + len + 31) / 32 + 2; if(need > c->cap_words){ u64 m = c->cap_words? c->cap_words : 4096; while(m < need) m <<= 1; c->bits = realloc(c->bits, m * 8); c->cap_words = m; } }
		r.off = c->nbases; r.len = (u32)len;
		{
			u64 w = c->nbases >> 5, pos = c->nbases; fill = (int)(c->nbases & 31);
			acc = fill? c->bits[w] >> (64 - 2 * fill) : 0;
			for(ls = q; ls < p; ){
				const char *e = memchr(b + ls, '\n', p - ls); size_t le = e? (size_t)(e - b) : p, k;
				for(k = ls; k < le; k++, pos++){
					u64 v = tab[(u8)b[k]];
					if(v > 3){ vec_push(c->bad, pos); v = 0; }
					acc = (acc << 2) | v;
					if(++fill == 32){ c->bits[w++] = acc; acc = 0; fill = 0; }
				}
				ls = e? le + 1 : p;
			}
			if(fill) c->bits[w] = acc << (64 - 2 * fill);
			c->nbases += len;
		}
		vec_push(c->reads, r);
	}
	return NULL;
}
typedef struct { pchunk_t *c; u64 *dst; u64 g0; pthread_t th; } pmerge_t;
static void* pmerge_run(void *arg){
	pmerge_t *m = arg; const pchunk_t *c = m->c; u64 nw = (c->nbases + 31) / 32, k, w0 = m->g0 >> 5; int r = (int)(m->g0 & 31);
	for(k=0;k<nw;k++){
		u64 v = c->bits[k];
		if(k == nw - 1 && (c->nbases & 31)) v &= ~0ULL << (64 - 2 * (c->nbases & 31));      /* bases past the end of the piece */
		if(r == 0) __sync_fetch_and_or(&m->dst[w0 + k], v);
		else { __sync_fetch_and_or(&m->dst[w0 + k], v >> (2 * r)); __sync_fetch_and_or(&m->dst[w0 + k + 1], v << (64 - 2 * r)); }
	}
	return NULL;
}
/* returns 1 if the files were loaded, 0 if the serial reader has to do it (nothing touched then) */
static int rs_load_parallel(readset_t *rs, char **files, int nfiles, int min_rdlen, int as_query, int nth){
	int f, t; size_t i;
	if(nth < 2) return 0;
	memset(g_base_tab, 4, 256); g_base_tab['A'] = g_base_tab['a'] = 0; g_base_tab['C'] = g_base_tab['c'] = 1; g_base_tab['G'] = g_base_tab['g'] = 2; g_base_tab['T'] = g_base_tab['t'] = 3;
	for(f=0;f<nfiles;f++){
		size_t l = strlen(files[f]); int fd; char ch = 0; struct stat st;
		if(!strcmp(files[f], "-") || (l > 3 && !strcmp(files[f] + l - 3, ".gz"))) return 0;
		if((fd = open(files[f], O_RDONLY)) < 0) return 0;
		if(fstat(fd, &st) || !S_ISREG(st.st_mode) || st.st_size < 2 || read(fd, &ch, 1) != 1 || ch != '>'){ close(fd); return 0; }
		close(fd);
	}
	for(f=0;f<nfiles;f++){
		int fd = open(files[f], O_RDONLY); struct stat st; const char *buf; pchunk_t *ch; pmerge_t *mg; u64 g0, tot = 0; size_t fend; int n;
		fstat(fd, &st); fend = (size_t)st.st_size;
		buf = mmap(NULL, fend, PROT_READ, MAP_PRIVATE, fd, 0);
		close(fd);
		if(buf == MAP_FAILED){ fprintf(stderr, " -- Cannot map %s --\n", files[f]); exit(1); }
		madvise((void*)buf, fend, MADV_SEQUENTIAL);
		n = nth; if((size_t)n > fend / 4096 + 1) n = (int)(fend / 4096 + 1);
		ch = calloc(n + 1, sizeof(pchunk_t)); mg = calloc(n + 1, sizeof(pmerge_t));
		for(t=0;t<=n;t++){
			/* piece t starts at the first header line at or after its share of the file */
			size_t p = t == n? fend : (size_t)((unsigned __int128)fend * t / n);
			if(t == 0) p = 0;
			else while(p < fend){ const char *e; if(b_is_header(buf, p)) break; e = memchr(buf + p, '\n', fend - p); p = e? (size_t)(e - buf) + 1 : fend; }
			ch[t].beg = p;
		}
		for(t=0;t<n;t++){ ch[t].buf = buf; ch[t].end = ch[t + 1].beg; ch[t].fend = fend; ch[t].min_rdlen = min_rdlen; if(ch[t].end < ch[t].beg) ch[t].end = ch[t].beg; }
		for(t=0;t<n;t++) if(pthread_create(&ch[t].th, NULL, pchunk_run, &ch[t]) != 0){ pchunk_run(&ch[t]); ch[t].th = 0; }
		for(t=0;t<n;t++){ if(ch[t].th) pthread_join(ch[t].th, NULL); tot += ch[t].nbases; }
		{	/* grow the global bank (zero-filled), then shift the pieces in */
			u64 need = (rs->nbases + tot + 31) / 32 + 2;
			if(need > rs->cap_words){ u64 m = rs->cap_words? rs->cap_words : 1024; while(m < need) m <<= 1; rs->bits = realloc(rs->bits, m * 8); memset(rs->bits + rs->cap_words, 0, (m - rs->cap_words) * 8); rs->cap_words = m; }
			/* words past the current end must be zero: the serial appender leaves garbage-free words (it zeroes on growth and writes whole words) */
			{ u64 w = (rs->nbases + 31) / 32; memset(rs->bits + w, 0, (rs->cap_words - w) * 8); if(rs->nbases & 31) rs->bits[rs->nbases >> 5] &= ~0ULL << (64 - 2 * (rs->nbases & 31)); }
		}
		g0 = rs->nbases;
		for(t=0;t<n;t++){ mg[t].c = &ch[t]; mg[t].dst = rs->bits; mg[t].g0 = g0; g0 += ch[t].nbases; }
		for(t=0;t<n;t++) if(pthread_create(&mg[t].th, NULL, pmerge_run, &mg[t]) != 0){ pmerge_run(&mg[t]); mg[t].th = 0; }
		for(t=0;t<n;t++) if(mg[t].th) pthread_join(mg[t].th, NULL);
		for(t=0;t<n;t++){
			/* reads in file order; non-ACGT bases drawn in file order (dna.h:405) */
			for(i=0;i<ch[t].reads.n;i++){ read_t r = ch[t].reads.a[i]; r.off += mg[t].g0; vec_push(rs->reads, r); if(as_query) rs->n_qr ++; else rs->n_rd ++; }
			for(i=0;i<ch[t].bad.n;i++) bank_or(rs->bits, mg[t].g0 + ch[t].bad.a[i], (u64)(lrand48() & 3));
			free(ch[t].bits); vec_free(ch[t].reads); vec_free(ch[t].bad);
		}
		rs->nbases = g0;
		free(ch); free(mg);
		munmap((void*)buf, fend);
	}
	return 1;
}

static void rs_load(readset_t *rs, char **files, int nfiles, int min_rdlen, int as_query){
	seqreader_t sr; u8v name, seq;
	{ char *env = getenv("ZMO_LOAD_THREADS"); long nc = sysconf(_SC_NPROCESSORS_ONLN); int nth = env? atoi(env) : (int)(nc > 16? 16 : nc); if(rs_load_parallel(rs, files, nfiles, min_rdlen, as_query, nth)) return; }
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



/* test hook (CPU suite): load the files with `threads` loader threads (0 / 1 = the serial reader) from the process-start state of lrand48 and
 * return {reads, bases, FNV-1a over (lengths, names, bank words)}: the parallel loader must reproduce the serial one bit for bit */
int wz_load_digest(int nfiles, char **files, int min_rdlen, int threads, uint64_t out[3]){
	readset_t rs; char tbuf[32]; u64 h = 1469598103934665603ULL; size_t i, k; const char *old = getenv("ZMO_LOAD_THREADS"); char *keep = old? strdup(old) : NULL;
	memset(&rs, 0, sizeof(rs));
	srand48(0x1234ABCD);      /* = the initial state of lrand48() in a fresh process */
	snprintf(tbuf, sizeof(tbuf), "%d", threads); setenv("ZMO_LOAD_THREADS", tbuf, 1);
	rs_load(&rs, files, nfiles, min_rdlen, 0);
	if(keep){ setenv("ZMO_LOAD_THREADS", keep, 1); free(keep); } else unsetenv("ZMO_LOAD_THREADS");
#define FNV(b) do { h ^= (u64)(b); h *= 1099511628211ULL; } while(0)
	for(i=0;i<rs.reads.n;i++){ const read_t *r = &rs.reads.a[i]; const char *s = r->name; FNV(r->len); FNV(r->off); while(*s) FNV((u8)*s++); FNV(0xFF); }
	for(k=0;k<(rs.nbases+31)/32;k++){ u64 w = rs.bits[k]; if(k == (rs.nbases >> 5) && (rs.nbases & 31)) w &= ~0ULL << (64 - 2 * (rs.nbases & 31)); FNV(w & 0xFFFFFFFFULL); FNV(w >> 32); }
#undef FNV
	out[0] = rs.reads.n; out[1] = rs.nbases; out[2] = h;
	for(i=0;i<rs.reads.n;i++) free(rs.reads.a[i].name);
	vec_free(rs.reads); free(rs.bits);
	return 0;
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

/* Turn the ordered event stream of one index partition (already filtered to ol >= kovl on the
 * device) into the candidate array (wtzmo.c:494-571).  `cands` may hold entries carried over from
 * earlier index partitions (-G).  Quirks kept: the heap-full test looks at the CURRENT ol but inserts
 * the PREVIOUS pending candidate; the final flush compares against ol==0; an empty stream pushes the
 * sentinel 0xFFFFFFFF00000000. */
static void candidates_from_events(const zmo_event_t *ev, size_t n, const zparams_t *par, u64v *cands){
	u64 x1 = 0xFFFFFFFF00000000ULL, x2; size_t i; u32 ol = 0;
	for(i=0;i<n;i++){
		ol = ev[i].ol;
		x2 = ((u64)(ev[i].tkey >> 1) << 32) | ol;
		if((x1 >> 32) == (x2 >> 32)){ x1 = (u32)x1 > (u32)x2? x1 : x2; }
		else if(x1 == 0xFFFFFFFF00000000ULL){ x1 = x2; }
		else {
			if(cands->n >= (size_t)par->ncand){ if((u32)cands->a[0] < ol) cheap_replace0(cands, x1); }
			else cheap_push(cands, x1);
			x1 = x2;
		}
	}
	ol = 0;
	if(cands->n >= (size_t)par->ncand){ if((u32)cands->a[0] < ol) cheap_replace0(cands, x1); }
	else cheap_push(cands, x1);
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


/* ------------------------------------------------------------------ records + output (wtzmo.c:1170-1249) */
typedef struct { u32 pb1, pb2; u8 dir2; int qb, qe, tb, te, score, mat, mis, ins, del, aln; const char *cigar; u32 n_cigar; int has_cigar; } hit_t;   /* cigar = device-formatted text, n_cigar bytes */
typedef VEC(hit_t) hitv;
typedef struct { u32 pb2; u32 ovl; u8 dir, closed; u32 cand_idx; } seed_t;
typedef VEC(seed_t) seedv;
#define WIN_OVL_MASK 0x1FFFFFFFU

#define WZ_MAX_CTX 8
#define WZ_PIN_WAVES 6
typedef struct {
	readset_t rs; zparams_t par; zmo_ctx *ctx;      /* ctx = root context (reads + index); ctxs[0] == ctx, ctxs[1..] = clones */
	int call_pairs;
	zmo_ctx *ctxs[WZ_MAX_CTX]; int n_ctx, depth; u8 ctx_busy[WZ_MAX_CTX];
	u8 *masked; u32 *rdcovs; u64set_t closed; u32 avg_rdlen; u32 kcut;
	u64v *rdhits;                       /* per-read candidate carry-over, only with -G > 1 */
	u64 n_records, aln_cols, n_tasks, n_tasks_used, n_pairs_seeded, n_batches, n_reads_batched, n_reads_late_masked, n_pairs_late_masked;
	FILE *out; char *obuf; size_t obuf_n, obuf_cap;
	double t_dev, t_replay, t_write;
	double tl[6];      /* timeline of the last job, seconds since its start: index done, first batch built, first replay starts, last batch built, last compute joined, end */
	int batch_reads, batch_pairs;
	/* page-locked result buffers reused across batches */
	/* one page-locked arena per device context: the result buffers of a batch's DP waves are carved from it (bump allocation, reset when the next
	 * batch starts on that context); allocated once on the start-up thread -- cudaHostAlloc costs ~45 ms per CALL next to running kernels */
	struct { u8 *base; size_t cap, used, want; } pin[WZ_MAX_CTX];
	pthread_mutex_t dev_mu[WZ_MAX_CTX];  /* one device call at a time per context */
	pthread_mutex_t stat_mu;
	u64 n_waves, n_wave_tasks; int wave_margin, wave_growth, wave_predict, wave_predict_b, ramp, drain_div, drain_min;
	int no_cigar;      /* ZMO_OVL_COLS=16: the 16 columns `cut -f1-16` keeps (smartdenovo.pl:58); no CIGAR text is formatted, copied or printed */
	u64 wv_tasks[8], wv_reads[8], wv_count[8];      /* ZMO_WAVE_DEBUG: tasks / reads / launches per DP wave index */
} wz_t;

static double now_s(void){ struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + 1e-6 * tv.tv_usec; }
static void die_zmo(const char *what){ fprintf(stderr, "wtzmo(b200): %s: %s\n", what, zmo_last_error()); exit(3); }

static inline void ob_reserve(wz_t *z, size_t more){
	if(z->obuf_n + more > z->obuf_cap){
		if(z->obuf_n){ fwrite(z->obuf, 1, z->obuf_n, z->out); z->obuf_n = 0; }
		if(more > z->obuf_cap){ z->obuf_cap = more + (1u << 20); z->obuf = realloc(z->obuf, z->obuf_cap); }
	}
}
static inline void ob_flush(wz_t *z){ if(z->obuf_n){ fwrite(z->obuf, 1, z->obuf_n, z->out); z->obuf_n = 0; } }
static inline char* put_u32(char *p, u32 v){ char t[12]; int n = 0; do { t[n++] = '0' + v % 10; v /= 10; } while(v); while(n) *p++ = t[--n]; return p; }

typedef struct { hitv hits; u32v masks; u64v closed; seedv seeds; u32 rd_id; } readout_t;

/* defer_masks: print the hits and merge rdcovs/closed now, but keep the read's masks pending; used at a
 * batch boundary, because the reference tests masked[next read] BEFORE merging the previous read's
 * masks (wtzmo.c:1315 vs 1322) */
static void flush_read(wz_t *z, readout_t *ro, int defer_masks){
	size_t i; const readset_t *rs = &z->rs;
	if(!z->par.do_align){   /* -N: seed lines only (wtzmo.c:1176-1181) */
		for(i=0;i<ro->seeds.n;i++){
			seed_t *s = &ro->seeds.a[i];
			if(s->closed) continue;
			ob_reserve(z, 1024 + strlen(rs->reads.a[ro->rd_id].name) + strlen(rs->reads.a[s->pb2].name));
			z->obuf_n += sprintf(z->obuf + z->obuf_n, "# %s\t%c\t%d\t%s\t%c\t%d\t%d\n", rs->reads.a[ro->rd_id].name, '+', rs->reads.a[ro->rd_id].len, rs->reads.a[s->pb2].name, "+-"[s->dir], rs->reads.a[s->pb2].len, s->ovl);
		}
	}
	/* with -N the reference prints hits only in step with the seed list (wtzmo.c:1176-1186): no seeds (dot-matrix mode) = no output, no rdcovs */
	for(i=0;z->par.do_align&&i<ro->hits.n;i++){
		hit_t *h = &ro->hits.a[i]; u32 x1, x2; int l1 = rs->reads.a[h->pb1].len, l2 = rs->reads.a[h->pb2].len; char *p;
		if(h->aln == 0) h->aln = 1;
		x1 = imin(h->tb, h->qb); x2 = imin(l1 - h->te, l2 - h->qe);
		if(x1 + x2 <= z->par.max_unalign_in_dovetail){ z->rdcovs[h->pb1] ++; z->rdcovs[h->pb2] ++; }
		ob_reserve(z, 512 + strlen(rs->reads.a[h->pb1].name) + strlen(rs->reads.a[h->pb2].name) + (size_t)h->n_cigar + 16);
		z->obuf_n += sprintf(z->obuf + z->obuf_n, "%s\t%c\t%d\t%d\t%d\t%s\t%c\t%d\t%d\t%d\t%d\t%0.3f\t%d\t%d\t%d\t%d", rs->reads.a[h->pb1].name, '+', l1, h->tb, h->te,
			rs->reads.a[h->pb2].name, "+-"[h->dir2], l2, h->qb, h->qe, h->score, 1.0 * h->mat / h->aln, h->mat, h->mis, h->ins, h->del);
		p = z->obuf + z->obuf_n;
		if(!z->no_cigar){
			*p++ = '\t';
			if(h->has_cigar){ memcpy(p, h->cigar, h->n_cigar); p += h->n_cigar; }      /* kswx_cigar2string (kswx.h:1093-1120), formatted by k_cig_text */
			else { *p++ = '0'; *p++ = 'M'; }
		}
		*p++ = '\n';
		z->obuf_n = p - z->obuf;
		z->n_records ++;
		z->aln_cols += z->par.dot_matrix? (u64)h->aln : (u64)(h->mat + h->mis + h->ins + h->del);
	}
	vec_clear(ro->hits); vec_clear(ro->seeds);
	if(!defer_masks){ for(i=0;i<ro->masks.n;i++) z->masked[ro->masks.a[i]] = 1; vec_clear(ro->masks); }
	for(i=0;i<ro->closed.n;i++) u64set_add(&z->closed, ro->closed.a[i]);
	vec_clear(ro->closed);
}
static void masks_put(u32v *m, u32 id){ size_t i; for(i=0;i<m->n;i++) if(m->a[i] == id) return; vec_push(*m, id); }

static int gt_cand_ol_desc(const void *a, const void *b, void *ctx){ (void)ctx; return (u32)(*(const u64*)b) > (u32)(*(const u64*)a); }
static int gt_seed_ovl_desc(const void *a, const void *b, void *ctx){ (void)ctx; return ((const seed_t*)b)->ovl > ((const seed_t*)a)->ovl; }

/* ------------------------------------------------------------------ batch engine
 * One batch = a run of consecutive eligible reads.  Device phases:
 *   A  zmo_candidates      -> ordered (target,strand,ol) events per read           (pure)
 *   B  zmo_pair_windows    -> z-matches, windows, chain weights per (read,cand)    (pure)
 *   C  zmo_pair_align      -> alignment record + CIGAR per (read,cand,strand)      (pure)
 * followed by the sequential replay of the reference state machine, which only CONSUMES results. */
typedef struct {
	u32 rd_id; int skip;                /* skip: bcov >= nbest at batch build time (nothing to compute) */
	u64v cands_raw;                     /* candidate array in heap order after this partition's events */
	u32v cand_pair;                     /* per raw candidate: pair index in the batch, 0xFFFFFFFF = none */
	u32 bcov0;                          /* rdcovs[rd_id] when the batch was built (lower bound of the replay-time value) */
	/* resumable seed walk (wtzmo.c:1005-1120): position and counters survive a pause for an on-demand DP wave */
	int walking; u32 wi, ncand, bcov, nbest;
} bread_t;
typedef struct {
	VEC(bread_t) reads;
	VEC(zmo_pair_t) pairs;
	zmo_pairseed_t *seeds; zmo_window_t *wins; size_t wins_cap;
	VEC(zmo_task_t) tasks;
	zmo_record_t *recs; u32 *cigars; size_t cig_cap;   /* first wave, in the session's pinned buffers */
	const zmo_record_t **pres; const u32 **pcig; u8 *pdir;   /* per pair: alignment of the chosen strand (NULL = not computed yet) */
	VEC(void*) extra;                   /* result buffers of on-demand waves */
	int slot;                           /* device batch slot holding this batch's windows/anchors */
	zmo_dotres_t *dots;
	int ci;                             /* device context (and pinned result buffer set) owned by this batch from build to the end of its replay */
	pthread_t th; int have_thread;
} batch_t;

static u32 read_nbest(const wz_t *z, u32 pbid){
	u32 nbest = (u32)(((size_t)z->par.nbest) * z->rs.reads.a[pbid].len / z->avg_rdlen);
	if(nbest < (u32)z->par.nbest) nbest = z->par.nbest;
	return nbest;
}

/* phase A for a list of reads: events -> raw candidate arrays (appended to `carry` arrays when -G) */
static void batch_candidates(wz_t *z, batch_t *b){
	size_t i, nq = 0; u32 *qids = malloc((b->reads.n + 1) * 4); u64 *off = malloc((b->reads.n + 2) * 8); u32 *qmap = malloc((b->reads.n + 1) * 4);
	zmo_event_t *ev = NULL; u64 cap = 0, need = 0; int rc;
	for(i=0;i<b->reads.n;i++) if(!b->reads.a[i].skip){ qmap[nq] = (u32)i; qids[nq++] = b->reads.a[i].rd_id; }
	if(nq){
		cap = 4096 + 512 * nq; ev = malloc(cap * sizeof(zmo_event_t));
		pthread_mutex_lock(&z->dev_mu[b->ci]);
		rc = zmo_candidates(z->ctxs[b->ci], qids, (u32)nq, off, ev, cap, &need);
		if(rc == ZMO_ERR_CAPACITY){ cap = need + 16; ev = realloc(ev, cap * sizeof(zmo_event_t)); rc = zmo_candidates(z->ctxs[b->ci], qids, (u32)nq, off, ev, cap, &need); }
		pthread_mutex_unlock(&z->dev_mu[b->ci]);
		if(rc) die_zmo("zmo_candidates");
		for(i=0;i<nq;i++){
			bread_t *r = &b->reads.a[qmap[i]];
			if(z->rdhits){ u64v *h = &z->rdhits[r->rd_id]; vec_clear(r->cands_raw); vec_reserve(r->cands_raw, h->n + 1); memcpy(r->cands_raw.a, h->a, h->n * 8); r->cands_raw.n = h->n; }
			candidates_from_events(ev + off[i], (size_t)(off[i + 1] - off[i]), &z->par, &r->cands_raw);
		}
	}
	free(qids); free(off); free(qmap); free(ev);
}

/* candidate post-filter with the CURRENT state (wtzmo.c:813-822); keeps cand_pair aligned with cands */
static void filter_sort_candidates(wz_t *z, u32 pbid, u64v *c, u32v *cp){
	size_t i, n = c->n; u64 *tmp;
	for(i=0;i<n;i++) if(u64set_has(&z->closed, pair_key(pbid, (u32)(c->a[i] >> 32)))) c->a[i] &= 0xFFFFFFFF00000000ULL;
	if(cp){
		/* sort (value, original index) records with the reference permutation: the comparator only
		 * looks at the low 32 bits of the value, so carrying the index along does not change it */
		typedef struct { u64 v; u32 idx; u32 pad; } rec_t;
		rec_t *r = malloc((n + 1) * sizeof(rec_t)); u32 *np = malloc((n + 1) * 4);
		for(i=0;i<n;i++){ r[i].v = c->a[i]; r[i].idx = (u32)i; r[i].pad = 0; }
		ref_sort(r, n, sizeof(rec_t), gt_cand_ol_desc, NULL);
		for(i=0;i<n;i++){ c->a[i] = r[i].v; np[i] = cp->a[r[i].idx]; }
		memcpy(cp->a, np, n * 4);
		free(r); free(np);
	} else ref_sort(c->a, n, 8, gt_cand_ol_desc, NULL);
	while(c->n && (u32)c->a[c->n - 1] == 0) c->n --;
	if(cp) cp->n = c->n;
	(void)tmp;
}

/* chosen strand of a seeded pair, or -1 (wtzmo.c:913-914) */
static inline int pair_dir(const zparams_t *par, const zmo_pairseed_t *ps){
	int dir;
	if(ps->n_zpair * (u32)par->zsize < (u32)par->ztot) return -1;
	dir = ((u32)ps->ovl[0] & WIN_OVL_MASK) < ((u32)ps->ovl[1] & WIN_OVL_MASK);
	return (((u32)ps->ovl[dir] & WIN_OVL_MASK) >= (u32)par->ztot)? dir : -1;
}

/* seeds of a read from its (already state-filtered and sorted) candidates: windeps, repeat weights, re-scoring and
 * the unstable sort by score (wtzmo.c:846-986).  Pure function of the candidate list and the batch's device results;
 * used by the replay and, on the un-filtered list, to PREDICT the alignment order for the first DP wave. */
static void build_seeds(const wz_t *z, const batch_t *b, u32 pbid, const u64 *cands, const u32 *cand_pair, size_t ncands, seedv *seeds){
	const zparams_t *par = &z->par; const readset_t *rs = &z->rs; u32 alen = rs->reads.a[pbid].len, i, j; u16 *windeps; float *weights;
	vec_clear(*seeds);
	windeps = calloc(alen + 2, sizeof(u16)); weights = malloc((alen + 1) * sizeof(float));
	for(i=0;i<ncands;i++){
		u32 id2 = (u32)(cands[i] >> 32), pi = cand_pair[i]; const zmo_pairseed_t *ps; int dir;
		if(pi == 0xFFFFFFFFU){ fprintf(stderr, "wtzmo(b200): internal error: pair (%u,%u) was not seeded\n", pbid, id2); exit(4); }
		ps = &b->seeds[pi];
		if(ps->n_zpair * (u32)par->zsize < (u32)par->ztot) continue;
		/* windeps[k]++ over every kept window (wtzmo.c:908) as a difference array; the uint16 wrap of the
		 * reference counter is reproduced by the mod-2^16 prefix sum below */
		for(dir=0;dir<2;dir++) for(j=0;j<ps->n_win[dir];j++){
			const zmo_window_t *w = &b->wins[ps->win_off[dir] + j];
			if(w->beg[0] < w->end[0]){ windeps[w->beg[0]] ++; windeps[w->end[0]] --; }
		}
		dir = pair_dir(par, ps);
		if(dir >= 0){
			seed_t sd; sd.pb2 = id2; sd.dir = dir; sd.ovl = (u32)ps->ovl[dir] & WIN_OVL_MASK; sd.closed = 0; sd.cand_idx = pi;
			vec_push(*seeds, sd);
		}
	}
	{ u16 acc = 0; for(i=0;i<alen;i++){ acc = (u16)(acc + windeps[i]); windeps[i] = acc; } }
	/* repeat weighting (wtzmo.c:933-980); float/double expression shapes kept */
	for(i=0;i<alen;i++)
		weights[i] = (windeps[i] <= par->wnorm)? 1.0 : ((windeps[i] >= par->wrep)? 0.0 : par->wnorm / (float)windeps[i]);
	for(i=0;i<alen;i++) weights[i] = weights[i] * (0.3 + 0.7 * (idiff(((int)i), (int)alen / 2) / ((int)alen / 2.0)));
	for(i=0;i<seeds->n;i++){
		seed_t *sd = &seeds->a[i]; const zmo_pairseed_t *ps = &b->seeds[sd->cand_idx]; int blen = rs->reads.a[sd->pb2].len; u32 ol = 0; double avg;
		for(j=0;j<ps->n_win[sd->dir];j++){
			const zmo_window_t *w = &b->wins[ps->win_off[sd->dir] + j];
			avg = (w->end[0] - w->beg[0]) * weights[(w->beg[0] + w->end[0]) / 2];
			avg = avg * (0.3 + 0.7 * (idiff(((int)((w->beg[1] + w->end[1]) / 2)), blen / 2) / (blen / 2.0)));
			ol += avg;
		}
		sd->ovl = ol & WIN_OVL_MASK;
		if(ol * par->wrep < par->ztot * par->wnorm) sd->closed = 1;
	}
	ref_sort(seeds->a, seeds->n, sizeof(seed_t), gt_seed_ovl_desc, NULL);
	free(windeps); free(weights);
}

/* containment / dovetail bookkeeping for one accepted hit (wtzmo.c:1064-1100).  Returns 0 = next seed, 2 = stop the
 * walk.  masks may be NULL (dry walk).  wt->skip_contained is always 1: -C never reaches it (wtzmo.c:168,1609). */
static int hit_rules(const zparams_t *par, int l1, int l2, u32 pb1, u32 pb2, const zmo_record_t *h, u32 *bcov, u32 nbest, u32 *ncand, u32v *masks){
	u32 x1, x2, x3, x4;
	x1 = imin(h->tb, h->qb); x2 = imin(l1 - h->te, l2 - h->qe);
	if(x1 + x2 <= par->max_unalign_in_dovetail){
		x3 = ((h->tb == 0 && h->qb) || (h->te == l1 && h->qe < l2));
		x4 = ((h->qb == 0 && h->tb) || (h->qe == l2 && h->te < l1));
		x1 = l2 + h->qb - h->qe; x2 = l1 + h->tb - h->te;
		if(x1 <= par->max_unalign_in_contained && x3 == 0){
			if(x2 <= par->max_unalign_in_contained && x4 == 0){
				if(l1 > l2){ if(masks) masks_put(masks, pb2); }
				else if(l1 < l2){ if(masks) masks_put(masks, pb1); return 2; }
				else if(pb2 > pb1){ if(masks) masks_put(masks, pb2); return 0; }
				else { if(masks) masks_put(masks, pb1); return 2; }
			} else { if(masks) masks_put(masks, pb2); return 0; }
			(*ncand) ++;
		} else if(x2 <= par->max_unalign_in_contained && x4 == 0){ if(masks) masks_put(masks, pb1); return 2; }
		(*bcov) ++;
		if(*bcov >= nbest) return 2;
	}
	return 0;
}

/* dry walk of a predicted seed list with the results computed so far: index of the first seed whose alignment is
 * needed but missing, or -1 if the walk would finish.  No state is touched; cross-read effects are ignored. */
static long dry_walk(const wz_t *z, const batch_t *b, const bread_t *r, const seedv *sv){
	const zparams_t *par = &z->par; const readset_t *rs = &z->rs; u32 ncand = par->ncand, bcov = r->bcov0, nbest = read_nbest(z, r->rd_id); size_t i;
	int alen = rs->reads.a[r->rd_id].len;
	if(bcov >= nbest) return -1;
	for(i=0;i<sv->n&&i<ncand;i++){
		const seed_t *s = &sv->a[i]; const zmo_record_t *x;
		if(s->closed){ ncand ++; continue; }
		x = b->pres[s->cand_idx];
		if(x == NULL) return (long)i;
		if(!x->ok){ ncand ++; continue; }
		if(x->score < par->min_score || x->mat < x->aln * par->min_id) continue;
		if(hit_rules(par, alen, rs->reads.a[s->pb2].len, r->rd_id, s->pb2, x, &bcov, nbest, &ncand, NULL) == 2) return -1;
	}
	return -1;
}

/* Predicted alignment of a seed that has not been aligned yet, from its kept windows: the bounding box of the chain, extended on both sides to
 * the nearer read end (what the end extensions reach when the overlap is real).  Only the speculation policy looks at it (how many seeds of a
 * read the next DP wave aligns); every record the replay uses is a computed one. */
static void predict_rec(const batch_t *b, const seed_t *s, int l1, int l2, zmo_record_t *x){
	const zmo_pairseed_t *ps = &b->seeds[s->cand_idx]; int b0 = l1, e0 = 0, b1 = l2, e1 = 0, lo, hi; u32 j;
	for(j=0;j<ps->n_win[s->dir];j++){
		const zmo_window_t *w = &b->wins[ps->win_off[s->dir] + j];
		b0 = imin(b0, w->beg[0]); e0 = imax(e0, w->end[0]); b1 = imin(b1, w->beg[1]); e1 = imax(e1, w->end[1]);
	}
	memset(x, 0, sizeof(*x));
	if(e0 <= b0 || e1 <= b1) return;
	lo = imin(b0, b1); hi = imin(l1 - e0, l2 - e1); if(hi < 0) hi = 0;
	x->ok = 1; x->tb = b0 - lo; x->te = e0 + hi; x->qb = b1 - lo; x->qe = e1 + hi;
}
/* dry walk in which a missing alignment is replaced by its prediction: index of the seed at which the walk is predicted to stop (the number of
 * seeds if it is predicted to run through).  The DP wave aligns the missing seeds up to there plus a margin, instead of a fixed number per read. */
static long predicted_stop(const wz_t *z, const batch_t *b, const bread_t *r, const seedv *sv, int *why){      /* *why: 0 the walk runs through, 1 the read is predicted contained, 2 enough dovetail hits */
	const zparams_t *par = &z->par; const readset_t *rs = &z->rs; u32 ncand = par->ncand, bcov = r->bcov0, nbest = read_nbest(z, r->rd_id); size_t i;
	int alen = rs->reads.a[r->rd_id].len;
	*why = 0;
	if(bcov >= nbest) return 0;
	for(i=0;i<sv->n&&i<ncand;i++){
		const seed_t *s = &sv->a[i]; const zmo_record_t *x; zmo_record_t px; u32 bc0 = bcov;
		if(s->closed){ ncand ++; continue; }
		x = b->pres[s->cand_idx];
		if(x == NULL){ predict_rec(b, s, alen, rs->reads.a[s->pb2].len, &px); x = &px; }
		else if(!x->ok){ ncand ++; continue; }
		else if(x->score < par->min_score || x->mat < x->aln * par->min_id) continue;
		if(!x->ok) continue;
		if(hit_rules(par, alen, rs->reads.a[s->pb2].len, r->rd_id, s->pb2, x, &bcov, nbest, &ncand, NULL) == 2){ *why = (bcov > bc0 && bcov >= nbest)? 2 : 1; return (long)i; }
	}
	return (long)i;
}

/* sort (candidate, pair) records by ol descending with the reference permutation (the payload rides along) */
static void ref_sort_pairs_desc(u64 *c, u32 *cp, size_t n){
	typedef struct { u64 v; u32 idx; u32 pad; } rec_t; size_t i;
	rec_t *r = malloc((n + 1) * sizeof(rec_t)); u32 *np = malloc((n + 1) * 4);
	for(i=0;i<n;i++){ r[i].v = c[i]; r[i].idx = (u32)i; r[i].pad = 0; }
	ref_sort(r, n, sizeof(rec_t), gt_cand_ol_desc, NULL);
	for(i=0;i<n;i++){ c[i] = r[i].v; np[i] = cp[r[i].idx]; }
	memcpy(cp, np, n * 4);
	free(r); free(np);
}

/* replay of one read (wtzmo.c:803-1134) using the batch's device results.  Returns 0 when the read is finished, 1 when
 * the walk needs an alignment that has not been computed yet (the caller runs an on-demand wave and calls again). */
static int replay_read(wz_t *z, batch_t *b, bread_t *br, u32 bcov_in, readout_t *ro){
	const zparams_t *par = &z->par; const readset_t *rs = &z->rs; u32 pbid = br->rd_id;
	u32 alen = rs->reads.a[pbid].len, i;
	if(!br->walking){
		ro->rd_id = pbid;
		br->nbest = read_nbest(z, pbid); br->bcov = bcov_in;
		if(br->bcov >= br->nbest) return 0;
		if(br->skip){ fprintf(stderr, "wtzmo(b200): internal error: read %u needed but was skipped at batch build\n", pbid); exit(4); }
		filter_sort_candidates(z, pbid, &br->cands_raw, &br->cand_pair);
		if(z->rdhits){ u64v *h = &z->rdhits[pbid]; vec_clear(*h); vec_reserve(*h, br->cands_raw.n + 1); memcpy(h->a, br->cands_raw.a, br->cands_raw.n * 8); h->n = br->cands_raw.n; }
		if(par->dot_matrix){
			for(i=0;i<br->cands_raw.n;i++){
				u32 id2 = (u32)(br->cands_raw.a[i] >> 32), pi = br->cand_pair.a[i]; const zmo_dotres_t *r; u32 ol;
				if(pi == 0xFFFFFFFFU){ fprintf(stderr, "wtzmo(b200): internal error: pair (%u,%u) was not seeded\n", pbid, id2); exit(4); }
				if(b->seeds[pi].n_zpair * (u32)par->zsize < (u32)par->ztot) continue;
				r = &b->dots[pi];
				vec_push(ro->closed, pair_key(id2, pbid));
				ol = imax(r->qe - r->qb, r->te - r->tb);
				if(r->score >= par->min_score && r->score >= (int)(par->min_id * ol)){
					hit_t h; memset(&h, 0, sizeof(h));
					h.pb1 = pbid; h.pb2 = id2; h.dir2 = r->strand; h.score = r->score; h.tb = r->tb; h.te = r->te; h.qb = r->qb; h.qe = r->qe;
					h.mat = r->score; h.aln = ol; h.has_cigar = 0;
					vec_push(ro->hits, h);
				}
			}
			return 0;
		}
		build_seeds(z, b, pbid, br->cands_raw.a, br->cand_pair.a, br->cands_raw.n, &ro->seeds);
		if(!par->do_align) return 0;
		br->walking = 1; br->wi = 0; br->ncand = par->ncand;
	}
	for(i=br->wi;i<ro->seeds.n&&i<br->ncand;i++){
		seed_t *s = &ro->seeds.a[i]; hit_t h; const zmo_record_t *x; int act;
		if(s->closed){ br->ncand ++; continue; }
		x = b->pres[s->cand_idx];
		if(x == NULL || b->pdir[s->cand_idx] != s->dir){ br->wi = i; return 1; }      /* not aligned yet: ask for a wave */
		vec_push(ro->closed, pair_key(s->pb2, pbid));
		z->n_tasks_used ++;
		if(!x->ok){ s->closed = 1; br->ncand ++; continue; }
		if(x->score < par->min_score || x->mat < x->aln * par->min_id) continue;
		memset(&h, 0, sizeof(h));
		h.pb1 = pbid; h.pb2 = s->pb2; h.dir2 = s->dir; h.score = x->score; h.tb = x->tb; h.te = x->te; h.qb = x->qb; h.qe = x->qe;
		h.mat = x->mat; h.mis = x->mis; h.ins = x->ins; h.del = x->del; h.aln = x->aln; h.cigar = (const char*)b->pcig[s->cand_idx] + x->cigar_off; h.n_cigar = x->n_cigar; h.has_cigar = 1;
		vec_push(ro->hits, h);
		act = hit_rules(par, alen, rs->reads.a[s->pb2].len, pbid, s->pb2, x, &br->bcov, br->nbest, &br->ncand, &ro->masks);
		if(act == 2) break;
	}
	br->walking = 0;
	return 0;
}

static void batch_free(batch_t *b){
	size_t i;
	for(i=0;i<b->reads.n;i++){ vec_free(b->reads.a[i].cands_raw); vec_free(b->reads.a[i].cand_pair); }
	vec_free(b->reads); vec_free(b->pairs); vec_free(b->tasks);
	free(b->seeds); free(b->wins); free(b->pres); free(b->pcig); free(b->pdir); free(b->dots);     /* first-wave recs / cigars live in the session's pinned buffers */
	{ size_t k; for(k=0;k<b->extra.n;k++) free(b->extra.a[k]); vec_free(b->extra); }
	memset(b, 0, sizeof(*b));
}

/* pair list of the batch: state dependent, runs on the main thread */
static void batch_pairs(wz_t *z, batch_t *b){
	size_t i, k;
	/* pairs = candidates not closed as of now (the closed set only grows, so this is a superset of what the replay will ask for) */
	for(i=0;i<b->reads.n;i++){
		bread_t *r = &b->reads.a[i];
		vec_reserve(r->cand_pair, r->cands_raw.n + 1); r->cand_pair.n = r->cands_raw.n;
		for(k=0;k<r->cands_raw.n;k++){
			u32 id2 = (u32)(r->cands_raw.a[k] >> 32), ol = (u32)r->cands_raw.a[k];
			r->cand_pair.a[k] = 0xFFFFFFFFU;
			if(r->skip || ol == 0 || id2 >= z->rs.n_rd + z->rs.n_qr) continue;
			if(u64set_has(&z->closed, pair_key(r->rd_id, id2))) continue;
			{ zmo_pair_t p; p.qid = r->rd_id; p.cid = id2; r->cand_pair.a[k] = (u32)b->pairs.n; vec_push(b->pairs, p); }
		}
	}
}

/* the context's previous batch has been replayed: its result buffers are dead; grow the arena if the last batch spilled */
#define WZ_PIN_ARENA0 ((size_t)160 << 20)
static void pin_arena_reset(wz_t *z, int ci){
	z->pin[ci].used = 0;
	if(z->pin[ci].base == NULL || z->pin[ci].want > z->pin[ci].cap){
		size_t cap = z->pin[ci].want > WZ_PIN_ARENA0? z->pin[ci].want : WZ_PIN_ARENA0;
		if(z->pin[ci].base) zmo_host_free(z->pin[ci].base);
		z->pin[ci].base = zmo_host_alloc(cap); if(!z->pin[ci].base) die_zmo("zmo_host_alloc");
		z->pin[ci].cap = cap; z->pin[ci].want = cap;
	}
}

/* first guess of the CIGAR text of a wave, in 4-byte words: an alignment has at most min(len q, len c) + end overhangs columns and its text
 * measured 0.77 bytes per column on 15%-error reads; 1.1 bytes per base of the shorter read + slack covers it (a too small buffer costs a
 * re-run of the wave: ZMO_ERR_CAPACITY) */
static size_t wave_text_words(const wz_t *z, const batch_t *b, const zmo_task_t *tk, size_t n){
	size_t i; u64 bytes = 0;
	for(i=0;i<n;i++){ const zmo_pair_t *p = &b->pairs.a[tk[i].pair_idx]; u32 l1 = z->rs.reads.a[p->qid].len, l2 = z->rs.reads.a[p->cid].len; bytes += (u64)(l1 < l2? l1 : l2) * 11 / 10 + 64; }
	return (size_t)(bytes / 4) + (1u << 16);
}

/* device phases B + C for the batch (touches only the batch and the device context: may run on the worker thread) */
static void batch_compute(wz_t *z, batch_t *b){
	const zparams_t *par = &z->par; size_t i; int rc; u64 need = 0;
	zmo_ctx *ctx = z->ctxs[b->ci]; pthread_mutex_t *mu = &z->dev_mu[b->ci];
	if(b->pairs.n == 0) return;
	pthread_mutex_lock(&z->stat_mu); z->n_pairs_seeded += b->pairs.n; pthread_mutex_unlock(&z->stat_mu);
	if(par->dot_matrix){
		b->dots = malloc(b->pairs.n * sizeof(zmo_dotres_t)); b->seeds = calloc(b->pairs.n, sizeof(zmo_pairseed_t));
		pthread_mutex_lock(mu);
		rc = zmo_pair_dotmatrix(ctx, b->pairs.a, (u32)b->pairs.n, b->dots);
		pthread_mutex_unlock(mu);
		if(rc) die_zmo("zmo_pair_dotmatrix");
		for(i=0;i<b->pairs.n;i++) b->seeds[i].n_zpair = b->dots[i].n_zpair;
		return;
	}
	b->seeds = malloc(b->pairs.n * sizeof(zmo_pairseed_t));
	b->wins_cap = 64 * b->pairs.n + 1024; b->wins = malloc(b->wins_cap * sizeof(zmo_window_t));
	pthread_mutex_lock(mu);
	rc = zmo_pair_windows(ctx, b->slot, b->pairs.a, (u32)b->pairs.n, b->seeds, b->wins, b->wins_cap, &need);
	if(rc == ZMO_ERR_CAPACITY && need > b->wins_cap){ b->wins_cap = need + 16; b->wins = realloc(b->wins, b->wins_cap * sizeof(zmo_window_t)); rc = zmo_pair_windows(ctx, b->slot, b->pairs.a, (u32)b->pairs.n, b->seeds, b->wins, b->wins_cap, &need); }
	pthread_mutex_unlock(mu);
	if(rc) die_zmo("zmo_pair_windows");
	if(!par->do_align) return;
	b->pres = calloc(b->pairs.n, sizeof(*b->pres)); b->pcig = calloc(b->pairs.n, sizeof(*b->pcig)); b->pdir = calloc(b->pairs.n, 1);
	pin_arena_reset(z, b->ci);
	/* DP waves: every read's seeds in the order the replay is predicted to walk them (build_seeds on the candidate list
	 * as it was at batch build time); wave k aligns, for every read whose DRY walk over the results so far still stops at
	 * a missing alignment, the missing seeds up to the PREDICTED end of the walk plus a margin (predicted_stop; at most
	 * 32, 128, ... per read).  Most reads finish early (the first containing candidate masks them, wtzmo.c:1079-1090),
	 * so this skips most of the speculative DP.  Whatever the real replay still misses is computed on demand (demand_wave). */
	{
		size_t nr = b->reads.n, k, chunk = z->wave_margin < 0? (size_t)1 << 30 : (size_t)(z->wave_margin > 0? z->wave_margin : 16); int wave = 0;
		seedv *sv = calloc(nr + 1, sizeof(seedv)); long *pos = calloc(nr + 1, sizeof(long));
		for(i=0;i<nr;i++){
			bread_t *r = &b->reads.a[i]; u64v cc; u32v cp;
			pos[i] = -1;
			if(r->skip || r->cands_raw.n == 0) continue;
			vec_init(cc); vec_init(cp);
			for(k=0;k<r->cands_raw.n;k++) if(r->cand_pair.a[k] != 0xFFFFFFFFU){ vec_push(cc, r->cands_raw.a[k]); vec_push(cp, r->cand_pair.a[k]); }
			ref_sort_pairs_desc(cc.a, cp.a, cc.n);
			build_seeds(z, b, r->rd_id, cc.a, cp.a, cc.n, &sv[i]);
			pos[i] = dry_walk(z, b, r, &sv[i]);
			vec_free(cc); vec_free(cp);
		}
		while(1){
			VEC(zmo_task_t) tk; zmo_record_t *recs; u32 *cig; size_t cap;
			vec_init(tk);
			for(i=0;i<nr;i++){
				size_t got = 0, lim = sv[i].n;
				if(pos[i] < 0) continue;
				if(z->wave_predict >= 0){
					/* up to the seed at which the walk is predicted to stop, plus a margin for alignments that fail or end short; never fewer than what
					 * the dry walk is waiting for */
					int why = 0; const long ps = predicted_stop(z, b, &b->reads.a[i], &sv[i], &why);
					lim = (size_t)(ps < pos[i]? pos[i] : ps) + 1 + (size_t)(why == 2? z->wave_predict_b : z->wave_predict);
				}
				for(k=(size_t)pos[i];k<sv[i].n&&k<lim&&got<chunk;k++){
					const seed_t *sd = &sv[i].a[k]; zmo_task_t t;
					if(sd->closed || b->pres[sd->cand_idx]) continue;
					t.pair_idx = sd->cand_idx; t.dir = sd->dir; vec_push(tk, t); got ++;
				}
			}
			if(tk.n == 0){ vec_free(tk); break; }
			{ size_t nrd = 0; for(i=0;i<nr;i++) if(pos[i] >= 0) nrd ++; pthread_mutex_lock(&z->stat_mu); z->n_tasks += tk.n; { const int wi = wave < 7? wave : 7; z->wv_tasks[wi] += tk.n; z->wv_reads[wi] += nrd; z->wv_count[wi] ++; } pthread_mutex_unlock(&z->stat_mu); }
			{
				/* results land in the context's page-locked arena (or, if it is full, in pageable memory until the next batch has grown it): they
				 * stay valid until the batch has been replayed */
				const int ps = b->ci; const size_t want = z->no_cigar? 16 : wave_text_words(z, b, tk.a, tk.n), rbytes = (tk.n * sizeof(zmo_record_t) + 63) & ~(size_t)63;
				int pinned = z->pin[ps].used + rbytes + want * 4 <= z->pin[ps].cap;
				if(pinned){ recs = (zmo_record_t*)(z->pin[ps].base + z->pin[ps].used); cig = (u32*)(z->pin[ps].base + z->pin[ps].used + rbytes); cap = want; z->pin[ps].used += (rbytes + want * 4 + 63) & ~(size_t)63; }
				else {
					recs = malloc(tk.n * sizeof(zmo_record_t)); cap = want; cig = malloc(cap * 4); vec_push(b->extra, (void*)recs); vec_push(b->extra, (void*)cig);
					if(z->pin[ps].want < 2 * (z->pin[ps].used + rbytes + want * 4)) z->pin[ps].want = 2 * (z->pin[ps].used + rbytes + want * 4);
				}
				pthread_mutex_lock(mu);
				if(z->no_cigar) rc = zmo_pair_align_records(ctx, b->slot, tk.a, (u32)tk.n, recs);
				else rc = zmo_pair_align_text(ctx, b->slot, tk.a, (u32)tk.n, recs, (char*)cig, cap * 4, &need);
				if(rc == ZMO_ERR_CAPACITY && need > cap * 4){
					/* the guess was too small: the text of this wave goes to pageable memory, the wave is run again */
					cap = (need + need / 4) / 4 + 16;
					cig = malloc(cap * 4); vec_push(b->extra, (void*)cig);
					if(z->pin[ps].want < z->pin[ps].cap + 2 * cap * 4) z->pin[ps].want = z->pin[ps].cap + 2 * cap * 4;
					rc = zmo_pair_align_text(ctx, b->slot, tk.a, (u32)tk.n, recs, (char*)cig, cap * 4, &need);
				}
				pthread_mutex_unlock(mu);
			}
			if(rc) die_zmo("zmo_pair_align");
			for(i=0;i<tk.n;i++){ b->pres[tk.a[i].pair_idx] = &recs[i]; b->pcig[tk.a[i].pair_idx] = cig; b->pdir[tk.a[i].pair_idx] = (u8)tk.a[i].dir; }
			vec_free(tk);
			for(i=0;i<nr;i++) if(pos[i] >= 0) pos[i] = dry_walk(z, b, &b->reads.a[i], &sv[i]);
			wave ++; if(chunk < ((size_t)1 << 20)) chunk *= (size_t)z->wave_growth;
		}
		for(i=0;i<nr;i++) vec_free(sv[i]);
		free(sv); free(pos);
	}
}

/* on-demand DP wave for the read being replayed: every not-yet-aligned seed from the current walk position on */
static void demand_wave(wz_t *z, batch_t *b, bread_t *br, readout_t *ro){
	VEC(zmo_task_t) tk; size_t i, cap; u64 need = 0; int rc; zmo_record_t *recs; u32 *cig;
	vec_init(tk);
	for(i=br->wi;i<ro->seeds.n;i++){
		seed_t *s = &ro->seeds.a[i]; zmo_task_t t;
		if(s->closed) continue;
		if(b->pres[s->cand_idx] && b->pdir[s->cand_idx] == s->dir) continue;
		t.pair_idx = s->cand_idx; t.dir = s->dir; vec_push(tk, t);
		if(tk.n >= (size_t)(br->nbest - br->bcov) + (br->nbest - br->bcov) / 4 + 8) break;
	}
	if(tk.n == 0){ fprintf(stderr, "wtzmo(b200): internal error: empty demand wave\n"); exit(4); }
	recs = malloc(tk.n * sizeof(zmo_record_t)); cap = 4096 * tk.n + 65536; cig = malloc(cap * 4);
	pthread_mutex_lock(&z->dev_mu[b->ci]);
	if(z->no_cigar) rc = zmo_pair_align_records(z->ctxs[b->ci], b->slot, tk.a, (u32)tk.n, recs);
	else rc = zmo_pair_align_text(z->ctxs[b->ci], b->slot, tk.a, (u32)tk.n, recs, (char*)cig, cap * 4, &need);
	if(rc == ZMO_ERR_CAPACITY && need > cap * 4){ cap = need / 4 + 16; cig = realloc(cig, cap * 4); rc = zmo_pair_align_text(z->ctxs[b->ci], b->slot, tk.a, (u32)tk.n, recs, (char*)cig, cap * 4, &need); }
	pthread_mutex_unlock(&z->dev_mu[b->ci]);
	if(rc) die_zmo("zmo_pair_align (demand wave)");
	for(i=0;i<tk.n;i++){ b->pres[tk.a[i].pair_idx] = &recs[i]; b->pcig[tk.a[i].pair_idx] = cig; b->pdir[tk.a[i].pair_idx] = (u8)tk.a[i].dir; }
	vec_push(b->extra, (void*)recs); vec_push(b->extra, (void*)cig);
	z->n_waves ++; z->n_wave_tasks += tk.n; z->n_tasks += tk.n;
	vec_free(tk);
}

typedef struct { wz_t *z; batch_t *b; } wk_arg_t;
static void* batch_compute_thread(void *arg){ wk_arg_t *wa = arg; batch_compute(wa->z, wa->b); free(wa); return NULL; }

static void run_overlap(wz_t *z){
	const zparams_t *par = &z->par; readset_t *rs = &z->rs; u32 j, beg, end, pbbeg = 0, pbend = 0, i_idx; u64 tot = 0; readout_t ro; zmo_index_stats_t st;
	memset(&ro, 0, sizeof(ro)); ro.rd_id = 0xFFFFFFFFU;
	if(rs->n_qr == 0 && rs->n_rd){ for(j=0;j<rs->n_rd;j++) tot += rs->reads.a[j].len; z->avg_rdlen = (u32)(tot / rs->n_rd); }
	else if(rs->n_qr){ for(j=0;j<rs->n_qr;j++) tot += rs->reads.a[j + rs->n_rd].len; z->avg_rdlen = (u32)(tot / rs->n_qr); }
	else z->avg_rdlen = 10000;
	if(z->avg_rdlen == 0) z->avg_rdlen = 1;
	if(par->n_idx > 1) z->rdhits = calloc(rs->n_rd + rs->n_qr, sizeof(u64v));
	z->kcut = par->kcut;
	if(rs->n_qr == 0){ beg = 0; end = rs->n_rd; } else { beg = rs->n_rd; end = beg + rs->n_qr; }
	const double tl0 = now_s(); int tl_first = 1;
	memset(z->tl, 0, sizeof(z->tl));
	for(i_idx=0;i_idx<(u32)par->n_idx;i_idx++){
		double t0 = now_s();
		pbbeg = pbend; pbend = pbbeg + (rs->n_rd + par->n_idx - 1) / par->n_idx;
		fprintf(stderr, "[wtzmo-b200] indexing %u/%u\n", i_idx + 1, par->n_idx);
		if(zmo_index_build(z->ctx, pbbeg, pbend, &z->kcut, &st)) die_zmo("zmo_index_build");
		fprintf(stderr, "[wtzmo-b200] - average kmer depth = %u\n[wtzmo-b200] - %llu high frequency kmers (>=%u)\n[wtzmo-b200] - indexing %llu kmers, %llu postings (%.3f s)\n", st.kavg,
			(unsigned long long)st.n_filtered_high, st.kcut, (unsigned long long)st.n_indexed, (unsigned long long)st.n_postings, now_s() - t0);
		if(i_idx + 1 >= (u32)par->n_idx) break;
		/* just_query passes (wtzmo.c:1289-1301): candidates of every eligible read against this partition; bcov is 0 there */
		for(j=0;j<rs->n_rd;){
			batch_t b; size_t i; memset(&b, 0, sizeof(b));
			for(;j<rs->n_rd&&b.reads.n<(size_t)z->batch_reads*8;j++){
				bread_t r; memset(&r, 0, sizeof(r));
				if((j % par->n_job) != (u32)par->i_job) continue;
				if(z->masked[j]) continue;
				r.rd_id = j; vec_push(b.reads, r);
			}
			batch_candidates(z, &b);
			for(i=0;i<b.reads.n;i++){ bread_t *r = &b.reads.a[i]; filter_sort_candidates(z, r->rd_id, &r->cands_raw, NULL); { u64v *h = &z->rdhits[r->rd_id]; vec_clear(*h); vec_reserve(*h, r->cands_raw.n + 1); memcpy(h->a, r->cands_raw.a, r->cands_raw.n * 8); h->n = r->cands_raw.n; } }
			batch_free(&b);
		}
	}
	{
		/* software pipeline: up to `depth` batches are in flight on their own device contexts (one worker thread each)
		 * while the host replays the oldest one.  A batch is built from the state as of its build time, i.e. batch k+depth
		 * from the state at the end of batch k-1, which only makes the speculation set larger (results are pure). */
		batch_t *q[WZ_MAX_CTX]; int qh = 0, qn = 0, ci; double t0, t1; unsigned n_built = 0;
		memset(z->ctx_busy, 0, sizeof(z->ctx_busy));
		j = beg; z->tl[0] = now_s() - tl0;
		while(1){
			size_t i; batch_t *cur;
			while(qn < z->depth && j < end){
				size_t est_pairs = 0; batch_t *nxt = calloc(1, sizeof(batch_t));
				for(ci=0;ci<z->n_ctx;ci++) if(!z->ctx_busy[ci]) break;
				if(ci == z->n_ctx){ fprintf(stderr, "wtzmo(b200): internal error: no free device context\n"); exit(4); }
				nxt->ci = ci; nxt->slot = 0;
				/* batch size: full batches in the steady state, smaller ones while the pipeline fills (the first replay can start sooner) and
				 * drains (the last reads are spread over all contexts instead of leaving the GPU to one of them) */
				size_t target = (size_t)z->batch_reads;
				if(z->ramp){
					size_t left = (size_t)(end - j) / (size_t)par->n_job + 1, up = (size_t)z->ramp << (n_built < 8? n_built : 8);
					if(up < target) target = up;
					{ size_t t2 = left / (size_t)z->drain_div; if(t2 < (size_t)z->drain_min) t2 = (size_t)z->drain_min; if(t2 < target) target = t2; }      /* geometric drain: every batch takes 1/drain_div of what is left */
				}
				n_built ++;
				for(;j<end&&nxt->reads.n<target&&est_pairs<(size_t)z->batch_pairs;j++){
					bread_t r; memset(&r, 0, sizeof(r));
					if((j % par->n_job) != (u32)par->i_job) continue;
					if(z->masked[j]) continue;
					r.rd_id = j; r.bcov0 = z->rdcovs[j]; r.skip = r.bcov0 >= read_nbest(z, j);
					vec_push(nxt->reads, r);
					if(!r.skip) est_pairs += 40;
				}
				if(nxt->reads.n == 0){ free(nxt); break; }
				z->ctx_busy[ci] = 1;
				t0 = now_s();
				batch_candidates(z, nxt);
				batch_pairs(z, nxt);
				{
					/* a device call takes at most 65,536 pairs (32,768 in dot-matrix mode): give the tail of an over-full batch
					 * back to the read cursor (only reached with large -A / very deep coverage) */
					const size_t lim = z->call_pairs? (size_t)z->call_pairs : (par->dot_matrix? 32768 : 65536);
					if(nxt->pairs.n > lim){
						size_t k, cum = 0, keep = 0, kr = 0;
						for(k=0;k<nxt->reads.n;k++){
							bread_t *r = &nxt->reads.a[k]; size_t c = 0, q;
							for(q=0;q<r->cand_pair.n;q++) if(r->cand_pair.a[q] != 0xFFFFFFFFU) c ++;
							if(cum + c > lim && k > 0) break;
							cum += c; keep = cum; kr = k + 1;
						}
						if(keep > lim){ fprintf(stderr, "wtzmo(b200): read with %zu candidate pairs exceeds the per-call limit of %zu (lower -A)\n", keep, lim); exit(1); }
						j = nxt->reads.a[kr].rd_id;
						for(k=kr;k<nxt->reads.n;k++){ vec_free(nxt->reads.a[k].cands_raw); vec_free(nxt->reads.a[k].cand_pair); }
						nxt->reads.n = kr; nxt->pairs.n = keep;
					}
				}
				if(z->depth > 1){
					wk_arg_t *wa = malloc(sizeof(wk_arg_t)); wa->z = z; wa->b = nxt;
					if(pthread_create(&nxt->th, NULL, batch_compute_thread, wa) != 0){ free(wa); batch_compute(z, nxt); } else nxt->have_thread = 1;
				} else batch_compute(z, nxt);
				z->t_dev += now_s() - t0;
				q[(qh + qn) % WZ_MAX_CTX] = nxt; qn ++;
				if(z->tl[1] == 0) z->tl[1] = now_s() - tl0;
				z->tl[3] = now_s() - tl0;
			}
			if(qn == 0) break;
			cur = q[qh]; qh = (qh + 1) % WZ_MAX_CTX; qn --;
			if(cur->have_thread){ double tj = now_s(); pthread_join(cur->th, NULL); z->t_dev += now_s() - tj; }
			t1 = now_s();
			if(tl_first){ z->tl[2] = t1 - tl0; tl_first = 0; }
			z->tl[4] = t1 - tl0;
			z->n_reads_batched += cur->reads.n;
			for(i=0;i<cur->reads.n;i++){
				bread_t *br = &cur->reads.a[i];
				if(z->masked[br->rd_id]){ z->n_reads_late_masked ++; z->n_pairs_late_masked += br->cands_raw.n; continue; }       /* checked BEFORE the previous read's masks are merged (wtzmo.c:1315 vs 1322) */
				flush_read(z, &ro, 0);
				while(replay_read(z, cur, br, z->rdcovs[br->rd_id], &ro)) demand_wave(z, cur, br, &ro);
			}
			flush_read(z, &ro, 1);  /* hits point into this batch's CIGAR buffers: print them before they are reused; masks stay pending */
			z->t_replay += now_s() - t1; z->n_batches ++;
			z->ctx_busy[cur->ci] = 0;
			batch_free(cur); free(cur);
		}
	}
	flush_read(z, &ro, 0);
	ob_flush(z);
	z->tl[5] = now_s() - tl0;
	vec_free(ro.hits); vec_free(ro.masks); vec_free(ro.closed); vec_free(ro.seeds);
}

/* ------------------------------------------------------------------ command line (wtzmo.c:1422-1812) */
static int file_exists(const char *f){ struct stat st; return stat(f, &st) == 0; }
static int usage(void){
	printf(
	"WTZMO: Overlaper of long reads using homopolymer compressed k-mer seeding\n"
	"SMARTdenovo: Ultra-fast de novo assembler for high noisy long reads\n"
	"B200 build: hot path on NVIDIA Blackwell (sm_100a); output identical to the reference `wtzmo -t 1`\n"
	"Usage: wtzmo [options]\n"
	"Options:\n"
	" -t <int>    Number of threads, [1] (accepted for compatibility; the GPU replaces the worker pool)\n"
	" -P <int>    Total parallel jobs, [1]\n"
	" -p <int>    Index of current job (0-based), [0]\n"
	" -i <string> Long reads sequences file, + *\n"
	" -I <string> Long reads sequence file, DON'T build index on them, +\n"
	" -b <string> Long reads retained region, often from wtobt/wtcyc, +\n"
	" -J <int>    Jack knife of original read length, [0]\n"
	" -L <string> Load pairs of read name from file, will avoid to calculate overlap them again, + [NULL]\n"
	" -o <string> Output file of alignments, *\n"
	" -9 <string> Record pairs of sequences have beed aligned regardless of successful, including pairs from '-L'\n"
	" -f          Force overwrite\n"
	" -H <int>    Option of homopolymer compression, [3]\n"
	" -k <int>    Kmer size, 5 <= <-k> <= 32, [16]\n"
	" -K <int>    Filter high frequency kmers, maybe repetitive, [0]\n"
	" -d <int>    Minimum size of total seeding region for kmer windows, [300]\n"
	" -S <int>    Subsampling kmers, 1/<-S> kmers are indexed, [4]\n"
	" -G <int>    Build kmer index in multiple iterations to save memory, 1: once, [1]\n"
	" -z <int>    Smaller kmer size (z-mer), 5 <= <-z> <= 16, [10]\n"
	" -Z <int>    Filter high frequency z-mers, maybe repetitive, [64]\n"
	" -U <float>  Ultra-fast dot matrix alignment (five values, or -U -1 for the defaults 128 64 160 1.0 0.05)\n"
	" -y <int>    Zmer window, [800]\n"
	" -R <int>    Minimum size of seeding region within zmer window, [200]\n"
	" -r <int>    Minimum size of total seeding region for zmer windows, [300]\n"
	" -l <int>    Maximum variant of uncompressed sizes between two matched hz-kmer, [2]\n"
	" -q <int>    THreshold of seed-window coverage along query, [100]\n"
	" -A <int>    Limit number of best candidates per read, [500]\n"
	" -B <int>    Limit number of best overlaps per read, [100]\n"
	" -C          Don't write the <output>.contained file\n"
	" -F <string> Reads from this file(s) are to be exclued, one line for one read name, + [NULL]\n"
	" -M <int>    Alignment penalty: match, [2]\n"
	" -X <int>    Alignment penalty: mismatch, [-5]\n"
	" -O <int>    Alignment penalty: insertion or deletion, [-3]\n"
	" -E <int>    Alignment penalty: gap extension, [-1]\n"
	" -T <int>    Alignment penalty: read end clipping, [-50]\n"
	" -w <int>    Minimum bandwidth, iteratively doubled to maximum [50]\n"
	" -W <int>    Maximum bandwidth, [3200]\n"
	" -e <int>    Maximum bandwidth at ending extension, [800]\n"
	" -s <int>    Minimum alignment score, [200]\n"
	" -m <float>  Minimum alignment identity, [0.5]\n"
	" -n          Refine the alignment\n"
	" -v          Verbose (accepted, ignored)\n"
	"Environment: ZMO_DEVICE (GPU ordinal, default 0; under torchrun LOCAL_RANK), ZMO_GPUS=n|all (GPU g runs job -P P*n -p p*n+g, records gathered\n"
	"             over NCCL and written in job order), ZMO_DEVICES=a,b,.. (their ordinals), ZMO_BATCH_READS, ZMO_BATCH_PAIRS, ZMO_STATS=file;\n"
	"             ZMO_OVL_COLS=16 (print the 16 columns `cut -f1-16` keeps: no CIGAR column);\n"
	"             A/B switches: ZMO_WA_BRIDGE=0 (window alignment by the sequential kernel), ZMO_SEED_LANES=G (G pairs per warp in the seeding kernel)\n"
	"\n");
	return 1;
}

/* ------------------------------------------------------------------ session API (used by main() and, through ctypes, by bench.py)
 * wz_open   parse the wtzmo command line, load + sort reads, side inputs, create the device context
 * wz_upload pack-upload the reads to the device (zmo_reads_upload)
 * wz_run    one complete overlap job `-P n_job -p i_job` from a clean state (index build included), .ovl to out_path
 * wz_stats  numbers of the last wz_run */
typedef struct {
	wz_t z; char *output, *pairoutf; int device, uploaded, is_fork; u8 *masked0; u64v closed0;
	double t_open, last_overlap_s, last_upload_s; u64 upload_bytes;
} wz_session_t;

/* device context(s) of a session, created on a helper thread while the main thread parses the reads: CUDA start-up (context creation,
 * module load) costs about as much as loading a 0.5 GB FASTA */
typedef struct { wz_t *z; int device, refine, rc; zmo_params_t zp; char err[600]; pthread_t th; int started; } ctx_boot_t;
static void* ctx_boot_thread(void *arg){
	ctx_boot_t *b = arg; wz_t *z = b->z; int q;
	b->rc = 0; b->err[0] = 0;
	if(zmo_ctx_create(&z->ctx, b->device, &b->zp)){ snprintf(b->err, sizeof(b->err), "zmo_ctx_create: %s", zmo_last_error()); b->rc = 3; return NULL; }
	if(b->refine && zmo_set_refine(z->ctx, 1)){ snprintf(b->err, sizeof(b->err), "zmo_set_refine: %s", zmo_last_error()); b->rc = 3; return NULL; }
	/* one context per queued batch: those in flight + the one being replayed (which may still ask for on-demand waves) */
	z->ctxs[0] = z->ctx; z->n_ctx = 1;
	for(q=1;q<z->depth;q++){ if(zmo_ctx_clone(z->ctx, &z->ctxs[z->n_ctx])){ snprintf(b->err, sizeof(b->err), "zmo_ctx_clone: %s", zmo_last_error()); b->rc = 3; return NULL; } z->n_ctx ++; }
	if(z->par.do_align && !z->par.dot_matrix) for(q=0;q<z->n_ctx;q++) pin_arena_reset(z, q);      /* page-locked result arenas, behind the FASTA parse */
	return NULL;
}
static int ctx_boot_join(ctx_boot_t *b){
	if(b->started){ pthread_join(b->th, NULL); b->started = 0; }
	if(b->rc) fprintf(stderr, "wtzmo(b200): %s\n", b->err);
	return b->rc;
}

wz_session_t* wz_open(int argc, char **argv, int *rc_out){
	wz_session_t *S = calloc(1, sizeof(wz_session_t)); wz_t *z = &S->z; zparams_t *par = &z->par; int c; float optval; zmo_params_t zp;
	char *env; double t_start = now_s(); ctx_boot_t boot;
	VEC(char*) pbs, flts, ovls, obts, tbas; u32 i; size_t k;
	zparams_default(par);
	vec_init(pbs); vec_init(flts); vec_init(ovls); vec_init(obts); vec_init(tbas);
	*rc_out = 0; optind = 1;
	setenv("CUDA_MODULE_LOADING", "EAGER", 0);      /* load every kernel with the context (helper thread, behind the FASTA parse) instead of at its first launch */
	while((c = getopt(argc, argv, "ht:P:p:Ni:b:J:I:o:9:S:fCH:k:G:z:Z:U:y:d:r:q:l:K:A:B:r:R:L:F:W:w:e:M:X:O:E:T:s:m:nv")) != -1){
		switch(c){
			case 'h': *rc_out = usage(); return NULL;
			case 't': par->ncpu = atoi(optarg); break;
			case 'P': par->n_job = atoi(optarg); break;
			case 'p': par->i_job = atoi(optarg); break;
			case 'N': par->do_align = 0; break;
			case 'i': vec_push(pbs, optarg); break;
			case 'b': vec_push(obts, optarg); break;
			case 'J': par->min_rdlen = atoi(optarg); break;
			case 'I': vec_push(tbas, optarg); break;
			case 'o': S->output = optarg; break;
			case '9': S->pairoutf = optarg; break;
			case 'S': par->ksave = atoi(optarg); break;
			case 'f': par->overwrite = 1; break;
			case 'C': par->write_contained = 0; break;   /* -C never reaches wt->skip_contained (wtzmo.c:168,1609,1781) */
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
			default: *rc_out = usage(); return NULL;
		}
	}
	if(S->output == NULL){ *rc_out = usage(); return NULL; }
	if(!par->overwrite && strcmp(S->output, "-") && file_exists(S->output)){ fprintf(stderr, "File exists! '%s'\n\n", S->output); *rc_out = usage(); return NULL; }
	if(pbs.n == 0 || par->ksize > 32 || par->ksize < 5 || par->zsize > 16 || par->zsize < 5 || par->ksave < 1){ *rc_out = usage(); return NULL; }
	if(par->n_job < 1 || par->n_idx < 1){ fprintf(stderr, "wtzmo(b200): -P and -G must be >= 1\n"); *rc_out = 2; return NULL; }
	par->max_overhang = 2 * par->xvar;
	par->kstep = par->kwin / 2;
	if((env = getenv("ZMO_DEVICE"))) S->device = atoi(env); else if((env = getenv("LOCAL_RANK"))) S->device = atoi(env);
	z->batch_reads = (env = getenv("ZMO_BATCH_READS"))? atoi(env) : 384;
	z->batch_pairs = (env = getenv("ZMO_BATCH_PAIRS"))? atoi(env) : 40000;
	z->depth = 1 + ((env = getenv("ZMO_DEPTH"))? atoi(env) : 2);      /* ZMO_DEPTH = batches in flight on the device while the host replays the oldest */
	if((env = getenv("ZMO_PIPELINE")) && atoi(env) == 0) z->depth = 1;  /* no pipeline: one batch at a time, one context */
	if(z->depth < 1) z->depth = 1;
	if(z->depth > WZ_MAX_CTX) z->depth = WZ_MAX_CTX;
	z->call_pairs = (env = getenv("ZMO_CALL_PAIRS"))? atoi(env) : 0;     /* test hook: pairs per device call (0 = the library's limits) */
	z->wave_margin = (env = getenv("ZMO_WAVE0"))? atoi(env) : 32;      /* most seeds of a read in the first DP wave (x ZMO_WAVE_GROWTH per wave); < 0: align every seed up front.  cfg2: 8 without the prediction below */
	z->wave_growth = (env = getenv("ZMO_WAVE_GROWTH"))? atoi(env) : 4; if(z->wave_growth < 2) z->wave_growth = 2;
	z->no_cigar = (env = getenv("ZMO_OVL_COLS")) && atoi(env) == 16;
	z->wave_predict = (env = getenv("ZMO_WAVE_PREDICT"))? atoi(env) : 2;
	z->wave_predict_b = (env = getenv("ZMO_WAVE_PREDICT_B"))? atoi(env) : z->wave_predict; if(z->wave_predict_b < 0) z->wave_predict_b = 0;      /* margin when the walk is predicted to end on the dovetail count */      /* >= 0: a wave aligns a read's seeds up to the predicted end of its walk + this margin (predict_rec); -1: fixed chunk per read.  cfg2: 85,432 -> 72,108 alignments issued for 64,797 consumed, 834 -> 761 ms per shard */
	z->drain_div = (env = getenv("ZMO_DRAIN_DIV"))? atoi(env) : z->depth; if(z->drain_div < 1) z->drain_div = 1;
	z->drain_min = (env = getenv("ZMO_DRAIN_MIN"))? atoi(env) : 48; if(z->drain_min < 1) z->drain_min = 1;
	z->ramp = (env = getenv("ZMO_RAMP"))? atoi(env) : 96;     /* first batch size of the pipeline ramp (doubles per batch up to ZMO_BATCH_READS); 0 = off.  cfg2: 1,025 -> 938 ms per shard */
	{ int q; for(q=0;q<WZ_MAX_CTX;q++) pthread_mutex_init(&z->dev_mu[q], NULL); pthread_mutex_init(&z->stat_mu, NULL); }
	if(z->batch_reads < 1) z->batch_reads = 1;
	memset(&zp, 0, sizeof(zp));
	zp.hk = par->hk; zp.hz = par->hz; zp.ksize = par->ksize; zp.zsize = par->zsize; zp.ksave = par->ksave; zp.kovl = par->kovl; zp.zcut = par->zcut; zp.kvar = par->kvar;
	zp.kwin = par->kwin; zp.kstep = par->kstep; zp.zovl = par->zovl; zp.ztot = par->ztot; zp.w = par->w; zp.ew = par->ew; zp.W = par->W;
	zp.M = par->M; zp.X = par->X; zp.O = par->O; zp.E = par->E; zp.T = par->T; zp.min_id = par->min_id;
	zp.xvar = par->xvar; zp.yvar = par->yvar; zp.min_block_len = par->min_block_len; zp.max_overhang = par->max_overhang; zp.deviation_penalty = par->deviation_penalty; zp.gap_penalty = par->gap_penalty;
	memset(&boot, 0, sizeof(boot)); boot.z = z; boot.device = S->device; boot.refine = par->refine; boot.zp = zp;
	if(pthread_create(&boot.th, NULL, ctx_boot_thread, &boot) == 0) boot.started = 1; else ctx_boot_thread(&boot);
	fprintf(stderr, "[wtzmo-b200] loading long reads\n");
	rs_load(&z->rs, pbs.a, (int)pbs.n, par->min_rdlen, 0);
	ref_sort(z->rs.reads.a, z->rs.reads.n, sizeof(read_t), gt_read_len_desc, NULL);      /* wtzmo.c:1708 */
	if(tbas.n) rs_load(&z->rs, tbas.a, (int)tbas.n, par->min_rdlen, 1);
	fprintf(stderr, "[wtzmo-b200] Done, %u reads (+%u query-only), %.3f s\n", z->rs.n_rd, z->rs.n_qr, now_s() - t_start);
	if(z->rs.n_rd == 0){ fprintf(stderr, "wtzmo(b200): no reads\n"); ctx_boot_join(&boot); *rc_out = 1; return NULL; }
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
	/* remember the initial state so that wz_run can start every job from it */
	S->masked0 = malloc(z->rs.n_rd + z->rs.n_qr + 1); memcpy(S->masked0, z->masked, z->rs.n_rd + z->rs.n_qr + 1);
	vec_init(S->closed0);
	for(k=0;k<z->closed.cap;k++) if(z->closed.tab[k] != ~0ULL) vec_push(S->closed0, z->closed.tab[k]);
	if((*rc_out = ctx_boot_join(&boot))) return NULL;
	S->t_open = now_s() - t_start;
	vec_free(pbs); vec_free(flts); vec_free(ovls); vec_free(obts); vec_free(tbas);
	return S;
}

int wz_upload(wz_session_t *S){
	wz_t *z = &S->z; u32 i, n = z->rs.n_rd + z->rs.n_qr; u64 *off = malloc((size_t)n * 8); u32 *len = malloc((size_t)n * 4); double t0 = now_s(); int rc;
	for(i=0;i<n;i++){ off[i] = z->rs.reads.a[i].off; len[i] = z->rs.reads.a[i].len; }
	rc = zmo_reads_upload(z->ctx, z->rs.bits, z->rs.nbases, off, len, n);
	free(off); free(len);
	if(rc){ fprintf(stderr, "wtzmo(b200): zmo_reads_upload: %s\n", zmo_last_error()); return 3; }
	S->uploaded = 1; S->last_upload_s = now_s() - t0; S->upload_bytes = ((z->rs.nbases + 31) / 32) * 8 + (u64)n * 12;
	return 0;
}

/* one complete overlap job `-P n_job -p i_job` from a clean state, records written to fp */
int wz_run_fp(wz_session_t *S, int n_job, int i_job, FILE *fp){
	wz_t *z = &S->z; size_t k, n = z->rs.n_rd + z->rs.n_qr; double t0;
	if(!S->uploaded){ int rc = wz_upload(S); if(rc) return rc; }
	if(n_job < 1 || i_job < 0 || i_job >= n_job) return 2;
	z->par.n_job = n_job; z->par.i_job = i_job;
	memcpy(z->masked, S->masked0, n + 1); memset(z->rdcovs, 0, (n + 1) * sizeof(u32));
	free(z->closed.tab); u64set_init(&z->closed);
	for(k=0;k<S->closed0.n;k++) u64set_add(&z->closed, S->closed0.a[k]);
	if(z->rdhits){ for(k=0;k<n;k++) vec_free(z->rdhits[k]); free(z->rdhits); z->rdhits = NULL; }
	z->n_records = z->aln_cols = z->n_tasks = z->n_tasks_used = z->n_pairs_seeded = z->n_batches = z->n_reads_batched = z->n_reads_late_masked = z->n_pairs_late_masked = z->n_waves = z->n_wave_tasks = 0; z->t_dev = z->t_replay = 0;
	memset(z->wv_tasks, 0, sizeof(z->wv_tasks)); memset(z->wv_reads, 0, sizeof(z->wv_reads)); memset(z->wv_count, 0, sizeof(z->wv_count));
	z->out = fp;
	if(z->obuf == NULL){ z->obuf_cap = 8u << 20; z->obuf = malloc(z->obuf_cap); }
	t0 = now_s();
	run_overlap(z);
	fflush(fp);
	S->last_overlap_s = now_s() - t0;
	if(getenv("ZMO_WAVE_DEBUG")){ int w; for(w=0;w<8;w++) if(z->wv_count[w]) fprintf(stderr, "[wtzmo-b200] DP wave %d: %llu launches, %llu reads, %llu alignments\n", w, (unsigned long long)z->wv_count[w], (unsigned long long)z->wv_reads[w], (unsigned long long)z->wv_tasks[w]); }
	return 0;
}
int wz_run(wz_session_t *S, int n_job, int i_job, const char *out_path){
	FILE *fp = strcmp(out_path, "-")? fopen(out_path, "w") : stdout; int rc;
	if(fp == NULL){ fprintf(stderr, "wtzmo(b200): cannot open %s\n", out_path); return 1; }
	rc = wz_run_fp(S, n_job, i_job, fp);
	if(fp != stdout) fclose(fp);
	return rc;
}

/* a second session on another GPU over the SAME parsed read set (shared, read-only) and the same initial state: own device
 * context(s), own replay state.  Multi-GPU runs: GPU g of n is the reference's job `-P n -p g` (wtzmo.c:1291,1314). */
wz_session_t* wz_fork(const wz_session_t *S0, int device, int *rc_out){
	wz_session_t *S = calloc(1, sizeof(wz_session_t)); wz_t *z = &S->z; const wz_t *z0 = &S0->z; zmo_params_t zp; const zparams_t *par = &z0->par; size_t n = z0->rs.n_rd + z0->rs.n_qr; int q;
	*rc_out = 0;
	z->rs = z0->rs; z->par = z0->par;      /* reads: shared pointers, never written after wz_open */
	z->batch_reads = z0->batch_reads; z->batch_pairs = z0->batch_pairs; z->depth = z0->depth; z->call_pairs = z0->call_pairs;
	z->wave_margin = z0->wave_margin; z->wave_growth = z0->wave_growth; z->wave_predict = z0->wave_predict; z->wave_predict_b = z0->wave_predict_b; z->no_cigar = z0->no_cigar; z->ramp = z0->ramp; z->drain_div = z0->drain_div; z->drain_min = z0->drain_min;
	for(q=0;q<WZ_MAX_CTX;q++) pthread_mutex_init(&z->dev_mu[q], NULL);
	pthread_mutex_init(&z->stat_mu, NULL);
	z->masked = calloc(n + 1, 1); z->rdcovs = calloc(n + 1, sizeof(u32)); u64set_init(&z->closed);
	S->masked0 = S0->masked0; S->closed0 = S0->closed0; S->output = S0->output; S->pairoutf = S0->pairoutf; S->device = device; S->is_fork = 1;
	memset(&zp, 0, sizeof(zp));
	zp.hk = par->hk; zp.hz = par->hz; zp.ksize = par->ksize; zp.zsize = par->zsize; zp.ksave = par->ksave; zp.kovl = par->kovl; zp.zcut = par->zcut; zp.kvar = par->kvar;
	zp.kwin = par->kwin; zp.kstep = par->kstep; zp.zovl = par->zovl; zp.ztot = par->ztot; zp.w = par->w; zp.ew = par->ew; zp.W = par->W;
	zp.M = par->M; zp.X = par->X; zp.O = par->O; zp.E = par->E; zp.T = par->T; zp.min_id = par->min_id;
	zp.xvar = par->xvar; zp.yvar = par->yvar; zp.min_block_len = par->min_block_len; zp.max_overhang = par->max_overhang; zp.deviation_penalty = par->deviation_penalty; zp.gap_penalty = par->gap_penalty;
	if(zmo_ctx_create(&z->ctx, device, &zp)){ fprintf(stderr, "wtzmo(b200): zmo_ctx_create(device %d): %s\n", device, zmo_last_error()); *rc_out = 3; return NULL; }
	if(par->refine && zmo_set_refine(z->ctx, 1)){ *rc_out = 3; return NULL; }
	z->ctxs[0] = z->ctx; z->n_ctx = 1;
	for(q=1;q<z->depth;q++){ if(zmo_ctx_clone(z->ctx, &z->ctxs[z->n_ctx])){ fprintf(stderr, "wtzmo(b200): zmo_ctx_clone: %s\n", zmo_last_error()); *rc_out = 3; return NULL; } z->n_ctx ++; }
	return S;
}

#define WZ_STATS_N 40   /* doubles written by wz_stats (callers size their buffer with wz_stats_n()) */
int wz_stats_n(void){ return WZ_STATS_N; }
void wz_stats(wz_session_t *S, double *out){
	wz_t *z = &S->z; double ms[12], m1[12]; uint64_t ct[8], c1[8]; int i, q; u64 nl = 0;
	memset(ms, 0, sizeof(ms)); memset(ct, 0, sizeof(ct)); for(i=0;i<WZ_STATS_N;i++) out[i] = 0;
	for(q=0;q<z->n_ctx;q++){ zmo_stage_ms(z->ctxs[q], m1); zmo_counters(z->ctxs[q], c1); for(i=0;i<12;i++) ms[i] += m1[i]; for(i=0;i<8;i++) ct[i] += c1[i]; nl += zmo_kernel_launches(z->ctxs[q]); }
	out[0] = (double)z->n_records; out[1] = (double)z->aln_cols; out[2] = S->last_overlap_s; out[3] = z->t_dev; out[4] = z->t_replay; out[5] = (double)z->n_batches;
	out[6] = (double)z->n_pairs_seeded; out[7] = (double)z->n_tasks; out[8] = (double)z->n_tasks_used; out[9] = (double)nl;
	for(i=0;i<8;i++) out[10 + i] = ms[i];
	for(i=0;i<7;i++) out[18 + i] = (double)ct[i];
	out[25] = (double)(z->rs.n_rd + z->rs.n_qr); out[26] = (double)z->rs.nbases; out[27] = S->last_upload_s; out[28] = (double)S->upload_bytes;
	out[29] = (double)z->n_waves; out[30] = (double)z->n_wave_tasks; out[31] = ms[8];
	out[32] = (double)z->n_reads_batched; out[33] = (double)z->n_reads_late_masked; out[34] = (double)z->n_pairs_late_masked;
	{ int k; for(k=0;k<5;k++) out[35 + k] = z->tl[k + (k >= 0)]; }      /* timeline [1..5] of the last job (s): first batch built, first replay, last batch built, last compute joined, end */      /* speculation: reads batched / masked by the time of their turn / their candidates */
}

void wz_close(wz_session_t *S){ if(S){ int q; for(q=0;q<WZ_MAX_CTX;q++) if(S->z.pin[q].base) zmo_host_free(S->z.pin[q].base); for(q=S->z.n_ctx-1;q>=0;q--) if(S->z.ctxs[q]) zmo_ctx_destroy(S->z.ctxs[q]); free(S); } }

/* ------------------------------------------------------------------ multi-GPU job: ZMO_GPUS=n
 * One process, one host thread per GPU.  GPU g runs the reference's job `-P (P*n) -p (p*n+g)` (query reads with rd_id % (P*n) == p*n+g
 * against the full index, private masked / tried-pair / counter state, wtzmo.c:1291,1314) into a memory stream; no communication during
 * compute.  At the end the record text of all jobs is gathered on GPU 0 over NVLink with NCCL (zmo_gather_records) and written in job
 * order, which is what `cat` of the per-job files gives (usage, wtzmo.c:1431-1433).  <out>.contained = union of the jobs' masked reads in
 * read-id order; -9 = union of the jobs' tried pairs. */
typedef struct { wz_session_t *S; int n_job, i_job, rc; char *buf; size_t len; } mg_arg_t;
typedef struct { zmo_ctx *ctxs[16]; int n, rc; } mg_prep_t;
static void* mg_prepare_thread(void *arg){ mg_prep_t *p = arg; p->rc = zmo_gather_prepare(p->ctxs, p->n); return NULL; }
static void* mg_thread(void *arg){
	mg_arg_t *a = arg; FILE *fp = open_memstream(&a->buf, &a->len);
	if(fp == NULL){ a->rc = 1; return NULL; }
	a->rc = wz_upload(a->S);
	if(!a->rc) a->rc = wz_run_fp(a->S, a->n_job, a->i_job, fp);
	fclose(fp);
	return NULL;
}
#define WZ_MAX_GPU 16
typedef struct { wz_session_t *S[WZ_MAX_GPU]; int n; double gather_ms, gather_wall_s; u64 gathered_bytes; } wz_multi_t;
static int run_multi(wz_session_t *S0, int ngpu, const int *devs, wz_multi_t *M){
	mg_arg_t a[WZ_MAX_GPU]; pthread_t th[WZ_MAX_GPU]; int g, rc = 0; zmo_ctx *ctxs[WZ_MAX_GPU]; const void *parts[WZ_MAX_GPU]; uint64_t sizes[WZ_MAX_GPU], total = 0; char *out; double t0;
	FILE *fp;
	memset(M, 0, sizeof(*M)); M->n = ngpu; M->S[0] = S0;
	for(g=0;g<ngpu;g++){ int h; if(devs[g] < 0 || devs[g] >= zmo_device_count()){ fprintf(stderr, "wtzmo(b200): ZMO_GPUS=%d: no GPU %d (%d visible)\n", ngpu, devs[g], zmo_device_count()); return 3; }
		for(h=0;h<g;h++) if(devs[h] == devs[g]){ fprintf(stderr, "wtzmo(b200): ZMO_DEVICES names GPU %d twice: one job per GPU\n", devs[g]); return 2; } }
	for(g=1;g<ngpu;g++){ M->S[g] = wz_fork(S0, devs[g], &rc); if(M->S[g] == NULL) return rc; }
	for(g=0;g<ngpu;g++){ memset(&a[g], 0, sizeof(a[g])); a[g].S = M->S[g]; a[g].n_job = S0->z.par.n_job * ngpu; a[g].i_job = S0->z.par.i_job * ngpu + g; }
	for(g=0;g<ngpu;g++) if(pthread_create(&th[g], NULL, mg_thread, &a[g]) != 0){ fprintf(stderr, "wtzmo(b200): cannot start the thread of GPU %d\n", devs[g]); return 3; }
	{	/* NCCL communicators while the jobs compute */
		mg_prep_t prep; pthread_t pt; int have;
		prep.n = ngpu; prep.rc = 0; for(g=0;g<ngpu;g++) prep.ctxs[g] = M->S[g]->z.ctx;
		have = pthread_create(&pt, NULL, mg_prepare_thread, &prep) == 0;
		for(g=0;g<ngpu;g++){ pthread_join(th[g], NULL); if(a[g].rc) rc = a[g].rc; }
		if(have) pthread_join(pt, NULL);
	}
	if(rc) return rc;
	for(g=0;g<ngpu;g++){ ctxs[g] = M->S[g]->z.ctx; parts[g] = a[g].buf; sizes[g] = a[g].len; total += a[g].len; }
	out = malloc(total + 1);
	t0 = now_s();
	if(zmo_gather_records(ctxs, ngpu, parts, sizes, out, total, &total, &M->gather_ms)){ fprintf(stderr, "wtzmo(b200): zmo_gather_records: %s\n", zmo_last_error()); return 3; }
	M->gather_wall_s = now_s() - t0; M->gathered_bytes = total;
	fp = strcmp(S0->output, "-")? fopen(S0->output, "w") : stdout;
	if(fp == NULL){ fprintf(stderr, "wtzmo(b200): cannot open %s\n", S0->output); return 1; }
	fwrite(out, 1, total, fp);
	if(fp != stdout) fclose(fp); else fflush(stdout);
	free(out);
	for(g=0;g<ngpu;g++) free(a[g].buf);
	return 0;
}

#ifndef WTZMO_LIB
int main(int argc, char **argv){
	int rc = 0, ngpu = 1, devs[WZ_MAX_GPU], g; double t_start = now_s(); char *env; u32 i;
	wz_session_t *S = wz_open(argc, argv, &rc); wz_t *z; wz_multi_t M;
	if(S == NULL) return rc;
	z = &S->z;
	memset(&M, 0, sizeof(M)); M.n = 1; M.S[0] = S;
	if((env = getenv("ZMO_GPUS"))){ ngpu = !strcmp(env, "all")? zmo_device_count() : atoi(env); if(ngpu < 1) ngpu = 1; if(ngpu > WZ_MAX_GPU) ngpu = WZ_MAX_GPU; }
	for(g=0;g<ngpu;g++) devs[g] = S->device + g;
	if((env = getenv("ZMO_DEVICES"))){ char *cp = strdup(env), *sv, *tok; for(g=0,tok=strtok_r(cp, ",", &sv);tok&&g<ngpu;tok=strtok_r(NULL, ",", &sv)) devs[g++] = atoi(tok); free(cp); }
	if(ngpu > 1){
		fprintf(stderr, "[wtzmo-b200] calculating overlaps on %d GPUs: GPU g runs job -P %d -p %d+g\n", ngpu, z->par.n_job * ngpu, z->par.i_job * ngpu);
		if((rc = run_multi(S, ngpu, devs, &M))) return rc;
	} else {
		if((rc = wz_upload(S))) return rc;
		fprintf(stderr, "[wtzmo-b200] calculating overlaps on GPU %d\n", S->device);
		if((rc = wz_run(S, z->par.n_job, z->par.i_job, S->output))) return rc;
	}
	for(g=0;g<M.n;g++){
		wz_session_t *Sg = M.S[g]; wz_t *zg = &Sg->z;
		fprintf(stderr, "[wtzmo-b200] %sDone, %llu records, %llu aligned columns, %.3f s (device calls %.3f s, replay+format %.3f s, %llu batches, %llu pairs seeded, %llu aligned, %llu consumed)\n", M.n > 1? "job " : "",
			(unsigned long long)zg->n_records, (unsigned long long)zg->aln_cols, Sg->last_overlap_s, zg->t_dev, zg->t_replay, (unsigned long long)zg->n_batches,
			(unsigned long long)zg->n_pairs_seeded, (unsigned long long)zg->n_tasks, (unsigned long long)zg->n_tasks_used);
	}
	if(M.n > 1) fprintf(stderr, "[wtzmo-b200] gathered %llu bytes of records on GPU %d over NCCL: %.3f ms on the device, %.3f s incl. staging\n", (unsigned long long)M.gathered_bytes, devs[0], M.gather_ms, M.gather_wall_s);
	if(z->par.write_contained && strcmp(S->output, "-")){
		char *maskf = malloc(strlen(S->output) + 16); FILE *mf;
		sprintf(maskf, "%s.contained", S->output); mf = fopen(maskf, "w");
		for(i=0;i<z->rs.n_rd;i++){ int m = 0; for(g=0;g<M.n;g++) m |= M.S[g]->z.masked[i]; if(m) fprintf(mf, "%s\n", z->rs.reads.a[i].name); }
		fclose(mf); free(maskf);
	}
	if(S->pairoutf){
		FILE *pf = fopen(S->pairoutf, "w"); size_t k; u64set_t seen; u64set_init(&seen);
		for(g=0;g<M.n;g++){
			const u64set_t *cl = &M.S[g]->z.closed;
			for(k=0;k<cl->cap;k++){
				u64 v = cl->tab[k];
				if(v == ~0ULL || (M.n > 1 && u64set_has(&seen, v))) continue;
				if(M.n > 1) u64set_add(&seen, v);
				fprintf(pf, "%s\t%s\n", z->rs.reads.a[v >> 33].name, z->rs.reads.a[(v & 0xFFFFFFFFU) >> 1].name);
			}
		}
		fclose(pf); free(seen.tab);
	}
	if((env = getenv("ZMO_STATS"))){
		FILE *sf = fopen(env, "w"); double st[WZ_STATS_N], s1[WZ_STATS_N]; const char *nm[8] = {"index", "candidates", "pair_windows", "window_align", "gap_global", "end_extend", "dotmatrix", "copy"};
		const char *cn[7] = {"cells_ext", "cells_win", "cells_gap", "zpairs", "postings", "h2d_bytes", "d2h_bytes"}; int k; double ovl_max = 0;
		u64 rb = 0, rl = 0, cl = 0; double al[6];
		zmo_alloc_stats(al);
		memset(st, 0, sizeof(st));
		for(g=0;g<M.n;g++){ wz_stats(M.S[g], s1); for(k=0;k<WZ_STATS_N;k++) if(k != 2 && k != 25 && k != 26) st[k] += s1[k]; if(s1[2] > ovl_max) ovl_max = s1[2]; st[25] = s1[25]; st[26] = s1[26];
			rb += M.S[g]->z.n_reads_batched; rl += M.S[g]->z.n_reads_late_masked; cl += M.S[g]->z.n_pairs_late_masked; }
		st[2] = ovl_max;
		if(sf){
			fprintf(sf, "{\"records\": %.0f, \"aligned_cols\": %.0f, \"overlap_s\": %.6f, \"total_s\": %.6f, \"device_call_s\": %.6f, \"replay_s\": %.6f, \"batches\": %.0f, \"pairs_seeded\": %.0f, \"tasks\": %.0f, \"tasks_used\": %.0f, \"launches\": %.0f, \"stage_ms\": {",
				st[0], st[1], st[2], now_s() - t_start, st[3], st[4], st[5], st[6], st[7], st[8], st[9]);
			for(k=0;k<8;k++) fprintf(sf, "%s\"%s\": %.3f", k? ", " : "", nm[k], st[10 + k]);
			fprintf(sf, "}, \"counters\": {");
			for(k=0;k<7;k++) fprintf(sf, "%s\"%s\": %.0f", k? ", " : "", cn[k], st[18 + k]);
			fprintf(sf, "}, \"n_reads\": %.0f, \"n_bases\": %.0f, \"reads_batched\": %llu, \"reads_late_masked\": %llu, \"cands_late_masked\": %llu, \"n_gpus\": %d, \"gather_device_ms\": %.3f, \"gather_wall_s\": %.6f, \"gathered_bytes\": %llu, \"load_s\": %.3f, \"alloc\": {\"device_s\": %.3f, \"device_calls\": %.0f, \"device_bytes\": %.0f, \"pinned_s\": %.3f, \"pinned_calls\": %.0f, \"pinned_bytes\": %.0f}}\n", st[25], st[26],
				(unsigned long long)rb, (unsigned long long)rl, (unsigned long long)cl, M.n, M.gather_ms, M.gather_wall_s, (unsigned long long)M.gathered_bytes, S->t_open, al[0], al[1], al[2], al[3], al[4], al[5]);
			fclose(sf);
		}
	}
	for(g=M.n-1;g>=0;g--) wz_close(M.S[g]);
	return 0;
}
#endif
