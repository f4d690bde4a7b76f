/*
 * gen_reads.c -- deterministic synthetic long-read generator.
 *
 *   gen_reads -n <reads> -L <mean_len> -G <genome_len> [-m pacbio|ont] [-s seed] [-o out.fa]
 *             [-r repeat_frac(0.05)] [-c <n_contained_extra>]
 *
 * Genome: i.i.d. uniform ACGT of length G; a fraction -r of it is overwritten with copies of
 * 2-6 kb elements at 2-10% divergence (exercises the -K filter and the repeat weighting).
 * Reads: uniform start, strand Bernoulli(1/2), length clamp(round(N(L, 0.15 L)), 1000, 2^24-1)
 * (and <= G).  Error models (per reference base):
 *   pacbio: ins 8.25% / del 4.5% / sub 2.25%  (15%, 55/30/15)
 *   ont   : ins 3% / del 5% / sub 4% (12%), indel rates doubled inside homopolymer runs >= 3
 * Output: FASTA, one line per sequence, names r0, r1, ... in generation order, pure ACGT.
 * PRNG: xoshiro256** seeded by splitmix64 -- identical output on every platform.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

static uint64_t s[4];
static inline uint64_t rotl(uint64_t x, int k){ return (x << k) | (x >> (64 - k)); }
static inline uint64_t rnd64(void){
	uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
	s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
	return r;
}
static void seed_rng(uint64_t x){
	int i;
	for(i=0;i<4;i++){ uint64_t z = (x += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; s[i] = z ^ (z >> 31); }
}
static inline double rndu(void){ return (rnd64() >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t rndn(uint64_t n){ return (uint64_t)(rndu() * n); }
static double rnd_gauss(void){
	double u1 = rndu(), u2 = rndu();
	if(u1 < 1e-300) u1 = 1e-300;
	return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

int main(int argc, char **argv){
	long n = 2000, L = 8000, G = 500000, i, j;
	uint64_t seed = 20240601;
	const char *model = "pacbio", *outf = NULL;
	double rfrac = 0.05, pins, pdel, psub;
	int ont;
	for(i=1;i<argc;i++){
		if(!strcmp(argv[i], "-n") && i + 1 < argc) n = atol(argv[++i]);
		else if(!strcmp(argv[i], "-L") && i + 1 < argc) L = atol(argv[++i]);
		else if(!strcmp(argv[i], "-G") && i + 1 < argc) G = atol(argv[++i]);
		else if(!strcmp(argv[i], "-m") && i + 1 < argc) model = argv[++i];
		else if(!strcmp(argv[i], "-s") && i + 1 < argc) seed = strtoull(argv[++i], NULL, 10);
		else if(!strcmp(argv[i], "-o") && i + 1 < argc) outf = argv[++i];
		else if(!strcmp(argv[i], "-r") && i + 1 < argc) rfrac = atof(argv[++i]);
		else { fprintf(stderr, "usage: gen_reads -n N -L len -G genome [-m pacbio|ont] [-s seed] [-r repfrac] [-o out.fa]\n"); return 1; }
	}
	ont = !strcmp(model, "ont");
	if(ont){ pins = 0.03; pdel = 0.05; psub = 0.04; } else { pins = 0.0825; pdel = 0.045; psub = 0.0225; }
	seed_rng(seed);
	FILE *out = outf? fopen(outf, "w") : stdout;
	if(out == NULL){ perror("open"); return 1; }
	char *g = malloc(G + 1);
	for(i=0;i<G;i++) g[i] = "ACGT"[rnd64() >> 62];
	/* repeat family: elements of 2-6 kb copied with 2-10% divergence until rfrac of G is covered */
	if(rfrac > 0 && G > 20000){
		long covered = 0, elen = 2000 + (long)rndn(4001), src = (long)rndn(G - elen);
		char *elem = malloc(elen);
		memcpy(elem, g + src, elen);
		while(covered < (long)(rfrac * G)){
			double dv = 0.02 + 0.08 * rndu();
			long dst = (long)rndn(G - elen);
			for(j=0;j<elen;j++){
				char c = elem[j];
				if(rndu() < dv) c = "ACGT"[rnd64() >> 62];
				g[dst + j] = c;
			}
			covered += elen;
			if(rndu() < 0.2){ /* start a new family now and then */
				elen = 2000 + (long)rndn(4001); src = (long)rndn(G - elen);
				elem = realloc(elem, elen); memcpy(elem, g + src, elen);
			}
		}
		free(elem);
	}
	long maxlen = 1L << 24; maxlen -= 1;
	char *rd = malloc(4 * (size_t)(maxlen < G? maxlen : G) + 64);
	char *seg = malloc((size_t)(maxlen < G? maxlen : G) + 1);
	for(i=0;i<n;i++){
		long len = (long)floor(L + 0.15 * L * rnd_gauss() + 0.5);
		if(len < 1000) len = 1000;
		if(len > maxlen) len = maxlen;
		if(len > G) len = G;
		long st = (long)rndn(G - len + 1);
		int rev = (int)(rnd64() >> 63);
		if(rev){ for(j=0;j<len;j++){ char c = g[st + len - 1 - j]; seg[j] = c == 'A'? 'T' : (c == 'C'? 'G' : (c == 'G'? 'C' : 'A')); } }
		else memcpy(seg, g + st, len);
		long m = 0; int run = 0;
		for(j=0;j<len;j++){
			double mul = 1.0, u;
			if(j && seg[j] == seg[j-1]) run ++; else run = 1;
			if(ont && run >= 3) mul = 2.0;
			/* insertion(s) before this base */
			while(rndu() < pins * mul && m < 4 * len) rd[m++] = "ACGT"[rnd64() >> 62];
			u = rndu();
			if(u < pdel * mul) continue;
			if(u < pdel * mul + psub){ char c; do { c = "ACGT"[rnd64() >> 62]; } while(c == seg[j]); rd[m++] = c; }
			else rd[m++] = seg[j];
		}
		if(m > maxlen) m = maxlen;
		rd[m] = 0;
		fprintf(out, ">r%ld\n%s\n", i, rd);
	}
	if(outf) fclose(out);
	free(g); free(rd); free(seg);
	return 0;
}
