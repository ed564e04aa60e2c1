/*
 * gen_reads.c -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A fast C restatement of ropebwt2_b200/synth.py:hash_reads (the counter-based generators of the
 * full-size workloads) used to FEED THE REFERENCE BINARY: it writes the
 * reads as `-L` text lines (main.c:180-186).  tests/test_synth.py pins it to the numpy and torch
 * versions.  pthreads over reads.
 */
#include <stdint.h>
#include <stddef.h>
#include <pthread.h>

static inline uint64_t mix64(uint64_t x)
{
	uint64_t z = x + 0x9E3779B97F4A7C15ULL;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

#define K_GENOME 0x1000003D1ULL
#define K_START  0x2000005A7ULL
#define K_STRAND 0x30000071BULL
#define K_ERR    0x400000963ULL
#define K_BASE   0x500000B3FULL

/* kind 0 = U, 1 = G.  Reads a..b-1, L bases each; out gets (b-a)*(L+1) bytes: when `lines` the
 * characters ACGT + '\n', otherwise nt6 codes 1..4 in forward orientation + 0. */
typedef struct { int kind, L, lines; int64_t a, r0, r1; uint64_t seed, gseed, glen, err_thr; uint8_t *out; } job_t;

static void *gen_part(void *p_)
{
	job_t *p = (job_t*)p_;
	int kind = p->kind, L = p->L, lines = p->lines;
	uint64_t seed = p->seed, gseed = p->gseed, glen = p->glen, err_thr = p->err_thr;
	int64_t r;
	for (r = p->r0; r < p->r1; ++r) {
		uint8_t *o = p->out + (size_t)(r - p->a) * (L + 1);
		int j;
		if (kind == 0) {
			for (j = 0; j < L; ++j)
				o[j] = 1 + (mix64(seed * K_BASE + (uint64_t)r * L + j) >> 62);
		} else {
			uint64_t start = (mix64(seed * K_START + r) >> 1) % (glen - L + 1);
			int rev = mix64(seed * K_STRAND + r) >> 63;
			for (j = 0; j < L; ++j) {
				uint64_t pos = rev? start + (L - 1) - j : start + j;
				uint8_t base = 1 + (mix64(gseed * K_GENOME + pos) >> 62);
				uint64_t e = mix64(seed * K_ERR + (uint64_t)r * L + j);
				if (rev) base = 5 - base;
				if ((e & 0xFFFFFF) < err_thr)
					base = (uint8_t)((base - 1 + 1 + ((e >> 24) & 0xFFFF) % 3) % 4 + 1);
				o[j] = base;
			}
		}
		if (lines) {
			for (j = 0; j < L; ++j) o[j] = "$ACGTN"[o[j]];
			o[L] = '\n';
		} else o[L] = 0;
	}
	return 0;
}

void gen_reads(int kind, int64_t a, int64_t b, int L, uint64_t seed, uint64_t gseed, uint64_t glen,
               uint64_t err_thr, uint8_t *out, int lines, int n_threads)
{
	pthread_t tid[64];
	job_t job[64];
	int t;
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 64) n_threads = 64;
	for (t = 0; t < n_threads; ++t) {
		job_t j = { kind, L, lines, a, a + (b - a) * t / n_threads, a + (b - a) * (t + 1) / n_threads, seed, gseed, glen, err_thr, out };
		job[t] = j;
		pthread_create(&tid[t], 0, gen_part, &job[t]);
	}
	for (t = 0; t < n_threads; ++t) pthread_join(tid[t], 0);
}
