/*
 * rle.h -- leaf-block byte format shared with lh3/ropebwt2 (reference rle.h).
 *
 * The engine stores the BWT in HBM in exactly this format, so callers written against the
 * reference (main.c:292-313 decodes blocks inline with rle_nptr/rle_dec1) keep working.
 * Only the format-level pieces a *caller* needs are provided here: the block header
 * accessor, the one-run decoder/encoder, and rle_count/rle_print as small host utilities.
 * The reference's in-place block mutators (rle_insert, rle_insert_cached, rle_split,
 * rle_rank2a; reference rle.c) have no host implementation in this library: that work is
 * what the CUDA kernels in ropebwt2_b200/csrc/rb2_engine.cu do.
 *
 * A block is  [uint16 nbytes][run][run]...  A run is 1, 2, 4 or 8 bytes (reference
 * rle.h:53-75): the low 3 bits of the first byte are the symbol, the length is spread over
 * the remaining bits of the first byte and 6 bits of every continuation byte (10xxxxxx).
 */
#ifndef RB2_RLE_H_
#define RB2_RLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* add the symbol counts of one block to cnt[6] (reference rle.c:109-118) */
void rle_count(const uint8_t *block, int64_t cnt[6]);
/* print one block, expanded or as symbol/length pairs (reference rle.c:120-132) */
void rle_print(const uint8_t *block, int expand);

#ifdef __cplusplus
}
#endif

#define RLE_MIN_SPACE 18 /* a leaf is split once nbytes + 18 exceeds the block length (reference rope.c:143) */

/* number of run bytes in a block */
#define rle_nptr(block) ((uint16_t*)(block))

/* decode the run at p into (c, l) and advance p; same contract as the reference macro (rle.h:39-51) */
#define rle_dec1(p, c, l) do { \
		const uint8_t rb2_h_ = *(p); \
		(c) = rb2_h_ & 7; \
		if (rb2_h_ < 0x80) { (l) = rb2_h_ >> 3; (p) += 1; } \
		else if (rb2_h_ < 0xE0) { (l) = ((int64_t)(rb2_h_ & 0x18) << 3) | ((p)[1] & 0x3f); (p) += 2; } \
		else { \
			int rb2_n_ = (rb2_h_ & 0x10)? 8 : 4, rb2_i_; \
			int64_t rb2_x_ = (rb2_h_ >> 3) & 1; \
			for (rb2_i_ = 1; rb2_i_ < rb2_n_; ++rb2_i_) rb2_x_ = rb2_x_ << 6 | ((p)[rb2_i_] & 0x3f); \
			(l) = rb2_x_; (p) += rb2_n_; \
		} \
	} while (0)

/* encode run (c, l) at p; returns the number of bytes written (reference rle.h:53-75) */
static inline int rle_enc1(uint8_t *p, int c, int64_t l)
{
	int n, i;
	if (l < 16) { p[0] = (uint8_t)(l << 3 | c); return 1; }
	if (l < 256) { p[0] = (uint8_t)(0xC0 | (l >> 6) << 3 | c); p[1] = (uint8_t)(0x80 | (l & 0x3f)); return 2; }
	n = l < (1LL << 19)? 4 : 8;
	p[0] = (uint8_t)((n == 4? 0xE0 : 0xF0) | (l >> (6 * (n - 1))) << 3 | c);
	for (i = 1; i < n; ++i) p[i] = (uint8_t)(0x80 | ((l >> (6 * (n - 1 - i))) & 0x3f));
	return n;
}

#endif
