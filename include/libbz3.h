/*
 * libbz3.h -- C ABI of the B200 block codec.  Binary-compatible with the reference header
 * (kspalaiologos/bzip3 v1.5.2, include/libbz3.h): same symbol names, argument order, integer types
 * and error numbers, so a program built against the reference header links against
 * libbzip3_b200.so unchanged.  Line numbers below cite the reference declaration each entry replaces.
 *
 * Every function that touches block data runs it through CUDA kernels on the current device; there
 * is no CPU implementation behind this header.  bz3_new() returns NULL when no usable GPU exists.
 */
#ifndef LIBBZ3_H
#define LIBBZ3_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define BZIP3_API __attribute__((visibility("default")))
#else
#define BZIP3_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* error numbers: reference include/libbz3.h:47-55 */
enum {
    BZ3_OK = 0,
    BZ3_ERR_OUT_OF_BOUNDS = -1,
    BZ3_ERR_BWT = -2,
    BZ3_ERR_CRC = -3,
    BZ3_ERR_MALFORMED_HEADER = -4,
    BZ3_ERR_TRUNCATED_DATA = -5,
    BZ3_ERR_DATA_TOO_BIG = -6,
    BZ3_ERR_INIT = -7,
    BZ3_ERR_DATA_SIZE_TOO_SMALL = -8
};

struct bz3_state; /* opaque; owns device buffers, a stream and pinned scratch */

BZIP3_API const char *bz3_version(void);                                   /* :62  */
BZIP3_API int8_t bz3_last_error(struct bz3_state *state);                  /* :67  */
BZIP3_API const char *bz3_strerror(struct bz3_state *state);               /* :72  */
BZIP3_API struct bz3_state *bz3_new(int32_t block_size);                   /* :79  65 KiB .. 511 MiB */
BZIP3_API void bz3_free(struct bz3_state *state);                          /* :84  */
BZIP3_API size_t bz3_bound(size_t input_size);                             /* :89  n + n/50 + 32 */

/* frame API, :101 and :110 */
BZIP3_API int bz3_compress(uint32_t block_size, const uint8_t *in, uint8_t *out, size_t in_size, size_t *out_size);
BZIP3_API int bz3_decompress(const uint8_t *in, uint8_t *out, size_t in_size, size_t *out_size);

BZIP3_API size_t bz3_min_memory_needed(int32_t block_size);                /* :167 host-equivalent figure */

/* block API, :176 and :194.  `buffer` is caller-owned HOST memory transformed in place. */
BZIP3_API int32_t bz3_encode_block(struct bz3_state *state, uint8_t *buffer, int32_t size);
BZIP3_API int32_t bz3_decode_block(struct bz3_state *state, uint8_t *buffer, size_t buffer_size,
                                   int32_t compressed_size, int32_t orig_size);

/* batch API, :206 and :212: n blocks at once, one host thread + one CUDA stream per block */
BZIP3_API void bz3_encode_blocks(struct bz3_state *states[], uint8_t *buffers[], int32_t sizes[], int32_t n);
BZIP3_API void bz3_decode_blocks(struct bz3_state *states[], uint8_t *buffers[], size_t buffer_sizes[],
                                 int32_t sizes[], int32_t orig_sizes[], int32_t n);

BZIP3_API int bz3_orig_size_sufficient_for_decode(const uint8_t *block, size_t block_size, int32_t orig_size); /* :235 */

#ifdef __cplusplus
}
#endif
#endif
