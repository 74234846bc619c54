/*
 * bz3_b200.h -- extensions of the B200 block codec beyond the reference ABI (libbz3.h).
 * Plain C types only.  None of these exist in the reference; they expose
 *   - device-resident operation (input uploaded once, codec timed without PCIe traffic),
 *   - per-stage device timings (CUDA events on the state's stream),
 *   - single-stage entry points used by the parity tests.
 */
#ifndef BZ3_B200_H
#define BZ3_B200_H

#include <stddef.h>
#include <stdint.h>
#include "libbz3.h"

#ifdef __cplusplus
extern "C" {
#endif

/* stage indices for bz3_b200_stage_ms() */
enum {
    BZ3_STAGE_H2D = 0, BZ3_STAGE_CRC, BZ3_STAGE_RLE, BZ3_STAGE_LZP, BZ3_STAGE_BWT, BZ3_STAGE_CM, BZ3_STAGE_D2H,
    BZ3_STAGE_COUNT
};

/* number of CUDA devices visible / device a state lives on */
BZIP3_API int bz3_b200_device_count(void);
BZIP3_API int bz3_b200_state_device(struct bz3_state *state);
/* Placement of new states.  Default 1: bz3_new() uses the calling thread's current device (one process per GPU).
 * devices = N (<= 0: all visible): bz3_new() deals states round-robin over N GPUs starting at the current device, so
 * that the reference's batch calls (bz3_encode_blocks / bz3_decode_blocks with n states, src/libbz3.c:845-870; the
 * -j loop of src/main.c:336-378) run block i on GPU i mod N.  Same as the environment variable BZ3_B200_DEVICES=N|all.
 * Returns the number of devices in effect. */
BZIP3_API int bz3_b200_set_devices(int devices);
/* device memory a state owns (three block buffers + LZP table) ... */
BZIP3_API size_t bz3_b200_device_bytes(struct bz3_state *state);
/* ... and the stage workspaces (mRLE / suffix sort / inverse BWT scratch, 48 B per byte of block) that all states of
 * the state's device share: BZ3_B200_ARENAS of them (default 2), leased per stage call, sized for the largest live
 * state, released with the last state.  Replaces the per-state `sais_array` of src/libbz3.c:498-504, 549-551. */
BZIP3_API size_t bz3_b200_workspace_bytes(struct bz3_state *state);

/* Device-resident codec: stage `size` bytes into the state, run the block codec without host copies,
 * fetch the result.  encode_resident/decode_resident return what bz3_encode_block/bz3_decode_block
 * would return and set last_error the same way. */
BZIP3_API int bz3_b200_upload(struct bz3_state *state, const uint8_t *host, int32_t size);
BZIP3_API int32_t bz3_b200_encode_resident(struct bz3_state *state, int32_t size);
BZIP3_API int32_t bz3_b200_decode_resident(struct bz3_state *state, int32_t compressed_size, int32_t orig_size);
BZIP3_API int bz3_b200_download(struct bz3_state *state, uint8_t *host, int32_t size);
/* n states at once (one host thread per state); results[i] receives the per-block return value */
BZIP3_API void bz3_b200_encode_resident_many(struct bz3_state *states[], int32_t sizes[], int32_t results[], int32_t n);
BZIP3_API void bz3_b200_decode_resident_many(struct bz3_state *states[], int32_t csizes[], int32_t osizes[],
                                             int32_t results[], int32_t n);

/* accumulated device milliseconds per stage since the last reset; launches = kernels launched */
BZIP3_API void bz3_b200_stats_reset(struct bz3_state *state);
BZIP3_API double bz3_b200_stage_ms(struct bz3_state *state, int stage, int decode);
BZIP3_API uint64_t bz3_b200_kernel_launches(struct bz3_state *state);
/* details of the last suffix sort: records sorted over all radix passes, rounds, device ms in passes */
BZIP3_API void bz3_b200_last_sort_stats(struct bz3_state *state, uint64_t *records, int32_t *rounds, double *ms);

/* single stages on HOST buffers (parity tests).  Each returns the stage's own return value. */
BZIP3_API uint32_t bz3_b200_stage_crc(struct bz3_state *state, const uint8_t *in, int32_t n);
BZIP3_API int32_t bz3_b200_stage_rle_encode(struct bz3_state *state, const uint8_t *in, int32_t n, uint8_t *out);
BZIP3_API int bz3_b200_stage_rle_decode(struct bz3_state *state, const uint8_t *in, int32_t maxin, uint8_t *out,
                                        int32_t outlen);
BZIP3_API int32_t bz3_b200_stage_lzp_encode(struct bz3_state *state, const uint8_t *in, int32_t n, uint8_t *out);
BZIP3_API int32_t bz3_b200_stage_lzp_decode(struct bz3_state *state, const uint8_t *in, int32_t n, uint8_t *out,
                                            int32_t max);
BZIP3_API int32_t bz3_b200_stage_bwt(struct bz3_state *state, const uint8_t *in, int32_t n, uint8_t *out);
BZIP3_API int32_t bz3_b200_stage_unbwt(struct bz3_state *state, const uint8_t *in, int32_t n, int32_t idx, uint8_t *out);
BZIP3_API int32_t bz3_b200_stage_cm_encode(struct bz3_state *state, const uint8_t *in, int32_t n, uint8_t *out);
BZIP3_API int bz3_b200_stage_cm_decode(struct bz3_state *state, const uint8_t *in, int32_t insize, uint8_t *out,
                                       int32_t n);
/* The ".bz3" container of the reference's command line tool (src/main.c:157-482: "BZ3v1", s32 LE block size, then per
 * block s32 LE coded size, s32 LE original size, coded bytes) over file descriptors, with a deep block queue instead of
 * the reference's read-J / code-J / write-J batches (:352-478, J <= 64): a reader, `in_flight` workers (one state and
 * stream each) and an in-order writer, each running as soon as its slot is ready.  in_flight <= 0 picks
 * min(SM count, what device memory and 16 GiB of pinned host buffers hold).  The bytes written equal `bzip3 -e -b <block_size>`'s.
 * decode_fd with out_fd < 0 only tests (`bzip3 -t`).  Return 0, a BZ3_ERR_* of the failing block (blocks before it have
 * been written, as by the reference's loop), or one of the codes below. */
#define BZ3_B200_ERR_IO (-20)         /* read / write failed */
#define BZ3_B200_ERR_SIGNATURE (-21)  /* "Invalid signature." (src/main.c:186) */
#define BZ3_B200_ERR_BLOCK_SIZE (-24) /* block size outside 65 KiB .. 511 MiB (:195) */
#define BZ3_B200_ERR_HEADERS (-22)    /* "Inconsistent headers." (:265) */
#define BZ3_B200_ERR_TRUNCATED (-23)  /* file ends inside a block (xread_noeof, :262-270) */
BZIP3_API int bz3_b200_encode_fd(int in_fd, int out_fd, int32_t block_size, int in_flight, uint64_t *bytes_in, uint64_t *bytes_out);
BZIP3_API int bz3_b200_decode_fd(int in_fd, int out_fd, int in_flight, uint64_t *bytes_in, uint64_t *bytes_out);
/* The same over several GPUs of one process: the workers are dealt round-robin over `devices` GPUs starting with the
 * current one (<= 0: all visible), i.e. block i goes to GPU (i mod in_flight) mod devices -- the static work queue of
 * independent blocks; nothing is exchanged between GPUs.  in_flight <= 0 then means the automatic depth PER call, so
 * pass e.g. 64 x devices.  The plain entry points above are devices = 1 (one process per GPU, as bench.py runs). */
BZIP3_API int bz3_b200_encode_fd2(int in_fd, int out_fd, int32_t block_size, int in_flight, int devices, uint64_t *bytes_in, uint64_t *bytes_out);
BZIP3_API int bz3_b200_decode_fd2(int in_fd, int out_fd, int in_flight, int devices, uint64_t *bytes_in, uint64_t *bytes_out);
#ifdef __cplusplus
}
#endif
#endif
