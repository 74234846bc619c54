/*
 * ref_stage_shim.c -- exposes the reference's file-static stage functions for differential tests.
 * TEST INFRASTRUCTURE ONLY.  It contains no reference code: it #includes the reference translation
 * unit from where it lies (REF_TU, set by oracle/Makefile to /root/reference/src/libbz3.c) and
 * wraps the statics in exported symbols.  Built only in the container that has /root/reference.
 */
#include REF_TU

#define SHIM __attribute__((visibility("default")))

SHIM uint32_t ref_crc32(uint32_t crc, uint8_t *buf, size_t n) { return crc32sum(crc, buf, n); }
SHIM int32_t ref_mrlec(uint8_t *in, int32_t n, uint8_t *out) { return mrlec(in, n, out); }
SHIM int ref_mrled(uint8_t *in, uint8_t *out, int32_t outlen, int32_t maxin) { return mrled(in, out, outlen, maxin); }
SHIM int32_t ref_lzp_compress(const uint8_t *in, uint8_t *out, int32_t n, int32_t *lut) {
    return lzp_compress(in, out, n, lut);
}
SHIM int32_t ref_lzp_decompress(const uint8_t *in, uint8_t *out, int32_t n, int32_t max, int32_t *lut) {
    return lzp_decompress(in, out, n, max, lut);
}
SHIM int32_t ref_bwt(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n) { return libsais_bwt(T, U, A, n, 0, NULL); }
SHIM int32_t ref_unbwt(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, int32_t idx) {
    return libsais_unbwt(T, U, A, n, NULL, idx);
}
SHIM int32_t ref_cm_encode(uint8_t *in, int32_t n, uint8_t *out) {
    state *s = malloc(sizeof(state));
    begin(s);
    s->out_queue = out;
    s->output_ptr = 0;
    encode_bytes(s, in, n);
    int32_t r = s->output_ptr;
    free(s);
    return r;
}
SHIM void ref_cm_decode(uint8_t *in, int32_t insize, uint8_t *out, int32_t n) {
    state *s = malloc(sizeof(state));
    begin(s);
    s->in_queue = in;
    s->input_ptr = 0;
    s->input_max = insize;
    decode_bytes(s, out, n);
    free(s);
}
