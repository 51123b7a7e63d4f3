/* TEST INFRASTRUCTURE ONLY -- stand-ins for the few OCaml runtime functions the alignment stubs use, and helpers
 * that build the custom blocks (`struct seq`, `struct cm`, `struct cm_3d`, `struct matrices`) those stubs take, so
 * that tests/test_stubs.py can EXECUTE the reference's own algn_CAML_* symbols (oracle/_ref/libpoyref.so) and the
 * drop-in ones of stubs/poyb200_stubs.c (oracle/_ref/libpoystubs.so) side by side on the same blocks.
 *
 * Layout conventions are those of the shim headers (oracle/shim/caml/caml_shim.h): a custom block is
 * [ops pointer][payload]; tuples / strings carry a header word in front (wosize << 10 | tag).
 * Nothing here is linked into the product library. */
#include <assert.h>
#include <setjmp.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <caml/mlvalues.h>
#include "matrices.h"
#include "seq.h"
#include "cm.h"

/* ---- failwith: OCaml's Failure exception.  Armed by camlrt_call: the message is kept and control returns there. ---- */
static jmp_buf g_jmp;
static int g_armed = 0;
static char g_msg[512];

void failwith(const char *msg) {
    if (g_armed) {
        snprintf(g_msg, sizeof g_msg, "%s", msg);
        g_armed = 0;
        longjmp(g_jmp, 1);
    }
    fprintf(stderr, "failwith outside camlrt_call: %s\n", msg);
    abort();
}
void caml_invalid_argument(const char *msg) { failwith(msg); }
const char *camlrt_last_failure(void) { return g_msg; }

/* Calls fn(args[0..nargs-1]) like the OCaml runtime would; *failed = 1 (and the Failure text kept) if it raised. */
value camlrt_call(void *fn, int nargs, const value *a, int *failed) {
    volatile value res = Val_unit;
    *failed = 0;
    g_msg[0] = 0;
    if (setjmp(g_jmp)) {
        *failed = 1;
        return Val_unit;
    }
    g_armed = 1;
    switch (nargs) {
        case 3: res = ((value (*)(value, value, value)) fn)(a[0], a[1], a[2]); break;
        case 4: res = ((value (*)(value, value, value, value)) fn)(a[0], a[1], a[2], a[3]); break;
        case 5: res = ((value (*)(value, value, value, value, value)) fn)(a[0], a[1], a[2], a[3], a[4]); break;
        case 6: res = ((value (*)(value, value, value, value, value, value)) fn)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case 7: res = ((value (*)(value, value, value, value, value, value, value)) fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
        case 8: res = ((value (*)(value, value, value, value, value, value, value, value)) fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
        case 9: res = ((value (*)(value, value, value, value, value, value, value, value, value)) fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
        default: g_armed = 0; *failed = 2; return Val_unit;
    }
    g_armed = 0;
    return res;
}

/* ---- allocation ------------------------------------------------------------------------------------------------------ */
static value alloc_words(mlsize_t wosize, int tag) {
    uintnat *p = (uintnat *) calloc(wosize + 1, sizeof(uintnat));
    p[0] = ((uintnat) wosize << 10) | (uintnat) tag;
    return (value) (p + 1);
}
value caml_alloc_tuple(mlsize_t n) {
    value v = alloc_words(n ? n : 1, 0);
    ((uintnat *) v)[-1] = ((uintnat) n << 10);
    for (mlsize_t i = 0; i < n; i++) Field(v, i) = Val_unit;
    return v;
}
value caml_alloc(mlsize_t n, int tag) { return alloc_words(n, tag); }
value caml_alloc_string(mlsize_t len) {
    const mlsize_t wosize = (len + sizeof(value)) / sizeof(value);
    value v = alloc_words(wosize, 252);
    ((unsigned char *) v)[wosize * sizeof(value) - 1] = (unsigned char) (wosize * sizeof(value) - 1 - len);
    return v;
}
value caml_copy_string(const char *s) {
    value v = caml_alloc_string(strlen(s));
    memcpy((char *) v, s, strlen(s));
    return v;
}
mlsize_t camlrt_string_length(value v) {
    const mlsize_t bytes = Wosize_val(v) * sizeof(value);
    return bytes - 1 - ((unsigned char *) v)[bytes - 1];
}
value caml_alloc_custom(struct custom_operations *ops, uintnat size, mlsize_t mem, mlsize_t max) {
    (void) mem; (void) max;
    value *p = (value *) calloc(1, sizeof(value) + size);
    p[0] = (value) ops;
    return (value) p;
}
void caml_register_custom_operations(struct custom_operations *ops) { (void) ops; }
value caml_copy_double(double d) { double *p = (double *) malloc(sizeof(double)); *p = d; return (value) p; }
void caml_modify(value *fp, value v) { *fp = v; }
void caml_serialize_int_4(int32_t i) { (void) i; }
void caml_serialize_block_1(void *d, intnat l) { (void) d; (void) l; }
void caml_serialize_block_4(void *d, intnat l) { (void) d; (void) l; }
int32_t caml_deserialize_sint_4(void) { return 0; }
uint32_t caml_deserialize_uint_4(void) { return 0; }
void caml_deserialize_block_1(void *d, intnat l) { (void) d; (void) l; }
void caml_deserialize_block_4(void *d, intnat l) { (void) d; (void) l; }

/* ---- custom blocks ------------------------------------------------------------------------------------------------------ */

/* A sequence block as seq_CAML_create makes it (src/seq.c:362-393): struct seq + cap bytes of storage (+ one zeroed
 * guard byte, see ref_driver.c mk_seq), holding `len` elements right aligned. */
value camlrt_seq(const unsigned char *data, int len, int cap) {
    if (cap < len) cap = len;
    value v = caml_alloc_custom(NULL, sizeof(struct seq) + (size_t) cap + 16, 0, 0);
    seqt s = Seq_pointer(v);
    s->magic_number = POY_SEQ_MAGIC_NUMBER;
    s->cap = cap;
    s->len = len;
    s->head = (SEQT *) (s + 1);
    s->end = s->head + cap - 1;
    s->begin = s->end - len + 1;
    if (len) memcpy(s->begin, data, (size_t) len);
    return v;
}
/* Deliberately stale pointers, as after a GC move: every stub must re-derive them (Seq_custom_val). */
void camlrt_seq_scramble(value v) {
    seqt s = Seq_pointer(v);
    s->head = s->begin = s->end = (SEQT *) 16;
}
int camlrt_seq_len(value v) { return Seq_pointer(v)->len; }
int camlrt_seq_read(value v, unsigned char *out) {
    seqt s;
    Seq_custom_val(s, v);
    memcpy(out, s->begin, (size_t) s->len);
    return s->len;
}
void camlrt_seq_clear(value v) { Seq_pointer(v)->len = 0; }

/* A cost-matrix block whose payload is a COPY of the struct (the tables stay shared with the handle it was made from). */
value camlrt_cm(const struct cm *c) {
    value v = caml_alloc_custom(NULL, sizeof(struct cm), 0, 0);
    memcpy(Data_custom_val(v), c, sizeof(struct cm));
    return v;
}
value camlrt_cm3(const struct cm_3d *c) {
    value v = caml_alloc_custom(NULL, sizeof(struct cm_3d), 0, 0);
    memcpy(Data_custom_val(v), c, sizeof(struct cm_3d));
    return v;
}
/* A fresh Matrix.m (mat_CAML_create_general, src/matrices.c:139-153: all zero). */
value camlrt_matrices(void) { return caml_alloc_custom(NULL, sizeof(struct matrices), 0, 0); }

/* arrays for the batched externals */
value camlrt_array(int n) { return caml_alloc_tuple((mlsize_t) n); }
void camlrt_array_set(value arr, int i, value v) { Field(arr, i) = v; }
value camlrt_array_get(value arr, int i) { return Field(arr, i); }
int camlrt_array_len(value arr) { return (int) Wosize_val(arr); }
value camlrt_val_int(int x) { return Val_int(x); }
int camlrt_int_val(value v) { return Int_val(v); }
void camlrt_bytes_read(value str, unsigned char *out, int n) { memcpy(out, (const void *) str, (size_t) n); }

/* ---- Powell's 3-D aligner in one C call (tests only): powell_3D_align (src/ukkCommon.c:110-145) on fresh blocks.
 * Sequences with the leading gap; r1..r3 need l1 + l2 + l3 bytes.  Returns the cost, -1 if the stub raised. */
extern value powell_3D_align(value, value, value, value, value, value, value, value, value);
int camlrt_powell(const unsigned char *s1, int l1, const unsigned char *s2, int l2, const unsigned char *s3, int l3, int mm, int go,
                  int ge, unsigned char *r1, unsigned char *r2, unsigned char *r3, int *rlen) {
    const int cap = l1 + l2 + l3;
    value a[9];
    int failed = 0;
    a[0] = camlrt_seq(s1, l1, l1); a[1] = camlrt_seq(s2, l2, l2); a[2] = camlrt_seq(s3, l3, l3);
    a[3] = camlrt_seq(NULL, 0, cap); a[4] = camlrt_seq(NULL, 0, cap); a[5] = camlrt_seq(NULL, 0, cap);
    a[6] = Val_int(mm); a[7] = Val_int(go); a[8] = Val_int(ge);
    value res = camlrt_call((void *) powell_3D_align, 9, a, &failed);
    if (failed) return -1;
    *rlen = camlrt_seq_read(a[3], r1);
    camlrt_seq_read(a[4], r2);
    camlrt_seq_read(a[5], r3);
    return Int_val(res);
}
