/* TEST INFRASTRUCTURE ONLY -- stand-in for the OCaml runtime headers.
 *
 * The reference's native alignment code (/root/reference/src/algn.c and the
 * files it textually includes) only needs the OCaml headers for its
 * `algn_CAML_*` stubs.  No OCaml toolchain exists in this image, so this file
 * declares just enough of <caml/...> for that translation unit to compile
 * unmodified.  The oracle driver calls the plain-C functions underneath the stubs; tests/test_stubs.py
 * also executes the algn_CAML_* stubs themselves (the reference's and stubs/poyb200_stubs.c's) on blocks
 * built by oracle/caml_runtime.c.
 */
#ifndef POYB200_CAML_SHIM_H
#define POYB200_CAML_SHIM_H
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>

typedef intptr_t value;
typedef uintptr_t uintnat;
typedef intptr_t intnat;
typedef int32_t int32;
typedef uint32_t uint32;
typedef int64_t int64;
typedef uintptr_t mlsize_t;

#define Val_long(x) ((value) (((intnat) (x) << 1) + 1))
#define Long_val(x) ((x) >> 1)
#define Val_int(x) Val_long(x)
#define Int_val(x) ((int) Long_val(x))
#define Val_unit Val_int(0)
#define Val_bool(x) Val_int((x) != 0)
#define Bool_val(x) Int_val(x)
#define Val_true Val_int(1)
#define Val_false Val_int(0)
#define Is_block(x) (((x) & 1) == 0)
#define Is_long(x) (((x) & 1) != 0)
#define Field(x, i) (((value *) (x))[i])
#define Store_field(b, i, v) (Field(b, i) = (v))
/* blocks made by the stand-in allocators (oracle/caml_runtime.c) carry an OCaml-style header word: wosize << 10 | tag */
#define Hd_val(v) (((uintnat *) (v))[-1])
#define Wosize_val(v) ((mlsize_t) (Hd_val(v) >> 10))
#define Double_val(v) (*((double *) (v)))
#define String_val(v) ((char *) (v))
#define Data_custom_val(v) ((void *) &Field((v), 1))
#define Custom_ops_val(v) (*((struct custom_operations **) (v)))

struct custom_operations {
    char *identifier;
    void (*finalize)(value v);
    int (*compare)(value v1, value v2);
    intnat (*hash)(value v);
    void (*serialize)(value v, uintnat *wsize_32, uintnat *wsize_64);
    uintnat (*deserialize)(void *dst);
    int (*compare_ext)(value v1, value v2);
};
#define custom_finalize_default NULL
#define custom_compare_default NULL
#define custom_hash_default NULL
#define custom_serialize_default NULL
#define custom_deserialize_default NULL
#define custom_compare_ext_default NULL

#define CAMLparam0()
#define CAMLparam1(a)
#define CAMLparam2(a, b)
#define CAMLparam3(a, b, c)
#define CAMLparam4(a, b, c, d)
#define CAMLparam5(a, b, c, d, e)
#define CAMLxparam1(a)
#define CAMLxparam2(a, b)
#define CAMLxparam3(a, b, c)
#define CAMLxparam4(a, b, c, d)
#define CAMLxparam5(a, b, c, d, e)
#define CAMLlocal1(a) value a = Val_unit
#define CAMLlocal2(a, b) value a = Val_unit, b = Val_unit
#define CAMLlocal3(a, b, c) value a = Val_unit, b = Val_unit, c = Val_unit
#define CAMLlocal4(a, b, c, d) value a = Val_unit, b = Val_unit, c = Val_unit, d = Val_unit
#define CAMLlocal5(a, b, c, d, e) value a = Val_unit, b = Val_unit, c = Val_unit, d = Val_unit, e = Val_unit
#define CAMLreturn(x) return (x)
#define CAMLreturn0 return

value caml_alloc_custom(struct custom_operations *ops, uintnat size, mlsize_t mem, mlsize_t max);
#define alloc_custom caml_alloc_custom
void caml_register_custom_operations(struct custom_operations *ops);
#define register_custom_operations caml_register_custom_operations
void failwith(const char *msg) __attribute__((noreturn));
#define caml_failwith failwith
void caml_invalid_argument(const char *msg) __attribute__((noreturn));
#define invalid_argument caml_invalid_argument
value caml_copy_double(double d);
#define copy_double caml_copy_double
value caml_alloc_tuple(mlsize_t n);
value caml_alloc_string(mlsize_t len);
#define Bytes_val(v) ((unsigned char *) (v))
#define alloc_tuple caml_alloc_tuple
value caml_alloc(mlsize_t n, int tag);
value caml_copy_string(const char *s);
void caml_modify(value *fp, value v);

void caml_serialize_int_1(int i);
void caml_serialize_int_2(int i);
void caml_serialize_int_4(int32_t i);
void caml_serialize_int_8(int64_t i);
void caml_serialize_block_1(void *data, intnat len);
void caml_serialize_block_4(void *data, intnat len);
void caml_serialize_block_8(void *data, intnat len);
int caml_deserialize_uint_1(void);
int caml_deserialize_sint_1(void);
int caml_deserialize_uint_2(void);
int caml_deserialize_sint_2(void);
uint32_t caml_deserialize_uint_4(void);
int32_t caml_deserialize_sint_4(void);
void caml_deserialize_block_1(void *data, intnat len);
void caml_deserialize_block_4(void *data, intnat len);
void caml_deserialize_block_8(void *data, intnat len);
#define serialize_int_1 caml_serialize_int_1
#define serialize_int_2 caml_serialize_int_2
#define serialize_int_4 caml_serialize_int_4
#define serialize_int_8 caml_serialize_int_8
#define serialize_block_1 caml_serialize_block_1
#define serialize_block_4 caml_serialize_block_4
#define serialize_block_8 caml_serialize_block_8
#define deserialize_uint_1 caml_deserialize_uint_1
#define deserialize_sint_1 caml_deserialize_sint_1
#define deserialize_uint_2 caml_deserialize_uint_2
#define deserialize_sint_2 caml_deserialize_sint_2
#define deserialize_uint_4 caml_deserialize_uint_4
#define deserialize_sint_4 caml_deserialize_sint_4
#define deserialize_block_1 caml_deserialize_block_1
#define deserialize_block_4 caml_deserialize_block_4
#define deserialize_block_8 caml_deserialize_block_8

/* bigarray: only the type names are needed for prototypes */
struct caml_ba_array { void *data; intnat num_dims; intnat flags; void *proxy; intnat dim[1]; };
#define Caml_ba_array_val(v) ((struct caml_ba_array *) Data_custom_val(v))
#define Bigarray_val(v) Caml_ba_array_val(v)
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
#define Data_bigarray_val(v) Caml_ba_data_val(v)
#define caml_bigarray caml_ba_array
#endif
