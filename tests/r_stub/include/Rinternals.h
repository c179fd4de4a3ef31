/* Minimal stand-in for R's C API, for COMPILING and DRIVING r/src/mcb_shim.c where R is not installed (test infrastructure
 * only; see tests/test_r_shim.py). It implements just the entry points the shim uses, with R's semantics where they matter
 * to the shim: vectors own their storage, matrices are column-major, PROTECT is a counter, Rf_error long-jumps to the
 * harness (R long-jumps to its top level). Not a re-implementation of R and not used by the product. */
#ifndef MCB_STUB_RINTERNALS_H
#define MCB_STUB_RINTERNALS_H
#include <setjmp.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22, CHARSXP = 9 } SEXPTYPE;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
typedef ptrdiff_t R_xlen_t;

typedef struct SEXPREC {
  SEXPTYPE type;
  R_xlen_t len;
  void* data;            /* double / int / SEXP* / char* payload, or the external pointer address */
  struct SEXPREC* attr_logdet;
  void (*finalizer)(struct SEXPREC*);
} SEXPREC, *SEXP;

static SEXPREC mcb_stub_nil = {NILSXP, 0, NULL, NULL, NULL};
#define R_NilValue (&mcb_stub_nil)

extern jmp_buf mcb_stub_error_jmp;      /* defined by the harness */
extern char mcb_stub_error_msg[1024];
extern int mcb_stub_protect_depth;

static inline SEXP mcb_stub_new(SEXPTYPE t, R_xlen_t n, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof(SEXPREC));
  s->type = t;
  s->len = n;
  s->data = n > 0 ? calloc((size_t)n, elt) : NULL;
  return s;
}
static inline SEXP Rf_allocVector(SEXPTYPE t, R_xlen_t n) {
  return mcb_stub_new(t, n, t == REALSXP ? sizeof(double) : (t == VECSXP || t == STRSXP) ? sizeof(SEXP) : sizeof(int));
}
static inline SEXP Rf_allocMatrix(SEXPTYPE t, int nr, int nc) { return Rf_allocVector(t, (R_xlen_t)nr * nc); }
static inline SEXP Rf_alloc3DArray(SEXPTYPE t, int a, int b, int c) { return Rf_allocVector(t, (R_xlen_t)a * b * c); }
static inline SEXP PROTECT(SEXP s) { ++mcb_stub_protect_depth; return s; }
static inline void UNPROTECT(int n) { mcb_stub_protect_depth -= n; }
static inline double* REAL(SEXP s) { return (double*)s->data; }
static inline int* INTEGER(SEXP s) { return (int*)s->data; }
static inline int* LOGICAL(SEXP s) { return (int*)s->data; }
static inline int Rf_length(SEXP s) { return (int)s->len; }
static inline R_xlen_t Rf_xlength(SEXP s) { return s->len; }
static inline int Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP; }
static inline int Rf_asInteger(SEXP s) { return s->type == REALSXP ? (int)REAL(s)[0] : INTEGER(s)[0]; }
static inline int Rf_asLogical(SEXP s) { return s->type == REALSXP ? REAL(s)[0] != 0.0 : INTEGER(s)[0] != 0; }
static inline double Rf_asReal(SEXP s) { return s->type == REALSXP ? REAL(s)[0] : (double)INTEGER(s)[0]; }
static inline SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
static inline SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
static inline SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = v; return s; }
static inline void SET_VECTOR_ELT(SEXP v, R_xlen_t i, SEXP x) { ((SEXP*)v->data)[i] = x; }
static inline SEXP VECTOR_ELT(SEXP v, R_xlen_t i) { return ((SEXP*)v->data)[i]; }
static inline SEXP Rf_mkChar(const char* c) {
  SEXP s = mcb_stub_new(CHARSXP, (R_xlen_t)strlen(c) + 1, 1);
  memcpy(s->data, c, strlen(c) + 1);
  return s;
}
static inline const char* CHAR(SEXP s) { return (const char*)s->data; }
static inline SEXP STRING_ELT(SEXP v, R_xlen_t i) { return ((SEXP*)v->data)[i]; }
static inline void SET_STRING_ELT(SEXP v, R_xlen_t i, SEXP x) { ((SEXP*)v->data)[i] = x; }
static inline SEXP Rf_mkString(const char* c) { SEXP s = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(s, 0, Rf_mkChar(c)); return s; }
static inline int TYPEOF(SEXP s) { return (int)s->type; }
static inline SEXP Rf_install(const char* name) { (void)name; return R_NilValue; }   /* only "logdet" is ever installed */
static inline SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val) { (void)name; x->attr_logdet = val; return val; }
static inline SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
  (void)tag; (void)prot;
  SEXP s = mcb_stub_new(EXTPTRSXP, 0, 1);
  s->data = p;
  return s;
}
static inline void* R_ExternalPtrAddr(SEXP s) { return s->data; }
static inline void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
static inline void R_RegisterCFinalizerEx(SEXP s, void (*fin)(SEXP), Rboolean onexit) { (void)onexit; s->finalizer = fin; }
static inline void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(mcb_stub_error_msg, sizeof mcb_stub_error_msg, fmt, ap);
  va_end(ap);
  longjmp(mcb_stub_error_jmp, 1);
}
#endif
