/* Minimal declarations of the R C API used by r_shim/src/bssm_shim.c: R itself is absent from the image, so the
 * shim is compile-checked (gcc -fsyntax-only) against these stubs in tests/test_abi.py.  Not part of the product. */
typedef struct SEXPREC *SEXP; typedef ptrdiff_t R_xlen_t; typedef unsigned char Rbyte;
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
#define RAWSXP 24
#define STRSXP 16
int TYPEOF(SEXP); R_xlen_t XLENGTH(SEXP); double *REAL(SEXP); int *INTEGER(SEXP); Rbyte *RAW(SEXP);
SEXP Rf_allocVector(int, R_xlen_t); SEXP Rf_allocMatrix(int,int,int); SEXP PROTECT(SEXP); void UNPROTECT(int);
int Rf_asInteger(SEXP); double Rf_asReal(SEXP); SEXP Rf_getAttrib(SEXP,SEXP); SEXP Rf_setAttrib(SEXP,SEXP,SEXP);
SEXP VECTOR_ELT(SEXP,R_xlen_t); SEXP SET_VECTOR_ELT(SEXP,R_xlen_t,SEXP); SEXP STRING_ELT(SEXP,R_xlen_t); const char* CHAR(SEXP);
int Rf_isMatrix(SEXP); int Rf_nrows(SEXP); int Rf_ncols(SEXP); SEXP Rf_mkNamed(int,const char**); SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarLogical(int); SEXP Rf_ScalarInteger(int); SEXP Rf_mkString(const char*);
