#include <stddef.h>
#include <stdint.h>
void *R_alloc(size_t, int); double unif_rand(void); void GetRNGstate(void); void PutRNGstate(void);
void Rf_error(const char*, ...);
extern double R_NaReal;
#define NA_REAL R_NaReal
