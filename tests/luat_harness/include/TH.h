/* TEST INFRASTRUCTURE: the TH names the glue uses (torch7/lib/TH): tensor structs with size / stride / data, their
 * accessor functions, the generic-file machinery (TH_CONCAT_*, THTensor, THTensor_(NAME)) and THError. */
#ifndef HARNESS_TH_H
#define HARNESS_TH_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TH_CONCAT_STRING_3(x, y, z) TH_CONCAT_STRING_3_EXPAND(x, y, z)
#define TH_CONCAT_STRING_3_EXPAND(x, y, z) #x #y #z
#define TH_CONCAT_3(x, y, z) TH_CONCAT_3_EXPAND(x, y, z)
#define TH_CONCAT_3_EXPAND(x, y, z) x##y##z
#define TH_CONCAT_4(x, y, z, w) TH_CONCAT_4_EXPAND(x, y, z, w)
#define TH_CONCAT_4_EXPAND(x, y, z, w) x##y##z##w

#define THTensor TH_CONCAT_3(TH, Real, Tensor)
#define THTensor_(NAME) TH_CONCAT_4(TH, Real, Tensor_, NAME)

void THError(const char *fmt, ...);

#define HARNESS_DECLARE_TENSOR(T, real)                                           \
  typedef struct T {                                                              \
    long *size;                                                                   \
    long *stride;                                                                 \
    int nDimension;                                                               \
    real *data_;                                                                  \
    long capacity; /* elements; harness only (resize1d of the CUDA stand-in) */   \
  } T;                                                                            \
  static inline real *T##_data(const T *t) { return t->data_; }                   \
  static inline int T##_nDimension(const T *t) { return t->nDimension; }          \
  static inline long T##_size(const T *t, int d) { return t->size[d]; }           \
  static inline long T##_stride(const T *t, int d) { return t->stride[d]; }       \
  static inline ptrdiff_t T##_nElement(const T *t) {                              \
    ptrdiff_t n = t->nDimension ? 1 : 0;                                          \
    int d;                                                                        \
    for (d = 0; d < t->nDimension; d++) n *= t->size[d];                          \
    return n;                                                                     \
  }                                                                               \
  static inline int T##_isContiguous(const T *t) {                                \
    long z = 1;                                                                   \
    int d;                                                                        \
    for (d = t->nDimension - 1; d >= 0; d--) {                                    \
      if (t->size[d] != 1) {                                                      \
        if (t->stride[d] != z) return 0;                                          \
        z *= t->size[d];                                                          \
      }                                                                           \
    }                                                                             \
    return 1;                                                                     \
  }

HARNESS_DECLARE_TENSOR(THFloatTensor, float)
HARNESS_DECLARE_TENSOR(THDoubleTensor, double)

#ifdef __cplusplus
}
#endif
#endif
