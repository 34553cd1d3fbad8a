/* TEST INFRASTRUCTURE: the THC names lua/init.cu uses (cutorch/lib/THC): THCState with a current stream, and
 * THCudaTensor accessors that take the state first. */
#ifndef HARNESS_THC_H
#define HARNESS_THC_H
#include "TH.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct THCState {
  void *current_stream; /* cudaStream_t */
} THCState;
struct CUstream_st;
static inline struct CUstream_st *THCState_getCurrentStream(THCState *s) { return (struct CUstream_st *)s->current_stream; }

HARNESS_DECLARE_TENSOR(THCudaTensorImpl, float)
typedef THCudaTensorImpl THCudaTensor;
static inline float *THCudaTensor_data(THCState *s, const THCudaTensor *t) { (void)s; return THCudaTensorImpl_data(t); }
static inline int THCudaTensor_nDimension(THCState *s, const THCudaTensor *t) { (void)s; return THCudaTensorImpl_nDimension(t); }
static inline long THCudaTensor_size(THCState *s, const THCudaTensor *t, int d) { (void)s; return THCudaTensorImpl_size(t, d); }
static inline long THCudaTensor_stride(THCState *s, const THCudaTensor *t, int d) { (void)s; return THCudaTensorImpl_stride(t, d); }
static inline ptrdiff_t THCudaTensor_nElement(THCState *s, const THCudaTensor *t) { (void)s; return THCudaTensorImpl_nElement(t); }
static inline int THCudaTensor_isContiguous(THCState *s, const THCudaTensor *t) { (void)s; return THCudaTensorImpl_isContiguous(t); }
/* the harness cannot allocate device memory: the test pre-allocates `capacity` elements and resize1d only re-shapes */
void THCudaTensor_resize1d(THCState *s, THCudaTensor *t, long size0);
#ifdef __cplusplus
}
#endif
#endif
