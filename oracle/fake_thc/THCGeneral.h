/* oracle/fake_thc/THCGeneral.h -- TEST INFRASTRUCTURE ONLY: the slice of cutorch's THC that the reference's
 * CUDA file uses (tensor struct with size[] / stride[] / data, the current stream, THError). */
#ifndef VOLSTN_FAKE_THC_H
#define VOLSTN_FAKE_THC_H
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
typedef struct THCState { cudaStream_t stream; } THCState;
typedef struct THCudaTensor { long size[5]; long stride[5]; float *data; } THCudaTensor;
static inline float *THCudaTensor_data(THCState *s, THCudaTensor *t) { (void)s; return t->data; }
static inline long THCudaTensor_stride(THCState *s, THCudaTensor *t, int d) { (void)s; return t->stride[d]; }
static inline long THCudaTensor_size(THCState *s, THCudaTensor *t, int d) { (void)s; return t->size[d]; }
static inline cudaStream_t THCState_getCurrentStream(THCState *s) { return s->stream; }
extern int volstn_ref_cuda_error;
static inline void THError(const char *msg) { fprintf(stderr, "THError: %s\n", msg); volstn_ref_cuda_error = 1; }
#endif
