/*
 * volstn_glue.h -- shared by lua/init.c (libvolstn: torch.FloatTensor / torch.DoubleTensor) and lua/init.cu
 * (libcuvolstn: torch.CudaTensor).  Nothing but argument marshalling: Torch7 tensors on the Lua stack ->
 * volstn_tensor5 descriptors -> one call into the C ABI of include/volstn_b200.h (libvolstn_b200.so).
 *
 * Replaces the compiled body of stn3dtorch/init.c + generic/BilinearSamplerVolumetric.c (CPU) and of
 * stn3dtorch/init.cu + BilinearSamplerVolumetric.cu + utils.c (CUDA); file:line citations are relative to the
 * reference root.  Valid C and valid C++ (init.cu is compiled by nvcc).
 */
#ifndef VOLSTN_GLUE_H
#define VOLSTN_GLUE_H

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include "volstn_b200.h"

/* Flag bits a Lua caller may pass in the optional trailing number of the sampler functions -- the slot the
 * reference reads as `focal_length` and never uses (generic/BilinearSamplerVolumetric.c:447,592); nil reads as 0 =
 * exactly the reference's contract (gradInputImages accumulated into, inputs re-read by updateGradInput). */
#define VOLSTN_GLUE_FLAG_MASK                                                                                     \
  (VOLSTN_BWD_ONLY_GRID | VOLSTN_BWD_ONLY_VOLUME | VOLSTN_BWD_ZERO_GRADVOL | VOLSTN_HOST_KEEP_INPUTS | \
   VOLSTN_HOST_REUSE_INPUTS | VOLSTN_ALGO_MASK | VOLSTN_FAST_ARITH)

/* The reference's error convention (BilinearSamplerVolumetric.cu:736-740, 1066-1070): print, then THError. */
#define VOLSTN_GLUE_CHECK(where, rc)                                                                  \
  do {                                                                                                \
    const int volstn_rc_ = (rc);                                                                      \
    if (volstn_rc_ != VOLSTN_OK) {                                                                    \
      printf("error in %s: %s%s%s\n", (where), volstn_b200_strerror(volstn_rc_),                      \
             volstn_rc_ == VOLSTN_ERR_CUDA ? " -- " : "",                                             \
             volstn_rc_ == VOLSTN_ERR_CUDA ? volstn_b200_last_cuda_error_string() : "");              \
      THError("aborting");                                                                            \
    }                                                                                                 \
  } while (0)

/* ---------------------------------------------------------------------------------------------------------
 * One staging context per OS thread for host tensors: the reference's loader threads call the sampler
 * concurrently (data.lua:11-24, dataset_2d.lua:293-306), and a volstn_host_ctx serves one thread at a time.
 * The device is $VOLSTN_DEVICE (default 0): libvolstn is the CPU-tensor library and has no device argument.
 * ------------------------------------------------------------------------------------------------------- */
static pthread_key_t volstn_glue_key;
static pthread_once_t volstn_glue_once = PTHREAD_ONCE_INIT;

static void volstn_glue_ctx_free(void *p) { volstn_b200_host_ctx_destroy((volstn_host_ctx *)p); }
static void volstn_glue_key_make(void) { pthread_key_create(&volstn_glue_key, volstn_glue_ctx_free); }

#if defined(__GNUC__)
__attribute__((unused))
#endif
static volstn_host_ctx *volstn_glue_ctx(void) {
  volstn_host_ctx *ctx;
  pthread_once(&volstn_glue_once, volstn_glue_key_make);
  ctx = (volstn_host_ctx *)pthread_getspecific(volstn_glue_key);
  if (!ctx) {
    const char *dev = getenv("VOLSTN_DEVICE");
    VOLSTN_GLUE_CHECK("volstn host context", volstn_b200_host_ctx_create(dev ? atoi(dev) : 0, &ctx));
    pthread_setspecific(volstn_glue_key, ctx);
  }
  return ctx;
}

static unsigned volstn_glue_flags(double number) {
  return number > 0 ? ((unsigned)number & VOLSTN_GLUE_FLAG_MASK) : 0u;
}

#endif /* VOLSTN_GLUE_H */
