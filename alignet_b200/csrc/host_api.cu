// host_api.cu -- HOST-buffer entry points (what `require 'libvolstn'` binds for torch.FloatTensor /
// torch.DoubleTensor in the reference, stn3dtorch/init.c:13-21).
//
// There is no CPU compute path: host tensors are cut into batch slices (samples are independent:
// batch is the outermost loop of every reference function, ...c:482,641) and each slice is
// H2D-copied, processed by the sm_100a kernels and D2H-copied on one of kSlots streams, so that the
// upload of slice i+1, the kernels of slice i and the download of slice i-1 overlap (the B200 has
// independent copy engines per direction).  The call returns after the last download.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "volstn_common.cuh"
#include "volstn_internal.h"

struct volstn_host_ctx {
  static constexpr int kSlots = 3;
  int device;
  cudaStream_t stream[kSlots];
  void *arena[kSlots];
  size_t cap[kSlots];
  // Device copies of the last forward call's inputs (VOLSTN_HOST_KEEP_INPUTS), reused by the next backward
  // call on the same host tensors (VOLSTN_HOST_REUSE_INPUTS).  nn.Module's contract makes this legal for
  // the Lua / Python modules: updateGradInput(input, gradOutput) assumes updateOutput(input) ran before on
  // the same input.  Keyed by host pointer, shape and dtype; invalid after any other host sampler call.
  void *keep_vol, *keep_grid;
  size_t keep_vol_cap, keep_grid_cap;
  bool keep_valid;
  int keep_dtype;
  volstn_tensor5 keep_vol_key, keep_grid_key;
};

namespace volstn {
namespace {

constexpr size_t kAlign = 256;
// per slice, all tensors together (VOLSTN_HOST_SLICE_MB overrides, for tuning)
static size_t slice_target_bytes() {
  static const size_t v = [] {
    const char *e = getenv("VOLSTN_HOST_SLICE_MB");
    const long long mb = e ? atoll(e) : 48;
    return (size_t)(mb > 0 ? mb : 48) << 20;
  }();
  return v;
}

size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

static bool ramp_slices() {  // VOLSTN_HOST_RAMP=0 switches the slice ramps off (A/B timing)
  static const bool v = [] {
    const char *e = getenv("VOLSTN_HOST_RAMP");
    return !(e && e[0] == '0');
  }();
  return v;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int ensure_arena(volstn_host_ctx *c, int slot, size_t bytes) {
  if (c->cap[slot] >= bytes) return VOLSTN_OK;
  VOLSTN_CUDA_TRY(cudaStreamSynchronize(c->stream[slot]));
  if (c->arena[slot]) VOLSTN_CUDA_TRY(cudaFree(c->arena[slot]));
  c->arena[slot] = nullptr;
  c->cap[slot] = 0;
  VOLSTN_CUDA_TRY(cudaMalloc(&c->arena[slot], bytes));
  c->cap[slot] = bytes;
  return VOLSTN_OK;
}

int sync_all(volstn_host_ctx *c) {
  int rc = VOLSTN_OK;
  for (int s = 0; s < volstn_host_ctx::kSlots; s++) {
    cudaError_t e = cudaStreamSynchronize(c->stream[s]);
    if (e != cudaSuccess && rc == VOLSTN_OK) rc = cuda_fail(e);
  }
  return rc;
}

// samples per slice: ~kSliceTargetBytes, and enough slices to keep all slots busy
int64_t slice_batch(int64_t B, size_t bytes_per_sample) {
  int64_t bs = (int64_t)(slice_target_bytes() / (bytes_per_sample ? bytes_per_sample : 1));
  const int64_t per_slot = (B + volstn_host_ctx::kSlots - 1) / volstn_host_ctx::kSlots;
  bs = std::min(bs, per_slot);
  return std::max<int64_t>(bs, 1);
}

// Slice schedule of the sampler calls: sizes ramp up 1, 2, 4, ... to `bs` and down again over the last 2 * bs samples.
// The forward is bound by its uploads and only its LAST slice's kernel + download stick out (a one-sample slice makes
// that 20 us instead of 0.15 ms); the backward is bound by its downloads, which can start as soon as the FIRST slice is
// through (one sample: 0.1 ms instead of 0.5 ms).
int64_t slice_len(int64_t b0, int64_t B, int64_t bs, int slice) {
  int64_t nb = bs;
  const int64_t up = (int64_t)1 << (slice < 30 ? slice : 30);
  if (up < nb) nb = up;
  const int64_t rem = B - b0;
  if (rem <= 2 * bs) {
    const int64_t half = (rem + 1) / 2;
    if (half < nb) nb = half;
  }
  if (nb < 1) nb = 1;
  return nb < rem ? nb : rem;
}

// a contiguous B-slice of a contiguous 5-D host tensor, re-homed at `dev`
volstn_tensor5 slice_desc(const volstn_tensor5 *t, int64_t b0, int64_t nb, void *dev) {
  volstn_tensor5 d = *t;
  d.data = dev;
  d.size[0] = nb;
  (void)b0;
  return d;
}

bool same_desc(const volstn_tensor5 &a, const volstn_tensor5 &b) {
  if (a.data != b.data) return false;
  for (int i = 0; i < 5; i++)
    if (a.size[i] != b.size[i] || a.stride[i] != b.stride[i]) return false;
  return true;
}

int ensure_keep(volstn_host_ctx *c, void **buf, size_t *cap, size_t bytes) {
  if (*cap >= bytes) return VOLSTN_OK;
  int rc = sync_all(c);
  if (rc) return rc;
  if (*buf) VOLSTN_CUDA_TRY(cudaFree(*buf));
  *buf = nullptr;
  *cap = 0;
  VOLSTN_CUDA_TRY(cudaMalloc(buf, bytes));
  *cap = bytes;
  return VOLSTN_OK;
}

size_t sample_elems(const volstn_tensor5 *t) { return (size_t)(t->size[1] * t->size[2] * t->size[3] * t->size[4]); }

// ---------------------------------------------------------------------------------------------
// generic batch-sliced staging for the fused modules: every tensor is contiguous with the batch outermost
// ---------------------------------------------------------------------------------------------
struct HostSpan {
  const void *src;  // host pointer (inputs) ...
  void *dst;        // ... or host pointer (outputs)
  size_t per;       // bytes per sample
  bool zero;        // outputs only: zero-fill the device slice before the launch
};

template <typename Launch>
int run_sliced(volstn_host_ctx *c, int64_t B, const HostSpan *in, int nin, const HostSpan *out, int nout, size_t scratch_per,
               Launch launch) {
  size_t per = scratch_per;
  for (int i = 0; i < nin; i++) per += in[i].per;
  for (int i = 0; i < nout; i++) per += out[i].per;
  const int64_t bs = slice_batch(B, per);
  size_t need = align_up(scratch_per * bs);
  for (int i = 0; i < nin; i++) need += align_up(in[i].per * bs);
  for (int i = 0; i < nout; i++) need += align_up(out[i].per * bs);
  int rc = VOLSTN_OK;
  int slice = 0;
  for (int64_t b0 = 0; b0 < B && rc == VOLSTN_OK; b0 += bs, slice++) {
    const int64_t nb = std::min(bs, B - b0);
    const int s = slice % volstn_host_ctx::kSlots;
    if ((rc = ensure_arena(c, s, need))) break;
    cudaStream_t st = c->stream[s];
    char *p = static_cast<char *>(c->arena[s]);
    void *din[8], *dout[8];
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < nin && e == cudaSuccess; i++) {
      din[i] = p;
      p += align_up(in[i].per * bs);
      e = cudaMemcpyAsync(din[i], static_cast<const char *>(in[i].src) + b0 * in[i].per, nb * in[i].per,
                          cudaMemcpyHostToDevice, st);
    }
    for (int i = 0; i < nout && e == cudaSuccess; i++) {
      dout[i] = p;
      p += align_up(out[i].per * bs);
      if (out[i].zero && out[i].per) e = cudaMemsetAsync(dout[i], 0, nb * out[i].per, st);
    }
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
    if ((rc = launch(nb, din, dout, static_cast<void *>(p), scratch_per * bs, st))) break;
    for (int i = 0; i < nout && e == cudaSuccess; i++)
      if (out[i].per)
        e = cudaMemcpyAsync(static_cast<char *>(out[i].dst) + b0 * out[i].per, dout[i], nb * out[i].per,
                            cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
  }
  const int rs = sync_all(c);
  return rc ? rc : rs;
}

volstn_tensor5 rehome(const volstn_tensor5 *t, int64_t nb, void *dev) {
  volstn_tensor5 d = *t;
  d.data = dev;
  d.size[0] = nb;
  return d;
}

int check_host_fused(volstn_host_ctx *c, int dtype, const volstn_tensor5 *vol, const volstn_tensor5 *outlike) {
  if (!c || !vol || !outlike) return VOLSTN_ERR_ARG;
  if (dtype == VOLSTN_F64) return VOLSTN_ERR_UNSUPPORTED;
  if (dtype != VOLSTN_F32) return VOLSTN_ERR_ARG;
  if (!desc_contiguous(vol) || !desc_contiguous(outlike)) return VOLSTN_ERR_LAYOUT;
  if (outlike->size[0] != vol->size[0] || outlike->size[4] != vol->size[4]) return VOLSTN_ERR_SHAPE;
  return VOLSTN_OK;
}

}  // namespace
}  // namespace volstn

using namespace volstn;

extern "C" {

int volstn_b200_host_ctx_create(int device, volstn_host_ctx **out) {
  if (!out) return VOLSTN_ERR_ARG;
  *out = nullptr;
  int rc = volstn_b200_device_check(device);
  if (rc) return rc;
  DeviceGuard g(device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  volstn_host_ctx *c = new (std::nothrow) volstn_host_ctx();
  if (!c) return VOLSTN_ERR_ARG;
  c->device = device;
  for (int s = 0; s < volstn_host_ctx::kSlots; s++) {
    c->arena[s] = nullptr;
    c->cap[s] = 0;
    c->stream[s] = nullptr;
  }
  c->keep_vol = c->keep_grid = nullptr;
  c->keep_vol_cap = c->keep_grid_cap = 0;
  c->keep_valid = false;
  for (int s = 0; s < volstn_host_ctx::kSlots; s++) {
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream[s], cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      volstn_b200_host_ctx_destroy(c);
      return cuda_fail(e);
    }
  }
  *out = c;
  return VOLSTN_OK;
}

int volstn_b200_host_ctx_destroy(volstn_host_ctx *c) {
  if (!c) return VOLSTN_OK;
  DeviceGuard g(c->device);
  for (int s = 0; s < volstn_host_ctx::kSlots; s++) {
    if (c->stream[s]) {
      cudaStreamSynchronize(c->stream[s]);
      cudaStreamDestroy(c->stream[s]);
    }
    if (c->arena[s]) cudaFree(c->arena[s]);
  }
  if (c->keep_vol) cudaFree(c->keep_vol);
  if (c->keep_grid) cudaFree(c->keep_grid);
  delete c;
  return VOLSTN_OK;
}

int volstn_b200_host_pin(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return VOLSTN_ERR_ARG;
  VOLSTN_CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return VOLSTN_OK;
}

int volstn_b200_host_unpin(void *ptr) {
  if (!ptr) return VOLSTN_ERR_ARG;
  VOLSTN_CUDA_TRY(cudaHostUnregister(ptr));
  return VOLSTN_OK;
}

int volstn_b200_host_sampler_forward(volstn_host_ctx *c, int dtype, int G, const volstn_tensor5 *vol,
                                     const volstn_tensor5 *grid, const volstn_tensor5 *out, unsigned flags) {
  if (!c || !vol || !grid || !out) return VOLSTN_ERR_ARG;
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  if (!desc_contiguous(vol) || !desc_contiguous(grid) || !desc_contiguous(out)) return VOLSTN_ERR_LAYOUT;
  if (grid->size[0] != vol->size[0] || out->size[0] != vol->size[0]) return VOLSTN_ERR_SHAPE;
  const size_t esz = dtype == VOLSTN_F32 ? 4 : 8;
  const int64_t B = vol->size[0];
  if (B == 0 || desc_numel(out) == 0) {
    // still validate shapes through the device entry point (no launch happens for empty outputs)
    return sampler_forward_impl(dtype, G, vol, grid, out, flags, nullptr);
  }
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const size_t nv = sample_elems(vol) * esz, ng = sample_elems(grid) * esz, no = sample_elems(out) * esz;
  const int64_t bs = slice_batch(B, nv + ng + no);
  const size_t need = align_up(nv * bs) + align_up(ng * bs) + align_up(no * bs);
  int rc = VOLSTN_OK;
  const bool keep = flags & VOLSTN_HOST_KEEP_INPUTS;
  c->keep_valid = false;
  if (keep) {
    if ((rc = ensure_keep(c, &c->keep_vol, &c->keep_vol_cap, nv * B))) return rc;
    if ((rc = ensure_keep(c, &c->keep_grid, &c->keep_grid_cap, ng * B))) return rc;
  }
  int slice = 0;
  int64_t nb = 0;
  for (int64_t b0 = 0; b0 < B && rc == VOLSTN_OK; b0 += nb, slice++) {
    nb = ramp_slices() ? slice_len(b0, B, bs, slice) : std::min(bs, B - b0);
    const int s = slice % volstn_host_ctx::kSlots;
    if ((rc = ensure_arena(c, s, need))) break;
    char *dv = static_cast<char *>(c->arena[s]);
    char *dg = dv + align_up(nv * bs);
    char *dout = dg + align_up(ng * bs);
    if (keep) {  // upload straight into the persistent copies: the next backward call reads them from there
      dv = static_cast<char *>(c->keep_vol) + b0 * nv;
      dg = static_cast<char *>(c->keep_grid) + b0 * ng;
    }
    cudaStream_t st = c->stream[s];
    cudaError_t e;
    if ((e = cudaMemcpyAsync(dv, static_cast<const char *>(vol->data) + b0 * nv, nb * nv, cudaMemcpyHostToDevice, st)) !=
            cudaSuccess ||
        (e = cudaMemcpyAsync(dg, static_cast<const char *>(grid->data) + b0 * ng, nb * ng, cudaMemcpyHostToDevice, st)) !=
            cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
    volstn_tensor5 tv = slice_desc(vol, b0, nb, dv), tg = slice_desc(grid, b0, nb, dg), to = slice_desc(out, b0, nb, dout);
    if ((rc = sampler_forward_impl(dtype, G, &tv, &tg, &to, flags, st))) break;
    if ((e = cudaMemcpyAsync(static_cast<char *>(out->data) + b0 * no, dout, nb * no, cudaMemcpyDeviceToHost, st)) !=
        cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
  }
  const int rs = sync_all(c);
  if (keep && rc == VOLSTN_OK && rs == VOLSTN_OK) {
    c->keep_valid = true;
    c->keep_dtype = dtype;
    c->keep_vol_key = *vol;
    c->keep_grid_key = *grid;
  }
  return rc ? rc : rs;
}

int volstn_b200_host_sampler_backward(volstn_host_ctx *c, int dtype, int G, const volstn_tensor5 *vol,
                                      const volstn_tensor5 *grid, const volstn_tensor5 *gradVol,
                                      const volstn_tensor5 *gradGrid, const volstn_tensor5 *gradOut,
                                      unsigned flags) {
  if (!c || !vol || !grid || !gradOut) return VOLSTN_ERR_ARG;
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  const bool og = flags & VOLSTN_BWD_ONLY_GRID, ov = flags & VOLSTN_BWD_ONLY_VOLUME;
  if ((og && ov) || (!og && !gradVol) || (!ov && !gradGrid)) return VOLSTN_ERR_ARG;
  if (!desc_contiguous(vol) || !desc_contiguous(grid) || !desc_contiguous(gradOut)) return VOLSTN_ERR_LAYOUT;
  if (!og && !desc_contiguous(gradVol)) return VOLSTN_ERR_LAYOUT;
  if (!ov && !desc_contiguous(gradGrid)) return VOLSTN_ERR_LAYOUT;
  if (grid->size[0] != vol->size[0] || gradOut->size[0] != vol->size[0]) return VOLSTN_ERR_SHAPE;
  if (!og && gradVol->size[0] != vol->size[0]) return VOLSTN_ERR_SHAPE;
  if (!ov && gradGrid->size[0] != vol->size[0]) return VOLSTN_ERR_SHAPE;
  const size_t esz = dtype == VOLSTN_F32 ? 4 : 8;
  const int64_t B = vol->size[0];
  if (B == 0 || desc_numel(gradOut) == 0) {
    if (!og && (flags & VOLSTN_BWD_ZERO_GRADVOL) && desc_numel(gradVol))
      memset(gradVol->data, 0, (size_t)desc_numel(gradVol) * esz);  // host zero-fill: data movement, not compute
    return sampler_backward_impl(dtype, G, vol, grid, gradVol, gradGrid, gradOut, flags & ~VOLSTN_BWD_ZERO_GRADVOL,
                                 nullptr);
  }
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const size_t nv = sample_elems(vol) * esz, ng = sample_elems(grid) * esz, no = sample_elems(gradOut) * esz;
  const size_t ngv = og ? 0 : nv, ngg = ov ? 0 : ng;
  const int64_t bs = slice_batch(B, nv + ng + no + ngv + ngg);
  const size_t need =
      align_up(nv * bs) + align_up(ng * bs) + align_up(no * bs) + align_up(ngv * bs) + align_up(ngg * bs);
  const bool upload_gv = !og && !(flags & VOLSTN_BWD_ZERO_GRADVOL);  // accumulate semantics of the reference
  const bool reuse = (flags & VOLSTN_HOST_REUSE_INPUTS) && c->keep_valid && c->keep_dtype == dtype &&
                     same_desc(c->keep_vol_key, *vol) && same_desc(c->keep_grid_key, *grid);
  c->keep_valid = false;  // one forward, one backward
  int rc = VOLSTN_OK;
  int slice = 0;
  int64_t nb = 0;
  for (int64_t b0 = 0; b0 < B && rc == VOLSTN_OK; b0 += nb, slice++) {
    nb = ramp_slices() ? slice_len(b0, B, bs, slice) : std::min(bs, B - b0);
    const int s = slice % volstn_host_ctx::kSlots;
    if ((rc = ensure_arena(c, s, need))) break;
    char *dv = static_cast<char *>(c->arena[s]);
    char *dg = dv + align_up(nv * bs);
    char *dgo = dg + align_up(ng * bs);
    char *dgv = dgo + align_up(no * bs);
    char *dgg = dgv + align_up(ngv * bs);
    cudaStream_t st = c->stream[s];
    cudaError_t e = cudaSuccess;
    auto h2d = [&](char *dst, const volstn_tensor5 *t, size_t per) {
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(dst, static_cast<const char *>(t->data) + b0 * per, nb * per, cudaMemcpyHostToDevice, st);
    };
    if (reuse) {
      dv = static_cast<char *>(c->keep_vol) + b0 * nv;
      dg = static_cast<char *>(c->keep_grid) + b0 * ng;
    } else {
      h2d(dv, vol, nv);
      h2d(dg, grid, ng);
    }
    h2d(dgo, gradOut, no);
    if (upload_gv) h2d(dgv, gradVol, nv);
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
    volstn_tensor5 tv = slice_desc(vol, b0, nb, dv), tg = slice_desc(grid, b0, nb, dg),
                   tgo = slice_desc(gradOut, b0, nb, dgo);
    volstn_tensor5 tgv, tgg;
    if (!og) tgv = slice_desc(gradVol, b0, nb, dgv);
    if (!ov) tgg = slice_desc(gradGrid, b0, nb, dgg);
    if ((rc = sampler_backward_impl(dtype, G, &tv, &tg, og ? nullptr : &tgv, ov ? nullptr : &tgg, &tgo, flags, st)))
      break;
    auto d2h = [&](const volstn_tensor5 *t, const char *src, size_t per) {
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(static_cast<char *>(t->data) + b0 * per, src, nb * per, cudaMemcpyDeviceToHost, st);
    };
    if (!og) d2h(gradVol, dgv, nv);
    if (!ov) d2h(gradGrid, dgg, ng);
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
  }
  const int rs = sync_all(c);
  return rc ? rc : rs;
}

int volstn_b200_host_gridgen_forward(volstn_host_ctx *c, int dtype, const void *T, void *out, int64_t B, int64_t D,
                                     int64_t H, int64_t W) {
  if (!c) return VOLSTN_ERR_ARG;
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  if (B < 0 || D < 2 || H < 2 || W < 2) return VOLSTN_ERR_SHAPE;
  if (B == 0) return VOLSTN_OK;
  if (!T || !out) return VOLSTN_ERR_ARG;
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const size_t esz = dtype == VOLSTN_F32 ? 4 : 8;
  const size_t nt = 16 * esz, no = (size_t)(D * H * W * 4) * esz;
  const int64_t bs = slice_batch(B, nt + no);
  const size_t need = align_up(nt * bs) + align_up(no * bs);
  int rc = VOLSTN_OK;
  int slice = 0;
  for (int64_t b0 = 0; b0 < B && rc == VOLSTN_OK; b0 += bs, slice++) {
    const int64_t nb = std::min(bs, B - b0);
    const int s = slice % volstn_host_ctx::kSlots;
    if ((rc = ensure_arena(c, s, need))) break;
    char *dt = static_cast<char *>(c->arena[s]);
    char *dout = dt + align_up(nt * bs);
    cudaStream_t st = c->stream[s];
    cudaError_t e = cudaMemcpyAsync(dt, static_cast<const char *>(T) + b0 * nt, nb * nt, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
    if ((rc = volstn_b200_gridgen_forward(dtype, dt, dout, nb, D, H, W, st))) break;
    e = cudaMemcpyAsync(static_cast<char *>(out) + b0 * no, dout, nb * no, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
  }
  const int rs = sync_all(c);
  return rc ? rc : rs;
}

int volstn_b200_host_gridgen_backward(volstn_host_ctx *c, int dtype, const void *gradGrid, void *gradT, int64_t B,
                                      int64_t D, int64_t H, int64_t W) {
  if (!c) return VOLSTN_ERR_ARG;
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  if (B < 0 || D < 2 || H < 2 || W < 2) return VOLSTN_ERR_SHAPE;
  if (B == 0) return VOLSTN_OK;
  if (!gradGrid || !gradT) return VOLSTN_ERR_ARG;
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const size_t esz = dtype == VOLSTN_F32 ? 4 : 8;
  const size_t nt = 16 * esz, ng = (size_t)(D * H * W * 4) * esz;
  const int64_t bs = slice_batch(B, nt + ng);
  const size_t ws = volstn_b200_gridgen_backward_workspace(bs, D, H, W);
  const size_t need = align_up(ng * bs) + align_up(nt * bs) + align_up(ws);
  int rc = VOLSTN_OK;
  int slice = 0;
  for (int64_t b0 = 0; b0 < B && rc == VOLSTN_OK; b0 += bs, slice++) {
    const int64_t nb = std::min(bs, B - b0);
    const int s = slice % volstn_host_ctx::kSlots;
    if ((rc = ensure_arena(c, s, need))) break;
    char *dg = static_cast<char *>(c->arena[s]);
    char *dt = dg + align_up(ng * bs);
    char *dws = dt + align_up(nt * bs);
    cudaStream_t st = c->stream[s];
    cudaError_t e =
        cudaMemcpyAsync(dg, static_cast<const char *>(gradGrid) + b0 * ng, nb * ng, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
    // workspace sized for `bs` samples is large enough for the (possibly shorter) last slice
    if ((rc = volstn_b200_gridgen_backward(dtype, dg, dt, nb, D, H, W, dws,
                                           std::max(ws, volstn_b200_gridgen_backward_workspace(nb, D, H, W)), st)))
      break;
    e = cudaMemcpyAsync(static_cast<char *>(gradT) + b0 * nt, dt, nb * nt, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
  }
  const int rs = sync_all(c);
  return rc ? rc : rs;
}

// ---------------------------------------------------------------------------------------------
// fused modules, HOST tensors (contiguous).  The loader-side affine augmentation of the reference runs on CPU
// tensors in the data threads (dataset_2d.lua:285-311): the fused call moves volume + T up and the warped volume
// down -- the 16 B/voxel grid never crosses PCIe.
// ---------------------------------------------------------------------------------------------
int volstn_b200_host_affine_sampler_forward(volstn_host_ctx *c, int dtype, const void *T, const volstn_tensor5 *vol,
                                            const volstn_tensor5 *out, unsigned flags) {
  int rc = check_host_fused(c, dtype, vol, out);
  if (rc) return rc;
  const int64_t B = vol->size[0];
  if (B == 0 || desc_numel(out) == 0) return volstn_b200_affine_sampler_forward(dtype, T, vol, out, flags, nullptr);
  if (!T || !vol->data || !out->data) return VOLSTN_ERR_ARG;
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const HostSpan in[2] = {{vol->data, nullptr, sample_elems(vol) * 4, false}, {T, nullptr, 64, false}};
  const HostSpan outs[1] = {{nullptr, out->data, sample_elems(out) * 4, false}};
  return run_sliced(c, B, in, 2, outs, 1, 0, [&](int64_t nb, void **din, void **dout, void *, size_t, cudaStream_t st) {
    volstn_tensor5 tv = rehome(vol, nb, din[0]), to = rehome(out, nb, dout[0]);
    return volstn_b200_affine_sampler_forward(dtype, din[1], &tv, &to, flags, st);
  });
}

int volstn_b200_host_affine_sampler_backward(volstn_host_ctx *c, int dtype, const void *T, const volstn_tensor5 *vol,
                                             const volstn_tensor5 *gradVol, void *gradT, const volstn_tensor5 *gradOut,
                                             unsigned flags) {
  int rc = check_host_fused(c, dtype, vol, gradOut);
  if (rc) return rc;
  const bool og = flags & VOLSTN_BWD_ONLY_GRID, ov = flags & VOLSTN_BWD_ONLY_VOLUME;
  if ((og && ov) || (!og && !gradVol) || (!ov && !gradT)) return VOLSTN_ERR_ARG;
  if (!og && (!desc_contiguous(gradVol) || gradVol->size[0] != vol->size[0])) return VOLSTN_ERR_LAYOUT;
  const int64_t B = vol->size[0];
  if (B == 0) return VOLSTN_OK;
  if (desc_numel(gradOut) == 0) return VOLSTN_ERR_SHAPE;
  if (!T || !vol->data || !gradOut->data) return VOLSTN_ERR_ARG;
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  // gradVol is zero-filled on the device (the module's :zero(), BilinearSamplerVolumetric.lua:103); accumulate
  // semantics are available through the device entry point only
  const HostSpan in[3] = {{vol->data, nullptr, sample_elems(vol) * 4, false}, {T, nullptr, 64, false},
                          {gradOut->data, nullptr, sample_elems(gradOut) * 4, false}};
  const HostSpan outs[2] = {{nullptr, og ? nullptr : gradVol->data, og ? 0 : sample_elems(vol) * 4, true},
                            {nullptr, gradT, ov ? (size_t)0 : (size_t)64, false}};
  const int64_t oD = gradOut->size[1], oH = gradOut->size[2], oW = gradOut->size[3];
  // the workspace of a slice is bounded by the one of a single-sample launch (CTAs per sample shrink as B grows)
  const size_t ws_per = volstn_b200_affine_sampler_backward_workspace(1, oD, oH, oW);
  return run_sliced(c, B, in, 3, outs, 2, ws_per,
                    [&](int64_t nb, void **din, void **dout, void *ws, size_t ws_bytes, cudaStream_t st) {
                      volstn_tensor5 tv = rehome(vol, nb, din[0]), tgo = rehome(gradOut, nb, din[2]), tgv;
                      if (!og) tgv = rehome(gradVol, nb, dout[0]);
                      return volstn_b200_affine_sampler_backward(dtype, din[1], &tv, og ? nullptr : &tgv, dout[1], &tgo, ws,
                                                                 ws_bytes, flags & ~VOLSTN_BWD_ZERO_GRADVOL, st);
                    });
}

int volstn_b200_host_cage_warp_forward(volstn_host_ctx *c, int dtype, const void *cage, const int64_t cage_dims[3],
                                       const volstn_tensor5 *vol, const volstn_tensor5 *out, unsigned flags) {
  int rc = check_host_fused(c, dtype, vol, out);
  if (rc) return rc;
  const int64_t B = vol->size[0];
  if (B == 0 || desc_numel(out) == 0 || !cage_dims)
    return volstn_b200_cage_warp_forward(dtype, cage, cage_dims, vol, out, flags, nullptr);
  if (!cage || !vol->data || !out->data) return VOLSTN_ERR_ARG;
  for (int i = 0; i < 3; i++)
    if (cage_dims[i] < 1 || cage_dims[i] > 64) return volstn_b200_cage_warp_forward(dtype, cage, cage_dims, vol, out, flags, nullptr);
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const size_t cage_per = (size_t)(3 * cage_dims[0] * cage_dims[1] * cage_dims[2]) * 4;
  const HostSpan in[2] = {{vol->data, nullptr, sample_elems(vol) * 4, false}, {cage, nullptr, cage_per, false}};
  const HostSpan outs[1] = {{nullptr, out->data, sample_elems(out) * 4, false}};
  return run_sliced(c, B, in, 2, outs, 1, 0, [&](int64_t nb, void **din, void **dout, void *, size_t, cudaStream_t st) {
    volstn_tensor5 tv = rehome(vol, nb, din[0]), to = rehome(out, nb, dout[0]);
    return volstn_b200_cage_warp_forward(dtype, din[1], cage_dims, &tv, &to, flags, st);
  });
}

int volstn_b200_host_cage_warp_backward(volstn_host_ctx *c, int dtype, const void *cage, const int64_t cage_dims[3],
                                        const volstn_tensor5 *vol, const volstn_tensor5 *gradVol, void *gradCage,
                                        const volstn_tensor5 *gradOut, unsigned flags) {
  int rc = check_host_fused(c, dtype, vol, gradOut);
  if (rc) return rc;
  const bool og = flags & VOLSTN_BWD_ONLY_GRID, ov = flags & VOLSTN_BWD_ONLY_VOLUME;
  if ((og && ov) || (!og && !gradVol) || (!ov && !gradCage) || !cage_dims) return VOLSTN_ERR_ARG;
  if (!og && (!desc_contiguous(gradVol) || gradVol->size[0] != vol->size[0])) return VOLSTN_ERR_LAYOUT;
  for (int i = 0; i < 3; i++)
    if (cage_dims[i] < 1) return VOLSTN_ERR_SHAPE;
  for (int i = 0; i < 3; i++)
    if (cage_dims[i] > 64) return VOLSTN_ERR_UNSUPPORTED;
  const int64_t B = vol->size[0];
  if (B == 0) return VOLSTN_OK;
  if (desc_numel(gradOut) == 0) return VOLSTN_ERR_SHAPE;
  if (!cage || !vol->data || !gradOut->data) return VOLSTN_ERR_ARG;
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const size_t cage_per = (size_t)(3 * cage_dims[0] * cage_dims[1] * cage_dims[2]) * 4;
  const HostSpan in[3] = {{vol->data, nullptr, sample_elems(vol) * 4, false}, {cage, nullptr, cage_per, false},
                          {gradOut->data, nullptr, sample_elems(gradOut) * 4, false}};
  const HostSpan outs[2] = {{nullptr, og ? nullptr : gradVol->data, og ? 0 : sample_elems(vol) * 4, true},
                            {nullptr, gradCage, ov ? 0 : cage_per, false}};
  return run_sliced(c, B, in, 3, outs, 2, 0, [&](int64_t nb, void **din, void **dout, void *, size_t, cudaStream_t st) {
    volstn_tensor5 tv = rehome(vol, nb, din[0]), tgo = rehome(gradOut, nb, din[2]), tgv;
    if (!og) tgv = rehome(gradVol, nb, dout[0]);
    return volstn_b200_cage_warp_backward(dtype, din[1], cage_dims, &tv, og ? nullptr : &tgv, dout[1], &tgo,
                                          flags & ~VOLSTN_BWD_ZERO_GRADVOL, st);
  });
}

static int host_cumsum(volstn_host_ctx *c, int dtype, const void *in, void *out, int64_t B, int N,
                       const int64_t *dims, bool reverse) {
  if (!c || !dims || N < 1) return VOLSTN_ERR_ARG;
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  int64_t S = 1;
  for (int i = 0; i < N; i++) {
    if (dims[i] < 0) return VOLSTN_ERR_ARG;
    S *= dims[i];
  }
  if (B <= 0 || S == 0) return B < 0 ? VOLSTN_ERR_ARG : VOLSTN_OK;
  if (!in || !out) return VOLSTN_ERR_ARG;
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const size_t esz = dtype == VOLSTN_F32 ? 4 : 8;
  const size_t bytes = (size_t)(B * N * S) * esz;
  int rc = ensure_arena(c, 0, align_up(bytes) * 2);
  if (rc) return rc;
  char *din = static_cast<char *>(c->arena[0]);
  char *dout = din + align_up(bytes);
  cudaStream_t st = c->stream[0];
  VOLSTN_CUDA_TRY(cudaMemcpyAsync(din, in, bytes, cudaMemcpyHostToDevice, st));
  rc = reverse ? volstn_b200_cumsum_backward(dtype, din, dout, B, N, dims, st)
               : volstn_b200_cumsum_forward(dtype, din, dout, B, N, dims, st);
  if (rc) return rc;
  VOLSTN_CUDA_TRY(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, st));
  VOLSTN_CUDA_TRY(cudaStreamSynchronize(st));
  return VOLSTN_OK;
}

int volstn_b200_host_cumsum_forward(volstn_host_ctx *c, int dtype, const void *in, void *out, int64_t B, int N,
                                    const int64_t *dims) {
  return host_cumsum(c, dtype, in, out, B, N, dims, false);
}

int volstn_b200_host_cumsum_backward(volstn_host_ctx *c, int dtype, const void *gradOut, void *gradIn, int64_t B,
                                     int N, const int64_t *dims) {
  return host_cumsum(c, dtype, gradOut, gradIn, B, N, dims, true);
}

}  // extern "C"
