// host_api.cu -- HOST-buffer entry points (what `require 'libvolstn'` binds for torch.FloatTensor /
// torch.DoubleTensor in the reference, stn3dtorch/init.c:13-21).
//
// There is no CPU compute path: host tensors are cut into batch slices (samples are independent:
// batch is the outermost loop of every reference function, ...c:482,641) and each slice is
// H2D-copied, processed by the sm_100a kernels and D2H-copied on one of kSlots streams, so that the
// upload of slice i+1, the kernels of slice i and the download of slice i-1 overlap (the B200 has
// independent copy engines per direction).  A synchronous call returns after the last download; with
// VOLSTN_HOST_ASYNC it returns after the last enqueue and volstn_b200_host_wait() completes it, so that the
// uploads of the next call overlap the downloads of this one.
//
// Host tensors may be STRIDED on every dimension, like the reference's CPU code honours them (...c:460-473,
// 607-630): a slice is moved with one cudaMemcpyAsync when it is dense, otherwise with pitched
// cudaMemcpy2DAsync copies over its contiguous runs (copy_view below).  The device copies are always
// contiguous, so the kernels are the packed ones whatever the host layout is.
#include <sched.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "volstn_common.cuh"
#include "volstn_internal.h"

// streams (= slices in flight) per context; 3 = upload of slice i+1, kernels of slice i, download of slice i-1
#ifndef VOLSTN_HOST_SLOTS
#define VOLSTN_HOST_SLOTS 3
#endif

struct volstn_host_ctx {
  static constexpr int kSlice = VOLSTN_HOST_SLOTS;   // slices in flight per call
  // Two sets of streams and staging arenas, used by alternate sampler calls: a stream serialises its work, so with
  // one set the uploads of an asynchronous call could only start as the previous call's downloads drained; with two,
  // the forward of step i+1 uploads while the backward of step i still downloads (8.9 instead of 13 ms per 64^3 x 64
  // step with input reuse, profiles/r02_e2e.md).
  static constexpr int kSlots = 2 * kSlice;
  int call_set;                                      // 0 / 1: the set the current sampler call uses
  static constexpr int kRing = 8;    // tickets of asynchronous calls that can be waited for individually
  static constexpr int kKeepSets = 2;
  int device;
  cudaStream_t stream[kSlots];
  void *arena[kSlots];
  size_t cap[kSlots];
  // Device copies of a forward call's inputs (VOLSTN_HOST_KEEP_INPUTS), reused by the next backward call on the same
  // host tensors (VOLSTN_HOST_REUSE_INPUTS; opt-in: the caller promises not to have modified them in between).
  // Two sets, used alternately, so that with asynchronous calls the forward of step i+1 can upload while the
  // backward of step i still reads; `written` / `read` events order the streams that touch a set.
  struct Keep {
    void *vol, *grid;
    size_t vol_cap, grid_cap;
    bool valid;
    int dtype;
    volstn_tensor5 vol_key, grid_key;
    cudaEvent_t written[kSlots], read[kSlots];
    bool has_written, has_read;
  } keep[kKeepSets];
  int keep_cur;  // set written by the most recent KEEP forward
  // asynchronous calls
  uint64_t seq;                          // ticket of the most recently enqueued call (0 = none)
  cudaEvent_t ring[kRing][kSlots];
  bool pending;                          // work enqueued and not yet waited for
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
  c->pending = false;
  return rc;
}

// end of a call: a synchronous call waits for everything; an asynchronous one takes a ticket (events on every stream)
int finish_call(volstn_host_ctx *c, bool async, int rc) {
  if (!async || rc != VOLSTN_OK) {
    const int rs = sync_all(c);
    return rc ? rc : rs;
  }
  c->seq++;
  cudaEvent_t *ev = c->ring[c->seq % volstn_host_ctx::kRing];
  for (int s = 0; s < volstn_host_ctx::kSlots; s++) {
    cudaError_t e = cudaEventRecord(ev[s], c->stream[s]);
    if (e != cudaSuccess) {
      sync_all(c);
      return cuda_fail(e);
    }
  }
  c->pending = true;
  return VOLSTN_OK;
}

// every stream of the context waits for `ev[0..kSlots)` (recorded on the streams of an earlier call)
int wait_events(volstn_host_ctx *c, cudaEvent_t *ev) {
  for (int s = 0; s < volstn_host_ctx::kSlots; s++)
    for (int t = 0; t < volstn_host_ctx::kSlots; t++)
      if (s != t) VOLSTN_CUDA_TRY(cudaStreamWaitEvent(c->stream[s], ev[t], 0));  // same stream: already ordered
  return VOLSTN_OK;
}
int record_events(volstn_host_ctx *c, cudaEvent_t *ev) {
  for (int s = 0; s < volstn_host_ctx::kSlots; s++) VOLSTN_CUDA_TRY(cudaEventRecord(ev[s], c->stream[s]));
  return VOLSTN_OK;
}

// samples per slice: ~kSliceTargetBytes, and enough slices to keep all slots busy
int64_t slice_batch(int64_t B, size_t bytes_per_sample) {
  int64_t bs = (int64_t)(slice_target_bytes() / (bytes_per_sample ? bytes_per_sample : 1));
  const int64_t per_slot = (B + volstn_host_ctx::kSlice - 1) / volstn_host_ctx::kSlice;
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

// ---------------------------------------------------------------------------------------------
// strided host tensors
// ---------------------------------------------------------------------------------------------
// The samples [b0, b0 + nb) of a host tensor as nested (count, element stride) levels, innermost first, adjacent
// dimensions merged when they are contiguous with each other (stride[outer] == count[inner] * stride[inner]).  A
// contiguous tensor collapses to ONE level of stride 1.
struct HostView {
  int n;
  int64_t cnt[6], str[6];
};

HostView host_view(const volstn_tensor5 *t, int64_t nb) {
  HostView v;
  v.n = 0;
  for (int d = 4; d >= 0; d--) {
    const int64_t c = d == 0 ? nb : t->size[d], st = t->stride[d];
    if (c == 1) continue;
    if (v.n && st == v.cnt[v.n - 1] * v.str[v.n - 1]) {
      v.cnt[v.n - 1] *= c;
    } else {
      v.cnt[v.n] = c;
      v.str[v.n] = st;
      v.n++;
    }
  }
  if (v.n == 0) {
    v.cnt[0] = 1;
    v.str[0] = 1;
    v.n = 1;
  }
  return v;
}

// an OUTPUT must not alias itself: positive strides, and sorted by stride every dimension steps over the whole extent
// of the previous one
bool non_overlapping(const volstn_tensor5 *t) {
  int order[5], m = 0;
  for (int d = 0; d < 5; d++)
    if (t->size[d] > 1) order[m++] = d;
  for (int i = 0; i < m; i++)
    for (int j = i + 1; j < m; j++)
      if (t->stride[order[j]] < t->stride[order[i]]) std::swap(order[i], order[j]);
  int64_t extent = 1;  // elements spanned so far
  for (int i = 0; i < m; i++) {
    const int d = order[i];
    if (t->stride[d] < extent) return false;
    extent = t->stride[d] * (t->size[d] - 1) + extent;
  }
  return true;
}
bool strides_non_negative(const volstn_tensor5 *t) {
  for (int d = 0; d < 5; d++)
    if (t->size[d] > 1 && t->stride[d] < 0) return false;
  return true;
}

constexpr int64_t kMaxCopiesPerSlice = 1 << 15;  // beyond this a layout is too fragmented for DMA: VOLSTN_ERR_LAYOUT

// number of copy calls copy_view() needs for this view (to refuse pathological layouts before anything is enqueued)
int64_t view_copies(const HostView &v, size_t esz) {
  int first = 0;                                    // first level above the contiguous run
  int64_t run = 1;
  if (v.str[0] == 1) run = v.cnt[0], first = 1;
  if (first >= v.n) return 1;
  int64_t outer = 1;
  int lo = first + 1;
  if ((size_t)v.str[first] * esz < (size_t)run * esz) lo = first;  // pitch < width (broadcast): looped, not pitched
  for (int l = lo; l < v.n; l++) outer *= v.cnt[l];
  return outer;
}

// host <-> CONTIGUOUS device copy of a view.  `host` points at the view's first element.
cudaError_t copy_view(bool to_device, char *dev, char *host, const HostView &v, size_t esz, cudaStream_t st) {
  int first = 0;
  int64_t run = 1;
  if (v.str[0] == 1) run = v.cnt[0], first = 1;
  const size_t width = (size_t)run * esz;
  const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  if (first >= v.n) return to_device ? cudaMemcpyAsync(dev, host, width, kind, st) : cudaMemcpyAsync(host, dev, width, kind, st);
  // one pitched copy per index of the levels above `first` (none for a two-level view)
  bool pitched = (size_t)v.str[first] * esz >= width;
  const int lo = pitched ? first + 1 : first;
  const int64_t rows = pitched ? v.cnt[first] : 1;
  const size_t hpitch = pitched ? (size_t)v.str[first] * esz : width;
  int64_t idx[6] = {0, 0, 0, 0, 0, 0};
  for (;;) {
    int64_t off = 0;
    for (int l = lo; l < v.n; l++) off += idx[l] * v.str[l];
    char *h = host + off * (int64_t)esz;
    cudaError_t e;
    if (rows == 1)
      e = to_device ? cudaMemcpyAsync(dev, h, width, kind, st) : cudaMemcpyAsync(h, dev, width, kind, st);
    else
      e = to_device ? cudaMemcpy2DAsync(dev, width, h, hpitch, width, (size_t)rows, kind, st)
                    : cudaMemcpy2DAsync(h, hpitch, dev, width, width, (size_t)rows, kind, st);
    if (e != cudaSuccess) return e;
    dev += width * (size_t)rows;
    int l = lo;
    for (; l < v.n; l++) {
      if (++idx[l] < v.cnt[l]) break;
      idx[l] = 0;
    }
    if (l >= v.n) break;
  }
  return cudaSuccess;
}

// a contiguous nb-sample tensor of t's shape at `dev`
volstn_tensor5 device_desc(const volstn_tensor5 *t, int64_t nb, void *dev) {
  volstn_tensor5 d = *t;
  d.data = dev;
  d.size[0] = nb;
  int64_t st = 1;
  for (int i = 4; i >= 0; i--) {
    d.stride[i] = st;
    st *= d.size[i];
  }
  return d;
}

char *sample_ptr(const volstn_tensor5 *t, int64_t b0, size_t esz) {
  return static_cast<char *>(t->data) + b0 * t->stride[0] * (int64_t)esz;
}

// layout admission of one host tensor; `output` tensors must not alias themselves
int check_host_layout(const volstn_tensor5 *t, bool output, int64_t bs, size_t esz) {
  if (!strides_non_negative(t)) return VOLSTN_ERR_LAYOUT;
  if (output && !non_overlapping(t)) return VOLSTN_ERR_LAYOUT;
  if (view_copies(host_view(t, bs), esz) > kMaxCopiesPerSlice) return VOLSTN_ERR_LAYOUT;
  return VOLSTN_OK;
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
    const int s = slice % volstn_host_ctx::kSlice;
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

// The fused modules are NEW entry points (no reference registration to mirror, so the Double registration of
// stn3dtorch/init.c:18-19 does not bind them): float, contiguous host tensors.
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
  memset(static_cast<void *>(c), 0, sizeof(*c));
  c->device = device;
  cudaError_t e = cudaSuccess;
  for (int s = 0; s < volstn_host_ctx::kSlots && e == cudaSuccess; s++)
    e = cudaStreamCreateWithFlags(&c->stream[s], cudaStreamNonBlocking);
  for (int r = 0; r < volstn_host_ctx::kRing && e == cudaSuccess; r++)
    for (int s = 0; s < volstn_host_ctx::kSlots && e == cudaSuccess; s++)
      e = cudaEventCreateWithFlags(&c->ring[r][s], cudaEventDisableTiming);
  for (int k = 0; k < volstn_host_ctx::kKeepSets && e == cudaSuccess; k++)
    for (int s = 0; s < volstn_host_ctx::kSlots && e == cudaSuccess; s++) {
      e = cudaEventCreateWithFlags(&c->keep[k].written[s], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->keep[k].read[s], cudaEventDisableTiming);
    }
  if (e != cudaSuccess) {
    volstn_b200_host_ctx_destroy(c);
    return cuda_fail(e);
  }
  *out = c;
  return VOLSTN_OK;
}

int volstn_b200_host_ctx_destroy(volstn_host_ctx *c) {
  if (!c) return VOLSTN_OK;
  DeviceGuard g(c->device);
  for (int s = 0; s < volstn_host_ctx::kSlots; s++)
    if (c->stream[s]) cudaStreamSynchronize(c->stream[s]);
  for (int s = 0; s < volstn_host_ctx::kSlots; s++) {
    if (c->stream[s]) cudaStreamDestroy(c->stream[s]);
    if (c->arena[s]) cudaFree(c->arena[s]);
  }
  for (int r = 0; r < volstn_host_ctx::kRing; r++)
    for (int s = 0; s < volstn_host_ctx::kSlots; s++)
      if (c->ring[r][s]) cudaEventDestroy(c->ring[r][s]);
  for (int k = 0; k < volstn_host_ctx::kKeepSets; k++) {
    if (c->keep[k].vol) cudaFree(c->keep[k].vol);
    if (c->keep[k].grid) cudaFree(c->keep[k].grid);
    for (int s = 0; s < volstn_host_ctx::kSlots; s++) {
      if (c->keep[k].written[s]) cudaEventDestroy(c->keep[k].written[s]);
      if (c->keep[k].read[s]) cudaEventDestroy(c->keep[k].read[s]);
    }
  }
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

uint64_t volstn_b200_host_ticket(volstn_host_ctx *c) { return c ? c->seq : 0; }

int volstn_b200_host_wait(volstn_host_ctx *c, uint64_t ticket) {
  if (!c) return VOLSTN_ERR_ARG;
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  if (ticket == 0 || ticket >= c->seq) return sync_all(c);
  // a ticket that has left the ring is covered by the oldest one still in it (streams are in order)
  uint64_t t = ticket;
  if (c->seq >= (uint64_t)volstn_host_ctx::kRing && t < c->seq - volstn_host_ctx::kRing + 1)
    t = c->seq - volstn_host_ctx::kRing + 1;
  cudaEvent_t *ev = c->ring[t % volstn_host_ctx::kRing];
  for (int s = 0; s < volstn_host_ctx::kSlots; s++) VOLSTN_CUDA_TRY(cudaEventSynchronize(ev[s]));
  return VOLSTN_OK;
}

// NUMA placement: the CPUs and the memory node next to `device` from sysfs (PCI bus id -> local_cpulist / numa_node).
int volstn_b200_host_numa_node(int device) {
  char bus[32] = {0}, path[128];
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  for (char *p = bus; *p; p++)
    if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

int volstn_b200_host_bind_thread(int device) {
  char bus[32] = {0}, path[128], list[4096];
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    (void)cudaGetLastError();
    return VOLSTN_ERR_NO_DEVICE;
  }
  for (char *p = bus; *p; p++)
    if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE *f = fopen(path, "r");
  if (!f) return VOLSTN_OK;  // no topology information: nothing to bind to
  const bool got = fgets(list, sizeof(list), f) != nullptr;
  fclose(f);
  if (!got) return VOLSTN_OK;
  cpu_set_t allowed, want;
  CPU_ZERO(&want);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return VOLSTN_OK;
  int any = 0;
  for (char *p = list; *p && *p != '\n';) {  // "0-31,64-95"
    char *end;
    long a = strtol(p, &end, 10), b = a;
    if (end == p) break;
    if (*end == '-') b = strtol(end + 1, &end, 10);
    for (long k = a; k <= b && k < CPU_SETSIZE; k++)
      if (CPU_ISSET(k, &allowed)) CPU_SET(k, &want), any = 1;
    p = (*end == ',') ? end + 1 : end;
  }
  if (any) sched_setaffinity(0, sizeof(want), &want);  // memory allocated (first touched) from here on is node-local
  return VOLSTN_OK;
}

int volstn_b200_host_sampler_forward(volstn_host_ctx *c, int dtype, int G, const volstn_tensor5 *vol,
                                     const volstn_tensor5 *grid, const volstn_tensor5 *out, unsigned flags) {
  if (!c || !vol || !grid || !out) return VOLSTN_ERR_ARG;
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  if (grid->size[0] != vol->size[0] || out->size[0] != vol->size[0]) return VOLSTN_ERR_SHAPE;
  const size_t esz = dtype == VOLSTN_F32 ? 4 : 8;
  const int64_t B = vol->size[0];
  // shapes are validated by the device entry point on contiguous descriptors of the same extents
  {
    volstn_tensor5 tv = device_desc(vol, B, vol->data), tg = device_desc(grid, B, grid->data), to = device_desc(out, B, out->data);
    if (B == 0 || desc_numel(out) == 0) return sampler_forward_impl(dtype, G, &tv, &tg, &to, flags, nullptr);
  }
  const size_t nv = sample_elems(vol) * esz, ng = sample_elems(grid) * esz, no = sample_elems(out) * esz;
  const int64_t bs = slice_batch(B, nv + ng + no);
  int rc;
  if ((rc = check_host_layout(vol, false, bs, esz)) || (rc = check_host_layout(grid, false, bs, esz)) ||
      (rc = check_host_layout(out, true, bs, esz)))
    return rc;
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const size_t need = align_up(nv * bs) + align_up(ng * bs) + align_up(no * bs);
  const bool async = flags & VOLSTN_HOST_ASYNC;
  const bool keep = flags & VOLSTN_HOST_KEEP_INPUTS;
  volstn_host_ctx::Keep *k = nullptr;
  if (keep) {
    c->keep_cur = (c->keep_cur + 1) % volstn_host_ctx::kKeepSets;
    k = &c->keep[c->keep_cur];
    k->valid = false;
    if ((rc = ensure_keep(c, &k->vol, &k->vol_cap, nv * B))) return rc;
    if ((rc = ensure_keep(c, &k->grid, &k->grid_cap, ng * B))) return rc;
    // the previous reader of this set (a backward call two forwards ago) must be through before it is overwritten
    if (c->pending && k->has_read && (rc = wait_events(c, k->read))) return rc;
  }
  int slice = 0;
  int64_t nb = 0;
  c->call_set ^= 1;
  for (int64_t b0 = 0; b0 < B && rc == VOLSTN_OK; b0 += nb, slice++) {
    nb = ramp_slices() ? slice_len(b0, B, bs, slice) : std::min(bs, B - b0);
    const int s = c->call_set * volstn_host_ctx::kSlice + slice % volstn_host_ctx::kSlice;
    if ((rc = ensure_arena(c, s, need))) break;
    char *dv = static_cast<char *>(c->arena[s]);
    char *dg = dv + align_up(nv * bs);
    char *dout = dg + align_up(ng * bs);
    if (keep) {  // upload straight into the persistent copies: the next backward call reads them from there
      dv = static_cast<char *>(k->vol) + b0 * nv;
      dg = static_cast<char *>(k->grid) + b0 * ng;
    }
    cudaStream_t st = c->stream[s];
    cudaError_t e;
    if ((e = copy_view(true, dv, sample_ptr(vol, b0, esz), host_view(vol, nb), esz, st)) != cudaSuccess ||
        (e = copy_view(true, dg, sample_ptr(grid, b0, esz), host_view(grid, nb), esz, st)) != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
    volstn_tensor5 tv = device_desc(vol, nb, dv), tg = device_desc(grid, nb, dg), to = device_desc(out, nb, dout);
    if ((rc = sampler_forward_impl(dtype, G, &tv, &tg, &to, flags, st))) break;
    if ((e = copy_view(false, dout, sample_ptr(out, b0, esz), host_view(out, nb), esz, st)) != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
  }
  if (keep && rc == VOLSTN_OK) {
    if ((rc = record_events(c, k->written)) == VOLSTN_OK) {
      k->has_written = true;
      k->valid = true;
      k->dtype = dtype;
      k->vol_key = *vol;
      k->grid_key = *grid;
    }
  }
  rc = finish_call(c, async, rc);
  if (rc != VOLSTN_OK && k) k->valid = false;
  return rc;
}

int volstn_b200_host_sampler_backward(volstn_host_ctx *c, int dtype, int G, const volstn_tensor5 *vol,
                                      const volstn_tensor5 *grid, const volstn_tensor5 *gradVol,
                                      const volstn_tensor5 *gradGrid, const volstn_tensor5 *gradOut,
                                      unsigned flags) {
  if (!c || !vol || !grid || !gradOut) return VOLSTN_ERR_ARG;
  if (dtype != VOLSTN_F32 && dtype != VOLSTN_F64) return VOLSTN_ERR_ARG;
  const bool og = flags & VOLSTN_BWD_ONLY_GRID, ov = flags & VOLSTN_BWD_ONLY_VOLUME;
  if ((og && ov) || (!og && !gradVol) || (!ov && !gradGrid)) return VOLSTN_ERR_ARG;
  if (grid->size[0] != vol->size[0] || gradOut->size[0] != vol->size[0]) return VOLSTN_ERR_SHAPE;
  if (!og && gradVol->size[0] != vol->size[0]) return VOLSTN_ERR_SHAPE;
  if (!ov && gradGrid->size[0] != vol->size[0]) return VOLSTN_ERR_SHAPE;
  const size_t esz = dtype == VOLSTN_F32 ? 4 : 8;
  const int64_t B = vol->size[0];
  if (B == 0 || desc_numel(gradOut) == 0) {
    if (!og && (flags & VOLSTN_BWD_ZERO_GRADVOL) && desc_numel(gradVol)) {
      if (!desc_contiguous(gradVol)) return VOLSTN_ERR_LAYOUT;
      memset(gradVol->data, 0, (size_t)desc_numel(gradVol) * esz);  // host zero-fill: data movement, not compute
    }
    volstn_tensor5 tv = device_desc(vol, B, vol->data), tg = device_desc(grid, B, grid->data),
                   tgo = device_desc(gradOut, B, gradOut->data), tgv, tgg;
    if (!og) tgv = device_desc(gradVol, B, gradVol->data);
    if (!ov) tgg = device_desc(gradGrid, B, gradGrid->data);
    return sampler_backward_impl(dtype, G, &tv, &tg, og ? nullptr : &tgv, ov ? nullptr : &tgg, &tgo,
                                 flags & ~VOLSTN_BWD_ZERO_GRADVOL, nullptr);
  }
  const size_t nv = sample_elems(vol) * esz, ng = sample_elems(grid) * esz, no = sample_elems(gradOut) * esz;
  const size_t ngv = og ? 0 : nv, ngg = ov ? 0 : ng;
  const int64_t bs = slice_batch(B, nv + ng + no + ngv + ngg);
  int rc;
  if ((rc = check_host_layout(vol, false, bs, esz)) || (rc = check_host_layout(grid, false, bs, esz)) ||
      (rc = check_host_layout(gradOut, false, bs, esz)))
    return rc;
  if (!og && (rc = check_host_layout(gradVol, true, bs, esz))) return rc;
  if (!ov && (rc = check_host_layout(gradGrid, true, bs, esz))) return rc;
  DeviceGuard g(c->device);
  if (!g.ok) return VOLSTN_ERR_NO_DEVICE;
  const size_t need =
      align_up(nv * bs) + align_up(ng * bs) + align_up(no * bs) + align_up(ngv * bs) + align_up(ngg * bs);
  const bool async = flags & VOLSTN_HOST_ASYNC;
  const bool upload_gv = !og && !(flags & VOLSTN_BWD_ZERO_GRADVOL);  // accumulate semantics of the reference
  volstn_host_ctx::Keep *k = &c->keep[c->keep_cur];
  const bool reuse = (flags & VOLSTN_HOST_REUSE_INPUTS) && k->valid && k->dtype == dtype &&
                     same_desc(k->vol_key, *vol) && same_desc(k->grid_key, *grid);
  k->valid = false;  // one forward, one backward
  // the forward's uploads into the kept copies ran on other streams of this context
  if (reuse && c->pending && (rc = wait_events(c, k->written))) return rc;
  // the gradVol slices are zero-filled (or uploaded) contiguous device buffers: the flag is consumed here
  const unsigned kflags = flags & ~VOLSTN_BWD_ZERO_GRADVOL;
  int slice = 0;
  int64_t nb = 0;
  c->call_set ^= 1;
  for (int64_t b0 = 0; b0 < B && rc == VOLSTN_OK; b0 += nb, slice++) {
    nb = ramp_slices() ? slice_len(b0, B, bs, slice) : std::min(bs, B - b0);
    const int s = c->call_set * volstn_host_ctx::kSlice + slice % volstn_host_ctx::kSlice;
    if ((rc = ensure_arena(c, s, need))) break;
    char *dv = static_cast<char *>(c->arena[s]);
    char *dg = dv + align_up(nv * bs);
    char *dgo = dg + align_up(ng * bs);
    char *dgv = dgo + align_up(no * bs);
    char *dgg = dgv + align_up(ngv * bs);
    cudaStream_t st = c->stream[s];
    cudaError_t e = cudaSuccess;
    auto h2d = [&](char *dst, const volstn_tensor5 *t) {
      if (e == cudaSuccess) e = copy_view(true, dst, sample_ptr(t, b0, esz), host_view(t, nb), esz, st);
    };
    if (reuse) {
      dv = static_cast<char *>(k->vol) + b0 * nv;
      dg = static_cast<char *>(k->grid) + b0 * ng;
    } else {
      h2d(dv, vol);
      h2d(dg, grid);
    }
    h2d(dgo, gradOut);
    if (upload_gv) h2d(dgv, gradVol);
    else if (!og && e == cudaSuccess) e = cudaMemsetAsync(dgv, 0, nb * nv, st);
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
    volstn_tensor5 tv = device_desc(vol, nb, dv), tg = device_desc(grid, nb, dg), tgo = device_desc(gradOut, nb, dgo);
    volstn_tensor5 tgv, tgg;
    if (!og) tgv = device_desc(gradVol, nb, dgv);
    if (!ov) tgg = device_desc(gradGrid, nb, dgg);
    if ((rc = sampler_backward_impl(dtype, G, &tv, &tg, og ? nullptr : &tgv, ov ? nullptr : &tgg, &tgo, kflags, st)))
      break;
    auto d2h = [&](const volstn_tensor5 *t, char *src) {
      if (e == cudaSuccess) e = copy_view(false, src, sample_ptr(t, b0, esz), host_view(t, nb), esz, st);
    };
    if (!og) d2h(gradVol, dgv);
    if (!ov) d2h(gradGrid, dgg);
    if (e != cudaSuccess) {
      rc = cuda_fail(e);
      break;
    }
  }
  if (reuse && rc == VOLSTN_OK && (rc = record_events(c, k->read)) == VOLSTN_OK) k->has_read = true;
  return finish_call(c, async, rc);
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
    const int s = slice % volstn_host_ctx::kSlice;
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
  // B * blocks-per-sample(B) is not monotone in B: the workspace region must hold the LARGEST requirement of any
  // slice length that can occur (1 .. bs), and the launch is told the real size of the region
  size_t ws = 0;
  for (int64_t nb = 1; nb <= bs; nb++) ws = std::max(ws, volstn_b200_gridgen_backward_workspace(nb, D, H, W));
  const size_t need = align_up(ng * bs) + align_up(nt * bs) + align_up(ws);
  int rc = VOLSTN_OK;
  int slice = 0;
  for (int64_t b0 = 0; b0 < B && rc == VOLSTN_OK; b0 += bs, slice++) {
    const int64_t nb = std::min(bs, B - b0);
    const int s = slice % volstn_host_ctx::kSlice;
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
    if ((rc = volstn_b200_gridgen_backward(dtype, dg, dt, nb, D, H, W, dws, ws, st))) break;
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
