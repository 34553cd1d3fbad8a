// cumsum.cu -- nn.Cumsum on sm_100a: the axial prefix sum that integrates ALIGNet's predicted cage
// deltas into absolute sampling coordinates (reference Cumsum.lua; used at models/alignet_2d.lua:102).
//
// Input B x N x d[0] x ... x d[N-1] (contiguous).  Channel i is scanned along spatial axis i
// (Cumsum.lua:17-19); backward is the suffix sum along the same axis (flip / cumsum / flip,
// Cumsum.lua:42-48).  The reference needs N x (split, squeeze, cumsum) launches forward and
// N x (index, cumsum, index, copy) backward; the whole tensor is <= ~1.3 MB (cage sizes 2..12,
// opts.lua:42), so the cost is launch latency.  Here: ONE launch for all channels.
//
//   * channel N-1 (scan axis = innermost, unit stride): warp-shuffle SEGMENTED scan.  A warp owns
//     floor(32 / L) whole lines when L <= 32 (lanes map to consecutive memory, so loads and
//     stores are coalesced), or walks one long line in 32-element chunks carrying the running
//     total.
//   * channels 0..N-2 (scan axis strided): lanes run along the contiguous inner index, each thread
//     walks its own line sequentially -- coalesced and shuffle-free.
//
// Prefixes accumulate in double and are rounded to `real` once per element: TH's CPU cumsum
// accumulates float tensors in accreal = double, which is what the drop-in must reproduce.
#include "volstn_common.cuh"

namespace volstn {

constexpr int kCSThreads = 256;
constexpr int kMaxN = 6;

struct CumsumArgs {
  const void *in;
  void *out;
  long long B;
  int N;
  long long S;             // prod d
  long long L[kMaxN];      // scan length of channel i
  long long inner[kMaxN];  // stride of the scan axis of channel i
};

template <typename real, bool REVERSE>
__global__ void __launch_bounds__(kCSThreads) k_cumsum(const CumsumArgs a) {
  const int ch = blockIdx.y;
  const long long L = a.L[ch], inner = a.inner[ch];
  const real *__restrict__ in = static_cast<const real *>(a.in);
  real *__restrict__ out = static_cast<real *>(a.out);
  const long long lines_per_sample = a.S / L;

  if (inner != 1) {
    // ---- strided axis: one thread per line, lanes along the contiguous inner index ----
    const long long total = a.B * lines_per_sample;
    for (long long ln = (long long)blockIdx.x * kCSThreads + threadIdx.x; ln < total;
         ln += (long long)gridDim.x * kCSThreads) {
      const long long b = ln / lines_per_sample, r = ln % lines_per_sample;
      const long long o = r / inner, i = r % inner;
      const long long base = (b * a.N + ch) * a.S + o * L * inner + i;
      // the loads of a line are independent: issue eight before the first (ordered) addition, otherwise
      // every step of the scan pays a full memory round trip (9 us for a 384 KB cage, profiles/r01_aux*)
      double acc = 0.0;
      for (long long k0 = 0; k0 < L; k0 += 8) {
        real v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const long long k = k0 + q;
          v[q] = k < L ? in[base + (REVERSE ? (L - 1 - k) : k) * inner] : real(0);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const long long k = k0 + q;
          if (k < L) {
            acc += (double)v[q];
            out[base + (REVERSE ? (L - 1 - k) : k) * inner] = (real)acc;
          }
        }
      }
    }
    return;
  }

  // ---- innermost axis: warp-shuffle segmented scan ----
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * kCSThreads + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * kCSThreads) >> 5;
  const long long total_lines = a.B * lines_per_sample;  // lines of this channel, L contiguous elements each
  if (L <= 32) {
    const int Li = (int)L;
    const int lpw = 32 / Li;          // whole lines per warp
    const int seg = lane / Li;        // which of the warp's lines this lane serves
    const int pos = lane - seg * Li;  // scan position inside the line (in scan direction)
    const bool active_lane = seg < lpw;
    for (long long l0 = warp * lpw; l0 < total_lines; l0 += nwarps * lpw) {
      const long long ln = l0 + seg;
      const bool active = active_lane && ln < total_lines;
      long long p = 0;
      double v = 0.0;
      if (active) {
        const long long b = ln / lines_per_sample, r = ln % lines_per_sample;
        p = (b * a.N + ch) * a.S + r * L + (REVERSE ? (Li - 1 - pos) : pos);
        v = (double)in[p];
      }
      // Hillis-Steele inside the segment: lane adds the value `d` positions earlier in ITS line
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, v, d);
        if (pos >= d) v += up;
      }
      if (active) out[p] = (real)v;
    }
  } else {
    for (long long ln = warp; ln < total_lines; ln += nwarps) {
      const long long b = ln / lines_per_sample, r = ln % lines_per_sample;
      const long long base = (b * a.N + ch) * a.S + r * L;
      double carry = 0.0;
      for (long long k0 = 0; k0 < L; k0 += 32) {
        const long long k = k0 + lane;
        const bool active = k < L;
        const long long p = base + (REVERSE ? (L - 1 - k) : k);
        double v = active ? (double)in[p] : 0.0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, v, d);
          if (lane >= d) v += up;
        }
        v += carry;
        if (active) out[p] = (real)v;
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
  }
}

template <bool REVERSE>
int cumsum_impl(int dtype, const void *in, void *out, int64_t B, int N, const int64_t *dims, cudaStream_t st) {
  if (N < 1 || N > kMaxN || B < 0 || !dims) return VOLSTN_ERR_ARG;
  CumsumArgs a;
  a.in = in;
  a.out = out;
  a.B = B;
  a.N = N;
  a.S = 1;
  for (int i = 0; i < N; i++) {
    if (dims[i] < 0) return VOLSTN_ERR_ARG;
    a.S *= dims[i];
  }
  if (B == 0 || a.S == 0) return VOLSTN_OK;
  if (!in || !out) return VOLSTN_ERR_ARG;
  long long max_blocks = 1;
  for (int i = 0; i < N; i++) {
    a.L[i] = dims[i];
    long long inner = 1;
    for (int k = i + 1; k < N; k++) inner *= dims[k];
    a.inner[i] = inner;
    const long long lines = B * (a.S / a.L[i]);
    long long threads;
    if (inner != 1)
      threads = lines;
    else if (a.L[i] <= 32)
      threads = ((lines + (32 / a.L[i]) - 1) / (32 / a.L[i])) * 32;
    else
      threads = lines * 32;
    const long long blocks = (threads + kCSThreads - 1) / kCSThreads;
    if (blocks > max_blocks) max_blocks = blocks;
  }
  const long long cap = (long long)kSMs * 16;
  if (max_blocks > cap) max_blocks = cap;
  dim3 grid((unsigned)max_blocks, (unsigned)N);
  if (dtype == VOLSTN_F32)
    k_cumsum<float, REVERSE><<<grid, kCSThreads, 0, st>>>(a);
  else if (dtype == VOLSTN_F64)
    k_cumsum<double, REVERSE><<<grid, kCSThreads, 0, st>>>(a);
  else
    return VOLSTN_ERR_ARG;
  return after_launch();
}

}  // namespace volstn

extern "C" {

int volstn_b200_cumsum_forward(int dtype, const void *in, void *out, int64_t B, int N, const int64_t *dims,
                               void *cuda_stream) {
  return volstn::cumsum_impl<false>(dtype, in, out, B, N, dims, static_cast<cudaStream_t>(cuda_stream));
}

int volstn_b200_cumsum_backward(int dtype, const void *gradOut, void *gradIn, int64_t B, int N, const int64_t *dims,
                                void *cuda_stream) {
  return volstn::cumsum_impl<true>(dtype, gradOut, gradIn, B, N, dims, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
