// HBM-bound operators on channels-last rows: Relu, Add, MaxPool, GlobalAveragePool, Softmax.
// Roofline for each is bytes-in + bytes-out over measured HBM bandwidth; the kernels
// use 128-bit accesses along the channel axis whenever C, the pitches and the base pointers allow it,
// grid-stride loops sized in multiples of the SM count, and warp-shuffle reductions.
#include <cfloat>
#include <cstdlib>

#include "internal.h"

namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kSMs = 148;

inline int grid_for(long long work, int threads, int blocks_per_sm = 8) {
  long long b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  const long long cap = (long long)kSMs * blocks_per_sm;
  if (b > cap) b = cap;
  return (int)b;
}
inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
inline bool vec_ok(const TView& v) { return v.C % 4 == 0 && v.ld % 4 == 0 && aligned16(v.p); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------ Relu (relu_op.rs:31-33)
template <bool VEC>
__global__ void relu_kernel(const float* __restrict__ x, float* __restrict__ y, int C, long long pixels, int ldx,
                            int ldy) {
  const int CV = VEC ? C / 4 : C;
  const long long total = pixels * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / CV;
    const int c = (int)(i - pix * CV);
    if (VEC) {
      float4 v = ldg4(x + pix * ldx + c * 4);
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      *reinterpret_cast<float4*>(y + pix * ldy + c * 4) = v;
    } else {
      y[pix * ldy + c] = fmaxf(__ldg(x + pix * ldx + c), 0.f);
    }
  }
}

// ------------------------------------------------------------------ Add (add_op.rs:75 / :84)
template <bool VEC>
__global__ void add_channel_kernel(const float* __restrict__ x, const float* __restrict__ b, float* __restrict__ y,
                                   int C, long long pixels, int ldx, int ldy) {
  const int CV = VEC ? C / 4 : C;
  const long long total = pixels * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / CV;
    const int c = (int)(i - pix * CV);
    if (VEC) {
      float4 v = ldg4(x + pix * ldx + c * 4);
      const float4 bb = ldg4(b + c * 4);
      v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      *reinterpret_cast<float4*>(y + pix * ldy + c * 4) = v;
    } else {
      y[pix * ldy + c] = __ldg(x + pix * ldx + c) + __ldg(b + c);
    }
  }
}

// y[r][c] = x[r][c] + b[(b_rows == 1 ? 0 : r)][c]
__global__ void add_same_kernel(const float* __restrict__ x, const float* __restrict__ b, float* __restrict__ y,
                                int C, long long pixels, int ldx, int ldb, int ldy, int b_bcast) {
  const long long total = pixels * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / C;
    const int c = (int)(i - pix * C);
    const long long bp = b_bcast ? 0 : pix;
    y[pix * ldy + c] = __ldg(x + pix * ldx + c) + __ldg(b + bp * ldb + c);
  }
}

// ------------------------------------------------------------------ MaxPool (max_pool_op.rs:157-360)
// Zero-fill padding (:265-276) and a fold that starts at -FLT_MAX (:337).  Thread per (output pixel, 4 channels).
template <bool VEC, typename IDX>
__global__ void maxpool_kernel(PoolArgs a) {
  // IDX = unsigned (all element counts < 2^31: 32-bit divisions) or long long
  const int CV = VEC ? a.C / 4 : a.C;
  const IDX out_pixels = (IDX)a.N * a.Ho * a.Wo;
  const IDX total = out_pixels * CV;
  for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
    const IDX pix = i / CV;
    const int c = (int)(i - pix * CV);
    const int wo = (int)(pix % a.Wo);
    const IDX t = pix / a.Wo;
    const int ho = (int)(t % a.Ho);
    const IDX n = t / a.Ho;
    const int h0 = ho * a.sh - a.pt, w0 = wo * a.sw - a.pl;
    const float* img = a.x + (long long)n * a.H * a.W * a.ldx + (VEC ? c * 4 : c);
    float4 m = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int r = 0; r < a.kh; ++r) {
      const int h = h0 + r;
      const bool hin = (unsigned)h < (unsigned)a.H;
      for (int s = 0; s < a.kw; ++s) {
        const int w = w0 + s;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // padded taps read as 0.0, like the reference
        if (hin && (unsigned)w < (unsigned)a.W) {
          const float* px = img + (long long)(h * a.W + w) * a.ldx;
          if (VEC) v = ldg4(px);
          else v.x = __ldg(px);
        }
        m.x = fmaxf(m.x, v.x);
        if (VEC) { m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w); }
      }
    }
    float* py = a.y + (long long)pix * a.ldy;
    if (VEC) *reinterpret_cast<float4*>(py + c * 4) = m;
    else py[c] = m.x;
  }
}

// The SqueezeNet pools (3x3, stride 2): consecutive output rows share an input row, so a thread walks down a strip
// of output rows for one (image, output column, 4 channels) and carries the shared row's column-max along: 6 loads
// per output instead of 9.  max is exact, so any association gives the reference's bits (zero-filled padding taps
// and the -FLT_MAX fold start included).
constexpr int kPoolStrip = 14;   // output rows per thread
template <typename IDX>   // unsigned when the thread count fits 31 bits (32-bit divisions), else long long
__global__ void maxpool3x3s2_kernel(PoolArgs a, int strips) {
  const int CV = a.C / 4;
  const IDX total = (IDX)a.N * strips * a.Wo * CV;
  for (IDX j = (IDX)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (IDX)gridDim.x * blockDim.x) {
    const IDX i = a.reverse ? total - 1 - j : j;
    const int c = (int)(i % CV);
    IDX t = i / CV;
    const int wo = (int)(t % a.Wo);
    t /= a.Wo;
    const int strip = (int)(t % strips);
    const int n = (int)(t / strips);
    const int ho_begin = strip * kPoolStrip, ho_end = min(a.Ho, ho_begin + kPoolStrip);
    const int w0 = wo * 2 - a.pl;
    const float* img = a.x + (long long)n * a.H * a.W * a.ldx + c * 4;
    const bool in0 = (unsigned)w0 < (unsigned)a.W, in1 = (unsigned)(w0 + 1) < (unsigned)a.W, in2 = (unsigned)(w0 + 2) < (unsigned)a.W;
    auto rowmax = [&](int h) {
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0;   // padded taps read as 0.0, like the reference
      if ((unsigned)h < (unsigned)a.H) {
        const float* px = img + ((long long)h * a.W + w0) * a.ldx;
        if (in0) v0 = ldg4(px);
        if (in1) v1 = ldg4(px + a.ldx);
        if (in2) v2 = ldg4(px + 2 * a.ldx);
      }
      float4 m;
      m.x = fmaxf(fmaxf(v0.x, v1.x), v2.x); m.y = fmaxf(fmaxf(v0.y, v1.y), v2.y);
      m.z = fmaxf(fmaxf(v0.z, v1.z), v2.z); m.w = fmaxf(fmaxf(v0.w, v1.w), v2.w);
      return m;
    };
    float4 carry = rowmax(ho_begin * 2 - a.pt);
    float* py = a.y + (((long long)n * a.Ho + ho_begin) * a.Wo + wo) * a.ldy + c * 4;
    const long long ystep = (long long)a.Wo * a.ldy;
#pragma unroll 2
    for (int ho = ho_begin; ho < ho_end; ++ho) {
      const int h0 = ho * 2 - a.pt;
      const float4 r1 = rowmax(h0 + 1), r2 = rowmax(h0 + 2);
      float4 m;
      m.x = fmaxf(-FLT_MAX, fmaxf(fmaxf(carry.x, r1.x), r2.x)); m.y = fmaxf(-FLT_MAX, fmaxf(fmaxf(carry.y, r1.y), r2.y));
      m.z = fmaxf(-FLT_MAX, fmaxf(fmaxf(carry.z, r1.z), r2.z)); m.w = fmaxf(-FLT_MAX, fmaxf(fmaxf(carry.w, r1.w), r2.w));
      *reinterpret_cast<float4*>(py) = m;
      py += ystep;
      carry = r2;
    }
  }
}

// ------------------------------------------------------------------ GlobalAveragePool (global_average_pool_op.rs:33-52)
// Thread per (image, channel): lanes walk consecutive channels (coalesced), each sums its H*W values in the
// reference's order (sequential over h, w) and divides by the count.
__global__ void gap_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int C, int HW, int ldx,
                           int ldy) {
  const long long total = (long long)N * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / C;
    const int c = (int)(i - n * C);
    const float* p = x + n * HW * (long long)ldx + c;
    float s = 0.f;
    for (int k = 0; k < HW; ++k) s += __ldg(p + (long long)k * ldx);
    y[n * ldy + c] = s / (float)HW;
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide reductions for kThreads threads; `red` is 32 floats of shared memory.
__device__ float block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = ((threadIdx.x & 31) < (blockDim.x >> 5)) ? red[threadIdx.x & 31] : -INFINITY;
  r = warp_max(r);
  __syncthreads();
  return r;
}
__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = ((threadIdx.x & 31) < (blockDim.x >> 5)) ? red[threadIdx.x & 31] : 0.f;
  r = warp_sum(r);
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------ Softmax (softmax_op.rs:45-57)
// One block per image; row = flatten (C,H,W) in NCHW order; x - max, expf, / sum.
__global__ void softmax_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW, int ldx) {
  __shared__ float red[32];
  const long long n = blockIdx.x;
  const int L = C * HW;
  const float* xin = x + n * HW * (long long)ldx;
  float* yo = y + n * (long long)L;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const int c = j / HW, hw = j - c * HW;
    m = fmaxf(m, __ldg(xin + (long long)hw * ldx + c));
  }
  m = block_max(m, red);
  float s = 0.f;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const int c = j / HW, hw = j - c * HW;
    const float e = expf(__ldg(xin + (long long)hw * ldx + c) - m);
    yo[j] = e;
    s += e;
  }
  s = block_sum(s, red);
  for (int j = threadIdx.x; j < L; j += blockDim.x) yo[j] = yo[j] / s;
}

// ------------------------------------------------------------------ GlobalAveragePool + Softmax fused (SqueezeNet tail)
// One 1024-thread block per image.  Phase 1: thread (g, sl) sums channel group g (4 channels, one 128-bit load per
// pixel) over pixel slice sl of kGapSlices -- lanes walk consecutive channel groups, so every load instruction reads
// complete 512-byte runs of one pixel row; the per-thread loads are independent and stay in flight.  Phase 2: the
// slices are combined in a fixed order, divided by H*W, and the softmax runs over the means in shared memory.
constexpr int kGapThreads = 1024;
constexpr int kGapSlices = 4;
template <bool VEC>
__global__ void __launch_bounds__(kGapThreads) gap_softmax_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW,
                                                                  int ldx) {
  extern __shared__ float sm[];  // [kGapSlices][C] partial sums, then [C] means
  __shared__ float red[32];
  float* mean = sm + (size_t)kGapSlices * C;
  const long long n = blockIdx.x;
  const float* xin = x + n * HW * (long long)ldx;
  const int sl = threadIdx.x >> 8;            // pixel slice 0..3
  const int G = VEC ? C / 4 : C;
  for (int g = threadIdx.x & 255; g < G; g += 256) {
    if (VEC) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 11
      for (int k = sl; k < HW; k += kGapSlices) {
        const float4 v = ldg4(xin + (long long)k * ldx + g * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(sm + (size_t)sl * C + g * 4) = acc;
    } else {
      float acc = 0.f;
#pragma unroll 4
      for (int k = sl; k < HW; k += kGapSlices) acc += __ldg(xin + (long long)k * ldx + g);
      sm[(size_t)sl * C + g] = acc;
    }
  }
  __syncthreads();
  float m = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s = ((sm[c] + sm[C + c]) + (sm[2 * C + c] + sm[3 * C + c])) / (float)HW;
    mean[c] = s;
    m = fmaxf(m, s);
  }
  m = block_max(m, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float e = expf(mean[c] - m);
    mean[c] = e;
    s += e;
  }
  s = block_sum(s, red);
  float* yo = y + n * (long long)C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) yo[c] = mean[c] / s;
}

}  // namespace

int launch_relu(TView x, TView y, cudaStream_t st) {
  const long long pixels = x.pixels();
  if (pixels == 0 || x.C == 0) return 0;
  if (vec_ok(x) && vec_ok(y))
    relu_kernel<true><<<grid_for(pixels * (x.C / 4), kThreads), kThreads, 0, st>>>(x.p, y.p, x.C, pixels, x.ld, y.ld);
  else
    relu_kernel<false><<<grid_for(pixels * x.C, kThreads), kThreads, 0, st>>>(x.p, y.p, x.C, pixels, x.ld, y.ld);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_add_channel(TView x, const float* b, TView y, cudaStream_t st) {
  const long long pixels = x.pixels();
  if (pixels == 0 || x.C == 0) return 0;
  if (vec_ok(x) && vec_ok(y) && aligned16(b))
    add_channel_kernel<true><<<grid_for(pixels * (x.C / 4), kThreads), kThreads, 0, st>>>(x.p, b, y.p, x.C, pixels,
                                                                                        x.ld, y.ld);
  else
    add_channel_kernel<false><<<grid_for(pixels * x.C, kThreads), kThreads, 0, st>>>(x.p, b, y.p, x.C, pixels, x.ld,
                                                                                   y.ld);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_add_same(TView x, TView b, TView y, cudaStream_t st) {
  const long long pixels = x.pixels();
  if (pixels == 0 || x.C == 0) return 0;
  const int bcast = (b.pixels() == 1 && pixels != 1) ? 1 : 0;
  add_same_kernel<<<grid_for(pixels * x.C, kThreads), kThreads, 0, st>>>(x.p, b.p, y.p, x.C, pixels, x.ld, b.ld,
                                                                        y.ld, bcast);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_maxpool(const PoolArgs& a, cudaStream_t st) {
  const long long out_pixels = (long long)a.N * a.Ho * a.Wo;
  if (out_pixels == 0 || a.C == 0) return 0;
  const bool vec = a.C % 4 == 0 && a.ldx % 4 == 0 && a.ldy % 4 == 0 && aligned16(a.x) && aligned16(a.y);
  const bool small = out_pixels * a.C < (1ll << 31) && (long long)a.N * a.H * a.W < (1ll << 31);
  const long long work = out_pixels * (vec ? a.C / 4 : a.C);
  const int grid = grid_for(work, kThreads, 32);
  static const int no_strip = [] { const char* e = getenv("B200_POOL_NO_STRIP"); return e ? atoi(e) : 0; }();   // A/B timing only
  if (!no_strip && vec && a.kh == 3 && a.kw == 3 && a.sh == 2 && a.sw == 2) {
    const int strips = (a.Ho + kPoolStrip - 1) / kPoolStrip;
    const long long threads = (long long)a.N * strips * a.Wo * (a.C / 4);
    if (threads < (1ll << 31)) maxpool3x3s2_kernel<unsigned><<<grid_for(threads, kThreads, 32), kThreads, 0, st>>>(a, strips);
    else maxpool3x3s2_kernel<long long><<<grid_for(threads, kThreads, 32), kThreads, 0, st>>>(a, strips);
    B200_CUDA(cudaGetLastError());
    return 0;
  }
  if (vec && small) maxpool_kernel<true, unsigned><<<grid, kThreads, 0, st>>>(a);
  else if (vec) maxpool_kernel<true, long long><<<grid, kThreads, 0, st>>>(a);
  else if (small) maxpool_kernel<false, unsigned><<<grid, kThreads, 0, st>>>(a);
  else maxpool_kernel<false, long long><<<grid, kThreads, 0, st>>>(a);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_global_avgpool(TView x, TView y, cudaStream_t st) {
  if (x.N == 0 || x.C == 0) return 0;
  gap_kernel<<<grid_for((long long)x.N * x.C, kThreads), kThreads, 0, st>>>(x.p, y.p, x.N, x.C, x.H * x.W, x.ld,
                                                                          y.ld);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_softmax(TView x, float* y, cudaStream_t st) {
  if (x.N == 0 || x.C == 0) return 0;
  softmax_kernel<<<x.N, kThreads, 0, st>>>(x.p, y, x.C, x.H * x.W, x.ld);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_gap_softmax(TView x, float* y, cudaStream_t st) {
  if (x.N == 0 || x.C == 0) return 0;
  const size_t smem = (size_t)(kGapSlices + 1) * x.C * sizeof(float);
  if (smem > 48 * 1024) B200_FAIL(B200_EUNSUPPORTED, "gap_softmax: C=%d too large for the fused tail", x.C);
  if (vec_ok(x))
    gap_softmax_kernel<true><<<x.N, kGapThreads, smem, st>>>(x.p, y, x.C, x.H * x.W, x.ld);
  else
    gap_softmax_kernel<false><<<x.N, kGapThreads, smem, st>>>(x.p, y, x.C, x.H * x.W, x.ld);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
