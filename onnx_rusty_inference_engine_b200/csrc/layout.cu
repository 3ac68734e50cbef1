// Layout kernels: logical NCHW (what the reference's ndarray holds and what crosses the ABI) <-> the
// backend's channels-last rows.  All HBM-bound; one coalesced side each, 64-bit indexing.
#include "internal.h"

namespace b200 {
namespace {

constexpr int kThreads = 256;

inline int grid_for(long long work, int threads, int max_blocks = 148 * 16) {
  long long b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

// dst[n][h][w][0..ld) <- src[n][c][h][w].  Thread per (pixel, 4 lanes): reads are coalesced per channel plane
// (consecutive threads = consecutive pixels), writes are 128-bit per pixel.  IDX = unsigned when counts fit 31 bits.
template <typename IDX>
__global__ void nchw_to_rows_vec4_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, IDX HW, IDX pixels, int ld,
                                         int lanes4 /* number of 4-lane groups written per pixel */) {
  const IDX total = pixels * lanes4;
  for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
    const IDX pix = i % pixels;          // pixel fastest: coalesced plane reads
    const int q = (int)(i / pixels);
    const IDX n = pix / HW, hw = pix - n * HW;
    const float* s = src + ((long long)n * C + q * 4) * HW + hw;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q * 4 + 0 < C) v.x = __ldg(s);
    if (q * 4 + 1 < C) v.y = __ldg(s + (long long)HW);
    if (q * 4 + 2 < C) v.z = __ldg(s + 2 * (long long)HW);
    if (q * 4 + 3 < C) v.w = __ldg(s + 3 * (long long)HW);
    *reinterpret_cast<float4*>(dst + (long long)pix * ld + q * 4) = v;
  }
}

__global__ void nchw_to_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, long long HW,
                                    long long pixels, int ld, int lanes /* C or ld (zero-filling) */) {
  const long long total = pixels * lanes;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / lanes;
    const int c = (int)(i - pix * lanes);
    const long long n = pix / HW, hw = pix - n * HW;
    float v = 0.f;
    if (c < C) v = __ldg(src + (n * C + c) * HW + hw);
    dst[pix * ld + c] = v;
  }
}

// dst[n][c][h][w] <- src rows; thread per output element: writes coalesced.
__global__ void rows_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, long long HW,
                                    long long pixels, int ld) {
  const long long total = pixels * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long hw = i % HW;
    const long long nc = i / HW;
    const int c = (int)(nc % C);
    const long long n = nc / C;
    dst[i] = __ldg(src + (n * HW + hw) * ld + c);
  }
}

__global__ void copy_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, long long pixels,
                                 int lds, int ldd) {
  const long long total = pixels * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / C;
    const int c = (int)(i - pix * C);
    dst[pix * ldd + c] = __ldg(src + pix * lds + c);
  }
}

__global__ void copy_rows_vec4_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int C4,
                                      long long pixels, int lds4, int ldd4) {
  const long long total = pixels * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / C4;
    const int c = (int)(i - pix * C4);
    dst[pix * ldd4 + c] = __ldg(src + pix * lds4 + c);
  }
}

// dst[(n + n0, h + h0, w + w0), c] <- src[(n, h, w), c]: a block of a larger tensor (Concat along N / H / W)
__global__ void copy_block_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W, long long pixels, int lds,
                                  int Hd, int Wd, int ldd, int n0, int h0, int w0) {
  const long long total = pixels * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / C;
    const int c = (int)(i - pix * C);
    const int w = (int)(pix % W);
    const long long t = pix / W;
    const int h = (int)(t % H);
    const long long n = t / H;
    dst[(((n + n0) * Hd + (h + h0)) * (long long)Wd + (w + w0)) * ldd + c] = __ldg(src + pix * lds + c);
  }
}

__global__ void transpose2d_kernel(const float* __restrict__ src, int R, int C, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R && c < C) ? src[(long long)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[(long long)c * R + r] = tile[threadIdx.x][j];
  }
}

__global__ void fill_zero_kernel(float* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = 0.f;
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

int launch_nchw_to_rows(const float* src, TView dst, bool zero_pad_lanes, cudaStream_t st) {
  const long long pixels = dst.pixels();
  if (pixels == 0 || dst.C == 0) return 0;
  const int lanes = zero_pad_lanes ? dst.ld : dst.C;
  if (lanes % 4 == 0 && dst.ld % 4 == 0 && aligned16(dst.p) && (zero_pad_lanes || dst.C % 4 == 0)) {
    // 128-bit row writes (the SqueezeNet input: C = 3 stored with a zero 4th lane)
    const int lanes4 = lanes / 4;
    const int grid = grid_for(pixels * lanes4, kThreads, 148 * 32);
    if (pixels * lanes4 < (1ll << 31))
      nchw_to_rows_vec4_kernel<unsigned><<<grid, kThreads, 0, st>>>(src, dst.p, dst.C, (unsigned)((long long)dst.H * dst.W),
                                                                   (unsigned)pixels, dst.ld, lanes4);
    else
      nchw_to_rows_vec4_kernel<long long><<<grid, kThreads, 0, st>>>(src, dst.p, dst.C, (long long)dst.H * dst.W, pixels, dst.ld, lanes4);
    B200_CUDA(cudaGetLastError());
    return 0;
  }
  nchw_to_rows_kernel<<<grid_for(pixels * lanes, kThreads), kThreads, 0, st>>>(
      src, dst.p, dst.C, (long long)dst.H * dst.W, pixels, dst.ld, lanes);
  B200_CUDA(cudaGetLastError());
  return 0;
}

// 2x2 space-to-depth of an NCHW tensor into channels-last rows: dst[((n*H2 + y)*W2 + x)*4C + (dy*2+dx)*C + c] =
// src[((n*C + c)*H + 2y+dy)*W + 2x+dx].  Thread per (output pixel, 16-byte chunk), pixel fastest, so the plane reads of a
// warp are runs of consecutive (dx = 0, 1) pairs.
__global__ void nchw_to_s2d_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W, long long pixels /* N*H2*W2 */) {
  const int W2 = W >> 1, H2 = H >> 1, CS = 4 * C, nq = CS >> 2;
  const long long total = pixels * nq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % pixels;
    const int q = (int)(i / pixels);
    const int x = (int)(pix % W2);
    const long long t = pix / W2;
    const int y = (int)(t % H2);
    const long long n = t / H2;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = q * 4 + e;
      const int d = j / C, c = j - d * C;
      v[e] = __ldg(src + ((n * C + c) * H + 2 * y + (d >> 1)) * (long long)W + 2 * x + (d & 1));
    }
    *reinterpret_cast<float4*>(dst + pix * CS + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// The same for C channels known at compile time (C = 3: the image stem): thread per output pixel, one 64-bit load per
// (c, dy) -- a warp reads 256 contiguous bytes of an input row -- and 4*C/4 128-bit stores, 16*C contiguous bytes per pixel.
template <int C>
__global__ void nchw_to_s2d_fixed_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, long long pixels,
                                         int* __restrict__ nonfinite) {
  const int W2 = W >> 1, H2 = H >> 1;
  uint32_t mx = 0;   // largest |bits| seen: >= 0x7f800000 means Inf / NaN (the split-precision paths are exact for finite data only)
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < pixels; pix += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % W2);
    const long long t = pix / W2;
    const int y = (int)(t % H2);
    const long long n = t / H2;
    float v[4 * C];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float2 pr = __ldg(reinterpret_cast<const float2*>(src + ((n * C + c) * H + 2 * y + dy) * (long long)W + 2 * x));
        v[(dy * 2 + 0) * C + c] = pr.x;
        v[(dy * 2 + 1) * C + c] = pr.y;
        mx = max(mx, max(__float_as_uint(pr.x) & 0x7fffffffu, __float_as_uint(pr.y) & 0x7fffffffu));
      }
    float4* o = reinterpret_cast<float4*>(dst + pix * (4 * C));
#pragma unroll
    for (int q = 0; q < C; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  if (nonfinite && mx >= 0x7f800000u) *nonfinite = 1;
}

// Any Inf / NaN in p[0, n)?  (input stages that have no kernel of their own to carry the check)
__global__ void nonfinite_scan_kernel(const float* __restrict__ p, size_t n, int* __restrict__ nonfinite) {
  uint32_t mx = 0;
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
    mx = max(max(mx, __float_as_uint(v.x) & 0x7fffffffu), max(__float_as_uint(v.y) & 0x7fffffffu, max(__float_as_uint(v.z) & 0x7fffffffu, __float_as_uint(v.w) & 0x7fffffffu)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) mx = max(mx, __float_as_uint(p[n4 * 4 + threadIdx.x]) & 0x7fffffffu);
  if (mx >= 0x7f800000u) *nonfinite = 1;
}

int launch_nonfinite_scan(const float* p, size_t n, int* flag, cudaStream_t st) {
  if (n == 0 || !flag) return 0;
  if ((((uintptr_t)p) & 15) != 0) B200_FAIL(B200_EINVAL, "non-finite scan: input must be 16-byte aligned");
  nonfinite_scan_kernel<<<grid_for((long long)(n / 4 + 1), kThreads, 148 * 8), kThreads, 0, st>>>(p, n, flag);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_nchw_to_s2d(const float* src, int N, int C, int H, int W, float* dst, cudaStream_t st, int* nonfinite) {
  const long long pixels = (long long)N * (H / 2) * (W / 2);
  if (pixels == 0 || C == 0) return 0;
  if (C == 3 && (((uintptr_t)src) & 7) == 0 && (((uintptr_t)dst) & 15) == 0) {   // W even: every row pair is 8-byte aligned
    nchw_to_s2d_fixed_kernel<3><<<grid_for(pixels, kThreads, 148 * 32), kThreads, 0, st>>>(src, dst, H, W, pixels, nonfinite);
    B200_CUDA(cudaGetLastError());
    return 0;
  }
  if (nonfinite) B200_TRY(launch_nonfinite_scan(src, (size_t)N * C * H * W, nonfinite, st));
  nchw_to_s2d_kernel<<<grid_for(pixels * C, kThreads, 148 * 32), kThreads, 0, st>>>(src, dst, C, H, W, pixels);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_rows_to_nchw(TView src, float* dst, cudaStream_t st) {
  const long long pixels = src.pixels();
  if (pixels == 0 || src.C == 0) return 0;
  rows_to_nchw_kernel<<<grid_for(pixels * src.C, kThreads), kThreads, 0, st>>>(
      src.p, dst, src.C, (long long)src.H * src.W, pixels, src.ld);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_copy_rows(TView src, TView dst, cudaStream_t st) {
  const long long pixels = src.pixels();
  if (pixels == 0 || src.C == 0) return 0;
  if (src.C % 4 == 0 && src.ld % 4 == 0 && dst.ld % 4 == 0 && aligned16(src.p) && aligned16(dst.p)) {
    copy_rows_vec4_kernel<<<grid_for(pixels * (src.C / 4), kThreads), kThreads, 0, st>>>(
        (const float4*)src.p, (float4*)dst.p, src.C / 4, pixels, src.ld / 4, dst.ld / 4);
  } else {
    copy_rows_kernel<<<grid_for(pixels * src.C, kThreads), kThreads, 0, st>>>(src.p, dst.p, src.C, pixels,
                                                                               src.ld, dst.ld);
  }
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_copy_block(TView src, TView dst, int n0, int h0, int w0, cudaStream_t st) {
  const long long pixels = src.pixels();
  if (pixels == 0 || src.C == 0) return 0;
  if (src.C != dst.C || n0 + src.N > dst.N || h0 + src.H > dst.H || w0 + src.W > dst.W) B200_FAIL(B200_EINVAL, "copy_block: block does not fit");
  copy_block_kernel<<<grid_for(pixels * src.C, kThreads), kThreads, 0, st>>>(src.p, dst.p, src.C, src.H, src.W, pixels, src.ld, dst.H, dst.W, dst.ld, n0, h0, w0);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_transpose2d(const float* src, int R, int C, float* dst, cudaStream_t st) {
  if (R == 0 || C == 0) return 0;
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  transpose2d_kernel<<<grid, block, 0, st>>>(src, R, C, dst);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_fill_zero(float* p, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  fill_zero_kernel<<<grid_for((long long)n, kThreads), kThreads, 0, st>>>(p, n);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
