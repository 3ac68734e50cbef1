// CUDA-core fp32 implicit-GEMM convolution on channels-last rows.
//
// Role: (1) the on-GPU cross-check for the tcgen05 3xTF32 path (exact fp32 FMA arithmetic),
//       (2) the kernel for shapes the tensor-core path does not take (tiny C such as MNIST's C=1, ragged M).
// Semantics: conv2d, convolution_op.rs:224-517 -- cross-correlation, group 1, dilation 1,
//   y[n,m,ho,wo] = bias[m] + sum_{r,s,c} x[n,c,ho*sh+r-pt,wo*sw+s-pl] * w[m,c,r,s]   (zero padding)
// optionally followed by the folded per-channel Add (add_op.rs:75) and Relu (relu_op.rs:31-33).
//
// GEMM view: D[P x M] = A[P x K] * B[M x K]^T with P = N*Ho*Wo pixels, K = KH*KW*C ordered (r, s, c) so that
// both A (an input pixel's channels) and B (a weight row) are contiguous along K.
// Tile: 128 pixels x BN channels per 256-thread CTA, BK = 16, register tile 8 x TN, register-staged prefetch.
#include <cstdlib>

#include "internal.h"

namespace b200 {
namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NT = 256;

template <int BN, bool VEC>
__global__ void __launch_bounds__(NT) conv_simt_kernel(ConvArgs a) {
  constexpr int TN = BN / 16;                    // channels per thread
  constexpr int B_PER_T = (BN * BK / 4 + NT - 1) / NT;  // float4 (or 4 scalars) of B per thread
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const long long P = (long long)a.N * a.Ho * a.Wo;
  const long long p0 = (long long)blockIdx.x * BM;
  const int m0 = blockIdx.y * BN;

  // ---- A loader: thread owns pixels (tid & 63) and (tid & 63) + 64, k-quad (tid >> 6)
  const int a_kq = tid >> 6;
  int a_h0[2], a_w0[2];
  const float* a_base[2];
  bool a_valid[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const long long p = p0 + (tid & 63) + 64 * i;
    a_valid[i] = p < P;
    const long long pp = a_valid[i] ? p : 0;
    const int wo = (int)(pp % a.Wo);
    const long long t = pp / a.Wo;
    const int ho = (int)(t % a.Ho);
    const long long n = t / a.Ho;
    a_h0[i] = ho * a.sh - a.pt;
    a_w0[i] = wo * a.sw - a.pl;
    a_base[i] = a.x + n * (long long)a.H * a.W * a.ldx;
  }
  // ---- B loader: float4 index f = tid + j*NT over BN x (BK/4): row n = f % BN, k-quad = f / BN
  float4 ra[2];
  float4 rb[B_PER_T];

  auto load_tiles = [&](int k0) {
    // A
    const int kk = k0 + a_kq * 4;
    if (VEC) {
      // C % 4 == 0: the 4 consecutive k share one tap (r, s) and are 4 consecutive channels
      const int tap = kk / a.C;
      const int c = kk - tap * a.C;
      const int r = tap / a.KW, s = tap - r * a.KW;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int h = a_h0[i] + r, w = a_w0[i] + s;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_valid[i] && kk < a.K && h >= 0 && h < a.H && w >= 0 && w < a.W)
          v = __ldg(reinterpret_cast<const float4*>(a_base[i] + ((long long)h * a.W + w) * a.ldx + c));
        ra[i] = v;
      }
    } else {
      // decode (r, s, c) once for the first of the 4 consecutive k and step it (two divisions per k-block per thread)
      int rr[4], ss[4], cc[4];
      {
        const int tap = kk / a.C;
        int c = kk - tap * a.C, r = tap / a.KW, s = tap - r * a.KW;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          rr[j] = r; ss[j] = s; cc[j] = c;
          if (++c == a.C) { c = 0; if (++s == a.KW) { s = 0; ++r; } }
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = 0.f;
          if (a_valid[i] && kk + j < a.K) {
            const int h = a_h0[i] + rr[j], w = a_w0[i] + ss[j];
            if ((unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W)
              v[j] = __ldg(a_base[i] + ((long long)h * a.W + w) * a.ldx + cc[j]);
          }
        }
        ra[i] = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    // B
#pragma unroll
    for (int j = 0; j < B_PER_T; ++j) {
      const int f = tid + j * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < BN * BK / 4) {
        const int n = f % BN, kq = f / BN;
        const int m = m0 + n, k = k0 + kq * 4;
        if (m < a.M) {
          const float* wrow = a.w + (long long)m * a.ldw;
          if (VEC) {
            if (k < a.K) v = __ldg(reinterpret_cast<const float4*>(wrow + k));
          } else {
            float t[4];
            int tap = k / a.C, c = k - tap * a.C;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              t[q] = 0.f;
              if (k + q < a.K) t[q] = __ldg(wrow + tap * a.wc + c);  // weight taps are pitched by wc >= C
              if (++c == a.C) { c = 0; ++tap; }
            }
            v = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
      }
      rb[j] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int px = (tid & 63) + 64 * i;
      As[a_kq * 4 + 0][px] = ra[i].x;
      As[a_kq * 4 + 1][px] = ra[i].y;
      As[a_kq * 4 + 2][px] = ra[i].z;
      As[a_kq * 4 + 3][px] = ra[i].w;
    }
#pragma unroll
    for (int j = 0; j < B_PER_T; ++j) {
      const int f = tid + j * NT;
      if (f < BN * BK / 4) {
        const int n = f % BN, kq = f / BN;
        Bs[kq * 4 + 0][n] = rb[j].x;
        Bs[kq * 4 + 1][n] = rb[j].y;
        Bs[kq * 4 + 2][n] = rb[j].z;
        Bs[kq * 4 + 3][n] = rb[j].w;
      }
    }
  };

  const int tx = tid & 15;   // channel group
  const int ty = tid >> 4;   // pixel group
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (a.K + BK - 1) / BK;
  load_tiles(0);
  for (int kb = 0; kb < nk; ++kb) {
    store_tiles();
    __syncthreads();
    if (kb + 1 < nk) load_tiles((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[8], bv[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
      av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: + bias, + folded channel add, relu, store at the (possibly channel-offset) destination
  float add[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int m = m0 + tx * TN + j;
    float b = 0.f;
    if (m < a.M) {
      if (a.bias) b = __ldg(a.bias + m);
    }
    add[j] = b;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long p = p0 + ty * 8 + i;
    if (p >= P) continue;
    float* yp = a.y + p * a.ldy;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int m = m0 + tx * TN + j;
      if (m < a.M) {
        float v = acc[i][j] + add[j];                       // conv + bias   (add_bias, convolution_op.rs:705)
        if (a.chan_add) v = v + __ldg(a.chan_add + m);      // separate Add node (add_op.rs:75), kept as a 2nd rounding
        if (a.relu) v = fmaxf(v, 0.f);
        yp[m] = v;
      }
    }
  }
}

template <int BN>
int launch_bn(const ConvArgs& a, bool vec, cudaStream_t st) {
  const long long P = (long long)a.N * a.Ho * a.Wo;
  dim3 grid((unsigned)((P + BM - 1) / BM), (unsigned)((a.M + BN - 1) / BN));
  if (vec)
    conv_simt_kernel<BN, true><<<grid, NT, 0, st>>>(a);
  else
    conv_simt_kernel<BN, false><<<grid, NT, 0, st>>>(a);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// Direct convolution for few output channels and a short reduction (MNIST's first layer: C = 1, 5x5, M = 8).
// The implicit-GEMM tile above wastes most of a 128 x 16 tile there and the layer is bandwidth-class work
// (1.6 KFLOP and 6.3 KB of traffic per image row): one thread per output pixel keeps all MT accumulators in
// registers, walks k = (r, s, c) in the same order as the GEMM kernels, reads its input taps coalesced across
// the warp (consecutive output columns) and the weights as broadcast 128-bit shared-memory loads.
template <int MT>
__global__ void __launch_bounds__(256) conv_direct_kernel(ConvArgs a) {
  extern __shared__ __align__(16) float wsm[];   // [K][MT], zero padded to MT filters
  for (int i = threadIdx.x; i < a.K * MT; i += blockDim.x) {
    const int k = i / MT, m = i - k * MT;
    const int tap = k / a.C, c = k - tap * a.C;
    wsm[i] = (m < a.M) ? __ldg(a.w + (long long)m * a.ldw + (long long)tap * a.wc + c) : 0.f;
  }
  __syncthreads();
  const long long P = (long long)a.N * a.Ho * a.Wo;
  const bool vec_out = (a.M % 4 == 0) && (a.ldy % 4 == 0) && ((((uintptr_t)a.y) & 15) == 0);
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(p % a.Wo);
    const long long t = p / a.Wo;
    const int ho = (int)(t % a.Ho);
    const long long n = t / a.Ho;
    const int h0 = ho * a.sh - a.pt, w0 = wo * a.sw - a.pl;
    const float* img = a.x + n * (long long)a.H * a.W * a.ldx;
    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.f;
    const float4* wrow = reinterpret_cast<const float4*>(wsm);
    for (int r = 0; r < a.KH; ++r) {
      const int h = h0 + r;
      const bool hin = (unsigned)h < (unsigned)a.H;
      for (int sx = 0; sx < a.KW; ++sx) {
        const int w = w0 + sx;
        const bool in = hin && (unsigned)w < (unsigned)a.W;
        const float* px = img + ((long long)h * a.W + w) * a.ldx;
        for (int c = 0; c < a.C; ++c) {
          const float x = in ? __ldg(px + c) : 0.f;   // zero padding, convolution_op.rs:560-663
#pragma unroll
          for (int q = 0; q < MT / 4; ++q) {
            const float4 w4 = wrow[q];
            acc[4 * q + 0] = fmaf(x, w4.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(x, w4.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(x, w4.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(x, w4.w, acc[4 * q + 3]);
          }
          wrow += MT / 4;
        }
      }
    }
    float* py = a.y + p * a.ldy;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < a.M) {
        float v = acc[m];
        if (a.bias) v += __ldg(a.bias + m);          // add_bias, convolution_op.rs:705
        if (a.chan_add) v += __ldg(a.chan_add + m);  // folded Add, add_op.rs:75
        if (a.relu) v = fmaxf(v, 0.f);               // relu_op.rs:31-33
        acc[m] = v;
      }
    }
    if (vec_out) {
#pragma unroll
      for (int q = 0; q < MT / 4; ++q)
        if (4 * q < a.M) *reinterpret_cast<float4*>(py + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    } else {
#pragma unroll
      for (int m = 0; m < MT; ++m)
        if (m < a.M) py[m] = acc[m];
    }
  }
}

// The same with a 4-pixel register tile along the output row (stride-1 columns, KW <= 8): the KW + 3 input values of
// a kernel row are loaded once and shared by the four pixels (2 loads per pixel and kernel row instead of KW), and
// every broadcast weight vector feeds 4 x 4 FMAs.
template <int MT>
__global__ void __launch_bounds__(256) conv_direct4_kernel(ConvArgs a) {
  extern __shared__ __align__(16) float wsm[];   // [K][MT], k = (r*KW + s)*C + c, zero padded to MT filters
  for (int i = threadIdx.x; i < a.K * MT; i += blockDim.x) {
    const int k = i / MT, m = i - k * MT;
    const int tap = k / a.C, c = k - tap * a.C;
    wsm[i] = (m < a.M) ? __ldg(a.w + (long long)m * a.ldw + (long long)tap * a.wc + c) : 0.f;
  }
  __syncthreads();
  const int WQ = (a.Wo + 3) >> 2;
  const long long G = (long long)a.N * a.Ho * WQ;
  const bool vec_out = (a.M % 4 == 0) && (a.ldy % 4 == 0) && ((((uintptr_t)a.y) & 15) == 0);
  const float4* wv = reinterpret_cast<const float4*>(wsm);
  for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < G; gi += (long long)gridDim.x * blockDim.x) {
    const int wq = (int)(gi % WQ);
    const long long t = gi / WQ;
    const int ho = (int)(t % a.Ho);
    const long long n = t / a.Ho;
    const int wo0 = wq * 4;
    const int h0 = ho * a.sh - a.pt, w0 = wo0 - a.pl;   // sw == 1
    const float* img = a.x + n * (long long)a.H * a.W * a.ldx;
    float acc[4][MT];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[q][m] = 0.f;
    for (int r = 0; r < a.KH; ++r) {
      const int h = h0 + r;
      if ((unsigned)h >= (unsigned)a.H) continue;       // a padded row contributes 0 to every tap
      const float* row = img + (long long)h * a.W * a.ldx;
      for (int c = 0; c < a.C; ++c) {
        float x[11];
#pragma unroll
        for (int j = 0; j < 11; ++j) {
          const int w = w0 + j;
          x[j] = (j < a.KW + 3 && (unsigned)w < (unsigned)a.W) ? __ldg(row + (long long)w * a.ldx + c) : 0.f;
        }
#pragma unroll
        for (int sx = 0; sx < 8; ++sx) {
          if (sx < a.KW) {
            const float4* wk = wv + ((r * a.KW + sx) * a.C + c) * (MT / 4);
#pragma unroll
            for (int v = 0; v < MT / 4; ++v) {
              const float4 w4 = wk[v];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                acc[q][4 * v + 0] = fmaf(x[sx + q], w4.x, acc[q][4 * v + 0]);
                acc[q][4 * v + 1] = fmaf(x[sx + q], w4.y, acc[q][4 * v + 1]);
                acc[q][4 * v + 2] = fmaf(x[sx + q], w4.z, acc[q][4 * v + 2]);
                acc[q][4 * v + 3] = fmaf(x[sx + q], w4.w, acc[q][4 * v + 3]);
              }
            }
          }
        }
      }
    }
    const long long p0 = (n * a.Ho + ho) * (long long)a.Wo + wo0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (wo0 + q < a.Wo) {
        float* py = a.y + (p0 + q) * a.ldy;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          if (m < a.M) {
            float v = acc[q][m];
            if (a.bias) v += __ldg(a.bias + m);          // add_bias, convolution_op.rs:705
            if (a.chan_add) v += __ldg(a.chan_add + m);  // folded Add, add_op.rs:75
            if (a.relu) v = fmaxf(v, 0.f);               // relu_op.rs:31-33
            acc[q][m] = v;
          }
        }
        if (vec_out) {
#pragma unroll
          for (int v = 0; v < MT / 4; ++v)
            if (4 * v < a.M) *reinterpret_cast<float4*>(py + 4 * v) = make_float4(acc[q][4 * v], acc[q][4 * v + 1], acc[q][4 * v + 2], acc[q][4 * v + 3]);
        } else {
#pragma unroll
          for (int m = 0; m < MT; ++m)
            if (m < a.M) py[m] = acc[q][m];
        }
      }
    }
  }
}

static bool direct_eligible(const ConvArgs& a) {
  static const int off = [] { const char* e = getenv("B200_NO_DIRECT_CONV"); return e ? atoi(e) : 0; }();   // A/B timing only
  return !off && a.M <= 16 && a.K <= 128 && a.C <= 4;
}

int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
  const long long P = (long long)a.N * a.Ho * a.Wo;
  if (P == 0 || a.M == 0) return 0;
  if (direct_eligible(a)) {
    const int MT = a.M <= 8 ? 8 : 16;
    long long blocks = (P + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const size_t smem = (size_t)a.K * MT * sizeof(float);
    static const int no4 = [] { const char* e = getenv("B200_NO_DIRECT4"); return e ? atoi(e) : 0; }();   // A/B timing only
    if (!no4 && a.sw == 1 && a.KW <= 8) {
      const long long G = (long long)a.N * a.Ho * ((a.Wo + 3) / 4);
      long long b4 = (G + 255) / 256;
      if (b4 > 148 * 16) b4 = 148 * 16;
      if (MT == 8) conv_direct4_kernel<8><<<(int)b4, 256, smem, st>>>(a);
      else conv_direct4_kernel<16><<<(int)b4, 256, smem, st>>>(a);
    } else if (MT == 8) conv_direct_kernel<8><<<(int)blocks, 256, smem, st>>>(a);
    else conv_direct_kernel<16><<<(int)blocks, 256, smem, st>>>(a);
    B200_CUDA(cudaGetLastError());
    return 0;
  }
  const bool vec = a.C % 4 == 0 && a.wc == a.C && a.ldx % 4 == 0 && a.ldw % 4 == 0 && (((uintptr_t)a.x) & 15) == 0 &&
                   (((uintptr_t)a.w) & 15) == 0;
  if (a.M <= 16) return launch_bn<16>(a, vec, st);
  if (a.M <= 32) return launch_bn<32>(a, vec, st);
  if (a.M <= 64) return launch_bn<64>(a, vec, st);
  const int rem = a.M % 128;
  if (rem == 0 || rem > 64) return launch_bn<128>(a, vec, st);
  return launch_bn<64>(a, vec, st);
}

}  // namespace b200
