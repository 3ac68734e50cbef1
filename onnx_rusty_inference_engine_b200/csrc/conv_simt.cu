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

int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
  const long long P = (long long)a.N * a.Ho * a.Wo;
  if (P == 0 || a.M == 0) return 0;
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
