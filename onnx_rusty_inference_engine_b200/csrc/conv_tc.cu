// tcgen05 3xTF32 implicit-GEMM convolution for sm_100a.
//
// Semantics: conv2d, convolution_op.rs:224-517 (+ folded Add add_op.rs:75 and Relu relu_op.rs:31-33), identical to
// conv_simt.cu; fp32 accuracy is kept by splitting every operand x into hi = tf32(x) (top 19 bits) and
// lo = x - hi and issuing three tensor-core products per k-step, lo*hi + hi*lo + hi*hi, into one fp32 TMEM
// accumulator (the lo*lo term is below fp32 rounding).
//
// GEMM view: D[P x M] = A[P x K] * W[M x K]^T, P = N*Ho*Wo output pixels, K = KH*KW*C ordered (r, s, c).
//   UMMA tile: 128 pixels (TMEM lanes) x BN output channels (TMEM columns, two accumulators: main and correction), k-block = 32 floats = one 128-byte
//   swizzle row, 4 k-steps of 8 per block, kind::tf32, cta_group::1, both operands K-major in shared memory.
// Warp roles (192 threads):
//   warps 0-3  A producers: gather the im2col rows straight from the channels-last activation (any stride /
//              padding / tap, 16-byte chunks), split hi/lo in registers, store both tiles in the 128B-swizzled
//              K-major layout UMMA expects, fence.proxy.async, arrive.  After the main loop the same warps run the
//              epilogue: tcgen05.ld the accumulator rows, + bias (+ channel add), Relu, store at the
//              (channel-offset) destination.
//   warp 4     allocates TMEM, initialises the mbarriers and issues the TMA loads of the pre-split weight tiles
//              (cp.async.bulk.tensor.2d, SWIZZLE_128B) -- weights are split and padded ONCE per model.
//   warp 5     one thread issues tcgen05.mma and tcgen05.commit (frees the stage / signals the epilogue).
// Pipeline: S stages of {A_hi, A_lo, B_hi, B_lo}; full_a / full_b / empty mbarriers per stage.
#include <cuda.h>

#include "internal.h"

namespace b200 {

struct TcWeights {
  float* buf = nullptr;     // [2*Mpad][Kpad]: hi rows then lo rows, zero padded
  CUtensorMap tmap;         // 2-D tiled map over buf, box = 32 floats x BN rows, SWIZZLE_128B
  int M = 0, K = 0, Mpad = 0, Kpad = 0, BN = 0;
  ~TcWeights() { if (buf) cudaFree(buf); }
};

namespace {

constexpr int BM = 128;          // pixels per tile (UMMA M)
constexpr int BK = 32;           // floats per k-block (128 bytes)
constexpr int A_TILE_BYTES = BM * BK * 4;  // 16 KB
constexpr int NTHREADS = 192;
constexpr int SMEM_BUDGET_2CTA = 112 * 1024;
constexpr int SMEM_MAX = 227 * 1024;

struct TcParams {
  ConvArgs a;
  int BN;        // output channels per tile (multiple of 16, <= 256)
  int S;         // pipeline stages
  int nkb;       // k-blocks
  int tmem_cols; // power of two >= max(32, 2*BN): main accumulator + correction accumulator
  int Mpad;      // weight rows per half (hi / lo)
  int vec_store; // destination allows 16-byte stores
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, single CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (= 1, unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups)   [46,48) version = 1   [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1, a/b format TF32 [7,10)=[10,13)=2, K-major both, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t instr_desc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Round-to-nearest TF32 (low 13 mantissa bits cleared).  hi = rna(x), lo = rna(x - hi): with a rounded hi the
// remainder has at most 12 significant bits, so the second rounding loses at most one bit and is unbiased.
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NTHREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap tmapB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const ConvArgs& a = p.a;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_tile_bytes = p.BN * BK * 4;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
  const uint32_t bars = smem_base + (uint32_t)p.S * stage_bytes;      // 8-byte mbarriers
  auto full_a = [&](int s) { return bars + 8u * s; };
  auto full_b = [&](int s) { return bars + 8u * (p.S + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * p.S + s); };
  const uint32_t mma_done = bars + 8u * (3 * p.S);
  const uint32_t tmem_slot = mma_done + 8u;
  auto a_hi = [&](int s) { return smem_base + (uint32_t)s * stage_bytes; };
  auto a_lo = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + A_TILE_BYTES; };
  auto b_hi = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + 2 * A_TILE_BYTES; };
  auto b_lo = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + 2 * A_TILE_BYTES + b_tile_bytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long P = (long long)a.N * a.Ho * a.Wo;
  const long long p0 = (long long)blockIdx.x * BM;
  const int m0 = blockIdx.y * p.BN;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < p.S; ++s) { mbar_init(full_a(s), 4); mbar_init(full_b(s), 1); mbar_init(empty(s), 1); }
      mbar_init(mma_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    // ================================================================ A producers
    const int chunk = lane & 7;      // 16-byte chunk within the 128-byte k-block row
    const int rsub = lane >> 3;      // 4 rows per warp-wide access
    long long pix_base[8];
    int hw0[8];                      // (h0 << 16) | (w0 & 0xffff), both may be negative (padding)
    uint32_t valid = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = warp * 32 + i * 4 + rsub;
      const long long pp = p0 + row;
      const bool ok = pp < P;
      const long long q = ok ? pp : 0;
      const int wo = (int)(q % a.Wo);
      const long long t = q / a.Wo;
      const int ho = (int)(t % a.Ho);
      const long long n = t / a.Ho;
      pix_base[i] = n * a.H * a.W;
      const int h0 = ho * a.sh - a.pt, w0 = wo * a.sw - a.pl;
      hw0[i] = (h0 << 16) | (w0 & 0xFFFF);
      valid |= (ok ? 1u : 0u) << i;
    }
    for (int kb = 0; kb < p.nkb; ++kb) {
      const int s = kb % p.S;
      const uint32_t ph = (uint32_t)(kb / p.S) & 1u;
      const int k = kb * BK + chunk * 4;
      const int tap = k / a.C;
      const int c = k - tap * a.C;
      const int r = tap / a.KW, sx = tap - r * a.KW;
      const bool kvalid = k < a.K;
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int h = (hw0[i] >> 16) + r;
        const int w = (int)(short)(hw0[i] & 0xFFFF) + sx;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kvalid && ((valid >> i) & 1u) && h >= 0 && h < a.H && w >= 0 && w < a.W)
          v[i] = __ldg(reinterpret_cast<const float4*>(a.x + (pix_base[i] + (long long)h * a.W + w) * a.ldx + c));
      }
      mbar_wait(empty(s), ph ^ 1u);
      const uint32_t hi_base = a_hi(s), lo_base = a_lo(s);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = warp * 32 + i * 4 + rsub;
        const uint32_t off = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
        float4 hi, lo;
        hi.x = tf32_rna(v[i].x); hi.y = tf32_rna(v[i].y); hi.z = tf32_rna(v[i].z); hi.w = tf32_rna(v[i].w);
        lo.x = tf32_rna(v[i].x - hi.x); lo.y = tf32_rna(v[i].y - hi.y); lo.z = tf32_rna(v[i].z - hi.z); lo.w = tf32_rna(v[i].w - hi.w);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hi_base + off), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_base + off), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
      }
      fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_a(s));
    }

    // ================================================================ epilogue (same 4 warps: TMEM lanes 32*warp ..)
    mbar_wait(mma_done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const long long pp = p0 + row;
    float* yrow = a.y + pp * a.ldy;
    for (int j0 = 0; j0 < p.BN; j0 += 16) {
      uint32_t acc[16], cor[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)j0, acc);
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(p.BN + j0), cor);
      tmem_ld_wait();
      if (pp < P) {
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = m0 + j0 + j;
          float val = __uint_as_float(acc[j]) + __uint_as_float(cor[j]);  // hi*hi + (lo*hi + hi*lo)
          if (m < a.M) {
            if (a.bias) val = val + __ldg(a.bias + m);          // add_bias, convolution_op.rs:705
            if (a.chan_add) val = val + __ldg(a.chan_add + m);  // folded Add node, add_op.rs:75
            if (a.relu) val = fmaxf(val, 0.f);
          }
          o[j] = val;
        }
        if (p.vec_store && m0 + j0 + 16 <= a.M) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(yrow + m0 + j0 + q * 4) = make_float4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (m0 + j0 + j < a.M) yrow[m0 + j0 + j] = o[j];
        }
      }
    }
  } else if (warp == 4) {
    // ================================================================ weight tiles via TMA
    if (lane == 0) {
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int s = kb % p.S;
        const uint32_t ph = (uint32_t)(kb / p.S) & 1u;
        mbar_wait(empty(s), ph ^ 1u);
        mbar_expect_tx(full_b(s), 2u * (uint32_t)b_tile_bytes);
        tma_load_2d(b_hi(s), &tmapB, full_b(s), kb * BK, m0);
        tma_load_2d(b_lo(s), &tmapB, full_b(s), kb * BK, p.Mpad + m0);
      }
    }
  } else {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = instr_desc_tf32(p.BN);
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int s = kb % p.S;
        const uint32_t ph = (uint32_t)(kb / p.S) & 1u;
        mbar_wait(full_a(s), ph);
        mbar_wait(full_b(s), ph);
        tc_fence_after();
        const int ksteps = min(4, (a.K - kb * BK + 7) >> 3);
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t ah = smem_desc_sw128(a_hi(s) + kk * 32);
          const uint64_t al = smem_desc_sw128(a_lo(s) + kk * 32);
          const uint64_t bh = smem_desc_sw128(b_hi(s) + kk * 32);
          const uint64_t bl = smem_desc_sw128(b_lo(s) + kk * 32);
          // The tensor core truncates when it folds products into the fp32 accumulator, so the main term and the
          // 2^-11-times smaller correction terms get separate accumulators (columns [0,BN) and [BN,2BN)) and are
          // added once, in the epilogue: the correction sum then loses nothing and the main sum sees 1/3 of the
          // accumulation steps.
          const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
          umma_tf32(tmem_base + (uint32_t)p.BN, al, bh, idesc, acc);
          umma_tf32(tmem_base + (uint32_t)p.BN, ah, bl, idesc, 1u);
          umma_tf32(tmem_base, ah, bh, idesc, acc);
        }
        umma_commit(empty(s));   // frees the stage when the MMAs above have read it
      }
      umma_commit(mma_done);     // accumulator complete -> epilogue
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Weight preparation: w [M][ldw] (K valid floats per row) -> out [2*Mpad][Kpad]: tf32 hi rows, then lo rows.
__global__ void tc_split_weights_kernel(const float* __restrict__ w, int M, int K, int ldw, float* __restrict__ out, int Mpad, int Kpad) {
  const long long total = (long long)Mpad * Kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / Kpad), k = (int)(i - (long long)m * Kpad);
    float x = 0.f;
    if (m < M && k < K) x = w[(long long)m * ldw + k];
    const float hi = tf32_rna(x);
    out[i] = hi;
    out[total + i] = tf32_rna(x - hi);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

int pick_bn(int M) {
  int bn = (M + 15) & ~15;
  return bn > 256 ? 256 : bn;
}

}  // namespace

int tc_supported(const ConvArgs& a) {
  if (a.C % 4 != 0 || a.ldx % 4 != 0 || a.wc != a.C) return B200_EUNSUPPORTED;
  if ((((uintptr_t)a.x) & 15) != 0) return B200_EUNSUPPORTED;
  if (a.M < 1 || a.K < 8) return B200_EUNSUPPORTED;
  if (a.H >= 32768 || a.W >= 32768) return B200_EUNSUPPORTED;  // (h0, w0) are packed in 16 bits each
  return 0;
}

int tc_prepare_weights(const float* w_dev, int M, int K, cudaStream_t st, std::shared_ptr<TcWeights>* out) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) B200_FAIL(B200_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  std::shared_ptr<TcWeights> t(new TcWeights());
  t->M = M; t->K = K;
  t->BN = pick_bn(M);
  t->Mpad = (M + t->BN - 1) / t->BN * t->BN;
  t->Kpad = (K + BK - 1) / BK * BK;
  const size_t bytes = (size_t)2 * t->Mpad * t->Kpad * sizeof(float);
  if (cudaMalloc((void**)&t->buf, bytes) != cudaSuccess) { cudaGetLastError(); B200_FAIL(B200_ENOMEM, "cudaMalloc(%zu) for split weights", bytes); }
  const long long total = (long long)t->Mpad * t->Kpad;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  tc_split_weights_kernel<<<blocks, 256, 0, st>>>(w_dev, M, K, K, t->buf, t->Mpad, t->Kpad);
  B200_CUDA(cudaGetLastError());
  cuuint64_t gdim[2] = {(cuuint64_t)t->Kpad, (cuuint64_t)(2 * t->Mpad)};
  cuuint64_t gstride[1] = {(cuuint64_t)t->Kpad * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)t->BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&t->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, t->buf, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) B200_FAIL(B200_ECUDA, "cuTensorMapEncodeTiled failed with %d (M=%d K=%d BN=%d)", (int)r, M, K, t->BN);
  *out = t;
  return 0;
}

int launch_conv_tc(const ConvArgs& a, const TcWeights& w, cudaStream_t st) {
  const long long P = (long long)a.N * a.Ho * a.Wo;
  if (P == 0 || a.M == 0) return 0;
  if (w.M != a.M || w.K != a.K) B200_FAIL(B200_EINVAL, "tcgen05 weights prepared for M=%d K=%d, launch has M=%d K=%d", w.M, w.K, a.M, a.K);
  TcParams p;
  p.a = a;
  p.BN = w.BN;
  p.Mpad = w.Mpad;
  p.nkb = w.Kpad / BK;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.BN) p.tmem_cols <<= 1;   // main + correction accumulators
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * p.BN * BK * 4;
  int S = SMEM_BUDGET_2CTA / stage_bytes;
  if (S < 2) S = 2;
  if (S > 4) S = 4;
  if (S > p.nkb) S = p.nkb;
  p.S = S;
  p.vec_store = (a.ldy % 4 == 0 && (((uintptr_t)a.y) & 15) == 0) ? 1 : 0;
  const size_t smem = (size_t)S * stage_bytes + 1024 + 8 * (3 * S + 2);
  if (smem > (size_t)SMEM_MAX) B200_FAIL(B200_EUNSUPPORTED, "tcgen05 conv needs %zu bytes of shared memory", smem);
  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    attr_set[dev] = true;
  }
  dim3 grid((unsigned)((P + BM - 1) / BM), (unsigned)(w.Mpad / w.BN));
  conv_tc_kernel<<<grid, NTHREADS, smem, st>>>(w.tmap, p);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
