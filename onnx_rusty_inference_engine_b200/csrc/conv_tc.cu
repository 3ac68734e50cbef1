// tcgen05 3xTF32 implicit-GEMM convolution for sm_100a -- persistent, warp-specialised.
//
// Semantics: conv2d, convolution_op.rs:224-517 (+ folded Add add_op.rs:75 and Relu relu_op.rs:31-33), identical to
// conv_simt.cu.  fp32 accuracy is kept by splitting every operand x into hi = rna_tf32(x) and lo = rna_tf32(x - hi)
// and issuing three tensor-core products per k-step: hi*hi into a main accumulator, lo*hi + hi*lo into a separate
// correction accumulator (the tensor core truncates when it folds products into the fp32 accumulator; keeping the
// 2^-11-times smaller terms apart costs no extra MMAs and removes two thirds of the truncation steps from the main
// sum).  lo*lo is below fp32 rounding.  The two accumulators are added once, in the epilogue.
//
// GEMM view: D[P x M] = A[P x K] * W[M x K]^T, P = N*Ho*Wo output pixels, K = KH*KW*C ordered (r, s, c).
//   UMMA tile 128 pixels (TMEM lanes) x BN <= 128 output channels (TMEM columns), k-block = 32 floats = one 128-byte
//   swizzle row = 4 k-steps of 8, kind::tf32, cta_group::1, both operands K-major in shared memory.
//   TMEM: 2 accumulator stages x {main, correction} x BN columns (<= 512), so the epilogue of tile i overlaps the
//   main loop of tile i+1.
// One persistent CTA per SM (grid = min(tiles, SMs)), static round-robin tile schedule, 27 warps:
//   warps 0-7   epilogue (pairs w, w+4 share TMEM lanes 32*(w%4).. and split the channels): tcgen05.ld accumulator rows (warp w owns TMEM lanes 32w..), + bias (+ channel add), Relu,
//               transpose 32 rows x 32 channels through swizzled shared memory and write complete 128-byte row
//               segments at the (channel-offset) destination.
//   warp 8      allocates TMEM, initialises mbarriers, issues the TMA loads of the pre-split weight tiles.
//   warp 9      one thread issues tcgen05.mma / tcgen05.commit.
//   warp 10     proxy-fence relay (see below)
//   warps 11-26 A producers: gather im2col rows straight from the channels-last activation (any stride / padding /
//               tap; 16-byte chunks; 4 k-blocks of loads in flight per thread), split hi/lo in registers, store both
//               tiles in the 128B-swizzled K-major layout UMMA expects, fence.proxy.async, arrive.
// Weights are split, padded and given a TMA descriptor ONCE per model (tc_prepare_weights).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "internal.h"

namespace b200 {

struct TcWeights {
  float* buf = nullptr;     // [2*Mpad][Kpad]: hi rows then lo rows, zero padded
  CUtensorMap tmap;         // 2-D tiled map over buf, box = 32 floats x BN rows, SWIZZLE_128B
  int M = 0, K = 0, Mpad = 0, Kpad = 0, BN = 0;
  ~TcWeights() { if (buf) cudaFree(buf); }
};

namespace {

constexpr int BM = 128;                    // pixels per tile (UMMA M)
constexpr int BK = 32;                     // floats per k-block (128 bytes)
constexpr int A_TILE_BYTES = BM * BK * 4;  // 16 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int TMA_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int NUM_PROD_WARPS = 8;
constexpr int FENCE_WARP = 10;
constexpr int ATMA_WARP = 11;   // pointwise layers: issues the TMA loads of the raw A tiles
constexpr int PROD_WARP0 = 12;
constexpr int NTHREADS = (PROD_WARP0 + NUM_PROD_WARPS) * 32;  // 640
constexpr int ROWS_PER_THREAD = BM / (NUM_PROD_WARPS * 4);   // 2 rows per producer thread per k-block
constexpr int PREFETCH = 3;                // k-blocks of A loads in flight per producer thread
constexpr int EPI_SLAB_BYTES = 32 * 32 * 4;  // per epilogue warp pair: 32 rows x 32 channels
constexpr int EPI_STAGING_BYTES = 4 * EPI_SLAB_BYTES;  // one slab per warp pair: 16 KB
constexpr int SMEM_MAX = 227 * 1024;
constexpr int MAX_STAGES = 6;
constexpr int MAX_RAW = 8;                 // raw A ring (pointwise / TMA-fed mode)

struct TcParams {
  ConvArgs a;
  int BN;          // output channels per tile (multiple of 16, <= 128)
  int S;           // smem pipeline stages
  int nkb;         // k-blocks per tile
  int n_tiles_n;   // channel tiles
  int total_tiles; // pixel tiles x channel tiles
  int tmem_cols;   // power of two >= 4*BN
  int Mpad;        // weight rows per half (hi / lo)
  int P;           // output pixels (fits in int32, checked on the host)
  int vec_store;   // destination base and pitch are 16-byte aligned: 128-bit stores
  int a_tma;       // 1: A tiles arrive by TMA into a raw ring (pointwise layers); 0: register gather
  int R;           // raw ring slots (a_tma)
  int debug;       // B200_TC_DEBUG bit mask (timing experiments only): 1 skip weight TMA, 2 skip A loads, 4 skip stores, 8 skip proxy fence, 16 skip MMAs, 32 skip the gather producers' st.shared
  uint32_t magicC, magicKW, magicWo, magicHo;  // ceil(2^32 / d), 0 when d == 1: exact n / d for n, d < 2^16
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
// the same descriptor from a precomputed low word (address >> 4 | LBO); the high word is constant
__device__ __forceinline__ uint64_t sw128_desc(uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(0x40004040u));   // SBO 1024 >> 4, version 1, SWIZZLE_128B
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
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

// The producer's version of the same split, 4 integer/FP instructions per element instead of the ~10 that two
// cvt.rna.tf32 expand to (the PTX conversion is emulated with NaN/Inf handling on sm_100):
//   hi = (bits(x) + 0x1000) & 0xFFFFE000      round-to-nearest (ties away) to 10 mantissa bits, exact TF32 value
//   lo = bits(x - hi) + 0x1000                the tensor core ignores the low 13 bits of a TF32 operand, so adding
//                                             half an ulp before that truncation IS the rounding; no mask needed
__device__ __forceinline__ float split_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float split_lo(float x, float hi) { return __uint_as_float(__float_as_uint(x - hi) + 0x1000u); }

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapA, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const ConvArgs& a = p.a;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_tile_bytes = p.BN * BK * 4;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
  const uint32_t raw_ring = smem_base + (uint32_t)p.S * stage_bytes;             // R x 16 KB raw fp32 A tiles (a_tma)
  const uint32_t epi_staging = raw_ring + (uint32_t)p.R * A_TILE_BYTES;          // 2 KB-aligned slabs
  const uint32_t sbias = epi_staging + EPI_STAGING_BYTES;                        // Mpad floats: bias, zero padded
  const uint32_t sadd = sbias + 4u * (uint32_t)p.Mpad;                           // Mpad floats: folded channel add
  const uint32_t bars = sadd + 4u * (uint32_t)p.Mpad;                            // 8-byte mbarriers
  auto full_a = [&](int s) { return bars + 8u * s; };
  auto full_b = [&](int s) { return bars + 8u * (MAX_STAGES + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * MAX_STAGES + s); };
  auto ready_a = [&](int s) { return bars + 8u * (3 * MAX_STAGES + s); };
  auto tmem_full = [&](int s) { return bars + 8u * (4 * MAX_STAGES + s); };
  auto tmem_empty = [&](int s) { return bars + 8u * (4 * MAX_STAGES + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (4 * MAX_STAGES + 4);
  auto raw_full = [&](int r) { return bars + 8u * (4 * MAX_STAGES + 5 + r); };
  auto raw_empty = [&](int r) { return bars + 8u * (4 * MAX_STAGES + 5 + MAX_RAW + r); };
  auto a_hi = [&](int s) { return smem_base + (uint32_t)s * stage_bytes; };
  auto a_lo = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + A_TILE_BYTES; };
  auto b_hi = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + 2 * A_TILE_BYTES; };
  auto b_lo = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + 2 * A_TILE_BYTES + b_tile_bytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // per-channel epilogue constants -> shared memory once per CTA (the epilogue then needs no global loads)
  for (int m = threadIdx.x; m < p.Mpad; m += NTHREADS) {
    const float b = (a.bias && m < a.M) ? __ldg(a.bias + m) : 0.f;
    const float c = (a.chan_add && m < a.M) ? __ldg(a.chan_add + m) : 0.f;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + 4u * m), "f"(b) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sadd + 4u * m), "f"(c) : "memory");
  }

  if (warp == TMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < p.S; ++s) { mbar_init(full_a(s), NUM_PROD_WARPS); mbar_init(full_b(s), 1); mbar_init(empty(s), 1); mbar_init(ready_a(s), 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(tmem_full(s), 1); mbar_init(tmem_empty(s), NUM_EPI_WARPS); }
      for (int r = 0; r < p.R; ++r) { mbar_init(raw_full(r), 1); mbar_init(raw_empty(r), NUM_PROD_WARPS); }
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

  if (warp >= PROD_WARP0) {
    // ================================================================ A producers (16 warps)
    // Per tile each thread fixes, for its ROWS_PER_THREAD im2col rows, a base pointer (tap (0,0), channel 0) and two
    // separable validity masks (bit 8*i+r: input row h0+r inside the image; bit 8*i+s: column w0+s inside).
    // Per k-block the chunk's (r, s, c) is decoded once (multiply-high division) into one element offset shared by
    // all rows, so a load costs an add, a mask test and the LDG.
    const int pw = warp - PROD_WARP0;
    const int chunk = lane & 7;      // 16-byte chunk within the 128-byte k-block row
    const int rsub = lane >> 3;      // 4 rows per warp-wide access
    int my_tiles = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) ++my_tiles;
    const int items = my_tiles * p.nkb;

    if (p.a_tma) {
      // ---- TMA-fed mode (pointwise layers): the raw fp32 tile of each k-block is already in shared memory, in the
      // same 128B-swizzled layout as the hi / lo tiles, so a thread converts in place: same offsets in, same out.
      // Up to R raw tiles (R x 16 KB) are in flight from HBM without costing a register.
      uint32_t off[ROWS_PER_THREAD];
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        const int row = pw * (4 * ROWS_PER_THREAD) + i * 4 + rsub;
        off[i] = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
      }
      int s = 0, r = 0;
      uint32_t ph = 0, rph = 0;
      for (int idx = 0; idx < items; ++idx) {
        mbar_wait(raw_full(r), rph);
        const uint32_t raw = raw_ring + (uint32_t)r * A_TILE_BYTES;
        float4 x[ROWS_PER_THREAD];
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w) : "r"(raw + off[i]) : "memory");
        float4 hi[ROWS_PER_THREAD], lo[ROWS_PER_THREAD];
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i) {
          hi[i].x = split_hi(x[i].x); hi[i].y = split_hi(x[i].y); hi[i].z = split_hi(x[i].z); hi[i].w = split_hi(x[i].w);
          lo[i].x = split_lo(x[i].x, hi[i].x); lo[i].y = split_lo(x[i].y, hi[i].y);
          lo[i].z = split_lo(x[i].z, hi[i].z); lo[i].w = split_lo(x[i].w, hi[i].w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(raw_empty(r));   // the raw slot has been read (values are in registers)
        if (++r == p.R) { r = 0; rph ^= 1u; }
        mbar_wait(empty(s), ph ^ 1u);
        const uint32_t hi_base = a_hi(s), lo_base = a_lo(s);
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i) {
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hi_base + off[i]), "f"(hi[i].x), "f"(hi[i].y), "f"(hi[i].z), "f"(hi[i].w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_base + off[i]), "f"(lo[i].x), "f"(lo[i].y), "f"(lo[i].z), "f"(lo[i].w) : "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(full_a(s));
        if (++s == p.S) { s = 0; ph ^= 1u; }
      }
    } else {

    int l_tile = blockIdx.x, l_kb = 0;   // load cursor
    const float* base[ROWS_PER_THREAD];
    uint32_t hmask = 0, wmask = 0;
    auto set_tile = [&](int tile) {
      // first pixel of the tile -> (n, ho, wo) with two real divisions; the thread's rows follow with
      // multiply-high divisions of small numbers (row < 128, so the carries stay below 2^16)
      const int p0 = (tile / p.n_tiles_n) * BM;
      const int t0 = p0 / a.Wo, wo0 = p0 - t0 * a.Wo;
      const int n0 = t0 / a.Ho, ho0 = t0 - n0 * a.Ho;
      hmask = 0; wmask = 0;
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        const int row = pw * (4 * ROWS_PER_THREAD) + i * 4 + rsub;
        const bool ok = p0 + row < p.P;
        const int wsum = wo0 + row;
        const int cw = p.magicWo ? (int)__umulhi((unsigned)wsum, p.magicWo) : wsum;   // wsum / Wo
        const int wo = wsum - cw * a.Wo;
        const int hsum = ho0 + cw;
        const int ch = p.magicHo ? (int)__umulhi((unsigned)hsum, p.magicHo) : hsum;   // hsum / Ho
        const int ho = hsum - ch * a.Ho;
        const int n = n0 + ch;
        const int h0 = ho * a.sh - a.pt, w0 = wo * a.sw - a.pl;
        base[i] = a.x + ((long long)n * a.H * a.W + (long long)h0 * a.W + w0) * a.ldx;
        // taps r with 0 <= h0 + r < H form the bit range [max(0,-h0), min(KH, H-h0)); same for s
        const int rlo = max(0, -h0), rhi = min(a.KH, a.H - h0);
        const int slo = max(0, -w0), shi = min(a.KW, a.W - w0);
        const uint32_t hm = (ok && rhi > rlo) ? (((1u << rhi) - 1u) & ~((1u << rlo) - 1u)) : 0u;
        const uint32_t wm = (shi > slo) ? (((1u << shi) - 1u) & ~((1u << slo) - 1u)) : 0u;
        hmask |= hm << (8 * i);
        wmask |= wm << (8 * i);
      }
    };
    float4 v[PREFETCH][ROWS_PER_THREAD];
    auto issue = [&](float4 (&dst)[ROWS_PER_THREAD]) {
      const int k = l_kb * BK + chunk * 4;
      const int tap = p.magicC ? (int)__umulhi((unsigned)k, p.magicC) : k;      // k / C
      const int c = k - tap * a.C;
      const int r = p.magicKW ? (int)__umulhi((unsigned)tap, p.magicKW) : tap;  // tap / KW
      const int sx = tap - r * a.KW;
      const int delta = (r * a.W + sx) * a.ldx + c;
      const uint32_t m = (k < a.K) ? ((hmask >> r) & (wmask >> sx)) : 0u;   // bit 8*i: row i valid for this tap
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (((m >> (8 * i)) & 1u) && !(p.debug & 2)) dst[i] = __ldg(reinterpret_cast<const float4*>(base[i] + delta));
      }
      if (++l_kb == p.nkb) {
        l_kb = 0;
        l_tile += gridDim.x;
        if (l_tile < p.total_tiles) set_tile(l_tile);
      }
    };
    if (items > 0) set_tile(l_tile);
#pragma unroll
    for (int d = 0; d < PREFETCH; ++d)
      if (d < items) issue(v[d]);

    // smem offsets of this thread's rows inside a tile (fixed for the whole kernel)
    uint32_t soff[ROWS_PER_THREAD];
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
      const int row = pw * (4 * ROWS_PER_THREAD) + i * 4 + rsub;
      soff[i] = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
    }
    int s = 0;
    uint32_t ph = 0;
    for (int base_i = 0; base_i < items; base_i += PREFETCH) {
#pragma unroll
      for (int d = 0; d < PREFETCH; ++d) {
        const int idx = base_i + d;
        if (idx < items) {
          // split in registers first (independent of the stage), then wait for the stage and store
          float4 hi[ROWS_PER_THREAD], lo[ROWS_PER_THREAD];
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i) {
            const float4 x = v[d][i];
            hi[i].x = split_hi(x.x); hi[i].y = split_hi(x.y); hi[i].z = split_hi(x.z); hi[i].w = split_hi(x.w);
            lo[i].x = split_lo(x.x, hi[i].x); lo[i].y = split_lo(x.y, hi[i].y);
            lo[i].z = split_lo(x.z, hi[i].z); lo[i].w = split_lo(x.w, hi[i].w);
          }
          mbar_wait(empty(s), ph ^ 1u);
          const uint32_t hi_base = a_hi(s), lo_base = a_lo(s);
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i) {
            if (p.debug & 32) break;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hi_base + soff[i]), "f"(hi[i].x), "f"(hi[i].y), "f"(hi[i].z), "f"(hi[i].w) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_base + soff[i]), "f"(lo[i].x), "f"(lo[i].y), "f"(lo[i].z), "f"(lo[i].w) : "memory");
          }
          // No proxy fence here: a fence in this thread would also wait for the PREFETCH-1 k-blocks of global
          // loads it still has in flight and serialise the pipeline.  The stores are released by the mbarrier
          // arrive below; the MMA thread acquires them with its wait on full_a and issues the generic->async
          // proxy fence itself, immediately before the tcgen05.mma that reads this stage.
          __syncwarp();
          if (lane == 0) mbar_arrive(full_a(s));
          if (idx + PREFETCH < items) issue(v[d]);
          if (++s == p.S) { s = 0; ph ^= 1u; }
        }
      }
    }
    }  // register-gather path
  } else if (warp == TMA_WARP) {
    // ================================================================ weight tiles via TMA
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int m0 = (t % p.n_tiles_n) * p.BN;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(empty(s), ph ^ 1u);
          if (p.debug & 1) { mbar_arrive(full_b(s)); if (++s == p.S) { s = 0; ph ^= 1u; } continue; }
          mbar_expect_tx(full_b(s), 2u * (uint32_t)b_tile_bytes);
          tma_load_2d(b_hi(s), &tmapB, full_b(s), kb * BK, m0);
          tma_load_2d(b_lo(s), &tmapB, full_b(s), kb * BK, p.Mpad + m0);
          if (++s == p.S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == ATMA_WARP) {
    // ================================================================ raw A tiles via TMA (pointwise layers only)
    if (p.a_tma && lane == 0) {
      int r = 0;
      uint32_t rph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int p0 = (t / p.n_tiles_n) * BM;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(raw_empty(r), rph ^ 1u);
          mbar_expect_tx(raw_full(r), (uint32_t)A_TILE_BYTES);
          // box = 32 channels x 128 pixel rows; rows past P and channels past C are zero-filled by the tensor map
          tma_load_2d(raw_ring + (uint32_t)r * A_TILE_BYTES, &tmapA, raw_full(r), kb * BK, p0);
          if (++r == p.R) { r = 0; rph ^= 1u; }
        }
      }
    }
  } else if (warp == FENCE_WARP) {
    // ================================================================ proxy-fence relay
    // The A tiles are written with ordinary st.shared (generic proxy) and read by tcgen05.mma (async proxy), so a
    // fence.proxy.async has to sit between them.  It compiles to MEMBAR.ALL.CTA: in a producer thread it would wait
    // for that thread's prefetched global loads, in the MMA thread it would wait for the MMAs in flight -- either
    // way one fence per k-block serialises the pipeline (measured: ~1.8K clk per k-block for every layer).  This
    // thread has nothing outstanding: it acquires full_a, fences, and releases ready_a to the MMA thread.
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(full_a(s), ph);
          if (!(p.debug & 8)) fence_proxy_async();
          mbar_arrive(ready_a(s));
          if (++s == p.S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================================================================ MMA issuer
    // The whole warp walks the pipeline (uniform control flow, barrier waits by all lanes); one elected lane issues.
    // The issue stream is the kernel's critical resource (measured: the time per k-block did not depend on BN while
    // descriptors were rebuilt from addresses inside a single-lane branch), so the 64-bit shared-memory descriptors
    // are kept as 32-bit low words that advance by adds: +2 per k-step (32 bytes >> 4), +stage_bytes/16 per stage.
    const uint32_t idesc = instr_desc_tf32(p.BN), idesc2 = instr_desc_tf32(2 * p.BN);
    const bool leader = elect_one();
    const uint32_t lo_first = ((a_hi(0) >> 4) & 0x3FFFu) | (1u << 16);   // [0,14) address >> 4, [16,30) LBO = 1
    const uint32_t lo_step = (uint32_t)stage_bytes >> 4;
    const uint32_t lo_wrap = lo_first + (uint32_t)p.S * lo_step;
    uint32_t lo = lo_first;
    int s = 0;
    uint32_t ph = 0;
    int tc = 0;
    const int tail_ksteps = ((a.K - (p.nkb - 1) * BK + 7) >> 3);   // k-steps of the last k-block (1..4)
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tc) {
      const int as = tc & 1;
      const uint32_t aph = (uint32_t)(tc >> 1) & 1u;
      mbar_wait(tmem_empty(as), aph ^ 1u);   // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_main = tmem_base + (uint32_t)(as * 2 * p.BN);
      const uint32_t d_corr = d_main + (uint32_t)p.BN;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(ready_a(s), ph);  // A stage written, released, and proxy-fenced by the fence warp
        mbar_wait(full_b(s), ph);
        tc_fence_after();
        if (leader && !(p.debug & 16)) {
          const uint32_t ah = lo, al = lo + (A_TILE_BYTES >> 4), bh = lo + 2 * (A_TILE_BYTES >> 4);
          const int ksteps = (kb == p.nkb - 1) ? tail_ksteps : 4;
          // B_hi and B_lo are adjacent in the stage ([2*BN rows] x 128 B), and so are the two accumulators
          // ([main | correction] = 2*BN TMEM columns): ONE N = 2*BN instruction computes A_hi*B_hi -> main and
          // A_hi*B_lo -> correction, reading A_hi once; a second N = BN instruction adds A_lo*B_hi.
          umma_tf32(d_main, sw128_desc(ah), sw128_desc(bh), idesc2, kb != 0 ? 1u : 0u);
          umma_tf32(d_corr, sw128_desc(al), sw128_desc(bh), idesc, 1u);
#pragma unroll
          for (int kk = 1; kk < 4; ++kk) {
            if (kk < ksteps) {
              umma_tf32(d_main, sw128_desc(ah + 2 * kk), sw128_desc(bh + 2 * kk), idesc2, 1u);
              umma_tf32(d_corr, sw128_desc(al + 2 * kk), sw128_desc(bh + 2 * kk), idesc, 1u);
            }
          }
        }
        __syncwarp();
        if (leader) umma_commit(empty(s));   // frees the smem stage once the MMAs above have read it
        lo += lo_step;
        if (++s == p.S) { s = 0; ph ^= 1u; lo = lo_first; }
      }
      if (leader) umma_commit(tmem_full(as));   // accumulator stage complete -> epilogue
      __syncwarp();
    }
    (void)lo_wrap;
  } else {
    // ================================================================ epilogue (warps 0-7)
    // Warp w may only read TMEM lanes 32*(w % 4) .. +31, so warps w and w+4 form a pair on the same 32 accumulator
    // rows and split every 32-channel group: `half` 0 drains channels [j0, j0+16), `half` 1 drains [j0+16, j0+32).
    // Both write their 16 channels into the pair's shared 32 x 128-byte swizzled slab (row-per-thread, conflict-free),
    // meet on a 64-thread named barrier, and then each stores 16 of the 32 rows with 8 lanes per row, so that every
    // global store instruction writes four complete 128-byte row segments of the channels-last destination.
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t slab = epi_staging + (uint32_t)quarter * EPI_SLAB_BYTES;
    const uint32_t pair_bar = 1u + (uint32_t)quarter;   // named barriers 1..4 (0 is __syncthreads)
    const bool has_add = a.chan_add != nullptr, do_relu = a.relu != 0;
    int tc = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tc) {
      const int as = tc & 1;
      const uint32_t aph = (uint32_t)(tc >> 1) & 1u;
      const int p0 = (t / p.n_tiles_n) * BM;
      const int m0 = (t % p.n_tiles_n) * p.BN;
      mbar_wait(tmem_full(as), aph);
      tc_fence_after();
      const uint32_t t_main = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * 2 * p.BN);
      for (int j0 = 0; j0 < p.BN; j0 += 32) {
        const int width = min(32, p.BN - j0);   // 32, or 16 for the last group when BN % 32 == 16
        const int h = half * 16;
        if (h < width) {
          uint32_t acc[16], cor[16];
          tmem_ld16(t_main + (uint32_t)(j0 + h), acc);
          tmem_ld16(t_main + (uint32_t)(p.BN + j0 + h), cor);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 b4, c4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const uint32_t off = 4u * (uint32_t)(m0 + j0 + h + q * 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(sbias + off));
            if (has_add)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c4.x), "=f"(c4.y), "=f"(c4.z), "=f"(c4.w) : "r"(sadd + off));
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
            const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = q * 4 + e;
              float val = __uint_as_float(acc[j]) + __uint_as_float(cor[j]);   // hi*hi + (lo*hi + hi*lo)
              val = val + bb[e];                     // add_bias, convolution_op.rs:705 (0 when the node has no bias)
              if (has_add) val = val + cc[e];        // folded Add node, add_op.rs:75 (a second rounding, as upstream)
              if (do_relu) val = fmaxf(val, 0.f);    // relu_op.rs:31-33
              o[e] = val;
            }
            const int c = (h >> 2) + q;   // 16-byte chunk within the 128-byte slab row
            const uint32_t addr = slab + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // both halves of the slab are written
        const int c = lane & 7;
        const int m = m0 + j0 + c * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = half * 16 + i * 4 + (lane >> 3);
          const int prow = p0 + quarter * 32 + rr;
          if (c * 4 < width && prow < p.P && m < a.M && !(p.debug & 4)) {
            float4 val;
            const uint32_t addr = slab + (uint32_t)rr * 128u + (uint32_t)((c ^ (rr & 7)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w) : "r"(addr));
            float* dst = a.y + (long long)prow * a.ldy + m;
            if (p.vec_store && m + 4 <= a.M) {
              *reinterpret_cast<float4*>(dst) = val;
            } else {
              dst[0] = val.x;
              if (m + 1 < a.M) dst[1] = val.y;
              if (m + 2 < a.M) dst[2] = val.z;
              if (m + 3 < a.M) dst[3] = val.w;
            }
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // the slab is rewritten by the next channel group
      }
      // all tcgen05.ld of this accumulator stage have completed: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TMA_WARP) {
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

// BN: the smallest multiple of 16 that covers M in ceil(M / 128) channel tiles.
int pick_bn(int M) {
  static const int force = [] { const char* e = getenv("B200_TC_BN"); return e ? atoi(e) : 0; }();   // experiments only
  if (force >= 16 && force <= 128 && force % 16 == 0 && M > force) return force;
  const int nt = (M + 127) / 128;
  const int per = (M + nt - 1) / nt;
  return (per + 15) & ~15;
}

}  // namespace

int tc_supported(const ConvArgs& a) {
  if (a.C % 4 != 0 || a.ldx % 4 != 0 || a.wc != a.C) return B200_EUNSUPPORTED;
  if ((((uintptr_t)a.x) & 15) != 0) return B200_EUNSUPPORTED;
  if (a.M < 1 || a.K < 8) return B200_EUNSUPPORTED;
  if (a.KH > 8 || a.KW > 8 || a.pt > 15 || a.pl > 15 || ROWS_PER_THREAD > 4) return B200_EUNSUPPORTED;                  // 16-bit validity masks per row
  if (a.Ho + BM >= 65536 || a.Wo + BM >= 65536) return B200_EUNSUPPORTED;                       // multiply-high division range
  if (a.K >= 65536 || a.C >= 65536) return B200_EUNSUPPORTED;                                 // multiply-high division range
  if ((long long)(a.KH + 1) * a.W * a.ldx >= (1ll << 31)) return B200_EUNSUPPORTED;           // 32-bit tap offsets
  const long long P = (long long)a.N * a.Ho * a.Wo, in_pix = (long long)a.N * a.H * a.W;
  if (P >= (1ll << 31) - BM || in_pix >= (1ll << 31)) return B200_EUNSUPPORTED;  // 32-bit pixel indices in the producer
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
  p.P = (int)P;
  p.n_tiles_n = w.Mpad / w.BN;
  const long long tiles = ((P + BM - 1) / BM) * p.n_tiles_n;
  if (tiles >= (1ll << 31)) B200_FAIL(B200_EUNSUPPORTED, "too many tiles");
  p.total_tiles = (int)tiles;
  p.tmem_cols = 32;
  while (p.tmem_cols < 4 * p.BN) p.tmem_cols <<= 1;   // 2 stages x (main + correction)
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * p.BN * BK * 4;
  const int fixed = 1024 + EPI_STAGING_BYTES + 8 * p.Mpad + 8 * (4 * MAX_STAGES + 5 + 2 * MAX_RAW);
  // Pointwise layers (1x1, stride 1, no padding): im2col row p IS input pixel p, so the A operand is a plain 2-D
  // matrix [P][C] and TMA can stream it; these layers are HBM-bound and want many bytes in flight.
  static const int no_atma = [] { const char* e = getenv("B200_TC_NO_ATMA"); return e ? atoi(e) : 0; }();
  p.a_tma = (!no_atma && a.KH == 1 && a.KW == 1 && a.sh == 1 && a.sw == 1 && a.pt == 0 && a.pl == 0 && a.H == a.Ho && a.W == a.Wo) ? 1 : 0;
  p.R = 0;
  int S = (SMEM_MAX - fixed) / stage_bytes;
  if (S > MAX_STAGES) S = MAX_STAGES;
  if (p.a_tma) {
    if (S > 3) S = 3;
    if (S > 2 && stage_bytes >= 64 * 1024) S = 2;
    int R = (SMEM_MAX - fixed - S * stage_bytes) / A_TILE_BYTES;
    if (R > MAX_RAW) R = MAX_RAW;
    if (S < 2 || R < 2) { p.a_tma = 0; S = (SMEM_MAX - fixed) / stage_bytes; if (S > MAX_STAGES) S = MAX_STAGES; }
    else p.R = R;
  }
  if (S < 2) B200_FAIL(B200_EUNSUPPORTED, "tcgen05 conv: not enough shared memory for 2 stages (BN=%d)", p.BN);
  p.S = S;
  const size_t smem = (size_t)S * stage_bytes + (size_t)p.R * A_TILE_BYTES + fixed;

  p.vec_store = (a.ldy % 4 == 0 && (((uintptr_t)a.y) & 15) == 0) ? 1 : 0;
  { static const int dbg = [] { const char* e = getenv("B200_TC_DEBUG"); return e ? atoi(e) : 0; }(); p.debug = dbg; }
  p.magicC = a.C == 1 ? 0u : (uint32_t)(((1ull << 32) + a.C - 1) / a.C);
  p.magicKW = a.KW == 1 ? 0u : (uint32_t)(((1ull << 32) + a.KW - 1) / a.KW);
  p.magicWo = a.Wo == 1 ? 0u : (uint32_t)(((1ull << 32) + a.Wo - 1) / a.Wo);
  p.magicHo = a.Ho == 1 ? 0u : (uint32_t)(((1ull << 32) + a.Ho - 1) / a.Ho);

  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  static int sm_count[64] = {0};
  if (dev < 64 && !attr_set[dev]) {
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    attr_set[dev] = true;
  }
  const int sms = (dev < 64 && sm_count[dev] > 0) ? sm_count[dev] : 148;
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  CUtensorMap tmapA;
  memset(&tmapA, 0, sizeof(tmapA));
  if (p.a_tma) {
    EncodeTiledFn enc = get_encode_fn();
    cuuint64_t gdim[2] = {(cuuint64_t)a.C, (cuuint64_t)P};
    cuuint64_t gstride[1] = {(cuuint64_t)a.ldx * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc ? enc(&tmapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                     : CUDA_ERROR_NOT_SUPPORTED;
    if (r != CUDA_SUCCESS) B200_FAIL(B200_ECUDA, "cuTensorMapEncodeTiled (activation map) failed with %d (C=%d P=%lld ldx=%d)", (int)r, a.C, P, a.ldx);
  }
  conv_tc_kernel<<<grid, NTHREADS, smem, st>>>(w.tmap, tmapA, p);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
