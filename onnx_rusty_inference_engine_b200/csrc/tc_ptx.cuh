// PTX wrappers shared by the tcgen05 kernels of this library (conv_tc.cu, mnist8_fused.cu): mbarriers, TMA, tensor-memory
// allocation / load / store, tcgen05.mma (kind::tf32, A operand in tensor memory), descriptors, the hi / lo split.
// Device code only; include inside namespace b200 { namespace { ... } }.
#pragma once
#include <cuda.h>
#include <cstdint>

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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32, single CTA.  A: lanes = rows, one 32-bit column per k element.
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 TMEM lanes x 16 columns from 8 registers per thread: thread t writes rows t/4 (r0 r1 | r4 r5) and t/4 + 8
// (r2 r3 | r6 r7), columns 2*(t%4) + {0, 1} and 8 + 2*(t%4) + {0, 1}  (layout measured with tools/exp/tmem_layout.cu)
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, float r0, float r1, float r2, float r3, float r4, float r5, float r6, float r7) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(r0), "f"(r1),
               "f"(r2), "f"(r3), "f"(r4), "f"(r5), "f"(r6), "f"(r7)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// The same, naming the registers the pending tcgen05.ld write: tcgen05.ld is asynchronous, and only a data dependency
// keeps the compiler from scheduling arithmetic on those registers above the wait (a "memory" clobber orders memory
// operations, not register uses).
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&a)[16], uint32_t (&b)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]),
                 "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]),
                 "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]),
                 "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
               :
               : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (= 1, unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups)   [46,48) version = 1   [61,64) layout = 2 (SW128)
// built from a precomputed low word (address >> 4 | LBO); the high word is constant
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
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// Round-to-nearest TF32 (low 13 mantissa bits cleared).  hi = rna(x), lo = rna(x - hi): with a rounded hi the
// remainder has at most 12 significant bits, so the second rounding loses at most one bit and is unbiased.
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// The producer's split, 2 instructions per element (the producers are issue-bound on the gather layers: in the conv1
// capture they were busy 80 % of the time and the split was a third of their k-block loop):
//   hi = bits(x) & 0xFFFFE000      truncation to 10 mantissa bits, an exact TF32 value
//   lo = x - hi                    exact in fp32 (13 significant bits, sign of x); the tensor core reads its top 10
//                                  mantissa bits
// x = hi + lo holds exactly whatever the rounding of hi, so the only cost against the round-to-nearest split used for
// the weights is the representation error of lo: <= 2^-21 |x| instead of 2^-23 |x|, i.e. a relative 2.4e-7 on a
// product (mean 1.2e-7, the same sign on every term: a scale factor on the output, four hundred times below the 1e-4
// relative tolerance) -- measured with tools/exp/merged_error.py: rms error against fp64 unchanged to two digits.
__device__ __forceinline__ float split_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float split_lo(float x, float hi) { return x - hi; }

// A producer's four 16-byte chunks (rows lane/4 + 8i of its 32-row quarter) are split into hi / lo and written to the
// warp's part of an A stage in tensor memory: hi -> columns [0,32), lo -> [32,64) of the stage; t_stage = lane
// 32*quarter, column of the warp's 16-column half.  Two 16-lane halves: rows i = 2j, 2j+1 (split per half, so only 16
// temporaries are live).
// B200_SPLIT_RAW_HI (experiment, default 0).  The tensor core reads only the top 10 mantissa bits of a kind::tf32 operand,
// i.e. it TRUNCATES, so the raw fp32 value can serve as the hi operand (parity tests green with it) and only
// lo = x - trunc(x) needs computing: one LOP3 per element for -trunc(x) and one packed FADD2 per two elements, 1.5
// instead of 2 issue slots per element.  Measured in one A/B run (profiles/README.md, round 2): slower, not faster --
// conv1 0.549 -> 0.557 ms, fire6 expand3x3 0.129 -> 0.133, step 83.4k -> 82.1k img/s, MNIST head likewise: the packed
// FADD2 needs aligned register pairs and ptxas answers with moves and spills under the producers' 72-register cap.
#ifndef B200_SPLIT_RAW_HI
#define B200_SPLIT_RAW_HI 0
#endif
__device__ __forceinline__ float neg_trunc_tf32(float x) { return __uint_as_float((__float_as_uint(x) & 0xFFFFE000u) ^ 0x80000000u); }
__device__ __forceinline__ void split_store(uint32_t t_stage, const float4 (&x)[4], uint32_t lo_cols = 32u) {   // lo half lo_cols columns after hi
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float4 a = x[2 * j], b = x[2 * j + 1];
    const uint32_t ta = t_stage + ((uint32_t)(j * 16) << 16);
#if B200_SPLIT_RAW_HI
    tmem_st_16x256b_x2(ta, a.x, a.y, b.x, b.y, a.z, a.w, b.z, b.w);
    const float2 a01 = __fadd2_rn(make_float2(a.x, a.y), make_float2(neg_trunc_tf32(a.x), neg_trunc_tf32(a.y)));
    const float2 a23 = __fadd2_rn(make_float2(a.z, a.w), make_float2(neg_trunc_tf32(a.z), neg_trunc_tf32(a.w)));
    const float2 b01 = __fadd2_rn(make_float2(b.x, b.y), make_float2(neg_trunc_tf32(b.x), neg_trunc_tf32(b.y)));
    const float2 b23 = __fadd2_rn(make_float2(b.z, b.w), make_float2(neg_trunc_tf32(b.z), neg_trunc_tf32(b.w)));
    tmem_st_16x256b_x2(ta + lo_cols, a01.x, a01.y, b01.x, b01.y, a23.x, a23.y, b23.x, b23.y);
#else
    float4 ah, bh, al, bl;
    ah.x = split_hi(a.x); ah.y = split_hi(a.y); ah.z = split_hi(a.z); ah.w = split_hi(a.w);
    bh.x = split_hi(b.x); bh.y = split_hi(b.y); bh.z = split_hi(b.z); bh.w = split_hi(b.w);
    tmem_st_16x256b_x2(ta, ah.x, ah.y, bh.x, bh.y, ah.z, ah.w, bh.z, bh.w);
    al.x = split_lo(a.x, ah.x); al.y = split_lo(a.y, ah.y); al.z = split_lo(a.z, ah.z); al.w = split_lo(a.w, ah.w);
    bl.x = split_lo(b.x, bh.x); bl.y = split_lo(b.y, bh.y); bl.z = split_lo(b.z, bh.z); bl.w = split_lo(b.w, bh.w);
    tmem_st_16x256b_x2(ta + lo_cols, al.x, al.y, bl.x, bl.y, al.z, al.w, bl.z, bl.w);
#endif
  }
}
