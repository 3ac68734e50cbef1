// tcgen05 3xTF32 implicit-GEMM convolution for sm_100a -- persistent, warp-specialised, A operand in tensor memory.
//
// Semantics: conv2d, convolution_op.rs:224-517 (+ folded Add add_op.rs:75 and Relu relu_op.rs:31-33), identical to
// conv_simt.cu.  fp32 accuracy is kept by splitting every operand x into hi + lo (weights: hi = rna_tf32(x),
// lo = rna_tf32(x - hi), once per model; activations: hi = x truncated to TF32, lo = x - hi, 2 instructions per element
// in the producers) and issuing three tensor-core products per k-step: hi*hi into a main accumulator, lo*hi + hi*lo
// into a separate correction accumulator (the tensor core truncates when it folds products into the fp32 accumulator;
// keeping the 2^-11-times smaller terms apart costs no extra MMAs and removes two thirds of the truncation steps from
// the main sum).  lo*lo is below fp32 rounding.  The two accumulators are added once, in the epilogue.  (Round 1 ran short
// reductions on wide tiles with ONE merged accumulator so that two stages fit; retired: see MERGED_MAX_K.)
//
// GEMM view: D[P x M] = A[P x K] * W[M x K]^T, P = N*Ho*Wo output pixels, K = KH*KW*C ordered (r, s, c).
//   UMMA tile 128 pixels (TMEM lanes) x BN <= 128 output channels (TMEM columns), k-block = 32 floats = 4 k-steps of
//   8, kind::tf32, cta_group::1.
//   A (activations): gathered to REGISTERS, split there, and written straight into TENSOR MEMORY with tcgen05.st
//     (16x256b: a thread owns 16-byte chunks of rows t/4 + 8i) -- the MMAs take A from TMEM, so the activations never
//     touch shared memory: no st.shared, no A operand reads on the shared-memory port, no generic->async proxy fence.
//   B (weights): pre-split (hi, lo) K-major tiles streamed by TMA into a SWIZZLE_128B ring.  The 16x256b store shape
//     puts float 4c+e of a 16-float group in TMEM column 2c+e (e < 2) or 8+2c+e-2 (e >= 2), so the weight
//     preparation applies the same permutation to K inside every group of 16 (a contraction does not care).
//   TMEM columns (512): [accumulator stages: {main | correction} x BN each][A stages: {hi 32 | lo 32} each].
//     BN <= 64: two accumulator stages (the epilogue of tile i overlaps the main loop of tile i+1) + four A stages.
//     64 < BN <= 96, gather, <= 8 k-blocks (HALFA, conv1): two accumulator stages + FOUR HALF-STAGES {hi 16 | lo 16},
//     one per (producer set, k half), released after two k-steps each.
//     Otherwise one accumulator stage + four A stages; the epilogue drains TMEM to a shared-memory slab first and
//     releases the accumulator before it touches global memory.
// One persistent CTA per SM (grid = min(tiles, SMs)), static round-robin tile schedule, 28 warps in one of two role
// layouts (struct Roles); the default one:
//   warps 0-7   epilogue: warps w, w+4 share TMEM lanes 32*(w%4).. and take the two halves of the tile's channels.
//               Phase 1 drains: tcgen05.ld main + correction, ONE add, row-per-thread into a private swizzled slab --
//               nothing else, it sits between two tiles' MMAs -- then the accumulator stage is handed back.  Phase 2
//               reads the slab back in 16-byte chunks, consecutive lanes along a row, applies bias (+ channel add) and
//               Relu, and stores complete 32-byte sectors of the channels-last rows.
//   warp 8      allocates TMEM, initialises mbarriers, issues the TMA loads of the weight tiles.
//   warp 9      MMA issuer (uniform control flow, one elected lane, descriptors advanced by adds).
//   warp 10     pointwise layers: TMA loads of raw fp32 A tiles into a shared-memory ring (POOL: of the 2R+1 input rows
//               the tile's 3x3 / 2 max-pool windows touch, as 4-D boxes of the tensor before the pool).
//   warps 10-11 gather layers: per-tile row decode ({offset, validity masks} of the tile's 128 im2col rows) into
//               shared-memory tables, one tile ahead of the producer set each of them serves.
//   warps 12-27 A producers, two sets of 8 that take alternate k-blocks: warp w owns TMEM lanes 32*(w%4).. (hardware
//               rule) and the 64-byte half ((w-12)/4)%2 of each k-block row.  Gather mode: im2col rows straight from
//               the channels-last activation (any stride / padding / tap; (r, s, offset) of a chunk from a per-CTA
//               table; 2 k-blocks of loads in flight per thread).  Pointwise mode: read the TMA-fed raw tile (POOL:
//               the maximum of the nine window chunks).
// The second layout (pointwise layers with BN > 64): 16 epilogue warps, 8 converter warps.  The 896 threads start at 72
// registers; setmaxnreg then gives the TMA/MMA warpgroup 40, the epilogue warpgroups 88 (80) and the converters 64.
// Weights are split, permuted, padded and given a TMA descriptor ONCE per model (tc_prepare_weights).
// Launches are programmatic dependent launches when the planner asks (ConvArgs::pdl): griddepcontrol.launch_dependents at
// entry, griddepcontrol.wait between the prologue and the first access to activations.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "tc_common.h"

namespace b200 {


namespace {

constexpr int BM = 128;                    // pixels per tile (UMMA M)
constexpr int BK = 32;                     // floats per k-block (128 bytes)
constexpr int A_TILE_BYTES = BM * BK * 4;  // 16 KB (raw fp32 A tile of the pointwise mode)
// Warp roles, two layouts of the same 28 warps (EPI16 = template parameter of the kernel):
//   EPI16 = false: 8 epilogue warps | TMA, MMA, A-TMA, idle | 16 A-producer warps (two sets of 8 that take alternate
//                  k-blocks) -- layers that gather, and pointwise layers with a long reduction
//   EPI16 = true : 16 epilogue warps | TMA, MMA, A-TMA, idle | 8 A-converter warps (every k-block) -- pointwise
//                  layers with BN > 64 (expand1x1, conv10), whose epilogue is exposed or is the bound
template <bool EPI16>
struct Roles {
  static constexpr int NEPI = EPI16 ? 16 : 8;       // epilogue warps (NEPI / 4 per TMEM lane quarter)
  static constexpr int TMA_WARP = NEPI;
  static constexpr int MMA_WARP = NEPI + 1;
  static constexpr int ATMA_WARP = NEPI + 2;         // pointwise layers: issues the TMA loads of the raw A tiles
  static constexpr int PROD_WARP0 = NEPI + 4;        // a multiple of 4: producer warp w writes TMEM lanes 32*(w%4)..
  static constexpr int NSETS = EPI16 ? 1 : 2;        // producer sets of 8 warps; set j fills k-blocks j, j + NSETS, ...
};
constexpr int PROD_SET_WARPS = 8;
constexpr int NTHREADS = 28 * 32;  // 896
constexpr int ROWS_PER_THREAD = 4;         // rows lane/4 + 8i of the warp's 32-row quarter
constexpr int PREFETCH = 2;                // k-blocks of A loads in flight per producer thread (4 per SM-wide k-block stream)
constexpr int A_STAGE_COLS = 64;           // TMEM columns per A stage: hi 32 | lo 32
constexpr int MAX_A_STAGES = 4;
// B200_TC_DEBUG role masks (timing experiments, tools/exp/sweep_*.sh) are compiled in only with -DB200_TC_DEBUG_MASKS=1
// (`make debug` -> lib/variants/libb200rt_dbg.so, selected with B200RT_LIB): the producers are instruction-bound, and
// the mask tests sat in their inner loop.
#ifndef B200_TC_DEBUG_MASKS
#define B200_TC_DEBUG_MASKS 0
#endif
#define TC_DBG(bit) (B200_TC_DEBUG_MASKS && (p.debug & (bit)))
// Role-mask build only, B200_TC_DEBUG bit 1024: CTA 0 stamps clock64() at the hand-over points of its first 96 tiles
// (g_ts[role][tile][event]); launch_conv_tc prints them after the launch (tools/exp/tile_timeline.sh).
#if B200_TC_DEBUG_MASKS
__device__ unsigned long long g_ts[3][96][4];
#define TC_TS(role, tile, ev) do { if ((p.debug & 1024) && blockIdx.x == 0 && (tile) < 96 && (threadIdx.x & 31) == 0) g_ts[role][tile][ev] = clock64(); } while (0)
#else
#define TC_TS(role, tile, ev) do { } while (0)
#endif
// Longest reduction that uses one merged accumulator (see launch_conv_tc).  0 = never: the merged accumulator triples the
// truncating additions into the large accumulator, and on post-Relu (same-sign) inputs -- what the 3x3 expands really
// see -- it measured 1.65x the tolerance at K = 288 and 0.84x at K = 144 (profiles/r2_accumulator_accuracy.txt;
// {main | correction}: 0.67 / 0.28).  B200_TC_MERGED=1 still forces it for experiments.
constexpr int MERGED_MAX_K = 0;
constexpr int SMEM_MAX = 227 * 1024;
constexpr int MAX_STAGES = 6;              // weight ring
constexpr int MAX_RAW = 8;                 // raw A ring (pointwise / TMA-fed mode)

struct TcParams {
  ConvArgs a;
  int BN;          // output channels per tile (multiple of 16, <= 128)
  int S;           // weight ring stages (shared memory)
  int SA;          // A stages (tensor memory)
  int nacc;        // accumulator stages (tensor memory)
  int merged;      // 1: hi*hi, hi*lo and lo*hi accumulate into ONE accumulator of BN columns (three N = BN MMAs per k-step);
                   // 0: {main | correction} accumulators of BN columns each (one N = 2*BN and one N = BN MMA per k-step)
  int nkb;         // k-blocks per tile
  int n_tiles_n;   // channel tiles
  int reverse;     // walk the tiles from the last to the first (see ConvArgs::reverse)
  int total_tiles; // pixel tiles x channel tiles
  int Mpad;        // weight rows per half (hi / lo)
  int P;           // output pixels (fits in int32, checked on the host)
  int vec_store;   // destination base and pitch are 16-byte aligned: 128-bit stores
  int a_tma;       // 1: A tiles arrive by TMA into a raw ring (pointwise layers); 0: register gather
  int R;           // raw ring slots (a_tma)
  int raw_slot;    // bytes per raw ring slot: A_TILE_BYTES, or 2 * pool_rows + 1 window-row boxes in the pooled mode
  int rowbox;      // pooled mode: bytes per window-row box (2*Wo+1 pixels x 128 B, rounded up to the 1 KB swizzle atom)
  int tile_rows;   // output pixels per tile: BM, or pool_rows * Wo in the pooled mode
  int pool_rows;   // pooled mode: a tile is pool_rows consecutive pooled rows of one image (MMA row = rho * Wo + j)
  int pool_tpi;    // pooled mode: tiles per image = ceil(Ho / pool_rows)
  int half_a;      // 1: half-stage A ring (template HALFA): four {hi 16 | lo 16} stages + two accumulator stages (64 < BN <= 96, gather)
  int slab_pitch;  // bytes per row of an epilogue warp's slab (128 or 256)
  uint32_t zero;   // always 0; opaque to the compiler (builds data dependencies that must survive optimisation)
  int debug;       // B200_TC_DEBUG bit mask (timing experiments only): 1 skip weight TMA, 2 skip A loads, 4 skip stores, 16 skip MMAs, 32 skip the producers' tcgen05.st, 256 issue every MMA with N = 16
  int nopad;       // every tap of every output pixel lies inside the image: no validity masks (gather mode)
  int skip_cols;   // fused Fire expand: leading filters that are zero in k-blocks >= skip_kb (0 = none), a multiple of 16
  int skip_kb;
  unsigned long long m64Wo, m64Ho, m64NT;   // ceil(2^64 / d), 0 when d == 1: exact n / d for n < 2^32 by one multiply-high
  uint32_t magicC, magicKW, magicWo, magicHo;  // ceil(2^32 / d), 0 when d == 1: exact n / d for n, d < 2^16
};

#include "tc_ptx.cuh"

// ------------------------------------------------------------------------------------------------ the kernel
// NOPAD (gather layers whose taps all lie inside the image, e.g. conv1): a compile-time variant, so that the padded path's
// masks and zero fill cost the no-padding producers neither registers nor instructions
template <bool HAS_ADD, bool EPI16, bool NOPAD, bool POOL, bool HALFA>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapA, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // Programmatic dependent launch (ConvArgs::pdl): let the next launch on the stream take SMs as this grid's CTAs leave.
  // A no-op when nothing was launched as a dependent.
  asm volatile("griddepcontrol.launch_dependents;");
  using R_ = Roles<EPI16>;
  constexpr int NUM_EPI_WARPS = R_::NEPI, TMA_WARP = R_::TMA_WARP, MMA_WARP = R_::MMA_WARP, ATMA_WARP = R_::ATMA_WARP,
                PROD_WARP0 = R_::PROD_WARP0, NSETS = R_::NSETS;
  const ConvArgs& a = p.a;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_tile_bytes = p.BN * BK * 4;
  const int stage_bytes = 2 * b_tile_bytes;                                      // B_hi | B_lo
  const uint32_t raw_ring = smem_base + (uint32_t)p.S * stage_bytes;             // R x 16 KB raw fp32 A tiles (a_tma)
  const uint32_t epi_slabs = raw_ring + (uint32_t)(p.R * p.raw_slot);            // 8 x 32 rows x slab_pitch
  const uint32_t sbias = epi_slabs + (uint32_t)(NUM_EPI_WARPS * 32 * p.slab_pitch);  // Mpad floats: bias, zero padded
  const uint32_t sadd = sbias + 4u * (uint32_t)p.Mpad;                           // Mpad floats: folded channel add
  const uint32_t bars = sadd + 4u * (uint32_t)p.Mpad;                            // 8-byte mbarriers
  auto full_b = [&](int s) { return bars + 8u * s; };
  auto empty_b = [&](int s) { return bars + 8u * (MAX_STAGES + s); };
  auto full_a = [&](int s) { return bars + 8u * (2 * MAX_STAGES + s); };
  auto empty_a = [&](int s) { return bars + 8u * (2 * MAX_STAGES + MAX_A_STAGES + s); };
  auto tmem_full = [&](int s) { return bars + 8u * (2 * MAX_STAGES + 2 * MAX_A_STAGES + s); };
  auto tmem_empty = [&](int s) { return bars + 8u * (2 * MAX_STAGES + 2 * MAX_A_STAGES + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * MAX_STAGES + 2 * MAX_A_STAGES + 4);
  auto raw_full = [&](int r) { return bars + 8u * (2 * MAX_STAGES + 2 * MAX_A_STAGES + 5 + r); };
  auto raw_empty = [&](int r) { return bars + 8u * (2 * MAX_STAGES + 2 * MAX_A_STAGES + 5 + MAX_RAW + r); };
  // gather mode: per producer set two row tables (128 rows x {element offset of tap (0,0) channel 0, validity masks}),
  // written one tile ahead by the set's decode warp (ROWDEC_WARP0 + set), and their full / empty barriers
  auto tab_full = [&](int set, int b) { return raw_empty(MAX_RAW) + 8u * (uint32_t)(set * 2 + b); };
  auto tab_empty = [&](int set, int b) { return raw_empty(MAX_RAW) + 8u * (uint32_t)(4 + set * 2 + b); };
  const uint32_t rowtab = (tab_empty(1, 1) + 8u + 15u) & ~15u;   // [set][buffer][128 rows] x 8 bytes
  auto rowtab_of = [&](int set, int b) { return rowtab + 1024u * (uint32_t)(set * 2 + b); };
  const uint32_t ktab = rowtab + 4096u;   // gather mode: nkb x 8 entries {delta, r, s, valid mask}
  // tile index -> (pixel tile, channel tile), one multiply-high instead of an integer division per tile and role
  auto tile_pt = [&](int t) { const int tt = p.reverse ? p.total_tiles - 1 - t : t; return p.m64NT ? (int)__umul64hi((unsigned long long)tt, p.m64NT) : tt; };
  auto tile_nt = [&](int t, int pt) { const int tt = p.reverse ? p.total_tiles - 1 - t : t; return tt - pt * p.n_tiles_n; };
  const int acc_cols = p.merged ? p.BN : 2 * p.BN;   // TMEM columns per accumulator stage
  const int a_col0 = p.nacc * acc_cols;   // first TMEM column of the A stages

  // warp index through a shuffle from lane 0: the compiler then knows it is warp-uniform and keeps everything derived
  // from it (role branches, TMEM addresses, barrier addresses) in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  // per-channel epilogue constants -> shared memory once per CTA (the epilogue then needs no global loads)
  for (int m = threadIdx.x; m < p.Mpad; m += NTHREADS) {
    const float b = (a.bias && m < a.M) ? __ldg(a.bias + m) : 0.f;
    const float c = (a.chan_add && m < a.M) ? __ldg(a.chan_add + m) : 0.f;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + 4u * m), "f"(b) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sadd + 4u * m), "f"(c) : "memory");
  }

  // gather mode: what a 16-byte chunk of a k-block means is the same for every tile, so the (r, s, c) decode is
  // done once per CTA: entry (kb, chunk) = {element offset of tap (r, s) channel c from a row's base, r, s,
  // 0x01010101 when k < K else 0}
  if (!p.a_tma) {
    for (int e = threadIdx.x; e < p.nkb * 8; e += NTHREADS) {
      const int k = e * 4;
      const int pos = k / a.C, c = k - pos * a.C;
      const int tap = a.tap_perm ? (int)((a.tap_perm >> (4 * pos)) & 15ull) : pos;   // K position -> tap (planner's Fire fusion)
      const int r = tap / a.KW, sx = tap - r * a.KW;
      const bool ok = k < a.K;
      const int delta = ok ? (r * a.W + sx) * a.ldx + c : 0;
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ktab + 16u * e), "r"(delta), "r"(ok ? r : 0), "r"(ok ? sx : 0),
                   "r"(ok ? 0x01010101u : 0u)
                   : "memory");
    }
  }

  if (warp == TMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < p.S; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
      // HALFA: four half-stages, each filled by the four warps (one per lane quarter) of one (set, k half)
      for (int s = 0; s < (HALFA ? 4 : p.SA); ++s) { mbar_init(full_a(s), HALFA ? PROD_SET_WARPS / 2 : PROD_SET_WARPS); mbar_init(empty_a(s), 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(tmem_full(s), 1); mbar_init(tmem_empty(s), NUM_EPI_WARPS); }
      for (int r = 0; r < p.R; ++r) { mbar_init(raw_full(r), 1); mbar_init(raw_empty(r), PROD_SET_WARPS); }
      for (int st = 0; st < 2; ++st)
        for (int b = 0; b < 2; ++b) { mbar_init(tab_full(st, b), 1); mbar_init(tab_empty(st, b), PROD_SET_WARPS); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512u);   // one CTA per SM: the whole tensor memory
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // Everything above touched parameters and constants only (bias, decode tables, barriers, tensor memory).  From here on
  // the roles read activations the previous launch wrote and overwrite buffers it may still be reading (the arena reuses
  // dead buffers): wait until the grids this one depends on have completed and their writes are visible.  Returns at
  // once when the launch was not a programmatic dependent.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);   // uniform for the compiler as well

  if (warp >= PROD_WARP0) {
    // ================================================================ A producers (16 warps, two sets of 8)
    // Warp w may only write TMEM lanes 32*(w % 4) .. +31: it owns rows 32*quarter + 8*i + lane/4 (i = 0..3) of the
    // tile and the 16-byte chunk 4*khalf + lane%4 of each 128-byte k-block row.  The two sets take alternate
    // k-blocks of the CTA's k-block stream (SA and R are even, so a set always meets the same A stages / raw
    // slots): a single warp's instruction stream was the pacing resource with 8 producer warps.
    if (EPI16) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");   // converters only: no prefetch registers
    const int pw = warp - PROD_WARP0;
    const int quarter = pw & 3, khalf = (pw >> 2) & 1, kpar = pw >> 3;   // kpar < NSETS
    const int chunk = khalf * 4 + (lane & 3);
    const int rsub = lane >> 2;
    int my_tiles = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) ++my_tiles;
    const int items = my_tiles * p.nkb;   // k-blocks of this CTA; this warp handles idx = kpar, kpar + 2, ...
    // TMEM address of this warp's part of A stage 0: lane 32*quarter, column a_col0 + 16*khalf
    const uint32_t t_a0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a_col0 + khalf * 16);
    int sa = kpar;          // A stage of the current k-block (idx % SA)
    uint32_t ph = 0;        // its phase

    if (p.a_tma) {
      // ---- TMA-fed mode (pointwise layers): the raw fp32 tile of each k-block is in shared memory (128B-swizzled
      // rows); up to R raw tiles (R x 16 KB) are in flight from HBM without costing a register.
      uint32_t off[ROWS_PER_THREAD];
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        const int row = quarter * 32 + i * 8 + rsub;
        off[i] = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
      }
      if (POOL) {
        // pooled mode: row rho * Wo + j of the tile is pixel j of the tile's pooled row rho; the slot holds the
        // 2 * pool_rows + 1 input rows those windows touch (boxes of 2*Wo+1 pixels x 32 channels, first pixel = column
        // -pool_pl), so the window of (rho, j) is pixels 2j .. 2j+2 of boxes 2 rho .. 2 rho + 2.  off[i] = pixel 2j of
        // box 2 rho (a box is a multiple of 1 KB: the swizzle term is still pixel & 7).
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i) {
          const int mrow = quarter * 32 + i * 8 + rsub;
          const int rho = mrow / a.Wo, j = mrow - rho * a.Wo;
          // rows past the tile's pixels are not computed (0xFFFFFFFF): with 54 of 128 rows in use two lane quarters idle
          off[i] = rho < p.pool_rows ? (uint32_t)(2 * rho * p.rowbox) + (uint32_t)(2 * j) * 128u : 0xFFFFFFFFu;
        }
      }
      int r = kpar;
      uint32_t rph = 0;
      for (int idx = kpar; idx < items; idx += NSETS) {
        mbar_wait(raw_full(r), rph);
        const uint32_t raw = raw_ring + (uint32_t)(r * p.raw_slot);
        float4 x[ROWS_PER_THREAD];
        if (POOL) {
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i) {
            if (off[i] == 0xFFFFFFFFu) { x[i] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
            // pixel 2j is even: the swizzle term of pixels 2j, 2j+1, 2j+2 is (2j & 7), +1, (+2) & 7
            const uint32_t px = off[i] >> 7;
            const uint32_t a0 = raw + off[i] + (((uint32_t)chunk ^ (px & 7u)) << 4);
            const uint32_t a1 = raw + off[i] + 128u + (((uint32_t)chunk ^ ((px + 1u) & 7u)) << 4);
            const uint32_t a2 = raw + off[i] + 256u + (((uint32_t)chunk ^ ((px + 2u) & 7u)) << 4);
            float4 m;
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
              const uint32_t ro = (uint32_t)(rr * p.rowbox);
              float4 v0, v1, v2;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(a0 + ro) : "memory");
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(a1 + ro) : "memory");
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v2.x), "=f"(v2.y), "=f"(v2.z), "=f"(v2.w) : "r"(a2 + ro) : "memory");
              // max_pool2d, max_pool_op.rs:157-360 (padding cells are zeros: the tensor map's out-of-bounds fill)
              v0.x = fmaxf(fmaxf(v0.x, v1.x), v2.x); v0.y = fmaxf(fmaxf(v0.y, v1.y), v2.y);
              v0.z = fmaxf(fmaxf(v0.z, v1.z), v2.z); v0.w = fmaxf(fmaxf(v0.w, v1.w), v2.w);
              if (rr == 0) m = v0;
              else { m.x = fmaxf(m.x, v0.x); m.y = fmaxf(m.y, v0.y); m.z = fmaxf(m.z, v0.z); m.w = fmaxf(m.w, v0.w); }
            }
            x[i] = m;   // depends on all nine loads: the release below waits for every one of them
          }
        } else {
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w) : "r"(raw + off[i]) : "memory");
        }
        float4 hi[ROWS_PER_THREAD], lo[ROWS_PER_THREAD];
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i) {
          hi[i].x = split_hi(x[i].x); hi[i].y = split_hi(x[i].y); hi[i].z = split_hi(x[i].z); hi[i].w = split_hi(x[i].w);
          lo[i].x = split_lo(x[i].x, hi[i].x); lo[i].y = split_lo(x[i].y, hi[i].y);
          lo[i].z = split_lo(x[i].z, hi[i].z); lo[i].w = split_lo(x[i].w, hi[i].w);
        }
        // The slot may be refilled by TMA as soon as raw_empty completes, so every ld.shared above must have READ it
        // before the arrive is issued.  Having issued them is not enough (observed: with the input resident in L2 the
        // refill overtook the last loads of a warp about once in 20 runs, rows 8i + lane/4 of one tile wrong), and the
        // compiler sinks the split below the arrive.  So the arrive's address is made to depend on one register of
        // each load: the instruction cannot issue before their scoreboards clear.  (The whole warp's loads are one
        // instruction each, so lane 0's dependency covers the other lanes.)
        // (p.zero is a kernel parameter that is always 0: an `& 0` literal would be folded away by ptxas)
        const uint32_t dep = (__float_as_uint(x[0].w) ^ __float_as_uint(x[1].w) ^ __float_as_uint(x[2].w) ^ __float_as_uint(x[3].w)) & p.zero;
        __syncwarp();
        if (lane == 0) mbar_arrive(raw_empty(r) + dep);
        r += NSETS;
        if (r >= p.R) { r -= p.R; rph ^= 1u; }
        mbar_wait(empty_a(sa), ph ^ 1u);
        tc_fence_after();
        if (!TC_DBG(32)) {
          const uint32_t t_stage = t_a0 + (uint32_t)(sa * A_STAGE_COLS);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t ta = t_stage + ((uint32_t)(j * 16) << 16);
            tmem_st_16x256b_x2(ta, hi[2 * j].x, hi[2 * j].y, hi[2 * j + 1].x, hi[2 * j + 1].y, hi[2 * j].z, hi[2 * j].w, hi[2 * j + 1].z, hi[2 * j + 1].w);
            tmem_st_16x256b_x2(ta + 32u, lo[2 * j].x, lo[2 * j].y, lo[2 * j + 1].x, lo[2 * j + 1].y, lo[2 * j].z, lo[2 * j].w, lo[2 * j + 1].z, lo[2 * j + 1].w);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_a(sa));
        sa += NSETS;
        if (sa >= p.SA) { sa -= p.SA; ph ^= 1u; }
      }
    } else {
      // ---- register gather.  Per tile each thread fixes, for its ROWS_PER_THREAD im2col rows, an element offset
      // (tap (0,0), channel 0) and two separable validity masks (bit 8*i+r: input row h0+r inside the image; bit
      // 8*i+s: column w0+s inside).  Per k-block the chunk's (r, s, offset) comes from the per-CTA table, shared by
      // all rows, so a load costs an add, a mask test and the LDG.
      // load cursor: the tile, and the ADDRESS of this thread's entry of the k-block's row of the decode table
      // (128 bytes per k-block) instead of a k-block index -- the index would have to be combined with the thread's
      // chunk for every lookup, and under the 72-register cap the compiler rebuilt the chunk from %tid each time
      int l_tile = blockIdx.x;
      uint32_t l_ent = ktab + 16u * (uint32_t)chunk + 128u * (uint32_t)kpar;
      const uint32_t ktab_end = ktab + 128u * (uint32_t)p.nkb, ktab_bytes = 128u * (uint32_t)p.nkb;   // warp-uniform
      while (l_ent >= ktab_end) { l_ent -= ktab_bytes; l_tile += gridDim.x; }
      uint32_t base[ROWS_PER_THREAD];         // element offsets from a.x (< 2^31, checked on the host; int arithmetic)
      uint32_t hmask = 0, wmask = 0;
      // Per tile the thread needs, for its four rows, the element offset of tap (0,0) channel 0 and the validity
      // masks.  They come from the set's row table in shared memory, filled one tile ahead by the set's decode warp
      // (below): the decode is ~160 instructions for four rows, and with every producer warp doing it for itself it was
      // a third of the producers' instructions on the short-reduction layers (conv1, fire2 / fire3).
      uint32_t tcount = 0;   // tiles this set has entered: table buffer = tcount & 1, barrier phase = (tcount >> 1) & 1
      auto set_tile = [&]() {
        const uint32_t tb = tcount & 1u;
        mbar_wait(tab_full(kpar, tb), (tcount >> 1) & 1u);
        const uint32_t rt = rowtab_of(kpar, tb) + 8u * (uint32_t)(quarter * 32 + rsub);
        uint32_t mk[ROWS_PER_THREAD];
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i)
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(base[i]), "=r"(mk[i]) : "r"(rt + 64u * i) : "memory");
        hmask = (mk[0] & 0xFFu) | ((mk[1] & 0xFFu) << 8) | ((mk[2] & 0xFFu) << 16) | (mk[3] << 24);
        wmask = ((mk[0] >> 8) & 0xFFu) | (mk[1] & 0xFF00u) | ((mk[2] & 0xFF00u) << 8) | ((mk[3] & 0xFF00u) << 16);
        // the buffer goes back to the decode warp only after the loads above have delivered (same device as raw_empty)
        const uint32_t dep = (mk[0] ^ mk[1] ^ mk[2] ^ mk[3]) & p.zero;
        __syncwarp();
        if (lane == 0) mbar_arrive(tab_empty(kpar, tb) + dep);
        ++tcount;
      };
      float4 v[PREFETCH][ROWS_PER_THREAD];
      auto issue = [&](float4 (&dst)[ROWS_PER_THREAD]) {
        uint32_t delta, r, sx, vm;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(delta), "=r"(r), "=r"(sx), "=r"(vm) : "r"(l_ent));
        if (NOPAD) {
          // every tap is inside the image: four plain loads.  (The chunks past K in the last k-block have delta = 0:
          // they re-read tap (0,0) and meet zero weights, exact for finite data like the fused Fire module's zeros.)
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i) {
            if (TC_DBG(2)) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            else dst[i] = __ldg(reinterpret_cast<const float4*>(a.x + (int)(base[i] + delta)));
          }
        } else {
          const uint32_t m = (hmask >> r) & (wmask >> sx) & vm;   // bit 8*i: row i valid for this tap
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i) {
            dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (((m >> (8 * i)) & 1u) && !TC_DBG(2)) dst[i] = __ldg(reinterpret_cast<const float4*>(a.x + (int)(base[i] + delta)));
          }
        }
        l_ent += 128u * NSETS;
        if (l_ent >= ktab_end) {
          do { l_ent -= ktab_bytes; l_tile += gridDim.x; } while (l_ent >= ktab_end);
          if (l_tile < p.total_tiles) set_tile();
        }
      };
      const int my_items = (items - kpar + NSETS - 1) / NSETS;   // k-blocks this warp handles
      if (my_items > 0) set_tile();
#pragma unroll
      for (int d = 0; d < PREFETCH; ++d)
        if (d < my_items) issue(v[d]);

      for (int base_i = 0; base_i < my_items; base_i += PREFETCH) {
#pragma unroll
        for (int d = 0; d < PREFETCH; ++d) {
          const int it = base_i + d;
          if (it < my_items) {
            if (HALFA) {
              // half-stage ring (BN = 96 gather layers, so that TWO accumulator stages fit tensor memory): this warp's
              // 16 k-floats of the k-block are a stage of their own -- {hi 16 | lo 16} columns at half-stage
              // 2 * set + k half, always the same one, released by the MMA warp after two k-steps instead of four
              const int hs = 2 * kpar + khalf;
              mbar_wait(empty_a(hs), ph ^ 1u);
              tc_fence_after();
              if (!TC_DBG(32)) split_store(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a_col0 + hs * 32), v[d], 16u);
              if (it + PREFETCH < my_items) issue(v[d]);
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(full_a(hs));
              ph ^= 1u;
            } else {
            mbar_wait(empty_a(sa), ph ^ 1u);   // the MMAs that read this A stage have completed
            tc_fence_after();
            if (!TC_DBG(32)) split_store(t_a0 + (uint32_t)(sa * A_STAGE_COLS), v[d]);
            if (it + PREFETCH < my_items) issue(v[d]);   // next loads go out before the store wait
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(sa));
            sa += NSETS;
            if (sa >= p.SA) { sa -= p.SA; ph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp >= NUM_EPI_WARPS) {
   // ================================================================ TMA / MMA / A-TMA / idle warpgroup
   // these four warps need few registers: hand the rest to the epilogue warpgroups (setmaxnreg is warpgroup-wide)
   asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
   if (warp == TMA_WARP) {
    // ================================================================ weight tiles via TMA
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int m0 = tile_nt(t, tile_pt(t)) * p.BN;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(empty_b(s), ph ^ 1u);
          const uint32_t dst = smem_base + (uint32_t)s * stage_bytes;
          if (TC_DBG(1)) { mbar_arrive(full_b(s)); if (++s == p.S) { s = 0; ph ^= 1u; } continue; }
          mbar_expect_tx(full_b(s), 2u * (uint32_t)b_tile_bytes);
          tma_load_2d(dst, &tmapB, full_b(s), kb * BK, m0);
          tma_load_2d(dst + b_tile_bytes, &tmapB, full_b(s), kb * BK, p.Mpad + m0);
          if (++s == p.S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (!EPI16 && !p.a_tma && (warp == ATMA_WARP || warp == ATMA_WARP + 1)) {
    // ================================================================ row decode for the gather producers
    // Warp ATMA_WARP + set serves producer set `set`: for every tile the set enters, in the set's order, the 128 rows'
    // {element offset of tap (0,0) channel 0 from a.x, validity masks} go to one of the set's two row tables.
    // (With one k-block per tile the sets take alternate tiles; otherwise both enter every tile.)
    const int set = warp - ATMA_WARP;
    const int first = (p.nkb == 1) ? set : 0, step = (p.nkb == 1) ? NSETS : 1;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x + first * gridDim.x; tile < p.total_tiles; tile += step * gridDim.x, ++tcount) {
      const uint32_t tb = tcount & 1u;
      mbar_wait(tab_empty(set, tb), ((tcount >> 1) & 1u) ^ 1u);
      // first pixel of the tile -> (n, ho, wo) with 64-bit multiply-highs; the rows follow with 32-bit multiply-high
      // divisions of small numbers (row < 128, so the carries stay below 2^16)
      const int p0 = tile_pt(tile) * BM;
      const int t0 = p.m64Wo ? (int)__umul64hi((unsigned long long)p0, p.m64Wo) : p0, wo0 = p0 - t0 * a.Wo;
      const int n0 = p.m64Ho ? (int)__umul64hi((unsigned long long)t0, p.m64Ho) : t0, ho0 = t0 - n0 * a.Ho;
      const int last_row = p.P - 1 - p0;   // rows past the last pixel (last tile only)
      const uint32_t tab = rowtab_of(set, tb);
#pragma unroll 1
      for (int i = 0; i < BM / 32; ++i) {
        int row = i * 32 + lane;
        const bool ok = row <= last_row;
        const int rr = NOPAD ? min(row, last_row) : row;   // no masks: a row past the end re-reads the last pixel (never stored)
        const int wsum = wo0 + rr;
        const int cw = p.magicWo ? (int)__umulhi((unsigned)wsum, p.magicWo) : wsum;   // wsum / Wo
        const int wo = wsum - cw * a.Wo;
        const int hsum = ho0 + cw;
        const int ch = p.magicHo ? (int)__umulhi((unsigned)hsum, p.magicHo) : hsum;   // hsum / Ho
        const int ho = hsum - ch * a.Ho;
        const int n = n0 + ch;
        const int h0 = ho * a.sh - a.pt, w0 = wo * a.sw - a.pl;
        const int rbase = ((n * a.H + h0) * a.W + w0) * a.ldx;   // may be "negative" for padded taps: never dereferenced then
        // taps r with 0 <= h0 + r < H form the bit range [max(0,-h0), min(KH, H-h0)); same for s
        const int rlo = max(0, -h0), rhi = min(a.KH, a.H - h0);
        const int slo = max(0, -w0), shi = min(a.KW, a.W - w0);
        const uint32_t hm = (ok && rhi > rlo) ? (((1u << rhi) - 1u) & ~((1u << rlo) - 1u)) : 0u;
        const uint32_t wm = (shi > slo) ? (((1u << shi) - 1u) & ~((1u << slo) - 1u)) : 0u;
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(tab + 8u * (uint32_t)row), "r"(rbase), "r"(hm | (wm << 8)) : "memory");
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(tab_full(set, tb));
    }
  } else if (warp == ATMA_WARP) {
    // ================================================================ raw A tiles via TMA (pointwise layers only)
    if (p.a_tma && lane == 0) {
      int r = 0;
      uint32_t rph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int pt = tile_pt(t);
        const int p0 = pt * BM;
        // pooled mode: tile = pooled rows pool_rows * th .. of image n; their windows read 2 * pool_rows + 1 input rows
        // from row 2 * pool_rows * th - pool_pt on
        const int pn = POOL ? pt / p.pool_tpi : 0;
        const int h0 = POOL ? 2 * p.pool_rows * (pt - pn * p.pool_tpi) - a.pool_pt : 0;
        const int nbox = 2 * p.pool_rows + 1;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(raw_empty(r), rph ^ 1u);
          if (POOL) {
            // boxes of 32 channels x (2 Wo + 1) pixels x one input row; rows / columns outside the image and channels
            // past C are zero-filled by the tensor map -- the reference's MaxPool pads with zeros
            const uint32_t box_tx = (uint32_t)(2 * a.Wo + 1) * 128u;
            mbar_expect_tx(raw_full(r), (uint32_t)nbox * box_tx);
            const uint32_t slot = raw_ring + (uint32_t)(r * p.raw_slot);
            for (int rr = 0; rr < nbox; ++rr) tma_load_4d(slot + (uint32_t)(rr * p.rowbox), &tmapA, raw_full(r), kb * BK, -a.pool_pl, h0 + rr, pn);
          } else {
            mbar_expect_tx(raw_full(r), (uint32_t)A_TILE_BYTES);
            // box = 32 channels x 128 pixel rows; rows past P and channels past C are zero-filled by the tensor map
            tma_load_2d(raw_ring + (uint32_t)r * A_TILE_BYTES, &tmapA, raw_full(r), kb * BK, p0);
          }
          if (++r == p.R) { r = 0; rph ^= 1u; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================================================================ MMA issuer
    // The whole warp walks the pipeline (uniform control flow, barrier waits by all lanes); one elected lane issues.
    // The issue stream is a critical resource (measured: while descriptors were rebuilt from addresses inside a
    // single-lane branch the time per k-block did not depend on BN), so the 64-bit shared-memory descriptors are
    // kept as 32-bit low words that advance by adds: +2 per k-step (32 bytes >> 4), +stage_bytes/16 per stage.
    // B200_TC_DEBUG bit 256 (timing experiment): issue every MMA with N = 16, whatever BN is
    const uint32_t idesc = instr_desc_tf32(TC_DBG(256) ? 16 : p.BN), idesc2 = instr_desc_tf32(TC_DBG(256) ? 16 : 2 * p.BN);
    const uint32_t idesc_skip = instr_desc_tf32(p.BN - p.skip_cols);
    const bool leader = elect_one();
    const uint32_t lo_first = ((smem_base >> 4) & 0x3FFFu) | (1u << 16);   // [0,14) address >> 4, [16,30) LBO = 1
    const uint32_t lo_step = (uint32_t)stage_bytes >> 4;
    uint32_t lo = lo_first;
    int s = 0, sa = 0;
    uint32_t ph = 0, pha = 0;
    int tc = 0;
    // k-steps of the last k-block: K is permuted inside groups of 16, so whole groups (2 k-steps) are issued
    const int tail_ksteps = 2 * ((a.K - (p.nkb - 1) * BK + 15) >> 4);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tc) {
      const int as = p.nacc == 2 ? (tc & 1) : 0;
      const uint32_t aph = (p.nacc == 2 ? (uint32_t)(tc >> 1) : (uint32_t)tc) & 1u;
      TC_TS(0, tc, 3);                       // arrived at the tile
      mbar_wait(tmem_empty(as), aph ^ 1u);   // epilogue has drained this accumulator stage
      tc_fence_after();
      TC_TS(0, tc, 0);                       // accumulator stage free
      const uint32_t d_main = tmem_base + (uint32_t)(as * acc_cols);
      const uint32_t d_corr = p.merged ? d_main : d_main + (uint32_t)p.BN;
      const uint32_t b_lo = (uint32_t)b_tile_bytes >> 4;   // B_lo follows B_hi in the stage (descriptor units of 16 bytes)
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (HALFA) {
          // half-stage ring: `sa` counts the CTA's k-blocks modulo 2 (the producer set that filled this one), `pha` is the
          // phase of that set's two half-stages.  Two k-steps per half-stage, each half released on its own.
          mbar_wait(full_b(s), ph);
          const int ksteps = (kb == p.nkb - 1) ? tail_ksteps : 4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int hs = 2 * sa + h;
            mbar_wait(full_a(hs), pha);
            tc_fence_after();
            if (kb == 0 && h == 0) TC_TS(0, tc, 1);
            if (leader && !TC_DBG(16)) {
              const uint32_t ah = tmem_base + (uint32_t)(a_col0 + hs * 32), al = ah + 16u;
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2) {
                const int kk = 2 * h + k2;
                if (kk < ksteps) {
                  umma_tf32_ts(d_main, ah + 8u * k2, sw128_desc(lo + 2 * kk), idesc2, (kb | kk) != 0 ? 1u : 0u);
                  umma_tf32_ts(d_corr, al + 8u * k2, sw128_desc(lo + 2 * kk), idesc, 1u);
                }
              }
            }
            __syncwarp();
            if (leader) umma_commit(empty_a(hs));
          }
          if (leader) umma_commit(empty_b(s));
          lo += lo_step;
          if (++s == p.S) { s = 0; ph ^= 1u; lo = lo_first; }
          if (++sa == 2) { sa = 0; pha ^= 1u; }
          continue;
        }
        mbar_wait(full_a(sa), pha);  // A stage written to tensor memory by all 8 producer warps
        mbar_wait(full_b(s), ph);
        tc_fence_after();
        if (kb == 0) TC_TS(0, tc, 1);   // first k-block's operands there
        if (leader && !TC_DBG(16)) {
          const uint32_t ah = tmem_base + (uint32_t)(a_col0 + sa * A_STAGE_COLS), al = ah + 32u;
          const uint32_t bh = lo;
          const int ksteps = (kb == p.nkb - 1) ? tail_ksteps : 4;
          // B_hi and B_lo are adjacent in the stage ([2*BN rows] x 128 B), and so are the two accumulators
          // ([main | correction] = 2*BN TMEM columns): ONE N = 2*BN instruction computes A_hi*B_hi -> main and
          // A_hi*B_lo -> correction; a second N = BN instruction adds A_lo*B_hi.
          // Merged mode (BN > 64, where {main | correction} x 2 stages would not fit tensor memory): three N = BN
          // instructions per k-step into one accumulator -- the same tensor-pipe time, and two accumulator stages fit.
          if (p.merged) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (kk < ksteps) {
                umma_tf32_ts(d_main, ah + 8u * kk, sw128_desc(bh + 2 * kk), idesc, (kb | kk) != 0 ? 1u : 0u);
                umma_tf32_ts(d_main, ah + 8u * kk, sw128_desc(bh + b_lo + 2 * kk), idesc, 1u);
                umma_tf32_ts(d_main, al + 8u * kk, sw128_desc(bh + 2 * kk), idesc, 1u);
              }
            }
          } else if (p.skip_cols > 0 && kb >= p.skip_kb) {
            // Fused Fire expand: in these k-blocks the first skip_cols filters (the 1x1 branch, whose only tap sits in the
            // first k-blocks) are exact zeros -- issue the other columns only: three N = BN - skip instructions per k-step
            // instead of N = 2 BN + N = BN (BN = 128, skip = 64: 124 instead of 211 clk).  hi*hi -> main, hi*lo and lo*hi
            // -> correction; B_hi / B_lo rows and accumulator columns offset by skip.
            const uint32_t rs = (uint32_t)p.skip_cols * 8u;   // skip rows x 128 B >> 4
            const uint32_t dm = d_main + (uint32_t)p.skip_cols, dc = d_corr + (uint32_t)p.skip_cols;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (kk < ksteps) {
                umma_tf32_ts(dm, ah + 8u * kk, sw128_desc(bh + rs + 2 * kk), idesc_skip, 1u);
                umma_tf32_ts(dc, ah + 8u * kk, sw128_desc(bh + b_lo + rs + 2 * kk), idesc_skip, 1u);
                umma_tf32_ts(dc, al + 8u * kk, sw128_desc(bh + rs + 2 * kk), idesc_skip, 1u);
              }
            }
          } else {
            umma_tf32_ts(d_main, ah, sw128_desc(bh), idesc2, kb != 0 ? 1u : 0u);
            umma_tf32_ts(d_corr, al, sw128_desc(bh), idesc, 1u);
#pragma unroll
            for (int kk = 1; kk < 4; ++kk) {
              if (kk < ksteps) {
                umma_tf32_ts(d_main, ah + 8u * kk, sw128_desc(bh + 2 * kk), idesc2, 1u);
                umma_tf32_ts(d_corr, al + 8u * kk, sw128_desc(bh + 2 * kk), idesc, 1u);
              }
            }
          }
        }
        __syncwarp();
        if (leader) {
          umma_commit(empty_b(s));    // frees the weight stage once the MMAs above have read it
          umma_commit(empty_a(sa));   // and the A stage in tensor memory
        }
        lo += lo_step;
        if (++s == p.S) { s = 0; ph ^= 1u; lo = lo_first; }
        if (++sa == p.SA) { sa = 0; pha ^= 1u; }
      }
      if (leader) umma_commit(tmem_full(as));   // accumulator stage complete -> epilogue
      __syncwarp();
      TC_TS(0, tc, 2);                           // last MMA issued
    }
   }
  } else {
    // register budget (64 K per SM): 16 producer warps x 72 + 4 x 40 + 8 epilogue warps x 88, or
    //                                  8 converter warps x 64 + 4 x 40 + 16 epilogue warps x 80
    if (EPI16) asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // ================================================================ epilogue (warps 0 .. NUM_EPI_WARPS-1)
    // Warp w may only read TMEM lanes 32*(w % 4) .. +31, so warps w, w+4, .. work on the same 32 accumulator rows and
    // take the 16-channel groups [g_begin, g_end) each.  Phase 1 (drain): per group tcgen05.ld main + correction,
    // add, + bias (+ channel add), Relu, and park the row in the warp's private slab (row per thread, 16-byte chunks
    // XOR-swizzled by row: conflict-free); then the accumulator stage goes back to the MMA warp.  Phase 2 (store):
    // consecutive lanes take consecutive 16-byte chunks of a row, so the global stores cover whole sectors of the
    // channels-last destination (which may be a channel slice of a Concat result).
    const int quarter = warp & 3, part = warp >> 2;   // NUM_EPI_WARPS / 4 warps share a lane quarter and split its channels
    constexpr int PARTS = NUM_EPI_WARPS / 4;
    const int G = p.BN >> 4;
    const int g_begin = (G * part + PARTS - 1) / PARTS, g_end = (G * (part + 1) + PARTS - 1) / PARTS;
    const int ng = g_end - g_begin;            // 0..4 groups of 16 channels
    const int nchunk = ng * 4;                 // 16-byte chunks per slab row
    const uint32_t slab = epi_slabs + (uint32_t)(warp * 32 * p.slab_pitch);
    const bool do_relu = a.relu != 0;
    const uint32_t my_row = slab + (uint32_t)(lane * p.slab_pitch);
    const uint32_t lane_sw = (uint32_t)(lane & 7) << 4;   // XOR swizzle of this lane's slab row
    const bool fast_store = p.vec_store && (a.M & 3) == 0;   // every 16-byte chunk is all-or-nothing
    int tc = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tc) {
      const int as = p.nacc == 2 ? (tc & 1) : 0;
      const uint32_t aph = (p.nacc == 2 ? (uint32_t)(tc >> 1) : (uint32_t)tc) & 1u;
      const int ptile = tile_pt(t);
      // first output pixel of the tile and one past its last; pooled mode: pooled rows pool_rows * th .. of image pn
      int p0, p_end;
      if (POOL) {
        const int pn = ptile / p.pool_tpi, th = ptile - pn * p.pool_tpi;
        p0 = (pn * a.Ho + p.pool_rows * th) * a.Wo;
        p_end = p0 + min(p.pool_rows, a.Ho - p.pool_rows * th) * a.Wo;
      } else {
        p0 = ptile * BM;
        p_end = min(p.P, p0 + BM);
      }
      const int m0 = tile_nt(t, ptile) * p.BN;
      if (warp == 0) TC_TS(1, tc, 3);   // arrived at the tile (previous store phase done)
      mbar_wait(tmem_full(as), aph);
      tc_fence_after();
      if (warp == 0) TC_TS(1, tc, 0);   // accumulator complete
      const uint32_t t_main = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * acc_cols);
      const bool two_acc = !p.merged;
      // software-pipelined drain (the extra registers come from setmaxnreg): the tcgen05.ld of group g+1 is in flight
      // while group g gets its bias / Relu and goes to the slab
      // (gather layers: conv1 0.635 -> 0.608 ms; the 16-epilogue-warp layout is better off without it: expand1x1
      // 0.084 -> 0.098 ms with it)
      constexpr bool PIPE = !EPI16;
      uint32_t acc[16], cor[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) cor[j] = 0u;   // merged mode: no correction accumulator, acc + 0 is exact
      if (PIPE && g_begin < g_end) {
        tmem_ld16(t_main + (uint32_t)(g_begin * 16), acc);
        if (two_acc) tmem_ld16(t_main + (uint32_t)(p.BN + g_begin * 16), cor);
        tmem_ld_wait(acc, cor);
      }
#pragma unroll 1   // unrolled, the short body makes ptxas spill the registers the pending tcgen05.ld write
      for (int g = g_begin; g < g_end; ++g) {
        if (!PIPE) {
          tmem_ld16(t_main + (uint32_t)(g * 16), acc);
          if (two_acc) tmem_ld16(t_main + (uint32_t)(p.BN + g * 16), cor);
          tmem_ld_wait(acc, cor);
        }
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(acc[j]) + __uint_as_float(cor[j]);   // hi*hi + (lo*hi + hi*lo)
        if (PIPE && g + 1 < g_end) {
          tmem_ld16(t_main + (uint32_t)((g + 1) * 16), acc);
          if (two_acc) tmem_ld16(t_main + (uint32_t)(p.BN + (g + 1) * 16), cor);
        }
        const uint32_t crow = my_row + ((uint32_t)(g - g_begin) << 6);
        // The drain sits between two tiles' MMAs when there is one accumulator stage, and it is not the tensor-memory
        // reads that make it long (tools/exp/ldtm_rates.cu: 8 warps drain 128 x 256 columns in ~400 clk) but its
        // instructions, issued next to 16 busy producer warps: so only the sum goes to the slab here; bias, folded
        // Add and Relu are applied by the store phase, after the accumulator has gone back to the MMA warp.
        // 16-byte chunk (g - g_begin) * 4 + q of the slab row, XOR-swizzled by the row
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((crow + 16u * q) ^ lane_sw), "f"(o[4 * q]), "f"(o[4 * q + 1]), "f"(o[4 * q + 2]), "f"(o[4 * q + 3]) : "memory");
        if (PIPE && g + 1 < g_end) tmem_ld_wait(acc, cor);
      }
      // all tcgen05.ld of this accumulator stage have completed: hand it back to the MMA warp before storing
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty(as));
      if (warp == 0) TC_TS(1, tc, 1);   // drained
      // bias (0 when the node has none: add_bias, convolution_op.rs:705), folded Add node (add_op.rs:75: a second
      // rounding, as upstream), Relu (relu_op.rs:31-33) on the four channels m .. m+3 of a slab chunk
      auto finish = [&](float4& v, const float4& bb, const float4& cc) {
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        if (HAS_ADD) { v.x += cc.x; v.y += cc.y; v.z += cc.z; v.w += cc.w; }
        if (do_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      };
      auto chan_consts = [&](int m, float4& bb, float4& cc) {   // m < Mpad: the shared-memory copies are zero padded
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w) : "r"(sbias + 4u * (uint32_t)m));
        if (HAS_ADD) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cc.x), "=f"(cc.y), "=f"(cc.z), "=f"(cc.w) : "r"(sadd + 4u * (uint32_t)m));
      };
      const int row0 = p0 + quarter * 32;                 // first pixel of this warp's 32 rows
      const int rows_valid = min(32, p_end - row0);        // <= 0 for a quarter past the end (of the tile's pixels)
      const int mbase = m0 + g_begin * 16;
      if (TC_DBG(4)) {
      } else if (fast_store) {
        // chunk columns in power-of-two blocks (16, 8 or 4 wide; 12 = 8 + 4): a lane keeps its chunk column and walks
        // down the rows, so an iteration is a bounds test, the swizzle, LDS.128, STG.128 and two pointer adds
        int cdone = 0;
        while (cdone < nchunk) {
          const int w = (nchunk - cdone >= 16) ? 16 : (nchunk - cdone >= 8) ? 8 : 4;
          const int lw = (w == 16) ? 4 : (w == 8) ? 3 : 2;
          const int c = cdone + (lane & (w - 1));
          const int step = 32 >> lw;                        // rows per iteration
          int rr = lane >> lw;
          const int m = mbase + c * 4;
          const int rows_ok = (m < a.M) ? rows_valid : 0;
          float4 bb, cc = make_float4(0.f, 0.f, 0.f, 0.f);
          chan_consts(m, bb, cc);   // the lane keeps its channels: one load for all rows
          float* dst = a.y + (long long)(row0 + rr) * a.ldy + m;
          const long long dstep = (long long)step * a.ldy;
          uint32_t src = slab + (uint32_t)(rr * p.slab_pitch) + ((uint32_t)c << 4);
          const uint32_t sstep = (uint32_t)(step * p.slab_pitch);
          // two rows per round: both loads, then both stores (the asm loads keep program order)
          for (int it = 0; it < (1 << lw); it += 2) {
            float4 v0, v1;
            const bool ok0 = rr < rows_ok, ok1 = rr + step < rows_ok;
            if (ok0) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(src ^ ((uint32_t)(rr & 7) << 4)));
            if (ok1) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"((src + sstep) ^ ((uint32_t)((rr + step) & 7) << 4)));
            if (ok0) { finish(v0, bb, cc); *reinterpret_cast<float4*>(dst) = v0; }
            if (ok1) { finish(v1, bb, cc); *reinterpret_cast<float4*>(dst + dstep) = v1; }
            rr += 2 * step; dst += 2 * dstep; src += 2 * sstep;
          }
          cdone += w;
        }
      } else {
        const uint32_t rcp = nchunk ? (65536u + (uint32_t)nchunk - 1u) / (uint32_t)nchunk : 0u;   // L / nchunk for L < 512
        const int total = 32 * nchunk;
        for (int L = lane; L < total; L += 32) {
          const int rr = (int)(((uint32_t)L * rcp) >> 16);
          const int c = L - rr * nchunk;
          const int m = mbase + c * 4;
          if (rr < rows_valid && m < a.M) {
            float4 val;
            const uint32_t addr = slab + (uint32_t)(rr * p.slab_pitch) + (uint32_t)((c ^ (rr & 7)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w) : "r"(addr));
            float4 bb, cc = make_float4(0.f, 0.f, 0.f, 0.f);
            chan_consts(m, bb, cc);
            finish(val, bb, cc);
            float* dst = a.y + (long long)(row0 + rr) * a.ldy + m;
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
      }
      if (warp == 0) TC_TS(1, tc, 2);   // stored
      __syncwarp();   // the slab is rewritten by the next tile's drain
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// Weight preparation: w [M][ldw] (K valid floats per row) -> out [2*Mpad][Kpad]: tf32 hi rows, then lo rows.
// Position 16g + j of an output row holds k = 16g + perm(j), the order in which tcgen05.st.16x256b lays the
// activations' 16-float groups out in tensor-memory columns (see tmem_st_16x256b_x2).
__global__ void tc_split_weights_kernel(const float* __restrict__ w, int M, int K, int ldw, float* __restrict__ out, int Mpad, int Kpad, int natural_k) {
  const long long total = (long long)Mpad * Kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / Kpad), kp = (int)(i - (long long)m * Kpad);
    const int j = kp & 15;
    const int k = natural_k ? kp : (kp & ~15) + (j < 8 ? 4 * (j >> 1) + (j & 1) : 4 * ((j - 8) >> 1) + 2 + (j & 1));
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
  if (a.K > 12288 || a.C >= 65536) return B200_EUNSUPPORTED;                                  // k decode table: 128 B per k-block in shared memory
  if ((long long)(a.KH + 1) * a.W * a.ldx >= (1ll << 31)) return B200_EUNSUPPORTED;           // 32-bit tap offsets
  if (a.pool) {
    // fused MaxPool 3x3 / 2: pointwise stride-1 convolution on the pooled map, one pooled row (<= 128 pixels) per tile,
    // one channel tile of <= 64 filters (the 8 + 16 role layout), window-row boxes of <= 256 pixels
    if (a.KH != 1 || a.KW != 1 || a.sh != 1 || a.sw != 1 || a.pt != 0 || a.pl != 0) return B200_EUNSUPPORTED;
    if (a.Wo < 1 || a.Wo > BM || a.Ho < 1 || 2 * a.Wo + 1 > 256 || a.M > 64 || a.pool_pt < 0 || a.pool_pl < 0 || a.pool_pt > 2 || a.pool_pl > 2) return B200_EUNSUPPORTED;
    if ((long long)a.N * a.Ho >= (1ll << 31) - 1) return B200_EUNSUPPORTED;
    // two raw slots (three window-row boxes each) + two weight stages + slabs, constants, barriers, tables must fit
    const int rowbox = ((2 * a.Wo + 1) * 128 + 1023) & ~1023, bn = pick_bn(a.M);
    if (2 * 3 * rowbox + 2 * (2 * bn * BK * 4) + 48 * 1024 > SMEM_MAX) return B200_EUNSUPPORTED;
  }
  const long long P = (long long)a.N * a.Ho * a.Wo, in_pix = (long long)a.N * a.H * a.W;
  if (P >= (1ll << 31) - BM || (in_pix + (long long)a.W * 16) * a.ldx >= (1ll << 31)) return B200_EUNSUPPORTED;  // 32-bit element offsets in the producer
  return 0;
}

int tc_prepare_weights(const float* w_dev, int M, int K, cudaStream_t st, std::shared_ptr<TcWeights>* out, bool natural_k) {
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
  tc_split_weights_kernel<<<blocks, 256, 0, st>>>(w_dev, M, K, K, t->buf, t->Mpad, t->Kpad, natural_k ? 1 : 0);
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
  p.reverse = a.reverse ? 1 : 0;
  p.zero = 0;
  if (a.pool && (tc_supported(a) != 0 || p.n_tiles_n != 1)) B200_FAIL(B200_EUNSUPPORTED, "tcgen05 conv: this shape cannot take a fused MaxPool (Wo=%d M=%d)", a.Wo, a.M);
  // pooled mode: as many pooled rows per tile as fit the 128 MMA rows and two raw slots of 2 R + 1 window-row boxes in
  // shared memory (R = 1 fits by tc_supported), evened out over the image (13 rows: 7 + 6, not 9 + 4)
  p.pool_rows = 1; p.pool_tpi = 1;
  p.rowbox = a.pool ? (((2 * a.Wo + 1) * 128 + 1023) & ~1023) : 0;
  if (a.pool) {
    // The launch is bound by the input stream, and what keeps HBM busy is bytes in flight: at least four raw slots
    // (each k-block of a tile is one slot of 2 R + 1 boxes) matter more than full MMA tiles.  Measured (batch 256):
    // pool1 -> fire2 squeeze R = 1 (four 42 KB slots) 0.206 ms, R = 2 (two 70 KB slots) 0.239; pool after fire4 ->
    // fire5 squeeze R = 2 (five 35 KB slots) 0.151, R = 1 0.221, R = 4 (two 63 KB slots) 0.181; pool after fire8 ->
    // fire9 squeeze R = 5 0.095, R = 7 0.116, R = 2 0.137 (MaxPool + Conv launches: 0.294 / 0.198 / 0.113)
    const int room = SMEM_MAX - 2 * (2 * p.BN * BK * 4) - 48 * 1024;
    int R = BM / a.Wo;
    while (R > 1 && 4 * (2 * R + 1) * p.rowbox > room) --R;
    static const int force_rows = [] { const char* e = getenv("B200_TC_POOL_ROWS"); return e ? atoi(e) : 0; }();   // experiments only
    if (force_rows > 0 && force_rows * a.Wo <= BM && 2 * (2 * force_rows + 1) * p.rowbox <= room) R = force_rows;
    if (R > a.Ho) R = a.Ho;
    p.pool_tpi = (a.Ho + R - 1) / R;
    p.pool_rows = (a.Ho + p.pool_tpi - 1) / p.pool_tpi;
  }
  p.tile_rows = a.pool ? p.pool_rows * a.Wo : BM;
  const long long tiles = a.pool ? (long long)a.N * p.pool_tpi : ((P + BM - 1) / BM) * p.n_tiles_n;
  if (tiles >= (1ll << 31)) B200_FAIL(B200_EUNSUPPORTED, "too many tiles");
  p.total_tiles = (int)tiles;
  // tensor memory: accumulator stages {main | correction} x BN, then A stages of 64 columns
  static const int force_nacc = [] { const char* e = getenv("B200_TC_NACC"); return e ? atoi(e) : 0; }();   // experiments only
  // two accumulator stages only when four A stages still fit beside them (BN <= 64): measured on BN = 96 (conv1,
  // fire6/7 expand3x3), four A stages + one accumulator stage beat two + two by 11-15 %
  p.nacc = (4 * p.BN + MAX_A_STAGES * A_STAGE_COLS <= 512) ? 2 : 1;
  // BN > 64: one merged accumulator per stage, so two stages of BN columns + four A stages fit (2*128 + 4*64 = 512)
  static const int force_merged = [] { const char* e = getenv("B200_TC_MERGED"); return e ? atoi(e) : -1; }();   // experiments only
  // The tensor core truncates when it adds into the fp32 accumulator (measured: error grows linearly with the number
  // of accumulating instructions, ~0.5 ulp of the accumulator each), and merging triples the additions into the large
  // accumulator.  So merge only while 3*K/8 additions stay near what the longest unmerged reduction of the model
  // performs (K = 576: 72): K <= 288 -> <= 108.  Longer reductions amortise the exposed drain anyway.
  p.merged = (p.BN > 64 && a.K <= MERGED_MAX_K) ? 1 : 0;
  if (force_merged == 0 || force_merged == 1) p.merged = force_merged;
  const int acc_cols = p.merged ? p.BN : 2 * p.BN;
  if (p.merged) p.nacc = 2;
  if (force_nacc == 1 || force_nacc == 2) p.nacc = force_nacc;
  if (p.nacc * acc_cols + 2 * A_STAGE_COLS > 512) p.nacc = 1;
  p.SA = (512 - p.nacc * acc_cols) / A_STAGE_COLS;
  if (p.SA > MAX_A_STAGES) p.SA = MAX_A_STAGES;
  p.SA &= ~1;   // the two producer sets alternate k-blocks: even stage counts keep a set on its own stages
  // shared memory: weight ring, raw A ring (pointwise mode), epilogue slabs, per-channel constants, barriers
  const int stage_bytes = 2 * p.BN * BK * 4;
  // Pointwise layers (1x1, stride 1, no padding): im2col row p IS input pixel p, so the A operand is a plain 2-D
  // matrix [P][C] and TMA can stream it; these layers are HBM-bound and want many bytes in flight.
  static const int no_atma = [] { const char* e = getenv("B200_TC_NO_ATMA"); return e ? atoi(e) : 0; }();
  p.raw_slot = a.pool ? (2 * p.pool_rows + 1) * p.rowbox : A_TILE_BYTES;
  p.a_tma = a.pool ? 1 : ((!no_atma && a.KH == 1 && a.KW == 1 && a.sh == 1 && a.sw == 1 && a.pt == 0 && a.pl == 0 && a.H == a.Ho && a.W == a.Wo) ? 1 : 0);
  // Role layout: pointwise layers with wide tiles (BN > 64: one accumulator stage, so the drain is exposed, and for
  // the expand1x1 layers a 128 x BN tile per ~1000 clk of MMAs) run with 16 epilogue warps and 8 converter warps --
  // measured: conv10 0.210 -> 0.175 ms, expand1x1 3-6 % faster; squeeze layers (BN <= 64, HBM-bound) and BN = 64
  // expands are better off with 8 + 16, like everything that gathers.
  static const int force_epi16 = [] { const char* e = getenv("B200_TC_EPI16"); return e ? atoi(e) : -1; }();   // experiments only
  // Half-stage A ring: gather layers with 64 < BN <= 96 (conv1, fire6 / fire7 expand3x3).  Two {main | correction} stages
  // (4 * BN columns) leave 128 columns for A: as two full stages that measured 5-13 % slower than one accumulator stage +
  // four A stages (a set must refill its only stage while the MMA warp consumes the other set's); as FOUR half-stages of
  // {hi 16 | lo 16} columns, one per (producer set, k half), a half is released after two k-steps and its four warps have
  // three half-steps to refill it -- and the drain of tile i runs under the MMAs of tile i + 1.  Measured (tc_bench, one
  // call): conv1 (7 k-blocks) 0.517 -> 0.497 ms, fire6 expand3x3 (14 k-blocks: the drain is 13 % of a tile) 0.130 -> 0.131:
  // short reductions only.
  static const int no_halfa = [] { const char* e = getenv("B200_TC_NO_HALFA"); return e ? atoi(e) : 0; }();   // experiments only
  p.half_a = 0;
  if (!no_halfa && !p.a_tma && !p.merged && p.nacc == 1 && force_nacc == 0 && 4 * p.BN + 4 * 32 <= 512 && p.nkb <= 8) {
    p.half_a = 1; p.nacc = 2; p.SA = 4;
  }
  bool epi16 = p.a_tma && p.BN > 64 && !a.pool;
  if (force_epi16 == 0) epi16 = false;
  if (force_epi16 == 1) epi16 = p.a_tma != 0 && !a.pool;
  const int n_epi = epi16 ? 16 : 8;
  const int groups_per_warp = ((p.BN >> 4) + n_epi / 4 - 1) / (n_epi / 4);
  p.slab_pitch = groups_per_warp <= 2 ? 128 : 256;
  int fixed = 1024 + n_epi * 32 * p.slab_pitch + 8 * p.Mpad + 8 * (2 * MAX_STAGES + 2 * MAX_A_STAGES + 5 + 2 * MAX_RAW + 8) + 16 + 4096;   // ... barriers, row tables
  p.R = 0;
  if (!p.a_tma) fixed += 128 * p.nkb;   // gather mode: the per-CTA k decode table
  int S = (SMEM_MAX - fixed) / stage_bytes;
  if (S > MAX_STAGES) S = MAX_STAGES;
  if (S > p.nkb + 1) S = p.nkb + 1;   // a ring deeper than a tile's k-blocks (+1 for the next tile) buys nothing
  if (S < 2) S = 2;
  if (p.a_tma) {
    if (S > 3) S = 3;
    if (a.pool && S > 2) S = 2;   // the raw slots are the large ones here; the whole weight matrix is a few stages anyway
    int R = (SMEM_MAX - fixed - S * stage_bytes) / p.raw_slot;
    if (R > MAX_RAW) R = MAX_RAW;
    if (!epi16) R &= ~1;   // same for the raw ring
    if (R < 2) B200_FAIL(B200_EUNSUPPORTED, "tcgen05 conv: not enough shared memory for the raw A ring (BN=%d)", p.BN);
    p.R = R;
  }
  if ((size_t)S * stage_bytes + fixed > (size_t)SMEM_MAX) B200_FAIL(B200_EUNSUPPORTED, "tcgen05 conv: not enough shared memory for 2 weight stages (BN=%d)", p.BN);
  p.S = S;
  const size_t smem = (size_t)S * stage_bytes + (size_t)p.R * p.raw_slot + fixed;

  p.vec_store = (a.ldy % 4 == 0 && (((uintptr_t)a.y) & 15) == 0) ? 1 : 0;
  {
    static const int dbg = [] {
      const char* e = getenv("B200_TC_DEBUG");
      const int v = e ? atoi(e) : 0;
      if (v && !B200_TC_DEBUG_MASKS) fprintf(stderr, "b200rt: B200_TC_DEBUG=%d ignored: this build has no role masks (make debug, B200RT_LIB=.../libb200rt_dbg.so)\n", v);
      return v;
    }();
    p.debug = dbg;
  }
  p.m64NT = p.n_tiles_n == 1 ? 0ull : ~0ull / (unsigned long long)p.n_tiles_n + 1ull;
  p.m64Wo = a.Wo == 1 ? 0ull : ~0ull / (unsigned long long)a.Wo + 1ull;
  p.m64Ho = a.Ho == 1 ? 0ull : ~0ull / (unsigned long long)a.Ho + 1ull;
  p.skip_cols = 0; p.skip_kb = 0;
  if (a.skip_m > 0 && a.skip_m % 16 == 0 && a.skip_m + 16 <= p.BN && p.n_tiles_n == 1 && !p.merged && a.skip_kb >= 1 && a.skip_kb < p.nkb) {
    p.skip_cols = a.skip_m; p.skip_kb = a.skip_kb;
  }
  if (p.half_a) { p.skip_cols = 0; p.skip_kb = 0; }   // the half-stage MMA loop issues every column (the skipped ones are zeros anyway)
  p.nopad = (a.pt == 0 && a.pl == 0 && (long long)(a.Ho - 1) * a.sh + a.KH <= a.H && (long long)(a.Wo - 1) * a.sw + a.KW <= a.W) ? 1 : 0;
  p.magicC = a.C == 1 ? 0u : (uint32_t)(((1ull << 32) + a.C - 1) / a.C);
  p.magicKW = a.KW == 1 ? 0u : (uint32_t)(((1ull << 32) + a.KW - 1) / a.KW);
  p.magicWo = a.Wo == 1 ? 0u : (uint32_t)(((1ull << 32) + a.Wo - 1) / a.Wo);
  p.magicHo = a.Ho == 1 ? 0u : (uint32_t)(((1ull << 32) + a.Ho - 1) / a.Ho);

  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  static int sm_count[64] = {0};
  if (dev < 64 && !attr_set[dev]) {
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, false, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, false, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    B200_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    attr_set[dev] = true;
  }
  const int sms = (dev < 64 && sm_count[dev] > 0) ? sm_count[dev] : 148;
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  CUtensorMap tmapA;
  memset(&tmapA, 0, sizeof(tmapA));
  if (a.pool) {
    // the tensor BEFORE the pool as [N][H][W][C]; box = 32 channels x (2 Wo + 1) pixels of one row.  Coordinates outside
    // the tensor (the pool's zero padding, channels past C in the last k-block) are filled with zeros.
    EncodeTiledFn enc = get_encode_fn();
    cuuint64_t gdim[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
    cuuint64_t gstride[3] = {(cuuint64_t)a.ldx * sizeof(float), (cuuint64_t)a.W * a.ldx * sizeof(float), (cuuint64_t)a.H * a.W * a.ldx * sizeof(float)};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(2 * a.Wo + 1), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc ? enc(&tmapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a.x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                     : CUDA_ERROR_NOT_SUPPORTED;
    if (r != CUDA_SUCCESS) B200_FAIL(B200_ECUDA, "cuTensorMapEncodeTiled (pooled activation map) failed with %d (C=%d W=%d H=%d N=%d ldx=%d)", (int)r, a.C, a.W, a.H, a.N, a.ldx);
  } else if (p.a_tma) {
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
  const bool nopad = p.nopad && !p.a_tma;
  void (*kernel)(const CUtensorMap, const CUtensorMap, const TcParams);
  if (a.pool) kernel = a.chan_add ? conv_tc_kernel<true, false, false, true, false> : conv_tc_kernel<false, false, false, true, false>;
  else if (epi16) kernel = a.chan_add ? conv_tc_kernel<true, true, false, false, false> : conv_tc_kernel<false, true, false, false, false>;
  else if (p.half_a && nopad) kernel = a.chan_add ? conv_tc_kernel<true, false, true, false, true> : conv_tc_kernel<false, false, true, false, true>;
  else if (p.half_a) kernel = a.chan_add ? conv_tc_kernel<true, false, false, false, true> : conv_tc_kernel<false, false, false, false, true>;
  else if (nopad) kernel = a.chan_add ? conv_tc_kernel<true, false, true, false, false> : conv_tc_kernel<false, false, true, false, false>;
  else kernel = a.chan_add ? conv_tc_kernel<true, false, false, false, false> : conv_tc_kernel<false, false, false, false, false>;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 1 : 0;
  B200_CUDA(cudaLaunchKernelEx(&cfg, kernel, w.tmap, tmapA, p));
  B200_CUDA(cudaGetLastError());
#if B200_TC_DEBUG_MASKS
  if (p.debug & 1024) {   // tile timeline of CTA 0 (clk relative to the first stamp)
    B200_CUDA(cudaStreamSynchronize(st));
    static unsigned long long h[3][96][4];
    B200_CUDA(cudaMemcpyFromSymbol(h, g_ts, sizeof(h)));
    const unsigned long long t0 = h[0][0][3];
    const int nt = std::min(96, (p.total_tiles + grid - 1) / grid);
    fprintf(stderr, "tile timeline BN=%d K=%d tiles/CTA=%d: tile | MMA arrive, acc free, operands, last issue | EPI arrive, acc full, drained, stored\n", p.BN, a.K, nt);
    for (int t = 0; t < nt; ++t)
      fprintf(stderr, "%3d | %8lld %8lld %8lld %8lld | %8lld %8lld %8lld %8lld\n", t, (long long)(h[0][t][3] - t0), (long long)(h[0][t][0] - t0), (long long)(h[0][t][1] - t0),
              (long long)(h[0][t][2] - t0), (long long)(h[1][t][3] - t0), (long long)(h[1][t][0] - t0), (long long)(h[1][t][1] - t0), (long long)(h[1][t][2] - t0));
    static unsigned long long zero[3][96][4];
    B200_CUDA(cudaMemcpyToSymbol(g_ts, zero, sizeof(zero)));
  }
#endif
  return 0;
}

}  // namespace b200
