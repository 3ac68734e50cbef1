// Fused path for the small-kernel regime (BASELINE.json config 5: MNIST-8, batch 65,536): the whole 12-node graph in
// TWO launches that never write an un-pooled activation to HBM.
//
//   mnist8_stem_kernel   Conv(1->8, 5x5, SAME_UPPER) + Add[8,1,1] + Relu + MaxPool 2x2/2            CUDA cores (C = 1:
//                        (Convolution28, Plus30, ReLU32, Pooling66)                                  K = 25, N = 8 is far
//                        convolution_op.rs:224-517 (C == 1 branch :416-419), add_op.rs:75,           below one UMMA tile)
//                        relu_op.rs:31-33, max_pool_op.rs:157-360
//   mnist8_head_kernel   Conv(8->16, 5x5, SAME_UPPER) + Add[16,1,1] + Relu + MaxPool 3x3/3           tcgen05 3xTF32,
//                        + Reshape + MatMul(256x10) + Add[1,10]                                      K = 200, N = 16
//                        (Convolution110, Plus112, ReLU114, Pooling160, Times212_reshape0, Times212, Plus214)
//                        + reshape_op.rs:66-92, mul_op.rs:23, add_op.rs:84
//
// Between the two: the pooled stem output in a zero-haloed channels-last layout [N][18][18][8] (+ 16 bytes per image, see
// IMG_FLOATS), 10.4 KB per image, written once and read once.  The regime is instruction-issue-bound, not HBM-bound
// (3.2 KB in, 40 B out per image), so the design minimises issued instructions:
//   * stem: one thread = one pooled pixel = a 6x6 input patch in registers, 2x2 conv pixels x 8 channels = 32
//     accumulators as packed pairs (fma.rn.f32x2), weights as broadcast 128-bit shared loads; 8 images per block
//     iteration = 7 full passes of 224 threads.
//   * head: conv2 as an implicit GEMM whose 128-row tiles are (8 images x 16 pool windows) at ONE of the 9 positions
//     inside a 3x3 pool window.  MaxPool 3x3/3 is then an element-wise max over the 9 tiles' accumulators in the epilogue
//     threads' registers -- no cross-lane traffic, no pooled-out pixels computed (rows / columns 12, 13 of the 14x14 map
//     never reach the output: max_pool_op.rs:215-246 floors) -- and the MatMul is 160 FMAs per thread plus a shuffle /
//     shared-memory reduction.  Tiles at neighbouring window positions read the same pooled-stem pixels under
//     different taps, so the A operand is produced ONCE per input offset (7 x 7 blocks of 8 channels per group, not
//     9 tiles x 25 taps): 8 producer warps read the group's 83 KB from shared memory (one bulk copy per 8 images), split
//     hi / lo in registers and write their own tensor-memory rows (tcgen05.st.32x32b: natural K order, weights prepared
//     with natural_k).  MMAs: per tap A_hi x [B_hi; B_lo] (N = 32) and A_lo x B_hi (N = 16): 25.5 + 17.7 clk measured
//     (profiles/r2_mma_issue_rates.txt); the kernel is bound by that issue stream.
// Numerics: 3xTF32 with {main | correction} accumulators like conv_tc.cu; bias / Add, Relu and the max commute
// (x -> fl(x + b) and Relu are monotone), so max-then-add equals the reference's add-then-max bit for bit.
#include <cstring>

#include "internal.h"
#include "tc_common.h"

namespace b200 {
namespace {

#include "tc_ptx.cuh"

// ------------------------------------------------------------------------------------------------ geometry (MNIST-8)
constexpr int IN_HW = 28;                 // input 1 x 28 x 28
constexpr int C1 = 8;                     // stem channels
constexpr int P1_HW = 14;                 // pooled stem map
constexpr int PADW = 18;                  // 14 + 2 * 2 halo
constexpr int IMG_FLOATS = PADW * PADW * C1 + 4;    // 2,596 floats = 10,384 B = 16 B mod 128: the eight rows a quarter-warp of
                                                     // the head's producers reads in one shared-memory phase are the eight
                                                     // images of a group at the SAME pixel, so they fall in eight different
                                                     // 16-byte bank groups
constexpr int G = 8;                      // images per group = one 128-row tile per pool-window position
constexpr int C2 = 16;                    // head conv channels
constexpr int K2 = 200;                   // 5 * 5 * 8
constexpr int NKB = 7;                    // k-blocks of 32 floats (4 taps) per tile
constexpr int NOUT = 10;

// ------------------------------------------------------------------------------------------------ stem (CUDA cores)
constexpr int STEM_THREADS = 224;         // 8 images x 196 pooled pixels = 7 x 224
constexpr int TILE_W = 32;                // 28 + 2 * 2 halo, padded rows of 32 floats

struct StemArgs {
  const float* x;        // [N][28*28]
  Mnist8StemConsts k;    // weights [25][8] (tap-major), Conv bias [8] (convolution_op.rs:705), folded Add [8] (add_op.rs:75):
                         // kernel parameters = constant bank, so the 50 weight loads per task cost no shared-memory bandwidth
  float* p1;             // [N][IMG_FLOATS], halo and tail pre-zeroed, interior written here
  int N;
  int* nonfinite;        // set to 1 when the input holds an Inf / NaN (or null)
};

// One stem task = one pooled pixel: 2 x 2 conv pixels x 8 channels from the 6 x 6 patch at t0 (haloed tile, row pitch 32).
// 32 accumulators as 16 channel pairs, one packed FFMA2 (fma.rn.f32x2: two IEEE fp32 FMAs per lane, each bit-identical to
// fmaf) per pair and tap.  Only two patch rows are live at a time (kernel row r reads patch rows r and r + 1; the next row
// is loaded one kernel row ahead) and the weights of tap t + 1 are loaded while tap t is computed: with the whole 6 x 6
// patch in registers (36 + 32 accumulators under an 80-register cap) the compiler could not hoist any load, and a team of
// few warps paid the shared-memory latency on every tap (ncu, one-launch kernel: 7.5 clk per instruction per stem warp).
__device__ __forceinline__ void stem_task(const float* __restrict__ t0, const Mnist8StemConsts& k, float (&o)[8], uint32_t& mx) {
  const float (*ws)[8] = k.w;
  const float* sb = k.bias;
  const float* sa = k.add;
  auto load_row = [&](float (&row)[6], int r) {
#pragma unroll
    for (int c = 0; c < 6; c += 2) {
      const float2 v = *reinterpret_cast<const float2*>(t0 + r * TILE_W + c);
      row[c] = v.x; row[c + 1] = v.y;
    }
  };
  float ra[6], rb[6];
  load_row(ra, 0);
  load_row(rb, 1);
  float4 w0 = *reinterpret_cast<const float4*>(&ws[0][0]), w1 = *reinterpret_cast<const float4*>(&ws[0][4]);
  float2 acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    float rn[6];
    if (r < 4) load_row(rn, r + 2);
    // finite guard: every input pixel is the own position of exactly one conv pixel -- patch rows 2, 3, columns 2, 3
    if (r == 1 || r == 2) mx = max(mx, max(__float_as_uint(rb[2]) & 0x7fffffffu, __float_as_uint(rb[3]) & 0x7fffffffu));
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const int t = r * 5 + s;
      float4 w0n = w0, w1n = w1;
      if (t < 24) { w0n = *reinterpret_cast<const float4*>(&ws[t + 1][0]); w1n = *reinterpret_cast<const float4*>(&ws[t + 1][4]); }
      const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float xv = (p >> 1) ? rb[(p & 1) + s] : ra[(p & 1) + s];
        const float2 xx = make_float2(xv, xv);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[p][c] = __ffma2_rn(xx, wv[c], acc[p][c]);
      }
      w0 = w0n; w1 = w1n;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) { ra[c] = rb[c]; rb[c] = rn[c]; }
  }
  // (conv + bias) + add, Relu, max over the 2 x 2 window (fold from -FLT_MAX like max_pool_op.rs:337)
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float m = -3.402823466e+38f;
#pragma unroll
    for (int p = 0; p < 4; ++p) m = fmaxf(m, fmaxf((((c & 1) ? acc[p][c >> 1].y : acc[p][c >> 1].x) + sb[c]) + sa[c], 0.f));
    o[c] = m;
  }
}

constexpr int STEM_TILE_FLOATS = G * TILE_W * TILE_W;                       // 8 zero-haloed 32 x 32 tiles
constexpr int STEM_SMEM = 2 * STEM_TILE_FLOATS * 4;                         // two tile buffers

__global__ void __launch_bounds__(STEM_THREADS, 3) mnist8_stem_kernel(const StemArgs a) {
  extern __shared__ __align__(16) float stem_smem[];
  float* const tiles = stem_smem;                                            // [2][G][32 * 32]
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * STEM_TILE_FLOATS; i += STEM_THREADS) tiles[i] = 0.f;      // halos stay zero: only interiors are rewritten
  __syncthreads();
  const int groups = (a.N + G - 1) / G;
  // Stage 8 images into a tile buffer with cp.async (8-byte pieces: the interior of a haloed row starts at column 2):
  // the copy of group i+1 is in flight while group i is computed (the ncu capture of the single-buffered version had
  // 31 % of its stall samples in the load / barrier phase).
  auto stage = [&](int g, int buf) {
    const int img0 = g * G;
    const int nimg = min(G, a.N - img0);
    const float* src = a.x + (size_t)img0 * (IN_HW * IN_HW);
    float* dstb = tiles + buf * STEM_TILE_FLOATS;
    for (int i = tid; i < nimg * 392; i += STEM_THREADS) {     // 392 float2 per image, 14 per row
      const int im = i / 392, q = i - im * 392, r = q / 14, c2 = q - r * 14;
      const uint32_t d = smem_u32(dstb + im * (TILE_W * TILE_W) + (r + 2) * TILE_W + 2 + c2 * 2);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + 2 * i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int cur = 0;
  if ((int)blockIdx.x < groups) stage(blockIdx.x, 0);
  for (int g = blockIdx.x; g < groups; g += gridDim.x, cur ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // buffer `cur` is complete for everyone, and everyone has finished computing from the other buffer
    if (g + (int)gridDim.x < groups) stage(g + gridDim.x, cur ^ 1);
    const int img0 = g * G;
    const int nimg = min(G, a.N - img0);
    const float (*tile)[TILE_W * TILE_W] = reinterpret_cast<const float (*)[TILE_W * TILE_W]>(tiles + cur * STEM_TILE_FLOATS);
    uint32_t mx = 0;   // finite guard: every input pixel is the own position of exactly one conv pixel, checked there
    // ---- 7 passes: task = (image, pooled pixel)
#pragma unroll 1
    for (int pass = 0; pass < 7; ++pass) {
      const int task = pass * STEM_THREADS + tid;
      const int im = task / 196, pp = task - im * 196, ph = pp / P1_HW, pw = pp - ph * P1_HW;
      if (im >= nimg) continue;
      // 6 x 6 patch: rows 2ph .. 2ph+5, columns 2pw .. 2pw+5 of the haloed tile
      const float* t0 = &tile[im][(2 * ph) * TILE_W + 2 * pw];
      float o[8];
      stem_task(t0, a.k, o, mx);
      float* dst = a.p1 + (size_t)(img0 + im) * IMG_FLOATS + ((ph + 2) * PADW + (pw + 2)) * C1;
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (a.nonfinite && mx >= 0x7f800000u) *a.nonfinite = 1;
  }
}

// ------------------------------------------------------------------------------------------------ head (tcgen05)
// Tiles: row rho = image rho % 8, pool window rho / 8 (4 x 4 windows of 3 x 3 conv outputs); tile (jh, jw) holds the conv
// output at position (jh, jw) of every window.  Tap (r, s) of tile (jh, jw) reads the pooled-stem pixel at offset
// (jh + r, jw + s) from the window's origin: the SAME 128 x 8-channel block of the A operand serves up to nine
// (tile, tap) pairs.  So the A operand is produced per INPUT-ROW offset r' = 0..6 as seven blocks s' = 0..6 (128 rows x 8
// channels, hi and lo), 49 blocks per group instead of 9 tiles x 25 taps = 225: the producers' work drops 4.6x, and the
// MMA warp walks r' and issues, for every tile with 0 <= r' - jh <= 4, the five taps of that kernel row.  All nine
// accumulators of a group are live at once (9 x 32 columns), the A blocks of two input rows are double-buffered
// (2 x 112 columns): 512 columns exactly.
constexpr int HEAD_WARPS = 16;
constexpr int HEAD_THREADS = HEAD_WARPS * 32;   // 512
constexpr int EPI_WARPS = 4;                    // warps 0-3: TMEM lane quarter = warp
constexpr int LOAD_WARP = 4, MMA_WARP = 5;      // warps 6, 7 idle (producers must start at a multiple of 4)
constexpr int PROD_WARP0 = 8;
constexpr int NSETS = 2, SET_WARPS = 4;         // 8 producer warps: set j fills the A buffer of input rows j, j + 2, ... (counted over groups)
constexpr int NTILES = 9;                       // window positions = accumulators, 32 columns each: main 16 | correction 16
constexpr int A_COL0 = NTILES * 32;             // 288
constexpr int A_BUF_COLS = 112;                 // 7 blocks x 8 columns hi | 7 x 8 lo
constexpr int NROWS = 7, NBLK = 7;              // input-row offsets r' and column offsets s' per group
constexpr uint32_t IMG_BYTES = IMG_FLOATS * 4;  // 10,384
constexpr uint32_t GROUP_BYTES = G * IMG_BYTES; // 83,072
constexpr uint32_t B_KB_BYTES = 2 * C2 * 128;   // one k-block (4 taps) of weights: [B_hi 16 rows | B_lo 16 rows] x 128 B
constexpr uint32_t SM_B = 0;                                    // 7 x 4 KB
constexpr uint32_t SM_P1 = SM_B + NKB * B_KB_BYTES;             // 2 x 83,072 B
constexpr uint32_t SM_WM = SM_P1 + 2 * GROUP_BYTES;             // matmul weights [10][256] + bias [10] (+ pad)
constexpr uint32_t SM_ADD = SM_WM + (NOUT * 256 + 16) * 4;      // conv bias [16] | add [16]
constexpr uint32_t SM_PART = SM_ADD + 32 * 4;                   // [2][4 warps][8 img][10] partial logits
constexpr uint32_t SM_BARS = SM_PART + 2 * 4 * 8 * NOUT * 4;
constexpr int NBARS = 1 + 2 + 2 + NSETS + NSETS + NTILES + NTILES;   // b_full | p1_full[2] | p1_empty[2] | full_a | empty_a | tmem_full | tmem_empty
constexpr uint32_t SM_SLOT = SM_BARS + NBARS * 8;
constexpr uint32_t HEAD_SMEM = SM_SLOT + 16 + 1024;             // + alignment slack
static_assert(A_COL0 + 2 * A_BUF_COLS == 512, "tensor memory map");
static_assert(IMG_BYTES % 128 == 16 && IMG_BYTES % 16 == 0, "image pitch: 16 bytes mod 128");

struct HeadArgs {
  const float* p1;       // [N][IMG_FLOATS]
  const float* bias2;    // [16] or null
  const float* add2;     // [16] or null
  const float* wm;       // [10][256], k = window * 16 + c (the activation's physical flatten order, see do_matmul)
  const float* bm;       // [10] or null
  float* out;            // [N][10]
  int N;
};

// One row per lane: 8 consecutive 32-bit columns of the lane's own TMEM row (natural K order: column j = register j)
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, float r0, float r1, float r2, float r3, float r4, float r5, float r6, float r7) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(r0), "f"(r1), "f"(r2),
               "f"(r3), "f"(r4), "f"(r5), "f"(r6), "f"(r7)
               : "memory");
}

__global__ void __launch_bounds__(HEAD_THREADS, 1) mnist8_head_kernel(const __grid_constant__ CUtensorMap tmapB, const HeadArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (sbase - smem_u32(smem_raw));   // generic pointer to the same place
  const uint32_t bars = sbase + SM_BARS;
  const uint32_t b_full = bars;
  auto p1_full = [&](int b) { return bars + 8u * (1 + b); };
  auto p1_empty = [&](int b) { return bars + 8u * (3 + b); };
  auto full_a = [&](int s) { return bars + 8u * (5 + s); };
  auto empty_a = [&](int s) { return bars + 8u * (5 + NSETS + s); };
  auto tmem_full = [&](int t) { return bars + 8u * (5 + 2 * NSETS + t); };
  auto tmem_empty = [&](int t) { return bars + 8u * (5 + 2 * NSETS + NTILES + t); };
  const uint32_t tmem_slot = sbase + SM_SLOT;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int groups = (a.N + G - 1) / G;
  int my_groups = 0;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) ++my_groups;

  // ---- per-CTA constants
  float* const wm_s = reinterpret_cast<float*>(gbase + SM_WM);
  float* const add_s = reinterpret_cast<float*>(gbase + SM_ADD);
  for (int i = threadIdx.x; i < NOUT * 256; i += HEAD_THREADS) wm_s[i] = __ldg(a.wm + i);
  if (threadIdx.x < NOUT) wm_s[NOUT * 256 + threadIdx.x] = a.bm ? __ldg(a.bm + threadIdx.x) : 0.f;
  if (threadIdx.x < 16) { add_s[threadIdx.x] = a.bias2 ? __ldg(a.bias2 + threadIdx.x) : 0.f; add_s[16 + threadIdx.x] = a.add2 ? __ldg(a.add2 + threadIdx.x) : 0.f; }

  if (warp == LOAD_WARP) {
    if (lane == 0) {
      mbar_init(b_full, 1);
      for (int b = 0; b < 2; ++b) { mbar_init(p1_full(b), 1); mbar_init(p1_empty(b), NSETS * SET_WARPS); }
      for (int s = 0; s < NSETS; ++s) { mbar_init(full_a(s), SET_WARPS); mbar_init(empty_a(s), 1); }
      for (int t = 0; t < NTILES; ++t) { mbar_init(tmem_full(t), 1); mbar_init(tmem_empty(t), EPI_WARPS); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  if (warp >= PROD_WARP0) {
    // ================================================================ A producers: 2 sets x 4 warps, a lane = a row
    // Warp w owns TMEM lanes 32 * (w % 4) ..; lane l is row rho = 32 (w % 4) + l: image l % 8, window 4 (w % 4) + l / 8.
    // Per input-row offset r' the thread reads the seven 32-byte pixels (r', 0..6) of its window from the group's buffer
    // (a quarter-warp's eight lanes are the eight images at one pixel: image pitch = 16 bytes mod 128, conflict-free),
    // splits them and writes its own TMEM row: hi block s' -> columns 8 s' .., lo block -> 56 + 8 s' ...
    const int pw_ = warp - PROD_WARP0;
    const int quarter = pw_ & 3, set = pw_ >> 2;
    const int img = lane & 7, win = quarter * 4 + (lane >> 3);
    const uint32_t rowoff = (uint32_t)img * IMG_BYTES + (uint32_t)(((3 * (win >> 2)) * PADW + 3 * (win & 3)) * C1 * 4);
    const uint32_t t_a = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(A_COL0 + set * A_BUF_COLS);
    const int total_rows = my_groups * NROWS;
    uint32_t ph = 0;                    // phase of this set's buffer (full_a / empty_a [set])
    for (int idx = set; idx < total_rows; idx += NSETS) {
      const int gl = idx / NROWS, rr = idx - gl * NROWS;
      // the first row this warp touches in a group waits for the group's bulk copy
      if (idx < NSETS || (idx - NSETS) / NROWS != gl) mbar_wait(p1_full(gl & 1), (uint32_t)(gl >> 1) & 1u);
      const uint32_t src = sbase + SM_P1 + (uint32_t)(gl & 1) * GROUP_BYTES + rowoff + (uint32_t)(rr * PADW * C1 * 4);
      float4 x[NBLK][2];
#pragma unroll
      for (int sp = 0; sp < NBLK; ++sp) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[sp][0].x), "=f"(x[sp][0].y), "=f"(x[sp][0].z), "=f"(x[sp][0].w) : "r"(src + 32u * sp));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[sp][1].x), "=f"(x[sp][1].y), "=f"(x[sp][1].z), "=f"(x[sp][1].w) : "r"(src + 32u * sp + 16u));
      }
      // leaving the group: hand its buffer back once the loads above have delivered (the arrive depends on them)
      const bool last_of_group = idx + NSETS >= total_rows || (idx + NSETS) / NROWS != gl;
      if (last_of_group) {
        const uint32_t dep = (__float_as_uint(x[0][0].w) ^ __float_as_uint(x[3][1].w) ^ __float_as_uint(x[6][1].w)) & (uint32_t)(a.N >> 31);   // always 0, opaque
        __syncwarp();
        if (lane == 0) mbar_arrive(p1_empty(gl & 1) + dep);
      }
      mbar_wait(empty_a(set), ph ^ 1u);   // the MMAs that read this buffer have completed
      tc_fence_after();
#pragma unroll
      for (int sp = 0; sp < NBLK; ++sp) {
        const float4 u = x[sp][0], v = x[sp][1];
        const float h0 = split_hi(u.x), h1 = split_hi(u.y), h2 = split_hi(u.z), h3 = split_hi(u.w);
        const float h4 = split_hi(v.x), h5 = split_hi(v.y), h6 = split_hi(v.z), h7 = split_hi(v.w);
        tmem_st_32x32b_x8(t_a + 8u * sp, h0, h1, h2, h3, h4, h5, h6, h7);
        tmem_st_32x32b_x8(t_a + 56u + 8u * sp, split_lo(u.x, h0), split_lo(u.y, h1), split_lo(u.z, h2), split_lo(u.w, h3),
                          split_lo(v.x, h4), split_lo(v.y, h5), split_lo(v.z, h6), split_lo(v.w, h7));
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_a(set));
      ph ^= 1u;
    }
  } else if (warp == LOAD_WARP) {
    // ================================================================ loader: weights once, then one bulk copy per group
    if (lane == 0) {
      mbar_expect_tx(b_full, NKB * B_KB_BYTES);
      for (int kb = 0; kb < NKB; ++kb) {
        tma_load_2d(sbase + SM_B + (uint32_t)kb * B_KB_BYTES, &tmapB, b_full, kb * 32, 0);               // B_hi rows [0,16)
        tma_load_2d(sbase + SM_B + (uint32_t)kb * B_KB_BYTES + C2 * 128, &tmapB, b_full, kb * 32, C2);   // B_lo rows [16,32)
      }
      int gl = 0;
      for (int g = blockIdx.x; g < groups; g += gridDim.x, ++gl) {
        const int b = gl & 1;
        mbar_wait(p1_empty(b), ((uint32_t)(gl >> 1) & 1u) ^ 1u);
        const int nimg = min(G, a.N - g * G);
        const uint32_t bytes = (uint32_t)nimg * IMG_BYTES;   // rows of images past N keep stale (finite or not) data: never stored
        mbar_expect_tx(p1_full(b), bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + SM_P1 + (uint32_t)b * GROUP_BYTES),
                     "l"(a.p1 + (size_t)g * G * IMG_FLOATS), "r"(bytes), "r"(p1_full(b))
                     : "memory");
      }
    }
  } else if (warp == MMA_WARP) {
    // ================================================================ MMA issuer (whole warp walks, one elected lane issues)
    const uint32_t idesc32 = instr_desc_tf32(2 * C2), idesc16 = instr_desc_tf32(C2);
    const bool leader = elect_one();
    const uint32_t b_lo0 = (((sbase + SM_B) >> 4) & 0x3FFFu) | (1u << 16);
    mbar_wait(b_full, 0);
    int buf = 0;
    uint32_t pha = 0;
    for (int gl = 0; gl < my_groups; ++gl) {
      const uint32_t gph = (uint32_t)gl & 1u;
#pragma unroll 1
      for (int rr = 0; rr < NROWS; ++rr) {
        mbar_wait(full_a(buf), pha);
        tc_fence_after();
        const uint32_t a_hi = tmem_base + (uint32_t)(A_COL0 + buf * A_BUF_COLS), a_lo = a_hi + 56u;
#pragma unroll
        for (int jh = 0; jh < 3; ++jh) {
          const int r = rr - jh;                 // kernel row of tile row jh that reads input row rr
          if (r < 0 || r > 4) continue;          // warp-uniform
          if (r == 0) {                          // first MMAs into these three accumulators in this group
#pragma unroll
            for (int jw = 0; jw < 3; ++jw) mbar_wait(tmem_empty(jh * 3 + jw), gph ^ 1u);
            tc_fence_after();
          }
          if (leader) {
#pragma unroll
            for (int jw = 0; jw < 3; ++jw) {
              const uint32_t d_main = tmem_base + (uint32_t)((jh * 3 + jw) * 32), d_corr = d_main + (uint32_t)C2;
#pragma unroll
              for (int s = 0; s < 5; ++s) {
                const int tap = r * 5 + s;        // weights of tap (r, s): k-block tap / 4, 32-byte step tap % 4 inside its 128-byte rows
                const uint32_t bl = b_lo0 + (uint32_t)(tap >> 2) * (B_KB_BYTES >> 4) + 2u * (uint32_t)(tap & 3);
                umma_tf32_ts(d_main, a_hi + 8u * (uint32_t)(jw + s), sw128_desc(bl), idesc32, (r | s) != 0 ? 1u : 0u);   // hi*hi -> main, hi*lo -> corr
                umma_tf32_ts(d_corr, a_lo + 8u * (uint32_t)(jw + s), sw128_desc(bl), idesc16, 1u);                         // lo*hi -> corr
              }
            }
          }
          __syncwarp();
          if (r == 4 && leader) {                 // last kernel row: these three tiles are complete
#pragma unroll
            for (int jw = 0; jw < 3; ++jw) umma_commit(tmem_full(jh * 3 + jw));
          }
        }
        __syncwarp();
        if (leader) umma_commit(empty_a(buf));
        if (++buf == NSETS) { buf = 0; pha ^= 1u; }
      }
    }
  } else if (warp < EPI_WARPS) {
    // ================================================================ epilogue: max over the 9 tiles, Add, Relu, MatMul
    // Lane = row rho = 32 warp + lane of every tile: image rho % 8 = lane % 8, pool window rho / 8 = 4 warp + lane / 8.
    const int win = warp * 4 + (lane >> 3);
    float* const part = reinterpret_cast<float*>(gbase + SM_PART);
    int gl = 0;
    for (int g = blockIdx.x; g < groups; g += gridDim.x, ++gl) {
      float m[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) m[c] = -3.402823466e+38f;   // fold start of max_pool_op.rs:337
      for (int t = 0; t < NTILES; ++t) {                        // tiles complete in this order (tile row jh after input row jh + 4)
        mbar_wait(tmem_full(t), (uint32_t)gl & 1u);
        tc_fence_after();
        uint32_t acc[16], cor[16];
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 32);
        tmem_ld16(ta, acc);
        tmem_ld16(ta + 16u, cor);
        tmem_ld_wait(acc, cor);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty(t));
#pragma unroll
        for (int c = 0; c < 16; ++c) m[c] = fmaxf(m[c], __uint_as_float(acc[c]) + __uint_as_float(cor[c]));
      }
      // (conv + bias) + add, Relu (monotone: applied after the max), then this thread's share of the 256-long dot products
      float f[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) f[c] = fmaxf((m[c] + add_s[c]) + add_s[16 + c], 0.f);
      float o[NOUT];
#pragma unroll
      for (int n = 0; n < NOUT; ++n) {
        const float4* wr = reinterpret_cast<const float4*>(wm_s + n * 256 + win * 16);
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = wr[q];
          s = fmaf(f[4 * q], w4.x, s); s = fmaf(f[4 * q + 1], w4.y, s); s = fmaf(f[4 * q + 2], w4.z, s); s = fmaf(f[4 * q + 3], w4.w, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        o[n] = s;
      }
      float* pb = part + (gl & 1) * (4 * 8 * NOUT);
      if (lane < 8) {
#pragma unroll
        for (int n = 0; n < NOUT; ++n) pb[(warp * 8 + lane) * NOUT + n] = o[n];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps
      const int tid = warp * 32 + lane;
      if (tid < 8 * NOUT) {
        const int im = tid / NOUT, n = tid - im * NOUT;
        const float s = (pb[(0 * 8 + im) * NOUT + n] + pb[(1 * 8 + im) * NOUT + n]) + (pb[(2 * 8 + im) * NOUT + n] + pb[(3 * 8 + im) * NOUT + n]);
        if (g * G + im < a.N) a.out[(size_t)(g * G + im) * NOUT + n] = s + wm_s[NOUT * 256 + n];   // Plus214, add_op.rs:84
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == LOAD_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------------ one launch: stem || head
// The stem is bound by the fp32 FMA pipe and the head by the tcgen05 issue stream: different pipes.  Run back to back they
// add up (0.47 + 0.35 ms per 65,536 images); inside ONE persistent CTA per SM the seven stem warps compute the pooled
// map of group i+1 on the CUDA cores, straight into a shared-memory group buffer, while the producer / MMA / epilogue
// warps run the head of group i from the other buffer -- the pooled map never exists in HBM, the 12-node graph is one
// launch, and the step approaches max(stem, head).  The head needs window origin (3 ph, 3 pw) + offsets 0..6, i.e. padded
// coordinates 0..15 only, so the group buffers hold 16 x 16 pixels per image (8,208 B: 16 B mod 128) and two of them,
// the 32 KB of haloed input tiles, a 25 KB staging area for the next group's raw images (bulk-copied by warp 12) and the
// weights fit 227 KB.  Roles (21 warps, 672 threads, 80 registers):
//   warps 0-3 epilogue | 4-11 A producers (2 sets x 4) | 12 weights + input bulk copies, TMEM allocation | 13 MMA issuer |
//   14-20 stem (seven warps, seven passes over the group's 49 warp-tasks of 32 pooled pixels).
// What bounds the fused kernel is the stem's FMA stream at roughly half the fp32 FMA peak (10.8 us per group of 8 images
// against 5.0 us of FFMA2 pipe time), next to a head that needs 6.3 us per group: ten stem warps instead of seven change
// nothing, keeping the stem off the MMA issuer's sub-partition (three of four FMA pipes) costs 0.65 -> 0.83 ms, a
// conflict-free patch-load mapping that idles 2 of 16 lanes costs 0.60 -> 0.65 ms.  What did help: the stem's weights as
// kernel parameters (constant bank: 50 fewer shared-memory loads per task, 0.65 -> 0.60 ms; the stand-alone stem kernel
// 0.47 -> 0.42 ms, 78 -> 64 registers).
namespace one {
constexpr int PW = 16;                                       // padded pooled map kept per image: 16 x 16 pixels
constexpr uint32_t IMG_B = (PW * PW * C1 + 4) * 4;           // 8,208 B
constexpr uint32_t GROUP_B = G * IMG_B;                      // 65,664 B
constexpr int STEM_WARPS = 7, STEM_T = STEM_WARPS * 32;       // 224: 7 passes x 224 = 1,568 tasks
constexpr int WARPS = 14 + STEM_WARPS, THREADS = WARPS * 32; // 672
constexpr int PROD0 = 4, LOADW = 12, MMAW = 13, STEM0 = 14;
constexpr uint32_t SM_B = 0;                                             // 7 x 4 KB weights of conv2
constexpr uint32_t SM_P1 = SM_B + NKB * B_KB_BYTES;                      // 2 group buffers
constexpr uint32_t SM_WM = SM_P1 + 2 * GROUP_B;                          // matmul weights [10][256] + bias
constexpr uint32_t SM_ADD = SM_WM + (NOUT * 256 + 16) * 4;               // conv2 bias [16] | add [16]
constexpr uint32_t SM_PART = SM_ADD + 32 * 4;                            // partial logits
constexpr uint32_t SM_TILE = SM_PART + 4 * 8 * NOUT * 4;                 // 8 haloed 32 x 32 input tiles
constexpr uint32_t SM_RAW = SM_TILE + G * TILE_W * TILE_W * 4;           // the NEXT group's 8 raw images (bulk copy by warp 12)
constexpr uint32_t RAW_B = G * IN_HW * IN_HW * 4;                        // 25,088 B
constexpr uint32_t SM_BARS = SM_RAW + RAW_B;
constexpr int NBARS = 1 + 2 + 2 + NSETS + NSETS + NTILES + NTILES + 2;   // ... | raw_full | raw_empty
constexpr uint32_t SM_SLOT = SM_BARS + NBARS * 8;
constexpr uint32_t SMEM = SM_SLOT + 16 + 1008;                           // + alignment slack (the base is 16-byte aligned)
static_assert(IMG_B % 128 == 16, "image pitch: 16 bytes mod 128");
static_assert(SMEM <= 227 * 1024, "shared memory");
}  // namespace one

struct OneArgs {
  const float* x;        // [N][28*28] the caller's input
  Mnist8StemConsts k;    // stem weights / bias / add as kernel parameters (constant bank)
  const float* bias2;    // [16] or null
  const float* add2;     // [16] or null
  const float* wm;       // [10][256], k = window * 16 + c
  const float* bm;       // [10] or null
  float* out;            // [N][10]
  int N;
  int* nonfinite;
};

__global__ void __launch_bounds__(one::THREADS, 1) mnist8_onepass_kernel(const __grid_constant__ CUtensorMap tmapB, const OneArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + one::SM_BARS;
  const uint32_t b_full = bars;
  auto p1_full = [&](int b) { return bars + 8u * (1 + b); };
  auto p1_empty = [&](int b) { return bars + 8u * (3 + b); };
  auto full_a = [&](int s) { return bars + 8u * (5 + s); };
  auto empty_a = [&](int s) { return bars + 8u * (5 + NSETS + s); };
  auto tmem_full = [&](int t) { return bars + 8u * (5 + 2 * NSETS + t); };
  auto tmem_empty = [&](int t) { return bars + 8u * (5 + 2 * NSETS + NTILES + t); };
  const uint32_t raw_full = bars + 8u * (5 + 2 * NSETS + 2 * NTILES), raw_empty = raw_full + 8u;
  const uint32_t tmem_slot = sbase + one::SM_SLOT;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int groups = (a.N + G - 1) / G;
  int my_groups = 0;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) ++my_groups;

  // ---- per-CTA constants; group buffers and input tiles start as zeros (their halos stay zero)
  float* const wm_s = reinterpret_cast<float*>(gbase + one::SM_WM);
  float* const add_s = reinterpret_cast<float*>(gbase + one::SM_ADD);
  float* const tiles = reinterpret_cast<float*>(gbase + one::SM_TILE);
  for (int i = threadIdx.x; i < NOUT * 256; i += one::THREADS) wm_s[i] = __ldg(a.wm + i);
  if (threadIdx.x < NOUT) wm_s[NOUT * 256 + threadIdx.x] = a.bm ? __ldg(a.bm + threadIdx.x) : 0.f;
  if (threadIdx.x < 16) { add_s[threadIdx.x] = a.bias2 ? __ldg(a.bias2 + threadIdx.x) : 0.f; add_s[16 + threadIdx.x] = a.add2 ? __ldg(a.add2 + threadIdx.x) : 0.f; }
  for (int i = threadIdx.x; i < (int)(2 * one::GROUP_B / 16); i += one::THREADS) reinterpret_cast<float4*>(gbase + one::SM_P1)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < G * TILE_W * TILE_W / 4; i += one::THREADS) reinterpret_cast<float4*>(tiles)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  if (warp == one::LOADW) {
    if (lane == 0) {
      mbar_init(b_full, 1);
      for (int b = 0; b < 2; ++b) { mbar_init(p1_full(b), one::STEM_WARPS); mbar_init(p1_empty(b), NSETS * SET_WARPS); }
      for (int s = 0; s < NSETS; ++s) { mbar_init(full_a(s), SET_WARPS); mbar_init(empty_a(s), 1); }
      for (int t = 0; t < NTILES; ++t) { mbar_init(tmem_full(t), 1); mbar_init(tmem_empty(t), EPI_WARPS); }
      mbar_init(raw_full, 1); mbar_init(raw_empty, one::STEM_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  if (warp >= one::STEM0) {
    // ================================================================ stem: seven warps on the CUDA cores, one group ahead
    const int tid = threadIdx.x - one::STEM0 * 32;
    uint32_t mx = 0;
    int gl = 0;
    for (int g = blockIdx.x; g < groups; g += gridDim.x, ++gl) {
      const int img0 = g * G;
      const int nimg = min(G, a.N - img0);
      // the group's raw images were bulk-copied into the staging area by warp 12 while the previous group was computed:
      // spread them into the haloed tiles (everyone has finished reading the previous group's tiles), hand the staging back
      mbar_wait(raw_full, (uint32_t)gl & 1u);
      asm volatile("bar.sync 2, %0;" ::"n"(one::STEM_T) : "memory");
      {
        const float4* raw = reinterpret_cast<const float4*>(gbase + one::SM_RAW);
        for (int i = tid; i < nimg * 196; i += one::STEM_T) {     // 196 float4 per image, 7 per row
          const float4 v = raw[i];
          const int im = i / 196, q = i - im * 196, r = q / 7, c4 = q - r * 7;
          float* d = tiles + im * (TILE_W * TILE_W) + (r + 2) * TILE_W + 2 + c4 * 4;   // 8-byte aligned
          *reinterpret_cast<float2*>(d) = make_float2(v.x, v.y);
          *reinterpret_cast<float2*>(d + 2) = make_float2(v.z, v.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(raw_empty);   // the stores above consumed the loaded values: the staging may be refilled
      }
      asm volatile("bar.sync 2, %0;" ::"n"(one::STEM_T) : "memory");
      // the group buffer must have been read out by the head's producers (two groups ago)
      mbar_wait(p1_empty(gl & 1), (((uint32_t)gl >> 1) & 1u) ^ 1u);
      uint8_t* const pbuf = gbase + one::SM_P1 + (uint32_t)(gl & 1) * one::GROUP_B;
#pragma unroll 1
      // Task mapping: 32 consecutive pooled pixels per warp-task, 7 passes.  (A conflict-free mapping -- a half-warp per row
      // of 14 pooled pixels, 2 lanes idle, 8 passes -- was measured: 0.598 -> 0.654 ms.  The 14 % extra FMA issue costs more
      // than the three-way bank conflicts of the patch loads: the stem's FMA stream is the critical path.)
      for (int pass = 0; pass < 7; ++pass) {
        const int task = pass * one::STEM_T + tid;
        const int im = task / 196, pp = task - im * 196, ph = pp / P1_HW, pw = pp - ph * P1_HW;
        if (im >= nimg) continue;
        const float* t0 = tiles + im * (TILE_W * TILE_W) + (2 * ph) * TILE_W + 2 * pw;
      float o[8];
      stem_task(t0, a.k, o, mx);
      float* dst = reinterpret_cast<float*>(pbuf + (uint32_t)im * one::IMG_B) + ((ph + 2) * one::PW + (pw + 2)) * C1;
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(p1_full(gl & 1));   // release: this warp's part of the group is written
    }
    if (a.nonfinite && mx >= 0x7f800000u) *a.nonfinite = 1;
  } else if (warp >= one::PROD0 && warp < one::PROD0 + NSETS * SET_WARPS) {
    // ================================================================ A producers (as in mnist8_head_kernel, 16-pixel rows)
    const int pw_ = warp - one::PROD0;
    const int quarter = pw_ & 3, set = pw_ >> 2;
    const int img = lane & 7, win = quarter * 4 + (lane >> 3);
    const uint32_t rowoff = (uint32_t)img * one::IMG_B + (uint32_t)(((3 * (win >> 2)) * one::PW + 3 * (win & 3)) * C1 * 4);
    const uint32_t t_a = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(A_COL0 + set * A_BUF_COLS);
    const int total_rows = my_groups * NROWS;
    uint32_t ph = 0;
    for (int idx = set; idx < total_rows; idx += NSETS) {
      const int gl = idx / NROWS, rr = idx - gl * NROWS;
      if (idx < NSETS || (idx - NSETS) / NROWS != gl) mbar_wait(p1_full(gl & 1), (uint32_t)(gl >> 1) & 1u);
      const uint32_t src = sbase + one::SM_P1 + (uint32_t)(gl & 1) * one::GROUP_B + rowoff + (uint32_t)(rr * one::PW * C1 * 4);
      float4 x[NBLK][2];
#pragma unroll
      for (int sp = 0; sp < NBLK; ++sp) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[sp][0].x), "=f"(x[sp][0].y), "=f"(x[sp][0].z), "=f"(x[sp][0].w) : "r"(src + 32u * sp));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[sp][1].x), "=f"(x[sp][1].y), "=f"(x[sp][1].z), "=f"(x[sp][1].w) : "r"(src + 32u * sp + 16u));
      }
      const bool last_of_group = idx + NSETS >= total_rows || (idx + NSETS) / NROWS != gl;
      if (last_of_group) {
        const uint32_t dep = (__float_as_uint(x[0][0].w) ^ __float_as_uint(x[3][1].w) ^ __float_as_uint(x[6][1].w)) & (uint32_t)(a.N >> 31);   // always 0, opaque
        __syncwarp();
        if (lane == 0) mbar_arrive(p1_empty(gl & 1) + dep);
      }
      mbar_wait(empty_a(set), ph ^ 1u);
      tc_fence_after();
#pragma unroll
      for (int sp = 0; sp < NBLK; ++sp) {
        const float4 u = x[sp][0], v = x[sp][1];
        const float h0 = split_hi(u.x), h1 = split_hi(u.y), h2 = split_hi(u.z), h3 = split_hi(u.w);
        const float h4 = split_hi(v.x), h5 = split_hi(v.y), h6 = split_hi(v.z), h7 = split_hi(v.w);
        tmem_st_32x32b_x8(t_a + 8u * sp, h0, h1, h2, h3, h4, h5, h6, h7);
        tmem_st_32x32b_x8(t_a + 56u + 8u * sp, split_lo(u.x, h0), split_lo(u.y, h1), split_lo(u.z, h2), split_lo(u.w, h3),
                          split_lo(v.x, h4), split_lo(v.y, h5), split_lo(v.z, h6), split_lo(v.w, h7));
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_a(set));
      ph ^= 1u;
    }
  } else if (warp == one::LOADW) {
    if (lane == 0) {
      mbar_expect_tx(b_full, NKB * B_KB_BYTES);
      for (int kb = 0; kb < NKB; ++kb) {
        tma_load_2d(sbase + one::SM_B + (uint32_t)kb * B_KB_BYTES, &tmapB, b_full, kb * 32, 0);
        tma_load_2d(sbase + one::SM_B + (uint32_t)kb * B_KB_BYTES + C2 * 128, &tmapB, b_full, kb * 32, C2);
      }
      // the input images, one group ahead of the stem: one bulk copy of 8 x 3,136 bytes per group
      int gl = 0;
      for (int g = blockIdx.x; g < groups; g += gridDim.x, ++gl) {
        mbar_wait(raw_empty, ((uint32_t)gl & 1u) ^ 1u);
        const uint32_t bytes = (uint32_t)min(G, a.N - g * G) * (IN_HW * IN_HW * 4);
        mbar_expect_tx(raw_full, bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + one::SM_RAW),
                     "l"(a.x + (size_t)g * G * (IN_HW * IN_HW)), "r"(bytes), "r"(raw_full)
                     : "memory");
      }
    }
  } else if (warp == one::MMAW) {
    // ================================================================ MMA issuer (as in mnist8_head_kernel)
    const uint32_t idesc32 = instr_desc_tf32(2 * C2), idesc16 = instr_desc_tf32(C2);
    const bool leader = elect_one();
    const uint32_t b_lo0 = (((sbase + one::SM_B) >> 4) & 0x3FFFu) | (1u << 16);
    mbar_wait(b_full, 0);
    int buf = 0;
    uint32_t pha = 0;
    for (int gl = 0; gl < my_groups; ++gl) {
      const uint32_t gph = (uint32_t)gl & 1u;
#pragma unroll 1
      for (int rr = 0; rr < NROWS; ++rr) {
        mbar_wait(full_a(buf), pha);
        tc_fence_after();
        const uint32_t a_hi = tmem_base + (uint32_t)(A_COL0 + buf * A_BUF_COLS), a_lo = a_hi + 56u;
#pragma unroll
        for (int jh = 0; jh < 3; ++jh) {
          const int r = rr - jh;
          if (r < 0 || r > 4) continue;
          if (r == 0) {
#pragma unroll
            for (int jw = 0; jw < 3; ++jw) mbar_wait(tmem_empty(jh * 3 + jw), gph ^ 1u);
            tc_fence_after();
          }
          if (leader) {
#pragma unroll
            for (int jw = 0; jw < 3; ++jw) {
              const uint32_t d_main = tmem_base + (uint32_t)((jh * 3 + jw) * 32), d_corr = d_main + (uint32_t)C2;
#pragma unroll
              for (int s = 0; s < 5; ++s) {
                const int tap = r * 5 + s;
                const uint32_t bl = b_lo0 + (uint32_t)(tap >> 2) * (B_KB_BYTES >> 4) + 2u * (uint32_t)(tap & 3);
                umma_tf32_ts(d_main, a_hi + 8u * (uint32_t)(jw + s), sw128_desc(bl), idesc32, (r | s) != 0 ? 1u : 0u);
                umma_tf32_ts(d_corr, a_lo + 8u * (uint32_t)(jw + s), sw128_desc(bl), idesc16, 1u);
              }
            }
          }
          __syncwarp();
          if (r == 4 && leader) {
#pragma unroll
            for (int jw = 0; jw < 3; ++jw) umma_commit(tmem_full(jh * 3 + jw));
          }
        }
        __syncwarp();
        if (leader) umma_commit(empty_a(buf));
        if (++buf == NSETS) { buf = 0; pha ^= 1u; }
      }
    }
  } else if (warp < EPI_WARPS) {
    // ================================================================ epilogue (as in mnist8_head_kernel)
    const int win = warp * 4 + (lane >> 3);
    float* const part = reinterpret_cast<float*>(gbase + one::SM_PART);
    int gl = 0;
    for (int g = blockIdx.x; g < groups; g += gridDim.x, ++gl) {
      float m[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) m[c] = -3.402823466e+38f;
      for (int t = 0; t < NTILES; ++t) {
        mbar_wait(tmem_full(t), (uint32_t)gl & 1u);
        tc_fence_after();
        uint32_t acc[16], cor[16];
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 32);
        tmem_ld16(ta, acc);
        tmem_ld16(ta + 16u, cor);
        tmem_ld_wait(acc, cor);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty(t));
#pragma unroll
        for (int c = 0; c < 16; ++c) m[c] = fmaxf(m[c], __uint_as_float(acc[c]) + __uint_as_float(cor[c]));
      }
      float f[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) f[c] = fmaxf((m[c] + add_s[c]) + add_s[16 + c], 0.f);
      float o[NOUT];
#pragma unroll
      for (int n = 0; n < NOUT; ++n) {
        const float4* wr = reinterpret_cast<const float4*>(wm_s + n * 256 + win * 16);
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = wr[q];
          s = fmaf(f[4 * q], w4.x, s); s = fmaf(f[4 * q + 1], w4.y, s); s = fmaf(f[4 * q + 2], w4.z, s); s = fmaf(f[4 * q + 3], w4.w, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        o[n] = s;
      }
      float* pb = part;   // single buffer (shared memory is full): a second barrier below separates the groups
      if (lane < 8) {
#pragma unroll
        for (int n = 0; n < NOUT; ++n) pb[(warp * 8 + lane) * NOUT + n] = o[n];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int tid = warp * 32 + lane;
      if (tid < 8 * NOUT) {
        const int im = tid / NOUT, n = tid - im * NOUT;
        const float s = (pb[(0 * 8 + im) * NOUT + n] + pb[(1 * 8 + im) * NOUT + n]) + (pb[(2 * 8 + im) * NOUT + n] + pb[(3 * 8 + im) * NOUT + n]);
        if (g * G + im < a.N) a.out[(size_t)(g * G + im) * NOUT + n] = s + wm_s[NOUT * 256 + n];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == one::LOADW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
size_t mnist8_p1_floats(int N) { return (size_t)((N + G - 1) / G) * G * IMG_FLOATS; }

int launch_mnist8_stem(const float* x, const Mnist8StemConsts& k, float* p1, int N, cudaStream_t st, int* nonfinite) {
  if (N <= 0) return 0;
  StemArgs a{x, k, p1, N, nonfinite};
  const int groups = (N + G - 1) / G;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = groups < sms * 3 ? groups : sms * 3;
  static bool attr_set[64] = {false};
  if (dev < 64 && !attr_set[dev]) {
    B200_CUDA(cudaFuncSetAttribute(mnist8_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM));
    attr_set[dev] = true;
  }
  mnist8_stem_kernel<<<grid, STEM_THREADS, STEM_SMEM, st>>>(a);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_mnist8_head(const float* p1, const TcWeights& w2, const float* bias2, const float* add2, const float* wm, const float* bm,
                       float* out, int N, cudaStream_t st) {
  if (N <= 0) return 0;
  if (w2.M != C2 || w2.K != K2 || w2.BN != C2 || w2.Mpad != C2 || w2.Kpad != NKB * 32)
    B200_FAIL(B200_EINVAL, "mnist8 head: weights prepared for M=%d K=%d BN=%d", w2.M, w2.K, w2.BN);
  static bool attr_set[64] = {false};
  static int sm_count[64] = {0};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    B200_CUDA(cudaFuncSetAttribute(mnist8_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM));
    B200_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    attr_set[dev] = true;
  }
  const int sms = (dev < 64 && sm_count[dev] > 0) ? sm_count[dev] : 148;
  const int groups = (N + G - 1) / G;
  HeadArgs a{p1, bias2, add2, wm, bm, out, N};
  mnist8_head_kernel<<<groups < sms ? groups : sms, HEAD_THREADS, HEAD_SMEM, st>>>(w2.tmap, a);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_mnist8_onepass(const float* x, const Mnist8StemConsts& k, const TcWeights& w2, const float* bias2, const float* add2, const float* wm,
                          const float* bm, float* out, int N, cudaStream_t st, int* nonfinite) {
  if (N <= 0) return 0;
  if (w2.M != C2 || w2.K != K2 || w2.BN != C2 || w2.Mpad != C2 || w2.Kpad != NKB * 32)
    B200_FAIL(B200_EINVAL, "mnist8: weights prepared for M=%d K=%d BN=%d", w2.M, w2.K, w2.BN);
  static bool attr_set[64] = {false};
  static int sm_count[64] = {0};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    B200_CUDA(cudaFuncSetAttribute(mnist8_onepass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one::SMEM));
    B200_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    attr_set[dev] = true;
  }
  const int sms = (dev < 64 && sm_count[dev] > 0) ? sm_count[dev] : 148;
  const int groups = (N + G - 1) / G;
  OneArgs a{x, k, bias2, add2, wm, bm, out, N, nonfinite};
  mnist8_onepass_kernel<<<groups < sms ? groups : sms, one::THREADS, one::SMEM, st>>>(w2.tmap, a);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
