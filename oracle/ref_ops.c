/*
 * oracle/ref_ops.c -- CPU restatement of the reference's fp32 operator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (libb200rt.so, the
 * onnx_rusty_inference_engine_b200 package) links, imports or calls this file.
 * It is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs as the checker and the timed CPU baseline.
 *
 * The reference is Rust and cannot be compiled in this image (no cargo/rustc),
 * so this is a restatement in plain C of the algorithm each function cites.
 * Parity pinning: the whole MNIST-8 graph built from these functions reproduces
 * the reference's bundled golden pair mnist_data_0.pb -> mnist_output_0.pb
 * (tests/test_oracle_golden.py).  Ops that only SqueezeNet uses (Concat, Dropout,
 * GlobalAveragePool, Softmax, strided / biased Conv) are "parity unpinned":
 * models/squeezenet1.0-8.onnx is not shipped with the reference, so they are
 * cross-checked against torch.nn.functional only.
 *
 * Third-party arithmetic restated here (crates are not vendored in /root/reference):
 *   ndarray 0.15.x  numeric_util::unrolled_fold  -- the 8-accumulator sum used by
 *                   ArrayBase::sum() on contiguous data (call site convolution_op.rs:480).
 *   ndarray 0.15.x  Array2::dot -> matrixmultiply sgemm (call site mul_op.rs:23); restated
 *                   as a plain k-ascending dot product (order differs, inside 1e-4 rel).
 *
 * All tensors are single images: the reference is batch-1 only
 * (convolution_op.rs:623,647,480; max_pool_op.rs:436,337), so callers loop over images.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

enum { PAD_VALID = 0, PAD_SAME_UPPER = 1, PAD_SAME_LOWER = 2, PAD_NOTSET = 3 };

/* ndarray 0.15 numeric_util::unrolled_fold specialised to f32 addition (see header). */
static float nd_sum(const float *xs, size_t n) {
  float acc = 0.f, p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f, p5 = 0.f, p6 = 0.f, p7 = 0.f;
  while (n >= 8) {
    p0 += xs[0]; p1 += xs[1]; p2 += xs[2]; p3 += xs[3];
    p4 += xs[4]; p5 += xs[5]; p6 += xs[6]; p7 += xs[7];
    xs += 8; n -= 8;
  }
  acc += (p0 + p4);
  acc += (p1 + p5);
  acc += (p2 + p6);
  acc += (p3 + p7);
  for (size_t i = 0; i < n; ++i) acc += xs[i];
  return acc;
}

/* get_padding_size, convolution_op.rs:519-557 and max_pool_op.rs:363-401.
 * Returns (top, bottom, left, right) AFTER the reference's deliberate swap (:547-556):
 * the larger half of an odd total goes to top/left. */
static void same_pads(size_t in_h, size_t in_w, size_t sh, size_t sw, size_t kh, size_t kw,
                      size_t *top, size_t *bottom, size_t *left, size_t *right) {
  size_t ph = (in_h % sh == 0) ? (kh - sh) : (kh - (in_h % sh));
  size_t pw = (in_w % sw == 0) ? (kw - sw) : (kw - (in_w % sw));
  size_t t = ph / 2, b = ph - t, l = pw / 2, r = pw - l;
  *top = b; *bottom = t; *left = r; *right = l;
}

/* Output geometry shared by conv2d (convolution_op.rs:293-324) and max_pool2d (max_pool_op.rs:215-246).
 * Returns 0 on success. */
static int out_dims(int auto_pad, size_t H, size_t W, size_t kh, size_t kw, size_t sh, size_t sw,
                    const size_t p[4] /* top,bottom,left,right (NOTSET only) */, size_t *Ho, size_t *Wo) {
  if (sh == 0 || sw == 0) return -1;
  switch (auto_pad) {
    case PAD_SAME_UPPER:
    case PAD_SAME_LOWER:
      *Ho = (size_t)ceilf((float)H / (float)sh);
      *Wo = (size_t)ceilf((float)W / (float)sw);
      return 0;
    case PAD_NOTSET:
      if (H + p[0] + p[1] < kh || W + p[2] + p[3] < kw) return -1; /* usize underflow panics upstream */
      *Ho = (H - kh + (p[0] + p[1])) / sh + 1;
      *Wo = (W - kw + (p[2] + p[3])) / sw + 1;
      return 0;
    default:
      if (H < kh || W < kw) return -1;
      *Ho = (H - kh) / sh + 1;
      *Wo = (W - kw) / sw + 1;
      return 0;
  }
}

/* Zero-padded copy of one image, convolution_op.rs:334-362 / max_pool_op.rs:248-276. */
static float *pad_image(const float *x, size_t C, size_t H, size_t W, size_t top, size_t bottom,
                        size_t left, size_t right, size_t *Hp, size_t *Wp) {
  *Hp = H + top + bottom;
  *Wp = W + left + right;
  float *out = (float *)calloc(C * (*Hp) * (*Wp) + 1, sizeof(float));
  if (!out) return NULL;
  for (size_t c = 0; c < C; ++c)
    for (size_t h = 0; h < H; ++h)
      memcpy(out + (c * (*Hp) + h + top) * (*Wp) + left, x + (c * H + h) * W, W * sizeof(float));
  return out;
}

/* im2col_ref, convolution_op.rs:560-663 (dilation==1 branches) and max_pool_op.rs:403-449:
 * rows ordered (c, ho, wo); each row is the kh x kw patch, row-major. */
static float *im2col(const float *xp, size_t C, size_t Hp, size_t Wp, size_t kh, size_t kw, size_t sh,
                     size_t sw, size_t *new_h, size_t *new_w) {
  *new_h = (Hp - kh) / sh + 1;
  *new_w = (Wp - kw) / sw + 1;
  size_t rows = C * (*new_h) * (*new_w), cols = kh * kw;
  float *col = (float *)malloc((rows * cols + 1) * sizeof(float));
  if (!col) return NULL;
  size_t cont = 0;
  for (size_t k = 0; k < C; ++k)
    for (size_t i = 0; i < *new_h; ++i)
      for (size_t j = 0; j < *new_w; ++j) {
        float *row = col + cont * cols;
        for (size_t a = 0; a < kh; ++a)
          for (size_t b = 0; b < kw; ++b) row[a * kw + b] = xp[(k * Hp + i * sh + a) * Wp + j * sw + b];
        ++cont;
      }
  return col;
}

/*
 * conv2d, convolution_op.rs:224-517 (+ new_onnx_tensor_flow :57-71, ker2col_ref :666-703, add_bias :705-726).
 *   x [C,H,W], w [M,C,kh,kw] (ONNX order; the reference permutes to (C,M,kw,kh) and ker2col_ref
 *   reads it back so that row m*C+c == w[m,c,:,:] row-major), bias [M] or NULL.
 *   pads = ONNX order [h_begin, w_begin, h_end, w_end] (:267-278), used when auto_pad==NOTSET.
 *   group must be 1 and dilations 1 (other values are broken upstream).
 * Accumulation order (:407-504): for m, for p (row-major ho,wo), for c ascending:
 *   out[m,p] += nd_sum(patch[c,p,:] * ker[m,c,:]); bias added last in a separate pass.
 * Returns 0 on success, -1 on a condition that panics upstream.
 */
int ref_conv2d(const float *x, size_t C, size_t H, size_t W, const float *w, size_t M, size_t kh, size_t kw,
               const float *bias, int auto_pad, const long long pads[4], const long long strides[2],
               float *out, size_t out_capacity, size_t *Ho_out, size_t *Wo_out) {
  size_t sh = (size_t)strides[0], sw = (size_t)strides[1];
  size_t p[4] = {0, 0, 0, 0};
  if (auto_pad == PAD_NOTSET) {
    p[0] = (size_t)pads[0]; p[1] = (size_t)pads[2]; p[2] = (size_t)pads[1]; p[3] = (size_t)pads[3];
  }
  size_t Ho, Wo;
  if (out_dims(auto_pad, H, W, kh, kw, sh, sw, p, &Ho, &Wo)) return -1;
  *Ho_out = Ho; *Wo_out = Wo;
  if (out == NULL) return 0; /* shape query */
  if (M * Ho * Wo > out_capacity) return -1;

  size_t top = 0, bottom = 0, left = 0, right = 0;
  if (auto_pad == PAD_SAME_UPPER || auto_pad == PAD_SAME_LOWER) {
    if (kh < sh || kw < sw) return -1; /* usize underflow panics upstream */
    same_pads(H, W, sh, sw, kh, kw, &top, &bottom, &left, &right);
  } else if (auto_pad == PAD_NOTSET) {
    top = p[0]; bottom = p[1]; left = p[2]; right = p[3];
  }
  size_t Hp, Wp, nh, nw;
  float *xp = pad_image(x, C, H, W, top, bottom, left, right, &Hp, &Wp);
  if (!xp) return -1;
  float *col = im2col(xp, C, Hp, Wp, kh, kw, sh, sw, &nh, &nw);
  free(xp);
  if (!col) return -1;
  if (nh * nw != Ho * Wo) { free(col); return -1; } /* upstream would index out of bounds */

  const size_t kk = kh * kw, P = Ho * Wo;
  float *row_mul = (float *)malloc((kk + 1) * sizeof(float));
  memset(out, 0, M * P * sizeof(float));
  for (size_t m = 0; m < M; ++m)
    for (size_t pix = 0; pix < P; ++pix) {
      float acc = 0.f;
      for (size_t c = 0; c < C; ++c) {
        const float *im_row = col + (c * P + pix) * kk;
        const float *ker_row = w + (m * C + c) * kk;
        for (size_t i = 0; i < kk; ++i) row_mul[i] = im_row[i] * ker_row[i];
        acc += nd_sum(row_mul, kk);
      }
      out[m * P + pix] = acc;
    }
  free(row_mul);
  free(col);
  if (bias)
    for (size_t m = 0; m < M; ++m)
      for (size_t pix = 0; pix < P; ++pix) out[m * P + pix] += bias[m];
  return 0;
}

/*
 * max_pool2d, max_pool_op.rs:157-360.  Padding is ZERO-fill (:265-276), the fold starts from
 * F::min_value() == -FLT_MAX (:337).  pads honoured only when auto_pad==NOTSET (:188-201).
 */
int ref_maxpool2d(const float *x, size_t C, size_t H, size_t W, size_t kh, size_t kw, int auto_pad,
                  const long long pads[4], const long long strides[2], float *out, size_t out_capacity,
                  size_t *Ho_out, size_t *Wo_out) {
  size_t sh = (size_t)strides[0], sw = (size_t)strides[1];
  size_t p[4] = {0, 0, 0, 0};
  if (auto_pad == PAD_NOTSET) {
    p[0] = (size_t)pads[0]; p[1] = (size_t)pads[2]; p[2] = (size_t)pads[1]; p[3] = (size_t)pads[3];
  }
  size_t Ho, Wo;
  if (out_dims(auto_pad, H, W, kh, kw, sh, sw, p, &Ho, &Wo)) return -1;
  *Ho_out = Ho; *Wo_out = Wo;
  if (out == NULL) return 0;
  if (C * Ho * Wo > out_capacity) return -1;
  size_t top = 0, bottom = 0, left = 0, right = 0;
  if (auto_pad == PAD_SAME_UPPER || auto_pad == PAD_SAME_LOWER) {
    if (kh < sh || kw < sw) return -1;
    same_pads(H, W, sh, sw, kh, kw, &top, &bottom, &left, &right);
  } else if (auto_pad == PAD_NOTSET) {
    top = p[0]; bottom = p[1]; left = p[2]; right = p[3];
  }
  size_t Hp, Wp, nh, nw;
  float *xp = pad_image(x, C, H, W, top, bottom, left, right, &Hp, &Wp);
  if (!xp) return -1;
  float *col = im2col(xp, C, Hp, Wp, kh, kw, sh, sw, &nh, &nw);
  free(xp);
  if (!col) return -1;
  if (nh * nw != Ho * Wo) { free(col); return -1; }
  const size_t kk = kh * kw, P = Ho * Wo;
  for (size_t c = 0; c < C; ++c)
    for (size_t pix = 0; pix < P; ++pix) {
      const float *row = col + (c * P + pix) * kk;
      float m = -FLT_MAX;
      for (size_t i = 0; i < kk; ++i) m = (row[i] > m) ? row[i] : m; /* f32::max */
      out[c * P + pix] = m;
    }
  free(col);
  return 0;
}

/* relu_wrapper, relu_op.rs:31-33: x.map(|v| v.max(0.0)). */
void ref_relu(const float *x, size_t n, float *out) {
  for (size_t i = 0; i < n; ++i) out[i] = x[i] > 0.f ? x[i] : 0.f;
}

/* add, add_op.rs:75: Array4 [1,C,H,W] + Array3 [C,1,1] broadcast. */
void ref_add_channel(const float *x, size_t C, size_t HW, const float *b, float *out) {
  for (size_t c = 0; c < C; ++c)
    for (size_t i = 0; i < HW; ++i) out[c * HW + i] = x[c * HW + i] + b[c];
}

/* add, add_op.rs:84: Array2 + Array2, same shape. */
void ref_add_same(const float *a, const float *b, size_t n, float *out) {
  for (size_t i = 0; i < n; ++i) out[i] = a[i] + b[i];
}

/* mul, mul_op.rs:23: Array2::dot, [R,K]x[K,N]. */
void ref_matmul(const float *a, const float *b, size_t R, size_t K, size_t N, float *out) {
  for (size_t r = 0; r < R; ++r)
    for (size_t n = 0; n < N; ++n) {
      float acc = 0.f;
      for (size_t k = 0; k < K; ++k) acc += a[r * K + k] * b[k * N + n];
      out[r * N + n] = acc;
    }
}

/* global_average_pool_wrapper, global_average_pool_op.rs:33-52: sequential iter().sum() / len. */
void ref_global_avgpool(const float *x, size_t C, size_t HW, float *out) {
  for (size_t c = 0; c < C; ++c) {
    float s = 0.f;
    for (size_t i = 0; i < HW; ++i) s += x[c * HW + i];
    out[c] = s / (float)HW;
  }
}

/* softmax_wrapper, softmax_op.rs:45-57 for one row of n = C*H*W elements. */
void ref_softmax_row(const float *x, size_t n, float *out) {
  float mx = -INFINITY;
  for (size_t i = 0; i < n; ++i) mx = x[i] > mx ? x[i] : mx;
  float s = 0.f;
  for (size_t i = 0; i < n; ++i) { out[i] = expf(x[i] - mx); s += out[i]; }
  for (size_t i = 0; i < n; ++i) out[i] = out[i] / s;
}
