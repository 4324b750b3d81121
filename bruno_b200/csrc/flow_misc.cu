// Memory-bound pieces of the Real NVP stack around the tensor-core convolutions (nn_extra_nvp.py):
//   pre-processing (ScaleLayer :41-55, LogitLayer :22-38), squeeze (:277-298), factor-out slices (:301-346),
//   the 1x1 input conv c1 of the s/t ResNet (:115) with the mask multiply fused in, and the two 1x1 output heads
//   (:129-139) fused with tanh, mask, the affine coupling (:162-187) and the per-image log-det reduction.
#include <cuda_bf16.h>

#include "common.cuh"
#include "slab.cuh"

namespace bruno {

// ------------------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
  // deterministic: fixed shuffle tree + fixed-order sum of the warp totals.  blockDim.x == 256.
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// a / d for 0 <= a < 2^22 through one FMUL + F2I (exact: (a + 0.5) / d is at least 0.5 / d away from an integer and the
// float product is off by less than that); a runtime-divisor integer division costs ~25 instructions, and the
// row -> (image, line, column) decomposition needs three of them per 16 bytes moved.
__device__ __forceinline__ int fdiv22(int a, int d, float inv_d, bool small) {
  return small ? static_cast<int>((static_cast<float>(a) + 0.5f) * inv_d) : a / d;
}

__device__ __forceinline__ float mask_b(int mask_type, int y, int x, int c, int C) {
  // nn_extra_nvp.py:143-160.  b = 1: pass-through / conditioning half.
  switch (mask_type) {
    case BRUNO_MASK_CHECKERBOARD0: return static_cast<float>((y + x) & 1);
    case BRUNO_MASK_CHECKERBOARD1: return static_cast<float>(1 - ((y + x) & 1));
    case BRUNO_MASK_CHANNEL0: return c < C / 2 ? 1.f : 0.f;
    default: return c >= C / 2 ? 1.f : 0.f;
  }
}

__device__ __forceinline__ void unpack8f(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2f(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : __expf(x) - 1.f; }
__device__ __forceinline__ float elu_grad_from_out(float o) { return o > 0.f ? 1.f : o + 1.f; }

// ------------------------------------------------------------------------------------------------------------
// pre-processing: one block per image, per-image log-det reduced in a fixed order
// ------------------------------------------------------------------------------------------------------------
enum PreOp { PRE_SCALE_FWD = 0, PRE_LOGIT_FWD = 1, PRE_LOGIT_INV = 2, PRE_SCALE_INV = 3, PRE_FUSED_FWD = 4, PRE_FUSED_INV = 5 };

__global__ void __launch_bounds__(256)
pre_kernel(int op, const float* __restrict__ x, const float* __restrict__ ld_in, float* __restrict__ y,
           float* __restrict__ ld_out, int D, float scale, float alpha, int conditional_variant) {
  __shared__ float red[8];
  const int n = blockIdx.x;
  const float* xi = x + static_cast<size_t>(n) * D;
  float* yi = y + static_cast<size_t>(n) * D;
  const float l1a = conditional_variant ? 0.f : log1pf(-alpha);  // nn_extra_nvp_conditional.py:13 omits ln(1-alpha)
  float acc = 0.f;
  for (int i = threadIdx.x; i < D; i += 256) {
    float v = xi[i];
    if (op == PRE_SCALE_FWD || op == PRE_FUSED_FWD) v = v / scale;                          // nvp:47
    if (op == PRE_LOGIT_FWD || op == PRE_FUSED_FWD) {                                       // nvp:27-30
      const float t = v * (1.f - alpha) + alpha * 0.5f;
      const float la = logf(t), lb = logf(1.f - t);
      acc += -la - lb + l1a;
      v = la - lb;
    }
    if (op == PRE_LOGIT_INV || op == PRE_FUSED_INV) {                                       // nvp:34-37
      const float yv = v;
      const float sg = 1.f / (1.f + expf(-yv));
      v = (sg - 0.5f * alpha) / (1.f - alpha);
      const float sp = yv > 15.f ? yv + log1pf(expf(-yv)) : log1pf(expf(yv));
      acc += yv - 2.f * sp - l1a;
    }
    if (op == PRE_SCALE_INV || op == PRE_FUSED_INV) v = v * scale;                          // nvp:53
    yi[i] = v;
  }
  const float tot = block_sum_256(acc, red);
  if (threadIdx.x == 0) {
    float l = ld_in ? ld_in[n] : 0.f;
    const float lsc = logf(scale) * static_cast<float>(D);
    if (op == PRE_SCALE_FWD || op == PRE_FUSED_FWD) l -= lsc;                               // nvp:48
    l += tot;
    if (op == PRE_SCALE_INV || op == PRE_FUSED_INV) l += lsc;                               // nvp:54
    ld_out[n] = l;
  }
}

// ------------------------------------------------------------------------------------------------------------
// squeeze / channel slices (pure index remaps on the small C-channel tensors)
// ------------------------------------------------------------------------------------------------------------
// tf.space_to_depth(x, 2): out[n,h,w,(dy*2+dx)*C + c] = in[n,2h+dy,2w+dx,c]      (nvp:284)
__global__ void space_to_depth_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C,
                                      int inverse) {
  const size_t total = static_cast<size_t>(N) * H * W * C;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i indexes the LARGE-resolution tensor [N,H,W,C]
  const int c = static_cast<int>(i % C);
  const int x = static_cast<int>((i / C) % W);
  const int y = static_cast<int>((i / (static_cast<size_t>(C) * W)) % H);
  const int n = static_cast<int>(i / (static_cast<size_t>(C) * W * H));
  const size_t j = ((static_cast<size_t>(n) * (H / 2) + y / 2) * (W / 2) + x / 2) * (4 * C) + ((y & 1) * 2 + (x & 1)) * C + c;
  if (inverse) out[i] = in[j];
  else out[j] = in[i];
}

// dst[r, dst_off + c] = src[r, src_off + c] for c < nch
__global__ void channel_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t rows, int src_C,
                                    int src_off, int dst_C, int dst_off, int nch) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * nch) return;
  const size_t r = i / nch;
  const int c = static_cast<int>(i % nch);
  dst[r * dst_C + dst_off + c] = src[r * src_C + src_off + c];
}

// ------------------------------------------------------------------------------------------------------------
// c1: 1x1 weight-normalised conv C -> 64 on the masked input + ELU, written as a swizzled bf16 slab.  nvp:113-115
// ------------------------------------------------------------------------------------------------------------
constexpr int MAXC = 16;
// rows of the c1 weight matrix: ks*ks*C.  ks = 1 everywhere (nvp:113) except the label-conditioned couplings, whose c1 is a
// 3x3 convolution (nn_extra_nvp_conditional.py:427 calls the inner conv2d_wn without forwarding filter_size) with C <= 8.
constexpr int MAXK = 72;
constexpr int HEADS_BWD_BLOCKS = 296;   // 2 resident blocks per SM

__device__ __forceinline__ void c1_weights_to_smem(const float* __restrict__ V1, const float* __restrict__ g1, int C,
                                                   float* w_s /*[C][64]*/) {
  // W1[c][co] = g1[co] * V1[c][co] / sqrt(max(sum_c V1^2, 1e-12));  C = all rows (ks*ks*Cin)
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    float n2 = 0.f;
    for (int c = 0; c < C; ++c) n2 = fmaf(V1[c * 64 + i], V1[c * 64 + i], n2);
    const float sc = g1[i] * rsqrtf(fmaxf(n2, 1e-12f));
    for (int c = 0; c < C; ++c) w_s[c * 64 + i] = V1[c * 64 + i] * sc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ V1, const float* __restrict__ g1,
              const float* __restrict__ b1, int bias_img_stride, uint8_t* __restrict__ out, SlabGeom g, int C,
              int mask_type, int f32_slab, int KS) {
  __shared__ float w_s[MAXK * 64];
  __shared__ float b_s[64];
  if (threadIdx.x < 64) b_s[threadIdx.x] = bias_img_stride ? 0.f : b1[threadIdx.x];
  c1_weights_to_smem(V1, g1, KS * KS * C, w_s);
  const int total = g.NT * TILE_M * 8;
  const bool small = g.NT * TILE_M < (1 << 22);
  const float inv_img = 1.0f / static_cast<float>(g.IMG), inv_p = 1.0f / static_cast<float>(g.P);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int data_row = i >> 3, j = i & 7;
    const size_t R = static_cast<size_t>(g.HALO) + data_row;
    uint4 packed = make_uint4(0, 0, 0, 0);
    float outf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int n = fdiv22(data_row, g.IMG, inv_img, small), q = data_row - n * g.IMG;
    const int y1 = fdiv22(q, g.P, inv_p, small);
    const int yy = y1 - 1, xx = q - y1 * g.P - 1;
    if (data_row < g.data_rows && yy >= 0 && xx >= 0 && xx < g.W) {
      const float* xp = x + ((static_cast<size_t>(n) * g.H + yy) * g.W + xx) * C;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        acc[e] = bias_img_stride ? b1[static_cast<size_t>(n) * bias_img_stride + j * 8 + e] : b_s[j * 8 + e];
      if (KS == 1) {
        for (int c = 0; c < C; ++c) {
          const float xm = xp[c] * mask_b(mask_type, yy, xx, c, C);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(xm, w_s[c * 64 + j * 8 + e], acc[e]);
        }
      } else {                               // 3x3, zero 'SAME' padding, rows of V1 ordered (kh, kw, cin) like HWIO
        for (int t = 0; t < 9; ++t) {
          const int y2 = yy + t / 3 - 1, x2 = xx + t % 3 - 1;
          if (y2 < 0 || y2 >= g.H || x2 < 0 || x2 >= g.W) continue;
          const float* xq = x + ((static_cast<size_t>(n) * g.H + y2) * g.W + x2) * C;
          for (int c = 0; c < C; ++c) {
            const float xm = xq[c] * mask_b(mask_type, y2, x2, c, C);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(xm, w_s[(t * C + c) * 64 + j * 8 + e], acc[e]);
          }
        }
      }
      if (f32_slab) {
#pragma unroll
        for (int e = 0; e < 8; ++e) outf[e] = acc[e] > 0.f ? acc[e] : expm1f(acc[e]);
      } else {
        packed = make_uint4(pack2f(elu_f(acc[0]), elu_f(acc[1])), pack2f(elu_f(acc[2]), elu_f(acc[3])),
                            pack2f(elu_f(acc[4]), elu_f(acc[5])), pack2f(elu_f(acc[6]), elu_f(acc[7])));
      }
    }
    if (f32_slab) {       // fp32 verification layout: [row][64] floats, no swizzle
      float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + R * 64 + j * 8);
      o4[0] = make_float4(outf[0], outf[1], outf[2], outf[3]);
      o4[1] = make_float4(outf[4], outf[5], outf[6], outf[7]);
    } else {
      *reinterpret_cast<uint4*>(out + slab_chunk_off(R, j)) = packed;
    }
  }
}

// c1 forward, C <= 4, 1x1, shared bias, bf16 slab (every unconditional config): eight lanes per PIXEL -- lane j owns the
// 16-byte chunk j of the slab row and keeps its 8 x C weights and 8 biases in registers; only pixel rows are visited (the
// padding rows of a slab are zero and stay zero).  Same arithmetic per output as c1_fwd_kernel (bias, then fmaf over c
// ascending, ELU, bf16): the two are bit-identical.  The generic kernel ran ~236 instructions per 16 bytes written (47 us
// for a 640 x 28 x 28 slab, issue-bound).
template <int C>
__global__ void __launch_bounds__(256)
c1_fwd8_kernel(const float* __restrict__ x, const float* __restrict__ V1, const float* __restrict__ g1,
               const float* __restrict__ b1, uint8_t* __restrict__ out, SlabGeom g, int mask_type) {
  __shared__ float w_s[MAXK * 64];
  c1_weights_to_smem(V1, g1, C, w_s);
  const int j = threadIdx.x & 7;
  float w[C][8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    b[e] = b1[j * 8 + e];
#pragma unroll
    for (int c = 0; c < C; ++c) w[c][e] = w_s[c * 64 + j * 8 + e];
  }
  const int HW = g.H * g.W;
  const int NP = g.N * HW;
  const bool small = NP < (1 << 22);
  const float inv_hw = 1.0f / static_cast<float>(HW), inv_w = 1.0f / static_cast<float>(g.W);
  for (int q = (blockIdx.x * 256 + threadIdx.x) >> 3; q < NP; q += gridDim.x * 32) {
    const int n = fdiv22(q, HW, inv_hw, small);
    const int p = q - n * HW;
    const int yy = fdiv22(p, g.W, inv_w, small);
    const int xx = p - yy * g.W;
    const size_t R = static_cast<size_t>(g.HALO) + static_cast<size_t>(n) * g.IMG + (yy + 1) * g.P + (xx + 1);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = b[e];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float xm = x[static_cast<size_t>(q) * C + c] * mask_b(mask_type, yy, xx, c, C);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(xm, w[c][e], acc[e]);
    }
    *reinterpret_cast<uint4*>(out + slab_chunk_off(R, j)) =
        make_uint4(pack2f(elu_f(acc[0]), elu_f(acc[1])), pack2f(elu_f(acc[2]), elu_f(acc[3])),
                   pack2f(elu_f(acc[4]), elu_f(acc[5])), pack2f(elu_f(acc[6]), elu_f(acc[7])));
  }
}

// c1 backward: one block per image.  g1 slab = dL/d(pre-activation of c1).
//   dx[n,y,x,c] += b * sum_co g1[co] W1[c][co];  dW1[c][co] += sum_p xm[p][c] g1[p][co];  db1[co] += sum_p g1[p][co]
__global__ void __launch_bounds__(256)
c1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ V1, const float* __restrict__ g1w,
              const uint8_t* __restrict__ gslab, float* __restrict__ dx, float* __restrict__ dW1, float* __restrict__ db1,
              SlabGeom g, int C, int mask_type, int KS) {
  extern __shared__ float sm[];
  const int K = KS * KS * C;             // rows of the weight matrix
  const int KP = K + 1;                  // + the constant-one column that yields db1
  float* w_s = sm;                       // [K][64]
  float* g_s = w_s + MAXK * 64;          // [256][65]
  float* x_s = g_s + 256 * 65;           // [256][KP]
  c1_weights_to_smem(V1, g1w, K, w_s);
  const int n = blockIdx.x;
  const int HW = g.H * g.W;
  const int nout = KP * 64;
  float acc[(MAXK + 1) * 64 / 256 + 1];
#pragma unroll
  for (int i = 0; i < (MAXK + 1) * 64 / 256 + 1; ++i) acc[i] = 0.f;
  for (int p0 = 0; p0 < HW; p0 += 256) {
    const int p = p0 + threadIdx.x;
    const bool ok = p < HW;
    float gv[64];
    if (ok) {
      const int yy = p / g.W, xx = p % g.W;
      const size_t R = static_cast<size_t>(g.HALO) + static_cast<size_t>(n) * g.IMG + (yy + 1) * g.P + (xx + 1);
      const size_t xo = (static_cast<size_t>(n) * HW + p) * C;
      if (KS == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 u = *reinterpret_cast<const uint4*>(gslab + slab_chunk_off(R, j));
          unpack8f(u, gv + j * 8);
        }
        for (int c = 0; c < C; ++c) {
          const float b = mask_b(mask_type, yy, xx, c, C);
          float d = 0.f;
#pragma unroll
          for (int k = 0; k < 64; ++k) d = fmaf(gv[k], w_s[c * 64 + k], d);
          dx[xo + c] += b * d;
          x_s[threadIdx.x * KP + c] = x[xo + c] * b;
        }
      } else {
        // 3x3: the pre-activation at q = p - shift(t) read input pixel p through tap t, so
        //   dx[p][c] = b(p,c) sum_t sum_co g[p - shift(t)][co] W1[t*C + c][co]
        // (gather: deterministic; rows of the gradient slab outside the image are its zero padding), and the row of the
        // im2col matrix for dW1 is xm at p + shift(t) (zero outside the image).
        float d[MAXC];
        for (int c = 0; c < C; ++c) d[c] = 0.f;
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3 - 1, dxx = t % 3 - 1;
          const size_t Rn = R - static_cast<size_t>(dy * g.P + dxx) ;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 u = *reinterpret_cast<const uint4*>(gslab + slab_chunk_off(Rn, j));
            unpack8f(u, gv + j * 8);
          }
          for (int c = 0; c < C; ++c) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 64; ++k) a = fmaf(gv[k], w_s[(t * C + c) * 64 + k], a);
            d[c] += a;
          }
          const int y2 = yy + dy, x2 = xx + dxx;
          const bool in = y2 >= 0 && y2 < g.H && x2 >= 0 && x2 < g.W;
          for (int c = 0; c < C; ++c)
            x_s[threadIdx.x * KP + t * C + c] =
                in ? x[((static_cast<size_t>(n) * g.H + y2) * g.W + x2) * C + c] * mask_b(mask_type, y2, x2, c, C) : 0.f;
        }
        for (int c = 0; c < C; ++c) dx[xo + c] += mask_b(mask_type, yy, xx, c, C) * d[c];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 u = *reinterpret_cast<const uint4*>(gslab + slab_chunk_off(R, j));
          unpack8f(u, gv + j * 8);
        }
      }
      x_s[threadIdx.x * KP + K] = 1.f;
    } else {
#pragma unroll
      for (int k = 0; k < 64; ++k) gv[k] = 0.f;
      for (int c = 0; c <= K; ++c) x_s[threadIdx.x * KP + c] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) g_s[threadIdx.x * 65 + k] = gv[k];
    __syncthreads();
    int ai = 0;
    for (int o = threadIdx.x; o < nout; o += 256, ++ai) {
      const int co = o & 63, c = o >> 6;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
      for (int pp = 0; pp < 256; pp += 4) {
        a0 = fmaf(x_s[pp * KP + c], g_s[pp * 65 + co], a0);
        a1 = fmaf(x_s[(pp + 1) * KP + c], g_s[(pp + 1) * 65 + co], a1);
        a2 = fmaf(x_s[(pp + 2) * KP + c], g_s[(pp + 2) * 65 + co], a2);
        a3 = fmaf(x_s[(pp + 3) * KP + c], g_s[(pp + 3) * 65 + co], a3);
      }
      acc[ai] += (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
  }
  int ai = 0;
  for (int o = threadIdx.x; o < nout; o += 256, ++ai) {
    const int co = o & 63, c = o >> 6;
    if (c < K) atomicAdd(dW1 + c * 64 + co, acc[ai]);
    else atomicAdd(db1 + co, acc[ai]);
  }
}

// c1 backward for C <= 4, eight lanes per pixel (lane j owns the 16-byte chunk j of the gradient slab row): coalesced
// 512-byte warp reads, dx through 8-channel partials + xor-shuffles, dW1 / db1 in registers over all pixels of the thread
// and ONE fixed-order block reduction + atomics at the end (see heads_coupling_bwd8_kernel).
template <int C>
__global__ void __launch_bounds__(256, C == 1 ? 3 : 2)
c1_bwd8_kernel(const float* __restrict__ x, const float* __restrict__ V1, const float* __restrict__ g1w,
               const uint8_t* __restrict__ gslab, float* __restrict__ dx, float* __restrict__ dW1, float* __restrict__ db1,
               SlabGeom g, int mask_type) {
  __shared__ float w_s[MAXC * 64];
  __shared__ float red[8][64][C + 2];         // [warp][channel][c or bias]
  c1_weights_to_smem(V1, g1w, C, w_s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = lane & 7;
  float w[C][8];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) w[c][e] = w_s[c * 64 + 8 * j + e];
  float acc[C + 1][8];
#pragma unroll
  for (int c = 0; c <= C; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
  const int HW = g.H * g.W;
  const int NP = g.N * HW;
  const int ngroups = gridDim.x * 32;
  const bool small = NP < (1 << 22);
  const float inv_hw = 1.0f / static_cast<float>(HW), inv_w = 1.0f / static_cast<float>(g.W);
  // the inputs of the NEXT pixel are fetched before the current one is processed: with one 16-byte load per thread and
  // iteration the kernel sat at ~1 TB/s (latency x 16 warps / SM)
  struct In { uint4 gq; float xv[C]; int yy, xx, q; };
  auto fetch = [&](int q) {
    In r;
    r.q = q;
    const int n = fdiv22(q, HW, inv_hw, small), p = q - n * HW;
    r.yy = fdiv22(p, g.W, inv_w, small);
    r.xx = p - r.yy * g.W;
    const size_t R = static_cast<size_t>(g.HALO) + static_cast<size_t>(n) * g.IMG + (r.yy + 1) * g.P + (r.xx + 1);
    r.gq = *reinterpret_cast<const uint4*>(gslab + slab_chunk_off(R, j));
#pragma unroll
    for (int c = 0; c < C; ++c) r.xv[c] = x[static_cast<size_t>(q) * C + c];
    return r;
  };
  int q0 = (blockIdx.x * 256 + threadIdx.x) >> 3;
  // C <= 2 has the registers for TWO pixels in flight per thread (and C = 1 for three blocks per SM): 16 resident warps with
  // one 16-byte load each ahead cover ~45 % of the bandwidth-delay product
  constexpr bool DEEP = C <= 2;
  In cur, nx1;
  if (q0 < NP) cur = fetch(q0);
  nx1 = cur;
  if (DEEP && q0 + ngroups < NP) nx1 = fetch(q0 + ngroups);
  for (int q = q0; q < NP; q += ngroups) {
    In nxt = DEEP ? nx1 : cur;
    if (DEEP) {
      if (q + 2 * ngroups < NP) nx1 = fetch(q + 2 * ngroups);
    } else {
      if (q + ngroups < NP) nxt = fetch(q + ngroups);
    }
    float gv[8];
    unpack8f(cur.gq, gv);
    const size_t xo = static_cast<size_t>(q) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float b = mask_b(mask_type, cur.yy, cur.xx, c, C);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(gv[e], w[c][e], d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      const float xm = cur.xv[c] * b;
      if (j == c) dx[xo + c] += b * d;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[c][e] = fmaf(xm, gv[e], acc[c][e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[C][e] += gv[e];
    cur = nxt;
  }
#pragma unroll
  for (int c = 0; c <= C; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = acc[c][e];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) red[warp][8 * j + e][c] = v;
    }
  __syncthreads();
  for (int o = threadIdx.x; o < (C + 1) * 64; o += 256) {
    const int co = o & 63, c = o >> 6;
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) t += red[wq][co][c];
    if (c < C) atomicAdd(dW1 + c * 64 + co, t);
    else atomicAdd(db1 + co, t);
  }
}

// ------------------------------------------------------------------------------------------------------------
// heads + coupling.  One block per image.
//   s = tanh(h . Ks + bs) (1-b),  t = (h . Kt + bt) (1-b)                                       nvp:129-139
//   forward : y = x b + (1-b)(x e^s + t),        ld += sum s                                      nvp:168-172
//   inverse : x = y b + (y (1-b) - t) e^{-s},    ld -= sum s                                      nvp:181-186
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
heads_coupling_kernel(const float* __restrict__ xin, const uint8_t* __restrict__ hslab, const float* __restrict__ Ks,
                      const float* __restrict__ bs, const float* __restrict__ Kt, const float* __restrict__ bt,
                      const float* __restrict__ label_s, const float* __restrict__ label_t,
                      const float* __restrict__ ld_in, float* __restrict__ yout, float* __restrict__ ld_out, SlabGeom g,
                      int C, int mask_type, int inverse, int f32_slab) {
  __shared__ float ks_s[64 * MAXC], kt_s[64 * MAXC], bs_s[MAXC], bt_s[MAXC], red[8];
  for (int i = threadIdx.x; i < 64 * C; i += 256) {
    ks_s[i] = Ks[i];
    kt_s[i] = Kt[i];
  }
  if (threadIdx.x < C) {
    bs_s[threadIdx.x] = bs[threadIdx.x];
    bt_s[threadIdx.x] = bt[threadIdx.x];
  }
  __syncthreads();
  const int n = blockIdx.x;
  const int HW = g.H * g.W;
  float ssum = 0.f;
  for (int p = threadIdx.x; p < HW; p += 256) {
    const int yy = p / g.W, xx = p % g.W;
    const size_t R = static_cast<size_t>(g.HALO) + static_cast<size_t>(n) * g.IMG + (yy + 1) * g.P + (xx + 1);
    float h[64];
    if (f32_slab) {
      const float4* hp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(hslab) + R * 64);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 t = hp[j];
        h[4 * j] = t.x; h[4 * j + 1] = t.y; h[4 * j + 2] = t.z; h[4 * j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 u = *reinterpret_cast<const uint4*>(hslab + slab_chunk_off(R, j));
        unpack8f(u, h + j * 8);
      }
    }
    const size_t xo = (static_cast<size_t>(n) * HW + p) * C;
    for (int c = 0; c < C; ++c) {
      const float b = mask_b(mask_type, yy, xx, c, C);
      const float xv = xin[xo + c];
      float out = xv;
      if (b == 0.f) {
        float sp = bs_s[c], tp = bt_s[c];
        if (label_s) {        // label-conditioned heads: + dense(y_label)[n, c]   (nn_extra_nvp_conditional.py:382-388)
          sp += label_s[n * C + c];
          tp += label_t[n * C + c];
        }
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          sp = fmaf(h[k], ks_s[k * C + c], sp);
          tp = fmaf(h[k], kt_s[k * C + c], tp);
        }
        const float s = tanhf(sp);
        ssum += s;
        out = inverse ? (xv - tp) * expf(-s) : fmaf(xv, expf(s), tp);
      }
      yout[xo + c] = out;
    }
  }
  const float tot = block_sum_256(ssum, red);
  if (threadIdx.x == 0) ld_out[n] = ld_in[n] + (inverse ? -tot : tot);
}

// Backward of the forward coupling.  Given dy, dld:
//   dx (direct part) = dy (b + (1-b) e^s);  dt = dy (1-b);  ds = dy (1-b) x e^s + dld[n]
//   d(s_pre) = ds (1 - tanh^2),  d(t_pre) = dt
//   dh = Ks d(s_pre) + Kt d(t_pre);   gslab = dh * ELU'(h)   (h = output of the last residual block)
//   dKs[k][c] += h[k] d(s_pre)[c] ...  (accumulated per block, then atomics)
__global__ void __launch_bounds__(256)
heads_coupling_bwd_kernel(const float* __restrict__ xin, const uint8_t* __restrict__ hslab, const float* __restrict__ Ks,
                          const float* __restrict__ bs, const float* __restrict__ Kt, const float* __restrict__ bt,
                          const float* __restrict__ dy, const float* __restrict__ dld, float* __restrict__ dx,
                          uint8_t* __restrict__ gslab, float* __restrict__ partial, SlabGeom g, int C, int mask_type,
                          const float* __restrict__ label_s, float* __restrict__ dlabel_s, float* __restrict__ dlabel_t) {
  extern __shared__ float sm[];
  float* ks_s = sm;                        // [64][C]
  float* kt_s = ks_s + 64 * MAXC;
  float* bs_s = kt_s + 64 * MAXC;
  float* bt_s = bs_s + MAXC;
  float* h_s = bt_s + MAXC;                // [256][65]  (column 64 = 1 -> bias gradient)
  float* d_s = h_s + 256 * 65;             // [256][2*MAXC + 1]  (odd row stride: the per-thread rows are bank-conflict free)
  for (int i = threadIdx.x; i < 64 * C; i += 256) {
    ks_s[i] = Ks[i];
    kt_s[i] = Kt[i];
  }
  if (threadIdx.x < C) {
    bs_s[threadIdx.x] = bs[threadIdx.x];
    bt_s[threadIdx.x] = bt[threadIdx.x];
  }
  __syncthreads();
  const int HW = g.H * g.W;
  const int nout = 65 * 2 * C;
  constexpr int NACC = 65 * 2 * MAXC / 256 + 1;
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;

  // persistent over images: the head-weight gradients are accumulated in registers across all images of the block and
  // leave as ONE partial vector per block (640 blocks x same-address atomics serialised for ~200 us at L2 before)
  for (int n = blockIdx.x; n < g.N; n += gridDim.x) {
  const float dl = dld[n];
  // label-conditioned heads (nn_extra_nvp_conditional.py:382-388): the per-image terms label_s/t[n, c] enter the
  // pre-activations, their gradients are the per-image sums of d(s_pre), d(t_pre)
  float lsum[2 * MAXC];
  if (dlabel_s) {
#pragma unroll
    for (int c = 0; c < 2 * MAXC; ++c) lsum[c] = 0.f;
  }
  for (int p0 = 0; p0 < HW; p0 += 256) {
    const int p = p0 + threadIdx.x;
    const bool ok = p < HW;
    float h[64];
    float* dpre = d_s + threadIdx.x * (2 * MAXC + 1);
    if (ok) {
      const int yy = p / g.W, xx = p % g.W;
      const size_t R = static_cast<size_t>(g.HALO) + static_cast<size_t>(n) * g.IMG + (yy + 1) * g.P + (xx + 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 u = *reinterpret_cast<const uint4*>(hslab + slab_chunk_off(R, j));
        unpack8f(u, h + j * 8);
      }
      const size_t xo = (static_cast<size_t>(n) * HW + p) * C;
      for (int c = 0; c < C; ++c) {
        const float b = mask_b(mask_type, yy, xx, c, C);
        const float g_y = dy[xo + c];
        float dsp = 0.f, dtp = 0.f, dxv = g_y;
        if (b == 0.f) {
          float sp = bs_s[c] + (label_s ? label_s[n * C + c] : 0.f);
#pragma unroll
          for (int k = 0; k < 64; ++k) sp = fmaf(h[k], ks_s[k * C + c], sp);
          const float s = tanhf(sp);
          const float es = expf(s);
          dxv = g_y * es;
          dtp = g_y;
          dsp = (g_y * xin[xo + c] * es + dl) * (1.f - s * s);
        }
        dx[xo + c] = dxv;
        dpre[c] = dsp;
        dpre[C + c] = dtp;
      }
      if (dlabel_s) {
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) {
            lsum[c] += dpre[c];
            lsum[MAXC + c] += dpre[C + c];
          }
      }
      // dh and the gradient w.r.t. the pre-activation of the last residual block
      uint32_t packed[32];
#pragma unroll
      for (int k = 0; k < 64; k += 2) {
        float d0 = 0.f, d1 = 0.f;
        for (int c = 0; c < C; ++c) {
          d0 = fmaf(dpre[c], ks_s[k * C + c], fmaf(dpre[C + c], kt_s[k * C + c], d0));
          d1 = fmaf(dpre[c], ks_s[(k + 1) * C + c], fmaf(dpre[C + c], kt_s[(k + 1) * C + c], d1));
        }
        packed[k >> 1] = pack2f(d0 * elu_grad_from_out(h[k]), d1 * elu_grad_from_out(h[k + 1]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(gslab + slab_chunk_off(R, j)) =
            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < 64; ++k) h[k] = 0.f;
      for (int c = 0; c < 2 * C; ++c) dpre[c] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) h_s[threadIdx.x * 65 + k] = h[k];
    h_s[threadIdx.x * 65 + 64] = ok ? 1.f : 0.f;
    __syncthreads();
    int ai = 0;
    for (int o = threadIdx.x; o < nout; o += 256, ++ai) {
      const int k = o % 65, c2 = o / 65;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // independent chains: the loop is LDS-latency bound otherwise
#pragma unroll 4
      for (int pp = 0; pp < 256; pp += 4) {
        a0 = fmaf(h_s[pp * 65 + k], d_s[pp * (2 * MAXC + 1) + c2], a0);
        a1 = fmaf(h_s[(pp + 1) * 65 + k], d_s[(pp + 1) * (2 * MAXC + 1) + c2], a1);
        a2 = fmaf(h_s[(pp + 2) * 65 + k], d_s[(pp + 2) * (2 * MAXC + 1) + c2], a2);
        a3 = fmaf(h_s[(pp + 3) * 65 + k], d_s[(pp + 3) * (2 * MAXC + 1) + c2], a3);
      }
      acc[ai] += (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
  }
  if (dlabel_s) {
    // fixed shuffle tree + fixed-order sum of the 8 warp totals (h_s is free between images)
#pragma unroll
    for (int c = 0; c < 2 * MAXC; ++c) {
      float v = lsum[c];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) h_s[(threadIdx.x >> 5) * (2 * MAXC) + c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 2 * MAXC) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += h_s[w * (2 * MAXC) + threadIdx.x];
      const int c = threadIdx.x & (MAXC - 1);
      if (c < C) (threadIdx.x < MAXC ? dlabel_s : dlabel_t)[n * C + c] = t;
    }
    __syncthreads();
  }
  }
  int ai = 0;
  for (int o = threadIdx.x; o < nout; o += 256, ++ai) partial[static_cast<size_t>(blockIdx.x) * nout + o] = acc[ai];
}

// Per-image channel sums of L stacked gradient slabs: out[n, l*64 + c] = sum over the rows of image n of slab l.
// (gradient of the label-conditioned per-image bias tables, nn_extra_nvp_conditional.py:422-433.)  grid (N, L).
__global__ void __launch_bounds__(256)
slab_image_colsum_kernel(const uint8_t* __restrict__ slabs, size_t slab_stride, SlabGeom g, float* __restrict__ out, int out_stride) {
  __shared__ float red[32][65];
  const int n = blockIdx.x, l = blockIdx.y;
  const uint8_t* slab = slabs + static_cast<size_t>(l) * slab_stride;
  const int j = threadIdx.x & 7, rg = threadIdx.x >> 3;       // 16-byte chunk (8 channels), row group
  float a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = 0.f;
  for (int q = rg; q < g.IMG; q += 32) {
    const size_t R = static_cast<size_t>(g.HALO) + static_cast<size_t>(n) * g.IMG + q;
    float v[8];
    unpack8f(*reinterpret_cast<const uint4*>(slab + slab_chunk_off(R, j)), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += v[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rg][j * 8 + e] = a[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
    for (int r = 0; r < 32; ++r) t += red[r][threadIdx.x];
    out[static_cast<size_t>(n) * out_stride + l * 64 + threadIdx.x] = t;
  }
}

// The same backward for C <= 4 (all layers of bn2 / en2; the coupling at C channels has 2C head outputs), eight lanes per
// pixel: lane j of a pixel's group owns the 16-byte chunk j (8 channels) of the slab row.
//   * a warp instruction reads / writes 4 consecutive rows = 512 contiguous bytes (thread-per-pixel touched 32 lines
//     with 16 bytes each: LSU-bound at ~10 us per 256 pixels);
//   * the head dot products are 8-channel partials + three xor-shuffles;
//   * the head-weight gradients dK[k][c2] = sum_p h[p][k] d[p][c2] stay in REGISTERS (8 channels x 2C per thread) over
//     all pixels of the thread: the shared-memory transpose + dot-product phase of the generic kernel (19 us of LDS per
//     14x14 layer) is gone.  One fixed-order reduction per block at the end -> `partial`, same layout as above.
//   * C >= 3 needs 174 / 212 registers: one block per SM without spills beats two with 176 bytes of stack (43.0 -> 37.8 us
//     at 640x14x14, C = 4)
template <int C>
__global__ void __launch_bounds__(256, C >= 3 ? 1 : 2)
heads_coupling_bwd8_kernel(const float* __restrict__ xin, const uint8_t* __restrict__ hslab, const float* __restrict__ Ks,
                           const float* __restrict__ bs, const float* __restrict__ Kt, const float* __restrict__ dy,
                           const float* __restrict__ dld, float* __restrict__ dx, uint8_t* __restrict__ gslab,
                           float* __restrict__ partial, SlabGeom g, int mask_type, const float* __restrict__ label_s,
                           float* __restrict__ dlabel_acc /* [N, 2C], zeroed, or null */) {
  __shared__ float red[8][64][2 * C + 1];   // per-warp sums: [warp][channel][c2]; column 2C unused (bank spread)
  __shared__ float redb[8][2 * C];
  // head kernels by chunk: [j][e][c] with an odd pitch per chunk (the 8 chunk lanes of a pixel hit 8 different banks)
  constexpr int KP = 8 * C + 1;
  __shared__ float ks_s[8 * KP], kt_s[8 * KP];
  for (int i = threadIdx.x; i < 64 * C; i += 256) {
    const int k = i / C, c = i % C;
    ks_s[(k >> 3) * KP + (k & 7) * C + c] = Ks[i];
    kt_s[(k >> 3) * KP + (k & 7) * C + c] = Kt[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = lane & 7;                     // chunk = channels 8j .. 8j+7
  const float* ks = ks_s + j * KP;            // ks[e * C + c]
  const float* kt = kt_s + j * KP;
  float bsv[C];
#pragma unroll
  for (int c = 0; c < C; ++c) bsv[c] = bs[c];
  float acc[8][2 * C], accb[2 * C];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int c = 0; c < 2 * C; ++c) acc[e][c] = 0.f;
#pragma unroll
  for (int c = 0; c < 2 * C; ++c) accb[c] = 0.f;

  const int HW = g.H * g.W;
  const int NP = g.N * HW;                    // < 2^31 (the slab geometry already requires it)
  const int ngroups = gridDim.x * 32;
  const bool small = NP < (1 << 22);
  const float inv_hw = 1.0f / static_cast<float>(HW), inv_w = 1.0f / static_cast<float>(g.W);
  // inputs of the NEXT pixel are fetched before the current one is processed (see c1_bwd8_kernel)
  struct In { uint4 hq; float gy[C], xv[C], ls[C], dl; int n, yy, xx; size_t R; };
  auto fetch = [&](int q) {
    In r;
    r.n = fdiv22(q, HW, inv_hw, small);
    const int p = q - r.n * HW;
    r.yy = fdiv22(p, g.W, inv_w, small);
    r.xx = p - r.yy * g.W;
    r.R = static_cast<size_t>(g.HALO) + static_cast<size_t>(r.n) * g.IMG + (r.yy + 1) * g.P + (r.xx + 1);
    r.hq = *reinterpret_cast<const uint4*>(hslab + slab_chunk_off(r.R, j));
    r.dl = dld[r.n];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      r.gy[c] = dy[static_cast<size_t>(q) * C + c];
      r.xv[c] = xin[static_cast<size_t>(q) * C + c];
      r.ls[c] = label_s ? label_s[r.n * C + c] : 0.f;
    }
    return r;
  };
  const int q0 = (blockIdx.x * 256 + threadIdx.x) >> 3;
  In cur;
  if (q0 < NP) cur = fetch(q0);
  for (int q = q0; q < NP; q += ngroups) {
    In nxt = cur;
    if (q + ngroups < NP) nxt = fetch(q + ngroups);       // (two pixels ahead was measured: no gain here, the chain is the math)
    const int n = cur.n, yy = cur.yy, xx = cur.xx;
    const size_t R = cur.R;
    float h[8];
    unpack8f(cur.hq, h);
    const size_t xo = static_cast<size_t>(q) * C;
    const float dl = cur.dl;
    // scale-head pre-activation: 8-channel partial, summed over the 8 lanes of the pixel
    float sp[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float v = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) v = fmaf(h[e], ks[e * C + c], v);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      sp[c] = v + bsv[c] + cur.ls[c];
    }
    float d[2 * C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float b = mask_b(mask_type, yy, xx, c, C);
      const float g_y = cur.gy[c];
      float dsp = 0.f, dtp = 0.f, dxv = g_y;
      if (b == 0.f) {
        const float s_ = tanhf(sp[c]);
        const float es = expf(s_);
        dxv = g_y * es;
        dtp = g_y;
        dsp = (g_y * cur.xv[c] * es + dl) * (1.f - s_ * s_);
      }
      if (j == c) dx[xo + c] = dxv;
      d[c] = dsp;
      d[C + c] = dtp;
    }
    if (dlabel_acc && j == 0) {
#pragma unroll
      for (int c = 0; c < 2 * C; ++c) atomicAdd(dlabel_acc + static_cast<size_t>(n) * 2 * C + c, d[c]);
    }
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) v = fmaf(d[c], ks[e * C + c], fmaf(d[C + c], kt[e * C + c], v));
      o[e] = v * elu_grad_from_out(h[e]);
#pragma unroll
      for (int c = 0; c < 2 * C; ++c) acc[e][c] = fmaf(h[e], d[c], acc[e][c]);
    }
    if (j == 0) {
#pragma unroll
      for (int c = 0; c < 2 * C; ++c) accb[c] += d[c];
    }
    *reinterpret_cast<uint4*>(gslab + slab_chunk_off(R, j)) =
        make_uint4(pack2f(o[0], o[1]), pack2f(o[2], o[3]), pack2f(o[4], o[5]), pack2f(o[6], o[7]));
    cur = nxt;
  }
  // block reduction in a fixed order: the 4 pixel groups of a warp (xor 8, 16), then the 8 warps through shared memory
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int c = 0; c < 2 * C; ++c) {
      float v = acc[e][c];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) red[warp][8 * j + e][c] = v;
    }
#pragma unroll
  for (int c = 0; c < 2 * C; ++c) {
    float v = accb[c];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (lane == 0) redb[warp][c] = v;
  }
  __syncthreads();
  const int nout = 65 * 2 * C;
  for (int o = threadIdx.x; o < nout; o += 256) {
    const int k = o % 65, c2 = o / 65;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += k < 64 ? red[w][k][c2] : redb[w][c2];
    partial[static_cast<size_t>(blockIdx.x) * nout + o] = t;
  }
}

// fixed-order sum of the per-block partials -> dKs, dbs, dKt, dbt (overwritten)
__global__ void __launch_bounds__(256)
heads_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ dKs,
                          float* __restrict__ dbs, float* __restrict__ dKt, float* __restrict__ dbt) {
  // one warp per output: lane l adds blocks l, l+32, ... in order, then a fixed xor-shuffle tree (deterministic; a single
  // thread per output walked the ~300 partials as 74 dependent L2 round trips = 15 us)
  const int nout = 65 * 2 * C;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= nout) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int b = lane; b < nblk; b += 32) s += partial[static_cast<size_t>(b) * nout + o];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane != 0) return;
  const int k = o % 65, c2 = o / 65;
  const bool is_t = c2 >= C;
  const int c = is_t ? c2 - C : c2;
  float* dst = k < 64 ? (is_t ? dKt : dKs) + k * C + c : (is_t ? dbt : dbs) + c;
  *dst = s;
}

// fp32 verification path: epilogue of a 3x3 convolution whose nine taps were accumulated by bruno_sgemm into
// acc [rows][64] fp32:  out = act(acc + bias (+ skip)) on pixel rows, 0 on padding rows (keeps the slab invariant).
__global__ void conv_epilogue_f32_kernel(float* __restrict__ acc, const float* __restrict__ bias,
                                         const float* __restrict__ skip, SlabGeom g, int elu) {
  const size_t total = static_cast<size_t>(g.NT) * TILE_M * 16;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int data_row = static_cast<int>(i >> 4), j = static_cast<int>(i & 15);
    float4* p = reinterpret_cast<float4*>(acc + (static_cast<size_t>(g.HALO) + data_row) * 64 + j * 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slab_row_is_pixel(data_row, g)) {
      v = *p;
      const float4 b = *reinterpret_cast<const float4*>(bias + j * 4);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      if (skip) {
        const float4 s = *reinterpret_cast<const float4*>(skip + (static_cast<size_t>(g.HALO) + data_row) * 64 + j * 4);
        v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
      }
      if (elu) {
        v.x = v.x > 0.f ? v.x : expm1f(v.x); v.y = v.y > 0.f ? v.y : expm1f(v.y);
        v.z = v.z > 0.f ? v.z : expm1f(v.z); v.w = v.w > 0.f ? v.w : expm1f(v.w);
      }
    }
    *p = v;
  }
}

}  // namespace bruno

using namespace bruno;

extern "C" {

int bruno_pre(int op, const float* x, const float* ld_in, float* y, float* ld_out, int N, int D, float scale, float alpha,
              int conditional_variant, void* stream) {
  BRUNO_REQUIRE(op >= 0 && op <= 5 && x && y && ld_out && N > 0 && D > 0, "bruno_pre: bad arguments");
  pre_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(op, x, ld_in, y, ld_out, D, scale, alpha, conditional_variant);
  return check_launch("pre_kernel");
}

int bruno_space_to_depth(const float* in, float* out, int N, int H, int W, int C, int inverse, void* stream) {
  BRUNO_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0,
                "bruno_space_to_depth: bad arguments (H, W are the LARGE resolution and must be even)");
  const size_t total = static_cast<size_t>(N) * H * W * C;
  space_to_depth_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, out, N, H, W, C, inverse);
  return check_launch("space_to_depth_kernel");
}

int bruno_channel_copy(const float* src, float* dst, long rows, int src_C, int src_off, int dst_C, int dst_off, int nch,
                       void* stream) {
  BRUNO_REQUIRE(src && dst && rows > 0 && nch > 0 && src_off + nch <= src_C && dst_off + nch <= dst_C,
                "bruno_channel_copy: bad arguments");
  const size_t total = static_cast<size_t>(rows) * nch;
  channel_copy_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, dst, static_cast<size_t>(rows), src_C, src_off, dst_C, dst_off, nch);
  return check_launch("channel_copy_kernel");
}

int bruno_c1_fwd(const float* x, const float* V1, const float* g1, const float* b1, int bias_img_stride, void* out_slab,
                 int f32_slab, int N, int H, int W, int C, int mask_type, int ks, void* stream) {
  BRUNO_REQUIRE(x && V1 && g1 && b1 && out_slab && C >= 1 && C <= MAXC && (ks == 1 || ks == 3) && ks * ks * C <= MAXK,
                "bruno_c1_fwd: bad arguments (C <= %d, ks in {1,3}, ks*ks*C <= %d)", MAXC, MAXK);
  const SlabGeom g = slab_geom(N, H, W);
  if (ks == 1 && C <= 4 && !bias_img_stride && !f32_slab) {      // eight lanes per pixel, weights in registers
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* o = static_cast<uint8_t*>(out_slab);
    const long long groups = (static_cast<long long>(N) * H * W + 31) / 32;
    const int nb = static_cast<int>(groups < 148 * 8 ? groups : 148 * 8);
    if (C == 1) c1_fwd8_kernel<1><<<nb, 256, 0, st>>>(x, V1, g1, b1, o, g, mask_type);
    else if (C == 2) c1_fwd8_kernel<2><<<nb, 256, 0, st>>>(x, V1, g1, b1, o, g, mask_type);
    else if (C == 3) c1_fwd8_kernel<3><<<nb, 256, 0, st>>>(x, V1, g1, b1, o, g, mask_type);
    else c1_fwd8_kernel<4><<<nb, 256, 0, st>>>(x, V1, g1, b1, o, g, mask_type);
    return check_launch("c1_fwd8_kernel");
  }
  const int total = g.NT * TILE_M * 8;
  int blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  c1_fwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, V1, g1, b1, bias_img_stride,
                                                                        static_cast<uint8_t*>(out_slab), g, C, mask_type, f32_slab, ks);
  return check_launch("c1_fwd_kernel");
}

int bruno_c1_bwd(const float* x, const float* V1, const float* g1, const void* g_slab, float* dx, float* dW1, float* db1,
                 int N, int H, int W, int C, int mask_type, int ks, void* stream) {
  BRUNO_REQUIRE(x && V1 && g1 && g_slab && dx && dW1 && db1 && C >= 1 && C <= MAXC && (ks == 1 || ks == 3) &&
                    ks * ks * C <= MAXK, "bruno_c1_bwd: bad arguments");
  const SlabGeom g = slab_geom(N, H, W);
  const size_t smem = (MAXK * 64 + 256 * 65 + 256 * (ks * ks * C + 1)) * sizeof(float);
  static OncePerDevice once;      // the opt-in is per device; not per call (the call may run under stream capture)
  once.run([&] {
    cudaFuncSetAttribute(c1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>((MAXK * 64 + 256 * 65 + 256 * (MAXK + 1)) * sizeof(float)));
  });
  if (C <= 4 && ks == 1) {
    cudaStream_t st8 = static_cast<cudaStream_t>(stream);
    const SlabGeom g8 = slab_geom(N, H, W);
    const uint8_t* gs8 = static_cast<const uint8_t*>(g_slab);
    if (C == 1) c1_bwd8_kernel<1><<<444, 256, 0, st8>>>(x, V1, g1, gs8, dx, dW1, db1, g8, mask_type);
    else if (C == 2) c1_bwd8_kernel<2><<<296, 256, 0, st8>>>(x, V1, g1, gs8, dx, dW1, db1, g8, mask_type);
    else if (C == 3) c1_bwd8_kernel<3><<<296, 256, 0, st8>>>(x, V1, g1, gs8, dx, dW1, db1, g8, mask_type);
    else c1_bwd8_kernel<4><<<296, 256, 0, st8>>>(x, V1, g1, gs8, dx, dW1, db1, g8, mask_type);
    return check_launch("c1_bwd8_kernel");
  }
  c1_bwd_kernel<<<N, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, V1, g1, static_cast<const uint8_t*>(g_slab), dx, dW1,
                                                                      db1, g, C, mask_type, ks);
  return check_launch("c1_bwd_kernel");
}

int bruno_conv_epilogue_f32(float* acc, const float* bias, const float* skip, int elu, int N, int H, int W, void* stream) {
  BRUNO_REQUIRE(acc && bias && N > 0 && H > 0 && W > 0, "bruno_conv_epilogue_f32: bad arguments");
  const SlabGeom g = slab_geom(N, H, W);
  conv_epilogue_f32_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(acc, bias, skip, g, elu);
  return check_launch("conv_epilogue_f32_kernel");
}

int bruno_heads_coupling(const float* x, const void* h_slab, const float* Ks, const float* bs, const float* Kt,
                         const float* bt, const float* label_s, const float* label_t, const float* ld_in, float* y,
                         float* ld_out, int f32_slab, int N, int H, int W, int C, int mask_type, int inverse, void* stream) {
  BRUNO_REQUIRE((label_s == nullptr) == (label_t == nullptr), "bruno_heads_coupling: label_s / label_t come together");
  BRUNO_REQUIRE(x && h_slab && Ks && bs && Kt && bt && ld_in && y && ld_out && C >= 1 && C <= MAXC,
                "bruno_heads_coupling: bad arguments");
  const SlabGeom g = slab_geom(N, H, W);
  // (an eight-lanes-per-pixel forward like heads_coupling_bwd8 was measured: 35 / 29 / 17 us vs 32 / 16 / 16 us for this
  // kernel at 640x28x28 C=1, 640x14x14 C=4 / C=2 -- the extra shuffles and the eightfold tanh / exp cost more than the
  // coalescing returns; not kept)
  heads_coupling_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<const uint8_t*>(h_slab), Ks, bs, Kt, bt, label_s, label_t, ld_in, y, ld_out, g, C, mask_type, inverse, f32_slab);
  return check_launch("heads_coupling_kernel");
}

size_t bruno_heads_bwd_scratch_floats(int C) { return static_cast<size_t>(HEADS_BWD_BLOCKS) * 65 * 2 * C; }

int bruno_heads_coupling_bwd(const float* x, const void* h_slab, const float* Ks, const float* bs, const float* Kt,
                             const float* bt, const float* dy, const float* dld, float* dx, void* g_slab, float* dKs,
                             float* dbs, float* dKt, float* dbt, float* scratch, int N, int H, int W, int C, int mask_type,
                             void* stream) {
  return bruno_heads_coupling_bwd_cond(x, h_slab, Ks, bs, Kt, bt, dy, dld, dx, g_slab, dKs, dbs, dKt, dbt, scratch, N, H, W, C,
                                       mask_type, nullptr, nullptr, nullptr, stream);
}

int bruno_slab_image_colsum(const void* slabs, int L, float* out, int out_stride, int N, int H, int W, void* stream) {
  BRUNO_REQUIRE(slabs && out && L >= 1 && out_stride >= L * 64, "bruno_slab_image_colsum: bad arguments");
  const SlabGeom g = slab_geom(N, H, W);
  slab_image_colsum_kernel<<<dim3(N, L), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(slabs), static_cast<size_t>(g.rows) * SLAB_ROW_BYTES, g, out, out_stride);
  return check_launch("slab_image_colsum_kernel");
}

int bruno_heads_coupling_bwd_cond(const float* x, const void* h_slab, const float* Ks, const float* bs, const float* Kt,
                                  const float* bt, const float* dy, const float* dld, float* dx, void* g_slab, float* dKs,
                                  float* dbs, float* dKt, float* dbt, float* scratch, int N, int H, int W, int C,
                                  int mask_type, const float* label_s, float* dlabel_s, float* dlabel_t, void* stream) {
  BRUNO_REQUIRE((dlabel_s == nullptr) == (dlabel_t == nullptr), "bruno_heads_coupling_bwd: dlabel_s and dlabel_t go together");
  BRUNO_REQUIRE(x && h_slab && Ks && bs && Kt && bt && dy && dld && dx && g_slab && dKs && dbs && dKt && dbt && scratch &&
                    C >= 1 && C <= MAXC, "bruno_heads_coupling_bwd: bad arguments");
  const SlabGeom g = slab_geom(N, H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = (2 * 64 * MAXC + 2 * MAXC + 256 * 65 + 256 * (2 * MAXC + 1)) * sizeof(float);
  static OncePerDevice once;
  once.run([&] {
    cudaFuncSetAttribute(heads_coupling_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  });
  // (no zeroing of g_slab's padding rows here: slabs are zero-initialised once by the caller and every writer -- this
  // kernel, the convolution epilogues -- either skips the padding rows or writes zeros there; bruno_sm100.h, slab contract.
  // The per-call slab_zero_padding_kernel launch cost 0.1 ms per training step.)
  if (C <= 4 && !dlabel_s) {
    // eight lanes per pixel, head-weight gradients in registers (see heads_coupling_bwd8_kernel)
    int nb = 0;
#define HB8(CC)                                                                                                           \
  do {                                                                                                                    \
    static int per_sm = 0;                                                                                                \
    if (per_sm == 0) {                                                                                                    \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, heads_coupling_bwd8_kernel<CC>, 256, 0);                     \
      if (per_sm < 1) per_sm = 1;                                                                                         \
    }                                                                                                                     \
    nb = 148 * per_sm < HEADS_BWD_BLOCKS ? 148 * per_sm : HEADS_BWD_BLOCKS;                                               \
    heads_coupling_bwd8_kernel<CC><<<nb, 256, 0, st>>>(x, static_cast<const uint8_t*>(h_slab), Ks, bs, Kt, dy, dld, dx,   \
                                                       static_cast<uint8_t*>(g_slab), scratch, g, mask_type, label_s,     \
                                                       nullptr);                                                          \
  } while (0)
    if (C == 1) HB8(1);
    else if (C == 2) HB8(2);
    else if (C == 3) HB8(3);
    else HB8(4);
#undef HB8
    int rc8 = check_launch("heads_coupling_bwd8_kernel");
    if (rc8) return rc8;
    heads_bwd_finalize_kernel<<<ceil_div(65 * 2 * C, 8), 256, 0, st>>>(scratch, nb, C, dKs, dbs, dKt, dbt);
    return check_launch("heads_bwd_finalize_kernel");
  }
  const int nblk = N < HEADS_BWD_BLOCKS ? N : HEADS_BWD_BLOCKS;
  heads_coupling_bwd_kernel<<<nblk, 256, smem, st>>>(x, static_cast<const uint8_t*>(h_slab), Ks, bs, Kt, bt, dy, dld, dx,
                                                     static_cast<uint8_t*>(g_slab), scratch, g, C, mask_type, label_s, dlabel_s,
                                                     dlabel_t);
  int rc = check_launch("heads_coupling_bwd_kernel");
  if (rc) return rc;
  heads_bwd_finalize_kernel<<<ceil_div(65 * 2 * C, 8), 256, 0, st>>>(scratch, nblk, C, dKs, dbs, dKt, dbt);
  return check_launch("heads_bwd_finalize_kernel");
}

}  // extern "C"
