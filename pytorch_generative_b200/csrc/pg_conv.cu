// pg_conv.cu — CausalConv2d input layers with tiny Cin (1 or 3 image channels):
// reference nn/convolution.py:41-43 called from pixel_cnn.py:86-92, pixel_snail.py:155-161,
// image_gpt.py:87-93.  K = Cin*kh*kw is 9..49, far below a tensor-core K slab, and the layer is
// <= 0.01 % of the step's FLOPs at the ImageGPT/PixelSNAIL configs, so this is a direct CUDA-core kernel
// that reads the NCHW fp32 image and writes the pixel-major activation the GEMM path consumes.
// Masked taps are zero in `w` (the caller zeroes the Parameter in place, as the reference does), so the
// forward simply runs all kh*kw taps; wgrad is dense over the taps, matching autograd in the reference.
#include "../../include/pg_b200.h"
#include "pg_common.cuh"

namespace {

constexpr int PIX_PER_BLOCK = 32;
constexpr int MAX_K = 160;  // Cin*kh*kw upper bound (3*7*7 = 147)

struct ConvArgs {
  int N, Cin, H, W, Cout, kh, kw, ph, pw, K;
  int pre_act;  // activation applied to the input before the convolution (act(0) = 0, so it commutes with padding)
};

// Gathers the K-vector of input values under the kernel window of pixel p (zero padding).
__device__ __forceinline__ float patch_value(const float* __restrict__ x, const ConvArgs& a, int n, int y, int xx, int k) {
  const int ci = k / (a.kh * a.kw);
  const int r = k % (a.kh * a.kw);
  const int i = r / a.kw, j = r % a.kw;
  const int yy = y + i - a.ph, xc = xx + j - a.pw;
  if (yy < 0 || yy >= a.H || xc < 0 || xc >= a.W) return 0.f;
  return pg_act_fwd(a.pre_act, x[(((size_t)n * a.Cin + ci) * a.H + yy) * a.W + xc]);
}

__global__ void __launch_bounds__(256)
conv_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const ConvArgs a, float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int act_bf16) {
  __shared__ float patch[PIX_PER_BLOCK][MAX_K + 1];
  const int HW = a.H * a.W;
  const long long P = (long long)a.N * HW;
  const long long p0 = (long long)blockIdx.x * PIX_PER_BLOCK;
  for (int t = threadIdx.x; t < PIX_PER_BLOCK * a.K; t += blockDim.x) {
    const int pi = t / a.K, k = t % a.K;
    const long long p = p0 + pi;
    float v = 0.f;
    if (p < P) {
      const int n = (int)(p / HW), rem = (int)(p % HW);
      v = patch_value(x, a, n, rem / a.W, rem % a.W, k);
    }
    patch[pi][k] = v;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < a.Cout; co += blockDim.x) {
    float acc[PIX_PER_BLOCK];
    const float b = bias ? bias[co] : 0.f;
#pragma unroll
    for (int pi = 0; pi < PIX_PER_BLOCK; ++pi) acc[pi] = b;
    const float* wr = w + (size_t)co * a.K;
    for (int k = 0; k < a.K; ++k) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int pi = 0; pi < PIX_PER_BLOCK; ++pi) acc[pi] = fmaf(wv, patch[pi][k], acc[pi]);
    }
#pragma unroll
    for (int pi = 0; pi < PIX_PER_BLOCK; ++pi) {
      const long long p = p0 + pi;
      if (p < P) {
        if (out_f32) out_f32[p * a.Cout + co] = acc[pi];
        if (out_bf16) out_bf16[p * a.Cout + co] = __float2bfloat16(pg_act_fwd(act_bf16, acc[pi]));
      }
    }
  }
}

// wgrad: dw[co, k] += sum_p dy[p, co] * patch[p, k].  Persistent blocks; each thread owns a strided set of
// (co, k) outputs and accumulates over the block's pixel chunks, then one atomic per output.
constexpr int WG_PIX = 32;
constexpr int WG_MAX_OUT_PER_THREAD = 64;

__global__ void __launch_bounds__(256)
conv_small_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, const ConvArgs a,
                        float* __restrict__ dw, const int o_base) {
  extern __shared__ float smw[];
  float* patch = smw;                          // [WG_PIX][K]
  float* dys = smw + WG_PIX * a.K;             // [WG_PIX][Cout]
  const int HW = a.H * a.W;
  const long long P = (long long)a.N * HW;
  const int n_out = a.Cout * a.K;
  float acc[WG_MAX_OUT_PER_THREAD];
#pragma unroll
  for (int i = 0; i < WG_MAX_OUT_PER_THREAD; ++i) acc[i] = 0.f;
  for (long long p0 = (long long)blockIdx.x * WG_PIX; p0 < P; p0 += (long long)gridDim.x * WG_PIX) {
    __syncthreads();
    for (int t = threadIdx.x; t < WG_PIX * a.K; t += blockDim.x) {
      const int pi = t / a.K, k = t % a.K;
      const long long p = p0 + pi;
      float v = 0.f;
      if (p < P) {
        const int n = (int)(p / HW), rem = (int)(p % HW);
        v = patch_value(x, a, n, rem / a.W, rem % a.W, k);
      }
      patch[pi * a.K + k] = v;
    }
    for (int t = threadIdx.x; t < WG_PIX * a.Cout; t += blockDim.x) {
      const int pi = t / a.Cout, co = t % a.Cout;
      const long long p = p0 + pi;
      dys[pi * a.Cout + co] = p < P ? dy[p * a.Cout + co] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < WG_MAX_OUT_PER_THREAD; ++i) {
      const int o = o_base + threadIdx.x + i * blockDim.x;
      if (o < n_out) {
        const int co = o / a.K, k = o % a.K;
        float s = 0.f;
#pragma unroll 8
        for (int pi = 0; pi < WG_PIX; ++pi) s = fmaf(dys[pi * a.Cout + co], patch[pi * a.K + k], s);
        acc[i] += s;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < WG_MAX_OUT_PER_THREAD; ++i) {
    const int o = o_base + threadIdx.x + i * blockDim.x;
    if (o < n_out) atomicAdd(dw + o, acc[i]);
  }
}

// dgrad w.r.t. the input image: dx[n,ci,y,x] = act'(x) * sum_{co,i,j} dy[(y-i+ph, x-j+pw), co] * w[co,ci,i,j].
// The weight is staged once per block in shared memory as [tap][ci][co] (co contiguous: conflict-free, coalesced with
// the dy rows); one warp per input pixel, lanes over output channels.
__global__ void __launch_bounds__(256)
conv_small_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ dy, const ConvArgs a,
                        const float* __restrict__ x, float* __restrict__ dx) {
  extern __shared__ float wt[];  // [kh*kw][Cin][Cout]
  const int taps = a.kh * a.kw;
  for (int t = threadIdx.x; t < taps * a.Cin * a.Cout; t += blockDim.x) {
    const int co = t % a.Cout, ci = (t / a.Cout) % a.Cin, tap = t / (a.Cout * a.Cin);
    wt[t] = w[((size_t)co * a.Cin + ci) * taps + tap];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int HW = a.H * a.W;
  const long long P = (long long)a.N * HW;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; gw < P; gw += warps) {
    const int n = (int)(gw / HW), rem = (int)(gw % HW);
    const int y = rem / a.W, xx = rem % a.W;
    for (int c0 = 0; c0 < a.Cin; c0 += 4) {  // input channels in groups of 4 accumulators
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < a.kh; ++i) {
        const int yo = y - i + a.ph;
        if (yo < 0 || yo >= a.H) continue;
        for (int j = 0; j < a.kw; ++j) {
          const int xo = xx - j + a.pw;
          if (xo < 0 || xo >= a.W) continue;
          const float* dyr = dy + ((size_t)n * HW + (size_t)yo * a.W + xo) * a.Cout;
          const float* wr = wt + ((size_t)(i * a.kw + j) * a.Cin + c0) * a.Cout;
          for (int co = lane; co < a.Cout; co += 32) {
            const float d = dyr[co];
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
              if (c0 + ci < a.Cin) acc[ci] = fmaf(d, wr[ci * a.Cout + co], acc[ci]);
          }
        }
      }
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (c0 + ci >= a.Cin) break;
        const float v = warp_sum(acc[ci]);
        if (lane == 0) {
          const size_t off = (((size_t)n * a.Cin + c0 + ci) * a.H + y) * a.W + xx;
          dx[off] = a.pre_act == PG_ACT_NONE ? v : v * pg_act_bwd(a.pre_act, x[off]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tap gather / scatter for wide-channel convolutions (CausalConv2d with Cin >= 8, the GatedPixelCNN 1xN / Nx1 and
// PixelSNAIL 2x2 convs: reference gated_pixel_cnn.py:63-99,115,121, pixel_snail.py:41-56, nn/convolution.py:41-43).
// conv(x)[p] = sum_t W_t . x[p + (dy_t, dx_t)] with zero fill outside the image (the reference's pad + crop, SURVEY
// Appendix A).  The contraction itself runs on the tcgen05 GEMM: gather builds X_cat[p, t*C + c] = act(x[p+off_t, c])
// once (bf16, 16-byte chunks), the GEMM contracts over K = T*C, and backward scatters dX_cat back with the mirrored
// offsets.  act(0) = 0 for every activation on the path (ReLU / ELU), so it commutes with the zero padding.
// ------------------------------------------------------------------------------------------------
struct TapArgs {
  int N, H, W, C, T;
  int dy[32], dx[32];
};

__global__ void __launch_bounds__(256)
tap_gather_kernel(const bf16* __restrict__ x, int64_t ld_x, const TapArgs a, int act, bf16* __restrict__ out) {
  const int c8n = a.C / 8;
  const long long P = (long long)a.N * a.H * a.W;
  const long long total = P * a.T * c8n;
  const int HW = a.H * a.W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % c8n);
    const int t = (int)((idx / c8n) % a.T);
    const long long p = idx / ((long long)c8n * a.T);
    const int n = (int)(p / HW), rem = (int)(p % HW);
    const int ys = rem / a.W + a.dy[t], xs = rem % a.W + a.dx[t];
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ys >= 0 && ys < a.H && xs >= 0 && xs < a.W) {
      v = *reinterpret_cast<const uint4*>(x + ((size_t)n * HW + (size_t)ys * a.W + xs) * ld_x + c8 * 8);
      if (act != PG_ACT_NONE) {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack_bf16x2(w[i]);
          w[i] = pack_bf16x2(pg_act_fwd(act, f.x), pg_act_fwd(act, f.y));
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    *reinterpret_cast<uint4*>(out + (size_t)p * a.T * a.C + (size_t)t * a.C + c8 * 8) = v;
  }
}

// dx[p, c] = act'(x_pre[p, c]) * sum_t dxcat[p - off_t, t*C + c]
__global__ void __launch_bounds__(256)
tap_scatter_kernel(const bf16* __restrict__ dxcat, const TapArgs a, int act, const bf16* __restrict__ x_pre,
                   int64_t ld_pre, float* __restrict__ dx_f32, bf16* __restrict__ dx_bf16, int64_t ld_dx) {
  const int c8n = a.C / 8;
  const long long P = (long long)a.N * a.H * a.W;
  const long long total = P * c8n;
  const int HW = a.H * a.W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % c8n);
    const long long p = idx / c8n;
    const int n = (int)(p / HW), rem = (int)(p % HW);
    const int y = rem / a.W, xx = rem % a.W;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int t = 0; t < a.T; ++t) {
      const int ys = y - a.dy[t], xs = xx - a.dx[t];
      if (ys < 0 || ys >= a.H || xs < 0 || xs >= a.W) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(dxcat + ((size_t)n * HW + (size_t)ys * a.W + xs) * a.T * a.C +
                                                      (size_t)t * a.C + c8 * 8);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16x2(w[i]);
        acc[2 * i] += f.x;
        acc[2 * i + 1] += f.y;
      }
    }
    if (act != PG_ACT_NONE) {
      const uint4 v = *reinterpret_cast<const uint4*>(x_pre + (size_t)p * ld_pre + c8 * 8);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16x2(w[i]);
        acc[2 * i] *= pg_act_bwd(act, f.x);
        acc[2 * i + 1] *= pg_act_bwd(act, f.y);
      }
    }
    if (dx_f32) {
      float* o = dx_f32 + (size_t)p * ld_dx + c8 * 8;
      *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (dx_bf16)
      *reinterpret_cast<uint4*>(dx_bf16 + (size_t)p * ld_dx + c8 * 8) =
          make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                     pack_bf16x2(acc[6], acc[7]));
  }
}

int fill_taps(TapArgs& a, int N, int H, int W, int C, int T, const int* dy, const int* dx, const char* who) {
  PG_REQUIRE(T >= 1 && T <= 32, "%s: %d taps (max 32)", who, T);
  PG_REQUIRE(C % 8 == 0, "%s: channel count %d must be a multiple of 8", who, C);
  a.N = N; a.H = H; a.W = W; a.C = C; a.T = T;
  for (int t = 0; t < T; ++t) { a.dy[t] = dy[t]; a.dx[t] = dx[t]; }
  return 0;
}

}  // namespace

extern "C" int pg_conv_small_fwd(const float* x_nchw, const float* w_oihw, const float* bias, int N, int Cin, int H,
                                 int W, int Cout, int kh, int kw, int pad_h, int pad_w, int pre_act, float* out_f32,
                                 void* out_bf16, int act_bf16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PG_REQUIRE(x_nchw && w_oihw && (out_f32 || out_bf16), "pg_conv_small_fwd: null argument");
  ConvArgs a = {N, Cin, H, W, Cout, kh, kw, pad_h, pad_w, Cin * kh * kw, pre_act};
  PG_REQUIRE(a.K <= MAX_K, "pg_conv_small_fwd: Cin*kh*kw = %d exceeds %d", a.K, MAX_K);
  const long long P = (long long)N * H * W;
  const unsigned blocks = (unsigned)((P + PIX_PER_BLOCK - 1) / PIX_PER_BLOCK);
  conv_small_fwd_kernel<<<blocks, 256, 0, stream>>>(x_nchw, w_oihw, bias, a, out_f32, (bf16*)out_bf16, act_bf16);
  return pg_check_launch("pg_conv_small_fwd");
}

extern "C" int pg_conv_small_bwd(const float* x_nchw, const float* w_oihw, const float* dy_pm, int N, int Cin, int H,
                                 int W, int Cout, int kh, int kw, int pad_h, int pad_w, int pre_act, float* dw_oihw,
                                 float* dbias, float* dx_nchw, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PG_REQUIRE(x_nchw && w_oihw && dy_pm, "pg_conv_small_bwd: null argument");
  ConvArgs a = {N, Cin, H, W, Cout, kh, kw, pad_h, pad_w, Cin * kh * kw, pre_act};
  PG_REQUIRE(a.K <= MAX_K, "pg_conv_small_bwd: Cin*kh*kw = %d exceeds %d", a.K, MAX_K);
  const long long P = (long long)N * H * W;
  if (dw_oihw) {
    const size_t smem = (size_t)WG_PIX * (a.K + a.Cout) * sizeof(float);
    PG_REQUIRE(smem <= 200 * 1024, "pg_conv_small_bwd: shared memory %zu too large", smem);
    PG_CUDA(cudaFuncSetAttribute(conv_small_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long blocks = (P + WG_PIX - 1) / WG_PIX;
    const long long cap = (long long)pg_num_sms() * 2;
    if (blocks > cap) blocks = cap;
    // each launch covers WG_MAX_OUT_PER_THREAD * 256 = 16384 of the Cout*K outputs (one launch at every BASELINE config)
    for (int o_base = 0; o_base < a.Cout * a.K; o_base += WG_MAX_OUT_PER_THREAD * 256) {
      conv_small_wgrad_kernel<<<(unsigned)blocks, 256, smem, stream>>>(x_nchw, dy_pm, a, dw_oihw, o_base);
      if (pg_check_launch("pg_conv_small_bwd(wgrad)")) return 1;
    }
  }
  if (dbias) {
    if (pg_colsum_f32(dy_pm, Cout, (int)P, Cout, dbias, 1, stream_)) return 1;
  }
  if (dx_nchw) {
    const size_t smem_w = (size_t)a.K * Cout * sizeof(float);
    PG_REQUIRE(smem_w <= 200 * 1024, "pg_conv_small_bwd: weight tile %zu B too large for shared memory", smem_w);
    PG_CUDA(cudaFuncSetAttribute(conv_small_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long blocks = (P * 32 + 255) / 256;
    const long long cap = (long long)pg_num_sms() * 4;
    if (blocks > cap) blocks = cap;
    conv_small_dgrad_kernel<<<(unsigned)blocks, 256, smem_w, stream>>>(w_oihw, dy_pm, a, x_nchw, dx_nchw);
    if (pg_check_launch("pg_conv_small_bwd(dgrad)")) return 1;
  }
  return 0;
}

extern "C" int pg_tap_gather(const void* x_pm, int64_t ld_x, int N, int H, int W, int C, int T, const int* dy,
                             const int* dx, int act, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PG_REQUIRE(x_pm && out && dy && dx, "pg_tap_gather: null argument");
  PG_REQUIRE(ld_x % 8 == 0, "pg_tap_gather: pitch must be a multiple of 8");
  TapArgs a;
  if (fill_taps(a, N, H, W, C, T, dy, dx, "pg_tap_gather")) return 1;
  const long long total = (long long)N * H * W * T * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)pg_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  tap_gather_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const bf16*)x_pm, ld_x, a, act, (bf16*)out);
  return pg_check_launch("pg_tap_gather");
}

extern "C" int pg_tap_scatter(const void* dxcat, int N, int H, int W, int C, int T, const int* dy, const int* dx,
                              int act, const void* x_pre, int64_t ld_pre, float* dx_f32, void* dx_bf16, int64_t ld_dx,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PG_REQUIRE(dxcat && dy && dx && (dx_f32 || dx_bf16), "pg_tap_scatter: null argument");
  PG_REQUIRE(act == PG_ACT_NONE || x_pre, "pg_tap_scatter: activation backward needs the pre-activation input");
  PG_REQUIRE(ld_dx % 8 == 0 && (act == PG_ACT_NONE || ld_pre % 8 == 0), "pg_tap_scatter: pitches must be multiples of 8");
  TapArgs a;
  if (fill_taps(a, N, H, W, C, T, dy, dx, "pg_tap_scatter")) return 1;
  const long long total = (long long)N * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)pg_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  tap_scatter_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const bf16*)dxcat, a, act, (const bf16*)x_pre, ld_pre, dx_f32,
                                                            (bf16*)dx_bf16, ld_dx);
  return pg_check_launch("pg_tap_scatter");
}
