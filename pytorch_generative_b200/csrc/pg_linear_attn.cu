// pg_linear_attn.cu — the numerator of LinearCausalAttention (reference nn/attention.py:168-200,
// `_UnnormalizedLinearCausalAttention`): out_i = Q_i . S_i with the running state S_i = sum_{j <= i} K_j^T V_j.
//
// The reference walks the sequence in a Python loop (three small matmuls per position, forward and backward); here one
// CTA owns one (image, head) and keeps the d x dv state in registers: a thread owns one COLUMN of S when the product it
// serves contracts over the key / query features (forward out_i[c] = sum_a q_i[a] S[a, c]; backward dV_i[c]), or one ROW
// when it contracts over the value features (dQ_i[a] = sum_c G_i[c] S[a, c]; dK_i[a]) -- so no position needs a
// cross-thread reduction, and the rows of Q / K / V / G stream through shared memory in blocks of 32 positions.
// O(L (d + dv)) memory like the reference, fp32 throughout.  d <= 64, dv <= 128.
#include "../../include/pg_b200.h"
#include "pg_common.cuh"

namespace {

constexpr int LA_BLOCK = 32;   // positions staged per step
constexpr int LA_MAX_D = 64;   // state rows held by a column owner
constexpr int LA_MAX_DV = 128; // state columns held by a row owner

// Column owners (threads = dv): out[i, c] = sum_a X[i, a] * S[a, c],  S[a, c] += Y[i, a] * Z[i, c]
// forward:  X = Q, Y = K, Z = V, ascending i  (state updated BEFORE the product: j <= i)
// dV:       X = K, Y = Q, Z = G, descending i
template <bool REVERSE>
__global__ void __launch_bounds__(LA_MAX_DV)
la_col_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Z, float* __restrict__ out,
              int L, int d, int dv) {
  __shared__ float sx[LA_BLOCK][LA_MAX_D], sy[LA_BLOCK][LA_MAX_D], sz[LA_BLOCK][LA_MAX_DV];
  const size_t base_d = (size_t)blockIdx.x * L * d, base_v = (size_t)blockIdx.x * L * dv;
  const int c = threadIdx.x;
  float S[LA_MAX_D];
#pragma unroll
  for (int a = 0; a < LA_MAX_D; ++a) S[a] = 0.f;
  const int nblk = (L + LA_BLOCK - 1) / LA_BLOCK;
  for (int b = 0; b < nblk; ++b) {
    const int blk = REVERSE ? nblk - 1 - b : b;
    const int i0 = blk * LA_BLOCK, cnt = min(LA_BLOCK, L - i0);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * d; t += blockDim.x) {
      sx[t / d][t % d] = X[base_d + (size_t)i0 * d + t];
      sy[t / d][t % d] = Y[base_d + (size_t)i0 * d + t];
    }
    for (int t = threadIdx.x; t < cnt * dv; t += blockDim.x) sz[t / dv][t % dv] = Z[base_v + (size_t)i0 * dv + t];
    __syncthreads();
    if (c < dv) {
      for (int s = 0; s < cnt; ++s) {
        const int ii = REVERSE ? cnt - 1 - s : s;
        const float z = sz[ii][c];
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < LA_MAX_D; ++a) {
          if (a < d) {
            S[a] = fmaf(sy[ii][a], z, S[a]);
            acc = fmaf(sx[ii][a], S[a], acc);
          }
        }
        out[base_v + (size_t)(i0 + ii) * dv + c] = acc;
      }
    }
  }
}

// Row owners (threads = d): out[i, a] = sum_c X[i, c] * S[a, c],  S[a, c] += Y[i, a] * Z[i, c]
// dQ:  X = G, Y = K, Z = V, ascending;   dK:  X = V, Y = Q, Z = G, descending
template <bool REVERSE>
__global__ void __launch_bounds__(LA_MAX_D)
la_row_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Z, float* __restrict__ out,
              int L, int d, int dv) {
  __shared__ float sx[LA_BLOCK][LA_MAX_DV], sy[LA_BLOCK][LA_MAX_D], sz[LA_BLOCK][LA_MAX_DV];
  const size_t base_d = (size_t)blockIdx.x * L * d, base_v = (size_t)blockIdx.x * L * dv;
  const int a = threadIdx.x;
  float S[LA_MAX_DV];
#pragma unroll
  for (int c = 0; c < LA_MAX_DV; ++c) S[c] = 0.f;
  const int nblk = (L + LA_BLOCK - 1) / LA_BLOCK;
  for (int b = 0; b < nblk; ++b) {
    const int blk = REVERSE ? nblk - 1 - b : b;
    const int i0 = blk * LA_BLOCK, cnt = min(LA_BLOCK, L - i0);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * dv; t += blockDim.x) {
      sx[t / dv][t % dv] = X[base_v + (size_t)i0 * dv + t];
      sz[t / dv][t % dv] = Z[base_v + (size_t)i0 * dv + t];
    }
    for (int t = threadIdx.x; t < cnt * d; t += blockDim.x) sy[t / d][t % d] = Y[base_d + (size_t)i0 * d + t];
    __syncthreads();
    if (a < d) {
      for (int s = 0; s < cnt; ++s) {
        const int ii = REVERSE ? cnt - 1 - s : s;
        const float y = sy[ii][a];
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < LA_MAX_DV; ++c) {
          if (c < dv) {
            S[c] = fmaf(y, sz[ii][c], S[c]);
            acc = fmaf(sx[ii][c], S[c], acc);
          }
        }
        out[base_d + (size_t)(i0 + ii) * d + a] = acc;
      }
    }
  }
}

int la_check(int B, int L, int d, int dv, const char* who) {
  PG_REQUIRE(B > 0 && L > 0 && d > 0 && dv > 0, "%s: empty problem", who);
  PG_REQUIRE(d <= LA_MAX_D && dv <= LA_MAX_DV, "%s: head sizes d=%d (<= %d), dv=%d (<= %d)", who, d, LA_MAX_D, dv, LA_MAX_DV);
  return 0;
}

}  // namespace

extern "C" int pg_linear_attn_fwd(const float* q, const float* k, const float* v, float* out, int B, int L, int d, int dv,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PG_REQUIRE(q && k && v && out, "pg_linear_attn_fwd: null argument");
  if (la_check(B, L, d, dv, "pg_linear_attn_fwd")) return 1;
  la_col_kernel<false><<<B, LA_MAX_DV, 0, stream>>>(q, k, v, out, L, d, dv);
  return pg_check_launch("pg_linear_attn_fwd");
}

extern "C" int pg_linear_attn_bwd(const float* q, const float* k, const float* v, const float* g, float* dq, float* dk,
                                  float* dv_out, int B, int L, int d, int dv, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PG_REQUIRE(q && k && v && g && dq && dk && dv_out, "pg_linear_attn_bwd: null argument");
  if (la_check(B, L, d, dv, "pg_linear_attn_bwd")) return 1;
  la_row_kernel<false><<<B, LA_MAX_D, 0, stream>>>(g, k, v, dq, L, d, dv);       // dQ_i = G_i S_i^T, S_i = sum_{j<=i} K_j^T V_j
  if (pg_check_launch("pg_linear_attn_bwd(dq)")) return 1;
  la_col_kernel<true><<<B, LA_MAX_DV, 0, stream>>>(k, q, g, dv_out, L, d, dv);   // dV_i = K_i R_i,   R_i = sum_{j>=i} Q_j^T G_j
  if (pg_check_launch("pg_linear_attn_bwd(dv)")) return 1;
  la_row_kernel<true><<<B, LA_MAX_D, 0, stream>>>(v, q, g, dk, L, d, dv);        // dK_i = V_i R_i^T
  return pg_check_launch("pg_linear_attn_bwd(dk)");
}
