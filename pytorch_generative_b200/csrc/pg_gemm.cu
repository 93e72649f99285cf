// pg_gemm.cu — the channel-contraction kernel: bf16 x bf16 -> fp32 on tcgen05 tensor cores.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0 (one lane)  TMA producer: global -> 128B-swizzled shared-memory stages, mbarrier expect_tx
//   warp 1 (one lane)  MMA issuer:   tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instruction,
//                                    accumulator in TMEM, double buffered (2*BN <= 512 columns)
//   warp 2             TMEM allocator / deallocator
//   warps 4..7         epilogue: tcgen05.ld (lane == output row), fused bias / act' / residual / activation.
// Epilogue I/O is staged through shared memory so that HBM only ever sees full lines: the fp32 residual
// tiles and the bf16 pre-activation tile arrive by TMA load (prefetched two 32-column chunks ahead), results
// leave by TMA store from swizzled slabs (double buffered, so chunk c+1 is computed while chunk c drains).
// Shapes TMA cannot express (pitch not 16-byte aligned, fp32-atomic accumulation for split-K wgrad) take the
// direct register->global path.
// Operand majors are template parameters so that forward (K,K), dgrad (K,MN) and wgrad (MN,MN) all read the
// tensors where they lie: no transposed copies of activations or weights are ever materialised.
//
// Replaces: every 1x1 nn.Conv2d forward/backward on the reference path (see include/pg_b200.h).
#include <stdlib.h>
#include <string.h>

#include "../../include/pg_b200.h"
#include "pg_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle span
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int SLAB_F32 = BM * 128;  // [128 rows][32 fp32], 128B swizzle
constexpr int SLAB_BF16 = BM * 64;  // [128 rows][32 bf16], 64B swizzle
constexpr int SMEM_LIMIT = 232448;  // 227 KB

struct EpiMaps {
  CUtensorMap res0, res1, aux, out_f32, out_bf16, out_pre;
};

// Tap-loop convolution (pg_gemm_bf16_conv): the shifted operand is read straight from the pixel-major activation
// tensor through a 4-D TMA map [C, W, H, N]; out-of-image coordinates are zero-filled by the TMA unit, which is the
// reference's zero padding.  mode 1: A is shifted (forward / dgrad, K = taps x channel slabs); mode 2: B is shifted
// (wgrad, the taps are extra N blocks).
struct ConvGeom {
  int mode, H, W, C, T;
  int cslabs;  // C / 64 (mode 1): K iterations per tap
  int nbpt;    // mode 2: N blocks per tap (C / BN)
  int8_t dy[32], dx[32];
};

struct GemmParams {
  ConvGeom conv;
  int M, N, K;
  int num_m_blk, num_n_blk;
  int k_iters;      // ceil(K / BK)
  int k_per_split;  // k iterations per split
  int splits;
  int stages;       // smem pipeline depth
  int vec_ok;       // all epilogue pointers/pitches allow 16-byte vector access (direct path)
  // staged (TMA) epilogue plan
  int staged;
  int store_deriv;  // out_pre receives act'(pre) (PG_ACT_STORE_DERIV)
  int res_bf16;     // res0 / res1 are bf16 matrices (PG_ACT_RES_BF16)
  float* a_rowsum;  // MN-major A only: fp32 [M] += sum_k A(m, k) (pg_gemm_epilogue.bias_grad), nullptr = off
  int epi_depth;    // staging stages per epilogue warpgroup (1 or 2)
  int epi_stage_bytes;
  int off_res0, off_res1, off_aux, off_outf, off_outb, off_outp;  // slab offsets inside an epilogue stage, -1 = absent
  int in_bytes;  // bytes TMA-loaded per chunk (res0 + res1 + aux slabs)
  pg_gemm_epilogue epi;
};

// ------------------------------------------------------------------------------------------------
// Direct epilogue (generic fallback): `acc` = 32 consecutive fp32 accumulator columns of output row `row`.
// ------------------------------------------------------------------------------------------------
template <bool BF16_RES = true>
__device__ __forceinline__ void epilogue_row32(const GemmParams& p, int row, int col0, int ncols, bool first_split,
                                               const uint32_t (&acc)[32]) {
  const pg_gemm_epilogue& e = p.epi;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * e.alpha;

  const bool full = (ncols == 32) && p.vec_ok;
  if (first_split && e.bias) {
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 b = __ldg(b4 + i);
        v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
      }
    } else {
      for (int i = 0; i < ncols; ++i) v[i] += __ldg(e.bias + col0 + i);
    }
  }
  if (e.dact != PG_ACT_NONE) {
    const bf16* aux = reinterpret_cast<const bf16*>(e.aux) + (size_t)row * e.ld_aux + col0;
    if (full) {
      const uint4* a4 = reinterpret_cast<const uint4*>(aux);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 a = __ldg(a4 + i);
        uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = unpack_bf16x2(w[j]);
          v[8 * i + 2 * j] *= pg_act_bwd(e.dact, f.x);
          v[8 * i + 2 * j + 1] *= pg_act_bwd(e.dact, f.y);
        }
      }
    } else {
      for (int i = 0; i < ncols; ++i) v[i] *= pg_act_bwd(e.dact, __bfloat162float(aux[i]));
    }
  }
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const float* rp = which == 0 ? e.res0 : e.res1;
    if (BF16_RES && first_split && rp && p.res_bf16) {
      const bf16* r = reinterpret_cast<const bf16*>(rp) + (size_t)row * e.ld_res + col0;
      for (int i = 0; i < ncols; ++i) v[i] += __bfloat162float(r[i]);
    } else if (first_split && rp) {
      const float* r = rp + (size_t)row * e.ld_res + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 b = __ldg(reinterpret_cast<const float4*>(r) + i);
          v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
      } else {
        for (int i = 0; i < ncols; ++i) v[i] += r[i];
      }
    }
  }
  if (e.out_f32) {
    float* o = e.out_f32 + (size_t)row * e.ld_out_f32 + col0;
    if (e.accumulate) {
      for (int i = 0; i < ncols; ++i) atomicAdd(o + i, v[i]);
    } else if (full) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
      for (int i = 0; i < ncols; ++i) o[i] = v[i];
    }
  }
  if (e.out_pre) {
    bf16* o = reinterpret_cast<bf16*>(e.out_pre) + (size_t)row * e.ld_out_pre + col0;
    float d[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = p.store_deriv ? pg_act_bwd(e.act, v[i]) : v[i];
    if (full) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint4*>(o)[i] =
            make_uint4(pack_bf16x2(d[8 * i], d[8 * i + 1]), pack_bf16x2(d[8 * i + 2], d[8 * i + 3]),
                       pack_bf16x2(d[8 * i + 4], d[8 * i + 5]), pack_bf16x2(d[8 * i + 6], d[8 * i + 7]));
    } else {
      for (int i = 0; i < ncols; ++i) o[i] = __float2bfloat16(d[i]);
    }
  }
  if (e.out_bf16) {
    bf16* o = reinterpret_cast<bf16*>(e.out_bf16) + (size_t)row * e.ld_out_bf16 + col0;
    if (e.act != PG_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = pg_act_fwd(e.act, v[i]);
    }
    if (full) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint4*>(o)[i] =
            make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                       pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
    } else {
      for (int i = 0; i < ncols; ++i) o[i] = __float2bfloat16(v[i]);
    }
  }
}

// ---- swizzled slab row access (thread == tile row r) ----
// fp32 slab: [128][128B], TMA SWIZZLE_128B: 16-byte unit u of row r lives at r*128 + ((u ^ (r & 7)) << 4).
__device__ __forceinline__ void slab_f32_add(const uint8_t* slab, int r, float (&v)[32]) {
  const uint8_t* row = slab + r * 128;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float4 x = *reinterpret_cast<const float4*>(row + ((u ^ (r & 7)) << 4));
    v[4 * u] += x.x; v[4 * u + 1] += x.y; v[4 * u + 2] += x.z; v[4 * u + 3] += x.w;
  }
}
__device__ __forceinline__ void slab_f32_store(uint8_t* slab, int r, const float (&v)[32]) {
  uint8_t* row = slab + r * 128;
#pragma unroll
  for (int u = 0; u < 8; ++u)
    *reinterpret_cast<float4*>(row + ((u ^ (r & 7)) << 4)) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
}
// bf16 slab: [128][64B], TMA SWIZZLE_64B: unit u (of 4) of row r lives at r*64 + ((u ^ ((r >> 1) & 3)) << 4).
__device__ __forceinline__ void slab_bf16_load(const uint8_t* slab, int r, float (&x)[32]) {
  const uint8_t* row = slab + r * 64;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 w = *reinterpret_cast<const uint4*>(row + ((u ^ ((r >> 1) & 3)) << 4));
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(ww[j]);
      x[8 * u + 2 * j] = f.x;
      x[8 * u + 2 * j + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void slab_bf16_add(const uint8_t* slab, int r, float (&v)[32]) {
  const uint8_t* row = slab + r * 64;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 w = *reinterpret_cast<const uint4*>(row + ((u ^ ((r >> 1) & 3)) << 4));
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(ww[j]);
      v[8 * u + 2 * j] += f.x;
      v[8 * u + 2 * j + 1] += f.y;
    }
  }
}
__device__ __forceinline__ void slab_bf16_store(uint8_t* slab, int r, const float (&v)[32]) {
  uint8_t* row = slab + r * 64;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    *reinterpret_cast<uint4*>(row + ((u ^ ((r >> 1) & 3)) << 4)) =
        make_uint4(pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                   pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
}

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
// TWO = cta_group::2: the two CTAs of a cluster pair compute one 256 x BN tile; each stages its own 128 rows of A and
// its half (BN/2 rows) of B, the leader issues M = 256 MMAs that read both CTAs' shared memory, each CTA owns the
// accumulator rows of its half in its own TMEM and runs its own epilogue.  Halves the L2 -> SM operand traffic per flop
// (64 instead of 96 B/cycle/SM at BN = 256), which is what bounds the 1-CTA kernel.
// Epilogue warpgroups per CTA (each 4 warps = the four TMEM lane quarters).  More groups = more warps to hide the
// TMEM-load / shared-memory / barrier latencies of the epilogue; the register budget per thread shrinks accordingly.
// Two groups (384 threads, 168 registers) is the default; the short-K pair kernels, whose tile time is mostly
// epilogue, are also built with three (512 threads, 126 registers) -- see dispatch_bn.
template <int BN, bool A_MN, bool B_MN, bool TWO, int EPI_GROUPS>
__global__ void __launch_bounds__(128 + 128 * EPI_GROUPS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ EpiMaps em, const GemmParams p) {
  static_assert(!TWO || (!A_MN && BN == 256), "the 2-CTA kernel is the K-major-A, 256-wide variant");
  constexpr int B_ROWS = TWO ? BN / 2 : BN;  // rows of B (N extent) staged by this CTA
  constexpr int B_STAGE_BYTES = B_ROWS * BK * 2;
  constexpr int MT = TWO ? 2 * BM : BM;      // M extent of a tile
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int ACC_STAGES = 2;
  constexpr int TMEM_COLS = (ACC_STAGES * BN <= 32) ? 32
                            : (ACC_STAGES * BN <= 64) ? 64
                            : (ACC_STAGES * BN <= 128) ? 128
                            : (ACC_STAGES * BN <= 256) ? 256 : 512;
  constexpr int MAX_STAGES = 8;

  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned stage bases.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int STAGES = p.stages;
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + EPI_GROUPS * p.epi_depth * p.epi_stage_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint64_t* in_full = tempty_bar + ACC_STAGES;  // [groups][2 stages] epilogue input slabs landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_full + 2 * EPI_GROUPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;          // 0 = leader of the CTA pair
  const int tile_first = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      // freed by the MMAs' commit and, when the side reduction over A is on, by its two warps as well
      mbar_init(&empty_bar[i], (A_MN && p.a_rowsum != nullptr) ? 3 : 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], (TWO ? 8 : 4) * EPI_GROUPS);  // one arrive per epilogue warp, of both CTAs if paired
    }
    for (int i = 0; i < 2 * EPI_GROUPS; ++i) mbar_init(&in_full[i], 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    if (TWO) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();  // both CTAs' barriers are initialised before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_blk * p.num_n_blk * p.splits;

  if (warp == 0) {
    {
      // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
      int s = 0;
      uint32_t ph = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
        const int n_blk = tile % p.num_n_blk;
        const int rest = tile / p.num_n_blk;
        const int m_blk = rest % p.num_m_blk;
        const int ks = rest / p.num_m_blk;
        const int k0 = ks * p.k_per_split;
        const int k1 = min(k0 + p.k_per_split, p.k_iters);
        for (int kit = k0; kit < k1; ++kit) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sA = smem + s * STAGE_BYTES;
          uint8_t* sB = sA + A_STAGE_BYTES;
          if (TWO) {
            // both CTAs load their halves; all bytes are accounted on the leader's barrier
            if (rank == 0) mbar_arrive_expect_tx_w(&full_bar[s], 2 * STAGE_BYTES);
            const int row0 = m_blk * MT + rank * BM, nb0 = n_blk * BN + rank * B_ROWS;
            int bcol = nb0, brow = kit * BK;  // MN-major B coordinates
            if (p.conv.mode == 1) {
              const int t = kit / p.conv.cslabs, cs = kit - t * p.conv.cslabs;
              const int hw = p.conv.H * p.conv.W, n = row0 / hw, h0 = (row0 - n * hw) / p.conv.W;
              tma_load_4d_2sm_w(sA, &tmA, &full_bar[s], cs * 64, p.conv.dx[t], h0 + p.conv.dy[t], n);
              bcol = t * p.N + nb0;  // dgrad: W^T of tap t starts at column t * Cin of the packed weight
              brow = cs * 64;
            } else {
              tma_load_2d_2sm_w(sA, &tmA, &full_bar[s], kit * BK, row0);
            }
            if (!B_MN) {
              tma_load_2d_2sm_w(sB, &tmB, &full_bar[s], kit * BK, nb0);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j)
                tma_load_2d_2sm_w(sB + j * (BK * 128), &tmB, &full_bar[s], bcol + j * 64, brow);
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx_w(&full_bar[s], STAGE_BYTES);
          if (p.conv.mode != 0) {
            const int hw = p.conv.H * p.conv.W;
            if (p.conv.mode == 1) {  // A = activations under tap t (forward / dgrad)
              const int t = kit / p.conv.cslabs, cs = kit - t * p.conv.cslabs;
              const int row0 = m_blk * BM, n = row0 / hw, h0 = (row0 - n * hw) / p.conv.W;
              tma_load_4d_w(sA, &tmA, &full_bar[s], cs * 64, p.conv.dx[t], h0 + p.conv.dy[t], n);
              if (!B_MN) {
                tma_load_2d_w(sB, &tmB, &full_bar[s], kit * BK, n_blk * BN);
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_2d_w(sB + j * (BK * 128), &tmB, &full_bar[s], t * p.N + n_blk * BN + j * 64, cs * 64);
              }
            } else {  // wgrad: A = dY (MN-major), B = activations under the tap this N block belongs to
              const int t = n_blk / p.conv.nbpt, nb = n_blk - t * p.conv.nbpt;
              const int pix0 = kit * BK, n = pix0 / hw, h0 = (pix0 - n * hw) / p.conv.W;
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d_w(sA + j * (BK * 128), &tmA, &full_bar[s], m_blk * BM + j * 64, kit * BK);
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_4d_w(sB + j * (BK * 128), &tmB, &full_bar[s], nb * BN + j * 64, p.conv.dx[t], h0 + p.conv.dy[t], n);
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
            continue;
          }
          if (!A_MN) {
            tma_load_2d_w(sA, &tmA, &full_bar[s], kit * BK, m_blk * BM);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)  // box {64 m, 64 k-rows} per 64-wide MN atom
              tma_load_2d_w(sA + j * (BK * 128), &tmA, &full_bar[s], m_blk * BM + j * 64, kit * BK);
          }
          if (!B_MN) {
            tma_load_2d_w(sB, &tmB, &full_bar[s], kit * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d_w(sB + j * (BK * 128), &tmB, &full_bar[s], n_blk * BN + j * 64, kit * BK);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA only when paired; whole warp converged, elected lane issues) =====================
      constexpr uint32_t idesc = umma_idesc_bf16(MT, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
        const int rest = tile / p.num_n_blk;
        const int ks = rest / p.num_m_blk;
        const int k0 = ks * p.k_per_split;
        const int k1 = min(k0 + p.k_per_split, p.k_iters);
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kit = k0; kit < k1; ++kit) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
          // one descriptor per operand stage; a K step (16 elements) only moves the start-address field:
          // K-major: 32 bytes inside the swizzle span; MN-major: 16 k-rows of 128 B = 2048 bytes (LBO = atom stride)
          const uint64_t a_base = A_MN ? umma_desc_sw128(a_addr, BK * 128, 1024) : umma_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_base = B_MN ? umma_desc_sw128(b_addr, BK * 128, 1024) : umma_desc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t a_desc = a_base + (uint64_t)(A_MN ? kk * 128 : kk * 2);
            const uint64_t b_desc = b_base + (uint64_t)(B_MN ? kk * 128 : kk * 2);
            if (TWO) umma_bf16_ss_2sm_w(d_tmem, a_desc, b_desc, idesc, (kit > k0 || kk > 0) ? 1u : 0u);
            else umma_bf16_ss_w(d_tmem, a_desc, b_desc, idesc, (kit > k0 || kk > 0) ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs when paired) once these MMAs have read it
          if (TWO) umma_commit_2sm_w(&empty_bar[s]);
          else umma_commit_w(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        // accumulator complete -> epilogue(s)
        if (TWO) umma_commit_2sm_w(&tfull_bar[as]);
        else umma_commit_w(&tfull_bar[as]);
        if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
      }
    }
  } else if (A_MN && warp < 4) {
    // ===================== warps 2-3, weight-gradient GEMMs: bias gradient from the staged A tiles =====================
    // A = dY read MN-major, so sum_k A(m, k) is the bias gradient of the layer whose weight gradient this launch
    // computes.  The two otherwise idle warps add up every A stage while the MMAs run (16 KB of shared-memory reads per
    // 48 KB operand stage) — the separate column-sum pass re-read dY from HBM.  Only the CTAs of N block 0 do it (every
    // (m block, split) is seen once); partial sums go to a_rowsum with fp32 atomics, like the split-K tiles.
    if (p.a_rowsum != nullptr) {
      const int t = (warp - 2) * 32 + lane;   // 0..63
      const int c = t & 15, g = t >> 4;        // 16-byte chunk (8 consecutive m) of the 128-wide tile, k-row group
      const uint32_t atom = (uint32_t)(c >> 3) * (BK * 128), cc = (uint32_t)(c & 7);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
        const int n_blk = tile % p.num_n_blk;
        const int rest = tile / p.num_n_blk;
        const int m_blk = rest % p.num_m_blk;
        const int ks = rest / p.num_m_blk;
        const int k0 = ks * p.k_per_split;
        const int k1 = min(k0 + p.k_per_split, p.k_iters);
        const bool mine = (n_blk == 0);
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        for (int kit = k0; kit < k1; ++kit) {
          mbar_wait(&full_bar[s], ph);
          if (mine) {
            const uint8_t* sA = smem + s * STAGE_BYTES + atom;
#pragma unroll
            for (int i = 0; i < BK / 4; ++i) {
              const uint32_t k = (uint32_t)(g + 4 * i);
              const uint4 w = *reinterpret_cast<const uint4*>(sA + k * 128 + ((cc ^ (k & 7u)) << 4));
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(ww[j]);
                acc[2 * j] += f.x;
                acc[2 * j + 1] += f.y;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[s]);   // release: this warp's reads of the stage are done
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (mine) {
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 16);  // the warp's two k-row groups
          if (lane < 16) {
            const int m0 = m_blk * BM + c * 8;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (m0 + q < p.M) atomicAdd(p.a_rowsum + m0 + q, acc[q]);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: EPI_GROUPS warpgroups, round-robin over 32-column chunks =====================
    // Warp w may only touch TMEM lanes [32*(w%4), +32), so warps 4..7 (group 0) and 8..11 (group 1) cover the
    // same 128 rows; group g owns every chunk whose position in this CTA's chunk sequence is g mod EPI_GROUPS, its own
    // staging stage, input barrier and named barrier.  Two groups double the issue slots of the epilogue math.
    const int grp = (warp - 4) >> 2;
    const int ew = warp & 3;        // TMEM lane quarter
    const int r = ew * 32 + lane;   // row inside the tile
    const bool leader = (r == 0);
    const pg_gemm_epilogue& e = p.epi;
    constexpr int NCH = BN / 32;
    int as = 0;
    uint32_t aph = 0;
    unsigned gc = 0;       // position of the next chunk in this CTA's chunk sequence (all tiles)
    unsigned used = 0;     // chunks this group has processed (parity of in_full[grp])
    const int depth = p.epi_depth;  // staging stages owned by this group
    uint8_t* const grp_base = epi_smem + grp * depth * p.epi_stage_bytes;
    const uint32_t bar_id = 1 + grp;

    auto issue_inputs = [&](int tile, int c, int st) {  // group leader only
      const int n_blk = tile % p.num_n_blk;
      const int m_blk = (tile / p.num_n_blk) % p.num_m_blk;
      const int col0 = n_blk * BN + c * 32, row0 = m_blk * MT + (int)rank * BM;
      uint8_t* b = grp_base + st * p.epi_stage_bytes;
      uint64_t* bar = &in_full[grp * 2 + st];
      mbar_arrive_expect_tx(bar, p.in_bytes);
      if (p.off_res0 >= 0) tma_load_2d(b + p.off_res0, &em.res0, bar, col0, row0);
      if (p.off_res1 >= 0) tma_load_2d(b + p.off_res1, &em.res1, bar, col0, row0);
      if (p.off_aux >= 0) tma_load_2d(b + p.off_aux, &em.aux, bar, col0, row0);
    };
    // chunk at sequence position `pos` -> (tile, chunk)
    auto prefetch_pos = [&](unsigned pos, int st) {
      const int t = tile_first + (int)(pos / NCH) * tile_stride;
      if (t < num_tiles) issue_inputs(t, (int)(pos % NCH), st);
    };
    if (p.staged && p.in_bytes > 0 && leader) {
      prefetch_pos(grp, 0);
      if (depth > 1) prefetch_pos(grp + EPI_GROUPS, 1);
    }
    __syncwarp();

    for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
      const int n_blk = tile % p.num_n_blk;
      const int rest = tile / p.num_n_blk;
      const int m_blk = rest % p.num_m_blk;
      const int ks = rest / p.num_m_blk;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row = m_blk * MT + (int)rank * BM + r;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c, ++gc) {
        if ((gc % (unsigned)EPI_GROUPS) != (unsigned)grp) continue;  // another group's chunk
        const int col0 = n_blk * BN + c * 32;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN + c * 32, acc);
        tmem_wait_ld();
        if (!p.staged) {
          if (row < p.M && col0 < p.N) epilogue_row32<(EPI_GROUPS < 3)>(p, row, col0, min(32, p.N - col0), ks == 0, acc);
          continue;
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * e.alpha;
        if (e.bias) {
          if (col0 + 32 <= p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col0) + i);
              v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) v[i] += __ldg(e.bias + col0 + i);
          }
        }
        const int st = (depth > 1) ? (int)(used & 1u) : 0;
        uint8_t* const base = grp_base + st * p.epi_stage_bytes;
        if (p.in_bytes > 0) mbar_wait(&in_full[grp * 2 + st], (depth > 1 ? (used >> 1) : used) & 1u);
        if (p.off_aux >= 0) {
          float x[32];
          slab_bf16_load(base + p.off_aux, r, x);
          if (e.dact == PG_ACT_GIVEN) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= x[i];
          } else if (e.dact == PG_ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= pg_act_bwd(PG_ACT_GELU, x[i]);
          } else if (e.dact == PG_ACT_ELU_OUT) {  // aux = elu(pre): elu' = 1 (a > 0) or a + 1
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= (x[i] > 0.f ? 1.f : x[i] + 1.f);
          } else if (e.dact == PG_ACT_RELU_OUT || e.dact == PG_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = x[i] > 0.f ? v[i] : 0.f;
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= pg_act_bwd(e.dact, x[i]);
          }
        }
        if (EPI_GROUPS < 3 && p.res_bf16) {  // (the register-tight three-group build never sees bf16 residuals: dispatch_bn)
          if (p.off_res0 >= 0) slab_bf16_add(base + p.off_res0, r, v);
          if (p.off_res1 >= 0) slab_bf16_add(base + p.off_res1, r, v);
        } else {
          if (p.off_res0 >= 0) slab_f32_add(base + p.off_res0, r, v);
          if (p.off_res1 >= 0) slab_f32_add(base + p.off_res1, r, v);
        }
        // the TMA store that last used this stage's output slabs must have finished reading them
        if (leader) {
          if (depth > 1) tma_store_wait_read<1>();
          else tma_store_wait_read<0>();
        }
        __syncwarp();  // bar.sync / tcgen05.ld are warp-aligned: reconverge after every leader-only section
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (p.off_outf >= 0) slab_f32_store(base + p.off_outf, r, v);
        if (p.store_deriv && e.act == PG_ACT_GELU && p.off_outp >= 0 && p.off_outb >= 0) {
          // MLP forward: GELU(pre) and GELU'(pre) from one tanh per element
          float d[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) pg_gelu_both(v[i], v[i], d[i]);
          slab_bf16_store(base + p.off_outp, r, d);
          slab_bf16_store(base + p.off_outb, r, v);
        } else {
          if (p.off_outp >= 0) {
            if (p.store_deriv) {
              float d[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) d[i] = pg_act_bwd(e.act, v[i]);
              slab_bf16_store(base + p.off_outp, r, d);
            } else {
              slab_bf16_store(base + p.off_outp, r, v);
            }
          }
          if (p.off_outb >= 0) {
            if (e.act == PG_ACT_GELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = pg_act_fwd(PG_ACT_GELU, v[i]);
            } else if (e.act == PG_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (e.act == PG_ACT_ELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = pg_elu_fast(v[i]);
            } else if (e.act != PG_ACT_NONE) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = pg_act_fwd(e.act, v[i]);
            }
            slab_bf16_store(base + p.off_outb, r, v);
          }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (leader) {
          const int row0 = m_blk * MT + (int)rank * BM;
          if (p.off_outf >= 0) {
            if (e.accumulate) tma_reduce_add_2d(&em.out_f32, base + p.off_outf, col0, row0);
            else tma_store_2d(&em.out_f32, base + p.off_outf, col0, row0);
          }
          if (p.off_outp >= 0) tma_store_2d(&em.out_pre, base + p.off_outp, col0, row0);
          if (p.off_outb >= 0) tma_store_2d(&em.out_bf16, base + p.off_outb, col0, row0);
          tma_store_commit();
          if (p.in_bytes > 0) prefetch_pos(gc + EPI_GROUPS * depth, st);  // refill this stage for this group's chunk `depth` ahead
        }
        ++used;
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (TWO && rank != 0) mbar_arrive_cluster(&tempty_bar[as], 0);  // the leader's MMA waits for both epilogues
        else mbar_arrive(&tempty_bar[as]);
      }
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
    }
    if (p.staged && leader) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();  // neither CTA may exit (or free TMEM) while its partner can still signal / read it
  if (warp == 2) {
    tc_fence_after();
    if (TWO) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT cross-check kernel (tests only): one thread per 1x32 output segment, same epilogue arithmetic.
// ------------------------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, int a_mn, int64_t lda, const bf16* __restrict__ B,
                                 int b_mn, int64_t ldb, const GemmParams p) {
  const int nseg = (p.N + 31) / 32;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.M * nseg) return;
  const int row = (int)(idx / nseg);
  const int col0 = (int)(idx % nseg) * 32;
  const int ncols = min(32, p.N - col0);
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int k = 0; k < p.K; ++k) {
    const float a = __bfloat162float(a_mn ? A[(size_t)k * lda + row] : A[(size_t)row * lda + k]);
    for (int i = 0; i < ncols; ++i) {
      const int n = col0 + i;
      const float b = __bfloat162float(b_mn ? B[(size_t)k * ldb + n] : B[(size_t)n * ldb + k]);
      acc[i] = fmaf(a, b, acc[i]);
    }
  }
  uint32_t accu[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) accu[i] = __float_as_uint(acc[i]);
  epilogue_row32(p, row, col0, ncols, true, accu);
}

// ------------------------------------------------------------------------------------------------
// Skinny forward GEMM (M <= 32 rows): the per-pixel step of incremental sampling multiplies a handful of rows (one per
// image) by the full weight matrices.  A 128-row tensor-core tile would run on N/256 SMs and be latency-bound; here
// every warp owns one output column, streams its weight row once (16-byte loads) against the A rows held in shared
// memory, and the whole chip participates.  Same fused epilogue semantics (bias, residuals, activation).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const bf16* __restrict__ A, int64_t lda, const bf16* __restrict__ B, int64_t ldb, const GemmParams p) {
  extern __shared__ uint4 sA4[];  // [M][K/8] 16-byte units
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K8 = p.K / 8;
  for (int i = threadIdx.x; i < p.M * K8; i += blockDim.x)
    sA4[i] = *reinterpret_cast<const uint4*>(A + (size_t)(i / K8) * lda + (i % K8) * 8);
  __syncthreads();
  const int n = blockIdx.x * 8 + warp;
  if (n >= p.N) return;
  float acc[32];
#pragma unroll
  for (int m = 0; m < 32; ++m) acc[m] = 0.f;
  const uint4* wrow = reinterpret_cast<const uint4*>(B + (size_t)n * ldb);
  for (int k8 = lane; k8 < K8; k8 += 32) {
    const uint4 wv = __ldg(wrow + k8);
    const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
    float wf[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(ww[j]);
      wf[2 * j] = f.x;
      wf[2 * j + 1] = f.y;
    }
#pragma unroll
    for (int m = 0; m < 32; ++m) {
      if (m < p.M) {
        const uint4 av = sA4[m * K8 + k8];
        const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(aw[j]);
          acc[m] = fmaf(f.x, wf[2 * j], fmaf(f.y, wf[2 * j + 1], acc[m]));
        }
      }
    }
  }
  float mine = 0.f;
#pragma unroll
  for (int m = 0; m < 32; ++m) {
    if (m < p.M) {
      const float v = warp_sum(acc[m]);
      if (lane == m) mine = v;
    }
  }
  if (lane < p.M) {
    const pg_gemm_epilogue& e = p.epi;
    const int m = lane;
    float t = mine * e.alpha;
    if (e.bias) t += e.bias[n];
    if (p.res_bf16) {
      if (e.res0) t += __bfloat162float(reinterpret_cast<const bf16*>(e.res0)[(size_t)m * e.ld_res + n]);
      if (e.res1) t += __bfloat162float(reinterpret_cast<const bf16*>(e.res1)[(size_t)m * e.ld_res + n]);
    } else {
      if (e.res0) t += e.res0[(size_t)m * e.ld_res + n];
      if (e.res1) t += e.res1[(size_t)m * e.ld_res + n];
    }
    if (e.out_f32) e.out_f32[(size_t)m * e.ld_out_f32 + n] = t;
    if (e.out_pre)
      reinterpret_cast<bf16*>(e.out_pre)[(size_t)m * e.ld_out_pre + n] =
          __float2bfloat16(p.store_deriv ? pg_act_bwd(e.act, t) : t);
    if (e.out_bf16) reinterpret_cast<bf16*>(e.out_bf16)[(size_t)m * e.ld_out_bf16 + n] = __float2bfloat16(pg_act_fwd(e.act, t));
  }
}

template <int BN, bool A_MN, bool B_MN, bool TWO = false, int EPI_GROUPS = 2>
int launch_tc(const void* A, int64_t lda, const void* B, int64_t ldb, GemmParams& p, cudaStream_t stream) {
  constexpr int B_ROWS = TWO ? BN / 2 : BN;
  if (TWO) p.num_m_blk = (p.M + 2 * BM - 1) / (2 * BM);
  CUtensorMap tmA, tmB;
  const ConvGeom& cg = p.conv;
  auto conv_map = [&](CUtensorMap* out, const void* base, int64_t ld, int pixels_per_box) {
    const int64_t n_img = (cg.mode == 1 ? (int64_t)p.M : (int64_t)p.K) / ((int64_t)cg.H * cg.W);
    uint64_t dims[4] = {(uint64_t)cg.C, (uint64_t)cg.W, (uint64_t)cg.H, (uint64_t)n_img};
    uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)cg.W * ld * 2, (uint64_t)cg.H * cg.W * ld * 2};
    uint32_t box[4] = {64, (uint32_t)cg.W, (uint32_t)(pixels_per_box / cg.W), 1};
    return pg_make_tmap_nd_bf16(out, base, 4, dims, strides, box, 1);
  };
  if (cg.mode == 1) {
    if (conv_map(&tmA, A, lda, BM)) return 1;
  } else if (!A_MN) {
    if (pg_make_tmap_2d_bf16(&tmA, A, p.M, p.K, lda, BM, BK)) return 1;
  } else {
    if (pg_make_tmap_2d_bf16(&tmA, A, p.K, p.M, lda, BK, 64)) return 1;
  }
  if (cg.mode == 2) {
    PG_REQUIRE(B_MN && cg.C % BN == 0, "pg_gemm_bf16_conv(wgrad): channels (%d) must be a multiple of the N tile (%d)", cg.C, BN);
    p.conv.nbpt = cg.C / BN;
    if (conv_map(&tmB, B, ldb, BK)) return 1;
  } else if (!B_MN) {
    if (pg_make_tmap_2d_bf16(&tmB, B, p.N, p.K, ldb, B_ROWS, BK)) return 1;
  } else if (cg.mode == 1) {  // dgrad: the packed weight [Cout, T * Cin] read K-rows x N-columns, tap t at column t * Cin
    if (pg_make_tmap_2d_bf16(&tmB, B, cg.C, (uint64_t)cg.T * p.N, ldb, BK, 64)) return 1;
  } else {
    if (pg_make_tmap_2d_bf16(&tmB, B, p.K, p.N, ldb, BK, 64)) return 1;
  }
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_ROWS * BK * 2;
  const pg_gemm_epilogue& e = p.epi;
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  // ---- epilogue staging plan ----
  p.off_res0 = p.off_res1 = p.off_aux = p.off_outf = p.off_outb = p.off_outp = -1;
  p.in_bytes = 0;
  p.epi_stage_bytes = 0;
  const bool pure_acc = e.accumulate && !e.bias && !e.res0 && !e.res1 && e.dact == PG_ACT_NONE && !e.out_bf16 &&
                        !e.out_pre && e.alpha == 1.0f;  // split-K / grad accumulation: TMA reduce-add of the raw tile
  p.staged = (p.vec_ok && (pure_acc || (!e.accumulate && p.splits == 1))) ? 1 : 0;
  if (p.staged) {
    int off = 0;
    auto add = [&](int& slot, int bytes) { slot = off; off += bytes; };
    const int slab_res = p.res_bf16 ? SLAB_BF16 : SLAB_F32;
    if (e.res0) { add(p.off_res0, slab_res); p.in_bytes += slab_res; }
    if (e.res1) { add(p.off_res1, slab_res); p.in_bytes += slab_res; }
    if (e.dact != PG_ACT_NONE) { add(p.off_aux, SLAB_BF16); p.in_bytes += SLAB_BF16; }
    if (e.out_f32) add(p.off_outf, SLAB_F32);
    if (e.out_pre) add(p.off_outp, SLAB_BF16);
    if (e.out_bf16) add(p.off_outb, SLAB_BF16);
    p.epi_stage_bytes = off;
    const int res_es = p.res_bf16 ? 2 : 4, res_sw = p.res_bf16 ? 64 : 128;
    if (e.res0 && pg_make_tmap_2d(&em.res0, e.res0, res_es, p.M, p.N, e.ld_res, BM, 32, res_sw)) return 1;
    if (e.res1 && pg_make_tmap_2d(&em.res1, e.res1, res_es, p.M, p.N, e.ld_res, BM, 32, res_sw)) return 1;
    if (e.dact != PG_ACT_NONE && pg_make_tmap_2d(&em.aux, e.aux, 2, p.M, p.N, e.ld_aux, BM, 32, 64)) return 1;
    if (e.out_f32 && pg_make_tmap_2d(&em.out_f32, e.out_f32, 4, p.M, p.N, e.ld_out_f32, BM, 32, 128)) return 1;
    if (e.out_pre && pg_make_tmap_2d(&em.out_pre, e.out_pre, 2, p.M, p.N, e.ld_out_pre, BM, 32, 64)) return 1;
    if (e.out_bf16 && pg_make_tmap_2d(&em.out_bf16, e.out_bf16, 2, p.M, p.N, e.ld_out_bf16, BM, 32, 64)) return 1;
  }
  // two staging stages per epilogue group when that still leaves a >= 3-deep operand pipeline
  // (short-K tiles only: there the epilogue is a large share of the tile time; long-K tiles want the smem for
  // a deeper operand pipeline instead)
  p.epi_depth = (p.staged && p.k_per_split <= 16 && (SMEM_LIMIT - 2 * EPI_GROUPS * p.epi_stage_bytes - 1536) / STAGE_BYTES >= 3) ? 2 : 1;
  const int fixed = EPI_GROUPS * p.epi_depth * p.epi_stage_bytes + 1024 /*align slack*/ + 512 /*barriers*/;
  int stages = (SMEM_LIMIT - fixed) / STAGE_BYTES;
  if (stages > 8) stages = 8;
  if (stages > p.k_iters + 1) stages = p.k_iters + 1 > 2 ? p.k_iters + 1 : 2;
  PG_REQUIRE(stages >= 2, "pg_gemm_bf16: epilogue staging leaves no room for the operand pipeline (BN=%d)", BN);
  p.stages = stages;
  const int smem_bytes = stages * STAGE_BYTES + fixed;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, TWO, EPI_GROUPS>;
  constexpr int GEMM_THREADS = 128 + 128 * EPI_GROUPS;
  PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int num_tiles = p.num_m_blk * p.num_n_blk * p.splits;
  if (!TWO) {
    const int grid = min(num_tiles, pg_num_sms());
    kern<<<grid, GEMM_THREADS, smem_bytes, stream>>>(tmA, tmB, em, p);
    return pg_check_launch("pg_gemm_bf16(tcgen05)");
  }
  // CTA pairs: clusters of 2 along x, one pair per tile stream
  int pairs = pg_num_sms() / 2;
  if (pairs > num_tiles) pairs = num_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, em, p));
  return pg_check_launch("pg_gemm_bf16(tcgen05, cta_group::2)");
}

template <bool A_MN, bool B_MN>
int dispatch_bn(const void* A, int64_t lda, const void* B, int64_t ldb, GemmParams& p, cudaStream_t stream) {
  // Tile width: 256 when N is wide enough to fill it, otherwise the smallest of {32,64,128} covering N.
  int bn;
  if (p.N > 128) bn = 256;
  else if (p.N > 64) bn = 128;
  else if (p.N > 32) bn = 64;
  else bn = 32;
  if (B_MN && bn < 64) bn = 64;  // MN-major operands are staged in 64-wide swizzle atoms
  if (p.conv.mode == 2) bn = (p.conv.C % 256 == 0) ? 256 : (p.conv.C % 128 == 0) ? 128 : 64;  // taps are whole N blocks
  const pg_gemm_epilogue& e = p.epi;
  const int slab_res = p.res_bf16 ? SLAB_BF16 : SLAB_F32;
  const int epi = (e.res0 ? slab_res : 0) + (e.res1 ? slab_res : 0) + (e.dact != PG_ACT_NONE ? SLAB_BF16 : 0) +
                  (e.out_f32 ? SLAB_F32 : 0) + (e.out_pre ? SLAB_BF16 : 0) + (e.out_bf16 ? SLAB_BF16 : 0);
  const bool pure_acc = e.accumulate && !e.bias && !e.res0 && !e.res1;
  const bool staged = p.vec_ok && (!e.accumulate || pure_acc);
  if constexpr (!A_MN) {
    // CTA pairs (cta_group::2, 256 x 256 tiles) for the big pixel-major GEMMs (forward / dgrad): enough tiles to fill
    // 74 pairs, and room for >= 3 operand stages of 32 KB next to the epilogue slabs.
    static const bool no_pairs = getenv("PG_GEMM_NO_PAIRS") != nullptr;
    const int tiles = ((p.M + 255) / 256) * ((p.N + 255) / 256);
    constexpr int PAIR_STAGE = A_STAGE_BYTES + 128 * BK * 2;
    const bool fits = !staged || (SMEM_LIMIT - 2 * epi - 1536) / PAIR_STAGE >= 3;
    if (!no_pairs && bn == 256 && p.splits == 1 && tiles >= pg_num_sms() / 2 && fits) {
      p.num_n_blk = (p.N + 255) / 256;
      // Short-K tiles (K <= 1024) spend most of their time in the epilogue: a third epilogue warpgroup hides more
      // of its latency, provided its two extra staging stages still leave a 3-deep operand pipeline.
      static const bool no_g3 = getenv("PG_GEMM_NO_G3") != nullptr;
      if (!no_g3 && staged && !p.res_bf16 && p.k_iters <= 16 && (SMEM_LIMIT - 6 * epi - 1536) / PAIR_STAGE >= 3)
        return launch_tc<256, false, B_MN, true, 3>(A, lda, B, ldb, p, stream);
      return launch_tc<256, false, B_MN, true>(A, lda, B, ldb, p, stream);
    }
  }
  if (bn == 256) {
    // fp32-heavy staged epilogues (residual stream in/out) need more slab space than a 256-wide tile leaves
    // next to a >= 3-deep operand pipeline; those GEMMs are HBM-bound anyway, so take the 128-wide tile.
    if (staged && (SMEM_LIMIT - 2 * epi - 1536) / (A_STAGE_BYTES + 256 * BK * 2) < 3 && (p.conv.mode != 2 || p.conv.C % 128 == 0)) bn = 128;
    // Narrow problems with few tiles prefer 128 to spread over more SMs.
    if (bn == 256 && p.conv.mode != 2 && ((p.M + BM - 1) / BM) * ((p.N + 255) / 256) * p.splits < pg_num_sms() && p.N % 256 != 0) bn = 128;
  }
  p.num_n_blk = (p.N + bn - 1) / bn;
  switch (bn) {
    case 256: return launch_tc<256, A_MN, B_MN>(A, lda, B, ldb, p, stream);
    case 128: return launch_tc<128, A_MN, B_MN>(A, lda, B, ldb, p, stream);
    case 64: return launch_tc<64, A_MN, B_MN>(A, lda, B, ldb, p, stream);
    default: return launch_tc<32, A_MN, B_MN>(A, lda, B, ldb, p, stream);
  }
}

}  // namespace

static int gemm_entry(const void* A, int a_mn_major, int64_t lda, const void* B, int b_mn_major, int64_t ldb, int M, int N,
                      int K, int split_k, const pg_gemm_epilogue* epi, int impl, const ConvGeom* conv, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PG_REQUIRE(A && B && epi, "pg_gemm_bf16: null operand");
  PG_REQUIRE(M > 0 && N > 0 && K > 0, "pg_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  PG_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "pg_gemm_bf16: pitches must be multiples of 8 elements (lda=%lld ldb=%lld)",
             (long long)lda, (long long)ldb);
  PG_REQUIRE(epi->out_bf16 || epi->out_pre || epi->out_f32, "pg_gemm_bf16: no output requested");
  PG_REQUIRE(epi->dact == PG_ACT_NONE || epi->aux, "pg_gemm_bf16: dact needs aux");
  if (split_k < 1) split_k = 1;
  if (split_k > 1)
    PG_REQUIRE(epi->accumulate && epi->out_f32 && !epi->out_bf16 && !epi->out_pre && epi->dact == PG_ACT_NONE,
               "pg_gemm_bf16: split_k > 1 requires accumulate=1 into out_f32 only");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  if (conv) p.conv = *conv;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blk = (M + BM - 1) / BM;
  p.num_n_blk = 0;
  p.k_iters = (K + BK - 1) / BK;
  if (split_k > p.k_iters) split_k = p.k_iters;
  p.k_per_split = (p.k_iters + split_k - 1) / split_k;
  p.splits = (p.k_iters + p.k_per_split - 1) / p.k_per_split;  // no empty split
  p.epi = *epi;
  p.store_deriv = (epi->act & PG_ACT_STORE_DERIV) ? 1 : 0;
  p.res_bf16 = (epi->act & PG_ACT_RES_BF16) ? 1 : 0;
  p.epi.act = epi->act & 0xff;
  p.a_rowsum = epi->bias_grad;
  PG_REQUIRE(!epi->bias_grad || (a_mn_major && impl == 0),
             "pg_gemm_bf16: bias_grad rides on the weight-gradient GEMM (a_mn_major = 1, impl 0)");
  PG_REQUIRE(!p.store_deriv || epi->out_pre, "pg_gemm_bf16: PG_ACT_STORE_DERIV needs out_pre");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vec_ok = 1;
  if (epi->bias && !al16(epi->bias)) p.vec_ok = 0;
  if (epi->aux && (!al16(epi->aux) || epi->ld_aux % 8)) p.vec_ok = 0;
  const int res_mult = p.res_bf16 ? 8 : 4;
  if (epi->res0 && (!al16(epi->res0) || epi->ld_res % res_mult)) p.vec_ok = 0;
  if (epi->res1 && (!al16(epi->res1) || epi->ld_res % res_mult)) p.vec_ok = 0;
  if (epi->out_f32 && (!al16(epi->out_f32) || epi->ld_out_f32 % 4)) p.vec_ok = 0;
  if (epi->out_pre && (!al16(epi->out_pre) || epi->ld_out_pre % 8)) p.vec_ok = 0;
  if (epi->out_bf16 && (!al16(epi->out_bf16) || epi->ld_out_bf16 % 8)) p.vec_ok = 0;

  if (impl == 1) {
    p.splits = 1;
    p.k_per_split = p.k_iters;
    const int nseg = (N + 31) / 32;
    const long long total = (long long)M * nseg;
    const int threads = 128;
    const long long blocks = (total + threads - 1) / threads;
    gemm_simt_kernel<<<(unsigned)blocks, threads, 0, stream>>>(reinterpret_cast<const bf16*>(A), a_mn_major, lda,
                                                                 reinterpret_cast<const bf16*>(B), b_mn_major, ldb, p);
    return pg_check_launch("pg_gemm_bf16(simt)");
  }
  if (impl == 2) {
    // explicit opt-in (incremental sampling): never chosen implicitly, so forward() keeps one summation order
    // whatever the number of rows
    PG_REQUIRE(!a_mn_major && !b_mn_major && M <= 32 && !epi->accumulate && epi->dact == PG_ACT_NONE && K % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                   (size_t)M * K * 2 <= 160 * 1024,
               "pg_gemm_bf16(skinny): needs a K-major forward GEMM with M <= 32 rows (M=%d K=%d)", M, K);
    const size_t smem = (size_t)M * K * 2;
    PG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    gemm_skinny_kernel<<<(N + 7) / 8, 256, smem, stream>>>(reinterpret_cast<const bf16*>(A), lda,
                                                           reinterpret_cast<const bf16*>(B), ldb, p);
    return pg_check_launch("pg_gemm_bf16(skinny)");
  }
  if (!a_mn_major && !b_mn_major) return dispatch_bn<false, false>(A, lda, B, ldb, p, stream);
  if (!a_mn_major && b_mn_major) return dispatch_bn<false, true>(A, lda, B, ldb, p, stream);
  if (a_mn_major && b_mn_major) return dispatch_bn<true, true>(A, lda, B, ldb, p, stream);
  return dispatch_bn<true, false>(A, lda, B, ldb, p, stream);
}

extern "C" int pg_gemm_bf16(const void* A, int a_mn_major, int64_t lda, const void* B, int b_mn_major, int64_t ldb,
                            int M, int N, int K, int split_k, const pg_gemm_epilogue* epi, int impl, void* stream_) {
  return gemm_entry(A, a_mn_major, lda, B, b_mn_major, ldb, M, N, K, split_k, epi, impl, nullptr, stream_);
}

// Tap-loop convolution on the same kernel (see ConvGeom): forward / dgrad read the activation (or output-gradient)
// tensor under each tap's shift, wgrad reads the shifted activations as its B operand.
extern "C" int pg_gemm_bf16_conv(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int split_k,
                                 const pg_gemm_epilogue* epi, const pg_conv_geom* g, void* stream_) {
  PG_REQUIRE(g, "pg_gemm_bf16_conv: null geometry");
  PG_REQUIRE(g->mode >= PG_CONV_FWD && g->mode <= PG_CONV_WGRAD, "pg_gemm_bf16_conv: mode %d", g->mode);
  PG_REQUIRE(g->n_taps >= 1 && g->n_taps <= 32, "pg_gemm_bf16_conv: 1..32 taps (got %d)", g->n_taps);
  PG_REQUIRE(g->C % 64 == 0 && g->W >= 1 && g->W <= 64 && 64 % g->W == 0 && ((int64_t)g->H * g->W) % 128 == 0,
             "pg_gemm_bf16_conv: needs C %% 64 == 0, W | 64 and H*W %% 128 == 0 (C=%d H=%d W=%d); use pg_tap_gather otherwise",
             g->C, g->H, g->W);
  const int64_t P = (int64_t)g->N * g->H * g->W;
  ConvGeom cg;
  memset(&cg, 0, sizeof(cg));
  cg.H = g->H; cg.W = g->W; cg.C = g->C; cg.T = g->n_taps;
  cg.cslabs = g->C / 64;
  for (int t = 0; t < g->n_taps; ++t) {
    PG_REQUIRE(g->dy[t] >= -64 && g->dy[t] <= 64 && g->dx[t] >= -64 && g->dx[t] <= 64, "pg_gemm_bf16_conv: tap offset out of range");
    cg.dy[t] = (int8_t)g->dy[t];
    cg.dx[t] = (int8_t)g->dx[t];
  }
  if (g->mode == PG_CONV_WGRAD) {
    // dW[cout, t*C + c] += sum_p dY[p, cout] * X[p + off_t, c]:  A = dY (MN-major), B = X (shifted), K = pixels
    PG_REQUIRE(K == P && N == g->n_taps * g->C, "pg_gemm_bf16_conv(wgrad): K must be N*H*W and N = taps * C");
    cg.mode = 2;
    return gemm_entry(A, 1, lda, B, 1, ldb, M, N, K, split_k, epi, 0, &cg, stream_);
  }
  // forward: A = X (shifted), B = W [Cout, T*C] K-major.  dgrad: A = dY (shifted by -off), B = W [C, T*N] read MN-major.
  PG_REQUIRE(M == P && K == g->n_taps * g->C, "pg_gemm_bf16_conv: M must be N*H*W and K = taps * C");
  cg.mode = 1;
  return gemm_entry(A, 0, lda, B, g->mode == PG_CONV_DGRAD ? 1 : 0, ldb, M, N, K, split_k, epi, 0, &cg, stream_);
}
