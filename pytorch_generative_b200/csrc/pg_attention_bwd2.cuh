// pg_attention_bwd2.cuh — causal-attention backward, persistent and stage-pipelined (included by pg_attention.cu).
//
// Same tile algebra, operand layouts and descriptors as attn_bwd_tc_kernel (pg_attention_tc.cuh): one work item is a
// 128-key tile j of one (image, head); the item loops over the query tiles i >= j,
//     S = Q_i K_j^T,  dP = dO_i V_j^T,  P = exp2(S c - lse),  dS = P (dP - delta),
//     dV_j += P^T dO_i,  dK_j += dS^T Q_i  (TMEM accumulators),   dQ_i += dS K_j  (fp32 reduce-add into dq_accum).
// What changed against the round-1 kernel (profiles/r01_attn_bwd_ncu.txt: tensor pipe 18 %, the softmax threads and
// the tensor core strictly alternating, ~12 % of the time in CTA start-up):
//   * the thread work of a tile is split in two stages, A: S -> P and B: dP -> dS, and the five products are
//     interleaved between them, so that every product the threads wait for was issued one stage earlier:
//         threads  | A(i)            | B(i)              | A(i+1)               | B(i+1) ...
//         tensor   | dP(i) ..        | S(i+1), dV(i)     | dK(i), dP(i+1), dQ(i)| S(i+2), dV(i+1) ...
//     S and dP stay single-buffered in TMEM (S is free again once every thread has read it = when P is published);
//   * dQ is drained by its own warpgroup (TMEM -> fp32 staging slab -> TMA reduce-add), off the softmax threads;
//   * the CTA is persistent: work items are dealt statically, balanced and L2-local (see Bwd2Cursor), K/V of the next item and
//     Q / dO of the next tiles are prefetched through multi-stage rings, TMEM and barriers are set up once.
//
// Warp roles (SW = 2: 448 threads, SW = 4: 704): 4 * SW softmax warps, group g = key columns [g * 128 / SW, ...), thread =
// query row; then 4 warps of dQ drain (thread = query row) + dV / dK read-out at the end of an item (thread = key row); 12 TMA producer; 13 TMEM owner + MMA issuer.
#pragma once

namespace {

template <int DV>
struct Bwd2 {
  static constexpr int NKV = (DV == 64) ? 2 : 1;  // K/V stages (items)
  static constexpr int NQ = (DV == 64) ? 3 : 2;   // Q stages (tiles)
  static constexpr int NDO = 2;                   // dO stages (tiles)
  static constexpr int NDQ = (DV == 64) ? 2 : 1;  // dQ accumulators in TMEM
  static constexpr int V_BYTES = DV * 256;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + NKV * ATOM_BYTES;
  static constexpr int OFF_Q = OFF_V + NKV * V_BYTES;
  static constexpr int OFF_DO = OFF_Q + NQ * ATOM_BYTES;
  static constexpr int OFF_P = OFF_DO + NDO * V_BYTES;
  static constexpr int OFF_DS = OFF_P + 2 * ATOM_BYTES;
  static constexpr int OFF_DQ = OFF_DS + 2 * ATOM_BYTES;  // one fp32 slab [128][32]
  static constexpr int OFF_BAR = OFF_DQ + ATOM_BYTES;
  static constexpr int SMEM = OFF_BAR + 256;
  static constexpr int COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 256 + DV, COL_DQ = COL_DK + 64;
  static_assert(COL_DQ + 64 * NDQ <= 512, "TMEM budget");
  static_assert(SMEM <= 232448, "shared memory budget");
};

// Position in the CTA's work stream: item = (key tile j, image n, head h), tile `it` of its T - j query tiles.
//
// Scheduling.  The grid is G groups of `slots` = ceil(T / 2) CTAs.  A group works on one (image, head) at a time
// (group g takes nh = g, g + G, ...); CTA r of the group takes key tiles r and T - 1 - r of it, i.e. T + 1 query tiles
// whatever r is, so the CTAs stay balanced, and the 148 CTAs touch only ~G (image, head) pairs at any time: Q, dO and
// the fp32 dQ accumulator of those pairs (0.5 MB each at S = 1024) stay in L2 across the key tiles that re-read /
// re-accumulate them.  (Dealing the items longest-first over the whole batch instead made every key tile re-stream
// its pair's Q / dO / dQ from HBM: 1.9 GB of DRAM traffic per call at N = 64, profiles/r02_attn_bwd2_v1_ncu.txt.)
struct Bwd2Cursor {
  int q;  // index in the CTA's item sequence: (round, sub) = (q / 2, q % 2)
  int n, h, j, niter, it;
  bool valid;
};
__device__ __forceinline__ void bwd2_item(Bwd2Cursor& c, int q, const AttnArgs& a, int T) {
  const int NH = a.N * a.H;
  const int slots = (T + 1) >> 1;
  const int G = (int)gridDim.x / slots, grp = (int)blockIdx.x / slots, r = (int)blockIdx.x % slots;
  c.it = 0;
  for (;; ++q) {
    const int nh = grp + G * (q >> 1);
    c.q = q;
    c.valid = nh < NH;
    if (!c.valid) {
      c.j = c.n = c.h = 0;
      c.niter = 1;
      return;
    }
    const int j = (q & 1) ? T - 1 - r : r;
    if ((q & 1) && j == r) continue;  // odd T: the middle key tile has no partner
    c.j = j;
    c.n = nh / a.H;
    c.h = nh % a.H;
    c.niter = T - j;
    return;
  }
}
// returns true when the step crossed an item boundary
__device__ __forceinline__ bool bwd2_next(Bwd2Cursor& c, const AttnArgs& a, int T) {
  if (++c.it < c.niter) return false;
  bwd2_item(c, c.q + 1, a, T);
  return true;
}

// SW = softmax warpgroups: each covers 128 / SW key columns of every query row (2: 64 columns per thread, 448 threads;
// 4: 32 columns per thread, 704 threads -- twice the warps per scheduler to hide the TMEM-load / fence / barrier latencies
// of a stage, which is what bounds the 2-group kernel: profiles/r02_attn_bwd2_v2_ncu.txt).
template <int DV, int SW>
__global__ void __launch_bounds__(SW * 128 + 192, 1)
attn_bwd2_kernel(const __grid_constant__ AttnTmaps tm, const AttnArgs a, const int T) {
  using C = Bwd2<DV>;
  constexpr int CPT = 128 / SW;       // key columns per softmax thread
  constexpr int NCH = CPT / 32;       // 32-column chunks per thread
  constexpr int W_DRAIN = 4 * SW;     // first warp of the dQ drain group
  constexpr int W_TMA = 4 * SW + 4, W_MMA = 4 * SW + 5;
  constexpr int V_BYTES = C::V_BYTES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + C::OFF_K;
  uint8_t* sV = smem + C::OFF_V;
  uint8_t* sQ = smem + C::OFF_Q;
  uint8_t* sdO = smem + C::OFF_DO;
  uint8_t* sP = smem + C::OFF_P;
  uint8_t* sdS = smem + C::OFF_DS;
  uint8_t* sdQ = smem + C::OFF_DQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* kv_full = bars;            // [2]
  uint64_t* kv_empty = bars + 2;       // [2]
  uint64_t* q_full = bars + 4;         // [3]
  uint64_t* q_empty = bars + 7;        // [3]
  uint64_t* do_full = bars + 10;       // [2]
  uint64_t* do_empty = bars + 12;      // [2]
  uint64_t* s_full = bars + 14;        // S of the tile complete in TMEM
  uint64_t* dp_full = bars + 15;       // dP ...
  uint64_t* p_full = bars + 16;        // P in smem, S read by every softmax thread (8 warp arrivals)
  uint64_t* ds_full = bars + 17;       // dS in smem, dP read (8 warp arrivals)
  uint64_t* p_free = bars + 18;        // dV product has finished reading P
  uint64_t* ds_free = bars + 19;       // dK and dQ products have finished reading dS
  uint64_t* dq_full = bars + 20;       // [2] dQ of the tile complete in TMEM
  uint64_t* dq_empty = bars + 22;      // [2] ... drained (4 warp arrivals)
  uint64_t* acc_full = bars + 24;      // dV / dK of the item complete
  uint64_t* acc_empty = bars + 25;     // ... read out by the drain warpgroup (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("pg attention: shared memory base not 1024B aligned\n"); __trap(); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
      mbar_init(&do_full[s], 1); mbar_init(&do_empty[s], 1);
      mbar_init(&dq_full[s], 1); mbar_init(&dq_empty[s], 4);
    }
    for (int s = 0; s < 3; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(dp_full, 1);
    mbar_init(p_full, 4 * SW); mbar_init(ds_full, 4 * SW);  // one elected arrival per softmax warp
    mbar_init(p_free, 1); mbar_init(ds_free, 1);
    mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == W_MMA) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  Bwd2Cursor c;
  bwd2_item(c, 0, a, T);
  // development timeline (build with -DPG_ATTN_TRACE; pg_debug_set_trace, tools/attn_trace.py): role r stamps clock64()
  // into trace[r * 4096 + k].  Compiled out by default: the volatile clock reads pin the instruction schedule around
  // every wait (measured: 572 vs 501 us per call with the stamps compiled in but switched off).
#ifdef PG_ATTN_TRACE
  int tr_n = 0;
  auto TR = [&](int role) {
    if (a.trace != nullptr && blockIdx.x == 0 && lane == 0 && tr_n < 4096) a.trace[role * 4096 + tr_n++] = clock64();
  };
#else
  auto TR = [](int) {};
#endif

  if (warp == W_TMA) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    {
      unsigned g = 0, m = 0;
      while (c.valid) {
        if (c.it == 0) {
          const unsigned ks = m % C::NKV;
          mbar_wait(&kv_empty[ks], ((m / C::NKV) & 1u) ^ 1u);
          mbar_arrive_expect_tx_w(&kv_full[ks], ATOM_BYTES + V_BYTES);
          tma_load_3d_w(sK + ks * ATOM_BYTES, &tm.k, &kv_full[ks], c.h * 64, c.j * AT, c.n);
#pragma unroll
          for (int v = 0; v < DV / 64; ++v)
            tma_load_3d_w(sV + ks * V_BYTES + v * ATOM_BYTES, &tm.v, &kv_full[ks], c.h * DV + v * 64, c.j * AT, c.n);
        }
        const int i = c.j + c.it;
        const unsigned qs = g % C::NQ, os = g % C::NDO;
        mbar_wait(&q_empty[qs], ((g / C::NQ) & 1u) ^ 1u);
        TR(0);
        mbar_arrive_expect_tx_w(&q_full[qs], ATOM_BYTES);
        tma_load_3d_w(sQ + qs * ATOM_BYTES, &tm.q, &q_full[qs], c.h * 64, i * AT, c.n);
        mbar_wait(&do_empty[os], ((g / C::NDO) & 1u) ^ 1u);
        TR(0);
        mbar_arrive_expect_tx_w(&do_full[os], V_BYTES);
#pragma unroll
        for (int v = 0; v < DV / 64; ++v)
          tma_load_3d_w(sdO + os * V_BYTES + v * ATOM_BYTES, &tm.d_o, &do_full[os], c.h * DV + v * 64, i * AT, c.n);
        ++g;
        if (bwd2_next(c, a, T)) ++m;
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues: see umma_bf16_ss_w) =====================
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S = Q K^T, dP = dO V^T
      constexpr uint32_t idesc_dv = umma_idesc_bf16(128, DV, 1, 1);   // dV += P^T dO
      constexpr uint32_t idesc_dk = umma_idesc_bf16(128, 64, 1, 1);   // dK += dS^T Q
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, 0, 1);   // dQ  = dS K
      // descriptors: one base per operand tile, the K step only moves the 14-bit start-address field
      //   K-major:  step kk -> atom kk / 4, 32 bytes per step inside the atom;   MN-major: 16 rows = 2048 bytes per step
      auto kstep = [](uint64_t base, int kk) { return base + (uint64_t)((kk >> 2) * (ATOM_BYTES >> 4) + (kk & 3) * 2); };
      auto mnstep = [](uint64_t base, int kk) { return base + (uint64_t)(kk * 128); };
      const uint64_t p_mn = umma_desc_sw128(smem_u32(sP), ATOM_BYTES, 1024);
      const uint64_t ds_mn = umma_desc_sw128(smem_u32(sdS), ATOM_BYTES, 1024);
      const uint64_t ds_k = umma_desc_sw128(smem_u32(sdS), 16, 1024);
      auto k_of = [&](unsigned m) { return smem_u32(sK + (m % C::NKV) * ATOM_BYTES); };
      auto v_of = [&](unsigned m) { return smem_u32(sV + (m % C::NKV) * V_BYTES); };
      auto q_of = [&](unsigned g) { return smem_u32(sQ + (g % C::NQ) * ATOM_BYTES); };
      auto do_of = [&](unsigned g) { return smem_u32(sdO + (g % C::NDO) * V_BYTES); };
      const bool mma_on = a.dbg != 1;
      auto issue_s = [&](unsigned g, unsigned m, bool first_of_item) {
        if (first_of_item) mbar_wait(&kv_full[m % C::NKV], (m / C::NKV) & 1u);
        mbar_wait(&q_full[g % C::NQ], (g / C::NQ) & 1u);
        tc_fence_after();
        const uint64_t qd = umma_desc_sw128(q_of(g), 16, 1024), kd = umma_desc_sw128(k_of(m), 16, 1024);
        if (mma_on) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_ss_w(tmem + C::COL_S, kstep(qd, kk), kstep(kd, kk), idesc_s, kk > 0);
        }
        umma_commit_w(s_full);
      };
      auto issue_dp = [&](unsigned g, unsigned m) {
        mbar_wait(&do_full[g % C::NDO], (g / C::NDO) & 1u);
        tc_fence_after();
        const uint64_t dd = umma_desc_sw128(do_of(g), 16, 1024), vd = umma_desc_sw128(v_of(m), 16, 1024);
        if (mma_on) {
#pragma unroll
          for (int kk = 0; kk < DV / 16; ++kk) umma_bf16_ss_w(tmem + C::COL_DP, kstep(dd, kk), kstep(vd, kk), idesc_s, kk > 0);
        }
        umma_commit_w(dp_full);
      };
      unsigned g = 0, m = 0;
      if (c.valid) {
        issue_s(0, 0, true);
        issue_dp(0, 0);
      }
      while (c.valid) {
        Bwd2Cursor nx = c;
        const bool crosses = bwd2_next(nx, a, T);
        const unsigned mn = crosses ? m + 1 : m;
        // with a single K/V stage the next item's operands cannot land before this item releases them
        const bool defer = (C::NKV == 1) && crosses;
        const bool first = c.it == 0, last = c.it == c.niter - 1;
        const uint64_t q_mn = umma_desc_sw128(q_of(g), ATOM_BYTES, 1024), do_mn = umma_desc_sw128(do_of(g), ATOM_BYTES, 1024);
        const uint64_t k_mn = umma_desc_sw128(k_of(m), ATOM_BYTES, 1024);

        TR(1);
        mbar_wait(p_full, g & 1u);  // P(g) published; S is free
        TR(1);
        tc_fence_after();
        if (nx.valid && !defer) issue_s(g + 1, mn, crosses);
        if (first) {  // the previous item's dV / dK have been read out
          mbar_wait(acc_empty, (m & 1u) ^ 1u);
          tc_fence_after();
        }
        if (mma_on) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // K = 128 queries
            umma_bf16_ss_w(tmem + C::COL_DV, mnstep(p_mn, kk), mnstep(do_mn, kk), idesc_dv, (!first || kk > 0));
        }
        umma_commit_w(p_free);
        umma_commit_w(&do_empty[g % C::NDO]);

        TR(1);
        mbar_wait(ds_full, g & 1u);  // dS(g) published; dP is free
        TR(1);
        tc_fence_after();
        if (mma_on) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16_ss_w(tmem + C::COL_DK, mnstep(ds_mn, kk), mnstep(q_mn, kk), idesc_dk, (!first || kk > 0));
        }
        umma_commit_w(&q_empty[g % C::NQ]);
        if (nx.valid && !defer) issue_dp(g + 1, mn);
        const unsigned b = g % C::NDQ;
        TR(1);
        mbar_wait(&dq_empty[b], ((g / C::NDQ) & 1u) ^ 1u);
        TR(1);
        tc_fence_after();
        if (mma_on) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // K = 128 keys
            umma_bf16_ss_w(tmem + C::COL_DQ + b * 64, kstep(ds_k, kk), mnstep(k_mn, kk), idesc_dq, kk > 0);
        }
        umma_commit_w(&dq_full[b]);
        umma_commit_w(ds_free);
        if (last) {
          umma_commit_w(acc_full);
          umma_commit_w(&kv_empty[m % C::NKV]);
        }
        if (nx.valid && defer) {
          issue_s(g + 1, mn, true);
          issue_dp(g + 1, mn);
        }
        c = nx;
        m = mn;
        ++g;
      }
    }
  } else if (warp >= W_DRAIN) {
    // ===================== dQ drain: TMEM -> fp32 slab -> TMA reduce-add =====================
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const bool leader = threadIdx.x == W_DRAIN * 32;
    unsigned g = 0, m = 0;
    while (c.valid) {
      const unsigned b = g % C::NDQ;
      const int i = c.j + c.it;
      TR(2);
      mbar_wait(&dq_full[b], (g / C::NDQ) & 1u);
      TR(2);
      tc_fence_after();
      if (a.dbg != 4) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem + C::COL_DQ + b * 64 + lane_base + half * 32, v);
          tmem_wait_ld();
          if (half == 1) {  // the accumulator has been copied out: the tensor core may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dq_empty[b]);
          }
          if (leader) tma_store_wait_read<0>();  // the previous reduce has finished reading the slab
          __syncwarp();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          slab32_store_scaled(sdQ, r, v, a.scale);
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (leader && a.dbg != 3) {
            tma_reduce_add_3d(&tm.dq, sdQ, c.h * 64 + half * 32, i * AT, c.n);
            tma_store_commit();
          }
          __syncwarp();
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dq_empty[b]);
      }
      if (c.it == c.niter - 1) {
        // ---- dV_j, dK_j read-out (thread == key row): off the softmax threads, which move on to the next item ----
        mbar_wait(acc_full, m & 1u);
        tc_fence_after();
        const int kj = c.j * AT + r;
        const bool key_ok = kj < a.S;
        bf16* dvrow = a.dv_out + ((size_t)c.n * a.S + kj) * a.ld_dv + c.h * DV;
#pragma unroll
        for (int col = 0; col < DV; col += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem + C::COL_DV + lane_base + col, v);
          tmem_wait_ld();
          if (key_ok) {
#pragma unroll
            for (int e = 0; e < 32; e += 8)
              *reinterpret_cast<uint4*>(dvrow + col + e) =
                  make_uint4(pack_bf16x2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])),
                             pack_bf16x2(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])),
                             pack_bf16x2(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])),
                             pack_bf16x2(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7])));
          }
        }
        bf16* dkrow = a.dk_out + ((size_t)c.n * a.S + kj) * a.ld_dk + c.h * 64;
#pragma unroll
        for (int col = 0; col < 64; col += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem + C::COL_DK + lane_base + col, v);
          tmem_wait_ld();
          if (key_ok) {
#pragma unroll
            for (int e = 0; e < 32; e += 8)
              *reinterpret_cast<uint4*>(dkrow + col + e) = make_uint4(
                  pack_bf16x2(__uint_as_float(v[e]) * a.scale, __uint_as_float(v[e + 1]) * a.scale),
                  pack_bf16x2(__uint_as_float(v[e + 2]) * a.scale, __uint_as_float(v[e + 3]) * a.scale),
                  pack_bf16x2(__uint_as_float(v[e + 4]) * a.scale, __uint_as_float(v[e + 5]) * a.scale),
                  pack_bf16x2(__uint_as_float(v[e + 6]) * a.scale, __uint_as_float(v[e + 7]) * a.scale));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        ++m;
      }
      ++g;
      bwd2_next(c, a, T);
    }
    if (leader) tma_store_wait<0>();
    __syncwarp();
  } else {
    // ===================== softmax groups: thread == query row, group == key-column half =====================
    const int grp = warp >> 2;              // key columns [grp * CPT, +CPT)
    const int r = (warp & 3) * 32 + lane;   // TMEM lane / tile row
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float sl2 = a.scale * 1.4426950408889634f;
    unsigned g = 0;
    float lse_next = 0.f, delta_next = 0.f;
    auto load_stats = [&](const Bwd2Cursor& cc) {
      lse_next = 0.f;
      delta_next = 0.f;
      if (cc.valid) {
        const int q = (cc.j + cc.it) * AT + r;
        if (q < a.S) {
          const size_t sb = ((size_t)cc.n * a.H + cc.h) * a.S + q;
          lse_next = a.lse_in[sb];
          delta_next = a.delta[sb];
        }
      }
    };
    load_stats(c);
    while (c.valid) {
      const int i = c.j + c.it;
      const int k0 = c.j * AT;
      const int qi = i * AT + r;
      const bool row_ok = qi < a.S;
      const float lse2 = lse_next * 1.4426950408889634f;
      const float delta = delta_next;
      const int qlim = row_ok ? qi - a.strict : -1;  // invalid rows see no keys
      const bool need_mask = (c.it == 0) || (i == T - 1);
      Bwd2Cursor nx = c;
      bwd2_next(nx, a, T);
      load_stats(nx);  // in flight underneath this tile

      // ---- stage A: S -> P ----
      if (warp == 0) TR(3);
      mbar_wait(s_full, g & 1u);
      if (warp == 0) TR(3);
      tc_fence_after();
      uint32_t pk[NCH][16];
      const bool work = a.dbg != 2;
      if (work) {
        uint32_t sv[NCH][32];
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) tmem_ld_32x32b_x32(tmem + C::COL_S + lane_base + (grp * NCH + cc) * 32, sv[cc]);
        tmem_wait_ld();
        auto p_tile = [&](auto masked) {
          constexpr bool MASK = decltype(masked)::value;
#pragma unroll
          for (int cc = 0; cc < NCH; ++cc) {
            const int cb = k0 + (grp * NCH + cc) * 32;
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              float p0 = fmaf(__uint_as_float(sv[cc][e]), sl2, -lse2);
              float p1 = fmaf(__uint_as_float(sv[cc][e + 1]), sl2, -lse2);
              if (a.dbg != 6) {
                p0 = fast_exp2(p0);
                p1 = fast_exp2(p1);
              }
              if (MASK) {
                if (cb + e > qlim) p0 = 0.f;
                if (cb + e + 1 > qlim) p1 = 0.f;
              }
              pk[cc][e >> 1] = pack_bf16x2(p0, p1);
            }
          }
        };
        if (need_mask) p_tile(std::true_type{});
        else p_tile(std::false_type{});
      }
      if (warp == 0) TR(3);
      if (g > 0) mbar_wait(p_free, (g - 1) & 1u);  // dV(g-1) no longer reads sP
      if (warp == 0) TR(3);
      if (work && a.dbg != 5) {
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) store_tile_row_chunk(sP, r, grp * NCH + cc, pk[cc]);
      }
      if (a.dbg != 7 && a.dbg != 5) fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();  // one arrival per warp: every arrival wakes the warps parked on this SM's barriers
      if (lane == 0) mbar_arrive(p_full);

      // ---- stage B: dP -> dS = P (dP - delta) ----
      if (warp == 0) TR(3);
      mbar_wait(dp_full, g & 1u);
      if (warp == 0) TR(3);
      tc_fence_after();
      if (work) {
        uint32_t dv[NCH][32];
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) tmem_ld_32x32b_x32(tmem + C::COL_DP + lane_base + (grp * NCH + cc) * 32, dv[cc]);
        tmem_wait_ld();
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 p = unpack_bf16x2(pk[cc][e >> 1]);
            const float d0 = p.x * (__uint_as_float(dv[cc][e]) - delta);
            const float d1 = p.y * (__uint_as_float(dv[cc][e + 1]) - delta);
            pk[cc][e >> 1] = pack_bf16x2(d0, d1);
          }
        }
      }
      if (warp == 0) TR(3);
      if (g > 0) mbar_wait(ds_free, (g - 1) & 1u);  // dK(g-1), dQ(g-1) no longer read sdS
      if (warp == 0) TR(3);
      if (work && a.dbg != 5) {
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) store_tile_row_chunk(sdS, r, grp * NCH + cc, pk[cc]);
      }
      if (a.dbg != 7 && a.dbg != 5) fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);

      c = nx;
      ++g;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int attn_bwd_tc2(const AttnArgs& a, cudaStream_t stream) {
  if (attn_check_tc(a, "pg_causal_attn_bwd")) return 1;
  PG_REQUIRE(a.dq_accum != nullptr, "pg_causal_attn_bwd: dq_accum scratch is required by the tcgen05 path");
  PG_REQUIRE(a.ld_dq % 8 == 0 && a.ld_dk % 8 == 0 && a.ld_dv % 8 == 0, "pg_causal_attn_bwd: pitches must be multiples of 8");
  AttnTmaps tm;
  if (make_attn_map(&tm.q, a.q, a.ld_q, a.H * 64, a.S, a.N)) return 1;
  if (make_attn_map(&tm.k, a.k, a.ld_k, a.H * 64, a.S, a.N)) return 1;
  if (make_attn_map(&tm.v, a.v, a.ld_v, a.H * a.dv, a.S, a.N)) return 1;
  if (make_attn_map(&tm.d_o, a.d_o, a.ld_do, a.H * a.dv, a.S, a.N)) return 1;
  {
    uint64_t dims[3] = {(uint64_t)a.H * 64, (uint64_t)a.S, (uint64_t)a.N};
    uint64_t strides[2] = {(uint64_t)a.H * 64 * 4, (uint64_t)a.S * a.H * 64 * 4};
    uint32_t box[3] = {32, (uint32_t)AT, 1};
    if (pg_make_tmap_nd(&tm.dq, a.dq_accum, 4, 3, dims, strides, box, 128)) return 1;
  }
  const int T = (a.S + AT - 1) / AT;
  const int slots = (T + 1) / 2;  // CTAs per (image, head) group, see Bwd2Cursor
  int groups = pg_num_sms() / slots;
  if (groups > a.N * a.H) groups = a.N * a.H;
  if (groups < 1) groups = 1;
  const unsigned grid = (unsigned)(groups * slots);
  static const int sw = getenv("PG_ATTN_BWD_SW") ? atoi(getenv("PG_ATTN_BWD_SW")) : 4;
  auto launch = [&](auto kern, int smem, int threads) -> int {
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, threads, smem, stream>>>(tm, a, T);
    return 0;
  };
  if (a.dv == 64) {
    if (sw == 2 ? launch(attn_bwd2_kernel<64, 2>, Bwd2<64>::SMEM, 448) : launch(attn_bwd2_kernel<64, 4>, Bwd2<64>::SMEM, 704)) return 1;
  } else {
    if (sw == 2 ? launch(attn_bwd2_kernel<128, 2>, Bwd2<128>::SMEM, 448) : launch(attn_bwd2_kernel<128, 4>, Bwd2<128>::SMEM, 704)) return 1;
  }
  if (pg_check_launch("pg_causal_attn_bwd(tcgen05, pipelined)")) return 1;
  const long long P = (long long)a.N * a.S;
  const int width = a.H * 64;
  long long blocks = (P * (width / 8) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  attn_dq_convert_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a.dq_accum, a.dq, a.ld_dq, P, width);
  return pg_check_launch("pg_causal_attn_bwd(dq convert)");
}

}  // namespace
