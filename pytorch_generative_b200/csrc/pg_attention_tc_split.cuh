// pg_attention_tc_split.cuh — EXPERIMENTAL causal-attention backward (impl = 2 of pg_causal_attn_bwd), not the
// product path: compiled and exported so that round 2 can validate it on hardware, never selected by default.
//
// Same tile algebra, shared-memory layout and TMEM columns as attn_bwd_tc_kernel<64> (pg_attention_tc.cuh); what
// changes is the hand-off granularity.  The 128-key tile is treated as two 64-key halves:
//   * S and dP of half h live in TMEM columns [64h, 64h + 64) of the S / dP blocks, each half has its own
//     s_full[h] (tensor core -> threads) and s_empty[h] (threads -> tensor core, raised as soon as the half has been
//     copied into registers);
//   * all 256 softmax threads work on half 0, then on half 1 (thread = query row x 32-column group), so the tensor
//     core recomputes S / dP of the NEXT query tile for half 0 while the threads are still busy with half 1 of the
//     current one, and for half 1 while they store P / dS;
//   * dQ(i) is flushed in the middle of tile i + 1 (between the halves), long after its MMA has completed.
// Per tile the serial chain is then the thread work alone; the five products run underneath it.
// Profile that motivated it: profiles/r01_attn_bwd_ncu.txt (softmax warps wait ~25-30 % of the time for S / dP).
#pragma once

namespace {

template <int DV>
__global__ void __launch_bounds__(320, 1)
attn_bwd_split_kernel(const __grid_constant__ AttnTmaps tm, const AttnArgs a, const int T) {
  static_assert(DV == 64, "the split-phase backward is written for 64-wide value slots");
  constexpr int V_BYTES = DV * 256;
  constexpr int TMEM_COLS = 512;
  constexpr int COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 384, COL_DQ = 448;
  constexpr int HALF_ROWS_BYTES = 64 * 128;  // 64 key rows of a [128 rows][128 B] swizzle atom
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + ATOM_BYTES;
  uint8_t* sQ = sV + V_BYTES;            // 2 stages
  uint8_t* sdO = sQ + 2 * ATOM_BYTES;    // 2 stages
  uint8_t* sP = sdO + 2 * V_BYTES;       // 2 atoms (atom h = key half h)
  uint8_t* sdS = sP + 2 * ATOM_BYTES;    // 2 atoms
  uint8_t* sdQ = sdS + 2 * ATOM_BYTES;   // fp32 [2 slabs][128][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdQ + 2 * ATOM_BYTES);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;     // [2] S / dP of key half h written by the tensor core
  uint64_t* s_empty = bars + 7;    // [2] ... copied to registers by all 256 threads
  uint64_t* pds_full = bars + 9;   // P and dS of the whole tile in smem (256 arrivals)
  uint64_t* pds_empty = bars + 10; // dQ / dV / dK products have finished reading P and dS
  uint64_t* dq_full = bars + 11;   // dQ of the tile complete in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = blockIdx.x % (a.N * a.H);
  const int j = blockIdx.x / (a.N * a.H);  // key tile; small j = most query tiles = scheduled first
  const int n = nh / a.H, h = nh % a.H;
  const int niter = T - j;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("pg attention: shared memory base not 1024B aligned\n"); __trap(); }
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 256);
    }
    mbar_init(pds_full, 256);
    mbar_init(pds_empty, 1);
    mbar_init(dq_full, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, ATOM_BYTES + V_BYTES);
      tma_load_3d(sK, &tm.k, kv_full, h * 64, j * AT, n);
      tma_load_3d(sV, &tm.v, kv_full, h * DV, j * AT, n);
      for (int it = 0; it < niter; ++it) {
        const int st = it & 1, i = j + it;
        mbar_wait(&qdo_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], ATOM_BYTES + V_BYTES);
        tma_load_3d(sQ + st * ATOM_BYTES, &tm.q, &qdo_full[st], h * 64, i * AT, n);
        tma_load_3d(sdO + st * V_BYTES, &tm.d_o, &qdo_full[st], h * DV, i * AT, n);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc_half = umma_idesc_bf16(128, 64, 0, 0);  // S_h = Q K_h^T, dP_h = dO V_h^T  (N = 64 keys)
      constexpr uint32_t idesc_dv = umma_idesc_bf16(128, DV, 1, 1);    // dV += P^T dO
      constexpr uint32_t idesc_dk = umma_idesc_bf16(128, 64, 1, 1);    // dK += dS^T Q
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, 0, 1);    // dQ  = dS K
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      // key half `half` of S and dP for the query tile staged in `st`
      auto issue_half = [&](int st, int half) {
        const uint32_t q_addr = smem_u32(sQ + st * ATOM_BYTES), do_addr = smem_u32(sdO + st * V_BYTES);
        const uint32_t kh = k_addr + half * HALF_ROWS_BYTES, vh = v_addr + half * HALF_ROWS_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ss(tmem + COL_S + half * 64, desc_kmajor(q_addr, kk), desc_kmajor(kh, kk), idesc_half, kk > 0);
#pragma unroll
        for (int kk = 0; kk < DV / 16; ++kk)
          umma_bf16_ss(tmem + COL_DP + half * 64, desc_kmajor(do_addr, kk), desc_kmajor(vh, kk), idesc_half, kk > 0);
        umma_commit(&s_full[half]);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&qdo_full[0], 0);
      tc_fence_after();
      issue_half(0, 0);
      issue_half(0, 1);
      for (int it = 0; it < niter; ++it) {
        const int st = it & 1;
        const uint32_t q_addr = smem_u32(sQ + st * ATOM_BYTES), do_addr = smem_u32(sdO + st * V_BYTES);
        if (it + 1 < niter) {
          // next query tile: each half as soon as the threads have lifted the current one out of TMEM
          mbar_wait(&qdo_full[st ^ 1], ((it + 1) >> 1) & 1);
          mbar_wait(&s_empty[0], it & 1);
          tc_fence_after();
          issue_half(st ^ 1, 0);
          mbar_wait(&s_empty[1], it & 1);
          tc_fence_after();
          issue_half(st ^ 1, 1);
        }
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // K = 128 keys
          umma_bf16_ss(tmem + COL_DQ, desc_kmajor(ds_addr, kk), desc_mnmajor(k_addr, kk), idesc_dq, kk > 0);
        umma_commit(dq_full);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // K = 128 queries
          umma_bf16_ss(tmem + COL_DV, desc_mnmajor(p_addr, kk), desc_mnmajor(do_addr, kk), idesc_dv, (it > 0 || kk > 0));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(tmem + COL_DK, desc_mnmajor(ds_addr, kk), desc_mnmajor(q_addr, kk), idesc_dk, (it > 0 || kk > 0));
        umma_commit(&qdo_empty[st]);  // Q_i / dO_i stage reusable once these complete
        umma_commit(pds_empty);       // ... and so are the P / dS tiles
      }
    }
  } else {
    // ===================== softmax threads: thread == (query row, 32-column group of the current half) ==========
    const int grp = warp >> 2;              // columns [32 grp, 32 grp + 32) of the half
    const int r = (warp & 3) * 32 + lane;   // TMEM lane / tile row
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float sl2 = a.scale * 1.4426950408889634f;
    const int k0 = j * AT;
    const size_t stat_base = ((size_t)n * a.H + h) * a.S;
    float lse_next = 0.f, delta_next = 0.f;
    if (j * AT + r < a.S) {
      lse_next = a.lse_in[stat_base + j * AT + r];
      delta_next = a.delta[stat_base + j * AT + r];
    }
    // dQ of query tile i: this group's 32 columns -> fp32 staging slab -> one bulk reduce-add per slab
    auto flush_dq = [&](int i, uint32_t parity) {
      mbar_wait(dq_full, parity);
      tc_fence_after();
      if (threadIdx.x == 0) tma_store_wait_read<0>();  // previous reduce has drained the staging slabs
      __syncwarp();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + COL_DQ + lane_base + grp * 32, v);
      tmem_wait_ld();
      slab32_store_scaled(sdQ + grp * ATOM_BYTES, r, v, a.scale);
      fence_proxy_async_smem();
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        tma_reduce_add_3d(&tm.dq, sdQ, h * 64, i * AT, n);
        tma_reduce_add_3d(&tm.dq, sdQ + ATOM_BYTES, h * 64 + 32, i * AT, n);
        tma_store_commit();
      }
      __syncwarp();
    };
    for (int it = 0; it < niter; ++it) {
      const int i = j + it;
      const int qi = i * AT + r;
      const bool row_ok = qi < a.S;
      const float lse2 = lse_next * 1.4426950408889634f;
      const float delta = delta_next;
      if (it + 1 < niter && qi + AT < a.S) {
        lse_next = a.lse_in[stat_base + qi + AT];
        delta_next = a.delta[stat_base + qi + AT];
      } else {
        lse_next = 0.f;
        delta_next = 0.f;
      }
      const int qlim = row_ok ? qi - a.strict : -1;  // invalid rows see no keys
      const bool need_mask = (it == 0) || (i == T - 1);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int c = half * 2 + grp;  // 32-column chunk of the 128-key tile
        mbar_wait(&s_full[half], it & 1);
        tc_fence_after();
        uint32_t sv[32], dv[32];
        tmem_ld_32x32b_x32(tmem + COL_S + lane_base + c * 32, sv);
        tmem_ld_32x32b_x32(tmem + COL_DP + lane_base + c * 32, dv);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&s_empty[half]);  // the tensor core may overwrite this half with the next query tile
        // the products of the previous tile read sP / sdS until pds_empty: wait before this tile's first store
        if (half == 0 && it > 0) mbar_wait(pds_empty, (it - 1) & 1);
        uint32_t pw[16], dw[16];
        auto p_ds = [&](auto masked) {
          constexpr bool MASK = decltype(masked)::value;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), sl2, -lse2));
            float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), sl2, -lse2));
            if (MASK) {
              if (k0 + c * 32 + e > qlim) p0 = 0.f;
              if (k0 + c * 32 + e + 1 > qlim) p1 = 0.f;
            }
            const float d0 = p0 * (__uint_as_float(dv[e]) - delta);
            const float d1 = p1 * (__uint_as_float(dv[e + 1]) - delta);
            pw[e >> 1] = pack_bf16x2(p0, p1);
            dw[e >> 1] = pack_bf16x2(d0, d1);
          }
        };
        if (need_mask) p_ds(std::true_type{});
        else p_ds(std::false_type{});
        store_tile_row_chunk(sP, r, c, pw);
        store_tile_row_chunk(sdS, r, c, dw);
        // between the halves: dQ of the previous query tile (its MMA finished long ago) -> reduce-add
        if (half == 0 && it > 0) flush_dq(i - 1, (it - 1) & 1);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
    }
    flush_dq(j + niter - 1, (niter - 1) & 1);
    if (threadIdx.x == 0) tma_store_wait<0>();
    __syncwarp();
    mbar_wait(pds_empty, (niter - 1) & 1);  // the whole dV / dK accumulation is complete
    tc_fence_after();
    // dV_j, dK_j: thread == key row; each group writes its half of the columns.  TMEM loads are warp-aligned:
    // every lane executes them, only the stores are predicated.
    const int kj = k0 + r;
    const bool key_ok = kj < a.S;
    bf16* dvrow = a.dv_out + ((size_t)n * a.S + kj) * a.ld_dv + h * DV;
    {
      const int c = grp;
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + COL_DV + lane_base + c * 32, v);
      tmem_wait_ld();
      if (key_ok) {
#pragma unroll
        for (int e = 0; e < 32; e += 8)
          *reinterpret_cast<uint4*>(dvrow + c * 32 + e) =
              make_uint4(pack_bf16x2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])),
                         pack_bf16x2(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])),
                         pack_bf16x2(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])),
                         pack_bf16x2(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7])));
      }
    }
    bf16* dkrow = a.dk_out + ((size_t)n * a.S + kj) * a.ld_dk + h * 64;
    {
      const int c = grp;
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + COL_DK + lane_base + c * 32, v);
      tmem_wait_ld();
      if (key_ok) {
#pragma unroll
        for (int e = 0; e < 32; e += 8)
          *reinterpret_cast<uint4*>(dkrow + c * 32 + e) = make_uint4(
              pack_bf16x2(__uint_as_float(v[e]) * a.scale, __uint_as_float(v[e + 1]) * a.scale),
              pack_bf16x2(__uint_as_float(v[e + 2]) * a.scale, __uint_as_float(v[e + 3]) * a.scale),
              pack_bf16x2(__uint_as_float(v[e + 4]) * a.scale, __uint_as_float(v[e + 5]) * a.scale),
              pack_bf16x2(__uint_as_float(v[e + 6]) * a.scale, __uint_as_float(v[e + 7]) * a.scale));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// impl = 2 of pg_causal_attn_bwd (experimental, dv slot 64 only); same argument checks and dq conversion as attn_bwd_tc.
int attn_bwd_tc_split(const AttnArgs& a, cudaStream_t stream) {
  if (attn_check_tc(a, "pg_causal_attn_bwd")) return 1;
  PG_REQUIRE(a.dv == 64, "pg_causal_attn_bwd(impl 2): the split-phase kernel handles 64-wide value slots only");
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
  const unsigned grid = (unsigned)(a.N * a.H * T);
  constexpr int SMEM = ATOM_BYTES * (1 + 1 + 2 + 2 + 2 + 2 + 2) + 256;
  PG_CUDA(cudaFuncSetAttribute(attn_bwd_split_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  attn_bwd_split_kernel<64><<<grid, 320, SMEM, stream>>>(tm, a, T);
  if (pg_check_launch("pg_causal_attn_bwd(tcgen05 split)")) return 1;
  const long long P = (long long)a.N * a.S;
  const int width = a.H * 64;
  long long blocks = (P * (width / 8) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  attn_dq_convert_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a.dq_accum, a.dq, a.ld_dq, P, width);
  return pg_check_launch("pg_causal_attn_bwd(dq convert)");
}

}  // namespace
