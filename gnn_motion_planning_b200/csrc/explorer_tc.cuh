// Tensor-core (tcgen05 / TMEM, 3xTF32) version of the edge-feature stage of the explorer forward, e = 32.
//
// Same contract as edge_feature_kernel (explorer.cu): edge_free_code / edge_code encoders + the three edge Blocks
// (model.py:120,123,130,153-218) -> P = W4 ef + W5 ec + b (loop-invariant part of lin_0[0], model.py:39) and
// Q = Wc ef + b (edge part of policy[0], model.py:145), written in CSR slot order.
//
// Organisation (one persistent 384-thread CTA per SM; warps 9-11 only complete the issuer's warpgroup for setmaxnreg):
//   * warps 0-3 and 4-7 each own one 128-edge tile (thread == edge row, TMEM lane == row); the two tiles of a CTA are
//     consecutive halves of one 256-slot unit of a graph, so they share that graph's obstacle tables;
//   * warp 8 is the MMA issuer: for every stage of the per-tile program it waits until the tile's 128 threads have
//     published the stage's A operand (tcgen05.st hi/lo planes in TMEM, mbarrier `ready`), issues the 3xTF32
//     tcgen05.mma group from ONE elected lane with all operands in uniform registers, and commits to the tile's
//     `done` mbarrier.  It alternates strictly between the two tiles, so one tile's epilogue (tcgen05.ld, softmax,
//     LayerNorm, TF32 split, tcgen05.st) overlaps the other tile's MMAs;
//   * activations never touch shared memory: A operands are read from TMEM, accumulators are read back with the
//     32x32b shape (thread == row) so softmax / LayerNorm / residuals are thread-local exactly as in the SIMT kernels;
//   * shared memory holds the B operands only: all weights of this stage as hi / lo TF32 planes (resident, ~146 KB)
//     and ONE obstacle-table buffer (<= 48 KB: scale*Wq^T Wk o and Wv o of <= 96 obstacles, hi / lo) that warp 8
//     refills with a TMA bulk copy as soon as both tiles' P.V MMAs have retired -- the refill overlaps the FFN stages.
//
// TMEM columns of a tile (256 of the CTA's 512): XH [0,32) XL [32,64) A operand; A1 [64,128) accumulators
// (Gx | Vx, FFN, Q | P); SC [128,224) scores -> probabilities hi (in place); PL [0,96) probabilities lo (overlays the
// dead XH/XL/Gx); PV [224,256).
#pragma once
#include "handle.h"
#include "rowtile.cuh"
#include "umma.cuh"

namespace gmp {

// obstacle chunking shared by host and device: a graph's O obstacles are processed in nch chunks of `per` (multiple
// of 16, <= 96) table rows, zero padded
__host__ __device__ inline int tc_nchunks(int O) { return (O + 95) / 96; }
__host__ __device__ inline int tc_per(int O, int nch) { return nch ? ((O + nch - 1) / nch + 15) / 16 * 16 : 0; }

template <int C>
struct TcCfg {
  static constexpr int E = 32;
  static constexpr int K0 = (2 * C + 7) / 8 * 8;        // encoder input width padded to the MMA K step
  static constexpr int kOcMax = 96;
  // float offsets in the TC weight image; every matrix is [hi plane | lo plane], a plane is float[K/4][N][4]
  static constexpr int EF0 = 0;                          // edge_free_code.0   N=32 K=K0
  static constexpr int EF2 = EF0 + 2 * E * K0;           // edge_free_code.2   N=32 K=32
  static constexpr int EC0 = EF2 + 2 * E * E;            // edge_code.0
  static constexpr int EC2 = EC0 + 2 * E * K0;           // edge_code.2
  static constexpr int BLK = EC2 + 2 * E * E;            // + b*kBlk: GV (N=64: G rows | Wv rows) | W1 | W2
  static constexpr int kBlk = 2 * 64 * E + 4 * E * E;
  static constexpr int oW1 = 2 * 64 * E, oW2 = oW1 + 2 * E * E;
  static constexpr int QP = BLK + 3 * kBlk;              // N=64: policy.0 edge_free cols (Q) | lin_0.0 edge_free cols (P)
  static constexpr int W5 = QP + 2 * 64 * E;             // lin_0.0 edge_code cols
  static constexpr int VEC = W5 + 2 * E * E;
  // vectors (offsets from VEC)
  static constexpr int vEF0b = 0, vEF2b = 32, vEC0b = 64, vEC2b = 96, vBLK = 128 /* +b*192: ln1g ln1b b1 b2 ln2g ln2b */,
                       vQb = vBLK + 3 * 192, vPb = vQb + 32, kVec = vPb + 32;
  static constexpr int kImage = VEC + kVec;              // floats
  static constexpr int kTab = 4 * E * kOcMax;            // floats: Mt hi | Mt lo | Vt hi | Vt lo
  static constexpr size_t kSmemBytes = (size_t)(kImage + kTab) * sizeof(float);
  // TMEM columns of a tile
  static constexpr int cXH = 0, cXL = 32, cA1 = 64, cSC = 128, cPL = 0, cPV = 224;
};

namespace tc_detail {

// old-format obstacle tables (obstacle_kernel, OT = 32 rows per tile: [Mt E x OT | V OT x E]) -> TC units.
// grid (graph, block); unit (g, blk, chunk) = [Mt_hi float[8][per][4] | Mt_lo | Vt_hi float[per/4][32][4] | Vt_lo].
__global__ void __launch_bounds__(256) obs_table_tc_kernel(const float* __restrict__ tables, int64_t table_stride,
                                                           const int32_t* __restrict__ obs_ptr,
                                                           const int32_t* __restrict__ obs_tile_ptr,
                                                           const int64_t* __restrict__ tc_tab_off, float* __restrict__ tc_tables,
                                                           int64_t tc_tab_stride) {
  constexpr int E = 32, OT = 32;
  const int g = blockIdx.x, blk = blockIdx.y;
  const int O = obs_ptr[g + 1] - obs_ptr[g];
  const int nch = tc_nchunks(O), per = tc_per(O, nch);
  const float* tab = tables + (size_t)(1 * 3 + blk) * table_stride + (size_t)obs_tile_ptr[g] * (2 * E * OT);
  float* out = tc_tables + (size_t)blk * tc_tab_stride + tc_tab_off[g];
  const int total = nch * per * E;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = i / (per * E), rem = i % (per * E);
    const int oo = rem / E, f = rem % E;      // obstacle row within the chunk, feature
    const int o = c * per + oo;
    float m = 0.f, vv = 0.f;
    if (o < O) {
      const float* tile = tab + (size_t)(o / OT) * (2 * E * OT);
      m = tile[f * OT + (o % OT)];              // M_o[k = f]
      vv = tile[E * OT + (o % OT) * E + f];     // V_o[n = f]
    }
    float* unit = out + (size_t)c * (4 * E * per);
    // Mt plane: B[n = oo][k = f]  -> [(f/4)][oo][f%4]
    const int im = ((f >> 2) * per + oo) * 4 + (f & 3);
    // Vt plane: B[n = f][k = oo]  -> [(oo/4)][f][oo%4]
    const int iv = ((oo >> 2) * E + f) * 4 + (oo & 3);
    const float mh = umma::tf32_rna(m), vh = umma::tf32_rna(vv);
    unit[im] = mh;
    unit[E * per + im] = umma::tf32_rna(m - mh);
    unit[2 * E * per + iv] = vh;
    unit[3 * E * per + iv] = umma::tf32_rna(vv - vh);
  }
}

template <int N>
__device__ __forceinline__ void ld_cols(uint32_t taddr, float* dst) {
  static_assert(N % 16 == 0, "N % 16");
#pragma unroll
  for (int c = 0; c < N; c += 32) {
    if (c + 32 <= N) umma::ld32(taddr + c, dst + c); else umma::ld16(taddr + c, dst + c);
  }
}

template <int N>
__device__ __forceinline__ void layernorm_row(float* x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  float mu = 0.f;
#pragma unroll
  for (int n = 0; n < N; ++n) mu += x[n];
  mu *= (1.0f / N);
  float var = 0.f;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const float d = x[n] - mu;
    var = fmaf(d, d, var);
  }
  var *= (1.0f / N);
  const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
  for (int n = 0; n < N; ++n) x[n] = (x[n] - mu) * rstd * gamma[n] + beta[n];
}

}  // namespace tc_detail

template <int C>
__global__ void __launch_bounds__(384, 1) edge_feature_tc_kernel(
    const float* __restrict__ tcw, const float* __restrict__ v, const int32_t* __restrict__ csr_src,
    const int32_t* __restrict__ csr_dst, const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ tile_ptr, int n_graphs,
    int n_units, const int32_t* __restrict__ obs_ptr, const int64_t* __restrict__ tc_tab_off, const float* __restrict__ tc_tables,
    int64_t tc_tab_stride, int use_obstacles, float* __restrict__ P, float* __restrict__ Q) {
  using Cf = TcCfg<C>;
  constexpr int E = 32, K0 = Cf::K0;
  extern __shared__ __align__(128) float smem_tc[];
  float* img = smem_tc;
  float* tabbuf = smem_tc + Cf::kImage;
  __shared__ uint64_t bar_ready[2], bar_done[2], bar_tabfull, bar_tabfree;
  __shared__ uint32_t tmem_slot;

  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  // ---- one-time setup: weights -> shared memory, barriers, TMEM
  {
    const float4* s4 = reinterpret_cast<const float4*>(tcw);
    float4* d4 = reinterpret_cast<float4*>(img);
    for (int i = threadIdx.x; i < Cf::kImage / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar_ready[0], 128); mbar_init(&bar_ready[1], 128);
    mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1);
    mbar_init(&bar_tabfull, 1); mbar_init(&bar_tabfree, 1);
  }
  if (warp_u == 8) umma::tmem_alloc(&tmem_slot, 512);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_slot;
  const int n_blocks = use_obstacles ? 3 : 0;

  if (warp_u >= 8) {
    // warpgroup 2 hands most of its registers to the two compute warpgroups (per SM sub-partition: 2 x 232 + 40 <= 512)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  }
  if (warp_u > 8) {
    // warps 9-11 only pad the issuer's warpgroup
  } else if (warp_u == 8) {
    // =================================================================== MMA issuer / table loader
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tm, 0);
    const uint32_t img_s = smem_u32(img), tab_s = smem_u32(tabbuf);
    uint32_t rph[2] = {0, 0}, full_ph = 0, free_ph = 0;
    // table cursor: next (unit, blk, chunk) to load
    int cu_unit = blockIdx.x, cu_blk = 0, cu_c = 0;
    auto cursor_load = [&]() {   // loads the cursor's table (skipping graphs without obstacles) and advances; uniform
      while (cu_unit < n_units) {
        int g = find_segment(tile_ptr, n_graphs, cu_unit);
        int O = obs_ptr[g + 1] - obs_ptr[g];
        g = __shfl_sync(0xffffffffu, g, 0);
        O = __shfl_sync(0xffffffffu, O, 0);
        const int nch = tc_nchunks(O), per = tc_per(O, nch);
        if (n_blocks == 0 || nch == 0) { cu_unit += gridDim.x; continue; }
        long long off = tc_tab_off[g];
        off = __shfl_sync(0xffffffffu, off, 0);
        const float* src = tc_tables + (size_t)cu_blk * tc_tab_stride + off + (size_t)cu_c * (4 * E * per);
        const uint32_t bytes = (uint32_t)(4 * E * per) * 4u;
        if (umma::elect_one()) {
          mbar_expect_tx(&bar_tabfull, bytes);
          tma_bulk_g2s(tabbuf, src, bytes, &bar_tabfull);
        }
        __syncwarp();
        if (++cu_c == nch) { cu_c = 0; if (++cu_blk == n_blocks) { cu_blk = 0; cu_unit += gridDim.x; } }
        return;
      }
    };
    cursor_load();
    auto b_hi = [&](int off) { return img_s + (uint32_t)off * 4u; };
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      int g = find_segment(tile_ptr, n_graphs, unit);
      int O = obs_ptr[g + 1] - obs_ptr[g];
      O = __shfl_sync(0xffffffffu, O, 0);
      const int nch = n_blocks ? tc_nchunks(O) : 0, per = tc_per(O, nch);
      // one stage for both tiles: wait for the tile's operands, issue, commit
#define GMP_TC_STAGE(BODY)                                                        \
  _Pragma("unroll") for (int t = 0; t < 2; ++t) {                                 \
    umma::mbar_wait_guard(&bar_ready[t], rph[t]);                                 \
    rph[t] ^= 1u;                                                                 \
    umma::fence_after_sync();                                                     \
    if (umma::elect_one()) {                                                      \
      const uint32_t tc = tm_u + (uint32_t)t * 256u;                              \
      BODY;                                                                       \
      umma::commit(&bar_done[t]);                                                 \
    }                                                                             \
    __syncwarp();                                                                 \
  }
      GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cA1, tc + Cf::cXH, tc + Cf::cXL, b_hi(Cf::EF0), b_hi(Cf::EF0 + E * K0), E, 0, E, K0, false));
      GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cA1, tc + Cf::cXH, tc + Cf::cXL, b_hi(Cf::EF2), b_hi(Cf::EF2 + E * E), E, 0, E, E, false));
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int wb = Cf::BLK + blk * Cf::kBlk;
        if (nch == 0) {
          GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cA1, tc + Cf::cXH, tc + Cf::cXL, b_hi(wb), b_hi(wb + 64 * E), 64, 0, 64, E, false));
        }
        for (int c = 0; c < nch; ++c) {
          umma::mbar_wait_guard(&bar_tabfull, full_ph);   // this (blk, chunk)'s tables have landed
          full_ph ^= 1u;
          const uint32_t mt_hi = tab_s, mt_lo = tab_s + (uint32_t)(E * per) * 4u, vt_hi = tab_s + (uint32_t)(2 * E * per) * 4u,
                         vt_lo = tab_s + (uint32_t)(3 * E * per) * 4u;
          GMP_TC_STAGE({
            if (c == 0) umma::gemm3_ts(tc + Cf::cA1, tc + Cf::cXH, tc + Cf::cXL, b_hi(wb), b_hi(wb + 64 * E), 64, 0, 64, E, false);
            umma::gemm3_ts(tc + Cf::cSC, tc + Cf::cXH, tc + Cf::cXL, mt_hi, mt_lo, per, 0, per, E, false);
          });
          GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cPV, tc + Cf::cSC, tc + Cf::cPL, vt_hi, vt_lo, E, 0, E, per, false));
          // both tiles' P.V issued: when they retire the table buffer is free -> refill with the next table
          if (umma::elect_one()) umma::commit(&bar_tabfree);
          __syncwarp();
          umma::mbar_wait_guard(&bar_tabfree, free_ph);
          free_ph ^= 1u;
          cursor_load();
        }
        GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cA1, tc + Cf::cXH, tc + Cf::cXL, b_hi(wb + Cf::oW1), b_hi(wb + Cf::oW1 + E * E), E, 0, E, E, false));
        GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cA1, tc + Cf::cXH, tc + Cf::cXL, b_hi(wb + Cf::oW2), b_hi(wb + Cf::oW2 + E * E), E, 0, E, E, false));
      }
      GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cA1, tc + Cf::cXH, tc + Cf::cXL, b_hi(Cf::QP), b_hi(Cf::QP + 64 * E), 64, 0, 64, E, false));
      GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cSC, tc + Cf::cXH, tc + Cf::cXL, b_hi(Cf::EC0), b_hi(Cf::EC0 + E * K0), E, 0, E, K0, false));
      GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cSC, tc + Cf::cXH, tc + Cf::cXL, b_hi(Cf::EC2), b_hi(Cf::EC2 + E * E), E, 0, E, E, false));
      GMP_TC_STAGE(umma::gemm3_ts(tc + Cf::cA1 + 32, tc + Cf::cXH, tc + Cf::cXL, b_hi(Cf::W5), b_hi(Cf::W5 + E * E), E, 0, E, E, true));
#undef GMP_TC_STAGE
    }
  } else {
    // =================================================================== compute warps: thread == edge row
    const int tile = warp_u >> 2;
    const int row = threadIdx.x & 127;
    const uint32_t tc = tm + ((uint32_t)((warp_u & 3) * 32) << 16) + (uint32_t)tile * 256u;
    const float* vec = img + Cf::VEC;
    uint32_t dph = 0;
    auto publish = [&]() {   // this thread's TMEM writes (and reads) of the stage are complete -> tell the issuer
      umma::wait_st();
      umma::fence_before_sync();
      umma::mbar_arrive(&bar_ready[tile]);
    };
    auto await = [&]() {
      umma::mbar_wait_guard(&bar_done[tile], dph);
      dph ^= 1u;
      umma::fence_after_sync();
    };
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int g = find_segment(tile_ptr, n_graphs, unit);
      const int slot = edge_ptr[g] + (unit - tile_ptr[g]) * 256 + tile * 128 + row;
      const bool valid = slot < edge_ptr[g + 1];
      const int O = obs_ptr[g + 1] - obs_ptr[g];
      const int nch = n_blocks ? tc_nchunks(O) : 0, per = tc_per(O, nch);
      const int s_node = valid ? csr_src[slot] : 0, d_node = valid ? csr_dst[slot] : 0;
      auto stage_inputs = [&]() {   // cat(v[src], v[dst]) zero padded to K0 -> XH / XL
        float in[K0];
#pragma unroll
        for (int k = 0; k < K0; ++k) in[k] = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          in[k] = __ldg(v + (size_t)s_node * C + k);
          in[C + k] = __ldg(v + (size_t)d_node * C + k);
        }
        umma::st_split<K0>(tc + Cf::cXH, tc + Cf::cXL, in);
      };
      float x[E];
      // ---- edge_free_code                                                  (model.py:123)
      stage_inputs();
      publish();
      await();
      tc_detail::ld_cols<E>(tc + Cf::cA1, x);
      umma::wait_ld();
#pragma unroll
      for (int n = 0; n < E; ++n) x[n] = fmaxf(x[n] + vec[Cf::vEF0b + n], 0.f);
      umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
      publish();
      await();
      tc_detail::ld_cols<E>(tc + Cf::cA1, x);
      umma::wait_ld();
#pragma unroll
      for (int n = 0; n < E; ++n) x[n] += vec[Cf::vEF2b + n];
      umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
      publish();
      // ---- three edge Blocks                                               (model.py:130, 153-218)
      for (int blk = 0; blk < n_blocks; ++blk) {
        const float* bv = vec + Cf::vBLK + blk * 192;
        float acc[E];
        float m, l = 1.0f;
        await();   // Gx | Vx (+ scores of chunk 0)
        {
          float u[E];
          tc_detail::ld_cols<E>(tc + Cf::cA1, u);
          tc_detail::ld_cols<E>(tc + Cf::cA1 + 32, acc);
          umma::wait_ld();
          float s = 0.f;
#pragma unroll
          for (int n = 0; n < E; ++n) s = fmaf(u[n], x[n], s);   // self score x^T G x   (model.py:175,177)
          m = s;
        }
        for (int c = 0; c < nch; ++c) {
          if (c > 0) {
            // the probabilities' lo plane overwrote XH / XL: restore the A operand for this chunk's scores
            umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
            publish();
            await();
          }
          const int cnt = min(per, O - c * per);
          float sc[Cf::kOcMax];
#pragma unroll
          for (int j = 0; j < Cf::kOcMax; j += 16)
            if (j < per) umma::ld16(tc + Cf::cSC + j, sc + j);
          umma::wait_ld();
          // padded table rows score 0: mask them to -inf (weight 0); only the 16-column pieces at / after `cnt`
#pragma unroll
          for (int j = 0; j < Cf::kOcMax; j += 16)
            if (j < per && j + 16 > cnt) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (j + i >= cnt) sc[j + i] = -INFINITY;
            }
          float mnew = m;
#pragma unroll
          for (int j = 0; j < Cf::kOcMax; j += 16)
            if (j < per) {
#pragma unroll
              for (int i = 0; i < 16; ++i) mnew = fmaxf(mnew, sc[j + i]);
            }
          const float corr = umma::ex2_approx(m - mnew);
          float lsum = l * corr;
#pragma unroll
          for (int j = 0; j < Cf::kOcMax; j += 16)
            if (j < per) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float p = umma::ex2_approx(sc[j + i] - mnew);
                lsum += p;
                sc[j + i] = p;
              }
              umma::st_split<16>(tc + Cf::cSC + j, tc + Cf::cPL + j, sc + j);
            }
          l = lsum;
          m = mnew;
          publish();
          await();   // P.V of this chunk
          float pv[E];
          tc_detail::ld_cols<E>(tc + Cf::cPV, pv);
          umma::wait_ld();
#pragma unroll
          for (int n = 0; n < E; ++n) acc[n] = fmaf(acc[n], corr, pv[n]);
        }
        // softmax normalisation, residual, attention.layer_norm               (model.py:181)
        {
          const float inv = 1.0f / l;
#pragma unroll
          for (int n = 0; n < E; ++n) acc[n] = fmaf(acc[n], inv, x[n]);
        }
        tc_detail::layernorm_row<E>(acc, bv, bv + 32, 1e-6f);
        umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, acc);
        publish();
        await();   // map_feed.w_1                                             (model.py:193-201)
        tc_detail::ld_cols<E>(tc + Cf::cA1, x);
        umma::wait_ld();
#pragma unroll
        for (int n = 0; n < E; ++n) x[n] = fmaxf(x[n] + bv[64 + n], 0.f);
        umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
        publish();
        await();   // map_feed.w_2
        tc_detail::ld_cols<E>(tc + Cf::cA1, x);
        umma::wait_ld();
#pragma unroll
        for (int n = 0; n < E; ++n) x[n] += bv[96 + n] + acc[n];
        tc_detail::layernorm_row<E>(x, bv + 128, bv + 160, 1e-6f);
        umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
        publish();
      }
      // ---- Q = Wc ef + b ; P_ef = W4 ef (stays in TMEM)                      (model.py:145-146 ; :39)
      await();
      tc_detail::ld_cols<E>(tc + Cf::cA1, x);
      umma::wait_ld();
      if (valid) {
        float4* q4 = reinterpret_cast<float4*>(Q + (size_t)slot * E);
#pragma unroll
        for (int n = 0; n < E; n += 4)
          q4[n >> 2] = make_float4(x[n] + vec[Cf::vQb + n], x[n + 1] + vec[Cf::vQb + n + 1], x[n + 2] + vec[Cf::vQb + n + 2],
                                   x[n + 3] + vec[Cf::vQb + n + 3]);
      }
      // ---- edge_code                                                       (model.py:120)
      stage_inputs();
      publish();
      await();
      tc_detail::ld_cols<E>(tc + Cf::cSC, x);
      umma::wait_ld();
#pragma unroll
      for (int n = 0; n < E; ++n) x[n] = fmaxf(x[n] + vec[Cf::vEC0b + n], 0.f);
      umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
      publish();
      await();
      tc_detail::ld_cols<E>(tc + Cf::cSC, x);
      umma::wait_ld();
#pragma unroll
      for (int n = 0; n < E; ++n) x[n] += vec[Cf::vEC2b + n];
      umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
      publish();
      await();   // P = W4 ef + W5 ec
      tc_detail::ld_cols<E>(tc + Cf::cA1 + 32, x);
      umma::wait_ld();
      if (valid) {
        float4* p4 = reinterpret_cast<float4*>(P + (size_t)slot * E);
#pragma unroll
        for (int n = 0; n < E; n += 4)
          p4[n >> 2] = make_float4(x[n] + vec[Cf::vPb + n], x[n + 1] + vec[Cf::vPb + n + 1], x[n + 2] + vec[Cf::vPb + n + 2],
                                   x[n + 3] + vec[Cf::vPb + n + 3]);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp_u == 8) umma::tmem_dealloc(tm, 512);
}

}  // namespace gmp
