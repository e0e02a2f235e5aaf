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
//   * shared memory holds the B operands only: all weights of this stage as hi / lo TF32 planes (resident, 139-158 KB)
//     and ONE obstacle-table buffer (<= 64 KB: scale*Wq^T Wk o and Wv o of <= 128 obstacles, hi / lo) that warp 8
//     refills with a TMA bulk copy once both tiles have published map_feed.w_1's operand (every P.V product of the block
//     has retired by then) -- the refill overlaps the two FFN stages.
//
// TMEM columns of a tile (256 of the CTA's 512): XH [0,32) XL [32,64) A operand; A1 [64,128) accumulators
// (Gx | Vx, FFN, Q | P); SC [128,224) scores -> probabilities hi (in place); PL [0,96) probabilities lo (overlays the
// dead XH/XL/Gx); PV [224,256); in the tail HH [160,192) HL [192,224) hold the hidden layer of edge_code.
//
// Per-tile program (MMA round trips): [edge_free_code.0 if 2c > 8, else plain FMAs] -> edge_free_code.2 -> 3 x { Gx|Vx +
// scores, P.V (per <=96-obstacle chunk), map_feed.w_1, map_feed.w_2 } -> [edge_code.0 if 2c > 8] -> { Q | P_ef, P += (W5
// W_ec2) hidden_ec }: edge_code.2 is folded into lin_0's edge_code columns on the host, so edge_code itself never exists.
#pragma once
#include <type_traits>
#include "handle.h"
#include "rowtile.cuh"
#include "umma.cuh"

namespace gmp {

// obstacle chunking shared by host and device: up to 96 obstacles are one sub-chunk; more are split into nch sub-chunks of
// `per` (multiple of 16, <= 64) table rows, zero padded; two consecutive sub-chunks (<= 128 obstacles) form one table load
// and share one scores stage
__host__ __device__ inline int tc_nchunks(int O) { return O <= 96 ? (O > 0) : (O + 63) / 64; }
__host__ __device__ inline int tc_per(int O, int nch) { return nch ? ((O + nch - 1) / nch + 15) / 16 * 16 : 0; }

template <int C>
struct TcCfg {
  static constexpr int E = 32;
  static constexpr int K0 = (2 * C + 7) / 8 * 8;        // encoder input width padded to the MMA K step
  static constexpr int K4 = (2 * C + 3) / 4 * 4;        // ... padded to a float4 (SIMT first layer)
  static constexpr bool kSimtIn = 2 * C <= 8;           // first encoder layers as plain FMAs (cheaper than an MMA round trip)
  static constexpr int kOcMax = 96;                      // obstacle rows of a lone sub-chunk
  static constexpr int kOcMax2 = 64;                     // ... of each sub-chunk of a pair
  // float offsets in the TC weight image; an MMA matrix is [hi plane | lo plane], a plane is float[K/4][N][4]
  static constexpr int ENC0 = 0;                         // N=64: edge_free_code.0 rows | edge_code.0 rows, K=K0
  static constexpr int ENC0F = ENC0 + 2 * 64 * K0;       // the same as plain fp32 float[64][K4]
  static constexpr int EF2 = ENC0F + 64 * K4;            // edge_free_code.2   N=32 K=32
  static constexpr int BLK = EF2 + 2 * E * E;            // + b*kBlk: GV (N=64: G rows | Wv rows) | W1 | W2
  static constexpr int kBlk = 2 * 64 * E + 4 * E * E;
  static constexpr int oW1 = 2 * 64 * E, oW2 = oW1 + 2 * E * E;
  static constexpr int QP = BLK + 3 * kBlk;              // N=64: policy.0 edge_free cols (Q) | lin_0.0 edge_free cols (P)
  static constexpr int W52 = QP + 2 * 64 * E;            // lin_0.0 edge_code cols . edge_code.2   (N=32 K=32)
  static constexpr int VEC = W52 + 2 * E * E;
  // vectors (offsets from VEC)
  static constexpr int vEF0b = 0, vEC0b = 32, vEF2b = 64, vBLK = 96 /* +b*192: ln1g ln1b b1 b2 ln2g ln2b */,
                       vQb = vBLK + 3 * 192, vPb = vQb + 32, kVec = vPb + 32;
  static constexpr int kImage = VEC + kVec;              // floats
  static constexpr int kTab = 2 * 4 * E * kOcMax2;       // floats: two sub-chunks of [Mt hi | Mt lo | Vt hi | Vt lo] (>= one of 96 rows)
  static constexpr size_t kSmemBytes = (size_t)(kImage + kTab) * sizeof(float);
  // TMEM columns of a tile
  static constexpr int cXH = 0, cXL = 32, cA1 = 64, cSC = 128, cSC1 = 192, cPL = 0;
  // P.V accumulator: lone sub-chunk -> cPV1 (PL may reach column 96); pair -> cPV2 (= dead Gx; SC1 occupies cPV1)
  static constexpr int cPV1 = 224, cPV2 = cA1;
  static constexpr int cHH = 160, cHL = 192;             // tail: hidden layer of edge_code (A operand of W52), also its inputs
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

// per 256-slot unit: (first CSR slot, end slot of its graph, obstacle count, float offset of the graph's table units)
// -- one 16-byte load per unit in the main kernel instead of a binary search plus four dependent loads
__global__ void __launch_bounds__(256) unit_meta_kernel(const int32_t* __restrict__ tile_ptr, int n_graphs, int n_units,
                                                        const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ obs_ptr,
                                                        const int64_t* __restrict__ tc_tab_off, int4* __restrict__ meta) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_units) return;
  const int g = find_segment(tile_ptr, n_graphs, u);
  meta[u] = make_int4(edge_ptr[g] + (u - tile_ptr[g]) * 256, edge_ptr[g + 1], obs_ptr[g + 1] - obs_ptr[g], (int)tc_tab_off[g]);
}

template <int N>
__device__ __forceinline__ void ld_cols(uint32_t taddr, float* dst) {
  static_assert(N % 16 == 0, "N % 16");
#pragma unroll
  for (int c = 0; c < N; c += 32) {
    if (c + 32 <= N) umma::ld32(taddr + c, dst + c); else umma::ld16(taddr + c, dst + c);
  }
}

// LayerNorm of one row held in registers (biased variance, torch.nn.LayerNorm); sums run as four independent chains so
// that a lone warp on a scheduler is not serialised on the 4-cycle FADD latency
template <int N>
__device__ __forceinline__ void layernorm_row(float* x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  static_assert(N % 4 == 0, "N % 4");
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int n = 0; n < N; n += 4) { s0 += x[n]; s1 += x[n + 1]; s2 += x[n + 2]; s3 += x[n + 3]; }
  const float mu = ((s0 + s1) + (s2 + s3)) * (1.0f / N);
  float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
  for (int n = 0; n < N; n += 4) {
    const float d0 = x[n] - mu, d1 = x[n + 1] - mu, d2 = x[n + 2] - mu, d3 = x[n + 3] - mu;
    v0 = fmaf(d0, d0, v0); v1 = fmaf(d1, d1, v1); v2 = fmaf(d2, d2, v2); v3 = fmaf(d3, d3, v3);
  }
  const float var = ((v0 + v1) + (v2 + v3)) * (1.0f / N);
  const float rstd = rsqrtf(var + eps);   // one MUFU (2 ulp) instead of the ~20-instruction IEEE sqrt + divide; tolerance 1e-4 on the logits
#pragma unroll
  for (int n = 0; n < N; ++n) x[n] = (x[n] - mu) * rstd * gamma[n] + beta[n];
}

}  // namespace tc_detail

// HALVES = 1: four warps per tile, thread == row (the round-1 organisation, 384 threads).
// HALVES = 2: EIGHT warps per tile: warps w and w + 8 own the same 32 TMEM lanes (rows) and each take half of the columns of
//             every vector (16 of the 32 features, every other 16-column piece of the scores); row-wise reductions (self score,
//             softmax maximum and sum, LayerNorm moments) are completed by exchanging one float per row through shared memory
//             under a 64-thread named barrier.  The softmax is rolled over 16-column pieces in two passes (maximum, then
//             exponentials) -- no 96-entry register array -- so a thread fits in 120 registers and FOUR epilogue warps share
//             each scheduler instead of two (round 1: issue slots 35 % busy, epilogues 81 % of a tile's time).  544 threads.
// RD = true: ONE ISSUER WARP PER TILE (requires 1 <= obstacles <= 128 in every graph and use_obstacles; 576 threads).  Each issuer
//             sleeps on its own tile's `ready` barrier and issues that tile's MMA groups only, so the two tiles are no longer
//             served stage by stage in the fixed order tile 0, tile 1 and drift apart instead of contending for the same
//             issue slots and the same tensor pipe at the same moments.  The only coupling left is the single obstacle-table
//             buffer: the second issuer to pass map_feed.w_1 of a Block loads the next table, and both wait for it to land
//             before they issue the next Block.  (A first version -- one warp polling both `ready` barriers and serving
//             whichever tile had published -- was bit-identical but slower than lockstep: every instruction of the poll loop
//             costs ~5 cycles next to four epilogue warps on the same scheduler, profiles/r2_rd_issuer.md.)
template <int C, int HALVES, bool RD = false>
__global__ void __launch_bounds__(HALVES == 1 ? 384 : (RD ? 576 : 544), 1) edge_feature_tc_kernel(
    const float* __restrict__ tcw, const float* __restrict__ v, const int32_t* __restrict__ csr_src,
    const int32_t* __restrict__ csr_dst, const int4* __restrict__ unit_meta, int n_units, const float* __restrict__ tc_tables,
    int64_t tc_tab_stride, int use_obstacles, float* __restrict__ P, float* __restrict__ Q) {
  using Cf = TcCfg<C>;
  constexpr int E = 32, K0 = Cf::K0;
  extern __shared__ __align__(128) float smem_tc[];
  float* img = smem_tc;
  float* tabbuf = smem_tc + Cf::kImage;
  __shared__ uint64_t bar_ready[2], bar_done[2], bar_tabfull;
  __shared__ uint32_t tmem_slot;
  __shared__ int tab_released;                                                 // RD: issuers that are done with the current table
  __shared__ float xch[HALVES == 2 ? 2 * 2 * 2 * 128 : 1];                    // [tile][parity][half][row]: row-reduction exchange
  constexpr int kIssuer = 8 * HALVES;                                          // warp index of the MMA issuer

  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  // ---- one-time setup: weights -> shared memory, barriers, TMEM
  {
    const float4* s4 = reinterpret_cast<const float4*>(tcw);
    float4* d4 = reinterpret_cast<float4*>(img);
    for (int i = threadIdx.x; i < Cf::kImage / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
  }
  if (threadIdx.x == 0) {
    tab_released = 0;
    mbar_init(&bar_ready[0], 128 * HALVES); mbar_init(&bar_ready[1], 128 * HALVES);
    mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1);
    mbar_init(&bar_tabfull, 1);
  }
  if (warp_u == kIssuer) umma::tmem_alloc(&tmem_slot, 512);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_slot;
  const int n_blocks = use_obstacles ? 3 : 0;

  if constexpr (HALVES == 1) {
    if (warp_u >= 8) {
      // warpgroup 2 hands most of its registers to the two compute warpgroups (per SM sub-partition: 2 x 224 + 56 <= 512)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    }
  }
  if (RD && (warp_u == kIssuer || warp_u == kIssuer + 1)) {
    // =================================================================== one MMA issuer warp PER TILE (RD)
    if constexpr (RD) {
      const int t = warp_u - kIssuer;
      const uint32_t tm_u = __shfl_sync(0xffffffffu, tm, 0);
      const uint32_t img_s = smem_u32(img), tab_s = smem_u32(tabbuf);
      const uint32_t tc = tm_u + (uint32_t)t * 256u, xh = tc + Cf::cXH, xl = tc + Cf::cXL;
      auto wd = [&](int off, int rows) { return umma::desc_lo32(img_s + (uint32_t)off * 4u, (uint32_t)rows); };
      auto raw_meta = [&](int u) { return __ldg(unit_meta + min(u, n_units - 1)); };   // clamped: no select waiting on the load
      auto load_table = [&](const int4 mm_raw, int blk) {          // table of (unit with metadata mm_raw, Block blk) -> tabbuf
        const int O = __shfl_sync(0xffffffffu, mm_raw.z, 0), off = __shfl_sync(0xffffffffu, mm_raw.w, 0);
        const int nch = tc_nchunks(O), per = tc_per(O, nch);
        const float* src = tc_tables + (size_t)blk * tc_tab_stride + (size_t)off;
        const uint32_t bytes = (uint32_t)(nch * 4 * E * per) * 4u;
        if (umma::elect_one()) {
          mbar_expect_tx(&bar_tabfull, bytes);
          tma_bulk_g2s(tabbuf, src, bytes, &bar_tabfull);
        }
        __syncwarp();
      };
      uint32_t rph = 0, full_ph = 0;
      int4 meta_cur = raw_meta(blockIdx.x), meta_next = raw_meta(blockIdx.x + gridDim.x);   // one unit ahead of its use
      if (t == 0 && (int)blockIdx.x < n_units) load_table(meta_cur, 0);
#ifdef GMP_TC_PROFILE
      long long prof_ready = 0, prof_tab = 0, prof_issue = 0, prof_t0 = clock64();
#define GMP_RD_T(VAR, ...) { const long long c0 = clock64(); __VA_ARGS__; VAR += clock64() - c0; }
#else
#define GMP_RD_T(VAR, ...) { __VA_ARGS__; }
#endif
      // one stage of THIS tile: wait for its operands, issue, commit (exactly one commit per `ready` phase)
#define GMP_RD_STAGE(...)                                                        \
  {                                                                              \
    GMP_RD_T(prof_ready, umma::mbar_wait_guard(&bar_ready[t], rph));             \
    rph ^= 1u;                                                                   \
    umma::fence_after_sync();                                                    \
    GMP_RD_T(prof_issue,                                                         \
      if (umma::elect_one()) {                                                   \
        __VA_ARGS__;                                                             \
        umma::commit(&bar_done[t]);                                              \
      }                                                                          \
      __syncwarp());                                                             \
  }
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int O = __shfl_sync(0xffffffffu, meta_cur.z, 0);
        const int nch = tc_nchunks(O), per = tc_per(O, nch);
        const uint32_t sub = (uint32_t)(4 * E * per) * 4u, pl = (uint32_t)(E * per) * 4u;   // bytes per sub-chunk / per plane
        const uint32_t mt_h0 = umma::desc_lo32(tab_s, (uint32_t)per), mt_l0 = umma::desc_lo32(tab_s + pl, (uint32_t)per),
                       vt_h0 = umma::desc_lo32(tab_s + 2 * pl, E), vt_l0 = umma::desc_lo32(tab_s + 3 * pl, E);
        const uint32_t mt_h1 = umma::desc_lo32(tab_s + sub, (uint32_t)per), mt_l1 = umma::desc_lo32(tab_s + sub + pl, (uint32_t)per),
                       vt_h1 = umma::desc_lo32(tab_s + sub + 2 * pl, E), vt_l1 = umma::desc_lo32(tab_s + sub + 3 * pl, E);
        if constexpr (!Cf::kSimtIn) {
          GMP_RD_STAGE(umma::gemm3_fixed<E, K0, 64>(tc + Cf::cA1, xh, xl, wd(Cf::ENC0, 64), wd(Cf::ENC0 + 64 * K0, 64), false));
        }
        GMP_RD_STAGE(umma::gemm3_fixed<E, E, E>(tc + Cf::cA1, xh, xl, wd(Cf::EF2, E), wd(Cf::EF2 + E * E, E), false));
        for (int blk = 0; blk < 3; ++blk) {
          const int wb = Cf::BLK + blk * Cf::kBlk;
          const uint32_t gv_h = wd(wb, 64), gv_l = wd(wb + 64 * E, 64);
          GMP_RD_T(prof_tab, umma::mbar_wait_guard(&bar_tabfull, full_ph));   // this Block's table has landed (both issuers watch it)
          full_ph ^= 1u;
          GMP_RD_STAGE({
            umma::gemm3_fixed<64, E, 64>(tc + Cf::cA1, xh, xl, gv_h, gv_l, false);
            umma::gemm3_n<E>(tc + Cf::cSC, xh, xl, mt_h0, mt_l0, per);
            if (nch == 2) umma::gemm3_n<E>(tc + Cf::cSC1, xh, xl, mt_h1, mt_l1, per);
          });
          GMP_RD_STAGE(umma::gemm3_k<E, E>(tc + (nch == 2 ? Cf::cPV2 : Cf::cPV1), tc + Cf::cSC, tc + Cf::cPL, vt_h0, vt_l0, per));
          if (nch == 2) {
            GMP_RD_STAGE(umma::gemm3_k<E, E>(tc + Cf::cPV2, tc + Cf::cSC1, tc + Cf::cPL, vt_h1, vt_l1, per));
          }
          GMP_RD_STAGE(umma::gemm3_fixed<E, E, E>(tc + Cf::cA1, xh, xl, wd(wb + Cf::oW1, E), wd(wb + Cf::oW1 + E * E, E), false));
          // this tile has published map_feed.w_1's operand, so its P.V products of the Block have retired.  The SECOND issuer to
          // get here refills the table buffer (next Block / next unit) behind the FFN stages; nobody blocks
          {
            int prev = 0;
            if ((threadIdx.x & 31) == 0) {
              __threadfence_block();
              prev = atomicAdd(&tab_released, 1);
              __threadfence_block();
            }
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev & 1) {
              if (blk < 2) load_table(meta_cur, blk + 1);
              else if (unit + (int)gridDim.x < n_units) load_table(meta_next, 0);
            }
          }
          GMP_RD_STAGE(umma::gemm3_fixed<E, E, E>(tc + Cf::cA1, xh, xl, wd(wb + Cf::oW2, E), wd(wb + Cf::oW2 + E * E, E), false));
        }
        if constexpr (!Cf::kSimtIn) {
          GMP_RD_STAGE(umma::gemm3_fixed<E, K0, 64>(tc + Cf::cSC, tc + Cf::cHH, tc + Cf::cHL, wd(Cf::ENC0, 64) + E,
                                                    wd(Cf::ENC0 + 64 * K0, 64) + E, false));
        }
        GMP_RD_STAGE({
          umma::gemm3_fixed<64, E, 64>(tc + Cf::cA1, xh, xl, wd(Cf::QP, 64), wd(Cf::QP + 64 * E, 64), false);
          umma::gemm3_fixed<E, E, E>(tc + Cf::cA1 + 32, tc + Cf::cHH, tc + Cf::cHL, wd(Cf::W52, E), wd(Cf::W52 + E * E, E), true);
        });
        meta_cur = meta_next;
        meta_next = raw_meta(unit + 2 * (int)gridDim.x);
      }
#undef GMP_RD_STAGE
#undef GMP_RD_T
#ifdef GMP_TC_PROFILE
      if (blockIdx.x == 0 && (threadIdx.x & 31) == 0)
        printf("tc rd issuer %d: total %lld cyc; waiting ready %lld tables %lld; issuing %lld\n", t, clock64() - prof_t0, prof_ready, prof_tab,
               prof_issue);
#endif
    }
  } else if (warp_u > kIssuer) {
    // warps 9-11 only pad the issuer's warpgroup (HALVES == 1)
  } else if (warp_u == kIssuer) {
    // =================================================================== MMA issuer / table loader
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tm, 0);
    const uint32_t img_s = smem_u32(img), tab_s = smem_u32(tabbuf);
    uint32_t rph[2] = {0, 0}, full_ph = 0;
    // unit metadata, one unit ahead of its use (uniform: every lane loads the same 16 bytes)
    // (raw, index clamped: the broadcast that makes the fields provably uniform happens where they are consumed, a unit later,
    //  so the issuer never waits on the load itself)
    auto load_meta = [&](int u) { return __ldg(unit_meta + min(u, n_units - 1)); };
    int4 meta_cur = load_meta(blockIdx.x), meta_next = load_meta(blockIdx.x + gridDim.x);
    int cur_unit = blockIdx.x;
    // table cursor: next (unit, blk, pair of sub-chunks) to load.  tab_busy: the buffer holds (or is receiving) a table
    // whose P.V products have not all retired yet
    int cu_unit = blockIdx.x, cu_blk = 0, cu_s = 0;
    bool tab_busy = false;
    auto cursor_load = [&]() {   // loads the cursor's table (skipping graphs without obstacles) and advances; uniform
      if (tab_busy) return;
      while (cu_unit < n_units) {
        const int4 mm = cu_unit == cur_unit ? meta_cur : (cu_unit == cur_unit + (int)gridDim.x ? meta_next : load_meta(cu_unit));
        const int O = __shfl_sync(0xffffffffu, mm.z, 0), tab_off = __shfl_sync(0xffffffffu, mm.w, 0);
        const int nch = tc_nchunks(O), per = tc_per(O, nch);
        if (n_blocks == 0 || nch == 0) { cu_unit += gridDim.x; continue; }
        const int ns = min(2, nch - 2 * cu_s);
        const float* src = tc_tables + (size_t)cu_blk * tc_tab_stride + (size_t)tab_off + (size_t)(2 * cu_s) * (4 * E * per);
        const uint32_t bytes = (uint32_t)(ns * 4 * E * per) * 4u;
        if (umma::elect_one()) {
          mbar_expect_tx(&bar_tabfull, bytes);
          tma_bulk_g2s(tabbuf, src, bytes, &bar_tabfull);
        }
        __syncwarp();
        tab_busy = true;
        if (2 * (++cu_s) >= nch) { cu_s = 0; if (++cu_blk == n_blocks) { cu_blk = 0; cu_unit += gridDim.x; } }
        return;
      }
    };
    cursor_load();
#ifdef GMP_TC_PROFILE
    long long prof_ready[2] = {0, 0}, prof_tab = 0, prof_t0 = clock64();
    long long prof_issue[5] = {0, 0, 0, 0, 0};   // K0-wide encoder layer, 32x32, GV (+scores), P.V, Q|P
#endif
    // lo words of the weight descriptors: hi / lo plane of the matrix at float offset `off` with `rows` rows and K columns
    auto wd = [&](int off, int rows) { return umma::desc_lo32(img_s + (uint32_t)off * 4u, (uint32_t)rows); };
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      if (unit != cur_unit) {
        cur_unit = unit;
        meta_cur = meta_next;
        meta_next = load_meta(unit + gridDim.x);
      }
      const int O = __shfl_sync(0xffffffffu, meta_cur.z, 0);
      const int nch = n_blocks ? tc_nchunks(O) : 0, per = tc_per(O, nch);
      // one stage for both tiles: wait for the tile's operands, issue, commit.  Exactly ONE commit per `ready` phase: a
      // `done` phase must be observed by all 128 waiters before the next one can complete (mbarrier waits are by parity; a
      // second commit without a publish in between could complete two phases behind a slow warp's back and hang it)
#ifdef GMP_TC_PROFILE
#define GMP_TC_T0 const long long w0 = clock64();
#define GMP_TC_T1 prof_ready[t] += clock64() - w0;
#define GMP_TC_I0 const long long i0 = clock64();
#define GMP_TC_I1(KIND) prof_issue[KIND] += clock64() - i0;
#else
#define GMP_TC_T0
#define GMP_TC_T1
#define GMP_TC_I0
#define GMP_TC_I1(KIND)
#endif
#define GMP_TC_STAGE(KIND, ...)                                                   \
  _Pragma("unroll") for (int t = 0; t < 2; ++t) {                                 \
    GMP_TC_T0                                                                     \
    umma::mbar_wait_guard(&bar_ready[t], rph[t]);                                 \
    GMP_TC_T1                                                                     \
    rph[t] ^= 1u;                                                                 \
    umma::fence_after_sync();                                                     \
    if (umma::elect_one()) {                                                      \
      const uint32_t tc = tm_u + (uint32_t)t * 256u;                              \
      const uint32_t xh = tc + Cf::cXH, xl = tc + Cf::cXL;                        \
      GMP_TC_I0                                                                   \
      __VA_ARGS__;                                                                \
      umma::commit(&bar_done[t]);                                                 \
      GMP_TC_I1(KIND)                                                             \
    }                                                                             \
    __syncwarp();                                                                 \
  }
      if constexpr (!Cf::kSimtIn) {   // hidden layer of edge_free_code on the tensor cores (rows 0..31 of the stacked ENC0)
        GMP_TC_STAGE(0, (umma::gemm3_fixed<E, K0, 64>(tc + Cf::cA1, xh, xl, wd(Cf::ENC0, 64), wd(Cf::ENC0 + 64 * K0, 64), false)));
      }
      GMP_TC_STAGE(1, (umma::gemm3_fixed<E, E, E>(tc + Cf::cA1, xh, xl, wd(Cf::EF2, E), wd(Cf::EF2 + E * E, E), false)));
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int wb = Cf::BLK + blk * Cf::kBlk;
        const uint32_t gv_h = wd(wb, 64), gv_l = wd(wb + 64 * E, 64);
        if (nch == 0) {
          GMP_TC_STAGE(2, (umma::gemm3_fixed<64, E, 64>(tc + Cf::cA1, xh, xl, gv_h, gv_l, false)));
        }
        for (int sc = 0; 2 * sc < nch; ++sc) {
          const int ns = min(2, nch - 2 * sc);
          if (sc > 0) {
            // more than 128 obstacles (rare): both tiles have retired the previous table's products once they publish the
            // restored A operand; only then may the buffer be refilled
            umma::mbar_wait_guard(&bar_ready[0], rph[0]); rph[0] ^= 1u;
            umma::mbar_wait_guard(&bar_ready[1], rph[1]); rph[1] ^= 1u;
            tab_busy = false;
            cursor_load();
          }
#ifdef GMP_TC_PROFILE
          const long long wt0 = clock64();
#endif
          umma::mbar_wait_guard(&bar_tabfull, full_ph);   // this table has landed
          full_ph ^= 1u;
#ifdef GMP_TC_PROFILE
          prof_tab += clock64() - wt0;
#endif
          const uint32_t sub = (uint32_t)(4 * E * per) * 4u, pl = (uint32_t)(E * per) * 4u;   // bytes per sub-chunk / per plane
          const uint32_t mt_h0 = umma::desc_lo32(tab_s, (uint32_t)per), mt_l0 = umma::desc_lo32(tab_s + pl, (uint32_t)per),
                         vt_h0 = umma::desc_lo32(tab_s + 2 * pl, E), vt_l0 = umma::desc_lo32(tab_s + 3 * pl, E);
          const uint32_t mt_h1 = umma::desc_lo32(tab_s + sub, (uint32_t)per), mt_l1 = umma::desc_lo32(tab_s + sub + pl, (uint32_t)per),
                         vt_h1 = umma::desc_lo32(tab_s + sub + 2 * pl, E), vt_l1 = umma::desc_lo32(tab_s + sub + 3 * pl, E);
          if (sc == 0) {
            GMP_TC_STAGE(2, {
              umma::gemm3_fixed<64, E, 64>(tc + Cf::cA1, xh, xl, gv_h, gv_l, false);
              umma::gemm3_n<E>(tc + Cf::cSC, xh, xl, mt_h0, mt_l0, per);
              if (ns == 2) umma::gemm3_n<E>(tc + Cf::cSC1, xh, xl, mt_h1, mt_l1, per);
            });
          } else {
#pragma unroll
            for (int t = 0; t < 2; ++t) {   // (the tiles' `ready` phases were consumed above)
              umma::fence_after_sync();
              if (umma::elect_one()) {
                const uint32_t tc = tm_u + (uint32_t)t * 256u;
                umma::gemm3_n<E>(tc + Cf::cSC, tc + Cf::cXH, tc + Cf::cXL, mt_h0, mt_l0, per);
                if (ns == 2) umma::gemm3_n<E>(tc + Cf::cSC1, tc + Cf::cXH, tc + Cf::cXL, mt_h1, mt_l1, per);
                umma::commit(&bar_done[t]);
              }
              __syncwarp();
            }
          }
          GMP_TC_STAGE(3, (umma::gemm3_k<E, E>(tc + (ns == 2 ? Cf::cPV2 : Cf::cPV1), tc + Cf::cSC, tc + Cf::cPL, vt_h0, vt_l0, per)));
          if (ns == 2) {
            GMP_TC_STAGE(3, (umma::gemm3_k<E, E>(tc + Cf::cPV2, tc + Cf::cSC1, tc + Cf::cPL, vt_h1, vt_l1, per)));
          }
        }
        GMP_TC_STAGE(1, (umma::gemm3_fixed<E, E, E>(tc + Cf::cA1, xh, xl, wd(wb + Cf::oW1, E), wd(wb + Cf::oW1 + E * E, E), false)));
        // both tiles have published map_feed.w_1's operand, so every P.V of this block has retired: refill the table buffer
        // (next block / next unit) behind the two FFN stages
        if (nch > 0) tab_busy = false;
        cursor_load();
        GMP_TC_STAGE(1, (umma::gemm3_fixed<E, E, E>(tc + Cf::cA1, xh, xl, wd(wb + Cf::oW2, E), wd(wb + Cf::oW2 + E * E, E), false)));
      }
      if constexpr (!Cf::kSimtIn) {   // hidden layer of edge_code: rows 32..63 of ENC0, inputs staged at cHH / cHL
        GMP_TC_STAGE(0, (umma::gemm3_fixed<E, K0, 64>(tc + Cf::cSC, tc + Cf::cHH, tc + Cf::cHL, wd(Cf::ENC0, 64) + E,
                                                      wd(Cf::ENC0 + 64 * K0, 64) + E, false)));
      }
      // Q | P_ef = [Wc ; W4] ef, then P += (W5 W_ec2) hidden_ec
      GMP_TC_STAGE(4, {
        umma::gemm3_fixed<64, E, 64>(tc + Cf::cA1, xh, xl, wd(Cf::QP, 64), wd(Cf::QP + 64 * E, 64), false);
        umma::gemm3_fixed<E, E, E>(tc + Cf::cA1 + 32, tc + Cf::cHH, tc + Cf::cHL, wd(Cf::W52, E), wd(Cf::W52 + E * E, E), true);
      });
#undef GMP_TC_STAGE
    }
#ifdef GMP_TC_PROFILE
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0)
      printf("tc profile: issuer total %lld cyc; waiting ready[0] %lld ready[1] %lld tables %lld; issuing: enc0 %lld 32x32 %lld "
             "GV+scores %lld PV %lld QP %lld\n", clock64() - prof_t0, prof_ready[0], prof_ready[1], prof_tab, prof_issue[0],
             prof_issue[1], prof_issue[2], prof_issue[3], prof_issue[4]);
#endif
  } else if constexpr (HALVES == 1) {
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
#ifdef GMP_TC_PROFILE
    long long prof_wait = 0, prof_t0 = clock64();
#endif
#ifdef GMP_TC_PROFILE
    long long prof_w[6] = {0, 0, 0, 0, 0, 0}, prof_e[6] = {0, 0, 0, 0, 0, 0}, prof_last = clock64();
    int prof_kind = 5;
#endif
    auto await = [&](int kind = 5) {   // kind (profiling only): 0 Gx|Vx+scores, 1 P.V, 2 w_1, 3 w_2, 4 tail, 5 encoders / other
#ifdef GMP_TC_PROFILE
      const long long w0 = clock64();
      prof_e[prof_kind] += w0 - prof_last;   // work since the previous await belongs to that stage's epilogue
#endif
      umma::mbar_wait_guard(&bar_done[tile], dph);
      dph ^= 1u;
      umma::fence_after_sync();
#ifdef GMP_TC_PROFILE
      prof_last = clock64();
      prof_wait += prof_last - w0;
      prof_w[kind] += prof_last - w0;
      prof_kind = kind;
#endif
    };
    // two units of metadata and one unit of endpoints are kept in flight, so that no load of the chain
    // meta -> CSR endpoints -> node coordinates ever waits on the one before it
    // (index clamped instead of a select on the loaded value: a select would wait ~700 cycles for a load whose result is
    //  needed two units later; the metadata of a unit past the end is never used -- the unit loop ends first)
    auto load_meta = [&](int u) { return __ldg(unit_meta + min(u, n_units - 1)); };
    int4 meta_nx = load_meta(blockIdx.x), meta_nx2 = load_meta(blockIdx.x + gridDim.x);
    int s_nx = 0, d_nx = 0;
    auto load_endpoints = [&](const int4 mm) {
      const int sl = mm.x + tile * 128 + row;
      const bool ok = sl < mm.y;
      s_nx = ok ? __ldg(csr_src + sl) : 0;
      d_nx = ok ? __ldg(csr_dst + sl) : 0;
    };
    load_endpoints(meta_nx);
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int4 meta = meta_nx;
      const int slot = meta.x + tile * 128 + row;
      const bool valid = slot < meta.y;
      const int O = meta.z;
      const int nch = n_blocks ? tc_nchunks(O) : 0, per = tc_per(O, nch);
      const int s_node = s_nx, d_node = d_nx;
      meta_nx = meta_nx2;                                  // next unit: its metadata arrived a unit ago ...
      load_endpoints(meta_nx);                             // ... so its endpoints can be requested right away
      meta_nx2 = load_meta(unit + 2 * gridDim.x);
      auto gather = [&](float* in) {   // cat(v[src], v[dst]), zero padded to K0
#pragma unroll
        for (int k = 0; k < K0; ++k) in[k] = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          in[k] = __ldg(v + (size_t)s_node * C + k);
          in[C + k] = __ldg(v + (size_t)d_node * C + k);
        }
      };
      // hidden = relu(W in + b) of encoder `which` (0 edge_free_code.0, 1 edge_code.0) as plain FMAs (2C <= 8)
      auto simt_hidden = [&](const float* in, int which, float* h) {
        const float* Wf = img + Cf::ENC0F + which * 32 * Cf::K4;
        const float* bb = vec + (which ? Cf::vEC0b : Cf::vEF0b);
#pragma unroll
        for (int n = 0; n < E; ++n) {
          float a = bb[n];
#pragma unroll
          for (int k4 = 0; k4 < Cf::K4; k4 += 4) {
            const float4 w = *reinterpret_cast<const float4*>(Wf + n * Cf::K4 + k4);
            a = fmaf(w.x, in[k4], a); a = fmaf(w.y, in[k4 + 1], a); a = fmaf(w.z, in[k4 + 2], a); a = fmaf(w.w, in[k4 + 3], a);
          }
          h[n] = fmaxf(a, 0.f);
        }
      };
      // A operand of the tail: hidden layer of edge_code (SIMT) or its inputs (tensor-core first layer) -> cHH / cHL
      auto stage_edge_code = [&]() {
        float in[K0];
        gather(in);
        if constexpr (Cf::kSimtIn) {
          float h[E];
          simt_hidden(in, 1, h);
          umma::st_split<E>(tc + Cf::cHH, tc + Cf::cHL, h);
        } else {
          umma::st_split<K0>(tc + Cf::cHH, tc + Cf::cHL, in);
        }
      };
      float x[E];
      // ---- edge_free_code                                                  (model.py:123)
      {
        float in[K0];
        gather(in);
        if constexpr (Cf::kSimtIn) {
          simt_hidden(in, 0, x);
        } else {
          umma::st_split<K0>(tc + Cf::cXH, tc + Cf::cXL, in);
          publish();
          await();
          tc_detail::ld_cols<E>(tc + Cf::cA1, x);
          umma::wait_ld();
#pragma unroll
          for (int n = 0; n < E; ++n) x[n] = fmaxf(x[n] + vec[Cf::vEF0b + n], 0.f);
        }
      }
      umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
      publish();
      await();
      tc_detail::ld_cols<E>(tc + Cf::cA1, x);
      umma::wait_ld();
#pragma unroll
      for (int n = 0; n < E; ++n) x[n] += vec[Cf::vEF2b + n];
      umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
      if (n_blocks == 0) stage_edge_code();
      publish();
      // ---- three edge Blocks                                               (model.py:130, 153-218)
      for (int blk = 0; blk < n_blocks; ++blk) {
        const float* bv = vec + Cf::vBLK + blk * 192;
        float acc[E];
        float m, l = 1.0f;
        await(0);   // Gx | Vx (+ scores of the first one or two sub-chunks)
        {
          float u[E];
          tc_detail::ld_cols<E>(tc + Cf::cA1, u);
          tc_detail::ld_cols<E>(tc + Cf::cA1 + 32, acc);   // value of the row itself, weight exp2(s_self - m) = 1
          umma::wait_ld();
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;               // self score x^T G x   (model.py:175,177)
#pragma unroll
          for (int n = 0; n < E; n += 4) {
            s0 = fmaf(u[n], x[n], s0); s1 = fmaf(u[n + 1], x[n + 1], s1); s2 = fmaf(u[n + 2], x[n + 2], s2); s3 = fmaf(u[n + 3], x[n + 3], s3);
          }
          m = (s0 + s1) + (s2 + s3);
        }
        // online softmax over one sub-chunk: scores at TMEM column `col` -> probabilities (hi plane in place, lo plane at
        // cPL); returns the factor by which everything accumulated so far must be rescaled
        auto softmax_sub = [&](uint32_t col, int cnt) -> float {
          constexpr int W = Cf::kOcMax;
          float sc[W];
#pragma unroll
          for (int j = 0; j < W; j += 16)
            if (j < per) umma::ld16(tc + col + j, sc + j);
          umma::wait_ld();
          // padded table rows score 0: mask them to -inf (weight 0); only the 16-column pieces at / after `cnt`
#pragma unroll
          for (int j = 0; j < W; j += 16)
            if (j < per && j + 16 > cnt) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (j + i >= cnt) sc[j + i] = -INFINITY;
            }
          float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;   // four independent chains (ILP)
#pragma unroll
          for (int j = 0; j < W; j += 16)
            if (j < per) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                m0 = fmaxf(m0, sc[j + i]); m1 = fmaxf(m1, sc[j + i + 1]); m2 = fmaxf(m2, sc[j + i + 2]); m3 = fmaxf(m3, sc[j + i + 3]);
              }
            }
          const float mnew = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          const float corr = umma::ex2_approx(m - mnew);
          float l0 = l * corr, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
          for (int j = 0; j < W; j += 16)
            if (j < per) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float p0 = umma::ex2_approx(sc[j + i] - mnew), p1 = umma::ex2_approx(sc[j + i + 1] - mnew),
                            p2 = umma::ex2_approx(sc[j + i + 2] - mnew), p3 = umma::ex2_approx(sc[j + i + 3] - mnew);
                l0 += p0; l1 += p1; l2 += p2; l3 += p3;
                sc[j + i] = p0; sc[j + i + 1] = p1; sc[j + i + 2] = p2; sc[j + i + 3] = p3;
              }
              umma::st_split<16>(tc + col + j, tc + Cf::cPL + j, sc + j);
            }
          l = (l0 + l1) + (l2 + l3);
          m = mnew;
          return corr;
        };
        for (int s2 = 0; 2 * s2 < nch; ++s2) {
          const int ns = min(2, nch - 2 * s2);
          if (s2 > 0) {
            // (> 128 obstacles) the probabilities' lo plane overwrote XH / XL: restore the A operand for these scores
            umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
            publish();
            await();
          }
          for (int sub = 0; sub < ns; ++sub) {
            const float corr = softmax_sub(sub ? Cf::cSC1 : Cf::cSC, max(0, min(per, O - (2 * s2 + sub) * per)));
            publish();   // -> P.V of this sub-chunk
            await(1);
            float pv[E];
            tc_detail::ld_cols<E>(tc + (ns == 2 ? Cf::cPV2 : Cf::cPV1), pv);
            umma::wait_ld();
#pragma unroll
            for (int n = 0; n < E; ++n) acc[n] = fmaf(acc[n], corr, pv[n]);
          }
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
        await(2);   // map_feed.w_1                                             (model.py:193-201)
        tc_detail::ld_cols<E>(tc + Cf::cA1, x);
        umma::wait_ld();
#pragma unroll
        for (int n = 0; n < E; ++n) x[n] = fmaxf(x[n] + bv[64 + n], 0.f);
        umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
        publish();
        await(3);   // map_feed.w_2
        tc_detail::ld_cols<E>(tc + Cf::cA1, x);
        umma::wait_ld();
#pragma unroll
        for (int n = 0; n < E; ++n) x[n] += bv[96 + n] + acc[n];
        tc_detail::layernorm_row<E>(x, bv + 128, bv + 160, 1e-6f);
        umma::st_split<E>(tc + Cf::cXH, tc + Cf::cXL, x);
        if (blk == n_blocks - 1) stage_edge_code();
        publish();
      }
      // ---- hidden layer of edge_code on the tensor cores (wide inputs)     (model.py:120)
      if constexpr (!Cf::kSimtIn) {
        await();
        float h[E];
        tc_detail::ld_cols<E>(tc + Cf::cSC, h);
        umma::wait_ld();
#pragma unroll
        for (int n = 0; n < E; ++n) h[n] = fmaxf(h[n] + vec[Cf::vEC0b + n], 0.f);
        umma::st_split<E>(tc + Cf::cHH, tc + Cf::cHL, h);
        publish();
      }
      // ---- Q = Wc ef + b (model.py:145-146);  P = W4 ef + W5 edge_code + b (model.py:39), edge_code.2 folded into W5
      await(4);
      float q[E];
      tc_detail::ld_cols<E>(tc + Cf::cA1, q);
      tc_detail::ld_cols<E>(tc + Cf::cA1 + 32, x);
      umma::wait_ld();
      if (valid) {
        float* qr = Q + (size_t)slot * E;
        float* pr = P + (size_t)slot * E;
        const float* qb = vec + Cf::vQb;
        const float* pb = vec + Cf::vPb;
#pragma unroll
        for (int n = 0; n < E; n += 8) {
          umma::st_global_v8(qr + n, q[n] + qb[n], q[n + 1] + qb[n + 1], q[n + 2] + qb[n + 2], q[n + 3] + qb[n + 3], q[n + 4] + qb[n + 4],
                             q[n + 5] + qb[n + 5], q[n + 6] + qb[n + 6], q[n + 7] + qb[n + 7]);
          umma::st_global_v8(pr + n, x[n] + pb[n], x[n + 1] + pb[n + 1], x[n + 2] + pb[n + 2], x[n + 3] + pb[n + 3], x[n + 4] + pb[n + 4],
                             x[n + 5] + pb[n + 5], x[n + 6] + pb[n + 6], x[n + 7] + pb[n + 7]);
        }
      }
    }
#ifdef GMP_TC_PROFILE
    if (blockIdx.x == 0 && row == 0)
      printf("tc profile: tile %d total %lld cyc, waiting for MMAs %lld cyc (%d units); wait/epilogue per stage kind: A %lld/%lld PV %lld/%lld "
             "w1 %lld/%lld w2 %lld/%lld tail %lld/%lld enc %lld/%lld\n", tile, clock64() - prof_t0, prof_wait,
             (n_units + gridDim.x - 1) / gridDim.x, prof_w[0], prof_e[0], prof_w[1], prof_e[1], prof_w[2], prof_e[2], prof_w[3], prof_e[3],
             prof_w[4], prof_e[4], prof_w[5], prof_e[5]);
#endif
  } else {
    // =================================================================== compute warps, column-split (HALVES == 2)
    const int qd = warp_u & 3, tile = (warp_u >> 2) & 1, half = warp_u >> 3;
    const int row = qd * 32 + (int)(threadIdx.x & 31);
    const uint32_t tc = tm + ((uint32_t)(qd * 32) << 16) + (uint32_t)tile * 256u;
    const float* vec = img + Cf::VEC;
    constexpr int H = E / 2;
    const int c0 = half * H;                               // this thread's 16 of the 32 features
    uint32_t dph = 0;
    int par = 0;
    float* xme = xch + tile * 512 + half * 128 + row;      // + par * 256;  the partner's slot is at (half ^ 1)
    const float* xot = xch + tile * 512 + (half ^ 1) * 128 + row;
    const int bar_id = 1 + tile * 4 + qd;                  // named barrier of this (tile, lane quarter): the two warps that share rows
#ifdef GMP_TC_PROFILE
    long long* prof_xp = nullptr;
    auto pair_sync = [&]() { const long long t0 = clock64(); asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); if (prof_xp) *prof_xp += clock64() - t0; };
#else
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };
#endif
    // complete a row reduction: returns (half 0 part) + (half 1 part) -- the same expression in both warps
    auto xsum = [&](float mine) -> float {
      xme[par * 256] = mine;
      pair_sync();
      const float other = xot[par * 256];
      par ^= 1;
      return half == 0 ? mine + other : other + mine;
    };
    auto xmax = [&](float mine) -> float {
      xme[par * 256] = mine;
      pair_sync();
      const float other = xot[par * 256];
      par ^= 1;
      return fmaxf(mine, other);
    };
    auto publish = [&]() {
      umma::wait_st();
      umma::fence_before_sync();
      umma::mbar_arrive(&bar_ready[tile]);
    };
#ifdef GMP_TC_PROFILE
    long long prof_w[6] = {0, 0, 0, 0, 0, 0}, prof_e[6] = {0, 0, 0, 0, 0, 0}, prof_x = 0, prof_last = clock64(), prof_t0 = prof_last;
    long long prof_q[4] = {0, 0, 0, 0};   // w_1 epilogue split: tcgen05.ld + wait | math + tcgen05.st issue | wait::st | fence + arrive
    int prof_kind = 5;
#endif
    auto await = [&](int kind = 5) {   // kind (profiling only): 0 Gx|Vx+scores, 1 P.V, 2 w_1, 3 w_2, 4 tail, 5 encoders / other
#ifdef GMP_TC_PROFILE
      const long long w0 = clock64();
      prof_e[prof_kind] += w0 - prof_last;   // work since the previous await belongs to that stage's epilogue
#endif
      umma::mbar_wait_guard(&bar_done[tile], dph);
      dph ^= 1u;
      umma::fence_after_sync();
#ifdef GMP_TC_PROFILE
      prof_last = clock64();
      prof_w[kind] += prof_last - w0;
      prof_kind = kind;
#endif
    };
    // LayerNorm of a row whose 32 features are split over the two warps (biased variance, two passes)
    auto layernorm_half = [&](float* y, const float* gamma, const float* beta) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int n = 0; n < H; n += 4) { s0 += y[n]; s1 += y[n + 1]; s2 += y[n + 2]; s3 += y[n + 3]; }
      const float mu = xsum((s0 + s1) + (s2 + s3)) * (1.0f / E);
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
      for (int n = 0; n < H; n += 4) {
        const float d0 = y[n] - mu, d1 = y[n + 1] - mu, d2 = y[n + 2] - mu, d3 = y[n + 3] - mu;
        v0 = fmaf(d0, d0, v0); v1 = fmaf(d1, d1, v1); v2 = fmaf(d2, d2, v2); v3 = fmaf(d3, d3, v3);
      }
      const float var = xsum((v0 + v1) + (v2 + v3)) * (1.0f / E);
      const float rstd = rsqrtf(var + 1e-6f);
#pragma unroll
      for (int n = 0; n < H; ++n) y[n] = (y[n] - mu) * rstd * gamma[c0 + n] + beta[c0 + n];
    };
#ifdef GMP_TC_PROFILE
    prof_xp = &prof_x;
#endif
    // (index clamped instead of a select on the loaded value: a select would wait ~700 cycles for a load whose result is
    //  needed two units later; the metadata of a unit past the end is never used -- the unit loop ends first)
    auto load_meta = [&](int u) { return __ldg(unit_meta + min(u, n_units - 1)); };
    int4 meta_nx = load_meta(blockIdx.x), meta_nx2 = load_meta(blockIdx.x + gridDim.x);
    int s_nx = 0, d_nx = 0;
    auto load_endpoints = [&](const int4 mm) {   // (slot clamped into the graph instead of a select on the loaded ids; rows past the end are never stored)
      const int sl = min(mm.x + tile * 128 + row, mm.y - 1);
      s_nx = __ldg(csr_src + sl);
      d_nx = __ldg(csr_dst + sl);
    };
    load_endpoints(meta_nx);
    // narrow inputs: the endpoints' coordinates of the NEXT unit are requested before the tail of the current one, so that the
    // unit starts with its inputs in registers instead of a dependent gather (ids -> coordinates, ~2 L2 latencies)
    float in_nx[Cf::kSimtIn ? 2 * C : 1];
    auto gather_next = [&]() {
      if constexpr (Cf::kSimtIn) {
#pragma unroll
        for (int k = 0; k < C; ++k) {
          in_nx[k] = __ldg(v + (size_t)s_nx * C + k);
          in_nx[C + k] = __ldg(v + (size_t)d_nx * C + k);
        }
      }
    };
    gather_next();
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int4 meta = meta_nx;
      const int slot = meta.x + tile * 128 + row;
      const bool valid = slot < meta.y;
      const int O = meta.z;
      const int nch = n_blocks ? tc_nchunks(O) : 0, per = tc_per(O, nch);
      const int s_node = s_nx, d_node = d_nx;
      meta_nx = meta_nx2;
      load_endpoints(meta_nx);
      meta_nx2 = load_meta(unit + 2 * gridDim.x);
      auto gather = [&](float* in) {   // cat(v[src], v[dst]), zero padded to K0
#pragma unroll
        for (int k = 0; k < K0; ++k) in[k] = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          in[k] = __ldg(v + (size_t)s_node * C + k);
          in[C + k] = __ldg(v + (size_t)d_node * C + k);
        }
      };
      // this thread's half of hidden = relu(W in + b) of encoder `which` (0 edge_free_code.0, 1 edge_code.0), plain FMAs
      auto simt_hidden = [&](const float* in, int which, float* h) {
        const float* Wf = img + Cf::ENC0F + which * 32 * Cf::K4;
        const float* bb = vec + (which ? Cf::vEC0b : Cf::vEF0b);
#pragma unroll
        for (int n = 0; n < H; ++n) {
          float a = bb[c0 + n];
#pragma unroll
          for (int k4 = 0; k4 < Cf::K4; k4 += 4) {
            const float4 w = *reinterpret_cast<const float4*>(Wf + (c0 + n) * Cf::K4 + k4);
            a = fmaf(w.x, in[k4], a); a = fmaf(w.y, in[k4 + 1], a); a = fmaf(w.z, in[k4 + 2], a); a = fmaf(w.w, in[k4 + 3], a);
          }
          h[n] = fmaxf(a, 0.f);
        }
      };
      auto stage_edge_code = [&]() {
        float in[K0];
        gather(in);
        if constexpr (Cf::kSimtIn) {
          float h[H];
          simt_hidden(in, 1, h);
          umma::st_split<H>(tc + Cf::cHH + c0, tc + Cf::cHL + c0, h);
        } else {
          umma::st_split<K0 / 2>(tc + Cf::cHH + half * (K0 / 2), tc + Cf::cHL + half * (K0 / 2), in + half * (K0 / 2));
        }
      };
      float x[H];
      // ---- edge_free_code                                                  (model.py:123)
      {
        float in[K0];
        if constexpr (Cf::kSimtIn) {
#pragma unroll
          for (int k = 0; k < K0; ++k) in[k] = k < 2 * C ? in_nx[k] : 0.f;   // requested before the previous unit's tail
          simt_hidden(in, 0, x);
        } else {
          gather(in);
          umma::st_split<K0 / 2>(tc + Cf::cXH + half * (K0 / 2), tc + Cf::cXL + half * (K0 / 2), in + half * (K0 / 2));
          publish();
          await();
          umma::ld16(tc + Cf::cA1 + c0, x);
          umma::wait_ld();
#pragma unroll
          for (int n = 0; n < H; ++n) x[n] = fmaxf(x[n] + vec[Cf::vEF0b + c0 + n], 0.f);
        }
      }
      umma::st_split<H>(tc + Cf::cXH + c0, tc + Cf::cXL + c0, x);
      publish();
      await();
      umma::ld16(tc + Cf::cA1 + c0, x);
      umma::wait_ld();
#pragma unroll
      for (int n = 0; n < H; ++n) x[n] += vec[Cf::vEF2b + c0 + n];
      umma::st_split<H>(tc + Cf::cXH + c0, tc + Cf::cXL + c0, x);
      if (n_blocks == 0) stage_edge_code();
      publish();
      // ---- three edge Blocks                                               (model.py:130, 153-218)
      for (int blk = 0; blk < n_blocks; ++blk) {
        const float* bv = vec + Cf::vBLK + blk * 192;
        float acc[H];
        float m, l = 1.0f;
        await(0);   // Gx | Vx (+ scores of the first one or two sub-chunks)
        {
          float u[H];
          umma::ld16(tc + Cf::cA1 + c0, u);
          umma::ld16(tc + Cf::cA1 + 32 + c0, acc);          // value of the row itself, weight exp2(s_self - m) = 1
          umma::wait_ld();
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;     // self score x^T G x   (model.py:175,177)
#pragma unroll
          for (int n = 0; n < H; n += 4) {
            s0 = fmaf(u[n], x[n], s0); s1 = fmaf(u[n + 1], x[n + 1], s1); s2 = fmaf(u[n + 2], x[n + 2], s2); s3 = fmaf(u[n + 3], x[n + 3], s3);
          }
          m = xsum((s0 + s1) + (s2 + s3));
        }
        // online softmax over one sub-chunk, two rolled passes over this warp's 16-column pieces (pieces half, half + 2, ...):
        // scores at TMEM column `col` -> probabilities (hi plane in place, lo plane at cPL); returns the rescale factor
        auto softmax_sub = [&](uint32_t col, int cnt) -> float {
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
          for (int j = half * 16; j < per; j += 32) {
            float sc[16];
            umma::ld16(tc + col + j, sc);
            umma::wait_ld();
            if (j + 16 > cnt) {                           // padded table rows score 0: weight 0
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (j + i >= cnt) sc[i] = -INFINITY;
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              mx0 = fmaxf(mx0, sc[i]); mx1 = fmaxf(mx1, sc[i + 1]); mx2 = fmaxf(mx2, sc[i + 2]); mx3 = fmaxf(mx3, sc[i + 3]);
            }
          }
          const float mnew = fmaxf(m, xmax(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3))));
          const float corr = umma::ex2_approx(m - mnew);
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
          for (int j = half * 16; j < per; j += 32) {
            float sc[16];
            umma::ld16(tc + col + j, sc);
            umma::wait_ld();
            if (j + 16 > cnt) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (j + i >= cnt) sc[i] = -INFINITY;
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float p0 = umma::ex2_approx(sc[i] - mnew), p1 = umma::ex2_approx(sc[i + 1] - mnew),
                          p2 = umma::ex2_approx(sc[i + 2] - mnew), p3 = umma::ex2_approx(sc[i + 3] - mnew);
              l0 += p0; l1 += p1; l2 += p2; l3 += p3;
              sc[i] = p0; sc[i + 1] = p1; sc[i + 2] = p2; sc[i + 3] = p3;
            }
            umma::st_split<16>(tc + col + j, tc + Cf::cPL + j, sc);
          }
          l = fmaf(l, corr, xsum((l0 + l1) + (l2 + l3)));
          m = mnew;
          return corr;
        };
        for (int s2 = 0; 2 * s2 < nch; ++s2) {
          const int ns = min(2, nch - 2 * s2);
          if (s2 > 0) {
            // (> 128 obstacles) the probabilities' lo plane overwrote XH / XL: restore the A operand for these scores
            umma::st_split<H>(tc + Cf::cXH + c0, tc + Cf::cXL + c0, x);
            publish();
            await();
          }
          for (int sub = 0; sub < ns; ++sub) {
            const float corr = softmax_sub(sub ? Cf::cSC1 : Cf::cSC, max(0, min(per, O - (2 * s2 + sub) * per)));
            publish();   // -> P.V of this sub-chunk
            await(1);
            float pv[H];
            umma::ld16(tc + (ns == 2 ? Cf::cPV2 : Cf::cPV1) + c0, pv);
            umma::wait_ld();
#pragma unroll
            for (int n = 0; n < H; ++n) acc[n] = fmaf(acc[n], corr, pv[n]);
          }
        }
        // softmax normalisation, residual, attention.layer_norm               (model.py:181)
        {
          const float inv = 1.0f / l;
#pragma unroll
          for (int n = 0; n < H; ++n) acc[n] = fmaf(acc[n], inv, x[n]);
        }
        layernorm_half(acc, bv, bv + 32);
        umma::st_split<H>(tc + Cf::cXH + c0, tc + Cf::cXL + c0, acc);
        publish();
        await(2);   // map_feed.w_1                                             (model.py:193-201)
#ifdef GMP_TC_PROFILE
        const long long q0 = clock64();
#endif
        umma::ld16(tc + Cf::cA1 + c0, x);
        umma::wait_ld();
#ifdef GMP_TC_PROFILE
        const long long q1 = clock64();
#endif
#pragma unroll
        for (int n = 0; n < H; ++n) x[n] = fmaxf(x[n] + bv[64 + c0 + n], 0.f);
        umma::st_split<H>(tc + Cf::cXH + c0, tc + Cf::cXL + c0, x);
#ifdef GMP_TC_PROFILE
        const long long q2 = clock64();
        umma::wait_st();
        const long long q3 = clock64();
#endif
        publish();
#ifdef GMP_TC_PROFILE
        prof_q[0] += q1 - q0; prof_q[1] += q2 - q1; prof_q[2] += q3 - q2; prof_q[3] += clock64() - q3;
#endif
        await(3);   // map_feed.w_2
        umma::ld16(tc + Cf::cA1 + c0, x);
        umma::wait_ld();
#pragma unroll
        for (int n = 0; n < H; ++n) x[n] += bv[96 + c0 + n] + acc[n];
        layernorm_half(x, bv + 128, bv + 160);
        umma::st_split<H>(tc + Cf::cXH + c0, tc + Cf::cXL + c0, x);
        if (blk == n_blocks - 1) stage_edge_code();
        publish();
      }
      // ---- hidden layer of edge_code on the tensor cores (wide inputs)     (model.py:120)
      if constexpr (!Cf::kSimtIn) {
        await();
        float h[H];
        umma::ld16(tc + Cf::cSC + c0, h);
        umma::wait_ld();
#pragma unroll
        for (int n = 0; n < H; ++n) h[n] = fmaxf(h[n] + vec[Cf::vEC0b + c0 + n], 0.f);
        umma::st_split<H>(tc + Cf::cHH + c0, tc + Cf::cHL + c0, h);
        publish();
      }
      // ---- Q = Wc ef + b (model.py:145-146);  P = W4 ef + W5 edge_code + b (model.py:39), edge_code.2 folded into W5
      gather_next();   // (the next unit's endpoint ids landed a unit ago)
      await(4);
      float q[H];
      umma::ld16(tc + Cf::cA1 + c0, q);
      umma::ld16(tc + Cf::cA1 + 32 + c0, x);
      umma::wait_ld();
      if (valid) {
        float* qr = Q + (size_t)slot * E + c0;
        float* pr = P + (size_t)slot * E + c0;
        const float* qb = vec + Cf::vQb + c0;
        const float* pb = vec + Cf::vPb + c0;
#pragma unroll
        for (int n = 0; n < H; n += 8) {
          umma::st_global_v8(qr + n, q[n] + qb[n], q[n + 1] + qb[n + 1], q[n + 2] + qb[n + 2], q[n + 3] + qb[n + 3], q[n + 4] + qb[n + 4],
                             q[n + 5] + qb[n + 5], q[n + 6] + qb[n + 6], q[n + 7] + qb[n + 7]);
          umma::st_global_v8(pr + n, x[n] + pb[n], x[n + 1] + pb[n + 1], x[n + 2] + pb[n + 2], x[n + 3] + pb[n + 3], x[n + 4] + pb[n + 4],
                             x[n + 5] + pb[n + 5], x[n + 6] + pb[n + 6], x[n + 7] + pb[n + 7]);
        }
      }
    }
#ifdef GMP_TC_PROFILE
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && qd == 0)
      printf("tc8 profile: tile %d half %d total %lld cyc (%d units), pair barriers %lld; wait/epilogue per stage kind: A %lld/%lld PV %lld/%lld "
             "w1 %lld/%lld w2 %lld/%lld tail %lld/%lld enc %lld/%lld; w1 epilogue: ld %lld math+st %lld wait_st %lld arrive %lld\n", tile, half, clock64() - prof_t0, (n_units + gridDim.x - 1) / gridDim.x, prof_x,
             prof_w[0], prof_e[0], prof_w[1], prof_e[1], prof_w[2], prof_e[2], prof_w[3], prof_e[3], prof_w[4], prof_e[4], prof_w[5], prof_e[5], prof_q[0], prof_q[1], prof_q[2], prof_q[3]);
#endif
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp_u == kIssuer) umma::tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------------------------------------
// messages + max aggregation for one round on the tensor cores, e = 32            (model.py:33,38-41)
// ------------------------------------------------------------------------------------------------
// Same contract and the same memory pipeline as edge_msg_kernel (explorer.cu): persistent CTAs, the loop-invariant edge
// term P streamed one tile ahead by TMA bulk copies, A[src] / B[dst] rows gathered from L2 with 8 lanes per row, segmented
// max with one RED.MAX per (segment, feature).  Only the per-edge 32x32 product m = lin_0[2](hidden) moves: the hidden rows
// go to TMEM as TF32 hi / lo planes (thread == row), one elected lane issues the 12 tcgen05.mma of the 3xTF32 product and the
// result comes back with tcgen05.ld.  Tiles are 128 rows (one MMA); each CTA owns 128 TMEM columns, four CTAs per SM.
template <int E_>
struct MsgTc {
  static constexpr int E = E_, R = 128, XP = E + 4;   // XP: row pitch of the hidden / message tile (floats)
  static constexpr int kPS = R * E;              // staged P tile (row-major)
  // hidden / message tile, ROW-major with a 4-float pad: the 8-lanes-per-row gather stores float4s, the thread-per-row
  // readers load float4s (a quarter warp = 8 rows, 16 B each, 144 B apart: all 32 banks), the (feature, row group) scan reads
  // 32 consecutive floats -- every access conflict-free, a quarter of the LSU instructions of a feature-major tile
  static constexpr int kX = R * XP;
  static constexpr int kW = 2 * E * E + 2 * E;   // lin_0[2] / policy[2]: hi plane | lo plane | bias | (policy[4] weight)
  static constexpr size_t kNeed = (size_t)(kPS + kX + kW) * sizeof(float) + 4 * R * sizeof(int);   // IDX / SRC double-buffered
  // tensor memory: XH | XL | D = 3E columns, allocated as a power of two; kCtas CTAs per SM share the 512 columns.  The
  // shared-memory request is padded so that exactly kCtas fit: one more could not allocate tensor memory and would stall
  static constexpr int kTmemCols = E == 32 ? 128 : 256;
  static constexpr int kCtas = 512 / kTmemCols;
  static constexpr size_t kPad = (size_t)(227 * 1024) / (kCtas + 1) + 1024;
  static constexpr size_t kBytes = kNeed > kPad ? kNeed : kPad;
  static_assert(kBytes * kCtas <= 227 * 1024, "kCtas CTAs must fit in shared memory");
  static constexpr int cXH = 0, cXL = E, cD = 2 * E;
};

// POLICY = false: messages, AGG = segmented max of lin_0[2](hidden)                              (model.py:33,38-41)
// POLICY = true : policy head, logit = policy[4](relu(policy[2](hidden))) with hidden = relu(G[src] + H[dst] + Q), written
//                 at the edge's COO position (and into the dense [N, N] matrix)               (model.py:145-150);
//                 the weight image is then [hi plane | lo plane | bias | policy.4 weight]
struct PolicyOut {
  const int32_t* csr_eid; const int32_t* edge_ptr; const int32_t* node_ptr; const int64_t* dense_off; int n_graphs;
  float* logits; float* dense;
};

template <int E_, bool POLICY>
__global__ void __launch_bounds__(128, MsgTc<E_>::kCtas) edge_msg_tc_kernel(const float* __restrict__ w_l02 /* planes + bias */, int n_slots,
                                                             const int32_t* __restrict__ csr_src, const int32_t* __restrict__ csr_dst,
                                                             const float* __restrict__ A, const float* __restrict__ B,
                                                             const float* __restrict__ P, float* __restrict__ AGG, PolicyOut po) {
  using M = MsgTc<E_>;
  constexpr int E = M::E, R = M::R, XP = M::XP;
  constexpr int LPR = E / 4, RPP = 128 / LPR;
  extern __shared__ __align__(128) float smem_msg[];
  float* PS = smem_msg;                       // first: 128 B aligned for the bulk copy
  float* X = PS + M::kPS;
  float* WB = X + M::kX;
  // target / source ids of the tile, DOUBLE-BUFFERED by tile parity: for E = 64 the segmented-max scan of a warp reads the ids
  // of rows another warp owns, and the next tile's ids are stored before the loop-top barrier -- with one buffer a fast warp
  // would overwrite ids a slow warp is still scanning.  With two, a buffer is rewritten only two barriers later.
  int* IDX2 = reinterpret_cast<int*>(WB + M::kW);
  __shared__ uint64_t bar_p, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int n_tiles = (n_slots + R - 1) / R;
  int tile = blockIdx.x;
  auto tile_bytes = [&](int t) { return (uint32_t)(min(R, n_slots - t * R) * E * (int)sizeof(float)); };
  if (threadIdx.x == 0) {
    mbar_init(&bar_p, 1);
    mbar_init(&bar_mma, 1);
    if (tile < n_tiles) {
      mbar_expect_tx(&bar_p, tile_bytes(tile));
      tma_bulk_g2s(PS, P + (size_t)tile * R * E, tile_bytes(tile), &bar_p);
    }
  }
  if (warp_u == 0) umma::tmem_alloc(&tmem_slot, M::kTmemCols);
  for (int i = threadIdx.x; i < M::kW / 4; i += 128) reinterpret_cast<float4*>(WB)[i] = __ldg(reinterpret_cast<const float4*>(w_l02) + i);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_slot;
  const uint32_t tm_u = __shfl_sync(0xffffffffu, tm, 0);
  const uint32_t trow = tm + ((uint32_t)(warp_u * 32) << 16);
  const uint32_t wh = umma::desc_lo32(smem_u32(WB), E), wl = umma::desc_lo32(smem_u32(WB + E * E), E);
  const float* bias = WB + 2 * E * E;
  uint32_t ph_p = 0, ph_m = 0;
  const int q = threadIdx.x % LPR, rsub = threadIdx.x / LPR;
  int nsrc, ndst;
  auto load_indices = [&](int t) {
    const int slot = t * R + threadIdx.x;
    const bool ok = t < n_tiles && slot < n_slots;
    nsrc = ok ? __ldg(csr_src + slot) : -1;
    ndst = ok ? __ldg(csr_dst + slot) : -1;
  };
  load_indices(tile);
  for (int par = 0; tile < n_tiles; tile += gridDim.x, par ^= 1) {
    int* IDX = IDX2 + par * 2 * R;
    int* SRC = IDX + R;
    SRC[threadIdx.x] = nsrc;
    IDX[threadIdx.x] = ndst;
    __syncthreads();           // indices visible; X free (previous scan done)
    load_indices(tile + gridDim.x);
    mbar_wait(&bar_p, ph_p);   // this tile's P block has landed
    ph_p ^= 1;
    constexpr int UB = 8;      // 8 rows per thread in flight: all 16 row gathers are issued before the first is consumed
#pragma unroll
    for (int base = 0; base < R; base += UB * RPP) {
      float4 av[UB], bv[UB];
      int sv[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int r = base + rsub + u * RPP;
        const int s = SRC[r], d = IDX[r];
        sv[u] = s;
        av[u] = __ldg(reinterpret_cast<const float4*>(A + (size_t)max(s, 0) * E) + q);
        bv[u] = __ldg(reinterpret_cast<const float4*>(B + (size_t)max(d, 0) * E) + q);
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int r = base + rsub + u * RPP;
        const float4 p = *reinterpret_cast<const float4*>(PS + r * E + 4 * q);
        const bool ok = sv[u] >= 0;
        float4 hv;
        hv.x = ok ? fmaxf(av[u].x + bv[u].x + p.x, 0.0f) : 0.0f;
        hv.y = ok ? fmaxf(av[u].y + bv[u].y + p.y, 0.0f) : 0.0f;
        hv.z = ok ? fmaxf(av[u].z + bv[u].z + p.z, 0.0f) : 0.0f;
        hv.w = ok ? fmaxf(av[u].w + bv[u].w + p.w, 0.0f) : 0.0f;
        *reinterpret_cast<float4*>(X + r * XP + 4 * q) = hv;
      }
    }
    __syncthreads();           // X complete; staging buffer consumed
    const int next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < n_tiles) {
      mbar_expect_tx(&bar_p, tile_bytes(next));
      tma_bulk_g2s(PS, P + (size_t)next * R * E, tile_bytes(next), &bar_p);
    }
    {
      // this thread's hidden row -> TF32 hi / lo planes in TMEM (the A operand)
      float h[E];
#pragma unroll
      for (int k = 0; k < E; k += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(X + threadIdx.x * XP + k);
        h[k] = t4.x; h[k + 1] = t4.y; h[k + 2] = t4.z; h[k + 3] = t4.w;
      }
      umma::st_split<E>(trow + M::cXH, trow + M::cXL, h);
      umma::wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();           // every row published (and every thread has read its row of X)
    if (warp_u == 0) {
      umma::fence_after_sync();
      if (umma::elect_one()) {
        umma::gemm3_fixed<E, E, E>(tm_u + M::cD, tm_u + M::cXH, tm_u + M::cXL, wh, wl, false);
        umma::commit(&bar_mma);
      }
      __syncwarp();
    }
    umma::mbar_wait_guard(&bar_mma, ph_m);
    ph_m ^= 1;
    umma::fence_after_sync();
    if constexpr (POLICY) {
      float mrow[E];
      tc_detail::ld_cols<E>(trow + M::cD, mrow);
      umma::wait_ld();
      const float* w4 = bias + E;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int n = 0; n < E; n += 4) {
        l0 = fmaf(fmaxf(mrow[n] + bias[n], 0.f), w4[n], l0);
        l1 = fmaf(fmaxf(mrow[n + 1] + bias[n + 1], 0.f), w4[n + 1], l1);
        l2 = fmaf(fmaxf(mrow[n + 2] + bias[n + 2], 0.f), w4[n + 2], l2);
        l3 = fmaf(fmaxf(mrow[n + 3] + bias[n + 3], 0.f), w4[n + 3], l3);
      }
      const float logit = (l0 + l1) + (l2 + l3);
      const int slot = tile * R + threadIdx.x;
      if (slot < n_slots) {
        po.logits[po.csr_eid[slot]] = logit;
        if (po.dense) {
          const int g = find_segment(po.edge_ptr, po.n_graphs, slot);
          const int n0 = po.node_ptr[g];
          const int64_t ng = po.node_ptr[g + 1] - n0;
          const int sn = SRC[threadIdx.x] - n0, dn = IDX[threadIdx.x] - n0;
          po.dense[po.dense_off[g] + (int64_t)dn * ng + sn] = logit;   // out[dst, src]   (model.py:149)
        }
      }
      umma::fence_before_sync();
      continue;   // (the loop-top barrier orders this tile's tensor-memory reads before the next tile's writes)
    }
    {
      float mrow[E];
      tc_detail::ld_cols<E>(trow + M::cD, mrow);
      umma::wait_ld();
      // (no bias here: x -> fl(x + b) is monotone, so max_i fl(m_i + b) == fl(max_i m_i + b) bit for bit -- the bias is added
      //  once per (segment, feature) when the maximum is flushed instead of once per (row, feature))
#pragma unroll
      for (int n = 0; n < E; n += 4)
        *reinterpret_cast<float4*>(X + threadIdx.x * XP + n) = make_float4(mrow[n], mrow[n + 1], mrow[n + 2], mrow[n + 3]);
    }
    umma::fence_before_sync();
    __syncthreads();
    // segmented max: thread = (feature n, row group); rows of a group are walked in CSR order, so a target's rows are
    // consecutive; one RED per (segment, feature), 128 B coalesced across the warp.  Every lane of a warp walks the same
    // rows: the segment heads are found once per 32 rows (one target id per lane, shuffle + ballot) instead of re-reading
    // the id of every row -- the scan is then one shared-memory load per row.
    constexpr int GROUPS = 128 / E;
    constexpr int ROWS = R / GROUPS;
    const int n = threadIdx.x % E;
    const int r0 = (threadIdx.x / E) * ROWS;
    const int lane = threadIdx.x & 31;
    int cur = -1;
    float run = -INFINITY;
    const float bn = bias[n];
#pragma unroll
    for (int c = 0; c < ROWS / 32; ++c) {
      const int dl = IDX[r0 + 32 * c + lane];
      int prev = __shfl_up_sync(0xffffffffu, dl, 1);
      if (lane == 0) prev = cur;   // (cur is warp-uniform: the target of the previous chunk's last row, -1 at the start)
      const uint32_t heads = __ballot_sync(0xffffffffu, dl != prev);
      float xv[32];                // the column slice first (32 independent loads), then the dependent max chain
#pragma unroll
      for (int i = 0; i < 32; ++i) xv[i] = X[(r0 + 32 * c + i) * XP + n];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if ((heads >> i) & 1u) {
          if (cur >= 0) atomic_max_f32(AGG + (size_t)cur * E + n, run + bn);
          cur = __shfl_sync(0xffffffffu, dl, i);
          run = -INFINITY;
        }
        run = fmaxf(run, xv[i]);
      }
    }
    if (cur >= 0) atomic_max_f32(AGG + (size_t)cur * E + n, run + bn);
    // (the loop-top barrier protects X, the parity buffers protect IDX; tensor memory is rewritten only after the next tile's barriers)
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp_u == 0) umma::tmem_dealloc(tm, M::kTmemCols);
}

}  // namespace gmp
