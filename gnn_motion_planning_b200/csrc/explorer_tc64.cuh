// Tensor-core (tcgen05 / TMEM, 3xTF32) edge-feature stage for embed 64 (kuka7), O <= 32 obstacles per graph.
//
// Same contract as edge_feature_kernel (explorer.cu) and the same machinery as edge_feature_tc_kernel (explorer_tc.cuh):
// one persistent 384-thread CTA per SM, two 128-edge tiles (thread == edge row == TMEM lane) ping-ponging on an MMA-issuer
// warp, A operands and accumulators in TMEM, weights / obstacle tables as TF32 hi / lo planes in shared memory.
// What differs at e = 64:
//   * the hi / lo planes of ONE Block (Gx|Vx 64 KB, w_1 32 KB, w_2 32 KB) already fill shared memory, so the stage is split
//     into PHASES, one launch each -- 0: edge_free_code, 1: one edge Block (launched three times), 2: Q / P tail -- with the
//     64-float activation row of every edge handed from phase to phase through the (L2-resident, later overwritten) Q buffer;
//   * a tile's 256 TMEM columns are X hi [0,64) | X lo [64,128) | A1 [128,256): `Gx | Vx` needs all of A1, so scores get their
//     own MMA stage afterwards (scores -> A1[0,per), probabilities hi in place, lo at A1[64, 64+per)), and the P.V product
//     lands in the dead X hi columns.
#pragma once
#include "explorer_tc.cuh"

namespace gmp {

template <int C>
struct Tc64Cfg {
  static constexpr int E = 64;
  static constexpr int K0 = (2 * C + 7) / 8 * 8;
  static constexpr int kPerMax = 32;                      // obstacle rows (one chunk); larger graphs use the SIMT kernel
  // phase 0 image: ENC0 planes (N=128: edge_free_code.0 rows | edge_code.0 rows, K=K0) | EF2 planes | ef0b ec0b ef2b
  static constexpr int p0ENC0 = 0, p0EF2 = p0ENC0 + 2 * 128 * K0, p0VEC = p0EF2 + 2 * E * E, kImg0 = p0VEC + 3 * E;
  // phase 1 image (per Block): GV planes (N=128: G rows | Wv rows) | W1 | W2 | ln1g ln1b b1 b2 ln2g ln2b
  static constexpr int p1GV = 0, p1W1 = p1GV + 2 * 128 * E, p1W2 = p1W1 + 2 * E * E, p1VEC = p1W2 + 2 * E * E, kImg1 = p1VEC + 6 * E;
  // phase 2 image: QP planes (N=128: policy.0 edge_free cols | lin_0.0 edge_free cols) | W52 | EC0 planes (N=64, K=K0) | ec0b qb pb
  static constexpr int p2QP = 0, p2W52 = p2QP + 2 * 128 * E, p2EC0 = p2W52 + 2 * E * E, p2VEC = p2EC0 + 2 * E * K0, kImg2 = p2VEC + 3 * E;
  static constexpr int kTab = 4 * E * kPerMax;            // floats: Mt hi | Mt lo | Vt hi | Vt lo
  static constexpr int cXH = 0, cXL = 64, cA1 = 128;
  template <int PHASE> static constexpr int image() { return PHASE == 0 ? kImg0 : (PHASE == 1 ? kImg1 : kImg2); }
  template <int PHASE> static constexpr size_t smem() { return (size_t)(image<PHASE>() + (PHASE == 1 ? kTab : 0)) * sizeof(float); }
};

namespace tc_detail {

// old-format obstacle tables of the edge stream (OT = 16 rows per tile at e = 64) -> [Mt_hi float[16][per][4] | Mt_lo |
// Vt_hi float[per/4][64][4] | Vt_lo] per (block, graph); one chunk of per = tc_per(O, 1) <= 32 rows
__global__ void __launch_bounds__(256) obs_table_tc64_kernel(const float* __restrict__ tables, int64_t table_stride,
                                                             const int32_t* __restrict__ obs_ptr,
                                                             const int32_t* __restrict__ obs_tile_ptr,
                                                             const int64_t* __restrict__ tc_tab_off, float* __restrict__ tc_tables,
                                                             int64_t tc_tab_stride) {
  constexpr int E = 64, OT = 16;
  const int g = blockIdx.x, blk = blockIdx.y;
  const int O = obs_ptr[g + 1] - obs_ptr[g];
  const int per = tc_per(O, O > 0);
  const float* tab = tables + (size_t)(1 * 3 + blk) * table_stride + (size_t)obs_tile_ptr[g] * (2 * E * OT);
  float* unit = tc_tables + (size_t)blk * tc_tab_stride + tc_tab_off[g];
  for (int i = threadIdx.x; i < per * E; i += blockDim.x) {
    const int oo = i / E, f = i % E;
    float m = 0.f, vv = 0.f;
    if (oo < O) {
      const float* tile = tab + (size_t)(oo / OT) * (2 * E * OT);
      m = tile[f * OT + (oo % OT)];
      vv = tile[E * OT + (oo % OT) * E + f];
    }
    const int im = ((f >> 2) * per + oo) * 4 + (f & 3);
    const int iv = ((oo >> 2) * E + f) * 4 + (oo & 3);
    const float mh = umma::tf32_rna(m), vh = umma::tf32_rna(vv);
    unit[im] = mh;
    unit[E * per + im] = umma::tf32_rna(m - mh);
    unit[2 * E * per + iv] = vh;
    unit[3 * E * per + iv] = umma::tf32_rna(vv - vh);
  }
}

}  // namespace tc_detail

// PHASE 0: xio[slot] = edge_free_code(v[src], v[dst]);  PHASE 1: xio[slot] = Block(xio[slot]) with this launch's image / tables;
// PHASE 2: Q[slot], P[slot] from xio[slot] (xio may alias Q: every thread reads its own row before it writes it).
template <int C, int PHASE>
__global__ void __launch_bounds__(384, 1) edge_feature64_tc_kernel(
    const float* __restrict__ tcw, const float* __restrict__ v, const int32_t* __restrict__ csr_src,
    const int32_t* __restrict__ csr_dst, const int4* __restrict__ unit_meta, int n_units, const float* __restrict__ tc_tables,
    float* xio, float* __restrict__ P, float* Q) {
  using Cf = Tc64Cfg<C>;
  constexpr int E = 64, K0 = Cf::K0, kImage = Cf::template image<PHASE>();
  extern __shared__ __align__(128) float smem_tc[];
  float* img = smem_tc;
  float* tabbuf = smem_tc + kImage;
  __shared__ uint64_t bar_ready[2], bar_done[2], bar_tabfull;
  __shared__ uint32_t tmem_slot;

  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  {
    const float4* s4 = reinterpret_cast<const float4*>(tcw);
    float4* d4 = reinterpret_cast<float4*>(img);
    for (int i = threadIdx.x; i < kImage / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar_ready[0], 128); mbar_init(&bar_ready[1], 128);
    mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1);
    mbar_init(&bar_tabfull, 1);
  }
  if (warp_u == 8) umma::tmem_alloc(&tmem_slot, 512);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_slot;

  if (warp_u >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
  }
  if (warp_u > 8) {
    // warps 9-11 only pad the issuer's warpgroup
  } else if (warp_u == 8) {
    // =================================================================== MMA issuer / table loader
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tm, 0);
    const uint32_t img_s = smem_u32(img), tab_s = smem_u32(tabbuf);
    uint32_t rph[2] = {0, 0}, full_ph = 0;
    auto load_meta = [&](int u) {
      int4 mm = make_int4(0, 0, 0, 0);
      if (u < n_units) mm = __ldg(unit_meta + u);
      mm.z = __shfl_sync(0xffffffffu, mm.z, 0);
      mm.w = __shfl_sync(0xffffffffu, mm.w, 0);
      return mm;
    };
    int4 meta_cur = load_meta(blockIdx.x), meta_next = load_meta(blockIdx.x + gridDim.x);
    // the table of unit u (phase 1): loaded one unit ahead, behind the FFN stages of the previous unit
    auto table_load = [&](const int4 mm) {
      const int O = mm.z;
      if (PHASE != 1 || O <= 0) return;
      const int per = tc_per(O, 1);
      const uint32_t bytes = (uint32_t)(4 * E * per) * 4u;
      if (umma::elect_one()) {
        mbar_expect_tx(&bar_tabfull, bytes);
        tma_bulk_g2s(tabbuf, tc_tables + (size_t)mm.w, bytes, &bar_tabfull);
      }
      __syncwarp();
    };
    table_load(meta_cur);
    auto wd = [&](int off, int rows) { return umma::desc_lo32(img_s + (uint32_t)off * 4u, (uint32_t)rows); };
#define GMP_TC_STAGE(...)                                                         \
  _Pragma("unroll") for (int t = 0; t < 2; ++t) {                                 \
    umma::mbar_wait_guard(&bar_ready[t], rph[t]);                                 \
    rph[t] ^= 1u;                                                                 \
    umma::fence_after_sync();                                                     \
    if (umma::elect_one()) {                                                      \
      const uint32_t tc = tm_u + (uint32_t)t * 256u;                              \
      const uint32_t xh = tc + Cf::cXH, xl = tc + Cf::cXL, a1 = tc + Cf::cA1;     \
      __VA_ARGS__;                                                                \
      umma::commit(&bar_done[t]);                                                 \
    }                                                                             \
    __syncwarp();                                                                 \
  }
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      if (unit != (int)blockIdx.x) {
        meta_cur = meta_next;
        meta_next = load_meta(unit + gridDim.x);
      }
      if constexpr (PHASE == 0) {
        GMP_TC_STAGE((umma::gemm3_fixed<E, K0, 128>(a1, xh, xl, wd(Cf::p0ENC0, 128), wd(Cf::p0ENC0 + 128 * K0, 128), false)));
        GMP_TC_STAGE((umma::gemm3_fixed<E, E, E>(a1, xh, xl, wd(Cf::p0EF2, E), wd(Cf::p0EF2 + E * E, E), false)));
      } else if constexpr (PHASE == 1) {
        const int O = meta_cur.z, per = tc_per(O, O > 0);
        GMP_TC_STAGE((umma::gemm3_fixed<128, E, 128>(a1, xh, xl, wd(Cf::p1GV, 128), wd(Cf::p1GV + 128 * E, 128), false)));
        if (O > 0) {
          umma::mbar_wait_guard(&bar_tabfull, full_ph);
          full_ph ^= 1u;
          const uint32_t pl = (uint32_t)(E * per) * 4u;
          const uint32_t mt_h = umma::desc_lo32(tab_s, (uint32_t)per), mt_l = umma::desc_lo32(tab_s + pl, (uint32_t)per),
                         vt_h = umma::desc_lo32(tab_s + 2 * pl, E), vt_l = umma::desc_lo32(tab_s + 3 * pl, E);
          GMP_TC_STAGE((umma::gemm3_n<E>(a1, xh, xl, mt_h, mt_l, per)));
          GMP_TC_STAGE((umma::gemm3_k<E, E>(xh, a1, a1 + 64, vt_h, vt_l, per)));
        }
        GMP_TC_STAGE((umma::gemm3_fixed<E, E, E>(a1, xh, xl, wd(Cf::p1W1, E), wd(Cf::p1W1 + E * E, E), false)));
        // both tiles have published w_1's operand: every P.V product has retired, the table buffer is free
        table_load(meta_next);
        GMP_TC_STAGE((umma::gemm3_fixed<E, E, E>(a1, xh, xl, wd(Cf::p1W2, E), wd(Cf::p1W2 + E * E, E), false)));
      } else {
        GMP_TC_STAGE((umma::gemm3_fixed<128, E, 128>(a1, xh, xl, wd(Cf::p2QP, 128), wd(Cf::p2QP + 128 * E, 128), false)));
        GMP_TC_STAGE((umma::gemm3_fixed<E, K0, E>(a1, xh, xl, wd(Cf::p2EC0, E), wd(Cf::p2EC0 + E * K0, E), false)));
        GMP_TC_STAGE((umma::gemm3_fixed<E, E, E>(a1, xh, xl, wd(Cf::p2W52, E), wd(Cf::p2W52 + E * E, E), false)));
      }
    }
#undef GMP_TC_STAGE
  } else {
    // =================================================================== compute warps: thread == edge row
    const int tile = warp_u >> 2;
    const int row = threadIdx.x & 127;
    const uint32_t tc = tm + ((uint32_t)((warp_u & 3) * 32) << 16) + (uint32_t)tile * 256u;
    const uint32_t xh = tc + Cf::cXH, xl = tc + Cf::cXL, a1 = tc + Cf::cA1;
    uint32_t dph = 0;
    auto publish = [&]() {
      umma::wait_st();
      umma::fence_before_sync();
      umma::mbar_arrive(&bar_ready[tile]);
    };
    auto await = [&]() {
      umma::mbar_wait_guard(&bar_done[tile], dph);
      dph ^= 1u;
      umma::fence_after_sync();
    };
    // rows are 256 B per lane: 256-bit accesses (whole 32-byte sectors per lane, half the instructions of float4s)
    auto load_row = [&](const float* p, float* x) {   // 64 contiguous floats
#pragma unroll
      for (int n = 0; n < E; n += 8) umma::ld_global_v8(p + n, x + n);
    };
    auto store_row = [&](float* p, const float* x) {
#pragma unroll
      for (int n = 0; n < E; n += 8) umma::st_global_v8(p + n, x[n], x[n + 1], x[n + 2], x[n + 3], x[n + 4], x[n + 5], x[n + 6], x[n + 7]);
    };
    int4 meta_nx = make_int4(0, 0, 0, 0);
    int s_nx = 0, d_nx = 0;
    auto prefetch_unit = [&](int u) {
      if (u < n_units) {
        meta_nx = __ldg(unit_meta + u);
        if (PHASE != 0) {
          // the next unit's activation row comes from HBM (written by the previous launch): pull it into L2 now
          const int sl = min(meta_nx.x + tile * 128 + row, meta_nx.y - 1);
          const float* nx = xio + (size_t)sl * E;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 32));
        }
        if (PHASE != 1) {
          const int sl = meta_nx.x + tile * 128 + row;
          const bool ok = sl < meta_nx.y;
          s_nx = ok ? __ldg(csr_src + sl) : 0;
          d_nx = ok ? __ldg(csr_dst + sl) : 0;
        }
      }
    };
    prefetch_unit(blockIdx.x);
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int4 meta = meta_nx;
      const int slot = meta.x + tile * 128 + row;
      const bool valid = slot < meta.y;
      const int O = meta.z;
      const int s_node = s_nx, d_node = d_nx;
      prefetch_unit(unit + gridDim.x);
      // rows past the end of the graph compute on row `meta.y - 1`'s data (any valid memory) and are never stored
      float* xrow = xio + (size_t)(valid ? slot : meta.x) * E;
      auto gather = [&](float* in) {
#pragma unroll
        for (int k = 0; k < K0; ++k) in[k] = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          in[k] = __ldg(v + (size_t)s_node * C + k);
          in[C + k] = __ldg(v + (size_t)d_node * C + k);
        }
      };
      if constexpr (PHASE == 0) {
        const float* vec = img + Cf::p0VEC;
        float x[E];
        {
          float in[K0];
          gather(in);
          umma::st_split<K0>(xh, xl, in);
        }
        publish();
        await();
        tc_detail::ld_cols<E>(a1, x);
        umma::wait_ld();
#pragma unroll
        for (int n = 0; n < E; ++n) x[n] = fmaxf(x[n] + vec[n], 0.f);
        umma::st_split<E>(xh, xl, x);
        publish();
        await();
        tc_detail::ld_cols<E>(a1, x);
        umma::wait_ld();
#pragma unroll
        for (int n = 0; n < E; ++n) x[n] += vec[2 * E + n];
        if (valid) store_row(xrow, x);
      } else if constexpr (PHASE == 1) {
        const float* bv = img + Cf::p1VEC;
        const int per = tc_per(O, O > 0);
        float acc[E];
        float x[E];   // the Block's input row: A operand, self score, residual
        float m, l = 1.0f;
        load_row(xrow, x);
        umma::st_split<E>(xh, xl, x);
        publish();
        await();   // Gx | Vx
        {
          float u[E];
          tc_detail::ld_cols<E>(a1, u);
          umma::wait_ld();
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int n = 0; n < E; n += 4) {
            s0 = fmaf(u[n], x[n], s0); s1 = fmaf(u[n + 1], x[n + 1], s1); s2 = fmaf(u[n + 2], x[n + 2], s2); s3 = fmaf(u[n + 3], x[n + 3], s3);
          }
          m = (s0 + s1) + (s2 + s3);
        }
        tc_detail::ld_cols<E>(a1 + 64, acc);   // value of the row itself
        umma::wait_ld();
        if (O > 0) {
          publish();   // A1 has been read: the scores may overwrite it
          await();
          float sc[Cf::kPerMax];
#pragma unroll
          for (int j = 0; j < Cf::kPerMax; j += 16)
            if (j < per) umma::ld16(a1 + j, sc + j);
          umma::wait_ld();
#pragma unroll
          for (int j = 0; j < Cf::kPerMax; ++j)
            if (j >= O) sc[j] = -INFINITY;
          float mnew = m;
#pragma unroll
          for (int j = 0; j < Cf::kPerMax; ++j) mnew = fmaxf(mnew, sc[j]);
          const float corr = umma::ex2_approx(m - mnew);
          float lsum = l * corr;
#pragma unroll
          for (int j = 0; j < Cf::kPerMax; ++j) {
            sc[j] = umma::ex2_approx(sc[j] - mnew);
            lsum += sc[j];
          }
          l = lsum;
          m = mnew;
#pragma unroll
          for (int j = 0; j < Cf::kPerMax; j += 16)
            if (j < per) umma::st_split<16>(a1 + j, a1 + 64 + j, sc + j);
          publish();
          await();   // P.V -> X hi columns
          float pv[E];
          tc_detail::ld_cols<E>(xh, pv);
          umma::wait_ld();
#pragma unroll
          for (int n = 0; n < E; ++n) acc[n] = fmaf(acc[n], corr, pv[n]);
        }
        {
          const float inv = 1.0f / l;
#pragma unroll
          for (int n = 0; n < E; ++n) acc[n] = fmaf(acc[n], inv, x[n]);
        }
        tc_detail::layernorm_row<E>(acc, bv, bv + E, 1e-6f);
        umma::st_split<E>(xh, xl, acc);
        publish();
        await();   // map_feed.w_1
        {
          float h[E];
          tc_detail::ld_cols<E>(a1, h);
          umma::wait_ld();
#pragma unroll
          for (int n = 0; n < E; ++n) h[n] = fmaxf(h[n] + bv[2 * E + n], 0.f);
          umma::st_split<E>(xh, xl, h);
        }
        publish();
        await();   // map_feed.w_2
        {
          float h[E];
          tc_detail::ld_cols<E>(a1, h);
          umma::wait_ld();
#pragma unroll
          for (int n = 0; n < E; ++n) h[n] += bv[3 * E + n] + acc[n];
          tc_detail::layernorm_row<E>(h, bv + 4 * E, bv + 5 * E, 1e-6f);
          if (valid) store_row(xrow, h);
        }
      } else {
        const float* vec = img + Cf::p2VEC;
        float pef[E];
        {
          float x[E];
          load_row(xrow, x);
          umma::st_split<E>(xh, xl, x);
        }
        publish();
        await();   // Q | P_ef
        {
          float q[E];
          tc_detail::ld_cols<E>(a1, q);
          tc_detail::ld_cols<E>(a1 + 64, pef);
          umma::wait_ld();
#pragma unroll
          for (int n = 0; n < E; ++n) q[n] += vec[E + n];
          if (valid) store_row(Q + (size_t)slot * E, q);
        }
        {
          float in[K0];
          gather(in);
          umma::st_split<K0>(xh, xl, in);
        }
        publish();
        await();   // edge_code.0
        {
          float h[E];
          tc_detail::ld_cols<E>(a1, h);
          umma::wait_ld();
#pragma unroll
          for (int n = 0; n < E; ++n) h[n] = fmaxf(h[n] + vec[n], 0.f);
          umma::st_split<E>(xh, xl, h);
        }
        publish();
        await();   // (W5 W_ec2) hidden
        {
          float h[E];
          tc_detail::ld_cols<E>(a1, h);
          umma::wait_ld();
#pragma unroll
          for (int n = 0; n < E; ++n) h[n] += pef[n] + vec[2 * E + n];
          if (valid) store_row(P + (size_t)slot * E, h);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp_u == 8) umma::tmem_dealloc(tm, 512);
}

}  // namespace gmp
