// Opaque handle shared by the translation units of libgnnmp.so.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace gmp {

// One attention Block of a map stream (model.py:204-218), device-packed.  All offsets are in floats
// into the packed weight image; every matrix is stored K-major (Wt[k][n]) and is immediately
// followed by the vectors its stage needs, so a stage is ONE contiguous global->shared copy.
struct BlockW {
  int Gt;    // [E][E]  scale * Wq^T Wk        (self score  s = x^T G x)
  int Wvt;   // [E][E] | ln_g[E] | ln_b[E]      (value, attention.layer_norm)
  int W1t;   // [E][E] | b1[E]                  (map_feed.w_1)
  int W2t;   // [E][E] | b2[E] | ln_g | ln_b    (map_feed.w_2, map_feed.layer_norm)
};
struct ObsBlockW {
  int Wkt;   // [E][E]                          key
  int WqS;   // [E][E]  scale * Wq              (M_o = scale * Wq^T key_o)
  int Wvt;   // [E][E]                          value
  int W1t;   // [E][E] | b1                     (obs_feed.w_1)
  int W2t;   // [E][E] | b2 | ln_g | ln_b       (obs_feed.w_2, obs_feed.layer_norm)
};
struct ExplorerW {
  BlockW node_blk[3], edge_blk[3];
  ObsBlockW obs_blk[2][3];           // [stream: 0 node, 1 edge]
  int obs0[2], obs2[2];              // obs_{node,edge}_code.{0,2}: [Wt | b]
  int nc0, nc2, nf0, nf2;            // node_code / node_free_code: [Wt | b]
  int ef0, ef2, ec0, ec2;            // edge_free_code / edge_code: [Wt | b]
  int enc_nc, enc_nf, enc_u3, enc_h; // encoder slices: We1t ; [We2t | b] ; We3*goal_encoder [E] ; We4t
  int dec_nc, dec_h;                 // decoder slices: [Wd1t | b] ; Wd2t
  int l0_ef, l0_ec;                  // lin_0.0 slices on edge_free / edge_code: W4t ; [W5t | b]
  int l0_A, l0_B;                    // (W1+W2)^T ; (W3-W1)^T
  int l0_2;                          // lin_0.2: [Wt | b]
  int l1_x, l1_a;                    // lin_1 slices: Wxt ; [Wat | b]
  int p0_ef, p0_G, p0_H;             // policy.0: [Wct | b] ; (Wa+Wb)^T ; (-Wb)^T
  int p2;                            // policy.2: [Wt | b | policy.4 weight [E]]
  int goal_enc;                      // [E]
  int tc_img;                        // tensor-core image of the edge-feature stage (explorer_tc.cuh), -1 if none
  int tc_l02;                        // lin_0.2 for the tensor-core message kernel: [hi plane | lo plane | bias | pad], -1 if none
  int tc_p2;                         // policy.2 for the tensor-core policy kernel: [hi plane | lo plane | bias | policy.4 weight]
  int tc64_img[5];                   // embed 64: phase images of explorer_tc64.cuh (encoder, Block 0..2, tail), -1 if none
};

struct ExplorerModel {
  int c = 0, e = 0, s = 0;
  bool ready = false;
  // -1 auto, 0 fp32 SIMT, 1 tcgen05 3xTF32 with eight epilogue warps per tile and one lockstep issuer, 2 the same with four warps
  // per tile, 3 eight warps per tile and one issuer warp per tile (what auto runs for narrow inputs with 1..128 obstacles)
  int edge_feature_mode = -1;
  std::map<std::string, std::vector<float>> tensors;  // reference state_dict (live entries)
  ExplorerW w{};
  float* d_weights = nullptr;
  int64_t n_weights = 0;
};

struct SmootherW {
  int nc0;   // node_code.0 with BatchNorm (eval) folded: [Wt (c+3 x E) | b]
  int nc3;   // node_code.3: [Wt | b]
  int l0_A, l0_B;  // lin_0.0 split: (W1+W2)^T ; [(W3-W1)^T | b]
  int l0_2;  // [Wt | b]
  int l1_0, l1_2;  // [Wt | b]
  int sm;    // smooth_node: [Wt (E x CP) | b (CP)], CP = c rounded up to 4
};

struct SmootherModel {
  int c = 0, e = 0;
  bool ready = false;
  std::map<std::string, std::vector<float>> tensors;
  SmootherW w{};
  float* d_weights = nullptr;
  int64_t n_weights = 0;
};

// Optional per-phase device timing: cudaEvents recorded on the launching stream around each phase of the
// last explorer forward (bench.py's roofline numbers come from here, not from a profiler).
enum Phase { kPhCsr = 0, kPhGoal, kPhObstacle, kPhNodePre, kPhEdgeFeature, kPhNodeLoop, kPhEdgeMsg, kPhPolicy, kNumPhases };

struct Timeline {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Span { int phase; cudaEvent_t a, b; };
  std::vector<Span> spans;
  cudaEvent_t next() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
  void reset() { used = 0; spans.clear(); }
  void begin(int phase, cudaStream_t st) {
    if (!enabled) return;
    Span s{phase, next(), nullptr};
    cudaEventRecord(s.a, st);
    spans.push_back(s);
  }
  void end(cudaStream_t st) {
    if (!enabled) return;
    spans.back().b = next();
    cudaEventRecord(spans.back().b, st);
  }
};

}  // namespace gmp

struct gmp_handle {
  int device = 0;
  gmp::ExplorerModel ex;
  gmp::SmootherModel sm;
  gmp::Timeline tl;
  // launch state of the explorer kernels on THIS handle's device (shared-memory opt-in done, persistent grid size)
  bool ex_attr_done = false;
  int ex_msg_grid = 0;
  // fork / join of the explorer forward: the CSR build (bound by L2 atomic latency) runs on this side stream next to the
  // node-side kernels (low occupancy) of the launching stream and is joined before the first kernel that reads the CSR
  cudaStream_t ex_side = nullptr;
  cudaEvent_t ex_fork = nullptr, ex_join = nullptr;
  const int32_t* ex_bad_edges = nullptr;   // device counter (in the caller's workspace) of out-of-range edge ids in the last forward
};
