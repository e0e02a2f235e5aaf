/* gnnmp.h -- C ABI of libgnnmp.so: the B200 (sm_100a) implementation of the data-parallel hot
 * path of rainorangelemon/gnn-motion-planning.
 *
 * The reference is pure Python and has no FFI; its seam for this path is a set of Python call
 * sites.  Each entry point below names the reference interface it replaces
 * (file:line, relative to the reference root).  The reference-side binding a maintainer would add
 * is a ctypes stub -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no torch types.  Pointers are DEVICE pointers unless the name ends in `_h`.
 *   - the caller allocates every input, output and workspace buffer; nothing is retained after a
 *     call returns except the weights owned by a handle.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only enqueue
 *     work; they do not synchronise unless stated.
 *   - return value: 0 = OK, negative = GMP_E_*; gmp_last_error() gives the message (thread local).
 *   - batches are PACKED: graph g owns nodes  node_ptr[g] .. node_ptr[g+1]  of `v`, edges
 *     edge_ptr[g] .. edge_ptr[g+1] of `edge_index`, obstacle rows obs_ptr[g] .. obs_ptr[g+1].
 *     The *_ptr offset arrays are HOST arrays (the caller knows its own sizes); node ids inside
 *     edge_index are LOCAL to their graph (0 .. N_g-1), int64 like torch_geometric's edge_index:
 *     row 0 = source j, row 1 = target i, stored as [2, E_total] (row stride = E_total).
 *   - there is NO CPU fallback anywhere in this library.
 */
#ifndef GNNMP_H_
#define GNNMP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMP_OK 0
#define GMP_E_INVALID (-1)   /* bad argument / shape */
#define GMP_E_CUDA (-2)      /* CUDA runtime error */
#define GMP_E_STATE (-3)     /* handle not ready (weights missing) */
#define GMP_E_UNSUPPORTED (-4)

#define GMP_DTYPE_F32 0
#define GMP_DTYPE_F64 1

typedef struct gmp_handle gmp_handle;

/* ---- library ------------------------------------------------------------------------------ */
const char* gmp_last_error(void);
const char* gmp_version(void);
/* 1 if a CUDA device of compute capability 10.x is visible; never throws. */
int gmp_device_ok(int device);

gmp_handle* gmp_create(int device);
void gmp_destroy(gmp_handle* h);

/* ---- explorer: EncoderProcessDecoder (model.py:48-150) -------------------------------------- */
/* Replaces nn.Module construction (model.py:49) : dims of the model this handle serves. */
int gmp_explorer_init(gmp_handle* h, int config_size /*c*/, int embed_size /*e: 32 or 64*/, int obs_size /*s*/);
/* Replaces load_state_dict (eval_gnn.py:101): hand over one reference state_dict tensor by its
 * reference name (e.g. "process.lin_0.0.weight"), host fp32, row-major as torch stores it.
 * Dead tensors of the reference dict are accepted and ignored. */
int gmp_explorer_set_tensor(gmp_handle* h, const char* name, const float* data_h, int64_t numel);
/* Validates that every live tensor arrived with the right size, builds the device-side packed
 * (transposed / algebraically pre-combined) weight image. */
int gmp_explorer_finalize(gmp_handle* h);
/* Arithmetic of the edge-feature stage (edge encoders + edge Blocks, model.py:120,123,130):
 *   -1 auto (default) / 1: tcgen05 tensor cores with 3xTF32 split operands (embed_size 64: graphs with more than 32
 *    obstacles fall back to fp32 FMA for the edge-feature stage);  0: fp32 FMA (SIMT) always;  2: tensor cores with the
 *    round-1 tile organisation of the embed-32 kernel (four epilogue warps per tile instead of eight);  3: eight warps per
 *    tile and one MMA issuer warp per tile instead of one for both (what auto runs for narrow inputs, 2c <= 8, when every
 *    graph has 1..128 obstacles; bit-identical to mode 1).  The environment variable GMP_TC_RD=0 / 1 overrides the issuer
 *    choice for A/B measurements.
 * All meet the 1e-4 logit tolerance; the switch exists for A/B parity tests and profiling. */
int gmp_explorer_set_edge_feature_mode(gmp_handle* h, int mode);

/* Bytes of scratch gmp_explorer_forward needs for a batch of these totals. */
int64_t gmp_explorer_workspace_bytes(const gmp_handle* h, int64_t n_graphs, int64_t n_nodes_total,
                                     int64_t n_edges_total, int64_t n_obs_total);

/* Replaces EncoderProcessDecoder.forward (model.py:115-150) for a packed batch of graphs.
 *   v            [N_total, c] f32        node configurations               (model.py:115 `v`)
 *   edge_index   [2, E_total] i64        local node ids                    (`edge_index`)
 *   edge_row_stride  elements between row 0 and row 1 of edge_index (>= E_total; = E_total when contiguous)
 *   goal         [B, c] f32                                                 (`goal`)
 *   obstacles    [O_total, s] f32        obstacle tokens, already viewed [-1, s] (model.py:126)
 *   node_ptr_h / edge_ptr_h / obs_ptr_h  [B+1] i32 host offset arrays
 *   loop         message-passing rounds (eval_gnn.py:14: 5)
 *   use_obstacles  model.use_obstacles (model.py:125)
 *   edge_logits_out [E_total] f32        logit of edge e, aligned with edge_index   (the `policy` vector, model.py:145)
 *   dense_out    nullable; sum_g N_g^2 f32, graph g at offset sum_{g'<g} N_g'^2, out[dst*N_g+src] = logit,
 *                zero elsewhere                                             (model.py:148-150)
 *   workspace    >= gmp_explorer_workspace_bytes(...)
 */
int gmp_explorer_forward(gmp_handle* h, int64_t n_graphs, const float* v, const int64_t* edge_index,
                         int64_t edge_row_stride, const float* goal, const float* obstacles, const int32_t* node_ptr_h,
                         const int32_t* edge_ptr_h, const int32_t* obs_ptr_h, int loop, int use_obstacles,
                         float* edge_logits_out, float* dense_out, void* workspace, int64_t workspace_bytes,
                         void* stream);

/* edge_index ids outside [0, N_g) cannot be rejected without a device round trip: gmp_explorer_forward CLAMPS them (memory
 * safe; the logits of such edges are meaningless) and counts them.  This call synchronises `stream` and returns the count of
 * the last forward on this handle (0 = every id was in range; the caller's workspace must still be alive). */
int gmp_explorer_bad_edges(gmp_handle* h, void* stream);

/* Optional device-side timing of the last gmp_explorer_forward on this handle (replaces the reference's
 * wall-clock Timer spans, environment/timer.py:6-25).  gmp_get_timings synchronises on the recorded events and
 * writes milliseconds per phase into ms_out_h[0..7]:
 *   0 csr build, 1 goal index, 2 obstacle stream, 3 node encoders+blocks, 4 edge encoders+blocks,
 *   5 node updates (sum over rounds), 6 edge messages + max aggregation (sum over rounds), 7 policy head.
 * Returns the number of phases (8) or a negative error. */
int gmp_set_timing(gmp_handle* h, int enable);
int gmp_get_timings(gmp_handle* h, float* ms_out_h, int n);

/* ---- k-NN random geometric graph: create_data (eval_gnn.py:150-165) ------------------------- */
/* Upper bound of edges graph g can emit: 4 * N_g * k1 (two k-NN sets, both directions). */
int64_t gmp_knn_graph_max_edges(int64_t n_nodes, int k1);
int64_t gmp_knn_graph_workspace_bytes(int64_t n_graphs, int64_t n_nodes_total, int64_t max_nodes_per_graph, int k1_max);
/* For every graph: S = kNN_k1(all N_g nodes) U kNN_k1(first n_free_g nodes), self included
 * (knn_graph(.., loop=True), eval_gnn.py:160,162); output = sorted unique of S U reverse(S) by key
 * src*N+dst (coalesce, eval_gnn.py:164).  Distance: fp32 squared L2, left-to-right over dims,
 * ties to the lower index.
 *   v [N_total,c] f32; node_ptr_h [B+1]; n_free_h [B]; k1_h [B] (per graph, eval_gnn.py:159)
 *   edge_index_out [2, edge_capacity] i64 (row stride = edge_capacity), graph g's edges are
 *   written contiguously from column edge_ptr_out[g]; edge_ptr_out [B+1] i32 DEVICE (filled by the call).
 * Returns GMP_E_INVALID if edge_capacity < sum_g 4*N_g*k1_g is not guaranteed to fit. */
int gmp_knn_graph(gmp_handle* h, int64_t n_graphs, const float* v, int c, const int32_t* node_ptr_h,
                  const int32_t* n_free_h, const int32_t* k1_h, int64_t* edge_index_out, int64_t edge_capacity,
                  int32_t* edge_ptr_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- maze collision: MazeEnv (environment/maze_env.py, dim == 2) --------------------------- */
/* Replaces MazeEnv._state_fp / _point_in_free_space (maze_env.py:270-277, 293-299), batched.
 *   states [n,2] f32|f64 (dtype = GMP_DTYPE_*); maps [P,15,15] u8 (1 = occupied);
 *   problem_of_state [n] i32 or NULL (= problem 0)
 *   free_out [n] u8; counted_out [n] u8 nullable: 1 iff the reference would have incremented
 *   collision_check_count (in-range states only, maze_env.py:272-276). */
int gmp_maze_state_fp(const void* states, int dtype, const uint8_t* maps, const int32_t* problem_of_state,
                      int64_t n, uint8_t* free_out, uint8_t* counted_out, void* stream);
/* Replaces MazeEnv._edge_fp + _iterative_check_segment (maze_env.py:301-325), batched.
 *   n_checks_out [n] i32 nullable: increments of collision_check_count the reference would make
 *   for this edge (DFS order, early exit). */
int gmp_maze_edge_fp(const void* a, const void* b, int dtype, const uint8_t* maps, const int32_t* problem_of_edge,
                     int64_t n, uint8_t* free_out, int32_t* n_checks_out, void* stream);
/* Same check for every edge of a packed batch of graphs, endpoints gathered from v (f32):
 * edge e of graph g tests v[node_ptr[g]+src] -> v[node_ptr[g]+dst] against maps[problem_of_graph[g]]
 * (what eval_gnn.py:215 does one edge at a time).  node_ptr / edge_ptr / problem_of_graph are DEVICE
 * arrays here ([B+1],[B+1],[B] i32; problem_of_graph NULL = graph index). */
int gmp_maze_edge_fp_graph(const float* v, const int64_t* edge_index, int64_t edge_row_stride, const int32_t* node_ptr,
                           const int32_t* edge_ptr, const int32_t* problem_of_graph, int64_t n_graphs,
                           int64_t n_edges_total, const uint8_t* maps, uint8_t* free_out, int32_t* n_checks_out,
                           void* stream);

/* ---- 3-D stick maze: MazeEnv(dim=3) (environment/maze_env.py:245-264, 279-291, 327-347) ------ */
/* states [n,3] = (x, y, theta) f32|f64.  _state_fp = _stick_in_free_space: valid state, both stick end points free, bisection of
 * the stick (float64 geometry whatever the state dtype).  _edge_fp: both states, then K = int(distance / 0.015) interpolated poses
 * (theta wrapped), each stick checked as a 2-D edge.  n_checks_out = collision_check_count increments, k_out = env.k afterwards
 * (both nullable). */
int gmp_maze3_state_fp(const void* states, int dtype, const uint8_t* maps, const int32_t* problem_of_state, int64_t n,
                       uint8_t* free_out, int32_t* n_checks_out, int32_t* k_out, void* stream);
int gmp_maze3_edge_fp(const void* a, const void* b, int dtype, const uint8_t* maps, const int32_t* problem_of_edge, int64_t n,
                      uint8_t* free_out, int32_t* n_checks_out, int32_t* k_out, void* stream);

/* ---- batched lazy tree search: the inner loop of explore() (eval_gnn.py:198-233), maze environments ------------ */
/* One CTA per problem replays the reference's search on the SPARSE logits of gmp_explorer_forward: masks of eval_gnn.py:198-202
 * (diagonal, explored columns, collided rows / columns = nodes >= n_free[g], the explored-edge list with the reference's
 * reshape(2,-1) quirk), then repeatedly: arg-max over the rows of the explored nodes (ties: earlier position in `explored`,
 * then lower column), env._edge_fp on that edge, grow the tree (and test in_goal_region) or zero the edge both ways.
 * Stops on success (status 1), when no candidate is left (status 2: the caller resamples, rebuilds the graph and calls again
 * with first_round = 0), or when a capacity is exceeded (status 3).
 *   v, node_ptr (DEVICE [B+1]), n_free (DEVICE [B]), edge_index / edge_ptr (DEVICE [B+1]), edge_logits: this round's packed batch
 *   goal [S,2] f64: env.goal_state per slot;  maps [P,15,15] u8;  problem_of_graph [B] (NULL = g);  slot_of_graph [B] (NULL = g):
 *   which row of the persistent search state graph g continues
 *   spec_k in [1,32]: edges checked per iteration (1 = the reference's one edge; more = speculative checks of the next-best
 *   candidates, remembered and committed later in the reference's order -- results are identical for every spec_k)
 *   persistent state, S rows: explored [S,cap_nodes], n_explored [S], prev [S,cap_nodes], explored_edges [S,cap_explored_edges]
 *   (the flat list of eval_gnn.py:184,214), n_explored_edges [S], n_checks [S] (collision_check_count increments of the search:
 *   edge checks + the in_goal_region state checks), n_spec_checks [S] (speculative checks never committed), status [S],
 *   path [S,cap_nodes] + path_len [S] (node ids from 0 to the goal-region node, eval_gnn.py:223-229), path_cost [S] f32 (nullable;
 *   path_cost(path), eval_gnn.py:53-58). */
int64_t gmp_tree_search_workspace_bytes(int64_t n_graphs, int64_t n_nodes_total, int64_t n_edges_total);
int gmp_maze_tree_search(const float* v, const int32_t* node_ptr, const int32_t* n_free, const int64_t* edge_index,
                         int64_t edge_row_stride, const int32_t* edge_ptr, const float* edge_logits, const double* goal,
                         const uint8_t* maps, const int32_t* problem_of_graph, const int32_t* slot_of_graph,
                         int64_t n_graphs, int64_t n_nodes_total, int64_t n_edges_total, int spec_k, int first_round,
                         int32_t* explored, int32_t* n_explored, int32_t* prev, int32_t* explored_edges,
                         int32_t* n_explored_edges, int32_t* n_checks, int32_t* n_spec_checks, int32_t* status,
                         int32_t* path, int32_t* path_len, float* path_cost, int32_t cap_nodes, int32_t cap_explored_edges,
                         void* workspace, int64_t workspace_bytes, void* stream);
/* The same search for the arm environments (KukaEnv / Kuka2Env / UR5Env / SnakeEnv): env._edge_fp = the arm model's edge check
 * on float32 states (kuka_env.py:389-411), in_goal_region = float64 distance to goal [S, dof] + one state check
 * (kuka_env.py:244-249).  Speculation (spec_k > 1) pays here: an edge check is K forward-kinematics passes. */
int gmp_arm_tree_search(int model, const float* v, const int32_t* node_ptr, const int32_t* n_free, const int64_t* edge_index,
                        int64_t edge_row_stride, const int32_t* edge_ptr, const float* edge_logits, const double* goal,
                        const double* boxes, const int32_t* box_ptr, const int32_t* problem_of_graph, double rrt_eps,
                        const int32_t* slot_of_graph, int64_t n_graphs, int64_t n_nodes_total, int64_t n_edges_total,
                        int spec_k, int first_round, int32_t* explored, int32_t* n_explored, int32_t* prev,
                        int32_t* explored_edges, int32_t* n_explored_edges, int32_t* n_checks, int32_t* n_spec_checks,
                        int32_t* status, int32_t* path, int32_t* path_len, float* path_cost, int32_t cap_nodes,
                        int32_t cap_explored_edges, void* workspace, int64_t workspace_bytes, void* stream);
/* The per-problem tuple the reference reduces at the end of eval_gnn (eval_gnn.py:120-134), as rows of 6 floats:
 * (first_problem_id + i, success, path_cost, n_checks, n_spec_checks, n_explored).  These rows are what the multi-GPU run
 * all-gathers (NCCL, host side). */
int gmp_search_result_rows(const int32_t* status, const float* path_cost, const int32_t* n_checks, const int32_t* n_spec_checks,
                           const int32_t* n_explored, int64_t n_problems, int32_t first_problem_id, float* rows_out, void* stream);

/* Batched env.sample_n_points(n, need_negative=True) (maze_env.py:85-100) with a counter-based RNG: slot s draws the fixed
 * sequence Philox4x32-10(key = seed, counter = (k, stream_of_slot[s])), k = first_draw[s], first_draw[s]+1, ... -> states
 * uniform in [-1,1)^2 (float64), and returns the prefix of that sequence up to its n_points-th free state: free_out
 * [S, n_points, 2], collided_out [S, cap_collided, 2] (the rejected draws, in draw order; n_collided_out may exceed
 * cap_collided, then the tail is dropped), n_draws_out [S] = draws consumed = collision_check_count increments.  A NEW stream:
 * parity with the reference's global NumPy stream is distributional only (the host mirror keeps the exact stream). */
int gmp_maze_sample_points(const uint8_t* maps, const int32_t* problem_of_slot, const int64_t* stream_of_slot,
                           const int64_t* first_draw, int64_t n_slots, int32_t n_points, int32_t cap_collided, uint64_t seed,
                           double* free_out, double* collided_out, int32_t* n_collided_out, int64_t* n_draws_out, void* stream);

/* Replaces proposed_path_smootherv2 (smoother.py:194-216) for a packed batch of 2-D maze paths: K = ceil(max ||old - new|| /
 * rrt_eps) rounds; per round every interior waypoint is steered at most rrt_eps toward its proposal and kept iff both adjacent
 * edges are collision free (against the already updated left neighbour); stops when the accepted waypoints have all reached
 * their proposals.  float32 arithmetic as NumPy evaluates it on the reference's float32 paths.
 *   old_path / new_path / path_out [P_total, 2] f32, path p owns rows path_ptr[p] .. path_ptr[p+1] (DEVICE [B+1]);
 *   n_checks_out [B]: collision_check_count increments;  n_rounds_out [B] (nullable);  path_cost_out [B] f32 (nullable):
 *   path_cost of the result (eval_gnn.py:53-58). */
int gmp_maze_steer_rounds(const float* old_path, const float* new_path, const int32_t* path_ptr, const uint8_t* maps,
                          const int32_t* problem_of_path, int64_t n_paths, double rrt_eps, float* path_out,
                          int32_t* n_checks_out, int32_t* n_rounds_out, float* path_cost_out, void* stream);

/* ---- arm collision: KukaEnv / Kuka2Env (environment/kuka_env.py, kuka_2arm_env.py) ------------- */
/* The reference queries PyBullet contact points; this library substitutes its own geometric model (DESIGN.md:
 * FK down the URDF joint chain, link hulls filled with inscribed spheres, boxes as AABBs) -- parity with PyBullet
 * is UNPINNED, parity with oracle/arm.c is bit-exact.  model: 0 = kuka7 (kuka_env.py, model_0.urdf), 1 = kuka14
 * (kuka_2arm_env.py: two arms at x = -0.5 / +0.5, arm-arm contacts), 2 = kuka13 (model_3.urdf).
 *   boxes [O_total, 6] f64 = (halfExtents[3], basePosition[3]) as stored in maze_files/kukas_*.pkl; problem p owns
 *   rows box_ptr[p] .. box_ptr[p+1] (DEVICE i32 array). */
int gmp_arm_model_count(void);
/* dof and joint limits (KukaEnv.pose_range, kuka_env.py:57-60) of a model; host outputs, nullable. */
int gmp_arm_model_info(int model, int32_t* dof_out, double* lower_h, double* upper_h);
/* Replaces _state_fp / _point_in_free_space (kuka_env.py:354-376).  counted_out: 1 iff collision_check_count
 * would be incremented (every state within joint limits, both outcomes). */
int gmp_arm_state_fp(int model, const void* states, int dtype, const double* boxes, const int32_t* box_ptr,
                     const int32_t* problem_of_state, int64_t n, uint8_t* free_out, uint8_t* counted_out, void* stream);
/* Replaces _edge_fp (kuka_env.py:389-411): endpoints valid + free, then K = int(||b-a|| / rrt_eps) interpolated
 * states k = 0..K-1, in the input dtype.  n_checks_out = collision_check_count increments. */
int gmp_arm_edge_fp(int model, const void* a, const void* b, int dtype, const double* boxes, const int32_t* box_ptr,
                    const int32_t* problem_of_edge, int64_t n, double rrt_eps, uint8_t* free_out, int32_t* n_checks_out,
                    void* stream);
/* Same for every edge of a packed batch of graphs (endpoints gathered from v f32, as eval_gnn.py:215 does one at a
 * time); node_ptr / edge_ptr / problem_of_graph are DEVICE arrays (problem_of_graph NULL = graph index). */
int gmp_arm_edge_fp_graph(int model, const float* v, const int64_t* edge_index, int64_t edge_row_stride,
                          const int32_t* node_ptr, const int32_t* edge_ptr, const int32_t* problem_of_graph,
                          int64_t n_graphs, int64_t n_edges_total, const double* boxes, const int32_t* box_ptr,
                          double rrt_eps, uint8_t* free_out, int32_t* n_checks_out, void* stream);
/* The same result (booleans and check counts, bit for bit) with the endpoint checks evaluated once per NODE instead of once per
 * incident edge: every edge check starts with both endpoint states and repeats the start state at k = 0 (kuka_env.py:394-409),
 * and in a k-NN graph each node is an endpoint of ~2k edges.  node_flags_ws: n_nodes_total bytes of device scratch. */
int gmp_arm_edge_fp_graph_cached(int model, const float* v, int64_t n_nodes_total, const int64_t* edge_index,
                                 int64_t edge_row_stride, const int32_t* node_ptr, const int32_t* edge_ptr,
                                 const int32_t* problem_of_graph, int64_t n_graphs, int64_t n_edges_total, const double* boxes,
                                 const int32_t* box_ptr, double rrt_eps, uint8_t* node_flags_ws, uint8_t* free_out,
                                 int32_t* n_checks_out, void* stream);

/* The same result again (booleans and check counts bit-identical to gmp_arm_edge_fp_graph / oracle/arm.c), reached faster:
 * every interpolated state goes through an fp32 evaluation of the model that decides it only when no sphere test is within
 * 2.5e-4 m of touching (the fp32 forward kinematics are within ~3e-5 m of the fp64 ones); the undecided states (~1 %) are
 * listed and decided by the exact fp64 test in a second kernel; lanes pull edges from a per-CTA queue instead of owning one.
 * max_boxes_per_problem: the largest box count of any problem referenced (> 160 routes to the fp64 form).
 * workspace: gmp_arm_edge_graph_workspace_bytes(n_nodes_total, n_edges_total) bytes of device scratch. */
int64_t gmp_arm_edge_graph_workspace_bytes(int64_t n_nodes_total, int64_t n_edges_total);
int gmp_arm_edge_fp_graph_fast(int model, const float* v, int64_t n_nodes_total, const int64_t* edge_index,
                               int64_t edge_row_stride, const int32_t* node_ptr, const int32_t* edge_ptr,
                               const int32_t* problem_of_graph, int64_t n_graphs, int64_t n_edges_total, const double* boxes,
                               const int32_t* box_ptr, int max_boxes_per_problem, double rrt_eps, void* workspace,
                               int64_t workspace_bytes, uint8_t* free_out, int32_t* n_checks_out, void* stream);

/* ---- smoother: ModelSmoother (model_smoother.py:46-142) --------------------------------------- */
int gmp_smoother_init(gmp_handle* h, int config_size /*c*/, int embed_size /*128*/);
/* load_state_dict (eval_gnn.py:104) by reference tensor name, e.g. "node_code.1.running_mean"; dead tensors ignored. */
int gmp_smoother_set_tensor(gmp_handle* h, const char* name, const float* data_h, int64_t numel);
int gmp_smoother_finalize(gmp_handle* h);
int64_t gmp_smoother_workspace_bytes(const gmp_handle* h, int64_t n_problems, int64_t n_path_total, int64_t n_sample_total,
                                     int64_t n_edges_total);
/* Replaces ModelSmoother.forward (model_smoother.py:104-142) for a packed batch of problems.
 *   path     [P_total, c] f32   waypoints, problem g owns rows path_ptr_h[g] .. path_ptr_h[g+1]     (`path`)
 *   samples  [S_total, c] f32   cat(free, collided) per problem, rows sample_ptr_h[g] ..; the first n_free_h[g]
 *                               rows of a problem are `free`, the rest `collided`                  (`free`, `collided`)
 *   edge_index [2, E_total] i64 caller's edges, node ids LOCAL to the problem's cat(path, free, collided)
 *                               (smoother.py:238-241 passes the path chain + self loops)            (`edge_index`)
 *   scale    ModelSmoother.scale (model_smoother.py:55,118-120,142);  loop: model_smoother.py:104 (model_smooth uses 1)
 *   path_out [P_total, c] f32   new path; rows 0 and P_g-1 of each problem keep their input value (model_smoother.py:139)
 * The caller's `path` is not modified (the reference divides first, model_smoother.py:118). */
int gmp_smoother_forward(gmp_handle* h, int64_t n_problems, const float* path, const float* samples,
                         const int64_t* edge_index, int64_t edge_row_stride, const int32_t* path_ptr_h,
                         const int32_t* sample_ptr_h, const int32_t* n_free_h, const int32_t* edge_ptr_h, float scale,
                         int loop, float* path_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- per-problem result rows: the final reduction of eval_gnn (eval_gnn.py:120-134) ----------- */
/* One row of 4 floats per graph: (first_problem_id + g, E_g, #edges with edge_free != 0, max edge logit).
 * edge_ptr is a DEVICE array [B+1]; edge_free is nullable.  These rows are the only payload of the multi-GPU
 * all-gather (torch.distributed / NCCL on the host side). */
int gmp_result_rows(const float* edge_logits, const uint8_t* edge_free, const int32_t* edge_ptr, int64_t n_graphs,
                    int32_t first_problem_id, float* rows_out, void* stream);

/* Local node ids of a packed edge_index as int16 (bits = 16: every N_g <= 32767) or int32 (bits = 32) for the
 * device->host read-back of the batched path: 4 B/edge instead of the 16 B/edge of torch_geometric's int64 layout
 * (eval_gnn.py:164).  out [2, out_row_stride] of that type. */
int gmp_edge_index_narrow(const int64_t* edge_index, int64_t edge_row_stride, int64_t n_edges, int bits, void* out,
                          int64_t out_row_stride, void* stream);

/* Small control data (<= 1 MiB, a multiple of 4 bytes) device -> pinned host memory that is mapped into the device's address
 * space (cudaHostAlloc / torch pin_memory under UVA), written by a kernel instead of a copy engine: the per-batch edge_ptr
 * [B+1] the host needs before it can enqueue the forward (eval_gnn.py:164 builds the graph on the host; here it is the one
 * host synchronisation of a batch) must not wait behind the previous batch's 100 MB read-back on the same copy engine. */
int gmp_post_to_host(const void* src_device, int64_t nbytes, void* dst_mapped_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GNNMP_H_ */
