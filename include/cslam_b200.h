/*
 * cslam_b200.h — C ABI of libcslam_b200.so, the B200 (sm_100a) loop-closure
 * front end for Swarm-SLAM's `cslam` package.
 *
 * The reference (lajoiepy/cslam) has NO FFI on this path: its boundary is the
 * Python class API.  Every entry point below therefore cites the reference
 * Python method whose arithmetic it replaces; the Python classes in
 * `cslam_b200/` keep the reference names/signatures and call these symbols
 * through ctypes (see INTEGRATION.md for the stub a cslam maintainer adds).
 *
 * Conventions
 *   - plain C types only: pointers, sizes, ints.  No torch / C++ types.
 *   - every function returns an int status (CSLAM_OK == 0, <0 = error) unless
 *     documented otherwise; `cslam_last_error()` returns a thread-local
 *     human-readable message for the last failure.  Nothing throws across
 *     the ABI.
 *   - "host" variants take host pointers and perform the H2D / D2H copies
 *     themselves; "device" variants take device pointers valid on the
 *     handle's device and are asynchronous on the given stream.  `stream`
 *     is a `cudaStream_t` passed as void*; NULL is CUDA's legacy default
 *     stream (which is what torch's default stream is), exactly as in the
 *     CUDA runtime API — never a private stream of the library.  The inputs
 *     must be ready in stream order on that stream; outputs are ready in
 *     stream order on it.  Operations on one handle issued on different
 *     streams (the caller's, or the handle's own stream that the "host"
 *     variants use) execute in call order: every entry point makes its
 *     stream wait for the handle's previous operation.
 *   - handles are not re-entrant (the reference is single-threaded:
 *     cslam/loop_closure_detection_node.py:106-110).
 *   - there is NO CPU fallback: without a CUDA device every compute call
 *     returns CSLAM_ERR_CUDA.
 */
#ifndef CSLAM_B200_H
#define CSLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSLAM_OK 0
#define CSLAM_ERR_INVALID -1   /* bad argument */
#define CSLAM_ERR_CUDA -2      /* CUDA runtime / driver failure, or no device */
#define CSLAM_ERR_OOM -3       /* allocation failure */
#define CSLAM_ERR_LIMIT -4     /* request exceeds a documented limit */
#define CSLAM_ERR_SINGULAR -5  /* Laplacian singular / graph disconnected
                                  (reference: SuperLU RuntimeError swallowed by
                                  cslam/algebraic_connectivity_maximization.py:448-466) */
#define CSLAM_ERR_NOCONV -6    /* iterative solver did not converge */

#define CSLAM_DTYPE_F32 0
#define CSLAM_DTYPE_F64 1

/* ---- library ---------------------------------------------------------- */
const char* cslam_version(void);
const char* cslam_last_error(void);
/* number of visible CUDA devices (0 if none / driver missing) */
int cslam_device_count(void);
/* total kernels launched by this library in this process (bench `gpu_launches`) */
int64_t cslam_launch_count(void);

/* ---- A6: cosine nearest-neighbour matching --------------------------- *
 * Replaces cslam/nns_matching.py:6-76 (NearestNeighborsMatching).
 * Pool rows are stored float32 exactly like the reference's `.data`
 * (nns_matching.py:21,39); scoring of the returned matches is
 *   sim = 1 - clip(1 - q.x / sqrt((q.q) (x.x)), 0, 2)          (float64)
 * i.e. scipy.spatial.distance.cosine as called at nns_matching.py:58,
 * with x.x accumulated in float32 as np.dot(float32,float32) does.
 */
typedef struct cslam_nns cslam_nns_t;

/* dim > 0; device = CUDA ordinal.  (NearestNeighborsMatching.__init__, :10-21) */
int cslam_nns_create(int dim, int device, cslam_nns_t** out);
int cslam_nns_destroy(cslam_nns_t* h);
/* Append `count` rows ([count, dim] row-major, dtype F32 or F64; F64 is
 * rounded to float32 on store like `self.data[self.n] = vector`, :39).
 * Row ids are assigned consecutively from the current size.  (add_item, :23-40) */
int cslam_nns_add_host(cslam_nns_t* h, const void* rows, int dtype, int64_t count);
/* Same, rows already on the device as float32. */
int cslam_nns_add_device(cslam_nns_t* h, const float* d_rows, int64_t count, void* stream);
int64_t cslam_nns_size(cslam_nns_t* h);
int cslam_nns_dim(cslam_nns_t* h);
/* Copy rows [start, start+count) of the float32 pool to host (`.data`). */
int cslam_nns_read_rows(cslam_nns_t* h, int64_t start, int64_t count, float* out);

/* Top-k search for `nq` queries ([nq, dim], F32 or F64).  (search, :42-61)
 * out_idx  [nq, k] int32 row ids, best first; unused slots = -1
 * out_sims [nq, k] float64 similarities; unused slots = NaN
 * Each query returns min(k, size) matches.  A pool row of norm 0 (similarity NaN in the
 * reference, ranked first by its argsort()[::-1]) scores 0 here and is not returned.
 * Exactly equal similarities are
 * ordered by DESCENDING row id (the reference's np.argsort(sim)[::-1],
 * nns_matching.py:60, leaves the order of equal keys unspecified).
 * 1 <= k <= 1024.
 * out_info (optional, may be NULL) [4] int64:
 *   [0] queries answered by the tensor-core coarse pass + exact re-rank
 *   [1] total pool rows that were exactly re-ranked (all queries)
 *   [2] queries re-run through the exact fp64 scan kernel
 *   [3] kernels launched by this call                                   */
int cslam_nns_search_host(cslam_nns_t* h, const void* queries, int dtype, int nq, int k,
                          int32_t* out_idx, double* out_sims, int64_t* out_info);
/* Device variant: queries/out_* are device pointers; async on `stream`
 * except for the (rare) escalation check, which synchronises the stream. */
int cslam_nns_search_device(cslam_nns_t* h, const void* d_queries, int dtype, int nq, int k,
                            int32_t* d_out_idx, double* d_out_sims, void* stream,
                            int64_t* out_info);
/* Tuning / test hooks.  mode: 0 = auto (tcgen05 coarse pass + exact re-rank,
 * escalating to the exact scan only when the error-bound check fails),
 * 1 = force the exact fp64 scan kernel (validation path, still GPU). */
int cslam_nns_set_mode(cslam_nns_t* h, int mode);
/* Rows sampled by the threshold pass (default 32768, multiple of 256); pools of at
 * most this many rows are scored exhaustively without a threshold. */
int cslam_nns_set_sample_rows(cslam_nns_t* h, int sample_rows);
/* Time (ms, CUDA events on the handle's stream) of the coarse kernel launches
 * of the last search call and how many there were. */
int cslam_nns_last_timing(cslam_nns_t* h, float* coarse_ms, int* coarse_launches,
                          float* total_ms);

/* ---- A11-A15: MAC sparsification (Frank-Wolfe on the Fiedler value) ------ *
 * Replaces cslam/mac/mac.py:19-233 (class MAC) and the Laplacian assembly of
 * cslam/mac/utils.py:47-126.  All arithmetic is float64.
 * Node ids are the rekeyed ids of
 * cslam/algebraic_connectivity_maximization.py:312-362 (0 <= i < num_poses).
 * A Laplacian whose graph (fixed + active candidate edges) is disconnected is
 * reported as CSLAM_ERR_SINGULAR, the condition under which the reference's
 * SuperLU factorisation raises (mac.py:52-58 -> networkx _LUSolver).          */
typedef struct cslam_mac cslam_mac_t;

/* MAC(fixed_measurements, candidate_measurements, num_poses)  (mac.py:21-33).
 * Edge arrays are host pointers (i, j, weight); copied. */
int cslam_mac_create(int num_poses, int64_t n_fixed, const int32_t* fixed_i,
                     const int32_t* fixed_j, const double* fixed_w, int64_t n_cand,
                     const int32_t* cand_i, const int32_t* cand_j, const double* cand_w,
                     int device, cslam_mac_t** out);
int cslam_mac_destroy(cslam_mac_t* h);
/* Eigen-solver options: relative residual tolerance ||Lx - theta x||_1 / ||L||_inf
 * (default 1e-10; the reference stops at 1e-8, mac.py:35), LOBPCG block size (1 or 2,
 * default 2) and iteration cap. */
int cslam_mac_set_options(cslam_mac_t* h, double tol, int block_size, int max_lobpcg_iters);
/* evaluate_fiedler_pair(w)  (mac.py:79-97, 61-77, 35-59): lambda_2 and the unit-norm
 * Fiedler vector [num_poses] (sign arbitrary, as in the reference) of
 * L(w) = L_fixed + sum_{w_e > 1e-10} w_e * weight_e * (e_i - e_j)(e_i - e_j)^T.
 * w: host [n_cand].  vec_out (host, nullable), iters_out (nullable). */
int cslam_mac_fiedler(cslam_mac_t* h, const double* w, double* lambda2, double* vec_out,
                      int* iters_out);
/* grad_from_fiedler(fiedler_vec)  (mac.py:112-130): grad_e = weight_e * (v_i - v_j)^2.
 * host in [num_poses], host out [n_cand]. */
int cslam_mac_grad(cslam_mac_t* h, const double* fiedler_vec, double* grad_out);
/* fw_subset(w_init, k, max_iters, duality_gap_tol)  (mac.py:191-233).
 * rounded_out [n_cand] 0/1 (round_solution_tiebreaker, mac.py:168-189), w_out [n_cand]
 * the unrounded iterate, *u_out the dual upper bound, *iters_out Frank-Wolfe iterations
 * run.  Optional traces: trace_sel [max_iters, k] the ascending edge ids of every
 * direction s_i (round_solution, mac.py:132-147; exact ties at the k-th value are taken
 * in index order), trace_f [max_iters] the objective values. */
int cslam_mac_fw_subset(cslam_mac_t* h, const double* w_init, int k, int max_iters,
                        double duality_gap_tol, double* rounded_out, double* w_out,
                        double* u_out, int* iters_out, int32_t* trace_sel, double* trace_f);
/* The same with a SPARSE start vector and sparse results - what select_candidates needs
 * (algebraic_connectivity_maximization.py:448-466,519-533): the start vector is the greedy 0/1
 * vector with k ones (:205-218) and only the k selected edges are used afterwards, so no dense
 * [n_cand] vector crosses the boundary.
 *   init_idx/init_val [n_init]  the non-zero entries of w_init (distinct ids)
 *   sel_out [k]                 ascending ids of the rounded selection (round_solution_tiebreaker)
 *   sup_idx_out/sup_val_out     (nullable) the non-zero entries of the unrounded iterate w in
 *                               ascending id order, *n_sup_out of them; sup_capacity must be at
 *                               least min(n_cand, n_init + k * max(max_iters, 1))
 * The support of w and the Laplacian of the active candidates are maintained and assembled on the
 * device (mac.py:61-77, mac/utils.py:86-126 rebuild them from Python lists per iteration). */
int cslam_mac_fw_subset_sparse(cslam_mac_t* h, int64_t n_init, const int32_t* init_idx,
                               const double* init_val, int k, int max_iters, double duality_gap_tol,
                               int32_t* sel_out, int64_t sup_capacity, int32_t* sup_idx_out,
                               double* sup_val_out, int64_t* n_sup_out, double* u_out, int* iters_out,
                               int32_t* trace_sel, double* trace_f);
/* Totals since creation: LOBPCG iterations, SpMV columns applied, and whether the last
 * solve had to fall back from the tridiagonal to the diagonal preconditioner. */
int cslam_mac_stats(cslam_mac_t* h, int64_t* lobpcg_iters, int64_t* spmv_columns,
                    int* jacobi_fallback);
/* Totals since creation for the persistent eigen-solver kernel (k_lobpcg_persist): summed
 * CUDA-event durations of its launches on the handle's stream, launches, LOBPCG iterations and
 * the algorithmic bytes of those iterations (one SpMM each: nnz*12 + n*4 + m*n*16, SURVEY.md
 * section 8d).  Used by bench.py for the roofline entry. */
int cslam_mac_solver_timing(cslam_mac_t* h, double* kernel_ms, int64_t* launches,
                            int64_t* iterations, int64_t* algorithmic_bytes);
/* find_fiedler_pair(L)  (mac.py:35-59) for a caller-assembled CSR Laplacian (host
 * arrays, int32 indices, float64 data; the diagonal entries are ignored and rebuilt as
 * minus the row sums of the off-diagonals). */
int cslam_fiedler_csr(int n, const int32_t* indptr, const int32_t* indices, const double* data,
                      double tol, int block_size, int device, double* lambda2, double* vec_out,
                      int* iters_out);

/* ---- (f3) candidate graph: hash index of the edge keys on the device ---------------------- *
 * Replaces the Python dict of cslam/algebraic_connectivity_maximization.py:58,150,174 for BULK
 * maintenance of the candidate graph (add_match :559-572, remove_candidate_edges :178-190,
 * candidate_edges_to_fixed :192-203): packed 64-bit edge key -> int32 slot of the columnar
 * candidate table.  Open addressing in HBM, one kernel per batch; host arrays in and out.
 * Keys 0xFFFFFFFFFFFFFFFE/F are reserved.                                                    */
typedef struct cslam_keymap cslam_keymap_t;
int cslam_keymap_create(int64_t capacity_hint, int device, cslam_keymap_t** out);
int cslam_keymap_destroy(cslam_keymap_t* h);
int64_t cslam_keymap_size(cslam_keymap_t* h);
/* values_out[t] = slot stored for keys[t], or -1 */
int cslam_keymap_lookup(cslam_keymap_t* h, const uint64_t* keys, int64_t n, int32_t* values_out);
/* insert or overwrite; the keys of one call must be distinct */
int cslam_keymap_insert(cslam_keymap_t* h, const uint64_t* keys, const int32_t* values, int64_t n);
/* erase (missing keys are ignored); values_out (nullable) gets the erased slot or -1 */
int cslam_keymap_erase(cslam_keymap_t* h, const uint64_t* keys, int64_t n, int32_t* values_out);

/* ---- (e) multi-robot round: filters on the all-gathered per-shard top-k ------------------ *
 * Device pointers; async on `stream`.  Replace the host loops of
 * cslam/loop_closure_sparse_matching.py:45-53,62-72 (similarity gate on the best match of a
 * keyframe in every OTHER robot's pool) and :74-92 (intra-robot matches among the rows that
 * were in the pool before the keyframe was added).
 * cslam_swarm_hits: g_kf/g_sims [R pools][R*B queries][kx] as gathered (column 0 = best match,
 *   keyframe id -1 / similarity NaN = empty pool), all_ids [R][B] keyframe ids of the queries.
 *   out[0] = number of hits n, out[1 + 5 t ..] = (query robot, query keyframe, pool robot,
 *   matched keyframe, similarity) of hit t in the reference's order (query robot, keyframe, pool
 *   robot); at most `cap` hits are written.
 * cslam_swarm_intra: idx/kf/sims [B][k_search] search results of the rank's own B keyframes
 *   (pool rows, keyframe ids, similarities; best first) in the pool that already contains them;
 *   out [B][1 + 2 k_keep] = (count, kept keyframe ids, kept similarities). */
int cslam_swarm_hits(int R, int B, int kx, const int64_t* d_g_kf, const double* d_g_sims,
                     const int64_t* d_all_ids, double threshold, double* d_out, int cap, void* stream);
int cslam_swarm_intra(int B, int k_search, int k_keep, int64_t rows_before, const int64_t* d_idx,
                      const int64_t* d_kf, const double* d_sims, double* d_out, void* stream);

/* Test hook: the small generalised eigenproblem GA y = theta GB y (s x s, row-major with a
 * leading dimension of 6; 1 <= m <= 2 smallest pairs) that the eigen-solver solves once per
 * iteration, by one warp, `reps` times.  impl 2 = the two-stage solve of the eigen-solver (per column
 * the 3 x 3 problem on {x, w, p}, then the m x m problem on the results; sweeps > 0: Rayleigh-quotient
 * iteration for the 3 x 3 problems with the Jacobi solve as its fallback, sweeps < 0: Jacobi only with
 * |sweeps| sweeps), 1 = register-resident 6 x 6 solver (entry-per-lane Jacobi), 0 = shared-memory
 * 6 x 6 solver.  c_out [6][2] (GB-orthonormal vectors), theta_out [2],
 * *ok_out 0 when GB is not positive definite, cycles_out (nullable) [6]: SM cycles per solve:
 * total, then set-up, Cholesky, triangular transforms, Jacobi sweeps, back substitution. */
int cslam_debug_rayleigh_ritz(const double* ga, const double* gb, int s, int m, int impl, int sweeps,
                              int reps, int device, double* c_out, double* theta_out, int* ok_out,
                              int64_t* cycles_out);

/* Test hook: cycles per barrier of the solver's grid barrier on an otherwise empty co-resident grid
 * (`ctas` x `threads`, `stores` global stores per thread before each barrier); `variant` selects the
 * memory-ordering recipe (csrc/mac.cu grid_barrier). */
int cslam_debug_grid_barrier(int ctas, int threads, int reps, int stores, int variant, int device,
                             int64_t* cycles_per_barrier);

/* ---- A1-A5: descriptor extraction around the PyTorch backbone ------------- *
 * All pointers are DEVICE pointers (float32 unless noted); `stream` as above. */

/* A1 preprocessing: transforms.Compose([CenterCrop(crop), Resize(out, interpolation=3),
 * ToTensor(), Normalize(IMAGENET mean/std)]) of cslam/vpr/netvlad.py:202-208 and
 * cslam/vpr/cosplace.py:73-79, for uint8 HWC RGB keyframes of a fixed size.  The resize
 * reproduces Pillow's fixed-point antialiased bicubic bit for bit. */
typedef struct cslam_preproc cslam_preproc_t;
int cslam_preproc_create(int in_h, int in_w, int crop, int out_size, int device,
                         cslam_preproc_t** out);
int cslam_preproc_destroy(cslam_preproc_t* h);
/* d_images uint8 [batch, in_h, in_w, 3] -> d_out float32 [batch, 3, out, out] */
int cslam_preproc_run(cslam_preproc_t* h, const uint8_t* d_images, int batch, float* d_out,
                      void* stream);

/* A3 NetVLADLayer.forward (cslam/vpr/netvlad.py:94-130), vladv2=False:
 * d_x [batch, 512, locations] (NCHW feature map, locations = H*W <= 224),
 * d_conv_w [64, 512] (pool.conv.weight), d_centroids [64, 512] (pool.centroids)
 * -> d_out [batch, 64*512], cluster-major, intra-normalised and L2-normalised. */
int cslam_vlad_forward(const float* d_x, int batch, int channels, int locations,
                       const float* d_conv_w, const float* d_centroids, int clusters,
                       float* d_out, void* stream);

/* A4 sklearn PCA.transform + preprocessing.normalize (cslam/vpr/netvlad.py:234-237):
 * out = l2_normalise_rows((x @ W^T - bias) * scale), W = components_ [dout, din],
 * bias = mean_ @ W^T [dout], scale = 1/sqrt(explained_variance_) [dout] when whiten
 * (NULL otherwise).  d_work: cslam_pca_workspace_floats(min(batch,64), dout) floats. */
int64_t cslam_pca_workspace_floats(int batch, int dout);
int cslam_pca_project_l2(const float* d_x, int batch, int din, const float* d_w,
                         const float* d_bias, const float* d_scale, int dout, float* d_out,
                         float* d_work, void* stream);

/* A5 CosPlace aggregation (cslam/vpr/cosplace_utils/network.py:23-29, layers.py:8-36):
 * L2Norm -> GeM(p, eps) -> Flatten -> Linear(channels, dout) -> L2Norm.
 * d_x [batch, channels, locations], d_fc_w [dout, channels], d_fc_b [dout] -> d_out [batch, dout] */
int cslam_gem_head_forward(const float* d_x, int batch, int channels, int locations, float p,
                           float eps, const float* d_fc_w, const float* d_fc_b, int dout,
                           float* d_out, void* stream);

/* ---- (f4) lidar place recognition: Scan Context matching ------------------- *
 * Replaces cslam/lidar_pr/scancontext_matching.py:6-104 (ScanContextMatching) with the
 * helpers cslam/lidar_pr/scancontext_utils.py:78-79 (sc2rk) and :81-113 (distance_sc).
 * Descriptors are [rings, sectors] row-major (the reference's `descriptor.reshape(shape)`),
 * float64 arithmetic throughout.  Host pointers; the library owns the pool in HBM. */
typedef struct cslam_sc cslam_sc_t;

/* ScanContextMatching(shape=[rings, sectors], num_candidates)  (:10-22).
 * 1 <= num_candidates <= 16. */
int cslam_sc_create(int rings, int sectors, int num_candidates, int device, cslam_sc_t** out);
int cslam_sc_destroy(cslam_sc_t* h);
int64_t cslam_sc_size(cslam_sc_t* h);
/* rows allocated: 1000, doubling (:18-19,33-37) */
int64_t cslam_sc_capacity(cslam_sc_t* h);
/* add_item for `count` descriptors ([count, rings*sectors], F32 or F64): stores the scan
 * context and its ring key (row means, numpy's summation order)  (:24-46). */
int cslam_sc_add_host(cslam_sc_t* h, const void* descriptors, int dtype, int64_t count);
/* `.scancontexts[start:start+count]` ([count, rings, sectors]) and `.ringkeys[...]`
 * ([count, rings]); either pointer may be NULL. */
int cslam_sc_read(cslam_sc_t* h, int64_t start, int64_t count, double* out_scancontexts,
                  double* out_ringkeys);
/* search for `nq` queries  (:48-89).  Per query: the num_candidates entries with the nearest
 * ring keys (Euclidean; equal distances in ascending row order -- scipy's KDTree leaves that
 * order unspecified), the column-shift distance to each, and the first candidate with the
 * smallest distance below 1.
 *   out_row        [nq] pool row of the match, -1 when no candidate is closer than 1 (the
 *                  reference then answers item 0 with similarity 0, :81-84)
 *   out_similarity [nq] 1 - distance (0.0 when out_row is -1)
 *   out_yaw_shift  [nq] nullable: the best column shift, 1..sectors (`yaw_diff`, :110)
 *   out_candidates [nq, num_candidates] nullable: candidate rows, nearest ring key first, -1 padded
 *   out_candidate_dist [nq, num_candidates] nullable: their column-shift distances
 * The pool must not be empty. */
int cslam_sc_search_host(cslam_sc_t* h, const void* queries, int dtype, int nq, int32_t* out_row,
                         double* out_similarity, int32_t* out_yaw_shift, int32_t* out_candidates,
                         double* out_candidate_dist);
/* CUDA-event times (ms) of the last search: ring-key kNN kernels, distance + pick kernels. */
int cslam_sc_last_timing(cslam_sc_t* h, float* knn_ms, float* distance_ms);

#ifdef __cplusplus
}
#endif
#endif /* CSLAM_B200_H */
