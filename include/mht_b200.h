/*
 * mht_b200.h -- C ABI of libmht_b200.so: the B200 (sm_100a) implementation of pyMHT's per-scan hot
 * path.  Plain pointers and sizes only; no torch / C++ types.  The reference (erikliland/pyMHT) is
 * pure Python with no FFI of its own, so each entry point below names the reference *method* whose
 * body it replaces (file:line into the reference tree) -- the ctypes stubs a maintainer would add
 * are shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns an int status: MHT_OK (0) or a negative MHT_E_* code; never throws;
 *     mht_last_error() returns a human-readable message for the last failure on the calling thread.
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch.Tensor.data_ptr());
 *     h_* are HOST pointers.  Nothing is allocated inside except by mht_forest_create().
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous on
 *     that stream unless documented otherwise.
 *   - dtype contract = what the reference holds: states, innovations, NIS, NLLR and
 *     cumulative scores float64; Phi,Q,C,R and the covariance chain (P_bar,S,S^-1,K,P_hat) float32,
 *     each product evaluated as an ascending-k fused-multiply-add chain (bit-identical to the
 *     reference's NumPy/OpenBLAS sgemm on the authoring host).
 */
#ifndef MHT_B200_H
#define MHT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHT_OK 0
#define MHT_E_INVALID (-1)    /* bad argument                                             */
#define MHT_E_CUDA (-2)       /* CUDA runtime error (see mht_last_error)                  */
#define MHT_E_CAPACITY (-3)   /* an output buffer / forest level is too small             */
#define MHT_E_NODEVICE (-4)   /* no sm_100 device visible: there is NO cpu fallback       */
#define MHT_E_NOTOPTIMAL (-5) /* association solved but optimality not certified          */

#define MHT_MAX_WINDOW 16     /* max N-scan window + 1 (path planes per hypothesis)       */

/* Linear-Gaussian model constants, row-major, float32 like reference pymht/models/pv.py:7-34. */
typedef struct mht_model {
    float A[16];       /* Phi(T)      pv.py:27-32  */
    float Q[16];       /* Q(T)        pv.py:17-23  */
    float C[8];        /* C_RADAR     pv.py:7-8    */
    float R[4];        /* R_RADAR()   pv.py:25-26  */
    double eta2;       /* gate        tracker.py:110 */
    double lambda_ex;  /* lambda_phi + lambda_nu  tracker.py:107 */
} mht_model;

int mht_version(void);
const char *mht_last_error(void);
/* number of visible sm_100 devices (0 => every compute entry point returns MHT_E_NODEVICE). */
int mht_device_count(void);
/* kernels this library has launched so far in this process (every launch site counts itself). */
int64_t mht_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Stateless operator: replaces the body of Tracker._processLeafNodes for ANY batch of leaves
 * (reference pymht/tracker.py:383-398,804-889 = kalman.predict/precalc/z_tilde/
 * normalizedInnovationSquared/numpyFilter/nllr, pymht/utils/kalman.py:14-101) plus the zero
 * hypothesis score of Target.createZeroHypothesis (pymht/pyTarget.py:319-328).
 *
 * in : d_x0[L,4] f64, d_P0[L,16] f32, d_Pd[L] f64, d_cnllr[L] f64, d_z[M,2] f64
 * out: d_x_bar[L,4] f64, d_P_bar[L,16] f32, d_P_hat[L,16] f32, d_miss_cnllr[L] f64,
 *      d_pair_off[L+1] i32 (pairs of leaf l are [off[l], off[l+1]), ascending measurement index),
 *      d_pair_meas[cap] i32 (0-based index), d_pair_cnllr[cap] f64 (parent cNLLR + NLLR),
 *      d_pair_xhat[cap,4] f64, d_meas_used[M] u8 (feeds tracker.py:331-332,266)
 * d_work: >= mht_gate_batch_workspace(L, M) bytes.  Returns MHT_E_CAPACITY if the pair count
 * exceeds `cap` (d_pair_off[L] still holds the required count).  Synchronises the stream once
 * (it has to read the pair count).
 * ---------------------------------------------------------------------------------------------- */
int64_t mht_gate_batch_workspace(int64_t L, int64_t M);
int mht_gate_batch(const mht_model *model, int64_t L, int64_t M, const double *d_x0, const float *d_P0,
                   const double *d_Pd, const double *d_cnllr, const double *d_z, double *d_x_bar,
                   float *d_P_bar, float *d_P_hat, double *d_miss_cnllr, int32_t *d_pair_off,
                   int32_t *d_pair_meas, double *d_pair_cnllr, double *d_pair_xhat, int64_t cap,
                   uint8_t *d_meas_used, void *d_work, void *stream);

/* Same operator with HOST buffers (copies in, runs, copies out, synchronous): the form a
 * reference-side ctypes stub in Tracker._processLeafNodes would call.  On MHT_E_CAPACITY
 * h_pair_off[L] holds the required pair count. */
int mht_gate_batch_host(const mht_model *model, int64_t L, int64_t M, const double *h_x0, const float *h_P0,
                        const double *h_Pd, const double *h_cnllr, const double *h_z, double *h_x_bar,
                        float *h_P_bar, float *h_P_hat, double *h_miss_cnllr, int32_t *h_pair_off,
                        int32_t *h_pair_meas, double *h_pair_cnllr, double *h_pair_xhat, int64_t cap,
                        uint8_t *h_meas_used);

/* ------------------------------------------------------------------------------------------------
 * Stateless operators on an explicit column list (one column = one leaf hypothesis).
 *
 * mht_cluster: Tracker._findClustersFromSets (tracker.py:961-974).  Columns carry their tree and
 * up to `width` measurement-row ids (row < 0 = unused slot).  out d_cluster_of_tree[T] = smallest
 * tree index of the tree's connected component.
 *
 * mht_assoc_solve: the cluster loop of Tracker.addMeasurementList (tracker.py:228-236) =
 * Target._selectBestHypothesis (pyTarget.py:446-459) for every singleton cluster and
 * Tracker._solveOptimumAssociation/_solveBLP_OR_TOOLS (tracker.py:979-1217) for the rest, all
 * clusters at once:   min sum c_j tau_j ; each tree exactly one column ; each row at most once.
 * Columns of a tree must be contiguous (col_tree non-decreasing).  Lagrangian dual ascent on the
 * row constraints + reduced-cost fixing + exact search on the surviving columns (enumeration for small
 * components, best-first Lagrangian branch & bound for the rest; wall-clock budget MHT_BB_MS, default 10 s).
 * out d_selected_col[T] (column index per tree, -1 for trees without columns),
 *     h_info[8] = {lower bound, objective, n_candidate_cols, n_components, bb_nodes, dual_iters,
 *                  certified(1/0), max_component_trees}.
 * Returns MHT_E_NOTOPTIMAL if the search budget ran out (d_selected_col is still feasible).
 * ---------------------------------------------------------------------------------------------- */
int64_t mht_assoc_workspace(int64_t n_cols, int64_t n_trees, int64_t n_rows, int32_t width);
int mht_cluster(int64_t n_cols, int64_t n_trees, int64_t n_rows, int32_t width, const int32_t *d_col_tree,
                const int32_t *d_col_rows /* [width][n_cols] */, int32_t *d_cluster_of_tree, void *d_work,
                void *stream);
int mht_assoc_solve(int64_t n_cols, int64_t n_trees, int64_t n_rows, int32_t width, const double *d_col_cost,
                    const int32_t *d_col_tree, const int32_t *d_col_rows /* [width][n_cols] */,
                    int32_t *d_selected_col, double *h_info, void *d_work, void *stream);

/* Same solve on persistent buffers: the column arrays have room for cap_cols columns (d_col_rows is [width][cap_cols])
 * and d_work holds mht_assoc_workspace(cap_cols, n_trees, n_rows, width) bytes.  With warm != 0, d_work still holds the
 * multipliers of the previous call with the SAME (cap_cols, n_trees, n_rows) and the solve starts from them; rows in
 * [clear_row_lo, clear_row_hi) start from zero (the measurement plane being recycled for the new scan).  Used by the
 * tree-sharded tracker, where measurement rows keep their ids for N+1 scans. */
int mht_assoc_solve_warm(int64_t n_cols, int64_t cap_cols, int64_t n_trees, int64_t n_rows, int32_t width,
                         const double *d_col_cost, const int32_t *d_col_tree, const int32_t *d_col_rows,
                         int32_t *d_selected_col, double *h_info, void *d_work, void *stream, int32_t warm,
                         int64_t clear_row_lo, int64_t clear_row_hi, double exact_ms /* <= 0: default 10 s */,
                         int32_t max_dual_iters /* per sifting round; <= 0: default 200 */);

/* ------------------------------------------------------------------------------------------------
 * Device-resident hypothesis forest: steps 1-3 + terminate + N-scan prune of
 * Tracker.addMeasurementList (tracker.py:194-259) with the forest kept in HBM across scans.
 * One forest per process / GPU; trees are independent so N ranks each own a shard of the trees.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mht_forest mht_forest; /* opaque */

typedef struct mht_forest_config {
    mht_model model;
    int32_t n_scan_window;   /* N          tracker.py:112-114 */
    int32_t max_trees;       /* tree slots                            */
    int32_t max_meas;        /* measurements per scan capacity        */
    int64_t max_nodes;       /* hypotheses per tree level capacity    */
    int64_t max_parents;     /* live leaves entering a scan capacity  */
    double default_Pd;       /* tracker.py:50 */
    double score_upper;      /* tracker.py:115 */
    double cnllr_upper;      /* tracker.py:116 */
    double radar_range;      /* tracker.py:45  */
    double position[2];      /* tracker.py:44  */
    int32_t max_dual_iters;  /* Lagrangian iterations per scan        */
    int32_t exact_ms;        /* wall-clock budget (ms) per scan of the exact branch & bound that closes the
                                components the dual loop leaves open (tracker.py:1155-1217 is an exact MILP);
                                0 = default (8 ms), < 0 = no exact search.  A scan whose budget ran out returns
                                a feasible selection with certified = 0 and the bound gap in mht_scan_info. */
} mht_forest_config;

/* Per-scan summary (host struct filled by mht_forest_scan). */
typedef struct mht_scan_info {
    int64_t n_parents;       /* live leaves gated this scan (L)            */
    int64_t n_children;      /* hypotheses created (L + G)                 */
    int64_t n_pairs;         /* gated (leaf,meas) pairs (G)                */
    int32_t n_trees;         /* live trees before termination              */
    int32_t n_clusters;      /* tracker.py:220                             */
    int32_t n_multi_clusters;/* clusters with >1 tree (nOptimSolved)       */
    int32_t n_dead;          /* tracks terminated this scan                */
    int32_t dual_iters;
    int32_t certified;       /* 1 = optimality certificate holds           */
    int64_t n_candidates;    /* columns surviving reduced-cost fixing      */
    int64_t bb_nodes;
    double lower_bound;      /* sum over trees of cost, unscaled by N      */
    double objective;
    float ms_gate, ms_cluster, ms_assoc, ms_prune; /* CUDA-event stage times */
    float ms_total, ms_h2d;  /* whole scan on the device (first kernel -> results on host); scan upload */
    int64_t n_active;        /* columns on the dual iteration's active list (0 = all columns iterate) */
    int32_t max_component;   /* trees in the largest component handed to the exact search */
    int32_t n_components;    /* multi-tree components handed to the exact search */
    float ms_dual;           /* time inside the persistent dual-loop kernel launches (part of ms_assoc)      */
    float ms_exact;          /* time of the exact repair stage: plan, compaction, branch & bound, write-back */
    int64_t nnz_active;      /* row incidences of the columns the dual loop iterates on (its last round)     */
    int32_t rows_active;     /* measurement rows carrying a multiplier                                        */
    int32_t bb_iters;        /* subgradient iterations spent inside the branch & bound                        */
    int32_t open_components; /* components whose exact search did not finish (0 when certified)               */
    int32_t repaired_trees;  /* trees the final feasibility check moved to their miss hypothesis (lost a row to
                                another track, or never got a primal solution); > 0 withdraws the certificate  */
} mht_scan_info;

int mht_forest_create(const mht_forest_config *cfg, mht_forest **out);
void mht_forest_destroy(mht_forest *f);
/* bytes of device memory the forest holds */
int64_t mht_forest_bytes(const mht_forest *f);

/* Tracker.initiateTarget (tracker.py:147-160): new single-node tree; returns its slot in *slot. */
int mht_forest_initiate(mht_forest *f, const double x0[4], const float P0[16], double Pd, int32_t *slot);

/* The host is done with a DEAD track's slot (it has read the history it wants to keep): the slot and the host-side
 * records behind it are recycled by a later mht_forest_initiate, so max_trees bounds the LIVE tracks, not the tracks
 * ever initiated (the reference has no such limit at all). */
int mht_forest_release(mht_forest *f, int32_t slot);

/* One Tracker.addMeasurementList hot path.  h_z[M,2] f64 host (pinned or pageable), copied H2D on
 * the forest's stream; returns after the per-track results are back on the host.
 * h_meas_used[M] (may be NULL) receives the used-measurement mask of tracker.py:331-332. */
int mht_forest_scan(mht_forest *f, int64_t M, const double *h_z, double scan_time, mht_scan_info *info,
                    uint8_t *h_meas_used);

/* Same, with the scan already resident in HBM (d_z[M,2] f64) and no host copies except the
 * 64-byte status word: the kernel-only leg bench.py times as `value`. */
int mht_forest_scan_device(mht_forest *f, int64_t M, const double *d_z, double scan_time, mht_scan_info *info);

/* ------------------------------------------------------------------------------------------------
 * Tree-sharded forests: N ranks each own a contiguous range of the trees of ONE
 * surveillance region and receive the same scan.  The gate stage (Tracker._growTarget,
 * tracker.py:207-209,309-351) is independent per tree; trees couple only through shared measurement
 * rows of the 0/1 program (Tracker._createA1, tracker.py:1042-1113).  One scan is therefore
 *   mht_forest_grow            (every rank, its own trees)
 *   mht_forest_export_columns  (every rank writes its columns into caller-owned GLOBAL column arrays,
 *                               at its column offset; the caller all-gathers the arrays, e.g. NCCL)
 *   mht_assoc_solve            (global columns -> selected column per global tree)
 *   mht_forest_select          (every rank, its slice of the selection: report, terminate, N-scan prune)
 * Measurement row ids are plane * max_meas + index with plane = scan mod (N+1): identical on every rank
 * that was created with the same configuration and has seen the same scans.
 * ---------------------------------------------------------------------------------------------- */
/* Gate stage only.  z = [M,2] f64, host (z_on_device = 0) or device.  info gets n_parents / n_children /
 * n_pairs / ms_gate.  The scan stays open until mht_forest_select. */
int mht_forest_grow(mht_forest *f, int64_t M, const double *z, int32_t z_on_device, double scan_time,
                    mht_scan_info *info, uint8_t *h_meas_used);
/* Columns of the open scan, in this forest's leaf order: d_cost[col_offset + j] = cNLLR_j - cNLLR_root
 * (= N * c_j of tracker.py:1124-1136), d_tree[col_offset + j] = tree slot + tree_offset,
 * d_rows[w * stride + col_offset + j] = measurement row of path plane w (< 0: none), w < N+1. */
int mht_forest_export_columns(mht_forest *f, int32_t tree_offset, int64_t col_offset, int64_t stride,
                              double *d_cost, int32_t *d_tree, int32_t *d_rows);
/* The same columns as ONE packed record per column {f64 cost, i32 tree, i32 rows[N+1]} (mht_record_bytes(N+1)
 * bytes each, 40 for N = 6): the payload of the single all-gather of SURVEY.md 8e.  mht_unpack_records turns the
 * gathered buffer ([world][max_per_rank] records, rank r holding h_counts[r]) into the structure-of-arrays columns
 * mht_assoc_solve takes, concatenated in rank order.  d_scratch: >= 1024 bytes of device memory. */
int32_t mht_record_bytes(int32_t width);
int mht_forest_export_records(mht_forest *f, int32_t tree_offset, void *d_records, int64_t cap_records);
int mht_unpack_records(int32_t world, const int64_t *h_counts, int64_t max_per_rank, int32_t width,
                       const void *d_gathered, int64_t stride_out, double *d_cost, int32_t *d_tree, int32_t *d_rows,
                       void *d_scratch, void *stream);
/* Close the open scan with an externally computed selection: d_selected_col[t] = LOCAL column index
 * (position in this forest's leaf order) for every tree slot t < n slots (ignored for dead slots).
 * h_assoc_info = the 8 doubles of mht_assoc_solve (may be NULL); fills info like mht_forest_scan. */
int mht_forest_select(mht_forest *f, const int32_t *d_selected_col, const double *h_assoc_info,
                      mht_scan_info *info);

/* Tracker.__dynamicWindow (tracker.py:918-950).  enabled: apply the SIZE criterion on the device in every following
 * scan -- a tree holding more than target_size_limit nodes after the scan's growth (Target.getNumOfNodes,
 * pyTarget.py:148-151; tracker.py:118 default 3000) loses one scan of its N-scan window before the scan's pruning.
 * window_roof (0 = N): upper limit for every tree's window, what the reference lowers when a whole iteration
 * exceeds 0.8 x radarPeriod (tracker.py:943-950); the wall-clock criteria themselves stay with the host, which
 * owns the clock.  mht_forest_windows reads the per-tree windows back (__targetWindowSize__, tracker.py:83). */
int mht_forest_set_dynamic_window(mht_forest *f, int32_t enabled, int32_t target_size_limit, int32_t window_roof);
int mht_forest_windows(mht_forest *f, int32_t cap, int32_t *n, int32_t *h_slot, int32_t *h_window);

/* Selected hypothesis per live tree after the last scan (Tracker.getTrackNodes, tracker.py:976):
 * h_slot[T] tree slot, h_x[T,4], h_P[T,16], h_cnllr[T], h_meas[T] (measurementNumber, 0 = miss),
 * h_status[T] (0 active, 1 out-of-range, 2 too-low-score; dead tracks are reported once, in the scan
 * that killed them, then dropped).  *n receives T. */
int mht_forest_tracks(mht_forest *f, int32_t cap, int32_t *n, int32_t *h_slot, double *h_x, float *h_P,
                      double *h_cnllr, int32_t *h_meas, int32_t *h_status);

/* measurementNumber history of one track's selected leaf back to its initial node
 * (helpFunctions.backtrackMeasurementNumbers, pymht/utils/helpFunctions.py:66-83), plus the states
 * (h_x[n,4], h_cnllr[n], h_P[n,16]) along it, initial node included.  Oldest first. */
int mht_forest_history(mht_forest *f, int32_t slot, int32_t cap, int32_t *n, int32_t *h_meas, double *h_x,
                       double *h_cnllr, float *h_P);

/* The same for EVERY live track with a handful of launches (helpFunctions.backtrackMeasurementNumbers over all tracks):
 * row i of the [cap_tracks][cap_len] output arrays belongs to slot h_slot[i] and holds h_len[i] nodes, oldest first.
 * MHT_E_CAPACITY: *n_tracks holds the number of live tracks (cap_tracks too small) or the longest history (cap_len). */
int mht_forest_histories(mht_forest *f, int32_t cap_tracks, int32_t cap_len, int32_t *n_tracks, int32_t *h_slot,
                         int32_t *h_len, int32_t *h_meas, double *h_x, double *h_cnllr, float *h_P);

/* mht_forest_history for several slots in one call (the tracks that died in a scan, read before their slots are released):
 * row i of the [n][cap_len] outputs belongs to h_slots[i] and holds h_len[i] nodes, oldest first.
 * MHT_E_CAPACITY: h_len[0] holds the longest history. */
int mht_forest_histories_of(mht_forest *f, int32_t n, const int32_t *h_slots, int32_t cap_len, int32_t *h_len,
                            int32_t *h_meas, double *h_x, double *h_cnllr, float *h_P);

/* Tracker.__associatedMeasurements__[i] (tracker.py:83,331-332,1226-1227): the (scanNumber, measurementNumber) pairs of
 * every node below the tree's current root (Target.getMeasurementSet, pyTarget.py:414-430), after the last scan's pruning.
 * MHT_E_CAPACITY: *n holds the required count. */
int mht_forest_measurement_set(mht_forest *f, int32_t slot, int32_t cap, int32_t *n, int32_t *h_scan, int32_t *h_meas);

/* Smallest distance from (px,py) to any live leaf's position: the test of
 * Target.haveNoNeightbours (pymht/pyTarget.py:181-189) used by Tracker.initiateTarget. */
int mht_forest_min_leaf_distance(mht_forest *f, double px, double py, double *dist);

/* Leaves of one tree in the reference's DFS order (Target.getLeafNodes, pyTarget.py:461-471). */
int mht_forest_leaves(mht_forest *f, int32_t slot, int64_t cap, int64_t *n, double *h_x, double *h_cnllr,
                      int32_t *h_meas);

/* ---- M-of-N track initiator (reference pymht/initiators/m_of_n.py:233-478) ----------------------------------------
 * The step after the hot path: it consumes the unused-measurement mask the gate stage returns
 * (tracker.py:266-277).  The host class pymht_b200.initiators.m_of_n.Initiator keeps the O(n) bookkeeping of the
 * preliminary tracks; the O(n1 x n2) parts run here:
 *   mht_gnn_assign  replaces _solve_global_nearest_neighbour (m_of_n.py:24-104) TOGETHER with the dense distance /
 *                   NIS matrix its two callers build first (_processInitiators :385-396, _processPreliminaryTracks
 *                   :284-303).  The reference pads the gated matrix to a square one and runs an O(n^3) Munkres on it; its
 *                   optimum is the maximum-cardinality, then minimum-distance matching of the gated pairs, which is what
 *                   this solves exactly on the sparse gated graph (csrc/gnn_core.h), one thread block per component.
 *   mht_gnn_similar replaces the all-pairs PreliminaryTrack.compareSimilarity loop of __spawn_preliminary_tracks
 *                   (m_of_n.py:196-201, 462-470). */
typedef struct mht_gnn mht_gnn; /* opaque: device buffers sized at creation */

typedef struct mht_gnn_info {
    int32_t n_edges;            /* gated pairs */
    int32_t n_components;       /* connected components with at least one pair */
    int32_t largest_component;  /* rows of the largest one */
    int32_t n_assigned;
    int32_t searches, rounds;   /* block-wide augmenting-path searches and the frontier rounds they took */
    int32_t batches;            /* batches of concurrent (one warp each) searches before them */
    int32_t spec_commits;       /* searches committed by those batches */
    int32_t spec_overflow;      /* rows whose search outgrew the per-warp table (left to the block-wide search) */
    int32_t reserved;
    float ms_gate, ms_solve;    /* device time: gating + CSR build, components + assignment */
} mht_gnn_info;

int mht_gnn_create(int64_t max_rows, int64_t max_cols, int64_t max_edges, mht_gnn **out);
void mht_gnn_destroy(mht_gnn *h);

/* mode 0 (m_of_n.py:385-401): rows = initiators (previous scan's leftover measurements), columns = unused measurements;
 *   a pair is gated when its Euclidean distance (float64 norm of the float32 difference) is <= gate (= v_max * dt).
 * mode 1 (m_of_n.py:284-303): rows = predicted measurements of the preliminary tracks with h_row_sinv[n_rows][4] = S^-1
 *   (float32, row-major); a pair is gated when its float32 NIS is <= gate (chi2 0.99 quantile); cost = float32 distance.
 * h_match[i] = assigned column of row i or -1.  HOST pointers; MHT_E_CAPACITY when the gated pairs exceed max_edges. */
int mht_gnn_assign(mht_gnn *h, int mode, int64_t n_rows, const float *h_row_xy, const float *h_row_sinv, int64_t n_cols,
                   const float *h_col_xy, double gate, int32_t *h_match, mht_gnn_info *info);

/* h_state[n_tracks + n_cand][4], h_sinv[n_tracks + n_cand][16] (inverse of covariance + R_ais of that entry, float32):
 * candidate k is compared with every entry in front of it (all existing tracks and the candidates 0..k-1);
 * pairs (k, entry) with NIS = d^T S_entry^-1 d <= threshold are returned in h_pairs[2 * n_pairs] (unordered). */
int mht_gnn_similar(mht_gnn *h, int64_t n_tracks, int64_t n_cand, const float *h_state, const float *h_sinv,
                    double threshold, int32_t *h_pairs, int64_t cap_pairs, int64_t *n_pairs);

#ifdef __cplusplus
}
#endif
#endif /* MHT_B200_H */
