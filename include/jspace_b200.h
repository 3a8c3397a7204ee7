/* jspace_b200.h — C ABI of libjspace_b200.so
 *
 * B200-native replacement for ONE hot path of BIMIB-DISCo/J-Space.jl: the spatial
 * clonal-dynamics Gillespie simulation `simulate_evolution` (reference
 * src/J_Space.jl:638-842 random drivers, :848-1104 driver tree) together with the graph
 * construction/seeding it needs (src/J_Space.jl:43-163).
 *
 * The reference has no FFI of its own (it is 100 % Julia); the boundary it exposes for this
 * path is the pair of exported Julia methods `spatial_graph` / `simulate_evolution`
 * (src/J_Space.jl:25).  Every entry point below names the reference code it replaces.  A
 * Julia wrapper (j-space.jl_b200/julia/JSpaceB200.jl) and a Python mirror
 * (j-space.jl_b200/jspace_b200/) bind exactly these symbols; see INTEGRATION.md.
 *
 * Conventions: plain C types only; every function returns 0 (JSB_OK) or a negative JSB_E*
 * code and never throws; the message for the last failure on the calling thread is
 * jsb_last_error().  Vertex ("site") numbers are 1-based everywhere in this ABI, exactly as
 * Graphs.jl numbers them.  There is NO CPU fallback: without a CUDA device jsb_run returns
 * JSB_ENODEVICE.
 */
#ifndef JSPACE_B200_H
#define JSPACE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JSB_ABI_VERSION 2

/* ---- error codes -------------------------------------------------------------------- */
#define JSB_OK 0
#define JSB_EINVAL (-1)    /* bad argument (message says which)                          */
#define JSB_ECUDA (-2)     /* CUDA runtime error                                          */
#define JSB_ENOMEM (-3)    /* host or device allocation failed                            */
#define JSB_ECAPACITY (-4) /* state too large for this build (e.g. > 2^28 sites)          */
#define JSB_ENODEVICE (-5) /* no CUDA device: there is no CPU fallback                    */
#define JSB_EIO (-6)       /* file could not be read                                      */

/* ---- contact rule, `model::String` of src/J_Space.jl:646,767-786 ---------------------- */
#define JSB_RULE_CONTACT 0 /* "contact": reject when the target site is occupied          */
#define JSB_RULE_VOTER 1   /* "voter": reject when occupied by the same subpopulation     */
#define JSB_RULE_HVOTER 2  /* "h_voter"/"hvoter": reject when Fit(cell) <= Fit(target)    */
#define JSB_RULE_NONE 3    /* any other string: the reference never rejects               */

/* ---- which `simulate_evolution` method ------------------------------------------------ */
#define JSB_METHOD_RANDOM 1 /* src/J_Space.jl:638  (random driver labels, N(avg,std) gain) */
#define JSB_METHOD_TREE 2   /* src/J_Space.jl:848  (Tree_Driver_Configure = 1)             */

/* ---- quirk flags -----------------------------------------------------------------------
 * By default the excision ("bottleneck") index behaves exactly as the reference does:
 * method 1 never advances id_bottleneck (src/J_Space.jl:813-838) so only t_bottleneck[1]
 * can fire; method 2 advances it at the end of EVERY iteration while it is smaller than the
 * vector length (src/J_Space.jl:1062-1064).  JSB_QUIRK_EXCISION_INTENT switches both
 * methods to the evident intent: advance after an excision fired.                         */
#define JSB_QUIRK_EXCISION_INTENT 0x1u
/* Test hook: evaluate every iteration through the one-at-a-time path instead of the
 * speculative 32-iteration batch.  Results must be identical.                              */
#define JSB_QUIRK_DEBUG_SERIAL 0x100u
/* Test hooks for the event-class selection: always use the exact prob_cum
 * table / widen the guard band to 1e-3 so the exact fallback runs constantly.  Results must be
 * identical.                                                                               */
#define JSB_QUIRK_DEBUG_EXACT_CLASSES 0x200u
#define JSB_QUIRK_DEBUG_WIDE_GUARD 0x400u
/* Rejection-free mode test hook: after every event recompute the rates around the changed sites
 * from scratch and compare with the incrementally maintained tree (status JSB_ST_INTERNAL on a
 * mismatch).                                                                                 */
#define JSB_QUIRK_DEBUG_VERIFY_TREE 0x800u
/* Test hook for the decoupled mode of the exact kernel: hand the exact fold and
 * time chain to a helper warp whenever one is available, also for replicates with few clones
 * (by default only replicates with >= 64 clones do, below that it does not pay).  Results must
 * be identical.                                                                               */
#define JSB_QUIRK_DEBUG_DECOUPLE 0x1000u

/* ---- output selection (jsb_run_opts.flags) -------------------------------------------- */
#define JSB_OUT_EVENTS 0x1u  /* keep the per-replicate event log (`df`)                   */
#define JSB_OUT_LATTICE 0x2u /* keep the final lattice (clone id + cell id per site)      */
#define JSB_OUT_LISTS 0x4u   /* keep the final ordered cs_alive / ca_subpop lists          */
#define JSB_OUT_SERIES 0x8u  /* keep the per-effective-event series (list_len_node_occ, CA_Alive_TOT;
                              * src/J_Space.jl:656-658,740-741,757-758,807-808,835-836) without shipping
                              * the event log to the host: the log is reduced on the device           */
/* ---- simulation mode (jsb_run_opts.flags) ------------------------------------------------
 * JSB_MODE_REJECTION_FREE: same stochastic law, different algorithm.  Null
 * events are not simulated: a per-site propensity tree selects only state-changing events and
 * time advances by Exp(total effective rate).  Trajectories are NOT comparable event for event
 * with the reference (validated by two-sample KS tests on ensemble statistics); n_iterations
 * then counts effective events; ordered lists (JSB_OUT_LISTS) do not exist in this mode.     */
#define JSB_MODE_REJECTION_FREE 0x100u

/* ---- event log record: one row of the reference's `df` (src/J_Space.jl:284-290) -------
 * Rows appear in exactly the reference's push! order (cell_birth :351-352 / :387-388,
 * cell_death :579, migration_cell :612).  Cell ids are sequential integers in `uuid1` call
 * order (seed cells first, then id, id2 per accepted birth, src/J_Space.jl:340-341)
 * instead of time-based UUIDs.                                                            */
#define JSB_EV_DUPLICATE 0 /* "Duplicate": Notes = [parent]                               */
#define JSB_EV_MUTATION 1  /* "Mutation":  Notes = [parent, new driver]                   */
#define JSB_EV_DEATH 2     /* "Death"                                                     */
#define JSB_EV_MIGRATION 3 /* "Migration"                                                 */

#define JSB_EVF_GROUP_START 0x1 /* first row of one effective event (one series push)     */
#define JSB_EVF_DISPLACED 0x2   /* Death row written by cell_birth :335-338 (voter rules) */
#define JSB_EVF_EXCISION 0x4    /* Death row written by the excision block :824-833       */

typedef struct jsb_event {
  double time;      /* df.Time                                                            */
  uint32_t cell;    /* df.Cell (sequential id)                                            */
  uint32_t parent;  /* df.Notes[1] for Duplicate/Mutation, 0 otherwise                    */
  uint32_t site;    /* vertex the cell sits on after the row (Death: the vacated vertex)  */
  uint32_t site2;   /* Migration: the vertex it left; otherwise 0                         */
  uint16_t clone;   /* :Subpop of the cell after the row (index into the clone table)     */
  uint16_t label;   /* Mutation: new driver label (method 1) / tree node + 1 (method 2)   */
  uint8_t type;     /* JSB_EV_*                                                           */
  uint8_t flags;    /* JSB_EVF_*                                                          */
  uint16_t pad;
} jsb_event; /* 32 bytes */

/* One push to list_len_node_occ / CA_Alive_TOT (one effective event).  The reference pushes n_cs_alive and
 * length.(ca_subpop); a row holds n_cs_alive and the one or two list lengths that changed, so that the
 * caller rebuilds the per-clone counts by running sums.  An excision that removes k > 1 cells is ONE push:
 * it is written as k rows, all but the last with JSB_SERIES_MORE set in n_alive.                       */
#define JSB_SERIES_MORE 0x80000000u
typedef struct jsb_series_row {
  uint32_t n_alive; /* n_cs_alive after the row (low 31 bits)                                        */
  uint16_t gain;    /* clone (1-based) whose list grew by one cell, 0 = none; the first time an index
                       appears it is a new clone of one cell (Mutation)                               */
  uint16_t loss;    /* clone whose list lost one cell, 0 = none                                       */
} jsb_series_row; /* 8 bytes */

/* ---- clone table: set_mut_pop + α of the 7-tuple (src/J_Space.jl:841) ------------------ */
typedef struct jsb_clone {
  double alpha;     /* α[i]                                                               */
  uint32_t count;   /* length(ca_subpop[i]) at the end of the run                         */
  uint16_t parent;  /* 1-based index of the parent clone, 0 for the founder               */
  uint16_t label;   /* last driver of the genotype: label (method 1; founder = 1) or tree
                       node index + 1 (method 2)                                          */
} jsb_clone; /* 16 bytes */

/* ---- per-replicate summary --------------------------------------------------------------*/
#define JSB_ST_OK 0
#define JSB_ST_NONTERMINATING 1 /* rate_death == 0 and the lattice filled: the reference loop
                                   (src/J_Space.jl:687-688) never exits; stopped at Tf     */
#define JSB_ST_CLONE_CAPACITY 2 /* all 999 driver labels used: reference errors at :371    */
#define JSB_ST_LOG_OVERFLOW 3   /* event-log pool exhausted; counts are valid, log is cut  */
#define JSB_ST_MAX_ITER 4       /* jsb_params.max_iterations reached                       */
#define JSB_ST_INTERNAL 5       /* internal capacity error (list arena)                    */

typedef struct jsb_summary {
  uint64_t n_iterations; /* iterations of the while loop, null events included            */
  uint64_t n_events;     /* rows of df                                                    */
  double t_final;        /* t_curr when the loop exited                                   */
  uint32_t n_effective;  /* length(list_len_node_occ) - 1                                 */
  uint32_t n_births;     /* accepted births                                               */
  uint32_t n_deaths;     /* Death class events                                            */
  uint32_t n_migrations; /* effective migrations                                          */
  uint32_t n_displaced;  /* cells overwritten under voter / hvoter                        */
  uint32_t n_excised;    /* cells removed by excision                                     */
  uint32_t n_alive;      /* n_cs_alive at exit                                            */
  uint32_t n_clones;     /* length(set_mut_pop)                                           */
  uint32_t next_cell_id; /* ids handed out + 1                                            */
  int32_t status;        /* JSB_ST_*                                                      */
} jsb_summary; /* 64 bytes */

/* Snapshot marker: the reference pushes copy(G) when t_curr > Time_of_sampling[k]
 * (src/J_Space.jl:708-714); the state at that moment is the replay of the first
 * `n_events_before` log rows.                                                             */
typedef struct jsb_snapshot {
  double time;            /* t_curr of the iteration that triggered it                    */
  uint64_t iteration;     /* 1-based loop iteration                                       */
  uint64_t n_events_before;
  uint32_t n_alive;
  uint32_t index;         /* k, 0-based                                                   */
} jsb_snapshot; /* 32 bytes */

/* ---- parameters of one sweep point: the positional/keyword arguments of
 * simulate_evolution (src/J_Space.jl:638-650 and :848-859) ------------------------------- */
typedef struct jsb_params {
  uint32_t size;   /* = sizeof(jsb_params); versions the struct                           */
  uint32_t method; /* JSB_METHOD_*                                                        */
  double Tf;
  double rate_birth;     /* method 1: α[1] (src/J_Space.jl:664); ignored by method 2      */
  double rate_death;
  double rate_migration;
  double mu_driver;      /* μ_dri                                                         */
  double adv_mean;       /* driv_average_advantage (method 1)                             */
  double adv_std;        /* driv_std_advantage (method 1)                                 */
  int32_t rule;          /* JSB_RULE_*                                                    */
  int32_t n_start_cells; /* n_cell of spatial_graph (src/J_Space.jl:47)                   */
  int32_t n_bottleneck;
  int32_t n_snapshots;
  const double* t_bottleneck;     /* [n_bottleneck]                                       */
  const double* ratio_bottleneck; /* [n_bottleneck], each in [0,1]                        */
  const double* t_snapshot;       /* Time_of_sampling [n_snapshots]                       */
  /* method 2: driver tree, rows of edgelist_treedriver in file order; node ids 0-based   */
  int32_t n_tree_nodes;
  int32_t n_tree_edges;
  const int32_t* edge_parent; /* [n_tree_edges]                                           */
  const int32_t* edge_child;  /* [n_tree_edges]                                           */
  const double* node_rate;    /* [n_tree_nodes] rate looked up by name (:503-504)         */
  int32_t root_node;          /* setdiff!(parents, children)[1] (:873)                    */
  uint32_t quirk_flags;       /* JSB_QUIRK_*                                              */
  double root_rate;           /* α[1] = rate in the FIRST row of driver_birth_rates (:881)*/
  uint64_t max_iterations;    /* 0 = unlimited                                            */
} jsb_params;

typedef struct jsb_run_opts {
  uint32_t size;  /* = sizeof(jsb_run_opts)                                               */
  uint32_t flags; /* JSB_OUT_*                                                            */
  uint64_t seed;  /* Philox key; replaces Config.toml `seed` (src/Start.jl:12-13)         */
  /* Global replicate numbering: g = point * reps_per_point + r.  This call simulates
   * g in [g_begin, g_end); replicate g always uses Philox stream g, so results do not
   * depend on how an ensemble is sharded over GPUs/ranks.  g_end = 0 means "all".        */
  int64_t g_begin;
  int64_t g_end;
  int32_t device;         /* CUDA device ordinal (used when n_devices == 0)               */
  int32_t warps_per_sm;   /* 0 = library default                                          */
  uint64_t log_pool_bytes; /* event-log pool size per device; 0 = library default         */
  /* Several GPUs of one box in ONE call: the replicate range [g_begin, g_end) is
   * cut into n_devices contiguous shards, one host thread + stream per device; replicate g still
   * uses Philox stream g, so the result is the same as on one device.  n_devices == 0: `device`. */
  int32_t n_devices;
  int32_t reserved;
  const int32_t* devices; /* [n_devices] CUDA device ordinals, all different                  */
} jsb_run_opts;

typedef struct jsb_graph jsb_graph;
typedef struct jsb_result jsb_result;

/* ---- library ---------------------------------------------------------------------------*/
int jsb_abi_version(void);
int jsb_device_count(void);
const char* jsb_last_error(void);

/* ---- graphs: replace spatial_graph / initialize_graph --------------------------------- */
/* spatial_graph(row, col, seed; dim) src/J_Space.jl:43-64: dim 2 -> Graphs.grid((row,col,1)),
 * anything else -> cube of side trunc((row*col)^(1/3))+1; seed vertex by :131-133.        */
int jsb_graph_lattice(int dim, int64_t row, int64_t col, jsb_graph** out);
/* Arbitrary undirected graph in CSR form (1-based vertex ids in colidx, neighbours ascending,
 * rowptr 0-based with nv+1 entries).  The intent of spatial_graph(path, seed)
 * src/J_Space.jl:67-71 for graphs too large for a dense text matrix.                      */
int jsb_graph_csr(int64_t nv, const int64_t* rowptr, const int32_t* colidx, jsb_graph** out);
/* spatial_graph(path, seed) src/J_Space.jl:67-71: whitespace-separated dense 0/1 matrix.  */
int jsb_graph_dense_file(const char* path, jsb_graph** out);
int64_t jsb_graph_nv(const jsb_graph* g);
int64_t jsb_graph_ne(const jsb_graph* g); /* undirected edge count                         */
/* Seed candidates for n_cell == 1: the lattice formula (one vertex) or all vertices of
 * maximal closeness centrality (src/J_Space.jl:83-85).  Computed lazily, cached.          */
int jsb_graph_seed_candidates(jsb_graph* g, const int64_t** sites, int64_t* n);
/* Override the candidates (e.g. a precomputed centre for very large graphs).              */
int jsb_graph_set_seed_candidates(jsb_graph* g, const int64_t* sites, int64_t n);
/* Neighbours of one vertex (ascending, 1-based); returns the degree, <0 on error.         */
int jsb_graph_neighbors(const jsb_graph* g, int64_t site, int64_t* out, int cap);
void jsb_graph_free(jsb_graph* g);

/* ---- run: replaces both simulate_evolution methods, for an ensemble -------------------- */
/* Thread safety: jsb_run is blocking and re-entrant.  Calls that use the same device are serialised
 * inside the library (one launch configuration and one set of cached device buffers per device);
 * calls on different devices run concurrently.  A jsb_graph may be shared between threads: its
 * per-device copies are created under a lock.  Results are independent objects.               */
int jsb_run(jsb_graph* g, const jsb_params* points, int64_t n_points, int64_t reps_per_point,
            const jsb_run_opts* opts, jsb_result** out);
/* Device buffers (workspace, log pool, outputs) and the pinned staging block are kept between calls;
 * jsb_trim() releases them (every device).                                                     */
void jsb_trim(void);
/* Same, but the per-replicate outputs stay on the device and only timing is returned
 * (used by benchmarks to separate kernel time from copies).  elapsed_ms = CUDA-event time
 * of the simulation kernel alone.                                                         */
int jsb_result_kernel_ms(const jsb_result* r, double* elapsed_ms);
int jsb_result_launches(const jsb_result* r, int64_t* n_kernel_launches);

/* ---- results (local replicate index i = g - g_begin) ----------------------------------- */
int64_t jsb_result_count(const jsb_result* r);
int jsb_result_summaries(const jsb_result* r, const jsb_summary** p);
int jsb_result_clones(const jsb_result* r, int64_t i, const jsb_clone** p, int32_t* n);
int jsb_result_events(jsb_result* r, int64_t i, const jsb_event** p, int64_t* n);
int jsb_result_lattice(const jsb_result* r, int64_t i, const uint16_t** clone,
                       const uint32_t** cell_id);
/* Final ordered lists: cs_alive (list 0) and ca_subpop[c] (list c, 1-based), as 1-based
 * vertex numbers.                                                                         */
int jsb_result_list(jsb_result* r, int64_t i, int32_t list, const uint32_t** p, int64_t* n);
int jsb_result_snapshots(const jsb_result* r, int64_t i, const jsb_snapshot** p, int32_t* n);
/* JSB_OUT_SERIES: the rows of replicate i (the initial entry of list_len_node_occ, src/J_Space.jl:658, is the
 * number of starting cells and is not a row), and for every snapshot the number of rows before it
 * (CA_Alive_TOT gets an extra row there, :712).                                                       */
int jsb_result_series(const jsb_result* r, int64_t i, const jsb_series_row** p, int64_t* n,
                      const int64_t** rows_before_snapshot, int32_t* n_snapshots);
/* Ensemble histograms over this shard (fixed bins, uint64 counts) for sweep point `point`:
 * kind 0 = final n_alive (bins of width nv/64), kind 1 = n_clones (bins 0..63+),
 * kind 2 = log2(n_iterations).  64 bins each; meant to be summed over ranks with NCCL.     */
int jsb_result_histogram(const jsb_result* r, int64_t point, int kind, uint64_t out[64]);
void jsb_result_free(jsb_result* r);

/* ---- genealogy of sampled cells (replaces create_matrix_relational,
 * src/sampling_phylogentic_relation_and_genotype.jl:60-104) ------------------------------ */
typedef struct jsb_relation {
  double time;
  uint32_t father;
  uint32_t child;
  uint16_t subpop_child;
  uint8_t type; /* JSB_EV_DUPLICATE or JSB_EV_MUTATION */
  uint8_t pad[5];
} jsb_relation; /* 24 bytes */
int jsb_result_genealogy(jsb_result* r, int64_t i, const uint32_t* sample_cells, int64_t n_samples,
                         jsb_relation** rows, int64_t* n_rows);
void jsb_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* JSPACE_B200_H */
