/* suchtree_b200 -- C ABI of the B200-native patristic-distance path.
 *
 * Drop-in boundary for the ONE hot path of ryneches/SuchTree:
 * batched MRCA / patristic distance, the all-pairs distance matrix, the
 * linked-tree pair enumeration + sampler and pearson().  The reference has no
 * FFI layer of its own -- the seam is the `def` -> `cdef` call inside
 * SuchTree/MuchTree.pyx -- so every entry point below cites the reference
 * `cdef`/`def` it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain C types only; opaque handles; every function returns an st_status
 *     (0 = ST_OK) and st_last_error() gives a thread-local message.
 *   - a tree handle is immutable after st_tree_create(): any number of host
 *     threads may query it concurrently.
 *   - node ids are the reference's ids: in-order ranks of the binarised tree
 *     (MuchTree.pyx:171-180), leaves even, internal nodes odd.
 *   - "host" entry points take host pointers (pageable or pinned) and do the
 *     H2D/D2H copies themselves; "_device" entry points take device pointers on
 *     the tree's device and a cudaStream_t (as void*), and never synchronise.
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with ST_ERR_CUDA.
 */
#ifndef SUCHTREE_B200_H
#define SUCHTREE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ST_API __attribute__((visibility("default")))
#else
#define ST_API
#endif

typedef enum st_status {
    ST_OK = 0,
    ST_ERR_INVALID_ARG = 1,   /* NULL pointer, negative size, bad dtype ...            */
    ST_ERR_CUDA = 2,          /* CUDA runtime / driver error, or no device             */
    ST_ERR_NOT_BINARY = 3,    /* a node with exactly one child, or child/parent mismatch */
    ST_ERR_NOT_INORDER = 4,   /* ids are not in-order ranks (MuchTree.pyx:171-180)     */
    ST_ERR_NODE_RANGE = 5,    /* a queried id is outside [0, n_nodes): see st_bad_node */
    ST_ERR_LENGTH_MISMATCH = 6,
    ST_ERR_NOMEM = 7
} st_status;

typedef struct st_tree st_tree; /* device-resident index of one tree on one GPU */

typedef struct st_tree_info {
    int64_t n_nodes;      /* SuchTree.size        MuchTree.pyx:236-240 */
    int64_t n_leaves;     /* SuchTree.num_leaves  MuchTree.pyx:247-251 */
    int32_t root;         /* SuchTree.root_node   MuchTree.pyx:265-269 */
    int32_t depth;        /* SuchTree.depth: max #nodes on a leaf->root path, MuchTree.pyx:218-225 */
    int32_t device;       /* CUDA device ordinal the index lives on */
    int32_t block_shift;  /* log2(ids per RMQ block)        */
    int32_t micro_shift;  /* log2(ids per RMQ micro block)  */
    int32_t n_blocks;     /* RMQ blocks (block table lives in shared memory) */
    int64_t index_bytes;  /* device bytes held by the index */
    int32_t query_smem_bytes;
    int32_t sm_count;
    int32_t layout;       /* 0 = wide records (32 B/node, double-double root distance) + 64-bit block tables,
                             1 = compact records (16 B/node; every root distance exact in fp64) + 32-bit tables,
                             2 = wide records + 32-bit block tables (inexact root distances) */
    int32_t paired_records; /* 1: the pair kernel fetches each endpoint's whole 32-byte sector (its record and
                               its slot neighbour's) -- chosen at build time when the probe below says that the
                               neighbour is often the MRCA (ladder-like trees); compact layout only */
    double probe_third_gather;  /* build-time probe, 16384 random leaf pairs: fraction whose MRCA is neither a
                                   block minimum nor an endpoint (the plain kernel gathers its root distance) */
    double probe_neighbour_hit; /* ... and fraction for which an endpoint's sector neighbour IS the MRCA */
} st_tree_info;

/* thread-local text of the last error raised on this thread ("" if none) */
ST_API const char *st_last_error(void);
ST_API int st_version(void);
/* hash of the CUDA sources this binary was compiled from (suchtree_b200/build.py); the
 * Python loader compares it with the sources on disk and rebuilds a stale library */
ST_API const char *st_build_id(void);
/* hash of the sources of the pair kernel alone (st_query.cu, st_device.cuh, st_internal.cuh):
 * the stamp on profiles/traffic.json, so that an ncu figure is never quoted for another kernel */
ST_API const char *st_pairs_kernel_id(void);
ST_API int st_device_count(int *count);

/* ---- tree: replaces `cdef struct Node` + SuchTree.__init__'s fill/depth passes
 *      (MuchTree.pyx:55-60, 160, 182-225) and __dealloc__ (:230-232).
 * parent/left/right/edge_len are HOST arrays of n_nodes entries in the
 * reference's Node layout split by field; the root has parent -1 (its edge_len
 * is ignored: the reference stores the -1 sentinel there), leaves have left ==
 * right == -1; edge_len is fp32 exactly as the reference quantises it, epsilon
 * substitution (MuchTree.pyx:136,188-194) already applied by the caller.
 * The structure is validated on the host (strictly binary, ids are in-order
 * ranks); depth, root distance (double-double) and the range-minimum index are
 * then built by kernels on `device`.
 * block_shift / micro_shift = 0 pick the defaults (tests force small blocks to
 * cover every query path on small trees). */
ST_API int st_tree_create(int device, int64_t n_nodes, const int32_t *parent, const int32_t *left,
                   const int32_t *right, const float *edge_len, int block_shift, int micro_shift,
                   st_tree **out);
/* same, with flags: ST_TREE_WIDE_LAYOUT keeps the 32-byte records even when the
 * compact 16-byte layout (chosen automatically when every root distance is exact
 * in fp64; results are bit-identical either way) would apply. */
#define ST_TREE_WIDE_LAYOUT 1
ST_API int st_tree_create_ex(int device, int64_t n_nodes, const int32_t *parent, const int32_t *left,
                      const int32_t *right, const float *edge_len, int block_shift, int micro_shift,
                      int flags, st_tree **out);
ST_API void st_tree_destroy(st_tree *tree);
ST_API int st_tree_get_info(const st_tree *tree, st_tree_info *info);
/* copies the device-built per-node arrays back (any pointer may be NULL):
 * node depth (root = 0) and root distance as a double-double (hi + lo). */
ST_API int st_tree_export(const st_tree *tree, int32_t *depth, double *rd_hi, double *rd_lo);

/* ---- NEWICK loader (host code): replaces the dendropy calls and the two fill passes
 *      of SuchTree.__init__ (MuchTree.pyx:138-157, 171-216): parse, resolve polytomies
 *      the way dendropy's resolve_polytomies() does, assign in-order ids, substitute
 *      epsilon for missing / zero lengths, quantise to fp32.  Iterative and O(n).
 * text: NEWICK bytes (UTF-8; need not be NUL-terminated).  Structural errors ->
 * ST_ERR_NOT_BINARY with the message in st_last_error().  The arrays it returns are
 * what st_tree_create() takes. */
typedef struct st_newick st_newick;
ST_API int st_newick_parse(const char *text, int64_t len, st_newick **out);
ST_API void st_newick_free(st_newick *nw);
ST_API int st_newick_info(const st_newick *nw, int64_t *n_nodes, int64_t *n_leaves, int32_t *root,
                   int64_t *names_bytes);
/* copies out the node arrays (n_nodes entries each; any pointer may be NULL) */
ST_API int st_newick_arrays(const st_newick *nw, int32_t *parent, int32_t *left, int32_t *right,
                     float *distance, float *support);
/* leaves in ascending id: ids [n_leaves], name offsets [n_leaves + 1] into `names`
 * (names_bytes bytes, no separators) */
ST_API int st_newick_leaves(const st_newick *nw, int32_t *leaf_ids, int64_t *name_offsets, char *names);

/* after ST_ERR_NODE_RANGE: the id SuchTree.distances_bulk would report in its
 * InvalidNodeError (max id if it is >= size, else min id; MuchTree.pyx:897-903) */
ST_API int64_t st_bad_node(void);

/* ---- batched distances, host buffers: replaces SuchTree._distances
 *      (MuchTree.pyx:911-943) as called from distances_bulk (:872-909).
 * pairs: int64 [n,2] with element strides (stride0, stride1) -- the reference's
 * `long[:,:]` memoryview accepts any strides; out: double [n]. */
ST_API int st_distances(const st_tree *tree, const int64_t *pairs, int64_t stride0, int64_t stride1,
                 int64_t n, double *out);

/* ---- batched MRCA, host buffers: replaces SuchTree._mrca (MuchTree.pyx:999-1030)
 *      as called from common_ancestor (:1128-1149); out: int32 [n]. */
ST_API int st_mrca(const st_tree *tree, const int64_t *pairs, int64_t stride0, int64_t stride1, int64_t n,
            int32_t *out);

/* ---- device-resident variant (the throughput path): d_pairs is a device array
 * of n (a,b) pairs, contiguous, of int32 (idx_bits = 32) or int64 (idx_bits =
 * 64); d_out double[n] and/or d_mrca int32[n] may be NULL.  Asynchronous on
 * `stream`.  Out-of-range ids produce NaN / -1 and are reported by the next
 * st_check_range(). */
ST_API int st_distances_device(const st_tree *tree, const void *d_pairs, int idx_bits, int64_t n,
                        double *d_out, int32_t *d_mrca, void *stream);
/* synchronises `stream`, returns ST_ERR_NODE_RANGE if any launch since the last
 * check saw an out-of-range id (st_bad_node() then reports it), and resets. */
ST_API int st_check_range(const st_tree *tree, void *stream);

/* ---- quartet topologies: replaces SuchTree._quartet_topologies (MuchTree.pyx:1331-1376)
 *      as called from quartet_topologies_bulk (:1271-1329).
 * quartets: int64 [n,4] node ids (any nodes, arbitrary order) with element strides
 * (stride0, stride1); out: int64 [n,4], contiguous: each quartet reordered so that
 * (out[i,0], out[i,1]) and (out[i,2], out[i,3]) are the sister pairs -- the pair
 * (of the six, in the order ab ac ad bc bd cd) whose MRCA occurs once among the six
 * MRCAs comes first; without one the reference's fall-through order (c d a b).
 * Out-of-range ids -> ST_ERR_NODE_RANGE (st_bad_node(): max id if >= size, else min). */
ST_API int st_quartet_topologies(const st_tree *tree, const int64_t *quartets, int64_t stride0,
                          int64_t stride1, int64_t n, int64_t *out);
/* device-resident variant: d_quartets / d_out contiguous int64 [n,4] on the tree's
 * device, asynchronous on `stream`; out-of-range ids give a row of -1 and are
 * reported by the next st_check_range(). */
ST_API int st_quartet_topologies_device(const st_tree *tree, const int64_t *d_quartets, int64_t n,
                                 int64_t *d_out, void *stream);
/* the same with int32 ids in and out (32 instead of 64 streamed bytes per quartet) */
ST_API int st_quartet_topologies_device32(const st_tree *tree, const int32_t *d_quartets, int64_t n,
                                   int32_t *d_out, void *stream);

/* deterministic synthetic input: n random leaf-id pairs (ids 2*k, k uniform in
 * [0, n_leaves)), Philox4x32-10 keyed by `seed`, counter = first_pair + i, as
 * int32 (idx_bits = 32) or int64 pairs on the device. */
ST_API int st_random_leaf_pairs_device(const st_tree *tree, uint64_t seed, int64_t first_pair, int64_t n,
                                void *d_pairs, int idx_bits, void *stream);

/* ---- all-pairs matrix: replaces SuchTree.pairwise_distances (MuchTree.pyx:1082-1124).
 * ids: HOST int64[n] node ids (NULL = all leaves in ascending id, the
 * reference's default, :1097-1099); writes rows [row_begin,row_end) of the
 * symmetric n x n fp64 matrix, row-major, to `out` (row_end-row_begin rows of n).
 * out_on_device != 0: `out` is a device pointer on the tree's device and the
 * call is asynchronous on `stream`; else `out` is a host pointer. */
ST_API int st_distance_matrix(const st_tree *tree, const int64_t *ids, int64_t n, int64_t row_begin,
                       int64_t row_end, double *out, int out_on_device, void *stream);

/* ---- linked trees: replaces the body of SuchLinkedTrees.linked_distances
 *      (MuchTree.pyx:2900-2934).  linklist: HOST int64 [n_links,2] rows
 * [TreeB leaf id, TreeA leaf id] (:2869-2870).  For k enumerating (i, j<i) in
 * the reference's order (:2919-2925): ids_a[k] = (ll[j,1], ll[i,1]),
 * ids_b[k] = (ll[j,0], ll[i,0]); out_a[k] / out_b[k] the distances in tree_a /
 * tree_b.  ids_a / ids_b may be NULL.  Both trees must be on the same device. */
ST_API int st_linked_distances(const st_tree *tree_a, const st_tree *tree_b, const int64_t *linklist,
                        int64_t n_links, double *out_a, double *out_b, int64_t *ids_a,
                        int64_t *ids_b);

/* ---- sampler, exact reference stream: one cycle of
 *      SuchLinkedTrees.sample_linked_distances (MuchTree.pyx:3024-3052):
 * buckets x n sampled link pairs drawn with the reference's xorshift64* stream
 * (:2936-2949) starting from *seed (updated on return as the reference's
 * self.seed would be), distances in both trees written to out_a/out_b
 * [buckets*n] (HOST), per-bucket sums of d and d^2 ADDED to sums_a/sumsq_a/
 * sums_b/sumsq_b [buckets] (HOST). */
ST_API int st_sample_linked_cycle(const st_tree *tree_a, const st_tree *tree_b, const int64_t *linklist,
                           int64_t n_links, uint64_t *seed, int32_t buckets, int32_t n,
                           double *out_a, double *out_b, double *sums_a, double *sumsq_a,
                           double *sums_b, double *sumsq_b);

/* ---- sampler, throughput path: n_samples link pairs (with replacement) drawn
 * by Philox4x32-10 (key = seed, counter = first_sample + i); nothing is
 * materialised.  moments[8] (HOST) receives
 *   { n, sum x, sum y, sum (x-x0)^2.. } -- see st_moments below.           */
typedef struct st_moments {
    double n;
    double x0, y0;        /* shift used for conditioning (first sample)   */
    double sx, sy;        /* sum (x - x0), sum (y - y0)                    */
    double sxx, syy, sxy; /* sum (x-x0)^2, sum (y-y0)^2, sum (x-x0)(y-y0)  */
} st_moments;
ST_API int st_sample_moments(const st_tree *tree_a, const st_tree *tree_b, const int64_t *linklist,
                      int64_t n_links, uint64_t seed, int64_t first_sample, int64_t n_samples,
                      double x0, double y0, st_moments *out);
/* ---- exhaustive variant: the same moments over link pairs [first_pair, first_pair +
 * n_pairs) of linked_distances()'s enumeration (MuchTree.pyx:2919-2925), i.e.
 * linked_distances() + pearson() (:2900-2934, :62-87) fused, nothing materialised.
 * Shard the pair range across GPUs and add the sums. */
ST_API int st_linked_moments(const st_tree *tree_a, const st_tree *tree_b, const int64_t *linklist,
                      int64_t n_links, int64_t first_pair, int64_t n_pairs, double x0, double y0,
                      st_moments *out);
/* ---- per-clade scan: the host loop of the reference's co-phylogeny examples
 *      (docs/examples/SuchLinkedTree_examples.md:286-310) -- for every internal node of one
 *      tree { subset_b(node) (MuchTree.pyx:2876-2886, which re-runs _build_linklist,
 *      :2845-2874); linked_distances() (:2900-2934); pearson() (:62-87) } -- as ONE launch
 *      sequence; nothing is materialised.
 * side = 0: clades of TreeB (selects on linklist[:,0]); side = 1: clades of TreeA
 * (linklist[:,1]).  A clade is the inclusive interval [clade_lo[c], clade_hi[c]] of in-order
 * node ids (the leaves SuchTree.get_leaves(node) returns, :427-463, are exactly the even ids
 * in the interval spanned by the node's subtree); HOST int64 [n_clades].  For clade c the
 * links whose id on `side` lies in the interval are the subset subset_b(node) would select
 * out of `linklist`; n_links_out[c] (may be NULL) = their number (subset_n_links).  When
 * max(min_links, 2) <= n <= max_links (max_links < 0: no upper bound), out[c] receives the
 * moments over the n(n-1)/2 link pairs of the subset with x0, y0 = the distances of its first
 * pair; otherwise out[c].n = 0 and the clade costs nothing.  HOST st_moments [n_clades]. */
ST_API int st_clade_moments(const st_tree *tree_a, const st_tree *tree_b, const int64_t *linklist,
                     int64_t n_links, int side, const int64_t *clade_lo, const int64_t *clade_hi,
                     int64_t n_clades, int64_t min_links, int64_t max_links, st_moments *out,
                     int64_t *n_links_out);
/* Pearson r from (possibly all-reduced) moments: sxy / sqrt(sxx*syy + 1e-20),
 * the reference's formula (MuchTree.pyx:79) on centred sums. */
ST_API double st_moments_pearson(const st_moments *m);

/* ---- pearson(): replaces _pearson (MuchTree.pyx:62-79); x, y HOST double[n], pageable or
 * page-locked; streamed to the device in chunks, fp64 shifted moments, fixed-order fold. */
ST_API int st_pearson(int device, const double *x, const double *y, int64_t n, double *r);

/* ---- measurement helper: rate (pairs/s) at which the host thread pool packs int64
 * id pairs into int32 pinned staging -- the host stage of medium-size calls and of
 * st_quartet_topologies() (long st_distances() / st_mrca() calls use the bit stream below). */
ST_API int st_bench_pack(int64_t n_pairs, int iters, double *pairs_per_s);

/* ---- the host stage of st_distances() / st_mrca() on its own (tests, measurements; no device):
 * packs n (a, b) int64 pairs (element strides as in st_distances) into the bit stream the pair
 * kernel reads -- pair i in bits [i * 2w, (i + 1) * 2w), a in the low w = id_bits bits, b above --
 * with the host thread pool.  out: ceil(n * 2w / 64) + 1 words.  *or_of_ids = OR of every id seen
 * (a bit at or above w set <=> some id is negative or >= 2^w). */
/* what the chunked host pipeline does with a page-locked input of this tree in this process: the
 * fraction of every chunk it bit-packs (the rest is DMA'd as int64) and the id width of the stream. */
ST_API int st_host_route_info(const st_tree *t, double *pack_fraction, int *id_bits);
ST_API int st_host_pack_pairs(const int64_t *pairs, int64_t stride0, int64_t stride1, int64_t n, int id_bits,
                       uint64_t *out, uint64_t *or_of_ids);

/* ---- measurement helper: what the host interface can carry -- `iters` rounds of an H2D copy
 * of h2d_bytes and a concurrent D2H copy of d2h_bytes between pinned host memory and the
 * device, in chunks of chunk_bytes (<= 0: 64 MiB) on two streams; *seconds = wall time per
 * round.  The roofline of st_distances(): 16 B/pair in, 8 B/pair out. */
ST_API int st_bench_copy(int device, int64_t h2d_bytes, int64_t d2h_bytes, int64_t chunk_bytes, int iters,
                  double *seconds);

/* ---- page-locked host memory for results: SuchTree.distances_bulk returns a FRESH array
 *      (MuchTree.pyx:907 `result = np.zeros(...)`); the shim takes it from this pool so that
 * the D2H copies of st_distances() / st_distance_matrix() land in it directly instead of in
 * staging + a host copy.  Freed blocks are cached (SUCHTREE_B200_PINNED_CACHE_MB, default
 * 4096) and recycled by later calls of the same size class. */
ST_API int st_host_alloc(int64_t bytes, void **out);
ST_API int st_host_free(void *p);
ST_API int st_host_trim(int64_t keep_bytes); /* drop cached blocks down to keep_bytes */
/* page-lock / unlock a caller-owned array in place (cudaHostRegister): the host pipeline
 * then DMAs straight out of it.  The shim registers large inputs it sees repeatedly and
 * unregisters them when the array is garbage-collected. */
ST_API int st_host_register(const void *p, int64_t bytes);
ST_API int st_host_unregister(const void *p);
ST_API int st_host_is_pinned(const void *p);

/* ---- link-list handle: the device-resident form of SuchLinkedTrees.linklist
 *      (MuchTree.pyx:2839-2874), built once and reused by every call below (the entry points
 * above that take the raw `linklist` build one per call).  Both trees must outlive it. */
typedef struct st_links st_links;
ST_API int st_links_create(const st_tree *tree_a, const st_tree *tree_b, const int64_t *linklist,
                    int64_t n_links, st_links **out);
ST_API void st_links_destroy(st_links *links);
/* = st_linked_distances / st_sample_linked_cycle on the handle */
ST_API int st_links_linked_distances(const st_links *links, double *out_a, double *out_b, int64_t *ids_a,
                              int64_t *ids_b);
ST_API int st_links_sample_cycle(const st_links *links, uint64_t *seed, int32_t buckets, int32_t n,
                          double *out_a, double *out_b, double *sums_a, double *sumsq_a,
                          double *sums_b, double *sumsq_b);
/* = st_sample_moments / st_linked_moments on the handle, plus the path's ONE collective:
 * nccl_comm != NULL (an ncclComm_t, e.g. from st_nccl_comm_create) -> the six sums
 * {n, sx, sy, sxx, syy, sxy} are all-reduced (ncclAllReduce, sum, fp64, in place) on the
 * stream the moment kernel ran on, before the one device->host read; every rank of the
 * communicator must make the call, with the same x0, y0.  *out then holds the moments of
 * the union of all ranks' samples / pairs. */
ST_API int st_links_sample_moments(const st_links *links, uint64_t seed, int64_t first_sample,
                            int64_t n_samples, double x0, double y0, void *nccl_comm, st_moments *out);
ST_API int st_links_linked_moments(const st_links *links, int64_t first_pair, int64_t n_pairs, double x0,
                            double y0, void *nccl_comm, st_moments *out);

/* = st_clade_moments on the handle: the links sorted by the scanned side's id and their
 * prefix counts are built once per handle and side and kept on the device; the per-call plan
 * (runs, eligible clades, work items) is made by kernels, not on the host. */
ST_API int st_links_clade_moments(const st_links *links, int side, const int64_t *clade_lo,
                           const int64_t *clade_hi, int64_t n_clades, int64_t min_links,
                           int64_t max_links, st_moments *out, int64_t *n_links_out);

/* ---- NCCL plumbing for the moment all-reduce (libnccl.so.2 is dlopen'ed on first use; a
 *      single-GPU process never loads it).  id128: 128 bytes (ncclUniqueId) produced on one
 * rank by st_nccl_unique_id and handed to the others by whatever rendezvous the host has
 * (torch.distributed broadcast, MPI, a file).  *comm is an ncclComm_t. */
ST_API int st_nccl_version(int *version);
ST_API int st_nccl_unique_id(void *id128);
ST_API int st_nccl_comm_create(int device, int world, int rank, const void *id128, void **comm);
ST_API int st_nccl_comm_destroy(void *comm);

#ifdef __cplusplus
}
#endif
#endif /* SUCHTREE_B200_H */
