/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the
 * shipped product path (suchtree_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as
 * the checker or the reported CPU baseline.
 *
 * Plain-C restatement of the reference's hot path.  Each function cites the
 * reference lines it follows (paths relative to /root/reference).  Parity is
 * PINNED: tests/test_oracle_golden.py checks these functions against
 *   (1) the reference's golden vector SuchTree/tests/test.matrix,
 *   (2) outputs of the unmodified reference itself (oracle/_ref, built by
 *       oracle/build_ref.py) committed under tests/golden/,
 *   (3) the published notebook vectors (linklist, pearson r).
 *
 * Trees arrive as the reference's Node fields split into arrays:
 *   parent[i], dist[i] (fp32, root = -1)      MuchTree.pyx:55-60
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <complex.h>

/* ---- MuchTree.pyx:999-1030  _mrca ------------------------------------------
 * Record a's ancestor chain in visited[], then walk b upward, scanning the
 * chain linearly at every step.  Returns -1 when no common ancestor exists. */
int oracle_mrca(const int32_t *parent, int64_t *visited, int a, int b)
{
    int n = a, i = 0, mrca = -1, a_depth;
    for (;;) {
        visited[i] = n;
        n = parent[n];
        i += 1;
        if (n == -1) break;
    }
    a_depth = i;
    n = b;
    for (;;) {
        i = 0;
        for (;;) {
            if (i >= a_depth) break;
            if (visited[i] == n) { mrca = (int)visited[i]; break; }
            i += 1;
        }
        if (mrca != -1) break;
        n = parent[n];
        if (n == -1) { mrca = n; break; }
    }
    return mrca;
}

void oracle_mrca_bulk(const int32_t *parent, int64_t *visited,
                      const int64_t *ids, uint64_t n, int32_t *out)
{
    for (uint64_t i = 0; i < n; ++i)
        out[i] = oracle_mrca(parent, visited, (int)ids[2 * i], (int)ids[2 * i + 1]);
}

/* ---- MuchTree.pyx:911-943  _distances, literal arithmetic (oracle O1') -----
 * fp32 accumulator `cdef float d` (:924), a->mrca first, then b->mrca, widened
 * to fp64 on store (:943). */
void oracle_distances_f32(const int32_t *parent, const float *dist, int64_t *visited,
                          const int64_t *ids, uint64_t n, double *result)
{
    for (uint64_t i = 0; i < n; ++i) {
        unsigned a = (unsigned)ids[2 * i], b = (unsigned)ids[2 * i + 1];
        unsigned mrca = (unsigned)oracle_mrca(parent, visited, (int)a, (int)b);
        unsigned k = a;
        float d = 0;
        while (k != mrca) { d += dist[k]; k = (unsigned)parent[k]; }
        k = b;
        while (k != mrca) { d += dist[k]; k = (unsigned)parent[k]; }
        result[i] = d;
    }
}

/* ---- same path summation, fp64 accumulator over the fp32-quantised edges
 * (oracle O2: "the reference's path summation" of BASELINE.json north_star;
 * the 1e-12 tolerance is stated against this one).  Also returns, when asked,
 * the L1 norm of the path (sum |edge|) so that tests can state tolerances for
 * trees with negative branch lengths (data/bigtrees/nj.tree has 1292). */
void oracle_distances_f64(const int32_t *parent, const float *dist, int64_t *visited,
                          const int64_t *ids, uint64_t n, double *result, double *l1_or_null)
{
    for (uint64_t i = 0; i < n; ++i) {
        unsigned a = (unsigned)ids[2 * i], b = (unsigned)ids[2 * i + 1];
        unsigned mrca = (unsigned)oracle_mrca(parent, visited, (int)a, (int)b);
        unsigned k = a;
        double d = 0, l1 = 0;
        while (k != mrca) { d += (double)dist[k]; l1 += fabs((double)dist[k]); k = (unsigned)parent[k]; }
        k = b;
        while (k != mrca) { d += (double)dist[k]; l1 += fabs((double)dist[k]); k = (unsigned)parent[k]; }
        result[i] = d;
        if (l1_or_null) l1_or_null[i] = l1;
    }
}

/* ---- MuchTree.pyx:1331-1376  _quartet_topologies ---------------------------
 * Six MRCAs per quartet (a,b),(a,c),(a,d),(b,c),(b,d),(c,d); C[j] = how many of
 * the six equal M[j]; j = first index with C[j] == 1 (5 if there is none: the
 * loop variable keeps its last value); row i of the output is the quartet
 * permuted by row j of the table I (:1319-1320; named PERM here, complex.h owns I).  `cdef int a, b, c, d` truncates the ids. */
void oracle_quartet_topologies(const int32_t *parent, int64_t *visited, const int64_t *quartets,
                               uint64_t n, int64_t *topologies)
{
    static const int PERM[6][4] = {{0, 1, 2, 3}, {0, 2, 1, 3}, {0, 3, 1, 2},
                                {1, 2, 0, 3}, {1, 3, 0, 2}, {2, 3, 0, 1}};
    int64_t M[6], C[6];
    for (uint64_t i = 0; i < n; ++i) {
        int a = (int)quartets[4 * i], b = (int)quartets[4 * i + 1];
        int c = (int)quartets[4 * i + 2], d = (int)quartets[4 * i + 3];
        int j, k;
        M[0] = oracle_mrca(parent, visited, a, b);
        M[1] = oracle_mrca(parent, visited, a, c);
        M[2] = oracle_mrca(parent, visited, a, d);
        M[3] = oracle_mrca(parent, visited, b, c);
        M[4] = oracle_mrca(parent, visited, b, d);
        M[5] = oracle_mrca(parent, visited, c, d);
        for (j = 0; j < 6; ++j) C[j] = 0;
        for (j = 0; j < 6; ++j)
            for (k = 0; k < 6; ++k)
                if (M[j] == M[k]) C[j] = C[j] + 1;
        for (j = 0; j < 6; ++j)
            if (C[j] == 1) break;
        if (j == 6) j = 5; /* Cython's `for j in range(6)` leaves j at 5 without a break */
        for (k = 0; k < 4; ++k) topologies[4 * i + k] = quartets[4 * i + PERM[j][k]];
    }
}

/* ---- NOT a restatement: O(depth) MRCA by depth-levelled climbing, for trees
 * on which the reference's O(depth^2) scan cannot finish (10^6-deep
 * caterpillar).  Cross-checked against oracle_mrca on every
 * small tree in tests/test_oracle_golden.py. */
int oracle_mrca_climb(const int32_t *parent, const int32_t *depth, int a, int b)
{
    while (depth[a] > depth[b]) a = parent[a];
    while (depth[b] > depth[a]) b = parent[b];
    while (a != b) {
        a = parent[a];
        b = parent[b];
        if (a == -1 || b == -1) return -1;
    }
    return a;
}

void oracle_distances_f64_climb(const int32_t *parent, const int32_t *depth, const float *dist,
                                const int64_t *ids, uint64_t n, double *result, int32_t *mrca_or_null)
{
    for (uint64_t i = 0; i < n; ++i) {
        int a = (int)ids[2 * i], b = (int)ids[2 * i + 1];
        int m = oracle_mrca_climb(parent, depth, a, b);
        double d = 0;
        for (int k = a; k != m; k = parent[k]) d += (double)dist[k];
        for (int k = b; k != m; k = parent[k]) d += (double)dist[k];
        result[i] = d;
        if (mrca_or_null) mrca_or_null[i] = m;
    }
}

/* ---- MuchTree.pyx:218-225  depth = max #nodes on a leaf->root path --------- */
unsigned oracle_tree_depth(const int32_t *parent, const int64_t *leaf_ids, uint64_t n_leaves)
{
    unsigned depth = 0;
    for (uint64_t i = 0; i < n_leaves; ++i) {
        int node = (int)leaf_ids[i];
        unsigned n = 1;
        for (;;) {
            if (parent[node] == -1) break;
            node = parent[node];
            n += 1;
        }
        if (n > depth) depth = n;
    }
    return depth;
}

/* per-node depth (root = 0), helper for oracle_mrca_climb */
void oracle_node_depths(const int32_t *parent, uint64_t n_nodes, int32_t *depth)
{
    for (uint64_t i = 0; i < n_nodes; ++i) {
        int32_t d = 0;
        for (int k = (int)i; parent[k] != -1; k = parent[k]) d++;
        depth[i] = d;
    }
}

/* ---- MuchTree.pyx:62-79  _pearson, literal (fp32 accumulators :66-67,
 * +1e-20 guard :79). */
double oracle_pearson_f32(const double *x, const double *y, unsigned n)
{
    float yt, xt;
    float syy = 0.0f, sxy = 0.0f, sxx = 0.0f, ay = 0.0f, ax = 0.0f;
    for (unsigned long j = 0; j < n; ++j) { ax += x[j]; ay += y[j]; }
    ax /= n;
    ay /= n;
    for (unsigned long j = 0; j < n; ++j) {
        xt = x[j] - ax;
        yt = y[j] - ay;
        sxx += xt * xt;
        syy += yt * yt;
        sxy += xt * yt;
    }
    /* Cython compiles `x**(0.5)` on a C double to a COMPLEX power and quotient
     * (MuchTree.c:22164-22172: __Pyx_c_pow_double = cpow, __Pyx_c_quot_double), which
     * can differ from sqrt() in the last bit -- restated literally. */
    double complex den = cpow(((double)(sxx * syy) + 1.0e-20) + 0.0 * I, 0.5 + 0.0 * I);
    return creal((sxy + 0.0 * I) / den);
}

/* same two-pass formula with fp64 accumulators (what the fp32 code rounds) */
double oracle_pearson_f64(const double *x, const double *y, uint64_t n)
{
    double ax = 0, ay = 0, sxx = 0, syy = 0, sxy = 0;
    for (uint64_t j = 0; j < n; ++j) { ax += x[j]; ay += y[j]; }
    ax /= (double)n;
    ay /= (double)n;
    for (uint64_t j = 0; j < n; ++j) {
        double xt = x[j] - ax, yt = y[j] - ay;
        sxx += xt * xt;
        syy += yt * yt;
        sxy += xt * yt;
    }
    return sxy / sqrt(sxx * syy + 1.0e-20);
}

/* ---- MuchTree.pyx:2918-2925  linked_distances pair enumeration -------------
 * linklist rows are [TreeB leaf id, TreeA leaf id] (:2869-2870). */
void oracle_linked_pairs(const int64_t *linklist, uint64_t n_links, int64_t *ids_a, int64_t *ids_b)
{
    uint64_t k = 0;
    for (uint64_t i = 0; i < n_links; ++i)
        for (uint64_t j = 0; j < i; ++j) {
            ids_a[2 * k + 1] = linklist[2 * i + 1];
            ids_a[2 * k + 0] = linklist[2 * j + 1];
            ids_b[2 * k + 1] = linklist[2 * i + 0];
            ids_b[2 * k + 0] = linklist[2 * j + 0];
            k += 1;
        }
}

/* ---- MuchTree.pyx:2936-2949  _random_int (xorshift64*, modulus :2573) ------ */
uint64_t oracle_random_int(uint64_t *seed, uint64_t n)
{
    *seed ^= *seed >> 12;
    *seed ^= *seed << 25;
    *seed ^= *seed >> 27;
    return (*seed * 2685821657736338717ULL) % n;
}

/* ---- MuchTree.pyx:3026-3038  one bucket of sampled link pairs --------------
 * l1, l2 are `cdef int` (:3009-3010). */
void oracle_sample_bucket(uint64_t *seed, const int64_t *linklist, uint64_t n_links,
                          unsigned n, int64_t *query_a, int64_t *query_b)
{
    for (unsigned j = 0; j < n; ++j) {
        int l1 = (int)oracle_random_int(seed, n_links);
        int l2 = (int)oracle_random_int(seed, n_links);
        query_a[2 * j + 0] = linklist[2 * l1 + 1];
        query_a[2 * j + 1] = linklist[2 * l2 + 1];
        query_b[2 * j + 0] = linklist[2 * l1 + 0];
        query_b[2 * j + 1] = linklist[2 * l2 + 0];
    }
}
