/* CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py) for the neighbour searches
 * the reference takes from pytorch-cluster 1.6.1 (environment.yml:18; not vendored under
 * /root/reference).  It restates the *CUDA* kernels' published semantics
 * (SURVEY.md App. B.1), which is what the reference runs on GPU:
 *
 *   radius  <- torch_cluster radius_kernel: one scan per query y_j over the same-example
 *              slice of x in ascending index, squared distance accumulated over d=0,1,2 in
 *              fp32 with FMA contraction (nvcc default), strict `<`, first max_num_neighbors
 *              hits kept.  Call sites: models/all_atom_score_model.py:545-550,563,607,627.
 *   knn     <- torch_cluster knn_kernel: insertion-sorted best_dist[k] initialised to 1e10,
 *              strict `>` on insert (ties keep the lower index first).
 *              Call site: models/all_atom_score_model.py:524 (through knn_graph).
 *
 * Build: gcc -O2 -shared -fPIC -ffp-contract=off oracle/cluster_c.c -o oracle/_build/libcluster_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>

static inline float sqdist(const float *a, const float *b) {
    float dist = 0.f;
    for (int d = 0; d < 3; ++d) {
        float t = a[d] - b[d];
        dist = fmaf(t, t, dist);
    }
    return dist;
}

/* returns number of edges written; row = index into y, col = index into x */
int64_t oracle_radius(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
                      int64_t num_examples, float r2, int64_t max_num_neighbors,
                      int64_t *row, int64_t *col) {
    int64_t e = 0;
    for (int64_t b = 0; b < num_examples; ++b) {
        for (int64_t j = ptr_y[b]; j < ptr_y[b + 1]; ++j) {
            int64_t count = 0;
            for (int64_t i = ptr_x[b]; i < ptr_x[b + 1]; ++i) {
                if (sqdist(x + 3 * i, y + 3 * j) < r2) {
                    row[e] = j;
                    col[e] = i;
                    ++e;
                    ++count;
                }
                if (count >= max_num_neighbors) break;
            }
        }
    }
    return e;
}

/* writes exactly k entries per query (col = -1 where fewer than k candidates exist) */
void oracle_knn(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
                int64_t num_examples, int64_t k, int64_t *row, int64_t *col) {
    float best_dist[100];
    int64_t best_idx[100];
    for (int64_t b = 0; b < num_examples; ++b) {
        for (int64_t j = ptr_y[b]; j < ptr_y[b + 1]; ++j) {
            for (int64_t e = 0; e < k; ++e) { best_dist[e] = 1e10f; best_idx[e] = -1; }
            for (int64_t i = ptr_x[b]; i < ptr_x[b + 1]; ++i) {
                float t = sqdist(x + 3 * i, y + 3 * j);
                for (int64_t e1 = 0; e1 < k; ++e1) {
                    if (best_dist[e1] > t) {
                        for (int64_t e2 = k - 1; e2 > e1; --e2) {
                            best_dist[e2] = best_dist[e2 - 1];
                            best_idx[e2] = best_idx[e2 - 1];
                        }
                        best_dist[e1] = t;
                        best_idx[e1] = i;
                        break;
                    }
                }
            }
            for (int64_t e = 0; e < k; ++e) { row[j * k + e] = j; col[j * k + e] = best_idx[e]; }
        }
    }
}
