/* Plain-C restatement of scipy.optimize.linear_sum_assignment (oracle; TEST
 * INFRASTRUCTURE, never linked into the product library).
 *
 * Call site in the reference: src/tracker/core/linear_assignment.py:62.
 * scipy is not vendored there; this restates the published algorithm
 * (D. F. Crouse, IEEE T-AES 52(4), 2016) with scipy's scan order and tie rule:
 *   - rows <= cols after an optional transpose,
 *   - `remaining` starts reversed (nc-1 .. 0) and shrinks by swap-removal,
 *   - relaxation uses strict `<`,
 *   - among equal minima the last unassigned column in scan order wins, else
 *     the first minimum.
 * All arithmetic is float64.  Pinned against scipy in tests/test_oracle_lsap.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int lsap_ref_solve(int nr_in, int nc_in, const double *cost_in, int64_t *rows_out, int64_t *cols_out)
{
    int nr = nr_in, nc = nc_in;
    if (nr == 0 || nc == 0) return 0;
    const int transpose = nc < nr;
    double *cost = (double *)malloc(sizeof(double) * (size_t)nr * (size_t)nc);
    if (!cost) return -1;
    if (transpose) {
        for (int i = 0; i < nr_in; ++i)
            for (int j = 0; j < nc_in; ++j)
                cost[(size_t)j * nr_in + i] = cost_in[(size_t)i * nc_in + j];
        nr = nc_in; nc = nr_in;
    } else {
        memcpy(cost, cost_in, sizeof(double) * (size_t)nr * (size_t)nc);
    }
    double *u = (double *)calloc((size_t)nr, sizeof(double));
    double *v = (double *)calloc((size_t)nc, sizeof(double));
    double *spc = (double *)malloc(sizeof(double) * (size_t)nc);
    int *path = (int *)malloc(sizeof(int) * (size_t)nc);
    int *col4row = (int *)malloc(sizeof(int) * (size_t)nr);
    int *row4col = (int *)malloc(sizeof(int) * (size_t)nc);
    int *remaining = (int *)malloc(sizeof(int) * (size_t)nc);
    unsigned char *SR = (unsigned char *)malloc((size_t)nr);
    unsigned char *SC = (unsigned char *)malloc((size_t)nc);
    int rc = 0;
    for (int i = 0; i < nr; ++i) col4row[i] = -1;
    for (int j = 0; j < nc; ++j) { row4col[j] = -1; path[j] = -1; }

    for (int cur = 0; cur < nr && rc == 0; ++cur) {
        double min_val = 0.0;
        int i = cur, num_remaining = nc, sink = -1;
        memset(SR, 0, (size_t)nr);
        memset(SC, 0, (size_t)nc);
        for (int it = 0; it < nc; ++it) { remaining[it] = nc - 1 - it; spc[it] = INFINITY; }
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < num_remaining; ++it) {
                const int j = remaining[it];
                const double r = min_val + cost[(size_t)i * nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
                    lowest = spc[j];
                    index = it;
                }
            }
            min_val = lowest;
            if (min_val == INFINITY) { rc = -2; break; }
            const int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (rc) break;
        u[cur] += min_val;
        for (int i2 = 0; i2 < nr; ++i2)
            if (SR[i2] && i2 != cur) u[i2] += min_val - spc[col4row[i2]];
        for (int j2 = 0; j2 < nc; ++j2)
            if (SC[j2]) v[j2] -= min_val - spc[j2];
        int j = sink;
        for (;;) {
            const int ii = path[j];
            row4col[j] = ii;
            const int t = col4row[ii];
            col4row[ii] = j;
            j = t;
            if (ii == cur) break;
        }
    }
    if (rc == 0) {
        if (transpose) {
            /* pairs (col4row[k], k) sorted by first element; col4row is a partial
             * permutation, so a counting pass over original rows is a stable sort */
            int *inv = (int *)malloc(sizeof(int) * (size_t)nc);
            for (int j = 0; j < nc; ++j) inv[j] = -1;
            for (int k = 0; k < nr; ++k) inv[col4row[k]] = k;
            int o = 0;
            for (int j = 0; j < nc; ++j)
                if (inv[j] >= 0) { rows_out[o] = j; cols_out[o] = inv[j]; ++o; }
            free(inv);
        } else {
            for (int k = 0; k < nr; ++k) { rows_out[k] = k; cols_out[k] = col4row[k]; }
        }
    }
    free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col);
    free(remaining); free(SR); free(SC);
    return rc;
}
