/* host/coo2csc.c — stable counting sort of coordinate entries by their `col_coo` key.
 * Same contract as final/coo2csc.c:22-64: on return col[0..n] are the pointers (col[n] = nnz) and
 * row[col[c] .. col[c+1]) are the row_coo values of the entries with col_coo == c, in input order. */
#include "coo2csc.h"
#include <stdlib.h>
#include <string.h>

void coo2csc(uint32_t *const row, uint32_t *const col,
             uint32_t const *const row_coo, uint32_t const *const col_coo,
             uint32_t const nnz, uint32_t const n, uint32_t const isOneBased)
{
    /* histogram shifted by one so that the inclusive running sum leaves the START of bucket c in col[c] */
    memset(col, 0, ((size_t)n + 1) * sizeof(uint32_t));
    for (uint32_t e = 0; e < nnz; ++e) {
        uint32_t c = col_coo[e] - isOneBased;
        if (c + 1 <= n) col[c + 1]++;                   /* entries outside [0,n) are dropped, never written out of bounds */
    }
    for (uint32_t c = 0; c < n; ++c) col[c + 1] += col[c];

    /* a private cursor array keeps col[] intact, so no "shift back" pass is needed afterwards */
    uint32_t *cursor = (uint32_t *)malloc(((size_t)n + 1) * sizeof(uint32_t));
    if (!cursor) abort();
    memcpy(cursor, col, ((size_t)n + 1) * sizeof(uint32_t));
    for (uint32_t e = 0; e < nnz; ++e) {
        uint32_t c = col_coo[e] - isOneBased;
        if (c < n) row[cursor[c]++] = row_coo[e] - isOneBased;
    }
    free(cursor);
}
