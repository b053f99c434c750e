/*
 * CPU ORACLE -- test infrastructure only (see gs_oracle_impl.h header).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may use it.
 * Build: oracle/Makefile  ->  oracle/liboracle.so   (gcc -O2 -fopenmp -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16 /* config.h:15-17 BLOCK_X = BLOCK_Y = 16 */

/* Stable LSD radix sort, 16-bit digits. */
static void gso_sort_pairs(uint64_t* k, uint32_t* v, size_t n, int bits) {
    if (n < 2) return;
    uint64_t* k2 = (uint64_t*)malloc(n * sizeof(uint64_t));
    uint32_t* v2 = (uint32_t*)malloc(n * sizeof(uint32_t));
    size_t* hist = (size_t*)malloc(65536 * sizeof(size_t));
    uint64_t *ka = k, *kb = k2;
    uint32_t *va = v, *vb = v2;
    for (int shift = 0; shift < bits; shift += 16) {
        memset(hist, 0, 65536 * sizeof(size_t));
        for (size_t i = 0; i < n; i++) hist[(ka[i] >> shift) & 0xFFFF]++;
        size_t run = 0;
        for (int d = 0; d < 65536; d++) { size_t c = hist[d]; hist[d] = run; run += c; }
        for (size_t i = 0; i < n; i++) {
            size_t pos = hist[(ka[i] >> shift) & 0xFFFF]++;
            kb[pos] = ka[i]; vb[pos] = va[i];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != k) { memcpy(k, ka, n * sizeof(uint64_t)); memcpy(v, va, n * sizeof(uint32_t)); }
    free(k2); free(v2); free(hist);
}

#define REAL float
#define SUFFIX f32
#define IS_F32 1
#include "gs_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef IS_F32

#define REAL double
#define SUFFIX f64
#define IS_F32 0
#include "gs_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef IS_F32

int gso_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void gso_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
