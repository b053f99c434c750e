/* CPU oracle of simple_knn.distCUDA2 -- TEST INFRASTRUCTURE ONLY (see oracle/README in DESIGN.md section 3).
 *
 * Restates WHAT SimpleKNN::knn computes (submodules/simple-knn/simple_knn.cu:147-183): for every point the mean of the
 * squared distances to its 3 nearest OTHER points (by index: duplicates at distance 0 count), with the reference's
 * float arithmetic: d = q - p per axis, dist = fma(dz, dz, fma(dx, dx, dy*dy)) (nvcc's contraction of :138, read off the SASS of the reference
 * build: FMUL on y, FFMA with x, FFMA with z; pinned by tests/golden/knn_golden.npz), the three
 * smallest kept in ascending order (updateKBest :135-146), result (b0 + b1 + b2) / 3.0f (:182); fewer than 3 other
 * points leave FLT_MAX entries.  The search itself is brute force, O(P^2): independent of the Morton/box machinery,
 * which only accelerates the reference and cannot change the 3 smallest values. */
#include <float.h>
#include <math.h>
#include <stdint.h>

void gso_knn_mean_dist2(int64_t P, const float* pts, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < P; i++) {
        const float px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
        float b[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        for (int64_t j = 0; j < P; j++) {
            if (j == i) continue;
            const float dx = pts[3 * j] - px, dy = pts[3 * j + 1] - py, dz = pts[3 * j + 2] - pz;
            float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            for (int k = 0; k < 3; k++)
                if (b[k] > d) { const float t = b[k]; b[k] = d; d = t; }
        }
        out[i] = (b[0] + b[1] + b[2]) / 3.0f;
    }
}
