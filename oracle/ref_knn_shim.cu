// C-ABI shim over the reference's SimpleKNN::knn (submodules/simple-knn/simple_knn.h:15-19), compiled together with
// the reference's own simple_knn.cu from where it lies (never copied).  TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime.h>
#include "simple_knn.h"

extern "C" int ref_knn(int P, const float* points_dev, float* mean_dist2_dev) {
    SimpleKNN::knn(P, (float3*)points_dev, mean_dist2_dev);
    return (int)cudaDeviceSynchronize();
}
