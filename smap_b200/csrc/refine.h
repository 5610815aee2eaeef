// RefineNet post-processing: model/refinenet.py + exps/stage3_root2/test_util.py:102-131.
#pragma once
#include <cuda_runtime.h>

namespace smapb {

constexpr int RF_LAYERS = 5;
constexpr int RF_IN = 75, RF_OUT = 45;

// BN-folded, transposed ([K][N], k-major) fp32 weights and biases of the five Linear layers (device pointers)
struct RefineWeights {
    const float* w[RF_LAYERS];
    const float* b[RF_LAYERS];
};

// raw network: in fp32 [n,75] -> out fp32 [n,45]   (what `refine_model(inp)` computes, model/refinenet.py:19-26)
cudaError_t launch_refine_mlp(const RefineWeights& W, const float* in, int n, float* out, cudaStream_t st);

// lift_and_refine_3d_pose for a batch of images: person p of image b reads pred2d[b*s2d + p*60 ..] (fp32 [15,4]),
// pred3d[b*s3d + p*60 ..] (fp64 [15,4]) for p < counts[b*sc], and writes the refined fp64 [15,4] (X,Y,Z,score) to
// out[b*so + p*60 ..].  out may alias pred3d (a person's inputs are staged before anything is written).
cudaError_t launch_refine_records(const RefineWeights& W, const float* pred2d, const double* pred3d, const int* counts, int B,
                                  int root_idx, long long s2d, long long s3d, long long sc, double* out, long long so,
                                  cudaStream_t st);

}  // namespace smapb
