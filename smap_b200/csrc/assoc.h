// Internal launch interface of the association kernels (assoc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smapb {
constexpr int NJ = 15;     // key points            (extensions/association.cpp:18)
constexpr int NL = 14;     // limbs                 (extensions/association.cpp:19)
constexpr int MAXP = 127;  // max peaks per channel (extensions/association.cpp:20)
constexpr int NC2D = 43;   // 15 + 2*14 channels of the 2D head

int assoc_configure(int h, int w, const char** err);
size_t nms_mask_words(int B, int h, int w);  // uint32 words of NMS scratch (one ballot bit per pixel of the 15 key-point planes)
cudaError_t launch_nms(const float* hms, int nchan, int B, int h, int w, float thr, float* peaks, uint32_t* masks,
                       cudaStream_t st);
cudaError_t launch_paf(const float* hms, int nchan, int B, int h, int w, const float* peaks, float* scores,
                       int dense_fill, cudaStream_t st);
cudaError_t launch_group(const float* peaks, const float* scores, const float* rdepth, int B, int h, int w,
                         int root_idx, int dist_flag, float* bodies, int* counts, cudaStream_t st);
cudaError_t launch_lift(const float* bodies, const int* counts, const float* det_d, const float* root_d,
                        const double* scales, int B, int h, int w, int root_n, float* pred2d, double* pred3d,
                        double* root_depth, int* counts_out, long long s2d, long long s3d, long long srd, long long scnt,
                        cudaStream_t st);
// register_pred WITH ground truth (test_util.py:21-39) + float64 lift; outputs in the natural [B,127,15,4] / [B,127] layout
cudaError_t launch_lift_gt(const float* bodies, const int* counts, const float* det_d, const float* root_d,
                           const double* scales, const double* gt_roots, const int* gt_counts, int gmax, double* dist_ws, int B,
                           int h, int w, int root_n, double* pred2d, double* pred3d, double* root_depth, int* counts_out,
                           cudaStream_t st);
}  // namespace smapb
