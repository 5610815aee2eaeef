// RefineNet (75 -> 160 -> 256 -> 256 -> 128 -> 45 MLP, BatchNorm folded) fused with the input assembly and the root
// re-addition of lift_and_refine_3d_pose (exps/stage3_root2/test_util.py:102-131, model/refinenet.py:5-38).
//
// Latency-bound: <= 127 persons per frame, 157 k MAC per person, 628 KB of fp32 weights that stay in L2.  One CTA takes
// RF_PP persons (weights are read once per CTA and reused across its persons), thread n owns output neuron n of the
// current layer and walks k in ascending order with one fmaf per step (fixed summation order -> run-to-run identical).
#include "refine.h"

#include "assoc.h"

namespace smapb {

constexpr int RF_PP = 4;         // persons per CTA
constexpr int RF_THREADS = 256;  // >= widest layer
__constant__ int c_rf_dims[RF_LAYERS + 1] = {75, 160, 256, 256, 128, 45};

template <int MODE>  // 0: raw rows in/out, 1: records
__global__ void __launch_bounds__(RF_THREADS) refine_kernel(RefineWeights W, const float* __restrict__ in, int n_rows,
                                                             float* __restrict__ out, const float* pred2d, const double* pred3d,
                                                             const int* __restrict__ counts, int root, long long s2d,
                                                             long long s3d, long long sc, double* out3d, long long so) {
    __shared__ float act[2][RF_PP][256];
    __shared__ double root3[RF_PP][3];
    __shared__ int root_scored[RF_PP];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * RF_PP;
    int np;
    const float* P2 = nullptr;
    const double* P3 = nullptr;
    if (MODE == 0) {
        np = min(RF_PP, n_rows - row0);
    } else {
        const int b = blockIdx.y;
        const int cnt = min(counts[(long long)b * sc], MAXP);
        if (row0 >= cnt) return;
        np = min(RF_PP, cnt - row0);
        P2 = pred2d + (long long)b * s2d + (long long)row0 * (NJ * 4);
        P3 = pred3d + (long long)b * s3d + (long long)row0 * (NJ * 4);
    }
    // ---- stage the inputs (test_util.py:106-114: fp64 assembly, then .float()) ----
    for (int idx = tid; idx < RF_PP * RF_IN; idx += RF_THREADS) {
        const int p = idx / RF_IN, e = idx - p * RF_IN;
        float v = 0.f;
        if (p < np) {
            if (MODE == 0) {
                v = in[(long long)(row0 + p) * RF_IN + e];
            } else {
                const int j = e / 5, c = e - j * 5;
                const float* q2 = P2 + p * (NJ * 4);
                const double* q3 = P3 + p * (NJ * 4);
                if (j == root)
                    v = c < 2 ? q2[root * 4 + c] : (float)q3[root * 4 + c - 2];
                else if (q3[j * 4 + 3] > 0.0)
                    v = c < 2 ? __fsub_rn(q2[j * 4 + c], q2[root * 4 + c])          // float32 - float32 (numpy float32 arrays)
                              : (float)(q3[j * 4 + c - 2] - q3[root * 4 + c - 2]);  // float64 - float64, then .float()
            }
        }
        act[0][p][e] = v;
    }
    if (MODE == 1 && tid < RF_PP * 4) {
        const int p = tid >> 2, c = tid & 3;
        if (p < np) {
            const double* q3 = P3 + p * (NJ * 4);
            if (c < 3) root3[p][c] = q3[root * 4 + c];
            else root_scored[p] = q3[root * 4 + 3] != 0.0;
        }
    }
    __syncthreads();
    // ---- the five layers ----
#pragma unroll 1
    for (int l = 0; l < RF_LAYERS; l++) {
        const int K = c_rf_dims[l], N = c_rf_dims[l + 1];
        const float(*src)[256] = act[l & 1];
        float(*dst)[256] = act[(l + 1) & 1];
        if (tid < N) {
            const float* w = W.w[l] + tid;
            float acc[RF_PP];
            const float bias = W.b[l][tid];
#pragma unroll
            for (int p = 0; p < RF_PP; p++) acc[p] = bias;
#pragma unroll 8
            for (int k = 0; k < K; k++) {
                const float wk = __ldg(w + (long long)k * N);
#pragma unroll
                for (int p = 0; p < RF_PP; p++) acc[p] = fmaf(wk, src[p][k], acc[p]);
            }
#pragma unroll
            for (int p = 0; p < RF_PP; p++) dst[p][tid] = (l < RF_LAYERS - 1) ? fmaxf(acc[p], 0.f) : acc[p];
        }
        __syncthreads();
    }
    const float(*res)[256] = act[RF_LAYERS & 1];
    if (MODE == 0) {
        for (int idx = tid; idx < np * RF_OUT; idx += RF_THREADS) {
            const int p = idx / RF_OUT, e = idx - p * RF_OUT;
            out[(long long)(row0 + p) * RF_OUT + e] = res[p][e];
        }
    } else {
        // test_util.py:122-130: pred[i,j] += root (float32 += float64: added in double, stored as float32), the root row
        // is the lifted root itself; score column 1 unless the root was unscored
        double* O = out3d + (long long)blockIdx.y * so + (long long)row0 * (NJ * 4);
        for (int idx = tid; idx < np * NJ * 4; idx += RF_THREADS) {
            const int p = idx / (NJ * 4), r = idx - p * (NJ * 4);
            const int j = r >> 2, c = r & 3;
            double v;
            if (c == 3) v = root_scored[p] ? 1.0 : 0.0;
            else if (j == root) v = (double)(float)root3[p][c];
            else v = (double)(float)((double)res[p][j * 3 + c] + root3[p][c]);
            O[p * (NJ * 4) + r] = v;
        }
    }
}

cudaError_t launch_refine_mlp(const RefineWeights& W, const float* in, int n, float* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    refine_kernel<0><<<(n + RF_PP - 1) / RF_PP, RF_THREADS, 0, st>>>(W, in, n, out, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0);
    return cudaGetLastError();
}

cudaError_t launch_refine_records(const RefineWeights& W, const float* pred2d, const double* pred3d, const int* counts, int B,
                                  int root_idx, long long s2d, long long s3d, long long sc, double* out, long long so,
                                  cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    dim3 grid((MAXP + RF_PP - 1) / RF_PP, B);
    refine_kernel<1><<<grid, RF_THREADS, 0, st>>>(W, nullptr, 0, nullptr, pred2d, pred3d, counts, root_idx, s2d, s3d, sc, out, so);
    return cudaGetLastError();
}

}  // namespace smapb
