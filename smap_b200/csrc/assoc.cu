// Depth-aware part association on sm_100a: batched, device-resident, bit-exact with the reference.
//
//   nms_kernel    <- extensions/gpu/nmsBase.cu:10-135 (register + thrust scan + write, fused)
//   paf_kernel    <- extensions/gpu/bodyPartConnectorBase.cu:11-63,104-150
//   group_kernel  <- extensions/association.cpp:123-233 (findConnectedJoints; CPU in the reference)
//   lift_kernel   <- exps/stage3_root2/test_util.py:60-99 + lib/utils/post_3d.py:4-27 (numpy in the reference)
//
// Data movement: every heat-map / PAF plane is staged ONCE into shared memory with 1-D bulk async copies
// (cp.async.bulk -> UBLKCP, completion on an mbarrier); all neighbourhood / line-integral gathers then hit
// shared memory.  Floating-point expressions whose rounding feeds a comparison are pinned with explicit
// __f*_rn intrinsics in the contraction pattern of the reference's sm_100 binary.
#include "assoc.h"
#include "common.cuh"

namespace smapb {

__constant__ int c_joint_pairs[2 * NL] = {0, 1, 0, 2, 0, 9, 9, 10, 10, 11, 0, 3, 3, 4,
                                          4, 5, 2, 12, 12, 13, 13, 14, 2, 6, 6, 7, 7, 8};
// extensions/association.cpp:27-31 (vector<float> initialised from double literals)
__constant__ float c_bone_length[NL] = {26.42178982f, 48.36980909f, 14.88291009f, 31.28002332f, 23.915707f,
                                        14.97674918f, 31.28002549f, 23.91570732f, 12.4644364f, 48.26604433f,
                                        39.03553194f, 12.4644364f, 48.19076948f, 39.03553252f};

// ---------------------------------------------------------------------------------------------
// plane staging: warp 0 issues the bulk copies (one 16 KB chunk per lane per round, so up to 32 copies are in
// flight per CTA) and is the ONLY warp that polls the mbarrier; everyone else parks on the CTA barrier.  (Round 1
// had one thread issue 32 KB chunks while all 1024 threads spun on mbarrier.try_wait: the polling traffic slowed the
// very shared-memory writes it was waiting for - 15 GB/s per SM.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_planes(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, bytes);
        __syncwarp();
        const uint32_t CH = 16384;
        for (uint32_t off = threadIdx.x * CH; off < bytes; off += 32 * CH) {
            const uint32_t n = bytes - off < CH ? bytes - off : CH;
            bulk_g2s((char*)dst + off, (const char*)src + off, n, bar);
        }
        mbar_wait(bar, 0);
    }
    __syncthreads();
}

// ---- thread-block-cluster helpers (NMS row bands exchange their peak counts through distributed shared memory) ----
__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ int cl_ld_s32(const int* local, uint32_t rank) {
    uint32_t a;
    int v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------
// NMS: a CLUSTER of `bands` CTAs per (image, key-point channel); CTA r owns the rows [r*rows, (r+1)*rows) of the
// plane and stages them (+3 halo rows on each side for the 7x7 centroid) with one bulk copy - the rows are
// contiguous in memory.  grid (bands, NJ, B), cluster (bands, 1, 1), block NMS_THREADS.
// Peak order is raster order (required: candidate indices are part of the parity contract): warp ballots over
// contiguous per-warp pixel segments + a block scan of the warp totals give the order inside a band, and the bands
// exchange their totals through distributed shared memory (band r starts at the sum of bands < r).  This replaces the
// reference's global thrust::exclusive_scan (nmsBase.cu:165-166) and works for any map size (config 5: 256x256).
// ---------------------------------------------------------------------------------------------
constexpr int NMS_THREADS = 256;
constexpr int NMS_WARPS = NMS_THREADS / 32;
constexpr int NMS_MAX_BANDS = 8;  // portable cluster size

__global__ void __launch_bounds__(NMS_THREADS)
nms_kernel(const float* __restrict__ hms, int nchan, int h, int w, int rows, float thr, float* __restrict__ peaks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* s_cnt = reinterpret_cast<int*>(smem_raw + 8);     // this band's peak count (read by the other bands)
    int* warp_tot = reinterpret_cast<int*>(smem_raw + 16);  // [NMS_WARPS]
    float* tile = reinterpret_cast<float*>(smem_raw + 64);  // staged rows [r_lo, r_hi)
    const int band = (int)cl_rank(), bands = (int)gridDim.x;
    const int c = blockIdx.y, img = blockIdx.z;
    const int r0 = band * rows, r1 = min(h, r0 + rows);           // owned rows (may be empty for the last bands)
    const int r_lo = max(0, r0 - 3), r_hi = min(h, r1 + 3);       // staged rows
    const int n_own = max(0, r1 - r0) * w;
    const int nwords = (n_own + 31) / 32;
    uint32_t* masks = reinterpret_cast<uint32_t*>(tile + (size_t)(rows + 6) * w);  // one ballot word per 32 owned pixels
    const float* src = hms + ((size_t)img * nchan + c) * h * w;
    float* out = peaks + ((size_t)img * NJ + c) * (MAXP + 1) * 3;

    pdl_wait();
    if (r_hi > r_lo) stage_planes(tile, src + (size_t)r_lo * w, (uint32_t)((r_hi - r_lo) * w) * 4u, bar);
    const float* plane = tile - (size_t)r_lo * w;  // plane[y * w + x] for staged y

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpw = (nwords + NMS_WARPS - 1) / NMS_WARPS;  // contiguous run of ballot words per warp
    const int w0 = min(nwords, warp * wpw), w1 = min(nwords, w0 + wpw);
    const int base = r0 * w;
    int cnt = 0;
    for (int wi = w0; wi < w1; wi++) {
        const int i = base + wi * 32 + lane;
        bool f = false;
        if (wi * 32 + lane < n_own) {
            const int y = i / w, x = i - y * w;
            if (x > 0 && x < w - 1 && y > 0 && y < h - 1) {  // nmsBase.cu:24
                const float v = plane[i];
                if (v > thr) {
                    const float* q0 = plane + i - w;
                    const float* q2 = plane + i + w;
                    f = v > q0[-1] && v > q0[0] && v > q0[1] && v > plane[i - 1] && v > plane[i + 1] &&
                        v > q2[-1] && v > q2[0] && v > q2[1];
                }
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) masks[wi] = m;
        cnt += __popc(m);
    }
    if (lane == 0) warp_tot[warp] = cnt;
    __syncthreads();
    // exclusive prefix over the warps of this band
    int t = (lane < NMS_WARPS) ? warp_tot[lane] : 0;
    int before = (lane < warp) ? t : 0;
    int band_total = t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        band_total += __shfl_xor_sync(0xffffffffu, band_total, o);
    }
    if (threadIdx.x == 0) *s_cnt = band_total;
    cl_sync();  // every band's count is published
    int band_off = 0, total = 0;
    for (int r = 0; r < bands; r++) {
        const int v = (r == band) ? band_total : cl_ld_s32(s_cnt, (uint32_t)r);
        if (r < band) band_off += v;
        total += v;
    }
    cl_sync();  // nobody leaves (or reuses s_cnt) while a peer may still read it
    int running = band_off + before;
    for (int wi = w0; wi < w1; wi++) {
        const uint32_t m = masks[wi];
        if (m == 0) continue;
        if ((m >> lane) & 1u) {
            const int peakIndex = running + __popc(m & ((1u << lane) - 1u));
            if (peakIndex < MAXP) {  // nmsBase.cu:92
                const int i = base + wi * 32 + lane;
                const int py = i / w, px = i - py * w;
                float xAcc = 0.f, yAcc = 0.f, sAcc = 0.f;
                for (int dy = -3; dy <= 3; dy++) {
                    const int y = py + dy;
                    if (y < 0 || y >= h) continue;
                    for (int dx = -3; dx <= 3; dx++) {
                        const int x = px + dx;
                        if (x < 0 || x >= w) continue;
                        const float sc = plane[y * w + x];
                        if (sc > 0) {
                            xAcc = __fmaf_rn((float)x, sc, xAcc);  // FFMA in the reference SASS
                            yAcc = __fmaf_rn((float)y, sc, yAcc);
                            sAcc = __fadd_rn(sAcc, sc);
                        }
                    }
                }
                float* o = out + (peakIndex + 1) * 3;
                o[0] = __fadd_rn(__fdiv_rn(xAcc, sAcc), 0.5f);
                o[1] = __fadd_rn(__fdiv_rn(yAcc, sAcc), 0.5f);
                o[2] = plane[i];
            }
        }
        running += __popc(m);
    }
    if (band == 0) {
        const int count = total < MAXP ? total : MAXP;
        if (threadIdx.x == 0) {
            out[0] = (float)count;
            out[1] = 0.f;
            out[2] = 0.f;
        }
        // deterministic tail: slots the reference leaves uninitialised are zeroed
        for (int k = (count + 1) * 3 + threadIdx.x; k < (MAXP + 1) * 3; k += NMS_THREADS) out[k] = 0.f;
    }
    pdl_trigger();
}

// ---------------------------------------------------------------------------------------------
// PAF line-integral scoring: one CTA per (image, limb); both PAF planes (adjacent channels 15+2l, 16+2l)
// are staged once into shared memory, then one thread per (peakA, peakB) candidate.
// ---------------------------------------------------------------------------------------------
constexpr int PAF_THREADS = 1024;

__device__ __forceinline__ float paf_process(float ax, float ay, float bx, float by, const float* __restrict__ mapX,
                                             const float* __restrict__ mapY, int w, int h, float near_thr) {
    const float dx = __fsub_rn(bx, ax);
    const float dy = __fsub_rn(by, ay);
    const float dmax = fmaxf(fabsf(dx), fabsf(dy));
    int n = (int)__fadd_rn(__fsqrt_rn(__fmul_rn(5.f, dmax)), 0.5f);
    n = max(5, min(25, n));
    const float norm = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    if ((double)norm > 1e-6) {
        const float ux = __fdiv_rn(dx, norm), uy = __fdiv_rn(dy, norm);
        const float fn = (float)n;
        const float stepX = __fdiv_rn(dx, fn), stepY = __fdiv_rn(dy, fn);
        float sum = 0.f;
        int count = 0;
        for (int lm = 0; lm < n; lm++) {
            const float flm = (float)lm;
            int mX = (int)__fadd_rn(__fmaf_rn(flm, stepX, ax), 0.5f);
            int mY = (int)__fadd_rn(__fmaf_rn(flm, stepY, ay), 0.5f);
            mX = min(w - 1, mX);
            mY = min(h - 1, mY);
            // the reference applies no lower clamp (coordinates are >= 0.5 by construction); clamp to keep
            // the shared-memory access in range for adversarial inputs without changing valid results
            mX = max(0, mX);
            mY = max(0, mY);
            const int idx = mY * w + mX;
            const float score = __fmaf_rn(ux, mapX[idx], __fmul_rn(uy, mapY[idx]));
            if (score > 0.05f) {
                sum = __fadd_rn(sum, score);
                count++;
            }
        }
        if (__fdiv_rn((float)count, fn) > 0.95f) return __fdiv_rn(sum, (float)count);
        if (norm < near_thr) return (float)(0.1f + 1e-6);
    }
    return -1.f;
}

// STAGED: both planes fit in shared memory (the parity configuration 128x208: 213 KB) and are staged once; otherwise
// (larger maps, e.g. 256x256 at a 1024x1024 input) the line integrals gather straight from global memory / L2.
template <bool STAGED>
__global__ void __launch_bounds__(PAF_THREADS, 1)
paf_kernel(const float* __restrict__ hms, int nchan, int h, int w, const float* __restrict__ peaks,
           float* __restrict__ scores, int dense_fill) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int hw = h * w;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* pk = reinterpret_cast<float*>(smem_raw + 16);  // [2][MAXP+1][2]  (x, y) of joint A then joint B
    float* planes = pk + 4 * (MAXP + 1);                  // [2][hw] when STAGED (16-byte aligned: 16 + 2048 bytes in)

    const int l = blockIdx.x, img = blockIdx.y;
    const int partA = c_joint_pairs[2 * l], partB = c_joint_pairs[2 * l + 1];
    const float* pA = peaks + ((size_t)img * NJ + partA) * (MAXP + 1) * 3;
    const float* pB = peaks + ((size_t)img * NJ + partB) * (MAXP + 1) * 3;
    float* out = scores + ((size_t)img * NL + l) * MAXP * MAXP;

    pdl_wait();
    const int nA = (int)pA[0], nB = (int)pB[0];
    if (nA > 0 && nB > 0) {
        const float* src = hms + ((size_t)img * nchan + NJ + 2 * l) * hw;
        if (STAGED) stage_planes(planes, src, (uint32_t)hw * 8u, bar);
        const float* mapX = STAGED ? planes : src;
        for (int i = threadIdx.x; i < nA; i += PAF_THREADS) {
            pk[2 * i] = pA[3 * (i + 1)];
            pk[2 * i + 1] = pA[3 * (i + 1) + 1];
        }
        for (int i = threadIdx.x; i < nB; i += PAF_THREADS) {
            pk[2 * (MAXP + 1) + 2 * i] = pB[3 * (i + 1)];
            pk[2 * (MAXP + 1) + 2 * i + 1] = pB[3 * (i + 1) + 1];
        }
        __syncthreads();
        const float near_thr = __fdiv_rn(__fsqrt_rn((float)(w * h)), 150.f);
        const int npairs = nA * nB;
        for (int p = threadIdx.x; p < npairs; p += PAF_THREADS) {
            const int a = p / nB, b = p - a * nB;
            out[a * MAXP + b] = paf_process(pk[2 * a], pk[2 * a + 1], pk[2 * (MAXP + 1) + 2 * b],
                                            pk[2 * (MAXP + 1) + 2 * b + 1], mapX, mapX + hw, w, h, near_thr);
        }
    }
    if (dense_fill) {  // pafScoreKernel writes -1 outside nA x nB; only the extract() API needs it
        for (int p = threadIdx.x; p < MAXP * MAXP; p += PAF_THREADS) {
            const int a = p / MAXP, b = p - a * MAXP;
            if (a >= nA || b >= nB) out[p] = -1.f;
        }
    }
    pdl_trigger();
}

// ---------------------------------------------------------------------------------------------
// `predRootDepth.sort(0, false)` (association.cpp:144): at::sort(stable=false) on a CPU tensor is libstdc++
// std::sort over (key, index) pairs with comp(a,b) = (!isnan(a) && isnan(b)) || a < b.  It is not stable, so
// equal depths come out in introsort order.  When all depths are distinct the order is unique and a parallel
// rank sort is used; otherwise one thread replays libstdc++'s algorithm (bits/stl_algo.h: introsort loop with
// median-of-3 to first + unguarded partition, threshold 16, heap-sort fallback at depth 2*floor(log2 n), final
// insertion sort) step by step so that the tie order is bit-identical to the reference.
// ---------------------------------------------------------------------------------------------
struct KV {
    float k;
    int v;
};
__device__ __forceinline__ bool kv_comp(const KV& a, const KV& b) { return (!isnan(a.k) && isnan(b.k)) || (a.k < b.k); }
__device__ __forceinline__ void kv_swap(KV& a, KV& b) {
    const KV t = a;
    a = b;
    b = t;
}
__device__ void kv_adjust_heap(KV* first, int hole, int len, KV value) {  // std::__adjust_heap + __push_heap
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (kv_comp(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && kv_comp(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
__device__ void kv_heap_sort(KV* first, int len) {  // std::__partial_sort(first, last, last)
    if (len >= 2) {                                  // __make_heap
        int parent = (len - 2) / 2;
        while (true) {
            const KV value = first[parent];
            kv_adjust_heap(first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    // __heap_select(first, last, last) has an empty tail; then __sort_heap
    for (int last = len; last > 1;) {
        --last;
        const KV value = first[last];  // __pop_heap(first, last, last)
        first[last] = first[0];
        kv_adjust_heap(first, 0, last, value);
    }
}
__device__ __forceinline__ void kv_unguarded_linear_insert(KV* a, int last) {
    const KV val = a[last];
    int next = last - 1;
    while (kv_comp(val, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}
__device__ void kv_insertion_sort(KV* a, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (kv_comp(a[i], a[first])) {
            const KV val = a[i];
            for (int j = i; j > first; --j) a[j] = a[j - 1];
            a[first] = val;
        } else {
            kv_unguarded_linear_insert(a, i);
        }
    }
}
__device__ void kv_std_sort(KV* a, int n) {
    if (n <= 0) return;
    int lg = 0;
    while ((1 << (lg + 1)) <= n) lg++;
    // explicit stack for the recursive half of __introsort_loop
    int st_first[32], st_last[32], st_depth[32];
    int sp = 0;
    st_first[0] = 0, st_last[0] = n, st_depth[0] = 2 * lg;
    sp = 1;
    while (sp > 0) {
        --sp;
        int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {
                kv_heap_sort(a + first, last - first);
                break;
            }
            --depth;
            // __unguarded_partition_pivot
            const int mid = first + (last - first) / 2;
            {
                const int ia = first + 1, ib = mid, ic = last - 1;
                if (kv_comp(a[ia], a[ib])) {
                    if (kv_comp(a[ib], a[ic])) kv_swap(a[first], a[ib]);
                    else if (kv_comp(a[ia], a[ic])) kv_swap(a[first], a[ic]);
                    else kv_swap(a[first], a[ia]);
                } else if (kv_comp(a[ia], a[ic])) kv_swap(a[first], a[ia]);
                else if (kv_comp(a[ib], a[ic])) kv_swap(a[first], a[ic]);
                else kv_swap(a[first], a[ib]);
            }
            int lo = first + 1, hi = last;
            while (true) {
                while (kv_comp(a[lo], a[first])) ++lo;
                --hi;
                while (kv_comp(a[first], a[hi])) --hi;
                if (!(lo < hi)) break;
                kv_swap(a[lo], a[hi]);
                ++lo;
            }
            const int cut = lo;
            // recurse on [cut, last), iterate on [first, cut): libstdc++ runs the right part first, but the two
            // ranges are disjoint, so the order of processing does not change the result
            st_first[sp] = cut, st_last[sp] = last, st_depth[sp] = depth;
            ++sp;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        kv_insertion_sort(a, 0, 16);
        for (int i = 16; i != n; ++i) kv_unguarded_linear_insert(a, i);
    } else {
        kv_insertion_sort(a, 0, n);
    }
}

// ---------------------------------------------------------------------------------------------
// Grouping: one CTA (5 warps) per image.  The 14 limbs form 5 independent chains hanging off the
// pelvis/neck (dst joints are disjoint, `used` is per limb), so the sequential reference order
//   1,0,2,3,...,13  (association.cpp:164-170)
// is reproduced exactly by running the chains concurrently in 4 barrier-separated phases:
//   phase0: 1 | 8 | 11      phase1: 0 | 2 | 5 | 9 | 12      phase2: 3 | 6 | 10 | 13      phase3: 4 | 7
// Within a limb, persons are visited serially in ascending root depth (the ordinal prior); the scan over
// destination candidates is a warp arg-max with "first index wins" (strict > in ascending k2).
// ---------------------------------------------------------------------------------------------
constexpr int GROUP_WARPS = 5;
__constant__ int c_phase_limb[4][GROUP_WARPS] = {
    {1, -1, -1, 8, 11}, {0, 2, 5, 9, 12}, {-1, 3, 6, 10, 13}, {-1, 4, 7, -1, -1}};
// root_idx != 2 (neck root): the leg chains depend on limb 1, so fall back to the reference's serial order.
__constant__ int c_serial_limb[NL] = {1, 0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13};

__global__ void __launch_bounds__(GROUP_WARPS * 32, 1)
group_kernel(const float* __restrict__ peaks, const float* __restrict__ scores, const float* __restrict__ rdepth,
             int h, int w, int root_idx, int dist_flag, float* __restrict__ bodies, int* __restrict__ counts) {
    __shared__ float s_depth[MAXP + 1];
    __shared__ float s_sorted[MAXP + 1];
    __shared__ int s_order[MAXP + 1];
    __shared__ int s_rank[MAXP + 1];
    __shared__ unsigned long long s_kv[MAXP + 1];
    __shared__ unsigned char s_remap[NJ][MAXP + 1];
    __shared__ float s_body[MAXP][NJ][3];  // x, y, score
    __shared__ unsigned char s_used[GROUP_WARPS][MAXP + 1];

    const int img = blockIdx.x;
    const float* pk = peaks + (size_t)img * NJ * (MAXP + 1) * 3;
    const float* sc_img = scores + (size_t)img * NL * MAXP * MAXP;
    const float* rd = rdepth + (size_t)img * h * w;
    float* outb = bodies + (size_t)img * MAXP * NJ * 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthr = GROUP_WARPS * 32;

    pdl_wait();
    const float* rootPeaks = pk + (size_t)root_idx * (MAXP + 1) * 3;
    const int P = (int)rootPeaks[0];
    if (tid == 0) counts[img] = P;

    for (int i = tid; i < MAXP * NJ * 3; i += nthr) (&s_body[0][0][0])[i] = 0.f;
    for (int i = tid; i < P; i += nthr) {  // association.cpp:139-142
        const int yy = (int)rootPeaks[3 * (i + 1) + 1], xx = (int)rootPeaks[3 * (i + 1)];
        s_depth[i] = rd[min(h - 1, max(0, yy)) * w + min(w - 1, max(0, xx))];
    }
    __syncthreads();
    // ascending depth order (association.cpp:144): unique keys -> parallel rank sort, ties/NaNs -> std::sort replay
    {
        int tie = 0;
        for (int i = tid; i < P; i += nthr) {
            const KV a = {s_depth[i], i};
            int r = 0;
            for (int j = 0; j < P; j++) {
                const KV b = {s_depth[j], j};
                const bool lt = kv_comp(b, a);
                r += lt;
                tie |= (j != i) && !lt && !kv_comp(a, b);
            }
            s_rank[i] = r;
        }
        const int any_tie = __syncthreads_or(tie);
        if (!any_tie) {
            for (int i = tid; i < P; i += nthr) {
                s_order[s_rank[i]] = i;
                s_sorted[s_rank[i]] = s_depth[i];
            }
        } else if (tid == 0) {
            KV* kv = reinterpret_cast<KV*>(s_kv);
            for (int i = 0; i < P; i++) {
                kv[i].k = s_depth[i];
                kv[i].v = i;
            }
            kv_std_sort(kv, P);
            for (int i = 0; i < P; i++) {
                s_order[i] = kv[i].v;
                s_sorted[i] = kv[i].k;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < NJ * P; i += nthr) {  // association.cpp:148-154
        const int j = i / P, p = i - j * P;
        s_remap[j][p] = (unsigned char)((j == root_idx) ? s_order[p] : p);
    }
    for (int p = tid; p < P; p += nthr) {  // association.cpp:156-162
        const float* r = rootPeaks + 3 * (s_order[p] + 1);
        s_body[p][root_idx][0] = r[0];
        s_body[p][root_idx][1] = r[1];
        s_body[p][root_idx][2] = r[2];
    }
    __syncthreads();

    if (P > 0) {
        const int nphase = (root_idx == 2) ? 4 : NL;
        for (int phase = 0; phase < nphase; phase++) {
            const int i = (root_idx == 2) ? c_phase_limb[phase][warp] : (warp == 0 ? c_serial_limb[phase] : -1);
            if (i >= 0) {
                int src, dst;
                bool flip = false;
                if (root_idx == 2 && i == 1) {  // association.cpp:171-174
                    src = c_joint_pairs[2 * i + 1];
                    dst = c_joint_pairs[2 * i];
                    flip = true;
                } else {
                    src = c_joint_pairs[2 * i];
                    dst = c_joint_pairs[2 * i + 1];
                }
                const float* dstPeaks = pk + (size_t)dst * (MAXP + 1) * 3;
                const int dstSize = (int)dstPeaks[0];
                if (dstSize > 0) {
                    const float* sc = sc_img + (size_t)i * MAXP * MAXP;
                    unsigned char* used = s_used[warp];
                    for (int k = lane; k < dstSize; k += 32) used[k] = 0;
                    // candidate coordinates in registers: lane holds k2 = lane + 32*q
                    float cx[4], cy[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int k2 = lane + 32 * q;
                        cx[q] = (k2 < dstSize) ? dstPeaks[3 * (k2 + 1)] : 0.f;
                        cy[q] = (k2 < dstSize) ? dstPeaks[3 * (k2 + 1) + 1] : 0.f;
                    }
                    __syncwarp();
                    const float bl = c_bone_length[i];
                    // Score rows are prefetched one person ahead: the row index (remap of the source joint) and the
                    // source scores are fixed for the whole limb, only `used` changes from person to person.
                    auto load_row = [&](int k1, float(&row)[4]) {
                        const int rs = s_remap[src][k1];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int k2 = lane + 32 * q;
                            row[q] = (k2 < dstSize) ? (flip ? sc[k2 * MAXP + rs] : sc[rs * MAXP + k2]) : -1.f;
                        }
                    };
                    auto next_valid = [&](int k1) {
                        while (k1 < P && (double)s_body[k1][src][2] < 1e-5) k1++;  // association.cpp:190
                        return k1;
                    };
                    int k1 = next_valid(0);
                    float cur[4], nxt[4];
                    if (k1 < P) load_row(k1, cur);
                    while (k1 < P) {
                        const int k1n = next_valid(k1 + 1);
                        if (k1n < P) load_row(k1n, nxt);
                        const float sx = s_body[k1][src][0], sy = s_body[k1][src][1];
                        const float bone_dist =
                            __double2float_rn(__ddiv_rn(__dmul_rn(1.2, (double)bl), (double)s_sorted[k1]));
                        float best = 0.0f;
                        int bestIdx = 0x7fffffff;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int k2 = lane + 32 * q;
                            if (k2 < dstSize && !used[k2]) {
                                float score = cur[q];
                                if (dist_flag && score > 0) {
                                    const float ddx = __fsub_rn(sx, cx[q]), ddy = __fsub_rn(sy, cy[q]);
                                    const double d2 = __dadd_rn(__dmul_rn((double)ddx, (double)ddx),
                                                                __dmul_rn((double)ddy, (double)ddy));
                                    const float limb_dist = __double2float_rn(__dsqrt_rn(d2));
                                    const float t =
                                        __fsub_rn(__fdiv_rn(__fdiv_rn(bone_dist, limb_dist), 4.0f), 1.0f);
                                    score = __fadd_rn(score, (0.0f < t) ? 0.0f : t);  // std::min(t, 0.0f)
                                }
                                if (score > best) {  // ascending k2 within the lane: strict > keeps the first
                                    best = score;
                                    bestIdx = k2;
                                }
                            }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                            const int oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
                            if (ob > best || (ob == best && oi < bestIdx)) {
                                best = ob;
                                bestIdx = oi;
                            }
                        }
                        if (best > 0) {  // association.cpp:220-228
                            if (lane == 0) {
                                s_body[k1][dst][0] = dstPeaks[3 * (bestIdx + 1)];
                                s_body[k1][dst][1] = dstPeaks[3 * (bestIdx + 1) + 1];
                                s_body[k1][dst][2] = dstPeaks[3 * (bestIdx + 1) + 2];
                                s_remap[dst][k1] = (unsigned char)bestIdx;
                                used[bestIdx] = 1;
                            }
                            __syncwarp();
                        }
                        k1 = k1n;
#pragma unroll
                        for (int q = 0; q < 4; q++) cur[q] = nxt[q];
                    }
                }
            }
            __syncthreads();
        }
    }
    // bodies [MAXP][NJ][4] = (x, y, 0, score); rows >= P zeroed
    for (int i = tid; i < MAXP * NJ; i += nthr) {
        const int p = i / NJ, j = i - p * NJ;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < P) v = make_float4(s_body[p][j][0], s_body[p][j][1], 0.f, s_body[p][j][2]);
        reinterpret_cast<float4*>(outb)[i] = v;
    }
    pdl_trigger();
}

// ---------------------------------------------------------------------------------------------
// 3D lift (rows C1-C4).  One CTA per image; a thread per (person, limb) for the relative-depth line
// statistics, then a thread per person for the bone chain + back-projection.  float32/float64 islands
// follow what the reference's numpy code does (see oracle/lift_numpy.py).
// ---------------------------------------------------------------------------------------------
constexpr int LIFT_THREADS = 256;

__device__ __forceinline__ float np_linspace10(float start, float stop, int i) {
    // numpy.linspace(start, stop, 10) in float32: y = i*step + start (two roundings), last = stop
    if (i == 9) return stop;
    const float delta = __fsub_rn(stop, start);
    const float step = __fdiv_rn(delta, 9.f);
    float y;
    if (step == 0.f)
        y = __fmul_rn(__fdiv_rn((float)i, 9.f), delta);
    else
        y = __fmul_rn((float)i, step);
    return __fadd_rn(y, start);
}

__global__ void __launch_bounds__(LIFT_THREADS, 1)
lift_kernel(const float* __restrict__ bodies, const int* __restrict__ counts, const float* __restrict__ det_d,
            const float* __restrict__ root_d, const double* __restrict__ scales, int h, int w, int root_n,
            float* __restrict__ pred2d_base, double* __restrict__ pred3d_base, double* __restrict__ root_depth_base,
            int* __restrict__ counts_out, long long s2d, long long s3d, long long srd, long long scnt) {
    __shared__ float s_b[MAXP][NJ][4];
    __shared__ double s_dz[MAXP][NL];
    __shared__ int s_keep[MAXP];
    __shared__ int s_np;
    const int img = blockIdx.x, tid = threadIdx.x;
    const int hw = h * w;
    // per-image output slices (strides in elements: natural layout or smapb_record fields)
    float* pred2d = pred2d_base + (size_t)img * s2d;
    double* pred3d = pred3d_base + (size_t)img * s3d;
    double* root_depth = root_depth_base + (size_t)img * srd;
    pdl_wait();
    const int P = counts[img];
    const float* b = bodies + (size_t)img * MAXP * NJ * 4;
    const float* dd = det_d + (size_t)img * NL * hw;
    const float* rd = root_d + (size_t)img * hw;
    const double* sc = scales + (size_t)img * 9;  // scale, img_w, img_h, net_w, net_h, fx, fy, cx, cy
    if (tid == 0) {  // register_pred without GT (test_util.py:41): keep persons whose root score != 0
        int n = 0;
        for (int p = 0; p < P; p++)
            if (b[(p * NJ + root_n) * 4 + 3] != 0.f) s_keep[n++] = p;
        s_np = n;
    }
    __syncthreads();
    const int NP = s_np;
    for (int i = tid; i < NP * NJ; i += LIFT_THREADS) {
        const int p = i / NJ, j = i - p * NJ;
        const float* s = b + (s_keep[p] * NJ + j) * 4;
        s_b[p][j][0] = __fmul_rn(s[0], 4.f);  // test.py:117
        s_b[p][j][1] = __fmul_rn(s[1], 4.f);
        s_b[p][j][2] = s[2];
        s_b[p][j][3] = s[3];
    }
    __syncthreads();
    // numpy percentile constants (method 'linear'): virtual index (n-1)*q, gamma = frac
    const double vi10 = 9.0 * (10.0 / 100.0), vi90 = 9.0 * (90.0 / 100.0);
    const double g10 = vi10 - floor(vi10), g90 = vi90 - floor(vi90);
    const int i10 = (int)floor(vi10), i90 = (int)floor(vi90);
    for (int i = tid; i < NP * NL; i += LIFT_THREADS) {
        const int p = i / NL, k = i - p * NL;
        const int ja = c_joint_pairs[2 * k], jb = c_joint_pairs[2 * k + 1];
        double dz = 0.0;
        if (s_b[p][root_n][3] > 0 && s_b[p][jb][3] > 0 && s_b[p][ja][3] > 0) {
            float v[10];
#pragma unroll
            for (int t = 0; t < 10; t++) {
                const int xx = (int)rintf(np_linspace10(s_b[p][ja][0], s_b[p][jb][0], t));
                const int yy = (int)rintf(np_linspace10(s_b[p][ja][1], s_b[p][jb][1], t));
                const int hx = min(w - 1, max(0, xx >> 2)), hy = min(h - 1, max(0, yy >> 2));
                v[t] = dd[(size_t)k * hw + hy * w + hx];
            }
            float s[10];
#pragma unroll
            for (int t = 0; t < 10; t++) s[t] = v[t];
#pragma unroll
            for (int a = 1; a < 10; a++) {  // insertion sort
                const float key = s[a];
                int q = a - 1;
                while (q >= 0 && s[q] > key) {
                    s[q + 1] = s[q];
                    q--;
                }
                s[q + 1] = key;
            }
            // numpy _lerp: a + (b-a)*t, or b - (b-a)*(1-t) where t >= 0.5; (b-a) in float32
            auto lerp = [](float a, float bb, double t) {
                const double diff = (double)__fsub_rn(bb, a);
                return (t >= 0.5) ? __dsub_rn((double)bb, __dmul_rn(diff, __dsub_rn(1.0, t)))
                                  : __dadd_rn((double)a, __dmul_rn(diff, t));
            };
            const double lo = lerp(s[i10], s[i10 + 1], g10);
            const double hi = lerp(s[i90], s[i90 + 1], g90);
            const float flo = __double2float_rn(lo), fhi = __double2float_rn(hi);
#pragma unroll
            for (int t = 0; t < 10; t++) {
                if ((double)v[t] < lo) v[t] = flo;
                if ((double)v[t] > hi) v[t] = fhi;
            }
            // numpy float32 pairwise sum for n = 10: 8-way unrolled block + 2 tail adds
            const float r = __fadd_rn(__fadd_rn(__fadd_rn(v[0], v[1]), __fadd_rn(v[2], v[3])),
                                      __fadd_rn(__fadd_rn(v[4], v[5]), __fadd_rn(v[6], v[7])));
            const float tot = __fadd_rn(__fadd_rn(r, v[8]), v[9]);
            dz = (double)__fdiv_rn(tot, 10.f);
        }
        s_dz[p][k] = dz;
    }
    __syncthreads();
    for (int p = tid; p < NP; p += LIFT_THREADS) {
        double rdep = 0.0;
        if (s_b[p][root_n][3] > 0) {
            const int ry = (int)s_b[p][root_n][1], rx = (int)s_b[p][root_n][0];
            const float r = rd[min(h - 1, max(0, ry >> 2)) * w + min(w - 1, max(0, rx >> 2))];
            rdep = __dmul_rn(__dmul_rn((double)r, sc[0]), sc[5]);  // test_util.py:66
            // chain_bones (test_util.py:45-57): float32 column written in place
            s_b[p][2][2] = 0.f;
            s_b[p][0][2] = __double2float_rn(__dsub_rn((double)s_b[p][2][2], s_dz[p][1]));
            s_b[p][1][2] = __double2float_rn(__dadd_rn((double)s_b[p][0][2], s_dz[p][0]));
            for (int k = 2; k < NL; k++) {
                const int ja = c_joint_pairs[2 * k], jb = c_joint_pairs[2 * k + 1];
                s_b[p][jb][2] = __double2float_rn(__dadd_rn((double)s_b[p][ja][2], s_dz[p][k]));
            }
        }
        root_depth[p] = rdep;
        // gen_3d_pose (test_util.py:89-99) + get_3d_points/back_projection (post_3d.py:4-27)
        const double s = sc[0];
        const double offx = __ddiv_rn(__dsub_rn(__ddiv_rn(sc[3], s), sc[1]), 2.0);
        const double offy = __ddiv_rn(__dsub_rn(__ddiv_rn(sc[4], s), sc[2]), 2.0);
        const bool has_root = s_b[p][root_n][3] != 0.f;
        for (int j = 0; j < NJ; j++) {
            float* o2 = pred2d + ((size_t)p * NJ + j) * 4;
            o2[0] = s_b[p][j][0];
            o2[1] = s_b[p][j][1];
            o2[2] = s_b[p][j][2];
            o2[3] = s_b[p][j][3];
            double* o3 = pred3d + ((size_t)p * NJ + j) * 4;
            double X = 0, Y = 0, Z = 0;
            const float score = s_b[p][j][3];
            if (has_root && score != 0.f) {
                const float bx = __double2float_rn(__dsub_rn(__ddiv_rn((double)s_b[p][j][0], s), offx));
                const float by = __double2float_rn(__dsub_rn(__ddiv_rn((double)s_b[p][j][1], s), offy));
                const float bz = __double2float_rn(__dadd_rn((double)s_b[p][j][2], rdep));
                const double d = (double)bz;
                X = __ddiv_rn(__dmul_rn(__dsub_rn((double)bx, sc[7]), d), sc[5]);
                Y = __ddiv_rn(__dmul_rn(__dsub_rn((double)by, sc[8]), d), sc[6]);
                Z = d;
            }
            o3[0] = X;
            o3[1] = Y;
            o3[2] = Z;
            o3[3] = (double)score;
        }
    }
    // zero the unused tail of the fixed-stride record (all-gather payload must be deterministic)
    for (int i = NP * NJ * 4 + tid; i < MAXP * NJ * 4; i += LIFT_THREADS) {
        pred2d[i] = 0.f;
        pred3d[i] = 0.0;
    }
    for (int i = NP + tid; i < MAXP; i += LIFT_THREADS) root_depth[i] = 0.0;
    if (tid == 0) {
        counts_out[(size_t)img * scnt] = NP;
        if (scnt > 1) counts_out[(size_t)img * scnt + 1] = 0;  // smapb_record::pad_
    }
    pdl_trigger();
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static int nms_bands(int h) {
    const int b = (h + 31) / 32;
    return b < 1 ? 1 : (b > NMS_MAX_BANDS ? NMS_MAX_BANDS : b);
}
static int nms_rows(int h) { return (h + nms_bands(h) - 1) / nms_bands(h); }
static size_t nms_smem(int h, int w) {
    const int rows = nms_rows(h);
    return 64 + (size_t)(rows + 6) * w * 4 + (size_t)((rows * w + 31) / 32) * 4;
}
static size_t paf_pk_bytes() { return 16 + 4 * (MAXP + 1) * 4; }
static size_t paf_smem(int h, int w) { return paf_pk_bytes() + (size_t)h * w * 8; }
static bool paf_staged(int h, int w) { return paf_smem(h, w) <= 232448; }

// Any map size with w % 4 == 0 (16-byte bulk copies) whose NMS row band fits in shared memory; the reference hard-codes
// 128 x 208 (extensions/association.cpp:21).
int assoc_configure(int h, int w, const char** err) {
    static const char* e_big = "association: heat-map too large (an NMS row band of h/8 + 6 rows must fit in 227 KB of shared memory)";
    static const char* e_align = "association: w must be a multiple of 4 (16-byte bulk copies)";
    if (w % 4 != 0 || h < 3 || w < 4) {
        *err = e_align;
        return -1;
    }
    if (nms_smem(h, w) > 232448) {
        *err = e_big;
        return -1;
    }
    cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem(h, w));
    if (e == cudaSuccess && paf_staged(h, w))
        e = cudaFuncSetAttribute(paf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)paf_smem(h, w));
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -2;
    }
    return 0;
}

cudaError_t launch_nms(const float* hms, int nchan, int B, int h, int w, float thr, float* peaks, cudaStream_t st) {
    const int bands = nms_bands(h);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(bands, NJ, B);
    cfg.blockDim = dim3(NMS_THREADS);
    cfg.dynamicSmemBytes = nms_smem(h, w);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = bands;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, nms_kernel, hms, nchan, h, w, nms_rows(h), thr, peaks);
}
cudaError_t launch_paf(const float* hms, int nchan, int B, int h, int w, const float* peaks, float* scores,
                       int dense_fill, cudaStream_t st) {
    if (paf_staged(h, w))
        paf_kernel<true><<<dim3(NL, B), PAF_THREADS, paf_smem(h, w), st>>>(hms, nchan, h, w, peaks, scores, dense_fill);
    else
        paf_kernel<false><<<dim3(NL, B), PAF_THREADS, paf_pk_bytes(), st>>>(hms, nchan, h, w, peaks, scores, dense_fill);
    return cudaGetLastError();
}
cudaError_t launch_group(const float* peaks, const float* scores, const float* rdepth, int B, int h, int w,
                         int root_idx, int dist_flag, float* bodies, int* counts, cudaStream_t st) {
    group_kernel<<<B, GROUP_WARPS * 32, 0, st>>>(peaks, scores, rdepth, h, w, root_idx, dist_flag, bodies, counts);
    return cudaGetLastError();
}
cudaError_t launch_lift(const float* bodies, const int* counts, const float* det_d, const float* root_d,
                        const double* scales, int B, int h, int w, int root_n, float* pred2d, double* pred3d,
                        double* root_depth, int* counts_out, long long s2d, long long s3d, long long srd, long long scnt,
                        cudaStream_t st) {
    lift_kernel<<<B, LIFT_THREADS, 0, st>>>(bodies, counts, det_d, root_d, scales, h, w, root_n, pred2d, pred3d,
                                            root_depth, counts_out, s2d, s3d, srd, scnt);
    return cudaGetLastError();
}

}  // namespace smapb
