// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   y[n, oy, ox, co] = epilogue( sum_{ky,kx,ci} x[n, oy*s+ky-p, ox*s+kx-p, ci] * w[co, ci, ky, kx] )
//
// GEMM view: M = output pixels (a th x tw spatial patch = 128 rows per tile), N = Cout, K = taps * Cin.
// One k-block = one filter tap x 64 input channels.  Operands are K-major bf16 tiles of 128-byte rows in
// SWIZZLE_128B shared-memory layout, written by TMA:
//   A: 5-D tensor map (C, W, H, N, term) over the NHWC activation planes; the tap shift is a coordinate
//      offset, zero padding is TMA out-of-bounds fill, conv stride is the map's elementStrides.
//   B: 4-D tensor map (Cin, Cout, tap, term) over the repacked (BN-folded) weights.
// Precision: "bf16x3" - every fp32 value v is carried as two bf16 planes (hi = bf16(v), lo = bf16(v - hi));
// a product a*b is issued as 3 MMAs  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  into the same fp32 TMEM accumulator
// (dropped terms are O(2^-16) relative) - fp32-faithful results at 3x the bf16 MMA count.  NTERMS = 1 is the
// plain bf16 fast mode.
//
// Warp roles (256 threads, persistent CTA, static tile schedule):
//   warp 0 : TMA producer (one elected lane)        warp 1 : MMA issuer (one elected lane)
//   warp 2 : TMEM allocator                         warps 4-7 : epilogue (TMEM -> regs -> global)
// Pipelines: smem full/empty ring (TMA <-> MMA), double-buffered TMEM accumulators (MMA <-> epilogue).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace smapb {

struct alignas(64) ConvParams {
    CUtensorMap tmA;
    CUtensorMap tmB;
    // geometry
    int Hout, Wout, Nimg;
    int tw_log2, th;  // tile: tw = 1 << tw_log2, tw * th == 128
    int tiles_x, tiles_y;
    int Cout;  // channel stride of the output / residual tensors (elements)
    int ksize, stride, pad;
    int kchunks;  // Cin / 64
    int n_tiles;  // Cout_pad / BLOCK_N
    int total_tiles;
    // epilogue
    const float* bias;           // [Cout_pad] folded bias
    const __nv_bfloat16* res;    // residual added before ReLU (hi plane; lo plane at + plane_stride) or null
    const __nv_bfloat16* post1;  // added after ReLU (x_k = layer_k + skip1 + skip2, model/smap.py:143) or null
    const __nv_bfloat16* post2;
    __nv_bfloat16* out;  // bf16 hi/lo planes NHWC, or null
    float* out_f32;      // fp32 NHWC (heads), or null
    long long plane_stride;  // elements between the hi and lo planes (= N*H*W*Cout)
    int relu;
};

// ---- tcgen05 / TMA PTX wrappers -----------------------------------------------------------------
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];" ::"r"(dst),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];" ::"r"(dst),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand descriptor: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <int BLOCK_N, int NTERMS>
struct ConvCfg {
    static constexpr int TA = (NTERMS == 3) ? 2 : 1;  // operand planes held per stage
    static constexpr int A_BYTES = 128 * 128;         // 128 rows x 64 bf16
    static constexpr int B_BYTES = BLOCK_N * 128;
    static constexpr int STAGE_BYTES = TA * (A_BYTES + B_BYTES);
    static constexpr int SMEM_BUDGET = 227 * 1024 - 2048;
    static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2048;  // 1 KB barriers + 1 KB alignment slack
    static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                     : (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(STAGES >= 2, "need at least a double buffer");
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");
};

template <int BLOCK_N, int NTERMS>
__global__ void __launch_bounds__(256, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<BLOCK_N, NTERMS>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ unsigned char smem_raw[];
    // control block at the front, operand ring 1024-aligned behind it
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const uint32_t ring = (smem_u32(smem_raw) + 1024u + 1023u) & ~1023u;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 128);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_kb = p.ksize * p.ksize * p.kchunks;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int tw = 1 << p.tw_log2;

    pdl_wait();  // inputs of this layer are produced by the previous kernel in the stream

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
                const int img = mt / tiles_per_img, r = mt - img * tiles_per_img;
                const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                const int x_in0 = (tx << p.tw_log2) * p.stride - p.pad;
                const int y_in0 = ty * p.th * p.stride - p.pad;
                for (int kb = 0; kb < num_kb; kb++) {
                    const int tap = kb / p.kchunks, kc = kb - tap * p.kchunks;
                    const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    const uint32_t sbase = ring + stage * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
                    for (int t = 0; t < Cfg::TA; t++) {
                        tma_load_5d(sbase + t * Cfg::A_BYTES, &p.tmA, &full_bar[stage], kc * 64, x_in0 + kx, y_in0 + ky,
                                    img, t);
                        tma_load_4d(sbase + Cfg::TA * Cfg::A_BYTES + t * Cfg::B_BYTES, &p.tmB, &full_bar[stage],
                                    kc * 64, nt * BLOCK_N, tap, t);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ==============================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N, M=128
            constexpr uint32_t idesc =
                (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < num_kb; kb++) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sbase = ring + stage * Cfg::STAGE_BYTES;
                    const uint64_t a0 = umma_desc_sw128(sbase);
                    const uint64_t b0 = umma_desc_sw128(sbase + Cfg::TA * Cfg::A_BYTES);
                    constexpr uint64_t A_STEP = (uint64_t)(Cfg::A_BYTES >> 4), B_STEP = (uint64_t)(Cfg::B_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < 4; k++) {  // 4 x UMMA_K(16) per 64-channel k-block; +32 B per step
                        const uint64_t ka = a0 + (uint64_t)(k * 2), kbd = b0 + (uint64_t)(k * 2);
                        tc_mma_bf16(tmem_d, ka, kbd, idesc, (kb | k) != 0);  // a_hi * b_hi
                        if (NTERMS == 3) {
                            tc_mma_bf16(tmem_d, ka + A_STEP, kbd, idesc, 1u);  // a_lo * b_hi
                            tc_mma_bf16(tmem_d, ka, kbd + B_STEP, idesc, 1u);  // a_hi * b_lo
                        }
                    }
                    tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        // ============================ epilogue ================================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
            const int img = mt / tiles_per_img, r = mt - img * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
            const int py = ty * p.th + (row >> p.tw_log2);
            const int px = (tx << p.tw_log2) + (row & (tw - 1));
            const bool valid = (py < p.Hout) && (px < p.Wout);
            const long long pix = ((long long)img * p.Hout + py) * p.Wout + px;
            const int n0 = nt * BLOCK_N;

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(acc * BLOCK_N) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t acc_r[32];
                tmem_ld32(taddr + (uint32_t)c0, acc_r);
                // residual operands for this chunk are fetched while the TMEM load is in flight
                const long long off = pix * p.Cout + n0 + c0;
                uint4 rh[4], rl[4];
                const bool has_res = valid && (p.res != nullptr);
                if (has_res) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        rh[j] = ldg_nc_v4(p.res + off + j * 8);
                        rl[j] = (NTERMS == 3) ? ldg_nc_v4(p.res + p.plane_stride + off + j * 8) : make_uint4(0, 0, 0, 0);
                    }
                }
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0) + j);
                    v[4 * j + 0] = __uint_as_float(acc_r[4 * j + 0]) + b.x;
                    v[4 * j + 1] = __uint_as_float(acc_r[4 * j + 1]) + b.y;
                    v[4 * j + 2] = __uint_as_float(acc_r[4 * j + 2]) + b.z;
                    v[4 * j + 3] = __uint_as_float(acc_r[4 * j + 3]) + b.w;
                }
                if (has_res) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t h[4] = {rh[j].x, rh[j].y, rh[j].z, rh[j].w};
                        const uint32_t l[4] = {rl[j].x, rl[j].y, rl[j].z, rl[j].w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            v[8 * j + 2 * e] += bf16lo_to_f(h[e]) + bf16lo_to_f(l[e]);
                            v[8 * j + 2 * e + 1] += bf16hi_to_f(h[e]) + bf16hi_to_f(l[e]);
                        }
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
                }
                if (valid && p.post1 != nullptr) {
#pragma unroll
                    for (int s = 0; s < 2; s++) {
                        const __nv_bfloat16* src = s == 0 ? p.post1 : p.post2;
                        if (src == nullptr) continue;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint4 hh = ldg_nc_v4(src + off + j * 8);
                            const uint4 ll =
                                (NTERMS == 3) ? ldg_nc_v4(src + p.plane_stride + off + j * 8) : make_uint4(0, 0, 0, 0);
                            const uint32_t h[4] = {hh.x, hh.y, hh.z, hh.w};
                            const uint32_t l[4] = {ll.x, ll.y, ll.z, ll.w};
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                v[8 * j + 2 * e] += bf16lo_to_f(h[e]) + bf16lo_to_f(l[e]);
                                v[8 * j + 2 * e + 1] += bf16hi_to_f(h[e]) + bf16hi_to_f(l[e]);
                            }
                        }
                    }
                }
                if (valid) {
                    if (p.out != nullptr) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            uint32_t hw_[4], lw_[4];
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                __nv_bfloat16 h0, l0, h1, l1;
                                split_bf16(v[8 * j + 2 * e], h0, l0);
                                split_bf16(v[8 * j + 2 * e + 1], h1, l1);
                                hw_[e] = pack_bf16x2(h0, h1);
                                lw_[e] = pack_bf16x2(l0, l1);
                            }
                            *reinterpret_cast<uint4*>(p.out + off + j * 8) = make_uint4(hw_[0], hw_[1], hw_[2], hw_[3]);
                            if (NTERMS == 3)
                                *reinterpret_cast<uint4*>(p.out + p.plane_stride + off + j * 8) =
                                    make_uint4(lw_[0], lw_[1], lw_[2], lw_[3]);
                        }
                    }
                    if (p.out_f32 != nullptr) {
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            *reinterpret_cast<float4*>(p.out_f32 + off + j * 4) =
                                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }

    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS)
                     : "memory");
    }
}

}  // namespace smapb
