// Compile-time specialised LDE passes (same algorithm and tile geometry as ntt.cuh; see the header there).
// The generic kernels in ntt.cuh spend more than half of their instructions on run-time shifts/masks for tile geometry and
// round structure (ncu r01: 20-30 instructions per butterfly, ALU pipe 55 % busy with address math).  Here the row-bit
// count RB, the low split NLO, the round plan (Q, B0) and the lane count are template parameters, so shared/global
// addressing folds into immediates and each kernel carries only its own two or three register rounds.
#pragma once
#include "bb31.cuh"
#include "ntt.cuh"

namespace nttf {

using ntt::brev;
// 256-thread CTAs on 64 KB tiles: two or three CTAs are co-resident per SM, so the load phase of one overlaps the butterfly
// phase of another (ncu r01b: one 512-thread / 128 KB CTA per SM kept issue utilisation at 40-60 % -- all 16 warps wait on the
// same barrier-separated phases).  Rows of 16 lanes are still 64-byte coalesced segments.
constexpr int THREADS = 256;
constexpr int LOG_TILE = 14;

__host__ __device__ constexpr int log_lc_of(int rb) { return rb <= LOG_TILE - 4 ? 4 : LOG_TILE - rb; }
__host__ __device__ constexpr int n_rounds(int rb) { return (rb + 4) / 5; }
__host__ __device__ constexpr int q_of(int rb, int i) { return rb / n_rounds(rb) + (i < rb % n_rounds(rb) ? 1 : 0); }
__host__ __device__ constexpr int b0_of(int rb, int i) { return i == 0 ? 0 : b0_of(rb, i - 1) + q_of(rb, i - 1); }

enum { SWZ_NONE = 0, SWZ_TOP = 1, SWZ_LOW = 2 };

template <int RB, int LLC, int SWZ>
__device__ __forceinline__ uint32_t sidx(uint32_t t, uint32_t j) {
    constexpr uint32_t M = (1u << LLC) - 1;
    if (SWZ == SWZ_TOP) return (t << LLC) + (j ^ ((t >> (RB >= LLC ? RB - LLC : 0)) & M));
    if (SWZ == SWZ_LOW) return (t << LLC) + (j ^ (t & M));
    return (t << LLC) + j;
}

// Q radix-2 stages in registers.  Twiddles are (w, w') pairs for Shoup's constant product (w canonical, w' = floor(w 2^32/p)):
// IMAD.HI + 2 IMAD = 8 FMA-pipe cycles instead of 10, lazy [0,2p) operands accepted, data stays in Montgomery form.
// KS = twiddle-index stride of local bit 0.
template <int Q, bool INV, int KS>
__device__ __forceinline__ void stages(uint32_t (&x)[1 << Q], const uint2* __restrict__ tw, int u0, uint32_t koff, bool last_inverse,
                                       uint2 ninv) {
    if (INV) {
#pragma unroll
        for (int s = Q - 1; s >= 0; s--) {
            if (s == 0 && last_inverse) {
#pragma unroll
                for (int e = 0; e < (1 << Q); e += 2) {
                    uint32_t a = x[e], b = x[e + 1];
                    x[e] = bb::mul_shoup(a + b, ninv);
                    x[e + 1] = bb::mul_shoup(a - b + bb::P, ninv);
                }
            } else {
                const uint2* t = tw + ((size_t)1 << (u0 + s)) + koff;
#pragma unroll
                for (int el = 0; el < (1 << s); el++) {
                    const uint2 w = __ldg(t + (size_t)el * KS);
#pragma unroll
                    for (int eh = 0; eh < (1 << (Q - 1 - s)); eh++) {
                        const int e = (eh << (s + 1)) | el;
                        uint32_t a = x[e], b = x[e | (1 << s)];
                        x[e] = bb::add(a, b);
                        x[e | (1 << s)] = bb::mul_shoup(a - b + bb::P, w);      // lazy difference in (0, 2p)
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int s = 0; s < Q; s++) {
            const uint2* t = tw + ((size_t)1 << (u0 + s)) + koff;
#pragma unroll
            for (int el = 0; el < (1 << s); el++) {
                const uint2 w = __ldg(t + (size_t)el * KS);
#pragma unroll
                for (int eh = 0; eh < (1 << (Q - 1 - s)); eh++) {
                    const int e = (eh << (s + 1)) | el;
                    uint32_t a = x[e], m = bb::mul_shoup(x[e | (1 << s)], w);
                    x[e] = bb::add(a, m);
                    x[e | (1 << s)] = bb::sub(a, m);
                }
            }
        }
    }
}

// one register round over a [2^RB rows][2^LLC lanes] tile; UB = global bit index of row bit 0 (NLO strided, 0 transposed)
template <int RB, int LLC, int Q, int B0, int UB, bool INV, int SWZ, bool LANE_K, bool G_IN, bool G_OUT>
__device__ __forceinline__ void round_dev(uint32_t* __restrict__ sm, const uint2* __restrict__ tw, uint2 ninv,
                                          const uint32_t* __restrict__ gsrc, uint32_t* __restrict__ gdst, uint32_t live) {
    constexpr uint32_t LC = 1u << LLC;
    constexpr uint32_t TASKS = (1u << (RB - Q)) << LLC;
    constexpr int KS = 1 << (B0 + UB);
#pragma unroll 1
    for (uint32_t id = threadIdx.x; id < TASKS; id += THREADS) {
        const uint32_t j = id & (LC - 1), g = id >> LLC;
        const uint32_t tl = g & ((1u << B0) - 1);
        const uint32_t tbase = ((g >> B0) << (B0 + Q)) | tl;
        if (j >= live) continue;
        uint32_t x[1 << Q];
#pragma unroll
        for (int e = 0; e < (1 << Q); e++) {
            const uint32_t t = tbase + ((uint32_t)e << B0);
            x[e] = G_IN ? __ldg(gsrc + ((size_t)t << UB) + j) : sm[sidx<RB, LLC, SWZ>(t, j)];
        }
        const uint32_t koff = (LANE_K ? j : 0u) + (tl << UB);
        stages<Q, INV, KS>(x, tw, UB + B0, koff, INV && UB == 0 && B0 == 0, ninv);
#pragma unroll
        for (int e = 0; e < (1 << Q); e++) {
            const uint32_t t = tbase + ((uint32_t)e << B0);
            if (G_OUT) gdst[((size_t)t << UB) + j] = x[e];
            else sm[sidx<RB, LLC, SWZ>(t, j)] = x[e];
        }
    }
}

// all rounds of a pass, unrolled at compile time in processing order (inverse: high bits first; forward: low bits first)
template <int RB, int LLC, int UB, bool INV, int SWZ, bool LANE_K, bool G_FIRST, bool G_LAST, int I = 0>
__device__ __forceinline__ void pass_rounds(uint32_t* sm, const uint2* tw, uint2 ninv, const uint32_t* gsrc, uint32_t* gdst,
                                            uint32_t live) {
    constexpr int NR = n_rounds(RB);
    if constexpr (I < NR) {
        constexpr int R = INV ? NR - 1 - I : I;
        constexpr bool gi = G_FIRST && I == 0, go = G_LAST && I == NR - 1;
        round_dev<RB, LLC, q_of(RB, R), b0_of(RB, R), UB, INV, SWZ, LANE_K, gi, go>(sm, tw, ninv, gsrc, gdst, live);
        if constexpr (I + 1 < NR) __syncthreads();
        pass_rounds<RB, LLC, UB, INV, SWZ, LANE_K, G_FIRST, G_LAST, I + 1>(sm, tw, ninv, gsrc, gdst, live);
    }
}

// K1 (INV) / K3 (!INV): strided tile, rows <-> position bits [NLO, NLO+RB), lane j <-> position j0 + j
template <int RB, int NLO, bool INV>
__global__ void __launch_bounds__(THREADS) strided_kernel(const uint32_t* __restrict__ src, size_t src_col_stride,
                                                          uint32_t* __restrict__ dst, size_t dst_col_stride, int log_blowup,
                                                          const uint2* __restrict__ tw_all) {
    extern __shared__ uint32_t sm[];
    constexpr int LLC = log_lc_of(RB);
    constexpr int n = RB + NLO;
    const uint32_t j0 = blockIdx.x << LLC;
    const int c = INV ? 0 : (int)blockIdx.z;
    const int cosets = INV ? 1 : (1 << log_blowup);
    const uint32_t* s = (INV ? src + (size_t)blockIdx.y * src_col_stride : src + (((size_t)blockIdx.y * cosets + c) << n)) + j0;
    const uint2* tw = tw_all + ((size_t)c << n) + j0;         // lane part of the twiddle index is the position j0 + j
    uint32_t* d1 = dst + (size_t)blockIdx.y * dst_col_stride + j0;
    pass_rounds<RB, LLC, NLO, INV, INV ? SWZ_NONE : SWZ_TOP, true, true, INV>(sm, tw, make_uint2(0u, 0u), s, d1, 1u << LLC);
    if (!INV) {
        __syncthreads();
        // natural evaluation index k = j0 + j + (t << NLO) of coset c lands at row  bitrev_b(c)*N + bitrev_n(k)
        constexpr uint32_t ROWS = 1u << RB, TILE = ROWS << LLC;
        uint32_t* dcol = dst + (size_t)blockIdx.y * dst_col_stride + ((size_t)brev((uint32_t)c, log_blowup) << n);
#pragma unroll 4
        for (uint32_t e = threadIdx.x; e < TILE; e += THREADS) {
            const uint32_t p = e & (ROWS - 1), j = e >> RB;
            const uint32_t t = __brev(p) >> (32 - RB);
            dcol[((size_t)(__brev(j0 + j) >> (32 - NLO)) << RB) + p] = sm[sidx<RB, LLC, SWZ_TOP>(t, j)];
        }
    }
}

// K2a (INV) / K2b (!INV): transposed tile, lane l <-> one contiguous block of 2^RB positions (RB == n_lo)
template <int RB, bool INV>
__global__ void __launch_bounds__(THREADS) transposed_kernel(const uint32_t* __restrict__ src, size_t src_col_stride,
                                                             uint32_t* __restrict__ dst, int n, int log_blowup, size_t total_blocks,
                                                             const uint2* __restrict__ tw_all, uint2 ninv) {
    extern __shared__ uint32_t sm[];
    constexpr int LLC = log_lc_of(RB);
    constexpr uint32_t LC = 1u << LLC, ROWS = 1u << RB;
    const int c = INV ? 0 : (int)blockIdx.z;
    const int cosets = INV ? 1 : (1 << log_blowup);
    const size_t b_first = (size_t)blockIdx.x << LLC;
    const uint32_t live = (uint32_t)min((size_t)LC, total_blocks - b_first);
    const int log_bpc = n - RB;
    const uint2* tw = tw_all + ((size_t)c << n);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t l = warp; l < live; l += THREADS / 32) {
        const size_t b = b_first + l, col = b >> log_bpc, blk = b & (((size_t)1 << log_bpc) - 1);
        const uint32_t* p = src + col * src_col_stride + (blk << RB);
#pragma unroll 8
        for (uint32_t t = lane; t < ROWS; t += 32) sm[sidx<RB, LLC, SWZ_LOW>(t, l)] = __ldg(p + t);
    }
    __syncthreads();
    pass_rounds<RB, LLC, 0, INV, SWZ_LOW, false, false, false>(sm, tw, ninv, nullptr, nullptr, live);
    __syncthreads();
    for (uint32_t l = warp; l < live; l += THREADS / 32) {
        const size_t b = b_first + l, col = b >> log_bpc, blk = b & (((size_t)1 << log_bpc) - 1);
        uint32_t* p = dst + ((col * cosets + c) << n) + (blk << RB);
#pragma unroll 8
        for (uint32_t t = lane; t < ROWS; t += 32) p[t] = sm[sidx<RB, LLC, SWZ_LOW>(t, l)];
    }
}

}  // namespace nttf

// ---------------- host-side dispatch: returns false when no specialisation covers (n_hi, n_lo) ----------------
namespace nttf {

template <auto Kernel>
inline void ensure_smem() {
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << LOG_TILE);
        done = true;
    }
}

inline bool launch_strided(bool inv, int n_hi, int n_lo, dim3 grid, cudaStream_t st, const uint32_t* src, size_t src_col_stride,
                           uint32_t* dst, size_t dst_col_stride, int log_blowup, const uint2* tw) {
#define PB_CASE(RB, NLO)                                                                                                        \
    if (n_hi == RB && n_lo == NLO) {                                                                                            \
        const size_t smem = (size_t)4 << (RB + log_lc_of(RB));                                                                  \
        if (inv) {                                                                                                              \
            ensure_smem<strided_kernel<RB, NLO, true>>();                                                                       \
            strided_kernel<RB, NLO, true><<<grid, THREADS, smem, st>>>(src, src_col_stride, dst, dst_col_stride, log_blowup, tw); \
        } else {                                                                                                                \
            ensure_smem<strided_kernel<RB, NLO, false>>();                                                                      \
            strided_kernel<RB, NLO, false><<<grid, THREADS, smem, st>>>(src, src_col_stride, dst, dst_col_stride, log_blowup, tw); \
        }                                                                                                                       \
        return true;                                                                                                            \
    }
    PB_CASE(8, 8) PB_CASE(8, 9) PB_CASE(9, 9) PB_CASE(9, 10) PB_CASE(10, 10) PB_CASE(10, 11) PB_CASE(11, 11) PB_CASE(11, 12) PB_CASE(12, 12)
#undef PB_CASE
    return false;
}

inline bool launch_transposed(bool inv, int n, int n_lo, dim3 grid, cudaStream_t st, const uint32_t* src, size_t src_col_stride,
                              uint32_t* dst, int log_blowup, size_t total_blocks, const uint2* tw, uint2 ninv) {
#define PB_CASE(RB)                                                                                                               \
    if (n_lo == RB) {                                                                                                             \
        const size_t smem = (size_t)4 << (RB + log_lc_of(RB));                                                                    \
        if (inv) {                                                                                                                \
            ensure_smem<transposed_kernel<RB, true>>();                                                                           \
            transposed_kernel<RB, true><<<grid, THREADS, smem, st>>>(src, src_col_stride, dst, n, log_blowup, total_blocks, tw, ninv); \
        } else {                                                                                                                  \
            ensure_smem<transposed_kernel<RB, false>>();                                                                          \
            transposed_kernel<RB, false><<<grid, THREADS, smem, st>>>(src, src_col_stride, dst, n, log_blowup, total_blocks, tw, ninv); \
        }                                                                                                                         \
        return true;                                                                                                              \
    }
    PB_CASE(8) PB_CASE(9) PB_CASE(10) PB_CASE(11) PB_CASE(12)
#undef PB_CASE
    return false;
}

inline bool supported(int n_hi, int n_lo) {
    return n_hi >= 8 && n_hi <= 12 && (n_lo == n_hi || n_lo == n_hi + 1) && n_lo <= 12;     // 16 <= n <= 24
}

}  // namespace nttf
