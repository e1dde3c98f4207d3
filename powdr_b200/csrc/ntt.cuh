// Stage 1 of the north-star path: batched coset low-degree extension of column-major trace columns
// (call site main_trace_commit behind /root/reference/openvm-riscv/src/lib.rs:332).
//
// Per column:  evals over H (natural order)  --iNTT-->  coefficients  --coset NTT x 2^b-->  evals over shift*H',
// rows bit-reversed.  B200 mapping:
//   * inverse = decimation-in-frequency (natural in, bit-reversed coefficient slots out), forward = decimation-in-time
//     on those slots (natural out): no standalone permutation pass; the one bit reversal is folded into the last store.
//   * the coset shift is folded into the forward twiddles (w_u[k] = shift^(2^(n-1-u)) * omega^k), the 1/N into the last
//     inverse stage: no scaling passes.
//   * three kernels per column batch -- K1 high inverse stages on strided [2^n_hi x 32] tiles, K2 low inverse + low
//     forward stages on contiguous tiles, K3 high forward stages on strided tiles + bit-reversed store -- so each column
//     crosses HBM exactly once in (4N bytes) and once out (4N*2^b); the two intermediates are sized to stay in the 126 MB L2.
//   * lanes always walk consecutive columns of the tile (128 B coalesced rows, conflict-free shared memory); the butterfly
//     network runs across rows only.
#pragma once
#include "bb31.cuh"

namespace ntt {

constexpr int K13_THREADS = 1024;
constexpr int K2_THREADS = 512;
constexpr int LOG_TILE_LO = 13;       // K2 tile: 2^13 contiguous elements (two 32 KB shared buffers)
constexpr int LOG_TILE_HI_MAX = 15;   // K1/K3 tile: at most 2^15 elements (128 KB)

__device__ __forceinline__ uint32_t brev(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// K1: inverse stages u = n-1 .. n_lo on a strided tile (rows t <-> bits [n_lo, n) of the position)
__global__ void __launch_bounds__(K13_THREADS) inv_hi_kernel(const uint32_t* __restrict__ in, size_t in_col_stride,
                                                             uint32_t* __restrict__ tmp, int n, int n_lo, int log_lc,
                                                             const uint32_t* __restrict__ tw_inv) {
    extern __shared__ uint32_t sm[];
    const int n_hi = n - n_lo;
    const uint32_t lc = 1u << log_lc, rows = 1u << n_hi, tile = rows << log_lc;
    const uint32_t j0 = blockIdx.x << log_lc;
    const uint32_t* src = in + (size_t)blockIdx.y * in_col_stride;
    uint32_t* dst = tmp + ((size_t)blockIdx.y << n);
    for (uint32_t e = threadIdx.x; e < tile; e += K13_THREADS) {
        uint32_t t = e >> log_lc, j = e & (lc - 1);
        sm[e] = __ldg(src + ((size_t)t << n_lo) + j0 + j);
    }
    __syncthreads();
    for (int v = n_hi - 1; v >= 0; v--) {
        const uint32_t* tw = tw_inv + ((size_t)1 << (n_lo + v));
        const uint32_t mask = (1u << v) - 1;
        for (uint32_t q = threadIdx.x; q < tile / 2; q += K13_THREADS) {
            uint32_t j = q & (lc - 1), pr = q >> log_lc;
            uint32_t t0 = ((pr >> v) << (v + 1)) | (pr & mask), t1 = t0 + (1u << v);
            uint32_t k = j0 + j + ((t0 & mask) << n_lo);
            uint32_t a = sm[(t0 << log_lc) + j], b = sm[(t1 << log_lc) + j];
            sm[(t0 << log_lc) + j] = bb::add(a, b);
            sm[(t1 << log_lc) + j] = bb::mul(bb::sub(a, b), __ldg(tw + k));
        }
        __syncthreads();
    }
    for (uint32_t e = threadIdx.x; e < tile; e += K13_THREADS) {
        uint32_t t = e >> log_lc, j = e & (lc - 1);
        dst[((size_t)t << n_lo) + j0 + j] = sm[e];
    }
}

// K2: contiguous tile. inverse stages u = n_lo-1 .. 0 (1/N folded into u = 0), then per coset forward stages u = 0 .. n_lo-1.
__global__ void __launch_bounds__(K2_THREADS) lo_kernel(const uint32_t* __restrict__ src, size_t src_col_stride,
                                                        uint32_t* __restrict__ tmp2, int n, int n_lo, int log_tile,
                                                        const uint32_t* __restrict__ tw_inv, const uint32_t* __restrict__ tw_fwd,
                                                        uint32_t ninv, int cosets) {
    extern __shared__ uint32_t sm[];
    const uint32_t tile = 1u << log_tile;
    uint32_t* A = sm;
    uint32_t* B = sm + tile;
    const size_t base = (size_t)blockIdx.x << log_tile;
    const uint32_t* s = src + (size_t)blockIdx.y * src_col_stride + base;
    for (uint32_t e = threadIdx.x; e < tile; e += K2_THREADS) A[e] = __ldg(s + e);
    __syncthreads();
    for (int u = n_lo - 1; u >= 0; u--) {
        const uint32_t* tw = tw_inv + ((size_t)1 << u);
        const uint32_t mask = (1u << u) - 1;
        for (uint32_t q = threadIdx.x; q < tile / 2; q += K2_THREADS) {
            uint32_t i0 = ((q >> u) << (u + 1)) | (q & mask), i1 = i0 + (1u << u);
            uint32_t a = A[i0], b = A[i1];
            if (u > 0) {
                A[i0] = bb::add(a, b);
                A[i1] = bb::mul(bb::sub(a, b), __ldg(tw + (i0 & mask)));
            } else {
                A[i0] = bb::mul(bb::add(a, b), ninv);
                A[i1] = bb::mul(bb::sub(a, b), ninv);
            }
        }
        __syncthreads();
    }
    for (int c = 0; c < cosets; c++) {
        const uint32_t* twc = tw_fwd + ((size_t)c << n);
        for (int u = 0; u < n_lo; u++) {
            const uint32_t* tw = twc + ((size_t)1 << u);
            const uint32_t mask = (1u << u) - 1;
            const uint32_t* rd = u == 0 ? A : B;
            for (uint32_t q = threadIdx.x; q < tile / 2; q += K2_THREADS) {
                uint32_t i0 = ((q >> u) << (u + 1)) | (q & mask), i1 = i0 + (1u << u);
                uint32_t a = rd[i0], t = bb::mul(rd[i1], __ldg(tw + (i0 & mask)));
                B[i0] = bb::add(a, t);
                B[i1] = bb::sub(a, t);
            }
            __syncthreads();
        }
        uint32_t* d = tmp2 + (((size_t)blockIdx.y * cosets + c) << n) + base;
        for (uint32_t e = threadIdx.x; e < tile; e += K2_THREADS) d[e] = B[e];
        __syncthreads();
    }
}

// K3: forward stages u = n_lo .. n-1 on a strided tile of coset c, then the bit-reversed store:
// natural evaluation index k of coset c lands at row  bitrev_b(c)*N + bitrev_n(k).
__global__ void __launch_bounds__(K13_THREADS) fwd_hi_kernel(const uint32_t* __restrict__ tmp2, uint32_t* __restrict__ out,
                                                             size_t out_col_stride, int n, int n_lo, int log_lc, int log_blowup,
                                                             const uint32_t* __restrict__ tw_fwd) {
    extern __shared__ uint32_t sm[];
    const int n_hi = n - n_lo;
    const uint32_t lc = 1u << log_lc, rows = 1u << n_hi, tile = rows << log_lc;
    const uint32_t j0 = blockIdx.x << log_lc;
    const int cosets = 1 << log_blowup;
    const int c = blockIdx.z;
    const uint32_t* src = tmp2 + (((size_t)blockIdx.y * cosets + c) << n);
    const int swz_shift = n_hi >= log_lc ? n_hi - log_lc : 0;
    const uint32_t swz_mask = n_hi >= log_lc ? lc - 1 : 0;
    for (uint32_t e = threadIdx.x; e < tile; e += K13_THREADS) {
        uint32_t t = e >> log_lc, j = e & (lc - 1);
        sm[(t << log_lc) + (j ^ ((t >> swz_shift) & swz_mask))] = src[((size_t)t << n_lo) + j0 + j];
    }
    __syncthreads();
    const uint32_t* twc = tw_fwd + ((size_t)c << n);
    for (int v = 0; v < n_hi; v++) {
        const uint32_t* tw = twc + ((size_t)1 << (n_lo + v));
        const uint32_t mask = (1u << v) - 1;
        for (uint32_t q = threadIdx.x; q < tile / 2; q += K13_THREADS) {
            uint32_t j = q & (lc - 1), pr = q >> log_lc;
            uint32_t t0 = ((pr >> v) << (v + 1)) | (pr & mask), t1 = t0 + (1u << v);
            uint32_t k = j0 + j + ((t0 & mask) << n_lo);
            uint32_t x0 = (t0 << log_lc) + (j ^ ((t0 >> swz_shift) & swz_mask));
            uint32_t x1 = (t1 << log_lc) + (j ^ ((t1 >> swz_shift) & swz_mask));
            uint32_t a = sm[x0], w = bb::mul(sm[x1], __ldg(tw + k));
            sm[x0] = bb::add(a, w);
            sm[x1] = bb::sub(a, w);
        }
        __syncthreads();
    }
    uint32_t* dcol = out + (size_t)blockIdx.y * out_col_stride + ((size_t)brev((uint32_t)c, log_blowup) << n);
    for (uint32_t e = threadIdx.x; e < tile; e += K13_THREADS) {
        uint32_t p = e & (rows - 1), j = e >> n_hi;             // consecutive threads -> consecutive output rows
        uint32_t t = brev(p, n_hi);
        uint32_t v = sm[(t << log_lc) + (j ^ ((t >> swz_shift) & swz_mask))];
        dcol[((size_t)brev(j0 + j, n_lo) << n_hi) + p] = v;
    }
}

// bit-reversal of rows inside each column (used to feed bit-reversed quotient chunks back into the LDE)
__global__ void bitrev_rows_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n, size_t n_cols) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_cols << n)) return;
    size_t col = i >> n;
    uint32_t r = (uint32_t)(i & (((size_t)1 << n) - 1));
    out[(col << n) + brev(r, n)] = in[i];
}

}  // namespace ntt
