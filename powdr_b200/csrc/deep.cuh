// Openings and the DEEP reduced opening (the first "next" row widened in round 1): what turns the committed
// trace / quotient LDEs into the single Ext4 codeword FRI folds.  Backend counterpart: the opening phase of pcs_opening behind
// sdk.app_prover(exe)?.prove(stdin) (/root/reference/openvm-riscv/src/lib.rs:327-332); APC AIRs are single-row
// (/root/reference/openvm/src/powdr_extension/chip.rs:99-108), so every column is opened at ONE point zeta.
//
//   y_k      = f_k(zeta)                      barycentric over the N trace-domain evaluations:  4 mulmod per element, one pass
//   ro(x_r)  = (sum_j gamma^j f_j(x_r) - sum_j gamma^j y_j) / (x_r - zeta)   over the 2N LDE rows: 4 Shoup products per element
// Both stream column-major matrices with lane = row (coalesced), like the leaf and quotient kernels.
#pragma once
#include "bb31.cuh"

namespace deep {

__device__ __forceinline__ bb::E4 e4_inv(bb::E4 a) {
    // norm to F_p[y]/(y^2-11) (y = x^2), then to F_p; all limbs Montgomery
    const uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    const uint32_t W = bb::W11_M;
    uint32_t A0 = bb::add(bb::mul(a0, a0), bb::mul(W, bb::mul(a2, a2)));
    uint32_t A1 = bb::dbl(bb::mul(a0, a2));
    uint32_t B0 = bb::add(bb::mul(a1, a1), bb::mul(W, bb::mul(a3, a3)));
    uint32_t B1 = bb::dbl(bb::mul(a1, a3));
    uint32_t n0 = bb::sub(A0, bb::mul(W, B1)), n1 = bb::sub(A1, B0);
    uint32_t d = bb::inv(bb::sub(bb::mul(n0, n0), bb::mul(W, bb::mul(n1, n1))));
    uint32_t i0 = bb::mul(n0, d), i1 = bb::neg(bb::mul(n1, d));
    bb::E4 conj = {{a0, bb::neg(a1), a2, bb::neg(a3)}};
    bb::E4 s = {{i0, 0u, i1, 0u}};
    return bb::e4_mul(conj, s);
}

// barycentric weights over the subgroup H of size 2^log_n at the point z:  w_i = omega^i / (z - omega^i)
__global__ void __launch_bounds__(256) bary_weights_kernel(uint4* __restrict__ w, int log_n, uint32_t omega_m, bb::E4 z) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_n)) return;
    const uint32_t x = bb::pow(omega_m, i);
    bb::E4 d = z;
    d.c[0] = bb::sub(d.c[0], x);
    bb::E4 r = bb::e4_scale(e4_inv(d), x);
    w[i] = make_uint4(r.c[0], r.c[1], r.c[2], r.c[3]);
}

constexpr int EV_COLS = 8;       // columns per CTA (each weight load is reused 8 times)
constexpr int EV_THREADS = 256;

// partial[k][split] = sum over this CTA's rows of mat[k][i] * w_i
__global__ void __launch_bounds__(EV_THREADS) eval_partial_kernel(const uint32_t* __restrict__ mat, size_t n, uint32_t width,
                                                                  const uint4* __restrict__ w, uint4* __restrict__ partial,
                                                                  uint32_t rows_per_cta) {
    const uint32_t k0 = blockIdx.x * EV_COLS;
    const size_t r0 = (size_t)blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    bb::E4 acc[EV_COLS];
#pragma unroll
    for (int c = 0; c < EV_COLS; c++) acc[c] = bb::E4{{0u, 0u, 0u, 0u}};
    // software-pipelined: the loads of row i + 256 are in flight while row i is multiplied (ncu r01c: without it the kernel
    // issued on 23 % of cycles, one dependent load round trip per iteration)
    size_t i = r0 + threadIdx.x;
    uint4 wi = make_uint4(0u, 0u, 0u, 0u);
    uint32_t f[EV_COLS];
    if (i < r1) {
        wi = __ldg(w + i);
#pragma unroll
        for (int c = 0; c < EV_COLS; c++) f[c] = (k0 + c < width) ? __ldg(mat + (size_t)(k0 + c) * n + i) : 0u;
    }
    while (i < r1) {
        const size_t in = i + EV_THREADS;
        uint4 wn = make_uint4(0u, 0u, 0u, 0u);
        uint32_t fn[EV_COLS];
        if (in < r1) {
            wn = __ldg(w + in);
#pragma unroll
            for (int c = 0; c < EV_COLS; c++) fn[c] = (k0 + c < width) ? __ldg(mat + (size_t)(k0 + c) * n + in) : 0u;
        }
#pragma unroll
        for (int c = 0; c < EV_COLS; c++) {
            acc[c].c[0] = bb::add(acc[c].c[0], bb::mul(f[c], wi.x));
            acc[c].c[1] = bb::add(acc[c].c[1], bb::mul(f[c], wi.y));
            acc[c].c[2] = bb::add(acc[c].c[2], bb::mul(f[c], wi.z));
            acc[c].c[3] = bb::add(acc[c].c[3], bb::mul(f[c], wi.w));
        }
        wi = wn;
#pragma unroll
        for (int c = 0; c < EV_COLS; c++) f[c] = fn[c];
        i = in;
    }
    __shared__ uint32_t red[EV_THREADS / 32][EV_COLS * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < EV_COLS; c++)
#pragma unroll
        for (int l = 0; l < 4; l++) {
            uint32_t v = acc[c].c[l];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = bb::add(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) red[warp][c * 4 + l] = v;
        }
    __syncthreads();
    if (threadIdx.x < EV_COLS * 4) {
        uint32_t v = 0;
#pragma unroll
        for (int q = 0; q < EV_THREADS / 32; q++) v = bb::add(v, red[q][threadIdx.x]);
        const uint32_t c = threadIdx.x >> 2, l = threadIdx.x & 3;
        if (k0 + c < width) reinterpret_cast<uint32_t*>(partial + ((size_t)(k0 + c) * gridDim.y + blockIdx.y))[l] = v;
    }
}

// y_k = pref * sum_splits partial[k][split]
__global__ void eval_finalize_kernel(const uint4* __restrict__ partial, uint32_t width, uint32_t splits, bb::E4 pref, uint4* __restrict__ ys) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= width) return;
    bb::E4 s = {{0u, 0u, 0u, 0u}};
    for (uint32_t q = 0; q < splits; q++) {
        const uint4 p = partial[(size_t)k * splits + q];
        s.c[0] = bb::add(s.c[0], p.x); s.c[1] = bb::add(s.c[1], p.y); s.c[2] = bb::add(s.c[2], p.z); s.c[3] = bb::add(s.c[3], p.w);
    }
    const bb::E4 y = bb::e4_mul(s, pref);
    ys[k] = make_uint4(y.c[0], y.c[1], y.c[2], y.c[3]);
}

// reduced opening over the LDE domain (bit-reversed rows).  gp[j*4+l] = Shoup pair of limb l of gamma^j (canonical), so
// gamma^j * f_j costs four 8-cycle constant products; `cols` has one base pointer per opened column.  The m rows passed
// are rows [row0, row0 + m) of the 2^log_m-row domain (row0 = 0, m = 2^log_m for the whole domain).
__global__ void __launch_bounds__(256) deep_quotient_kernel(const uint32_t* const* __restrict__ cols, uint32_t n_cols, size_t m, int log_m, size_t row0,
                                                            uint32_t shift_m, uint32_t omega_m, const uint2* __restrict__ gp,
                                                            bb::E4 ysum, bb::E4 zeta, uint4* __restrict__ out) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
    for (uint32_t j = 0; j < n_cols; j++) {
        const uint32_t f = __ldg(cols[j] + r);
        const uint4 g01 = __ldg(reinterpret_cast<const uint4*>(gp) + 2 * j), g23 = __ldg(reinterpret_cast<const uint4*>(gp) + 2 * j + 1);
        a0 = bb::add(a0, bb::mul_shoup(f, make_uint2(g01.x, g01.y)));
        a1 = bb::add(a1, bb::mul_shoup(f, make_uint2(g01.z, g01.w)));
        a2 = bb::add(a2, bb::mul_shoup(f, make_uint2(g23.x, g23.y)));
        a3 = bb::add(a3, bb::mul_shoup(f, make_uint2(g23.z, g23.w)));
    }
    bb::E4 acc = {{bb::sub(a0, ysum.c[0]), bb::sub(a1, ysum.c[1]), bb::sub(a2, ysum.c[2]), bb::sub(a3, ysum.c[3])}};
    const uint32_t x = bb::mul(shift_m, bb::pow(omega_m, (uint64_t)(__brev((uint32_t)(row0 + r)) >> (32 - log_m))));
    bb::E4 d = {{bb::sub(x, zeta.c[0]), bb::neg(zeta.c[1]), bb::neg(zeta.c[2]), bb::neg(zeta.c[3])}};
    const bb::E4 v = bb::e4_mul(acc, e4_inv(d));
    out[r] = make_uint4(v.c[0], v.c[1], v.c[2], v.c[3]);
}

}  // namespace deep
