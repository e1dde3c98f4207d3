// Stage 3b of the north-star path: arity-2 FRI folding on bit-reversed Ext4 evaluations:
//   f'[j] = (lo + hi)/2 + beta * (lo - hi) / (2 x_j),   lo = f[2j], hi = f[2j+1],  x_j = shift * w_len^{bitrev(j)}.
// Codewords are [len][4] (16 B per Ext4 element, fold partners adjacent => two 128-bit loads per output).
// inv_tw[j] = w_len^{-bitrev(j)}; the table of the first layer serves every later layer (its prefix is the next
// layer's table), and the per-layer scalar (2*shift)^-1 is folded into beta on the host.
#pragma once
#include "bb31.cuh"

namespace fri {

__global__ void __launch_bounds__(256) fold_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t half,
                                                   const uint32_t* __restrict__ inv_tw, bb::E4 beta_c, uint32_t half_inv) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= half) return;
    uint4 l = __ldg(in + 2 * j), h = __ldg(in + 2 * j + 1);
    bb::E4 lo = {{l.x, l.y, l.z, l.w}}, hi = {{h.x, h.y, h.z, h.w}};
    bb::E4 s = bb::e4_scale(bb::e4_add(lo, hi), half_inv);
    bb::E4 d = bb::e4_scale(bb::e4_sub(lo, hi), __ldg(inv_tw + j));
    bb::E4 r = bb::e4_add(s, bb::e4_mul(beta_c, d));
    out[j] = make_uint4(r.c[0], r.c[1], r.c[2], r.c[3]);
}

// FRI input codeword from the two committed quotient-chunk LDEs (8 base columns, column-major, height m):
// f[r] = Q0[r] + gamma * Q1[r]
__global__ void __launch_bounds__(256) combine_chunks_kernel(const uint32_t* __restrict__ qlde, size_t m, bb::E4 gamma, uint4* __restrict__ f) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    bb::E4 a, b;
#pragma unroll
    for (int l = 0; l < 4; l++) { a.c[l] = __ldg(qlde + (size_t)l * m + r); b.c[l] = __ldg(qlde + (size_t)(4 + l) * m + r); }
    bb::E4 v = bb::e4_add(a, bb::e4_mul(gamma, b));
    f[r] = make_uint4(v.c[0], v.c[1], v.c[2], v.c[3]);
}

}  // namespace fri
