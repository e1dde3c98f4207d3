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

// same fold with the challenge read from device memory (Montgomery Ext4, written by p2::fri_challenge_kernel); c = (2*shift)^-1
__global__ void __launch_bounds__(256) fold_dev_beta_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t half,
                                                            const uint32_t* __restrict__ inv_tw, const uint32_t* __restrict__ beta4, uint32_t c,
                                                            uint32_t half_inv) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= half) return;
    const bb::E4 beta_c = {{bb::mul(__ldg(beta4), c), bb::mul(__ldg(beta4 + 1), c), bb::mul(__ldg(beta4 + 2), c), bb::mul(__ldg(beta4 + 3), c)}};
    uint4 l = __ldg(in + 2 * j), h = __ldg(in + 2 * j + 1);
    bb::E4 lo = {{l.x, l.y, l.z, l.w}}, hi = {{h.x, h.y, h.z, h.w}};
    bb::E4 s = bb::e4_scale(bb::e4_add(lo, hi), half_inv);
    bb::E4 d = bb::e4_scale(bb::e4_sub(lo, hi), __ldg(inv_tw + j));
    bb::E4 r = bb::e4_add(s, bb::e4_mul(beta_c, d));
    out[j] = make_uint4(r.c[0], r.c[1], r.c[2], r.c[3]);
}

// f[i] += g[i] over n words (multi-chip FRI: the reduced-opening codeword of a shorter height joins the folded codeword)
__global__ void __launch_bounds__(256) add_words_kernel(uint32_t* __restrict__ f, const uint32_t* __restrict__ g, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = bb::add(f[i], g[i]);
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

// ---------------- query phase: gather the openings a verifier needs for each sampled index ----------------
// One CTA per query.  Output (canonical words) per query:
//   [ r | trace row (W) | trace path (log_m x 8) | perm row (Wp) | perm path (log_m x 8)  (when Wp > 0) | quotient row (8) | quotient path (log_m x 8) |
//     per FRI layer i: pair row (8), path ((log_m-1-i) x 8) ]
// Merkle trees are node-major with level k at node offset 2^(h+1) - 2^(h-k+1); the sibling of leaf idx at level k is (idx>>k)^1.
namespace fri {

struct QueryDesc {
    const uint32_t* lde;        // trace LDE, column-major [W][M]
    const uint32_t* qlde;       // quotient chunk LDEs, column-major [8][M]
    const uint32_t* tree_t;     // digest layers of the trace commitment
    const uint32_t* tree_q;     // digest layers of the quotient commitment
    const uint32_t* plde;       // LogUp permutation-trace LDE, column-major [perm_width][M] (unused when perm_width = 0)
    const uint32_t* tree_p;     // digest layers of the permutation-trace commitment
    uint32_t perm_width;
    const uint32_t* fri_words;  // all FRI codewords back to back (layer i at word_off[i])
    const uint32_t* fri_trees;  // all FRI layer trees back to back (layer i at tree_off[i], in words)
    size_t m;
    uint32_t width;
    int log_m;
    int n_layers;
    size_t word_off[32];
    size_t tree_off[32];
};

__device__ __forceinline__ void copy_path(const uint32_t* tree, int log_h, size_t idx, uint32_t* out) {
    for (int k = 0; k < log_h; k++) {
        const size_t level_off = ((size_t)2 << log_h) - ((size_t)2 << (log_h - k));
        const uint32_t* node = tree + 8 * (level_off + ((idx >> k) ^ 1));
        for (int e = threadIdx.x; e < 8; e += blockDim.x) out[8 * k + e] = bb::from_monty(node[e]);
    }
}

__global__ void __launch_bounds__(256) gather_queries_kernel(QueryDesc d, const uint32_t* __restrict__ indices, uint32_t* __restrict__ out,
                                                             size_t words_per_query) {
    const size_t r = indices[blockIdx.x];
    uint32_t* o = out + (size_t)blockIdx.x * words_per_query;
    if (threadIdx.x == 0) o[0] = (uint32_t)r;
    o += 1;
    for (uint32_t c = threadIdx.x; c < d.width; c += blockDim.x) o[c] = bb::from_monty(d.lde[(size_t)c * d.m + r]);
    o += d.width;
    copy_path(d.tree_t, d.log_m, r, o);
    o += 8 * d.log_m;
    if (d.perm_width) {
        for (uint32_t c = threadIdx.x; c < d.perm_width; c += blockDim.x) o[c] = bb::from_monty(d.plde[(size_t)c * d.m + r]);
        o += d.perm_width;
        copy_path(d.tree_p, d.log_m, r, o);
        o += 8 * d.log_m;
    }
    for (uint32_t c = threadIdx.x; c < 8; c += blockDim.x) o[c] = bb::from_monty(d.qlde[(size_t)c * d.m + r]);
    o += 8;
    copy_path(d.tree_q, d.log_m, r, o);
    o += 8 * d.log_m;
    for (int i = 0; i < d.n_layers; i++) {
        const int log_h = d.log_m - 1 - i;
        const size_t j = r >> (i + 1);
        const uint32_t* row = d.fri_words + d.word_off[i] + 8 * j;
        for (uint32_t e = threadIdx.x; e < 8; e += blockDim.x) o[e] = bb::from_monty(row[e]);
        o += 8;
        copy_path(d.fri_trees + d.tree_off[i], log_h, j, o);
        o += 8 * log_h;
    }
}

}  // namespace fri

// ---------------- query gathering for the multi-chip prover: generic pieces, one launch per (commitment, matrix) / tree ----------------
namespace fri {
// out[q * wpq + off + c] = canonical(mat[c * m + (idx[q] >> shift)])
__global__ void gather_rows_kernel(const uint32_t* __restrict__ mat, size_t m, uint32_t width, int shift, const uint32_t* __restrict__ idx, uint32_t* __restrict__ out,
                                   size_t wpq, size_t off) {
    const size_t r = idx[blockIdx.x] >> shift;
    uint32_t* o = out + (size_t)blockIdx.x * wpq + off;
    for (uint32_t c = threadIdx.x; c < width; c += blockDim.x) o[c] = bb::from_monty(mat[(size_t)c * m + r]);
}
__global__ void gather_path_kernel(const uint32_t* __restrict__ tree, int log_h, int shift, const uint32_t* __restrict__ idx, uint32_t* __restrict__ out, size_t wpq,
                                   size_t off) {
    copy_path(tree, log_h, idx[blockIdx.x] >> shift, out + (size_t)blockIdx.x * wpq + off);
}
__global__ void gather_index_kernel(const uint32_t* __restrict__ idx, uint32_t* __restrict__ out, size_t wpq) {
    if (threadIdx.x == 0) out[(size_t)blockIdx.x * wpq] = idx[blockIdx.x];
}
struct FriDesc { const uint32_t* fri_words; const uint32_t* fri_trees; int log_m; int n_layers; size_t word_off[32]; size_t tree_off[32]; };
__global__ void gather_fri_kernel(FriDesc d, const uint32_t* __restrict__ idx, uint32_t* __restrict__ out, size_t wpq, size_t off) {
    const size_t r = idx[blockIdx.x];
    uint32_t* o = out + (size_t)blockIdx.x * wpq + off;
    for (int i = 0; i < d.n_layers; i++) {
        const int log_h = d.log_m - 1 - i;
        const size_t j = r >> (i + 1);
        const uint32_t* row = d.fri_words + d.word_off[i] + 8 * j;
        for (uint32_t e = threadIdx.x; e < 8; e += blockDim.x) o[e] = bb::from_monty(row[e]);
        o += 8;
        copy_path(d.fri_trees + d.tree_off[i], log_h, j, o);
        o += 8 * log_h;
    }
}
}  // namespace fri
