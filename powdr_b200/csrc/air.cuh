// Stage 2 of the north-star path: per-row evaluation of an autoprecompile's constraint polynomials on the LDE domain,
// folded with powers of alpha into the quotient chunks.  What is evaluated is exactly what
// PowdrAir::eval feeds to builder.assert_zero -- every machine.constraints[i] on row_slice(0), no next-row access, no
// selectors (/root/reference/openvm/src/powdr_extension/chip.rs:94-130) -- expressed in the reference's own stack-machine
// bytecode (/root/reference/openvm/cuda/src/expr_eval.cuh:12-89), re-packed by the host into one word per instruction.
//
// One thread per LDE row; column-major LDE => lane = row => every PUSH_APC is a coalesced 128 B warp load.  The program
// is warp-uniform (no divergence); the evaluation stack lives in shared memory at [slot][thread] (conflict-free) with the
// top of stack cached in a register.  HBM-bound by design: 4*N'*W bytes in, 16*N' out.
#pragma once
#include "bb31.cuh"

namespace air {

// packed instruction: op << 28 | arg   (arg = column index | constant-pool index)
enum : uint32_t { OP_PUSH_APC = 0, OP_PUSH_CONST = 1, OP_ADD = 2, OP_SUB = 3, OP_MUL = 4, OP_NEG = 5, OP_INV_OR_ZERO = 6 };
constexpr int STACK_CAPACITY = 16;   // STACK_CAPACITY, expr_eval.cuh:22
constexpr int THREADS = 256;

struct Span { uint32_t off, len; };

// evaluates one packed expression for row r; `stk` points at this thread's column of the shared stack (stride THREADS)
__device__ __forceinline__ uint32_t eval_packed(const uint32_t* __restrict__ code, uint32_t off, uint32_t len,
                                                const uint32_t* __restrict__ pool, const uint32_t* __restrict__ mat,
                                                size_t col_stride, size_t r, uint32_t* stk) {
    uint32_t tos = 0;
    int sp = 0;                                   // number of entries BELOW tos that live in shared memory
    bool have = false;
    for (uint32_t ip = off; ip < off + len; ip++) {
        uint32_t w = __ldg(code + ip);
        uint32_t op = w >> 28, arg = w & 0x0fffffffu;
        switch (op) {
        case OP_PUSH_APC:
        case OP_PUSH_CONST: {
            uint32_t v = op == OP_PUSH_APC ? __ldg(mat + (size_t)arg * col_stride + r) : __ldg(pool + arg);
            if (have) { stk[sp * THREADS] = tos; sp++; }
            tos = v; have = true;
            break;
        }
        case OP_ADD: { sp--; tos = bb::add(stk[sp * THREADS], tos); break; }
        case OP_SUB: { sp--; tos = bb::sub(stk[sp * THREADS], tos); break; }
        case OP_MUL: { sp--; tos = bb::mul(stk[sp * THREADS], tos); break; }
        case OP_NEG: tos = bb::neg(tos); break;
        default: tos = bb::inv(tos); break;       // OP_INV_OR_ZERO: inv(0) = 0
        }
    }
    return tos;
}

// quotient over the LDE domain (log_blowup 1: quotient domain == LDE domain).  alpha_pows[k] = alpha^(C-1-k) so that
// sum_k alpha_pows[k]*c_k equals the Horner fold acc = acc*alpha + c_k in declaration order.
__global__ void __launch_bounds__(THREADS) quotient_kernel(const uint32_t* __restrict__ code, const Span* __restrict__ spans,
                                                           uint32_t n_constraints, const uint32_t* __restrict__ pool,
                                                           const uint32_t* __restrict__ lde, int log_n,
                                                           const uint32_t* __restrict__ alpha_pows /* [C][4] */,
                                                           uint32_t zinv0, uint32_t zinv1, uint32_t* __restrict__ out,
                                                           int apply_zinv) {
    __shared__ uint32_t stack[STACK_CAPACITY * THREADS];
    const size_t m = (size_t)2 << log_n;
    const size_t r = (size_t)blockIdx.x * THREADS + threadIdx.x;
    if (r >= m) return;
    uint32_t* stk = stack + threadIdx.x;
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    for (uint32_t k = 0; k < n_constraints; k++) {
        Span s = spans[k];
        uint32_t c = eval_packed(code, s.off, s.len, pool, lde, m, r, stk);
        const uint4 ap = __ldg(reinterpret_cast<const uint4*>(alpha_pows) + k);
        acc0 = bb::add(acc0, bb::mul(c, ap.x));
        acc1 = bb::add(acc1, bb::mul(c, ap.y));
        acc2 = bb::add(acc2, bb::mul(c, ap.z));
        acc3 = bb::add(acc3, bb::mul(c, ap.w));
    }
    const size_t n = (size_t)1 << log_n;
    const size_t chunk = r >> log_n, j = r & (n - 1);
    if (apply_zinv) {
        const uint32_t z = chunk ? zinv1 : zinv0;
        acc0 = bb::mul(acc0, z); acc1 = bb::mul(acc1, z); acc2 = bb::mul(acc2, z); acc3 = bb::mul(acc3, z);
        out[(chunk * 4 + 0) * n + j] = acc0;
        out[(chunk * 4 + 1) * n + j] = acc1;
        out[(chunk * 4 + 2) * n + j] = acc2;
        out[(chunk * 4 + 3) * n + j] = acc3;
    } else {                                   // raw fold, [4][m] (parity on random traces)
        out[0 * m + r] = acc0; out[1 * m + r] = acc1; out[2 * m + r] = acc2; out[3 * m + r] = acc3;
    }
}

// generic height version of the raw fold (any matrix height, e.g. the trace itself): out [4][height]
__global__ void __launch_bounds__(THREADS) constraint_fold_kernel(const uint32_t* __restrict__ code, const Span* __restrict__ spans,
                                                                  uint32_t n_constraints, const uint32_t* __restrict__ pool,
                                                                  const uint32_t* __restrict__ mat, size_t height,
                                                                  const uint32_t* __restrict__ alpha_pows, uint32_t* __restrict__ out) {
    __shared__ uint32_t stack[STACK_CAPACITY * THREADS];
    const size_t r = (size_t)blockIdx.x * THREADS + threadIdx.x;
    if (r >= height) return;
    uint32_t* stk = stack + threadIdx.x;
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    for (uint32_t k = 0; k < n_constraints; k++) {
        Span s = spans[k];
        uint32_t c = eval_packed(code, s.off, s.len, pool, mat, height, r, stk);
        const uint4 ap = __ldg(reinterpret_cast<const uint4*>(alpha_pows) + k);
        acc0 = bb::add(acc0, bb::mul(c, ap.x));
        acc1 = bb::add(acc1, bb::mul(c, ap.y));
        acc2 = bb::add(acc2, bb::mul(c, ap.z));
        acc3 = bb::add(acc3, bb::mul(c, ap.w));
    }
    out[0 * height + r] = acc0; out[1 * height + r] = acc1; out[2 * height + r] = acc2; out[3 * height + r] = acc3;
}

}  // namespace air
