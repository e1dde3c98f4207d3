// Host-side Poseidon2 (BabyBear, width 16, x^7, 8 + 13 rounds) for the transcript -- see transcript_host.h.
// Compiled by the host compiler alone (no CUDA headers); linked into libpowdr_b200.so.
//
// AVX-512 version: the 16 state words are the 16 lanes of one register.
//   * Montgomery product of two vectors: even and odd lanes through vpmuldq / vpmuludq (three multiplies each), t - m p has a zero
//     low word so the signed quotient is the high half.  Inside the S-box the three inner products stay SIGNED in (-p, p) (the
//     input is s + rc - p), so only the last one is made canonical (one add + unsigned min).
//   * external layer: y_i = S + x_i + 2 x_(i+1) inside every 4-lane group (S = the group's sum) is circ(2,3,1,1); the column sums
//     across the four groups are two 128-bit-lane rotations.
//   * internal rounds: only lane 0 goes through the S-box, so lane 0 lives in a scalar register for those 13 rounds and the
//     serial chain is scalar S-box -> sum -> next S-box; the diagonal product and the horizontal sum of lanes 1..15 are
//     independent of it and overlap.  With the Plonky3 diagonal (d0 = -2) lane 0's update sum - 2 s0 is (sum of the other lanes) - s0:
//     one subtraction on the chain instead of a product.
// Measured (2.1 GHz Xeon of the build container): 0.57 us per permutation, scalar version 2.0-2.2 us.
#include "transcript_host.h"

#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace pbhost {

namespace {
constexpr uint32_t P = 0x78000001u;         // 2013265921
constexpr uint32_t NEG_PINV = 0x77ffffffu;  // -p^-1 mod 2^32
constexpr uint32_t PINV = 0x88000001u;      //  p^-1 mod 2^32

inline uint32_t red2p(uint32_t x) { const uint32_t y = x - P; return y < x ? y : x; }
inline uint32_t add(uint32_t a, uint32_t b) { return red2p(a + b); }
inline uint32_t mul(uint32_t a, uint32_t b) {
    const uint64_t t = (uint64_t)a * b;
    const uint32_t m = (uint32_t)t * NEG_PINV;
    return red2p((uint32_t)(((uint64_t)m * P + t) >> 32));
}
inline uint32_t sbox(uint32_t x) {
    const uint32_t x2 = mul(x, x), x3 = mul(x2, x), x4 = mul(x2, x2);
    return mul(x3, x4);
}
// signed Montgomery product: |a|, |b| < p  ->  a b R^-1 in (-p, p), no correction step (t - m p has a zero low word)
inline int32_t smul(int32_t a, int32_t b) {
    const int64_t t = (int64_t)a * b;
    const int32_t m = (int32_t)((uint32_t)t * PINV);
    return (int32_t)((t - (int64_t)m * (int64_t)P) >> 32);
}
// (s + rc)^7 with s, rc canonical: the sum is taken in [-p, p), the three inner products stay signed, one fix-up at the end
inline uint32_t sbox_rc(uint32_t s, uint32_t rc) {
    const int32_t x = (int32_t)(s + rc - P);
    const int32_t x2 = smul(x, x), x3 = smul(x2, x), x4 = smul(x2, x2);
    const uint32_t r = (uint32_t)smul(x3, x4);
    const uint32_t v = r + P;
    return v < r ? v : r;
}
void external_linear(uint32_t s[16]) {
    // M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]] on each 4-chunk as an add chain, then add the column sums
    for (int c = 0; c < 16; c += 4) {
        const uint32_t x0 = s[c], x1 = s[c + 1], x2 = s[c + 2], x3 = s[c + 3];
        const uint32_t t01 = add(x0, x1), t23 = add(x2, x3), t0123 = add(t01, t23);
        const uint32_t t01123 = add(t0123, x1), t01233 = add(t0123, x3);
        s[c + 3] = add(t01233, add(x0, x0));
        s[c + 1] = add(t01123, add(x2, x2));
        s[c] = add(t01123, t01);
        s[c + 2] = add(t01233, t23);
    }
    uint32_t q[4];
    for (int i = 0; i < 4; i++) q[i] = add(add(s[i], s[4 + i]), add(s[8 + i], s[12 + i]));
    for (int i = 0; i < 16; i++) s[i] = add(s[i], q[i & 3]);
}
}  // namespace

void permute_scalar(uint32_t s[16], const P2Host& k) {
    external_linear(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox(add(s[i], k.rc_ext[r][i]));
        external_linear(s);
    }
    for (int r = 0; r < 13; r++) {
        s[0] = sbox(add(s[0], k.rc_int[r]));
        uint32_t sum = 0;
        for (int i = 0; i < 16; i++) sum = add(sum, s[i]);
        for (int i = 0; i < 16; i++) s[i] = add(sum, mul(s[i], k.diag[i]));
    }
    for (int r = 4; r < 8; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox(add(s[i], k.rc_ext[r][i]));
        external_linear(s);
    }
}

#if defined(__x86_64__)
namespace {
#define PB_AVX512 __attribute__((target("avx512f,avx512dq"), always_inline)) inline

PB_AVX512 __m512i vadd(__m512i a, __m512i b) {
    const __m512i r = _mm512_add_epi32(a, b);
    return _mm512_min_epu32(r, _mm512_sub_epi32(r, _mm512_set1_epi32((int)P)));
}
PB_AVX512 __m512i vmul(__m512i a, __m512i b) {
    const __m512i vp = _mm512_set1_epi32((int)P), vmu = _mm512_set1_epi32((int)PINV);
    const __m512i pe = _mm512_mul_epu32(a, b);
    const __m512i po = _mm512_mul_epu32(_mm512_srli_epi64(a, 32), _mm512_srli_epi64(b, 32));
    const __m512i qe = _mm512_mul_epu32(_mm512_mul_epu32(pe, vmu), vp);      // (lo(t) * p^-1 mod 2^32) * p
    const __m512i qo = _mm512_mul_epu32(_mm512_mul_epu32(po, vmu), vp);
    const __m512i de = _mm512_sub_epi64(pe, qe), dn = _mm512_sub_epi64(po, qo);   // low words zero, high words in (-p, p)
    const __m512i t = _mm512_mask_blend_epi32((__mmask16)0xAAAA, _mm512_srli_epi64(de, 32), dn);
    return _mm512_min_epu32(t, _mm512_add_epi32(t, vp));
}
// signed lanes in (-p, p) in, signed lanes in (-p, p) out: the product without the closing correction
PB_AVX512 __m512i vsmul(__m512i a, __m512i b) {
    const __m512i vp = _mm512_set1_epi32((int)P), vmu = _mm512_set1_epi32((int)PINV);
    const __m512i pe = _mm512_mul_epi32(a, b);
    const __m512i po = _mm512_mul_epi32(_mm512_srli_epi64(a, 32), _mm512_srli_epi64(b, 32));
    const __m512i qe = _mm512_mul_epi32(_mm512_mul_epu32(pe, vmu), vp);      // the low word of lo(t) p^-1 is what matters, then signed times p
    const __m512i qo = _mm512_mul_epi32(_mm512_mul_epu32(po, vmu), vp);
    const __m512i de = _mm512_sub_epi64(pe, qe), dn = _mm512_sub_epi64(po, qo);
    return _mm512_mask_blend_epi32((__mmask16)0xAAAA, _mm512_srli_epi64(de, 32), dn);
}
// (s + rc)^7 per lane, s and rc canonical: x = s + rc - p in [-p, p), three signed products, one fix-up
PB_AVX512 __m512i vsbox_rc(__m512i s, __m512i rc) {
    const __m512i vp = _mm512_set1_epi32((int)P);
    const __m512i x = _mm512_sub_epi32(_mm512_add_epi32(s, rc), vp);
    const __m512i x2 = vsmul(x, x), x3 = vsmul(x2, x), x4 = vsmul(x2, x2);
    const __m512i t = vsmul(x3, x4);
    return _mm512_min_epu32(t, _mm512_add_epi32(t, vp));
}
PB_AVX512 __m512i vexternal(__m512i x) {
    const __m512i r1 = _mm512_shuffle_epi32(x, (_MM_PERM_ENUM)0x39);            // lane i <- x_(i+1) of its group
    const __m512i t = vadd(x, r1);
    const __m512i S = vadd(t, _mm512_shuffle_epi32(t, (_MM_PERM_ENUM)0x4E));    // the group's sum in every lane
    const __m512i y = vadd(vadd(S, x), vadd(r1, r1));
    const __m512i a = vadd(y, _mm512_shuffle_i32x4(y, y, 0x4E));                // groups (0+2, 1+3, 2+0, 3+1)
    const __m512i q = vadd(a, _mm512_shuffle_i32x4(a, a, 0xB1));                // column sums in every group
    return vadd(y, q);
}

__attribute__((target("avx512f,avx512dq"))) void permute_avx512(uint32_t s[16], const P2Host& k) {
    __m512i v = _mm512_loadu_si512((const void*)s);
    v = vexternal(v);
    for (int r = 0; r < 4; r++) {
        v = vsbox_rc(v, _mm512_loadu_si512((const void*)k.rc_ext[r]));
        v = vexternal(v);
    }
    const __m512i diag = _mm512_loadu_si512((const void*)k.diag), lo32 = _mm512_set1_epi64(0xffffffffll);
    uint32_t s0 = (uint32_t)_mm_cvtsi128_si32(_mm512_castsi512_si128(v));
    const uint32_t d0 = k.diag[0];
    const bool d0_minus2 = d0 == 0x58000005u;                    // Montgomery form of p - 2
    for (int r = 0; r < 13; r++) {
        s0 = sbox_rc(s0, k.rc_int[r]);
        // off the serial chain: sum of lanes 1..15 (64-bit lanes cannot overflow: 15 p < 2^35) and the diagonal product
        const __m512i vz = _mm512_maskz_mov_epi32((__mmask16)0xFFFE, v);
        const uint64_t tot = (uint64_t)_mm512_reduce_add_epi64(_mm512_add_epi64(_mm512_and_si512(vz, lo32), _mm512_srli_epi64(vz, 32)));
        const __m512i dv = vmul(v, diag);
        const uint32_t part = (uint32_t)(tot % P);
        const uint32_t sum = add(part, s0);
        v = vadd(dv, _mm512_set1_epi32((int)sum));              // lane 0 is not used until it is set below
        // lane 0: sum + d0 s0.  For the Plonky3 diagonal d0 = -2 and sum = part + s0, so it is just part - s0 (one subtraction on
        // the serial chain instead of a product)
        if (d0_minus2) { const uint32_t d = part - s0, e = d + P; s0 = e < d ? e : d; }
        else s0 = add(sum, mul(s0, d0));
    }
    v = _mm512_mask_set1_epi32(v, (__mmask16)0x0001, (int)s0);
    for (int r = 4; r < 8; r++) {
        v = vsbox_rc(v, _mm512_loadu_si512((const void*)k.rc_ext[r]));
        v = vexternal(v);
    }
    _mm512_storeu_si512((void*)s, v);
}

bool detect() {
    if (getenv("PB_HOST_P2_SCALAR")) return false;
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq");
}
const bool g_avx512 = detect();
}  // namespace

int uses_avx512() { return g_avx512 ? 1 : 0; }
void permute(uint32_t s[16], const P2Host& k) {
    if (g_avx512) permute_avx512(s, k);
    else permute_scalar(s, k);
}
#else
int uses_avx512() { return 0; }
void permute(uint32_t s[16], const P2Host& k) { permute_scalar(s, k); }
#endif

}  // namespace pbhost
