// built and run by tests/test_host.py: the AVX-512 host permutation against the scalar one with a NON-default internal diagonal (the
// generic lane-0 branch; the default diagonal takes the d0 = -2 shortcut and is covered through pb_host_poseidon2_permute)
#include "../powdr_b200/csrc/transcript_host.h"
#include "../include/pb_poseidon2_constants.h"
#include <stdio.h>
#include <string.h>
static const uint32_t P = 2013265921u;
static uint32_t to_m(uint32_t c) { return (uint32_t)(((uint64_t)c << 32) % P); }
int main() {
    pbhost::P2Host k;
    for (int r = 0; r < 8; r++) for (int i = 0; i < 16; i++) k.rc_ext[r][i] = to_m(PB_P2_RC_EXT[r][i]);
    for (int r = 0; r < 13; r++) k.rc_int[r] = to_m(PB_P2_RC_INT[r]);
    for (int i = 0; i < 16; i++) k.diag[i] = to_m((PB_P2_DIAG_M1[i] * 7ull + 3) % P);      // a different diagonal: generic lane-0 branch
    uint32_t a[16], b[16];
    for (int i = 0; i < 16; i++) a[i] = b[i] = to_m(i * 1234567u % P);
    int bad = 0;
    for (int it = 0; it < 50000; it++) { pbhost::permute(a, k); pbhost::permute_scalar(b, k); if (memcmp(a, b, 64)) bad++; a[it & 15] = b[it & 15] = (uint32_t)((a[it & 15] * 2654435761ull) % P); }
    printf("generic diag: avx512=%d bad=%d\n", pbhost::uses_avx512(), bad);
    return bad != 0;
}
