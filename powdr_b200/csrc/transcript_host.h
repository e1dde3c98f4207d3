// Host side of the Fiat-Shamir transcript: the Poseidon2 permutation of the duplex challenger, on Montgomery words.
// The sponge over the opened values is inherently serial (keccak shape: 4363 dependent permutations between zeta and gamma; a
// 50-chip segment: 41 k), so it runs on the CPU while the GPU waits -- one AVX-512 register holds the whole width-16 state
// (transcript_host.cpp), with a portable scalar version behind the same call.
#pragma once
#include <stdint.h>

namespace pbhost {

struct P2Host {   // Montgomery constants of the permutation
    uint32_t rc_ext[8][16], rc_int[13], diag[16];
};

// in place; every word in [0, p), Montgomery form
void permute(uint32_t s[16], const P2Host& k);
void permute_scalar(uint32_t s[16], const P2Host& k);
// 1 when the AVX-512 version is in use (CPU support and PB_HOST_P2_SCALAR unset)
int uses_avx512();

}  // namespace pbhost
