"""A second, independent restatement of Poseidon2 (width 16, x^7, 4+13+4 rounds) in pure Python big-integer arithmetic, written
from the Poseidon2 paper's round structure and driven only by constants/poseidon2_babybear_w16.json -- it shares no code with
oracle/poseidon2.c.  tests/test_oracle.py checks the C oracle against it (permutation, sponge, compression, a Merkle root)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = json.load(open(os.path.join(ROOT, "constants", "poseidon2_babybear_w16.json")))
P = K["p"]


def external(s):
    """circ(2 M4, M4, M4, M4): every 4-chunk is multiplied by M4, then the sum of the four chunks is added to each"""
    m4 = K["m4"]
    chunks = [[sum(m4[i][j] * s[4 * c + j] for j in range(4)) % P for i in range(4)] for c in range(4)]
    tot = [sum(chunks[c][i] for c in range(4)) % P for i in range(4)]
    return [(chunks[c][i] + tot[i]) % P for c in range(4) for i in range(4)]


def internal(s):
    """(J + diag(V)): y_i = V_i x_i + sum(x)"""
    tot = sum(s) % P
    return [(K["internal_diag_m1"][i] * s[i] + tot) % P for i in range(16)]


def permute(state):
    s = [int(x) % P for x in state]
    assert len(s) == 16
    s = external(s)
    for rc in K["external_initial"]:
        s = external([pow((x + c) % P, 7, P) for x, c in zip(s, rc)])
    for rc in K["internal"]:
        s[0] = pow((s[0] + rc) % P, 7, P)
        s = internal(s)
    for rc in K["external_terminal"]:
        s = external([pow((x + c) % P, 7, P) for x, c in zip(s, rc)])
    return s


def hash_row(row):
    """PaddingFreeSponge<16, 8, 8>: overwrite-mode absorption of 8 elements at a time, permute after every (partial) block"""
    s = [0] * 16
    row = [int(x) for x in row]
    for i in range(0, len(row), 8):
        blk = row[i:i + 8]
        s[:len(blk)] = blk
        s = permute(s)
    return s[:8]


def compress(left, right):
    """TruncatedPermutation<16 -> 8> on left || right"""
    return permute([int(x) for x in left] + [int(x) for x in right])[:8]


def merkle_root(rows):
    level = [hash_row(r) for r in rows]
    while len(level) > 1:
        level = [compress(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
    return level[0]


class DuplexChallenger:
    """DuplexChallenger<F, Perm16, WIDTH 16, RATE 8> restated from its published behaviour: observing clears pending outputs and
    buffers the value, a full input buffer is absorbed (overwrite) and permuted; sampling absorbs any pending input first and
    pops from the END of the 8-element output buffer."""

    def __init__(self):
        self.state, self.inp, self.out = [0] * 16, [], []

    def _duplex(self):
        self.state[:len(self.inp)] = self.inp
        self.inp = []
        self.state = permute(self.state)
        self.out = self.state[:8]

    def observe(self, vals):
        for v in vals:
            self.out = []
            self.inp.append(int(v))
            if len(self.inp) == 8:
                self._duplex()

    def sample(self):
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()
