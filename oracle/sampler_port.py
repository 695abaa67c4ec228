"""
CPU restatement of the device trajectory sampler (irl-maxent_b200/csrc/trajectories.cu).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The rollout loop is the reference's (`/root/reference/src/trajectory.py:52-128`: start state from
the start distribution, `while state not in final`, action from the policy row, successor from the
transition row, both by numpy's inverse-cdf rule `cdf.searchsorted(u, side='right')`).  What is NOT
the reference's is the random stream: the reference consumes numpy's global Mersenne Twister one
draw at a time, which cannot be reproduced by independent device threads, so the device sampler uses
Philox4x32-10 (Salmon et al., SC'11; 10 rounds, multipliers 0xD2511F53 / 0xCD9E8D57, Weyl constants
0x9E3779B9 / 0xBB67AE85) keyed by the seed with the counter (step, 0, trajectory, 0).  This file
restates that generator and the selection arithmetic so that device trajectories can be checked
bit for bit; the statistical agreement with the reference's sampler is a separate test.
Parity: the Philox block function is pinned to the known-answer vectors of the Random123
distribution (tests/test_sampler_oracle.py).
"""

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def u53(a, b):
    """numpy's random_sample recipe: 27 + 26 bits -> [0, 1)."""
    return ((a >> 5) * 67108864.0 + (b >> 6)) * (1.0 / 9007199254740992.0)


def sample_trajectory(succ, policy, start_cdf, terminal, i, seed, max_len):
    """Trajectory i.  `succ(s, a)` -> (states ascending, probabilities) of the non-zero entries of
    p_transition[s, :, a] PADDED the way the ELL row is (zeros allowed).  Returns (states, actions,
    truncated)."""
    key = (seed & MASK, (seed >> 32) & MASK)
    S = len(start_cdf)
    w = philox4x32_10((MASK, MASK, i, 0), key)
    u0 = u53(w[0], w[1]) * start_cdf[S - 1]
    s = int(np.searchsorted(start_cdf, u0, side='right'))
    s = min(s, S - 1)
    states, actions = [s], []
    term = set(int(t) for t in terminal)
    while s not in term:
        if len(actions) >= max_len:
            return states, actions, True
        w = philox4x32_10((len(actions), 0, i, 0), key)
        row = policy[s]
        tot = 0.0
        for a in range(len(row)):
            tot += row[a]
        ua = u53(w[0], w[1]) * tot
        act, cum = len(row) - 1, 0.0
        for a in range(len(row)):
            cum += row[a]
            if cum > ua:
                act = a
                break
        idx, p = succ(s, act)
        ptot = 0.0
        for pj in p:
            ptot += pj
        us = u53(w[2], w[3]) * ptot
        nxt, cum = s, 0.0
        for j in range(len(p)):
            cum += p[j]
            if p[j] > 0.0:
                nxt = int(idx[j])
            if cum > us:
                break
        actions.append(act)
        s = nxt
        states.append(s)
    return states, actions, False
