"""CPU tests of the sampler restatement (oracle/sampler_port.py): the Philox4x32-10 block function
against the known-answer vectors of the Random123 distribution (kat_vectors, `philox4x32 10`), and
the rollout statistics against the reference-style host sampler of trajectory.py."""
import numpy as np

from oracle import dense_port as D
from oracle import sampler_port as SPL

import gridworld as W
import solver as S
import trajectory as T


def test_philox4x32_10_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert SPL.philox4x32_10(ctr, key) == out


def test_u53_range():
    assert SPL.u53(0, 0) == 0.0
    assert SPL.u53(0xffffffff, 0xffffffff) == 1.0 - 2.0 ** -53


def test_oracle_sampler_agrees_statistically_with_the_reference_style_sampler():
    """Same world, policy, start and terminal as main.py (reference: main.py:32-51): visit frequencies
    and mean length of 3 000 Philox rollouts vs 3 000 numpy rollouts agree within sampling error."""
    world = W.IcyGridWorld(size=5, p_slip=0.2)
    reward = np.zeros(25); reward[24] = 1.0; reward[8] = 0.65
    value, _ = D.value_iteration(world.p_transition, reward, 0.7)      # the oracle: no GPU in this test
    policy = S.stochastic_policy_from_value(world, value, w=lambda x: x ** 5)
    n = 3000
    np.random.seed(0)
    ref = list(T.generate_trajectories(n, world, T.stochastic_policy_adapter(policy), 0, [24]))
    len_ref = np.array([len(t.transitions()) for t in ref], dtype=float)
    vis_ref = np.zeros(25)
    for t in ref:
        for s in t.states():
            vis_ref[s] += 1
    cdf = np.cumsum(np.eye(25)[0])
    lens, vis = [], np.zeros(25)
    for i in range(n):
        st, ac, trunc = SPL.sample_trajectory(world.successors, policy, cdf, [24], i, 12345, 10000)
        assert not trunc and st[0] == 0 and st[-1] == 24 and len(ac) == len(st) - 1
        lens.append(len(ac))
        for s in st:
            vis[s] += 1
    lens = np.array(lens, dtype=float)
    se = np.sqrt(lens.var() / n + len_ref.var() / n)
    assert abs(lens.mean() - len_ref.mean()) < 5 * se
    f, f_ref = vis / vis.sum(), vis_ref / vis_ref.sum()
    assert np.max(np.abs(f - f_ref)) < 0.01
