"""N > 1 path on CPU: world_size-2 (and 3) `gloo` runs of the slab protocol
(irl-maxent_b200/slab.py) with the test-only numpy backend, against the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, size, chunk, out_dir):
    for p in (ROOT, os.path.join(ROOT, "irl-maxent_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        import slab
        from slab_cpu_backend import NumpyBackend
        S = size * size
        g = slab.SlabGrid(size, 0.2, icy=True, backend=NumpyBackend(), chunk=chunk)
        r = np.full(S, -0.1); r[S - 1] = 1.0
        phi = np.full(S, -np.inf); phi[S - 1] = 0.0
        p0 = np.zeros(S); p0[0] = 1.0
        pol, v = g.soft_vi(g.local(r), g.local(phi), 0.9, 1e-5)
        n_lap, st_lap = g.last_n_iter, g.last_status
        d = g.svf(g.local(p0), [S - 1], pol, 1e-5)
        n_svf = g.last_n_iter
        val = g.value_iteration(g.local(r), 0.95, 1e-5)
        n_vi = g.last_n_iter
        d_cap = g.svf(g.local(p0), [S - 1], pol, 1e-5, max_sweeps=37)
        n_cap, st_cap = g.last_n_iter, g.last_status
        full = dict(pol=g.gather(pol).numpy(), d=g.gather(d).numpy(), val=g.gather(val).numpy(),
                    d_cap=g.gather(d_cap).numpy())
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), n_lap=n_lap, st_lap=st_lap, n_svf=n_svf, n_vi=n_vi,
                     n_cap=n_cap, st_cap=st_cap, exchanges=g.n_exchanges, **full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,size,chunk", [(2, 8, 32), (2, 7, 5), (3, 9, 16)])
def test_slab_protocol_matches_single_process_oracle(tmp_path, world, size, chunk):
    from oracle import dense_port as D
    port = _free_port()
    mp.spawn(_worker, args=(world, port, size, chunk, str(tmp_path)), nprocs=world, join=True)
    out = np.load(os.path.join(str(tmp_path), "out.npz"))
    S = size * size
    P = D.icy_gridworld_table(size, 0.2)
    r = np.full(S, -0.1); r[S - 1] = 1.0
    p0 = np.zeros(S); p0[0] = 1.0
    pol, n_lap = D.local_causal_action_probabilities(P, [S - 1], r, 0.9, 1e-5)
    d, n_svf = D.expected_svf_from_policy(P, p0, [S - 1], pol, 1e-5)
    val, n_vi = D.value_iteration(P, r, 0.95, 1e-5)
    d_cap, _ = D.expected_svf_from_policy(P, p0, [S - 1], pol, 1e-5, max_sweeps=37)
    assert int(out["n_lap"]) == n_lap and int(out["st_lap"]) == 0
    assert int(out["n_svf"]) == n_svf and int(out["n_vi"]) == n_vi
    assert int(out["n_cap"]) == 37 and int(out["st_cap"]) == 2
    np.testing.assert_allclose(out["pol"], pol, rtol=1e-11)
    np.testing.assert_allclose(out["d"], d, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(out["val"], val, rtol=1e-11)
    np.testing.assert_allclose(out["d_cap"], d_cap, rtol=1e-11, atol=1e-300)


def test_row_partition():
    for p in (os.path.join(ROOT, "irl-maxent_b200"),):
        if p not in sys.path:
            sys.path.insert(0, p)
    import slab
    assert slab.row_partition(8, 2) == [(0, 4), (4, 8)]
    assert slab.row_partition(9, 4) == [(0, 3), (3, 5), (5, 7), (7, 9)]
    for n, w in ((2048, 8), (7, 3), (5, 5)):
        parts = slab.row_partition(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


# ------------------------------------------------------------ batch sharding ---

def _shard_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "irl-maxent_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        import sharding
        B = 11
        b0, b1 = sharding.shard_range(B, rank, world)
        local = torch.arange(b0, b1, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)
        full = sharding.gather_rows(local, B)
        if rank == 1:
            np.save(os.path.join(out_dir, "g.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


def test_batch_sharding_gloo(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
    import sharding
    for B, w in ((4096, 8), (11, 2), (5, 8)):
        r = [sharding.shard_range(B, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == B and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    mp.spawn(_shard_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g = np.load(os.path.join(str(tmp_path), "g.npy"))
    assert g.shape == (11, 3) and np.array_equal(g[:, 0], np.arange(11.0))
