"""Worker of tests/test_gpu_multi.py (one process per GPU under torch.distributed.run): the slab kernels on
N ranks against the plain-C oracle (1e-10, identical sweep counts) and, bitwise, against the 1-GPU kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "irl-maxent_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import _irlb200 as E
import slab
from oracle import c_port as C
from oracle import sparse_port as SP


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    # 50 rows do not divide evenly over 4 or 8 ranks (slabs of different height), 96 x 96 is sized for edge CTAs
    for n, flow, chunk in ((64, True, 0), (96, True, 7), (50, True, 5), (64, False, 0)):
        S = n * n
        rng = np.random.default_rng(n)
        r = -0.1 + 0.02 * rng.standard_normal(S); r[S - 1] = 1.0
        phi = np.full(S, -np.inf); phi[S - 1] = 0.0
        p0 = np.zeros(S); p0[0] = 0.5; p0[S // 2 + 1] = 0.5
        g = slab.PeerSlabGrid(n, 0.2, flow=flow, chunk_sweeps=chunk)
        pol, v = g.soft_vi(g.local(r), g.local(phi), 0.9, 1e-5)
        n_lap = g.last_n_iter
        d = g.svf(g.local(p0), [S - 1], pol, 1e-5, max_sweeps=4000)
        n_fw, st_fw = g.last_n_iter, g.last_status
        val = g.value_iteration(g.local(r), 0.95, 1e-4)
        n_vi = g.last_n_iter
        pol_full, d_full, val_full = (g.gather(x).cpu().numpy() for x in (pol, d, val))
        # 1-GPU kernels, bitwise (every rank checks its own slab)
        t1 = E.gridworld_tables(n, 0.2, slots=4)
        pol1 = E.soft_vi(t1, E.terminal_phi([S - 1], S), r, 0.9, mode=E.MODE_GRID)
        assert int(E.last_info.n_iter.item()) == n_lap
        d1 = E.svf(t1, p0, E.terminal_mask([S - 1], S), pol1, 1e-5, max_sweeps=4000, mode=E.MODE_GRID)
        assert int(E.last_info.n_iter.item()) == n_fw
        assert bool((pol == pol1[0, g.lo:g.hi]).all()) and bool((d == d1[0, g.lo:g.hi]).all()), "slab != 1-GPU kernel"
        if rank == 0:
            sidx, sp = C.ell_from_sparse(SP.icy_gridworld_sparse(n, 0.2))
            pol_c, val_c, n_c = C.soft_vi(sidx, sp, phi, r, 0.9, 1e-5)
            assert n_c == n_lap, (n_c, n_lap)
            np.testing.assert_allclose(pol_full, pol_c, rtol=1e-10, atol=1e-300)
            d_c, n_dc = C.svf(sidx, sp, p0, [S - 1], pol_c, 1e-5, max_sweeps=4000)
            assert n_dc == n_fw, (n_dc, n_fw)
            np.testing.assert_allclose(d_full, d_c, rtol=1e-10, atol=1e-300)
            v_c, n_vc = C.value_iteration(sidx, sp, r, 0.95, 1e-4)
            assert n_vc == n_vi
            np.testing.assert_allclose(val_full, v_c, rtol=1e-10, atol=1e-300)
            print("slab parity ok: n=%d ranks=%d flow=%s soft-VI %d forward %d (status %d) VI %d sweeps"
                  % (n, world, flow, n_lap, n_fw, st_fw, n_vi), flush=True)
        g.close()
    # C5 at FULL size (2048 x 2048, 4.2 M states): the N-rank dataflow kernel must be bitwise the 1-GPU
    # cooperative-grid kernel, which tests/test_gpu_parity.py::test_c5_full_size_sweeps_against_sparse_oracle
    # holds to 1e-10 of the oracle at this size (fixed budgets: convergence takes 10^3 / 10^7 sweeps)
    n = 2048
    S = n * n
    g = slab.PeerSlabGrid(n, 0.2, flow=True)
    rng = np.random.default_rng(5)
    r = -0.1 + 0.05 * rng.standard_normal(S); r[S - 1] = 1.0
    phi = np.full(S, -np.inf); phi[S - 1] = 0.0
    p0 = np.zeros(S); p0[0] = 0.5; p0[S // 2 + n // 2] = 0.5
    k = 40
    pol, v = g.soft_vi(g.local(r), g.local(phi), 0.9, 1e-5, max_sweeps=k)
    assert g.last_n_iter == k and g.last_status == slab.ST_MAXSWEEPS
    uniform = torch.full((g.cnt, 4), 0.25, dtype=torch.float64, device="cuda")
    d = g.svf(g.local(p0), [S - 1], uniform, 1e-5, max_sweeps=k)
    assert g.last_n_iter == k
    t1 = E.gridworld_tables(n, 0.2, slots=4)
    _, v1 = E.soft_vi(t1, E.terminal_phi([S - 1], S), r, 0.9, max_sweeps=k, mode=E.MODE_GRID, want_value=True)
    d1 = E.svf(t1, p0, E.terminal_mask([S - 1], S), torch.full((S, 4), 0.25, dtype=torch.float64, device="cuda"),
               1e-5, max_sweeps=k, mode=E.MODE_GRID)
    assert bool((v == v1[0, g.lo:g.hi]).all()) and bool((d == d1[0, g.lo:g.hi]).all()), "C5 full size: slab != 1-GPU kernel"
    g.close()
    if rank == 0:
        print("slab parity ok: 2048x2048 over %d ranks, %d sweeps of soft-VI and of the forward pass bitwise the 1-GPU kernel"
              % (world, k), flush=True)
    dist.barrier()
    if rank == 0:
        print("SLAB_MULTI_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
