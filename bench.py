#!/usr/bin/env python
"""
bench.py -- IRL grad-steps/sec of the MaxEnt gradient-step body on B200.

    python bench.py --gpus N --steps K --warmup W            (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    (the reference's CPU arithmetic, host cores)

Workload (BASELINE.json configs[3], SURVEY section 8d "C4"): per GPU a batch of
B = 4096 independent 32x32 IcyGridWorlds (S = 1024 states, A = 4,
p_slip_b = 0.1 + 0.2 b / B), identity features, terminal = [S-1], start state 0.
One step = one gradient-step body of `maxent.irl` (maxent.py:240-252) for every
world of the batch: reward = omega -> backward pass (2 S partition sweeps) ->
forward state-visitation pass to eps = 1e-5 -> grad = e_features - svf ->
Sga step on the device-resident omega.  Rewards start at -ln 4 + 0.01 N(0,1), the
regime in which the reference's raw backward pass is finite (SURVEY TL;DR 2), so
the CPU arm computes the same thing.  Scaling is weak: every rank owns its own
4096 worlds, no collective on the data path; value = all ranks' grad-steps / max
over ranks of the device time.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "irl-maxent_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "irl_grad_steps_per_sec"
UNIT = "grad-steps/s"
SVF_BYTES_PER_STATE_SWEEP = 84      # 20 + 8 w, w = 8: SURVEY section 8(d), merged predecessor weights, Kp = 5
BWD_BYTES_PER_STATE_SWEEP = 216     # 64 + 19 w


def ncu_constants():
    """Figures read off ncu captures, kept in a tracked file next to the summaries they come from."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_constants.json")))
    except Exception:
        return {}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="worlds per GPU")
    ap.add_argument("--size", type=int, default=32, help="grid side length")
    ap.add_argument("--lr", type=float, default=1e-5)
    ap.add_argument("--max-sweeps", type=int, default=2000000,
                    help="guard per fixed point (the reference has none); worlds that hit it are reported")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C1/C2/C3 side measurements")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds for the whole reference arm")
    return ap.parse_args()


def workload(args, rank):
    """Synthetic inputs of one rank (host arrays)."""
    B, n = args.batch, args.size
    S = n * n
    gb = rank * B + np.arange(B)                       # global world ids: weak scaling, distinct worlds per rank
    p_slip = 0.1 + 0.2 * (gb % 4096) / 4096.0
    rng = np.random.default_rng(1000 + rank)
    theta0 = -np.log(4.0) + 0.01 * rng.standard_normal((B, S))
    r_true = -np.log(4.0) + 0.01 * np.random.default_rng(2000 + rank).standard_normal((B, S))
    p0 = np.zeros(S)
    p0[0] = 1.0
    return dict(B=B, n=n, S=S, p_slip=p_slip, theta0=theta0, r_true=r_true, p0=p0, terminal=[S - 1])


def config_dict(args, n_gpus):
    return {"workload": "C4: batched independent %dx%d IcyGridWorlds, MaxEnt gradient-step body "
                        "(2S backward sweeps + SVF to eps 1e-5 + Sga step)" % (args.size, args.size),
            "worlds_per_gpu": args.batch, "states": args.size ** 2, "actions": 4,
            "eps_svf": 1e-5, "optimizer": "Sga(lr=%g)" % args.lr, "features": "identity",
            "parallelism": "batch-sharded x%d, no data-path collective" % n_gpus,
            "l2": "inputs larger than L2: %.2f GB of tables + 0.1 GB of vectors streamed per step"
                  % (args.batch * args.size ** 2 * 360 / 1e9)}


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_mark = index, [], None, 0.0

    def mark(self):
        """Start of the timed region: only samples that arrive after this count."""
        self.t_mark = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, smax, reasons = [], None, set()
        for stamp, r in self.rows:
            if stamp < self.t_mark:
                continue
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arm: the reference's dense numpy arithmetic (oracle/dense_port.py)
# ---------------------------------------------------------------------------

def blas_threads(set_to=None):
    """BLAS threads numpy's dgemv really uses.  `set_to`: raise the pool to that many threads first --
    torch.distributed.run exports OMP_NUM_THREADS=1 to every rank, which would silently make the CPU
    arm single-threaded at N >= 2."""
    try:
        import threadpoolctl
        if set_to:
            threadpoolctl.threadpool_limits(limits=int(set_to), user_api="blas")
        return max([i.get("num_threads", 1) for i in threadpoolctl.threadpool_info()
                    if i.get("user_api") == "blas"] + [1])
    except Exception:
        return 1


def cpu_full_body(w, b):
    """One complete gradient-step body of world b with the dense port (kind: "port").
    Table construction is excluded (the reference builds it once, gridworld.py:52)."""
    from oracle import dense_port as D
    P = D.icy_gridworld_table(w["n"], w["p_slip"][b])
    ef = np.zeros(w["S"])
    t0 = time.perf_counter()
    d, n_svf = D.compute_expected_svf(P, w["p0"], w["terminal"], w["theta0"][b], 1e-5)
    grad = ef - d
    _ = w["theta0"][b] + 1e-3 * grad
    dt = time.perf_counter() - t0
    return dt, n_svf


def cpu_sparse_body(w, b):
    """The same body with the scipy-CSR restatement (oracle/sparse_port.py): NOT the reference's cost
    profile -- the "fair sparse CPU" line of BASELINE.md section 3, one thread."""
    from oracle import sparse_port as SP
    mdp = SP.icy_gridworld_sparse(w["n"], w["p_slip"][b])
    mdp.transposed()
    t0 = time.perf_counter()
    pa = SP.local_action_probabilities(mdp, w["terminal"], w["theta0"][b], rescale=False)
    d, n_svf = SP.expected_svf_from_policy(mdp, w["p0"], w["terminal"], pa, 1e-5)
    _ = w["theta0"][b] + 1e-3 * (0.0 - d)
    return time.perf_counter() - t0, n_svf


def cpu_c_openmp(w, n_worlds=None):
    """The same body for a slice of the batch with the plain-C restatement (oracle/c/irl_oracle.c), one
    world per OpenMP thread on all host cores: the "optimised CPU" line (not the reference's arithmetic)."""
    from oracle import c_port as C
    from oracle import sparse_port as SP
    threads = os.cpu_count() or 1
    nw = n_worlds or 2 * threads
    tabs = [C.ell_from_sparse(SP.icy_gridworld_sparse(w["n"], w["p_slip"][b])) for b in range(nw)]
    sidx = np.stack([t[0] for t in tabs]); sp = np.stack([t[1] for t in tabs])
    t0 = time.perf_counter()
    out, n_svf = C.batch_maxent_step(sidx, sp, w["terminal"], w["p0"], w["theta0"][:nw], 1e-5)
    dt = time.perf_counter() - t0
    return nw / dt, threads, nw, dt, float(n_svf.mean())


def cpu_sampled_body(w, b, frac):
    """Bounded sample of world b's gradient-step body with the reference's dense arithmetic
    (oracle/dense_port.py restates maxent.py line by line; these are the same statements with the loops
    cut short): the per-call dense slicing / copy (maxent.py:98-102,143) in full, then the first
    `frac` of the 2S backward sweeps and the first `frac` of this world's forward sweeps.  The world's
    exact forward sweep count comes from the sparse restatement (not timed).
    Returns (measured seconds, seconds extrapolated to the whole body, detail)."""
    from oracle import dense_port as D
    from oracle import sparse_port as SP
    S, A = w["S"], 4
    P = D.icy_gridworld_table(w["n"], w["p_slip"][b])
    r = w["theta0"][b]
    mdp = SP.icy_gridworld_sparse(w["n"], w["p_slip"][b])
    pa = SP.local_action_probabilities(mdp, w["terminal"], r)
    _, n_svf = SP.expected_svf_from_policy(mdp, w["p0"], w["terminal"], pa, 1e-5)
    n_bw, n_fw = max(1, int(round(frac * 2 * S))), max(1, int(round(frac * n_svf)))
    t0 = time.perf_counter()
    er = np.exp(r)
    per_action = [np.array(P[:, :, a]) for a in range(A)]              # maxent.py:143
    zs = np.zeros(S)
    zs[w["terminal"]] = 1.0
    t_setup_b = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(n_bw):
        za = np.array([er * per_action[a].dot(zs) for a in range(A)]).T    # :155
        zs = za.sum(axis=1)                                                # :156
    t_bw = time.perf_counter() - t0
    t0 = time.perf_counter()
    pt = np.copy(P)                                                    # :98
    pt[w["terminal"], :, :] = 0.0
    per_action_t = [np.array(pt[:, :, a]) for a in range(A)]           # :102
    t_setup_f = time.perf_counter() - t0
    d = np.zeros(S)
    t0 = time.perf_counter()
    for _ in range(n_fw):
        parts = [per_action_t[a].T.dot(pa[:, a] * d) for a in range(A)]    # :109
        d_new = w["p0"] + np.array(parts).sum(axis=0)                      # :110
        _delta, d = np.max(np.abs(d_new - d)), d_new                       # :112
    t_fw = time.perf_counter() - t0
    measured = t_setup_b + t_bw + t_setup_f + t_fw
    whole = t_setup_b + t_setup_f + t_bw * (2.0 * S / n_bw) + t_fw * (float(n_svf) / n_fw)
    return measured, whole, dict(setup_s=t_setup_b + t_setup_f, backward_sweeps_timed=n_bw, forward_sweeps_timed=n_fw,
                                 forward_sweeps_of_the_world=int(n_svf),
                                 backward_ms_per_sweep=1e3 * t_bw / n_bw, forward_ms_per_sweep=1e3 * t_fw / n_fw)


def run_reference_arm(args):
    """The reference's own CPU arithmetic for the same metric / config on the host cores.  Each step is a
    bounded sample of one world's gradient-step body (see cpu_sampled_body): `ms_per_step` is what was
    really timed, `value` = 1 / (that sample extrapolated to the whole body); one complete body is timed
    beside it so that the extrapolation can be checked."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = workload(args, 0)
    cores = blas_threads(set_to=os.cpu_count() or 1)
    total_steps = args.steps + args.warmup
    # ~60 % of the budget for the timed samples, the rest for the sparse sweep counts and one full body
    per_step = 0.6 * args.cpu_budget / max(1, total_steps)
    frac = float(min(1.0, max(0.02, per_step / 16.0)))       # a whole 32 x 32 body takes ~12-18 s on 16-32 cores
    measured, whole, detail = [], [], None
    for i in range(total_steps):
        b = (i * 409) % w["B"]
        m, e, detail = cpu_sampled_body(w, b, frac)
        if i == 0 and args.warmup > 0:                       # size the samples from the first (untimed) one
            frac = float(min(1.0, max(0.02, frac * per_step / max(m, 1e-3))))
        if i >= args.warmup:
            measured.append(m)
            whole.append(e)
    mean_whole = float(np.mean(whole))
    value = 1.0 / mean_whole
    full_check = None
    try:
        b = (args.warmup * 409) % w["B"]                     # the first timed step's world, complete
        t_full, n_svf = cpu_full_body(w, b)
        full_check = {"world": int(b), "seconds": t_full, "extrapolated_seconds_same_world": whole[0],
                      "forward_sweeps": int(n_svf)}
    except Exception as e:                                   # pragma: no cover
        full_check = {"error": repr(e)}
    sample = ("one world of the batch per step: per-call dense slicing in full + the first %.0f %% of the 2S backward "
              "and of the world's forward dense sweeps timed (ms_per_step), value = 1 / (sample extrapolated to the "
              "whole body by the exact sweep counts); one complete body timed beside it (full_body_check)" % (100 * frac))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(measured)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "sample_fraction": frac, "extrapolated_body_seconds": mean_whole, "detail": detail,
                             "full_body_check": full_check,
                             "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"), "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# side measurements of the other BASELINE configs (not the headline; N = 1 only)
# ---------------------------------------------------------------------------

def other_configs():
    """C1 / C2: the seeded 5x5 `irl` / `irl_causal` runs of main.py end to end through the public API
    (fixture trajectories, numpy in / numpy out); C3: one causal gradient-step body of a single 128x128
    world; C4 causal: the headline batch with the MaxCausalEnt body.  Wall-clock, synchronize on both sides."""
    import torch
    import _irlb200 as E
    import gridworld as W
    import maxent as M
    import optimizer as O
    import trajectory as T
    out = {}
    g = np.load(os.path.join(ROOT, "tests", "golden", "e2e_5x5.npz"))
    # the 200 seeded expert trajectories of main.py (fixture), as product Trajectory objects
    tjs, off = [], 0
    for length in g["traj_len"]:
        tjs.append(T.Trajectory([tuple(int(v) for v in row) for row in g["traj_flat"][off:off + length]]))
        off += length
    world = W.IcyGridWorld(5, 0.2)
    F = W.state_features(world)

    runs = (("C1_irl_5x5", lambda o: M.irl(world.p_transition, F, [24], tjs, o, O.Constant(1.0)), "irl_steps"),
            ("C2_irl_causal_5x5_g0.9",
             lambda o: M.irl_causal(world.p_transition, F, [24], tjs, o, O.Constant(1.0), 0.9), "irl_causal_0.9_steps"))
    for name, fn, key in runs:
        best = None
        for _ in range(3):
            o = O.ExpSga(lr=O.linear_decay(lr0=0.2))       # plain built-in optimizer: the outer loop runs on the device
            torch.cuda.synchronize()
            t = time.perf_counter()
            fn(o)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        out[name] = {"outer_steps": o.k, "reference_outer_steps": int(g[key]), "seconds": best,
                     "grad_steps_per_s": o.k / best,
                     "note": "end to end through maxent.irl / irl_causal (numpy in, numpy out); outer loop inside "
                             "irlb200_irl_small, one launch per 1024 steps"}
    n = 128
    S = n * n
    tabs = E.gridworld_tables(n, 0.2)
    p0 = np.zeros(S); p0[0] = 1.0
    r = np.full(S, -0.1); r[S - 1] = 1.0
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    rd, p0d = E.to_device(r), E.to_device(p0)
    best = None
    for _ in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        pol = E.soft_vi(tabs, phi, rd, 0.9, mode=E.MODE_GRID)
        n_lap = E.last_info.n_iter
        d = E.svf(tabs, p0d, mask, pol, 1e-5, mode=E.MODE_AUTO)        # -> thread-block-cluster kernel
        n_svf = E.last_info.n_iter
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    n_lap, n_svf = int(n_lap.item()), int(n_svf.item())
    out["C3_causal_step_128x128"] = {
        "soft_vi_sweeps": n_lap, "svf_sweeps": n_svf, "seconds": best, "grad_steps_per_s": 1.0 / best,
        "us_per_sweep": 1e6 * best / (n_lap + n_svf),
        "algorithmic_GBps": (n_lap * BWD_BYTES_PER_STATE_SWEEP + n_svf * SVF_BYTES_PER_STATE_SWEEP) * S / best / 1e9,
        "note": "sync-latency bound: 3.5 MB of tables are register resident; soft-VI in cooperative-grid mode (one "
                "grid barrier per sweep), forward pass in thread-block-cluster mode (8 CTAs; boundary rows and "
                "stop-rule votes pushed by st.async into the peers' shared memory, one mbarrier wait per sweep)"}
    # the same world, non-causal body (2S partition sweeps + forward pass), reward near -ln 4
    rm = E.to_device(-np.log(4.0) + 0.01 * np.random.default_rng(0).standard_normal(S))
    best = None
    for _ in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        pol = E.backward(tabs, mask, rm)                                   # AUTO -> cluster kernel
        torch.cuda.synchronize()
        t_bw = time.perf_counter() - t
        d = E.svf(tabs, p0d, mask, pol, 1e-5, mode=E.MODE_AUTO)
        n_svf = E.last_info.n_iter
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if best is None or dt < best[0]:
            best = (dt, t_bw)
    n_svf = int(n_svf.item())
    out["C3_maxent_step_128x128"] = {
        "backward_sweeps": 2 * S, "svf_sweeps": n_svf, "seconds": best[0], "backward_seconds": best[1],
        "grad_steps_per_s": 1.0 / best[0], "us_per_backward_sweep": 1e6 * best[1] / (2 * S),
        "us_per_forward_sweep": 1e6 * (best[0] - best[1]) / n_svf,
        "note": "both passes in thread-block-cluster mode (push exchange); the cooperative-grid backward pass "
                "took 1.7 us per sweep"}
    # C2 as a batch: 1024 independent copies of the 5x5 causal IRL problem in lockstep (irl_batch)
    try:
        Bc = 1024
        best = None
        for _ in range(2):
            torch.cuda.synchronize()
            t = time.perf_counter()
            _, st = M.irl_batch(E.gridworld_tables(5, 0.2), [24], np.tile(g["e_features"], (Bc, 1)), g["p_initial"],
                                O.ExpSga(lr=O.linear_decay(lr0=0.2)), O.Constant(1.0), causal=True, discount=0.9)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        out["C2_batched_x1024"] = {"candidates": Bc, "outer_steps_each": int(st[0]), "seconds": best,
                                   "grad_steps_per_s": float(st.sum()) / best}
    except Exception as e:
        out["C2_batched_x1024"] = {"error": repr(e)}
    out["C4_causal_batch"] = causal_batch()
    out["roofline_stream"] = stream_roofline()
    try:
        out["dense_batch"] = dense_batch_line()
    except Exception as e:
        out["dense_batch"] = {"error": repr(e)}
    return out


def dense_batch_line(S=1024, A=4, B=4096, sweeps=24, B_gather=256):
    """BASELINE configs[3], dense case: B reward candidates over ONE dense random table (K = S successors per
    state), soft-VI sweeps as FP64 tensor-core contractions [A S x S] . [S x B] (csrc/dense_batch.cu) against the
    run-time-K ELL gather on the same table (timed on B_gather candidates).  Fixed sweep budget; CUDA events."""
    import torch
    import _irlb200 as E
    rng = np.random.default_rng(5)
    P = torch.as_tensor(rng.random((S, S, A))).cuda() ** 3
    P[S - 1] = 0.0
    P[S - 1, S - 1, :] = 1.0
    P /= P.sum(dim=1, keepdim=True)
    rewards = torch.as_tensor(-0.2 + 0.1 * rng.standard_normal((B, S))).cuda()
    phi = E.terminal_phi([S - 1], S)
    dt = E.DenseTables(P)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 1e3

    t_dense = timed(lambda: E.dense_soft_vi(dt, phi, rewards, 0.9, 1e-30, max_sweeps=sweeps))
    # yardstick: the library DGEMM of the same shape (no epilogue, no stop rule), sustained over as many calls
    Pa, X = dt.packed[:A * S * S].view(A * S, S), rewards.t().contiguous()
    t_blas = timed(lambda: [torch.matmul(Pa, X) for _ in range(sweeps)])
    tabs = E.compress_dense(P)
    t_gather = timed(lambda: E.soft_vi(tabs, phi, rewards[:B_gather], 0.9, 1e-30, max_sweeps=sweeps))
    flops = 2.0 * A * S * S * B * sweeps
    per_cs_dense, per_cs_gather = t_dense / (B * sweeps), t_gather / (B_gather * sweeps)
    ncu = ncu_constants().get("dense_gemm_kernel", {})
    return {"workload": "dense random MDP S=%d A=%d, %d candidates, %d soft-VI sweeps" % (S, A, B, sweeps),
            "dense_seconds": t_dense, "dense_us_per_sweep_all_candidates": 1e6 * t_dense / sweeps,
            "dense_TFLOPs_fp64": flops / t_dense / 1e12,
            "fp64_peak_note": "B200 FP64: 33.6 TFLOP/s measured with DFMA (scripts/ubench.cu), 40 nominal (vector and tensor); the "
                              "whole-loop figure includes the epilogue launches and runs at the sustained (power-capped) clock",
            "cublas_dgemm_same_shape_TFLOPs": flops / t_blas / 1e12,
            "fraction_of_cublas_dgemm": t_blas / t_dense,
            "tensor_pipe_cycles_active_pct_ncu": ncu.get("tensor_pipe_cycles_active_pct"),
            "gemm_TFLOPs_in_ncu_capture": ncu.get("TFLOPs_in_capture"), "ncu_source": ncu.get("source"),
            "ell_gather_seconds_for_%d_candidates" % B_gather: t_gather,
            "ns_per_candidate_sweep": {"dense": 1e9 * per_cs_dense, "ell_gather": 1e9 * per_cs_gather},
            "speedup_over_ell_gather": per_cs_gather / per_cs_dense}


def causal_batch(B=4096, n=32):
    """The headline batch with the MaxCausalEnt gradient-step body (soft-VI gamma = 0.9 to eps 1e-5 +
    forward pass to eps 1e-5), goal-directed reward -0.1 / +1 at the terminal plus 0.01 N(0,1)."""
    import torch
    import _irlb200 as E
    import maxent as M
    S = n * n
    tabs = E.gridworld_tables(n, 0.1 + 0.2 * np.arange(B) / B)
    r = np.full((B, S), -0.1) + 0.01 * np.random.default_rng(7).standard_normal((B, S))
    r[:, S - 1] = 1.0
    p0 = np.zeros(S); p0[0] = 1.0
    rd = E.to_device(r)
    best, info = None, None
    for _ in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        d, _ = M.compute_expected_svf_batch(tabs, p0, [S - 1], rd, causal=True, discount=0.9, fused=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if best is None or dt < best:
            best, info = dt, E.last_info.counts()
    return {"worlds": B, "states": S, "seconds": best, "grad_steps_per_s": B / best,
            "soft_vi_sweeps_mean": float(info[:, 0].mean()), "svf_sweeps_mean": float(info[:, 1].mean())}


def stream_roofline(n=2048, fw_sweeps=400, lap_sweeps=150):
    """The physically HBM-bound regime (BASELINE configs[4] on ONE GPU): a single 2048x2048 world
    (4.2 M states, 1.5 GB of tables >> 126 MB L2) in cooperative-grid mode with streamed table rows,
    fixed sweep budgets.  achieved = algorithmic bytes / CUDA-event time of the launch."""
    import torch
    import _irlb200 as E
    S = n * n
    tabs = E.gridworld_tables(n, 0.2, slots=4)             # what World.tables() builds above 128 x 128
    dev = tabs.succ_idx.device
    p0 = torch.zeros(S, dtype=torch.float64, device=dev); p0[0] = 1.0
    r = torch.full((S,), -0.1, dtype=torch.float64, device=dev); r[S - 1] = 1.0
    mask = torch.zeros(S, dtype=torch.uint8, device=dev); mask[S - 1] = 1
    phi = torch.full((S,), -float("inf"), dtype=torch.float64, device=dev); phi[S - 1] = 0.0
    peak = float(measured_peaks().get("hbm_gbs", 6650.0))
    ncu = ncu_constants()
    res = {"workload": "single %dx%d IcyGridWorld, streamed cooperative-grid kernels, fixed sweep budgets, "
                       "compact 4-slot tables" % (n, n),
           "states": S, "table_bytes": tabs.nbytes(), "peak": peak, "unit": "GB/s"}

    def timed(t):
        for rep in range(2):                               # first pass warms up (workspace allocation)
            E.launch_log = []
            pol = E.soft_vi(t, phi, r, 0.9, max_sweeps=lap_sweeps, mode=E.MODE_GRID)
            E.svf(t, p0, mask, pol, 1e-5, max_sweeps=fw_sweeps, mode=E.MODE_GRID)
            torch.cuda.synchronize()
            log, E.launch_log = E.launch_log, None
        return {name: a.elapsed_time(b) for name, a, b in log}

    ms = timed(tabs)
    K = tabs.Ks
    # bytes the kernels really move per state and sweep with K slots: forward K*(4 + 8) + 3*8 (merged weights),
    # soft-VI K*4 + A*K*8 + 3*8; `achieved` stays on SURVEY 8(d)'s algorithmic figures (84 / 216)
    moved = {"forward": 12 * K + 24, "soft_vi": 4 * K + 32 * K + 24}
    for key, name, sweeps, bps in (("forward", "svf", fw_sweeps, SVF_BYTES_PER_STATE_SWEEP),
                                   ("soft_vi", "soft_vi", lap_sweeps, BWD_BYTES_PER_STATE_SWEEP)):
        ach = bps * S * sweeps / (ms[name] / 1e3) / 1e9
        phys = moved[key] * S * sweeps / (ms[name] / 1e3) / 1e9
        res[key] = {"sweeps": sweeps, "launch_ms": ms[name], "us_per_sweep": 1e3 * ms[name] / sweeps,
                    "bound": "hbm", "moved_bytes_per_state_sweep": moved[key], "achieved": phys, "frac": phys / peak,
                    "survey_bytes_per_state_sweep": bps, "survey_equivalent_GBps": ach,
                    "note": "achieved / frac count the bytes the kernel really moves with %d-slot tables; the SURVEY 8(d) "
                            "figure assumes 5-slot tables and is kept as an equivalent only" % K}
    # physical DRAM traffic of the same launches from ncu --set full captures (profiles/ncu_constants.json)
    c4 = ncu.get("svf_streamed_2048x2048_4slot", {})
    res["forward"]["traffic_per_sweep"] = c4.get("dram_bytes_per_sweep") if n == 2048 else None
    res["forward"]["traffic_source"] = c4.get("source")
    res["forward"]["moved_bytes_per_sweep"] = float(moved["forward"]) * S
    res["forward"]["algorithmic_bytes_per_sweep"] = float(SVF_BYTES_PER_STATE_SWEEP) * S
    del tabs
    torch.cuda.empty_cache()
    ms5 = timed(E.gridworld_tables(n, 0.2, slots=5))
    res["five_slot_tables"] = {"forward_us_per_sweep": 1e3 * ms5["svf"] / fw_sweeps,
                               "soft_vi_us_per_sweep": 1e3 * ms5["soft_vi"] / lap_sweeps,
                               "forward_traffic_per_sweep_ncu":
                                   ncu.get("svf_streamed_2048x2048_5slot", {}).get("dram_bytes_per_sweep") if n == 2048 else None}
    return res


def c5_slab(world, rank, dev, n=2048, fw_budget=10000, parity_n=256):
    """BASELINE configs[4] under torchrun: ONE 2048x2048 IcyGridWorld (4.2 M states) sharded by state-row slabs
    over the ranks (slab.PeerSlabGrid: one persistent dataflow kernel per GPU and fixed point, boundary rows
    through NVLink mailboxes, no collective on the sweep path).  Causal soft-VI gamma = 0.9 to convergence and the
    forward pass for a fixed sweep budget (convergence at this size needs 10^7+ sweeps), kernel
    time from CUDA events on each rank, max over ranks.  Beside it: the barrier-per-sweep kernel of round 1, the
    NCCL send/recv baseline, and a parity flag -- the N-rank result on a 256x256 world must be BITWISE the
    1-GPU cooperative-grid kernel's (soft-VI policy and sweep count, forward SVF)."""
    import torch
    import torch.distributed as dist
    import _irlb200 as E
    import slab
    S = n * n
    peak = float(measured_peaks().get("hbm_gbs", 6650.0))
    ref1 = ncu_constants().get("c5_single_gpu_reference", {})
    out = {"workload": "single %dx%d IcyGridWorld, state-row slabs over %d GPUs, soft-VI gamma=0.9 eps=1e-5 to convergence + "
                       "forward pass for %d sweeps (uniform policy: the diffuse worst case)" % (n, n, world, fw_budget),
           "states": S, "states_per_gpu": S // world, "table_slots": 4}

    def run(g, lap_budget, fw_b, reps):
        r = np.full(g.cnt, -0.1); phi = np.full(g.cnt, -np.inf); p0 = np.zeros(g.cnt)
        if g.hi == S:
            r[-1] = 1.0; phi[-1] = 0.0
        if g.lo == 0:
            p0[0] = 1.0
        uniform = torch.full((g.cnt, 4), 0.25, dtype=torch.float64, device=dev)
        best = None
        for _ in range(reps):
            E.launch_log = []
            torch.cuda.synchronize(); dist.barrier(); w0 = time.perf_counter()
            g.soft_vi(r, phi, 0.9, 1e-5, max_sweeps=lap_budget)
            n_lap = g.last_n_iter
            torch.cuda.synchronize(); dist.barrier(); w1 = time.perf_counter()
            d = g.svf(p0, [S - 1], uniform, 1e-5, max_sweeps=fw_b)
            n_fw = g.last_n_iter
            torch.cuda.synchronize(); dist.barrier(); w2 = time.perf_counter()
            log, E.launch_log = E.launch_log, None
            k = [a.elapsed_time(b) / 1e3 for nm, a, b in log if nm == "slab_persistent"]
            t = torch.tensor(k[:2] if len(k) >= 2 else [w1 - w0, w2 - w1], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t = t.tolist()
            if best is None or t[0] + t[1] < best[0] + best[1]:
                best = (t[0], t[1], n_lap, n_fw, float(d.sum()))
        return best

    def describe(t_lap, t_fw, n_lap, n_fw, mass):
        K = 4
        moved_fw, moved_lap = 12 * K + 24, 4 * K + 32 * K + 24      # bytes per state and sweep with 4-slot tables
        res = {"soft_vi_sweeps": n_lap, "soft_vi_us_per_sweep": 1e6 * t_lap / n_lap,
               "forward_sweeps": n_fw, "forward_us_per_sweep": 1e6 * t_fw / n_fw,
               "soft_vi_moved_GBps_aggregate": moved_lap * S * n_lap / t_lap / 1e9,
               "forward_moved_GBps_aggregate": moved_fw * S * n_fw / t_fw / 1e9,
               "forward_mass_local_rank0": mass}
        res["soft_vi_frac_of_N_x_hbm_peak"] = res["soft_vi_moved_GBps_aggregate"] / (world * peak)
        res["forward_frac_of_N_x_hbm_peak"] = res["forward_moved_GBps_aggregate"] / (world * peak)
        if ref1:
            res["soft_vi_strong_scaling_efficiency"] = ref1["soft_vi_us_per_sweep"] / (world * res["soft_vi_us_per_sweep"])
            res["forward_strong_scaling_efficiency"] = ref1["forward_us_per_sweep"] / (world * res["forward_us_per_sweep"])
        return res

    g = slab.PeerSlabGrid(n, 0.2, flow=True)
    out["dataflow_kernel"] = describe(*run(g, None, fw_budget, 2))
    out["dataflow_kernel"]["note"] = ("csrc/slab_flow.cu; per-GPU working set %.0f MB (forward) / %.0f MB (soft-VI): at 8 GPUs it is "
                                      "L2-resident, so the fraction of N x HBM peak is an equivalent there, not DRAM traffic"
                                      % (72.0 * S / world / 1e6, 168.0 * S / world / 1e6))
    g.close()
    g = slab.PeerSlabGrid(n, 0.2, flow=False)
    out["barrier_per_sweep_kernel_r01"] = describe(*run(g, 600, 1000, 2))
    g.close()
    g = slab.SlabGrid(n, 0.2, chunk=50)
    t_lap, t_fw, n_lap, n_fw, _ = run(g, 100, 200, 2)
    out["nccl_baseline"] = {"soft_vi_us_per_sweep": 1e6 * t_lap / n_lap, "forward_us_per_sweep": 1e6 * t_fw / n_fw,
                            "note": "one launch + NCCL send/recv of the ghost rows per sweep, votes all-reduced per 50 sweeps; wall time"}
    out["single_gpu_reference"] = ref1

    # ---- parity: N ranks == the 1-GPU kernel, bitwise ------------------------------------------
    m = parity_n
    Sm = m * m
    g = slab.PeerSlabGrid(m, 0.2, flow=True)
    rng = np.random.default_rng(11)
    rf = -0.1 + 0.02 * rng.standard_normal(Sm); rf[Sm - 1] = 1.0
    phif = np.full(Sm, -np.inf); phif[Sm - 1] = 0.0
    p0f = np.zeros(Sm); p0f[0] = 0.6; p0f[Sm // 2 + 3] = 0.4
    pol, _ = g.soft_vi(g.local(rf), g.local(phif), 0.9, 1e-5)
    n_lap = g.last_n_iter
    d = g.svf(g.local(p0f), [Sm - 1], pol, 1e-5, max_sweeps=3000)
    n_fw = g.last_n_iter
    t1 = E.gridworld_tables(m, 0.2, slots=4)
    pol1 = E.soft_vi(t1, E.terminal_phi([Sm - 1], Sm), rf, 0.9, mode=E.MODE_GRID)
    n_lap1 = int(E.last_info.n_iter.item())
    d1 = E.svf(t1, p0f, E.terminal_mask([Sm - 1], Sm), pol1, 1e-5, max_sweeps=3000, mode=E.MODE_GRID)
    n_fw1 = int(E.last_info.n_iter.item())
    ok = torch.tensor([int(bool((pol == pol1[0, g.lo:g.hi]).all()) and bool((d == d1[0, g.lo:g.hi]).all())
                           and n_lap == n_lap1 and n_fw == n_fw1)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["parity"] = {"world": "%dx%d over %d ranks vs the 1-GPU cooperative-grid kernel" % (m, m, world),
                     "bitwise_equal_policy_and_svf_on_every_rank": bool(ok.item()),
                     "soft_vi_sweeps": [n_lap, n_lap1], "forward_sweeps": [n_fw, n_fw1],
                     "oracle_parity": "tests/test_gpu_multi.py (-m gpu, >= 2 devices): same run against oracle/c to 1e-10"}
    g.close()
    return out


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------

def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import _irlb200 as E
    import maxent as M
    import optimizer as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = workload(args, rank)
    B, S = w["B"], w["S"]
    tabs = E.gridworld_tables(w["n"], w["p_slip"])          # setup, outside the timed region
    p0 = E.to_device(w["p0"])
    # synthetic expert statistics: the visitation frequencies of a hidden reward
    e_features, _ = M.compute_expected_svf_batch(tabs, p0, w["terminal"], w["r_true"], fused=False,
                                                 max_sweeps=args.max_sweeps)
    theta = E.to_device(w["theta0"])
    optim = O.Sga(lr=args.lr)
    optim.reset(theta)

    def step_device():
        # reward = features . omega with identity features is omega itself (maxent.py:244)
        svf, grad = M.compute_expected_svf_batch(tabs, p0, w["terminal"], theta, e_features=e_features, fused=False,
                                                 max_sweeps=args.max_sweeps)
        optim.step(grad)                                    # maxent.py:251, in place on the device omega
        return svf, grad

    # ---- device-resident timing ---------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                     # nvidia-smi needs a moment to come up: start it before the warm-up
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler.mark()
    E.launch_log = []
    l0 = E.n_launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sweeps, stati = [], []
    t0.record()
    for _ in range(args.steps):
        step_device()
        sweeps.append(E.last_info.n_iter)                   # device tensors, read after the region
        stati.append(E.last_info.status)
    t1.record()
    barrier()
    launches = E.n_launches - l0
    log, E.launch_log = E.launch_log, None
    clocks = sampler.stop() if rank == 0 else None
    ms_rank = t0.elapsed_time(t1)
    ms = torch.tensor([ms_rank], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * args.steps / (ms_total / 1e3)

    # per-kernel durations (CUDA events on the launching stream) and algorithmic bytes
    dur = {}
    for name, a, b in log:
        dur.setdefault(name, []).append(a.elapsed_time(b))
    n_fw = np.array([s[:, 1].sum().item() for s in sweeps], dtype=np.float64)      # forward sweeps per launch
    svf_ms = float(np.mean(dur.get("svf", [float("nan")])))
    bwd_ms = float(np.mean(dur.get("backward", [float("nan")])))
    svf_bytes = float(n_fw.mean()) * S * SVF_BYTES_PER_STATE_SWEEP
    bwd_bytes = float(B) * 2 * S * S * BWD_BYTES_PER_STATE_SWEEP
    peaks = measured_peaks()
    ncu = ncu_constants().get("svf_grid5_kernel", {})
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_equiv = svf_bytes / (svf_ms / 1e3) / 1e9
    # The dominant kernel keeps tables, weights and the iterate on chip for the 10^4-10^5 sweeps of a launch, so
    # HBM is not what bounds it (ncu: DRAM busy 0.2 %).  Its roofline is the shared-memory datapath: one
    # wavefront of 128 B per clock per SM (scripts/ubench.cu, scripts/ubench_shfl.cu); a 2x4-tile thread moves
    # 12 halo loads + 8 stores of 8 B per sweep = 160 wavefronts per 1024-state world-sweep (0 bank conflicts).
    props = torch.cuda.get_device_properties(local)
    n_sm = props.multi_processor_count
    sm_clock_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0)
    wf_per_ws = float(ncu.get("smem_wavefronts_per_1024_state_world_sweep", 160.0)) * S / 1024.0
    smem_bytes = float(n_fw.mean()) * wf_per_ws * 128.0                   # per launch
    smem_achieved = smem_bytes / (svf_ms / 1e3) / 1e9
    smem_peak = 128.0 * n_sm * sm_clock_hz / 1e9
    cyc = (svf_ms / 1e3) * sm_clock_hz * n_sm / float(n_fw.mean())
    traffic = float(ncu["dram_bytes_per_world"]) * B if (S == 1024 and "dram_bytes_per_world" in ncu) else None
    roofline = {"bound": "smem",
                "kernel": "svf_grid5_kernel (forward state-visitation sweeps, stencil-tiled, one CTA per world)",
                "achieved": smem_achieved, "peak": smem_peak, "unit": "GB/s", "frac": smem_achieved / smem_peak,
                "peak_source": "128 B per clock per SM x %d SMs x the SM clock sampled during the timed region "
                               "(%.0f MHz); shared-memory datapath width measured by scripts/ubench.cu" % (
                                   n_sm, sm_clock_hz / 1e6),
                "achieved_source": "shared-memory wavefronts per launch (%.0f per world-sweep x the launch's world-sweeps, "
                                   "all counted live) x 128 B / CUDA-event time of the launch" % wf_per_ws,
                "traffic": traffic,
                "traffic_source": ncu.get("source"),
                "sm_cycles_per_world_sweep": cyc, "smem_wavefronts_per_world_sweep": wf_per_ws,
                "fp64_pipe_frac": float(ncu.get("fp64_pipe_cycles_per_1024_state_world_sweep", 110.0)) * S / 1024.0 / cyc,
                "launch_ms": svf_ms, "share_of_step": svf_ms * args.steps / ms_total if ms_total else None,
                "forward_sweeps_per_world_mean": float(n_fw.mean()) / B,
                "forward_sweeps_per_world_max": int(max(int(s[:, 1].max().item()) for s in sweeps)),
                "worlds_stopped_by_guard": int(sum(int((s[:, 1] == 2).sum().item()) for s in stati)),
                "launch_order": "longest-first by the previous step's sweep counts (IRLB200_LPT=%s)" % os.environ.get("IRLB200_LPT", "1"),
                "hbm_equivalent": {
                    "note": "NOT a roofline fraction: the algorithmic bytes of SURVEY 8(d) (84 B per state and sweep) that a "
                            "streaming implementation would move, divided by the launch time, against the measured HBM copy "
                            "bandwidth; the physically HBM-bound regime is other_configs.roofline_stream / c5_slab",
                    "algorithmic_bytes_per_launch": svf_bytes, "GBps": hbm_equiv, "hbm_peak_GBps": hbm_peak,
                    "ratio_to_hbm_peak": hbm_equiv / hbm_peak},
                "backward": {"launch_ms": bwd_ms, "algorithmic_bytes_per_launch": bwd_bytes}}

    # ---- end to end through the public API with host buffers --------------------
    theta_h = torch.empty((B, S), dtype=torch.float64).pin_memory()
    theta_h.copy_(torch.as_tensor(w["theta0"]))
    grad_h = torch.empty((B, S), dtype=torch.float64).pin_memory()
    ef_dev = e_features

    def step_e2e():
        th = theta_h.to(dev, non_blocking=True)                             # H2D: this step's omega
        _, grad = M.compute_expected_svf_batch(tabs, p0, w["terminal"], th, e_features=ef_dev, fused=False,
                                               max_sweeps=args.max_sweeps)
        grad_h.copy_(grad, non_blocking=False)                              # D2H: the gradient
        theta_h.add_(grad_h, alpha=args.lr)                                 # host-side Sga step (optimizer.py:107)

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    wall = time.perf_counter() - tw0
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), 1e3 * wall)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(e2e_ms.item()) / 1e3)

    # ---- per-rank breakdown: why weak scaling is not exactly 1.0 ---------------------------
    # (no collective on the data path: a rank's time is its own work; `value` divides by the slowest rank)
    mine = torch.tensor([ms_rank / args.steps, svf_ms, bwd_ms, float(n_fw.mean()),
                         float(max(int(s[:, 1].max().item()) for s in sweeps))], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    per_rank = [{"rank": r, "ms_per_step": float(t[0]), "forward_launch_ms": float(t[1]), "backward_launch_ms": float(t[2]),
                 "forward_world_sweeps_per_step": float(t[3]), "forward_sweeps_max_world": int(t[4]),
                 "ns_per_world_sweep": 1e6 * float(t[1]) / float(t[3])} for r, t in enumerate(allr)]
    tot = np.array([r["forward_world_sweeps_per_step"] for r in per_rank])
    tms = np.array([r["ms_per_step"] for r in per_rank])
    rank_balance = {"work_max_over_mean": float(tot.max() / tot.mean()), "ms_max_over_mean": float(tms.max() / tms.mean()),
                    "ms_max_over_rank0": float(tms.max() / tms[0]),
                    "note": "every rank draws its own 4096 rewards (weak scaling), so the ranks' total forward sweeps differ; "
                            "the driver's efficiency v_N / (N v_1) compares the slowest rank with rank 0's own batch"}

    c5 = None
    if world > 1 and not args.no_other_configs:
        del tabs, e_features, theta
        torch.cuda.empty_cache()
        try:
            c5 = c5_slab(world, rank, dev)
        except Exception as e:                                             # side measurements never sink the line
            import traceback
            sys.stderr.write("[rank %d] C5 slab block failed:\n%s\n" % (rank, traceback.format_exc()))
            c5 = {"error": repr(e)}

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args, world), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * S * 8,
                        "d2h_bytes_per_step": B * S * 8},
                "gpu_launches": int(launches), "roofline": roofline,
                "per_rank": per_rank, "rank_balance": rank_balance}
        if c5 is not None:
            line["other_configs"] = {"C5_slab": c5}
        if world == 1 and not args.no_other_configs:
            try:
                line["other_configs"] = other_configs()
            except Exception as e:                                         # side measurements never sink the line
                line["other_configs"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            t, n_svf = cpu_full_body(w, 0)
            line["cpu_baseline"] = {"value": 1.0 / t, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                                    "sample": "world 0 of the batch, one full gradient-step body with the dense "
                                              "numpy restatement (2S = %d backward + %d forward dense sweeps), "
                                              "%.1f s" % (2 * S, n_svf, t)}
            ts, _ = cpu_sparse_body(w, 0)
            line["cpu_baseline"]["sparse_restatement"] = {
                "value": 1.0 / ts, "unit": UNIT, "cores": 1,
                "note": "scipy-CSR restatement of the same body (not the reference's dense arithmetic), %.1f s" % ts}
            try:
                v, thr, nw, dt, ns = cpu_c_openmp(w)
                line["cpu_baseline"]["c_openmp_restatement"] = {
                    "value": v, "unit": UNIT, "cores": thr,
                    "note": "plain-C sparse restatement, one world per OpenMP thread: worlds 0..%d of the batch in "
                            "%.1f s (%.0f forward sweeps per world); the optimised-CPU line, not the reference" % (
                                nw - 1, dt, ns)}
            except Exception as e:
                line["cpu_baseline"]["c_openmp_restatement"] = {"error": repr(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference_arm(a) if a.impl == "reference" else run_b200_arm(a))
