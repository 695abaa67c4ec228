"""The dense batch path's contraction against cuBLAS DGEMM of the same shape (torch.matmul, float64), sustained."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
S, A, B, sweeps = 1024, 4, 4096, 24
rng = np.random.default_rng(5)
P = torch.as_tensor(rng.random((S, S, A))).cuda() ** 3
P[S - 1] = 0.0; P[S - 1, S - 1, :] = 1.0
P /= P.sum(dim=1, keepdim=True)
rewards = torch.as_tensor(-0.2 + 0.1 * rng.standard_normal((B, S))).cuda()
phi = E.terminal_phi([S - 1], S)
dt = E.DenseTables(P)
def timed(fn, reps=1):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 1e3 / reps
flops = 2.0 * A * S * S * B
t = timed(lambda: E.dense_soft_vi(dt, phi, rewards, 0.9, 1e-30, max_sweeps=sweeps))
print("dense soft-VI loop: %.1f us per sweep, %.1f TFLOP/s (GEMM + epilogue + bookkeeping launches)" % (1e6 * t / sweeps, flops * sweeps / t / 1e12))
Pa = dt.packed[:A * S * S].view(A * S, S)
X = rewards.t().contiguous()                          # [S, B]
t = timed(lambda: torch.matmul(Pa, X), reps=sweeps)
print("cuBLAS DGEMM [%d x %d] . [%d x %d]: %.1f us, %.1f TFLOP/s sustained over %d calls" % (A * S, S, S, B, 1e6 * t, flops / t / 1e12, sweeps))
v = torch.empty((B, S), dtype=torch.float64, device="cuda")
E.launch_log = []
E.dense_value_iteration(dt, rewards, 0.9, 1e-30, max_sweeps=sweeps)
torch.cuda.synchronize()
log, E.launch_log = E.launch_log, None
ms = log[-1][1].elapsed_time(log[-1][2])
print("dense VI loop (cheap epilogue): %.1f us per sweep, %.1f TFLOP/s" % (1e3 * ms / sweeps, flops * sweeps / (ms / 1e3) / 1e12))
