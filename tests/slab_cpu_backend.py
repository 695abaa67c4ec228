"""TEST-ONLY arithmetic backend for irl-maxent_b200/slab.py: numpy sweeps on torch CPU tensors,
so that the slab protocol (partition, ghost-row exchange, chunked votes, snapshot/replay) can run
under `gloo` without a GPU.  Mirrors the per-state expressions of csrc/phases.cuh / slab_kernels.cu.
Not part of the product (which has no CPU fallback)."""
import numpy as np
import torch

from oracle import dense_port as D


class NumpyBackend:
    device = torch.device("cpu")

    def local_tables(self, size, p_slip, icy, lo, cnt):
        P = D.icy_gridworld_table(size, p_slip) if icy else D.gridworld_table(size)
        S, A, K = size * size, 4, 5
        nz = (P != 0).any(axis=2)
        si = np.zeros((K, cnt), np.int32); sp = np.zeros((A, K, cnt))
        pi = np.zeros((K, cnt), np.int32); pp = np.zeros((A, K, cnt))
        for i in range(cnt):
            s = lo + i
            succ, pred = np.nonzero(nz[s])[0], np.nonzero(nz[:, s])[0]
            si[:, i], pi[:, i] = s, s
            si[:len(succ), i] = succ
            pi[:len(pred), i] = pred
            sp[:, :len(succ), i] = P[s, succ, :].T
            pp[:, :len(pred), i] = P[pred, s, :].T
        return dict(A=A, K=K, succ_idx=torch.as_tensor(si), succ_p=torch.as_tensor(sp),
                    pred_idx=torch.as_tensor(pi), pred_p=torch.as_tensor(pp))

    def sweep(self, op, lo, cnt, A, K, idx, p, c0, c1, discount, eps, vi_mean, x_in, x_out, vote, policy):
        idx, p, c0 = idx.numpy(), p.numpy(), c0.numpy()
        xin, xout = x_in.numpy(), x_out.numpy()
        g = xin[idx]                                           # [K, cnt]
        with np.errstate(all="ignore"):
            if op == 3:
                acc = np.zeros(cnt)
                for j in range(K):
                    acc = p[j] * g[j] + acc
                x = c0 + acc
            else:
                q = np.zeros((A, cnt))
                for a in range(A):
                    dot = np.zeros(cnt)
                    for j in range(K):
                        dot = p[a, j] * g[j] + dot
                    q[a] = c0 + discount * dot if op == 1 else discount * dot
                if op == 1:
                    x = c1.numpy().copy()
                    for a in range(A):
                        x = D.softmax(x, q[a])
                    if policy is not None:
                        policy.numpy()[:] = np.exp(q - x[None, :]).T
                else:
                    x = c0 + (q.sum(axis=0) / A if vi_mean else q.max(axis=0))
            diff = np.abs(x - xin[lo:lo + cnt])
        if (diff > eps).any():
            vote.numpy()[0] |= 1
        if np.isnan(diff).any():
            vote.numpy()[1] |= 1
        xout[lo:lo + cnt] = x

    def weights(self, cnt, A, K, pred_idx, pred_p, policy_full, mask_full, W):
        idx, p, pol, mask = pred_idx.numpy(), pred_p.numpy(), policy_full.numpy(), mask_full.numpy()
        w = np.zeros((K, cnt))
        for j in range(K):
            acc = np.zeros(cnt)
            for a in range(A):
                acc = p[a, j] * pol[idx[j], a] + acc
            w[j] = np.where(mask[idx[j]] != 0, 0.0, acc)
        W.numpy()[:] = w
