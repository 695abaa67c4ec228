"""
plot -- the plotting entry points `main.py` calls (`/root/reference/src/plot.py`),
as thin matplotlib wrappers.  Visualisation is outside the accelerated path; this
module exists so that the reference's driver runs unchanged.  Values may arrive
as numpy arrays or CUDA tensors.
"""

import numpy as np


def _np(x):
    return x.cpu().numpy() if type(x).__module__.startswith("torch") else np.asarray(x)


def plot_state_values(ax, world, values, border, **kwargs):
    """Heat map of per-state values (reference: plot.py:74-97)."""
    img = ax.imshow(np.reshape(_np(values), (world.size, world.size)), origin='lower', **kwargs)
    if border is not None:
        for i in range(world.size + 1):
            ax.plot([i - 0.5, i - 0.5], [-0.5, world.size - 0.5], **border, label=None)
            ax.plot([-0.5, world.size - 0.5], [i - 0.5, i - 0.5], **border, label=None)
    return img


def plot_deterministic_policy(ax, world, policy, **kwargs):
    """One arrow per state (reference: plot.py:100-118)."""
    arrows = [(0.33, 0), (-0.33, 0), (0, 0.33), (0, -0.33)]
    policy = _np(policy)
    for s in range(world.n_states):
        cx, cy = world.state_index_to_point(s)
        dx, dy = arrows[int(policy[s])]
        ax.arrow(cx - 0.5 * dx, cy - 0.5 * dy, dx, dy, head_width=0.1, **kwargs)


def plot_stochastic_policy(ax, world, policy, border=None, **kwargs):
    """Four triangles per cell coloured by p(a|s) (reference: plot.py:121-178)."""
    policy = _np(policy)
    n = world.size
    # vertices: cell corners then cell centres
    corners = [(x - 0.5, y - 0.5) for y in range(n + 1) for x in range(n + 1)]
    centres = [(x, y) for y in range(n) for x in range(n)]
    xy = np.array(corners + centres)
    tris, vals = [], []
    for s in range(world.n_states):
        cx, cy = world.state_index_to_point(s)
        bl, br = cy * (n + 1) + cx, cy * (n + 1) + cx + 1
        tl, tr = (cy + 1) * (n + 1) + cx, (cy + 1) * (n + 1) + cx + 1
        c = (n + 1) ** 2 + s
        # action order: right, left, up, down
        tris += [(br, tr, c), (tl, bl, c), (tr, tl, c), (bl, br, c)]
        vals += [policy[s, 0], policy[s, 1], policy[s, 2], policy[s, 3]]
    ax.set_aspect('equal')
    ax.set_xlim(-0.5, n - 0.5)
    ax.set_ylim(-0.5, n - 0.5)
    p = ax.tripcolor(xy[:, 0], xy[:, 1], tris, facecolors=np.array(vals), vmin=0.0, vmax=1.0, **kwargs)
    if border is not None:
        ax.triplot(xy[:, 0], xy[:, 1], tris, **border)
    return p


def plot_trajectory(ax, world, trajectory, **kwargs):
    """Poly-line through the visited cells (reference: plot.py:181-197)."""
    pts = [world.state_index_to_point(s) for s in trajectory.states()]
    x, y = zip(*pts)
    return ax.plot(x, y, **kwargs)
