"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's MaxEnt / MaxCausalEnt IRL hot path
(`/root/reference/src/maxent.py`, `solver.py:9-52`, `gridworld.py`,
`optimizer.py`).  Nothing in the product path (`irl-maxent_b200/`) may import,
link or execute anything in this directory.  The only permitted users are
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py`, and there only as the checker / CPU baseline.

Parity status: PINNED.  The reference's own tests hold no golden vectors for
this path (`src/test_gridworld.py` pins only the zero pattern of
`p_transition`), so the oracle is pinned against outputs of the reference
itself: `tests/golden/generate_golden.py` imports the unmodified reference from
`/root/reference/src` in the build container and writes the fixtures in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every oracle
function against every fixture, `tests/test_oracle_c.py` does the same for the
plain-C restatement.
"""

from . import dense_port, sparse_port  # noqa: F401
# oracle.c_port (plain-C restatement, oracle/c/irl_oracle.c) is imported on demand: it builds
# oracle/_build/libirl_oracle.so with `make -C oracle` when missing.
