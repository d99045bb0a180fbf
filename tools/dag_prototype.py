"""Prototype of the static conflict DAG of one mode-S step (round 2 design study).

For a synthetic S2-like state (agents at density d on a torus, range r cyclic in 1..6, footprint
dilation k) it builds the DAG "A -> B iff priority(A) < priority(B) and the footprints intersect"
and reports edges per agent, the level histogram (longest path) and the width profile -- the
numbers the dataflow scheduler of ssb_dataflow.cuh is sized on.
"""
import sys
import numpy as np


def crosses_conflict(dx, dy, ra, rb, ka, kb):
    """L1 distance between cross(A) (radius ra at the origin) and cross(B) (radius rb at (dx, dy)) <= ka + kb."""
    adx, ady = abs(dx), abs(dy)
    # distance between two axis-aligned segments; arms: A-NS x=0,|y|<=ra; A-EW y=0,|x|<=ra; B-NS x=dx,|y-dy|<=rb; B-EW y=dy,|x-dx|<=rb
    def gap(a, r):          # distance from coordinate a to interval [-r, r]
        return max(a - r, 0)
    d_ns_ns = adx + gap(ady, ra + rb)                       # parallel vertical arms
    d_ew_ew = ady + gap(adx, ra + rb)
    d_ns_ew = gap(adx, rb) + gap(ady, ra)                   # A's NS arm against B's EW arm
    d_ew_ns = gap(adx, ra) + gap(ady, rb)
    return min(d_ns_ns, d_ew_ew, d_ns_ew, d_ew_ns) <= ka + kb


def main(n_side=512, density=0.29, rmax=6, k=0, seed=1):
    rng = np.random.default_rng(seed)
    W = H = n_side
    n = int(W * H * density)
    cells = rng.choice(W * H, n, replace=False)
    xs, ys = cells // H, cells % H
    r = 1 + rng.permutation(n) % rmax
    prio = rng.permutation(n)
    grid = -np.ones((W, H), dtype=np.int64)
    grid[xs, ys] = np.arange(n)
    R = 2 * rmax + 2 * k
    preds = [[] for _ in range(n)]
    n_edges = 0
    for a in range(n):
        x, y, ra = xs[a], ys[a], r[a]
        for dx in range(-R, R + 1):
            col = grid[(x + dx) % W]
            for dy in range(-R, R + 1):
                if dx == 0 and dy == 0:
                    continue
                if abs(dx) > rmax + 2 * k and abs(dy) > rmax + 2 * k:
                    continue
                b = col[(y + dy) % H]
                if b < 0 or prio[b] > prio[a]:
                    continue
                if crosses_conflict(dx, dy, ra, r[b], k, k):
                    preds[a].append(b)
                    n_edges += 1
    order = np.argsort(prio)
    level = np.zeros(n, dtype=np.int64)
    for a in order:
        if preds[a]:
            level[a] = 1 + max(level[b] for b in preds[a])
    hist = np.bincount(level)
    print(f"side {n_side} density {density} rmax {rmax} k {k}: agents {n}, edges/agent {n_edges / n:.2f}, depth {level.max() + 1}")
    print("level histogram (fraction):", " ".join(f"{h / n:.3f}" for h in hist))


if __name__ == "__main__":
    args = [float(a) if "." in a else int(a) for a in sys.argv[1:]]
    main(*args)
