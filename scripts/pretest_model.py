"""NumPy model of the quantised phase-1 pre-test proposed for round 2:
is it conservative (no true neighbour rejected) and how many false survivors does it send to
the exact phase-2 test?  No GPU needed.

Stencil of a 9 x 4 x 4-cell tile: 13 x 8 x 8 cells of 1.5 dx; positions relative to the stencil
origin, quantised per axis to one byte with ONE unit for all axes (u = x-extent / 255); the test
is   sum_axes (q_i - q_j)^2  <=  T   on integers (vabsdiff4 + dp4a on the device).
Each quantised coordinate is off by less than u/2 (round to nearest), a difference by less than
u, so T = ceil((cutoff / u + sqrt(3))^2) can never reject a pair with |r_ij| < cutoff.
"""

import numpy as np

rng = np.random.default_rng(0)
dx, cs = 1.0, 1.5
cut = 3.0 * dx
ext = np.array([13, 8, 8]) * cs
u = ext[0] / 255.0
T = int(np.ceil((cut / u + np.sqrt(3.0)) ** 2))
print(f"unit {u:.4f} dx, cutoff {cut / u:.2f} units, threshold T = {T} (sqrt {np.sqrt(T):.2f})")

for name, jitter in (("lattice", 0.0), ("disordered", 0.35)):
    n = np.round(ext / dx).astype(int)
    g = np.stack(np.meshgrid(*[np.arange(k) for k in n], indexing="ij"), -1).reshape(-1, 3)
    x = (g + 0.5) * dx + jitter * dx * rng.standard_normal((len(g), 3))
    x = np.clip(x, 0, ext - 1e-6)
    q = np.rint(x / u).astype(np.int64)
    own = np.nonzero(np.all((x >= 2 * cs) & (x < ext - 2 * cs), axis=1))[0]  # the tile's particles
    own = rng.choice(own, size=min(400, len(own)), replace=False)
    miss = fp = true = cand = 0
    for i in own:
        ci = np.floor(x[i] / cs)
        win = np.all(np.abs(np.floor(x / cs) - ci) <= 2, axis=1)  # the 5 x 5 x 5-cell window
        d2 = ((x[win] - x[i]) ** 2).sum(1)
        dq = ((q[win] - q[i]) ** 2).sum(1)
        inside, keep = d2 < cut**2, dq <= T
        miss += int((inside & ~keep).sum())
        fp += int((~inside & keep).sum())
        true += int(inside.sum())
        cand += int(win.sum())
    print(f"{name:10s}: candidates/particle {cand / len(own):6.1f}  neighbours {true / len(own):6.1f}  "
          f"rejected neighbours {miss}  false survivors {fp / len(own):5.1f} "
          f"(+{100.0 * fp / true:.1f} % of the phase-2 work)")
