"""Dam-break validation run on the engine (GPU box).

    python scripts/validate_dambreak.py [dx] [t_end] > profiles/r02_validation_dambreak.txt

The reference's validation (validation/db2d.sh, validate.py:462-519 `val_DB`) runs cases/db.yaml
(2D water column L x H = 2 x 1 in a tank 5.366 x 2, g = 1, Colagrossi & Landrini 2003) and plots
pressure snapshots at t = 1.62, 2.38, 4.0, 5.21, 6.02, 7.23; it asserts nothing.  This script
runs the same case -- particles laid out as cases/db.py:94-127 does (bench.db2d_state), the
db.yaml solver: SPH + generalized wall BC + density evolution + artificial viscosity 0.1,
dt = 0.0003 -- through the C ABI and records what those figures show:

  * the surge front x_front(t) (right-most fluid particle, measured from the tank's left wall),
    against the shallow-water (Ritter) bound 2 sqrt(g H) and the front speed SPH and experiments
    agree on for this geometry (~1.6-2.0 sqrt(g H) once the column has collapsed);
  * the column height at the left wall h(0, t), which must fall monotonically;
  * the arrival of the front at the right wall (between the reference's first and second
    snapshot, 1.62 < t < 2.38: in val_DB's first figure the tongue is still travelling, in the
    second the run-up on the wall has begun);
  * at the reference's time stamps: particles inside the tank, maximum fluid pressure, the
    fraction of the fluid in the right half of the tank.

`run()` is what tests/test_gpu_validation.py asserts on.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

L_WALL, H_WALL, L, H = 5.366, 2.0, 2.0, 1.0
STAMPS = (1.62, 2.38, 4.0, 5.21, 6.02, 7.23)  # validate.py:487


def run(dx=0.02, t_end=2.6, dt=0.0003, every=100, log=None):
    import torch

    import bench
    from jax_sph_b200 import Engine, make_config

    state, meta = bench.db2d_state(dx)
    meta = dict(meta, dt=dt)
    cfg = make_config(2, meta["box"], dx, dt, tvf=0.0, c_ref=meta["c_ref"], p_ref=meta["p_ref"],
                      **meta["cfg_kwargs"])
    n = len(state["r"])
    eng = Engine(cfg, n)
    eng.upload({k: torch.from_numpy(v) for k, v in state.items()})
    fluid = state["tag"] == 0
    wall = meta["n_walls"] * dx
    steps = int(round(t_end / dt))
    curve, snaps = [], {}
    stamp_steps = {int(round(t / dt)): t for t in STAMPS if t <= t_end + 1e-9}
    done = 0
    while done < steps:
        nxt = min(steps, (done // every + 1) * every)
        for s in stamp_steps:
            if done < s < nxt:
                nxt = s
        eng.step(dt, nxt - done)
        done = nxt
        got = eng.download(keys=("r", "p", "u"), host=True)
        r = got["r"].numpy()[fluid]
        t = done * dt
        x_front = float(r[:, 0].max()) - wall
        near = r[:, 0] < wall + 2 * dx
        h0 = float(r[near, 1].max()) - wall if near.any() else 0.0
        curve.append((t, x_front, h0))
        if done in stamp_steps:
            p = got["p"].numpy()[fluid]
            inside = ((r[:, 0] > wall - dx) & (r[:, 0] < wall + L_WALL + dx) & (r[:, 1] > wall - dx)).mean()
            snaps[stamp_steps[done]] = dict(
                inside=float(inside), p_max=float(p.max()), p_mean=float(p.mean()),
                right_half=float((r[:, 0] - wall > L_WALL / 2).mean()),
                umax=float(np.abs(got["u"].numpy()[fluid]).max()))
        if log:
            log(f"t={t:6.3f} x_front={x_front:6.3f} h0={h0:5.3f}")
    err = eng.error()
    return dict(n=n, n_fluid=int(fluid.sum()), dx=dx, dt=dt, steps=steps, err=err,
                curve=np.array(curve), snaps=snaps, counters=eng.counters())


def summary(res):
    c = res["curve"]
    t, x, h0 = c[:, 0], c[:, 1], c[:, 2]
    arrive = t[np.argmax(x >= L_WALL - 3 * res["dx"])] if (x >= L_WALL - 3 * res["dx"]).any() else None
    travelling = (x > 2.6) & (x < 4.8)  # collapsed column, tongue well before the wall
    speed = float(np.polyfit(t[travelling], x[travelling], 1)[0]) if travelling.sum() > 3 else None
    return dict(arrival=arrive, front_speed=speed, h0_end=float(h0[-1]),
                h0_monotone=bool((np.diff(h0) <= 2.5 * res["dx"]).all()))


def main():
    dx = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
    t_end = float(sys.argv[2]) if len(sys.argv) > 2 else 7.3
    res = run(dx, t_end)
    s = summary(res)
    print(f"# 2D dam break, cases/db.yaml (SPH + wall BC + density evolution + alpha 0.1), dx={dx} "
          f"dt={res['dt']} N={res['n']} ({res['n_fluid']} fluid), {res['steps']} steps, "
          f"device error word {res['err']}, {res['counters']}")
    print(f"# front arrival at the right wall: t = {s['arrival']} (reference snapshots: travelling at "
          f"1.62, run-up at 2.38); front speed while travelling: {s['front_speed']:.3f} sqrt(g H) "
          f"(Ritter bound 2); column height at the left wall at the end: {s['h0_end']:.3f} H, "
          f"monotone: {s['h0_monotone']}")
    for ts, v in sorted(res["snaps"].items()):
        print(f"# t={ts}: " + " ".join(f"{k}={x:.4f}" for k, x in v.items()))
    print("# t  x_front/H  h(0)/H")
    for t, x, h in res["curve"]:
        print(f"{t:.3f} {x:.4f} {h:.4f}")


if __name__ == "__main__":
    main()
