"""One-off converter: the reference ships its fetal-motion trajectories as pickled
``scipy.interpolate.interp1d`` objects (``svort/data/traj.npy``, read by ``sample_motion``,
``svort/data/fetal_motion.py:15-19``).  All 154 + 154 of them are *linear* interpolants on the
integer knots 0..T, so they are fully described by their knot values: this script stores those as
plain arrays (no pickle) in ``fetalsyngen_b200/generator/artifacts/traj_knots.npz``.

    python tools/convert_traj.py [/root/reference]
"""
import sys
from pathlib import Path

import numpy as np

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
rot, trans = np.load(ref / "fetalsyngen/generator/artifacts/svort/data/traj.npy", allow_pickle=True)
out = {}
for name, trajs in (("rot", rot), ("trans", trans)):
    ys, offs, T, dT = [], [0], [], []
    for f, t, dt in trajs:
        assert f._kind == "linear" and np.array_equal(f.x, np.arange(len(f.x))) and f.y.shape == (len(f.x), 3)
        ys.append(np.asarray(f.y, dtype=np.float64))
        offs.append(offs[-1] + len(f.x))
        T.append(t)
        dT.append(dt)
    out[f"{name}_y"] = np.concatenate(ys)
    out[f"{name}_off"] = np.asarray(offs, dtype=np.int64)
    out[f"{name}_T"] = np.asarray(T, dtype=np.float64)
    out[f"{name}_dT"] = np.asarray(dT, dtype=np.float64)
dst = Path(__file__).resolve().parent.parent / "fetalsyngen_b200/generator/artifacts/traj_knots.npz"
np.savez_compressed(dst, **out)
print(dst, dst.stat().st_size, {k: v.shape for k, v in out.items()})
