"""Timing of fibers_stream on a cfg2-shaped peak field (145 x 174 x 145, three smooth random orientation fields with holes,
ellipsoid mask, 3 sub-voxel samples per seed voxel): whole call through the host-pointer C ABI and the kernel part alone
(wall clock around the call minus nothing: the call is blocking).  One JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import fibers_jl_b200 as F

shape = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "145,174,145").split(","))
g = np.random.default_rng(0)
nx, ny, nz = shape
xs, ys, zs = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(ny, dtype=np.float32), np.arange(nz, dtype=np.float32), indexing="ij")
vols = []
for i in range(3):
    ph = g.uniform(0, 2 * np.pi, 6)
    th = 0.6 * np.sin(xs / 15 + ph[0]) + 0.5 * np.cos(ys / 12 + ph[1]) + 0.3 * np.sin(zs / 9 + ph[2]) + i * 1.1
    el = 0.5 * np.sin(xs / 18 + ph[3]) * np.cos(ys / 21 + ph[4]) + 0.2 * np.sin(zs / 8 + ph[5])
    v = np.stack([np.cos(th) * np.cos(el), np.sin(th) * np.cos(el), np.sin(el)], axis=-1).astype(np.float32)
    v[g.random(shape) < 0.03 + 0.2 * i] = 0
    vols.append(np.asfortranarray(v))
ax = [np.linspace(-1, 1, n) for n in shape]
mask = np.asfortranarray(((ax[0][:, None, None] ** 2 + ax[1][None, :, None] ** 2 + ax[2][None, None, :] ** 2) <= 0.8).astype(np.uint8))
sub = F.draw_sublist(3, rng=1)
mr = [F.MRI(v) for v in vols]
t, tc, tf = [], [], []
for it in range(4):
    tm = {}
    t0 = time.perf_counter()
    tr = F.stream(mr, mask=F.MRI(mask), sublist=sub, timing=tm)
    t.append(time.perf_counter() - t0); tc.append(tm["call_s"]); tf.append(tm["fetch_s"])
npts = int(tr.npts.sum())
print(json.dumps({"kernel": "stream " + "x".join(map(str, shape)), "seeds": int(mask.sum()), "lines_started": int(mask.sum()) * 3,
                  "streamlines_kept": int(tr.n_count), "points": npts, "s_per_call_best": min(t[1:]), "s_first_call": t[0],
                  "fibers_stream_s": min(tc[1:]), "fibers_stream_fetch_s": min(tf[1:]),
                  "streamlines_per_s": tr.n_count / min(tc[1:]), "points_per_s": npts / min(tc[1:]),
                  "note": "fibers_stream_s = the blocking C call (H2D of the three vector volumes, pack, count pass, scans, write pass); fetch = D2H of the points into pageable numpy arrays; s_per_call also builds the Python list of per-line views"}))
