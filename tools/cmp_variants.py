"""Debug aid: render the same inputs with the direct and the staged forward kernels and report where they differ."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import ml_gmpi_b200 as g
from ml_gmpi_b200 import _lib, synth

lib = _lib.load()
N, res, V = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (32, 256, 8)
d = torch.device("cuda:0")
case = synth.make_case(n_planes=N, tex=res, img=res, n_mpi=V, seed=1234, device=d)
out = {}
for name, var in (("direct", 1), ("staged", 2)):
    lib.gmpi_debug_set_fwd_variant(var)
    c, dp = g.render_views(case.rgba, case.dhw, case.view2mpi, case.ray_dir, case.eye, case.z_dir, check_last_plane=True)
    torch.cuda.synchronize()
    out[name] = (c.cpu().numpy(), dp.cpu().numpy())
for k in (0, 1):
    a, b = out["direct"][k], out["staged"][k]
    diff = np.abs(a - b)
    idx = np.unravel_index(np.argmax(diff), diff.shape)
    print(("color", "depth")[k], "max|direct-staged| =", diff.max(), "at", idx, "direct", a[idx], "staged", b[idx], "rel", diff.max() / np.abs(a).max())
    bad = np.argwhere(diff > 1e-5 * np.abs(a).max())
    print("  #elements off by >1e-5:", len(bad), "of", diff.size)
    if len(bad):
        vs, cs, ys, xs = bad[:, 0], bad[:, 1], bad[:, 2], bad[:, 3]
        print("  views", np.unique(vs), "y range", ys.min(), ys.max(), "x range", xs.min(), xs.max())
        print("  y mod 30 histogram:", np.bincount(ys % 30, minlength=30))
        print("  x mod 64 hist (8 bins):", np.bincount((xs % 64) // 8, minlength=8))

# which one is right?  oracle on the worst view
sys.path.insert(0, "oracle")
import mpi_oracle
n = lambda t: t.cpu().numpy()
v = int(idx[0]) if 'idx' in dir() else 0
for v in sorted(set([0, V - 1])):
    rc, rd, fl = mpi_oracle.forward(n(case.rgba[v:v + 1]), np.zeros(1, np.int32), n(case.dhw[v:v + 1]), n(case.ray_dir[v:v + 1]),
                                    n(case.eye[v:v + 1]), n(case.z_dir[v:v + 1]), nthreads=32)
    for name in ("direct", "staged"):
        dc = np.abs(out[name][0][v] - rc[0]); dd = np.abs(out[name][1][v] - rd[0])
        wi = np.unravel_index(np.argmax(dc), dc.shape)
        print(f"view {v} {name}: color max err {dc.max():.3e} at {wi}, depth max err {dd.max():.3e}, #color>1e-5: {(dc > 1e-5).sum()}")
