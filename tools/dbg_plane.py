import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import ml_gmpi_b200 as g
from ml_gmpi_b200 import _lib, synth
import mpi_oracle
lib = _lib.load()
N, res, V = 32, 256, 8
d = torch.device("cuda:0")
case = synth.make_case(n_planes=N, tex=res, img=res, n_mpi=V, seed=1234, device=d)
y, x = 134, 3
sl = slice(0, 1)
co = mpi_oracle.coords(np.zeros(1, np.int32), case.dhw[sl].cpu().numpy(), case.ray_dir[sl].cpu().numpy(), case.eye[sl].cpu().numpy(), res, res, True)
# corner coords of tile (0,4): x 0..63, y 120..149
for k in range(N):
    rg = case.rgba[sl].clone(); a = rg[:, :, 3].clone(); rg[:, :, 3] = 0; rg[:, k, 3] = a[:, k]
    res_ = {}
    for name, var in (("direct", 1), ("staged", 2)):
        lib.gmpi_debug_set_fwd_variant(var)
        c, dp = g.render_views(rg, case.dhw[sl], case.view2mpi[sl], case.ray_dir[sl], case.eye[sl], case.z_dir[sl])
        res_[name] = c[0, :, 120:150, 0:64].cpu().numpy()
    diff = np.abs(res_["direct"] - res_["staged"])
    ix = co[0, k, 0, 120:150, 0:64]; iy = co[0, k, 1, 120:150, 0:64]
    print(f"plane {k:2d}: max diff {diff.max():.2e} nbad {(diff>1e-6).sum():5d} | tile ix [{ix.min():8.2f},{ix.max():8.2f}] iy [{iy.min():8.2f},{iy.max():8.2f}] corners ix {ix[0,0]:.1f} {ix[0,-1]:.1f} {ix[-1,0]:.1f} {ix[-1,-1]:.1f}")
