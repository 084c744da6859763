import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, euc_b200 as e
from euc_b200 import scenes
n=int(sys.argv[1]); w,h=640,64
r = scenes.u01(1234 + n, n * 3 * 8).reshape(n, 3, 8)
v = np.zeros((n, 3), dtype=e.VERTEX_P4C4)
v["pos"][:, :, 0] = -0.9 + r[:, :, 0] * 0.08
v["pos"][:, :, 1] = 0.2 + r[:, :, 1] * 0.6
v["pos"][:, :, 2] = 0.9 - 0.8 * (np.arange(n)[:, None] / n) + r[:, :, 2] * 1e-4
v["pos"][:, :, 3] = 1.0
v["rgba"][:, :, :3] = r[:, :, 3:6]; v["rgba"][:, :, 3] = 0.3
px = e.Buffer2d.fill([w, h], 0xFF000000, dtype=np.uint32); z = e.Buffer2d.fill([w, h], 1.0)
e.BlendTris().render(v.reshape(-1), px, z)
print(n, "ok", np.count_nonzero(px.raw()!=0xFF000000))
