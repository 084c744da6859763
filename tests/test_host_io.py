"""Row N4: host I/O around the render call (euc_b200/io.py)."""
import io as _io
import os

import numpy as np
import pytest

import euc_b200 as e
from euc_b200 import scenes

OBJ = """# tiny mesh
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vn 0 0 1
vn 0 1 0
f 1//1 2//1 3//2
f 1//1 3//2 4//1 2//1
f -4//-2 -3//-1 -2//1
"""


def test_load_obj_face_order_and_triangulation():
    s = e.io.load_obj(_io.StringIO(OBJ))
    assert s.dtype == e.VERTEX_PN and s.shape == (3 + 6 + 3,)          # triangle, quad fan (2 triangles), negative indices
    assert s["pos"][:3].tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0]]  # file order, like wavefront's Obj::vertices()
    assert s["normal"][2].tolist() == [0, 1, 0]
    assert s["pos"][3:9].tolist() == [[0, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 0], [0, 1, 0], [1, 0, 0]]
    assert s["pos"][9:].tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0]] and s["normal"][9].tolist() == [0, 0, 1]


def test_load_obj_matches_the_teapot_fixture(tmp_path):
    # write the fixture back out as OBJ and read it through the loader: same stream as scenes.teapot_stream()
    d = np.load(os.path.join(os.path.dirname(e.__file__), "data", "teapot.npz"))
    p = tmp_path / "t.obj"
    with open(p, "w") as f:
        for v in d["positions"]:
            f.write("v %r %r %r\n" % tuple(float(x) for x in v))
        for n in d["normals"]:
            f.write("vn %r %r %r\n" % tuple(float(x) for x in n))
        for tri in d["faces"]:
            f.write("f " + " ".join("%d//%d" % (a + 1, b + 1) for a, b in tri) + "\n")
    s = e.io.load_obj(str(p))
    ref = scenes.teapot_stream()
    assert s.shape == ref.shape and np.array_equal(s["pos"], ref["pos"]) and np.array_equal(s["normal"], ref["normal"])


@pytest.mark.gpu
def test_texture_from_image_and_readback_ring():
    ctx = e.default_context()
    img = scenes.rust_texture()
    tex = e.io.texture_from_image(img[:, :, :3])           # RGB input gets an opaque alpha channel
    back = tex.raw().view(np.uint8).reshape(img.shape)
    assert np.array_equal(back[:, :, :3], img[:, :, :3]) and (back[:, :, 3] == 255).all()
    # three frames in flight through the pinned ring: every frame comes back intact and in order
    w, h = 640, 480
    color = e.Buffer2d.fill([w, h], 0, dtype=np.uint32)
    ring = e.io.ReadbackRing(color, depth=3)
    slots = []
    for k in range(5):
        color.clear(0x01010101 * (k + 1))
        slots.append((ring.submit(color), k))
        if len(slots) == 3:
            s, kk = slots.pop(0)
            assert (ring.wait(s) == 0x01010101 * (kk + 1)).all()
    for s, kk in slots:
        assert (ring.wait(s) == 0x01010101 * (kk + 1)).all()
    ring.close()
