"""Converts the reference's example assets into binary fixtures that travel with the repo (the reference tree
does not exist on the GPU box).  Run in the build container only:  python tools/make_data_fixtures.py

  /root/reference/examples/data/teapot.obj -> euc_b200/data/teapot.npz   (positions, normals, face index triples)
  /root/reference/examples/data/rust.png   -> euc_b200/data/rust_rgba.npz (860x899 RGBA8, `image::open(..).to_rgba8()`)

The OBJ is all-triangle `f v//vn` faces; wavefront 0.2's `Obj::vertices()` yields the three face-vertices of every
`f` line in file order (benches/teapot.rs:151-152, :189), which is the order the stream is stored in here.
"""
import os
import numpy as np
from PIL import Image

REF = "/root/reference/examples/data"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "euc_b200", "data")


def main():
    v, vn, faces = [], [], []
    with open(os.path.join(REF, "teapot.obj")) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                v.append([float(x) for x in t[1:4]])
            elif t[0] == "vn":
                vn.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                assert len(t) == 4, "teapot.obj is all triangles"
                tri = []
                for c in t[1:]:
                    a, b, n = c.split("/")
                    assert b == ""
                    tri.append((int(a) - 1, int(n) - 1))
                faces.append(tri)
    v = np.array(v, dtype=np.float32)
    vn = np.array(vn, dtype=np.float32)
    faces = np.array(faces, dtype=np.int32)  # (2256, 3, 2): position index, normal index
    assert v.shape == (1202, 3) and vn.shape == (1202, 3) and faces.shape == (2256, 3, 2)
    np.savez_compressed(os.path.join(OUT, "teapot.npz"), positions=v, normals=vn, faces=faces)

    img = np.array(Image.open(os.path.join(REF, "rust.png")).convert("RGBA"), dtype=np.uint8)
    assert img.shape == (899, 860, 4)
    np.savez_compressed(os.path.join(OUT, "rust_rgba.npz"), rgba=img)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
