"""float32 matrix helpers with the conventions of the `vek` crate the reference's benches use
(benches/teapot.rs:154-179, examples/texture_mapping.rs:127-133).  Matrices are numpy (4,4) arrays indexed
m[row, col]; uniforms are computed once on the host and handed bit-identically to the oracle and to the device,
so the exact rounding of these helpers is not parity-relevant."""
import numpy as np

f32 = np.float32


def identity():
    return np.eye(4, dtype=f32)


def translation_3d(v):
    m = identity()
    m[0, 3], m[1, 3], m[2, 3] = f32(v[0]), f32(v[1]), f32(v[2])
    return m


def scaling_3d(v):
    v = (v, v, v) if np.isscalar(v) else v
    m = identity()
    m[0, 0], m[1, 1], m[2, 2] = f32(v[0]), f32(v[1]), f32(v[2])
    return m


def rotation_x(a):
    c, s = f32(np.cos(f32(a))), f32(np.sin(f32(a)))
    m = identity()
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    return m


def rotation_y(a):
    c, s = f32(np.cos(f32(a))), f32(np.sin(f32(a)))
    m = identity()
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


def rotation_z(a):
    c, s = f32(np.cos(f32(a))), f32(np.sin(f32(a)))
    m = identity()
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def mul(*ms):
    out = ms[0].astype(f32)
    for m in ms[1:]:
        out = (out.astype(f32) @ m.astype(f32)).astype(f32)
    return out


def perspective_fov_lh_zo(fov_y, width, height, near, far):
    """Left-handed, depth zero-to-one."""
    fov_y, width, height, near, far = map(f32, (fov_y, width, height, near, far))
    h = f32(np.cos(fov_y / f32(2)) / np.sin(fov_y / f32(2)))
    w = f32(h * height / width)
    m = np.zeros((4, 4), dtype=f32)
    m[0, 0], m[1, 1] = w, h
    m[2, 2] = far / (far - near)
    m[2, 3] = -(far * near) / (far - near)
    m[3, 2] = f32(1)
    return m


def perspective_fov_rh_no(fov_y, width, height, near, far):
    """Right-handed, depth negative-one-to-one (OpenGL style)."""
    fov_y, width, height, near, far = map(f32, (fov_y, width, height, near, far))
    h = f32(np.cos(fov_y / f32(2)) / np.sin(fov_y / f32(2)))
    w = f32(h * height / width)
    m = np.zeros((4, 4), dtype=f32)
    m[0, 0], m[1, 1] = w, h
    m[2, 2] = -(far + near) / (far - near)
    m[2, 3] = -(f32(2) * far * near) / (far - near)
    m[3, 2] = f32(-1)
    return m


def _norm(v):
    v = np.asarray(v, dtype=f32)
    return (v / f32(np.sqrt(f32(np.dot(v, v))))).astype(f32)


def look_at_lh(eye, target, up):
    eye, target, up = (np.asarray(x, dtype=f32) for x in (eye, target, up))
    f = _norm(target - eye)
    s = _norm(np.cross(up, f))
    u = np.cross(f, s).astype(f32)
    m = identity()
    m[0, :3], m[1, :3], m[2, :3] = s, u, f
    m[0, 3], m[1, 3], m[2, 3] = -np.dot(s, eye), -np.dot(u, eye), -np.dot(f, eye)
    return m.astype(f32)


def inverted(m):
    return np.linalg.inv(m.astype(np.float64)).astype(f32)


def mul_point(m, p):
    v = m.astype(f32) @ np.array([p[0], p[1], p[2], 1.0], dtype=f32)
    return (v[:3] / v[3]).astype(f32)
