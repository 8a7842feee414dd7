"""The few cgmath 0.18 constructors the reference's scene uses (tracing.rs:383,393,403), in f32.

Matrices are numpy float32 (4,4) in maths convention M[row, col]; `colmajor()` gives cgmath's
memory layout, which is what the C ABI takes.  Products are evaluated column by column, left to
right, the way cgmath's Matrix4 * Matrix4 does, so the bits match a Rust caller's.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def identity() -> np.ndarray:
    return np.eye(4, dtype=f32)


def from_translation(v) -> np.ndarray:
    m = identity()
    m[0, 3], m[1, 3], m[2, 3] = f32(v[0]), f32(v[1]), f32(v[2])
    return m


def from_scale(s) -> np.ndarray:
    return from_nonuniform_scale(s, s, s)


def from_nonuniform_scale(x, y, z) -> np.ndarray:
    m = identity()
    m[0, 0], m[1, 1], m[2, 2] = f32(x), f32(y), f32(z)
    return m


def _sin_cos_deg(deg):
    rad = f32(deg) * f32(np.pi / 180.0)  # Rad::from(Deg): deg * (PI/180 rounded to f32)
    return f32(np.sin(rad, dtype=f32)), f32(np.cos(rad, dtype=f32))


def from_angle_x(deg) -> np.ndarray:
    s, c = _sin_cos_deg(deg)
    m = identity()
    m[1, 1], m[2, 1], m[1, 2], m[2, 2] = c, s, -s, c
    return m


def from_angle_y(deg) -> np.ndarray:
    s, c = _sin_cos_deg(deg)
    m = identity()
    m[0, 0], m[2, 0], m[0, 2], m[2, 2] = c, -s, s, c
    return m


def from_angle_z(deg) -> np.ndarray:
    s, c = _sin_cos_deg(deg)
    m = identity()
    m[0, 0], m[1, 0], m[0, 1], m[1, 1] = c, s, -s, c
    return m


def mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a * b with cgmath's evaluation order: column j = a0*b[0,j] + a1*b[1,j] + a2*b[2,j] + a3*b[3,j]."""
    a = np.asarray(a, dtype=f32)
    b = np.asarray(b, dtype=f32)
    out = np.empty((4, 4), dtype=f32)
    for j in range(4):
        col = a[:, 0] * b[0, j]
        col = col + a[:, 1] * b[1, j]
        col = col + a[:, 2] * b[2, j]
        col = col + a[:, 3] * b[3, j]
        out[:, j] = col
    return out


def chain(*ms) -> np.ndarray:
    """m0 * m1 * m2 ... (left-associative, like the Rust expression)."""
    r = np.asarray(ms[0], dtype=f32)
    for m in ms[1:]:
        r = mul(r, m)
    return r


def inverse_transform(m: np.ndarray) -> np.ndarray | None:
    """Matrix4::inverse_transform: general inverse by cofactors in f32; None when singular."""
    a = np.asarray(m, dtype=f32)
    cof = np.empty((4, 4), dtype=f32)
    for r in range(4):
        for c in range(4):
            rows = [i for i in range(4) if i != r]
            cols = [j for j in range(4) if j != c]
            s = a[np.ix_(rows, cols)]
            det3 = (s[0, 0] * (s[1, 1] * s[2, 2] - s[1, 2] * s[2, 1])
                    - s[0, 1] * (s[1, 0] * s[2, 2] - s[1, 2] * s[2, 0])
                    + s[0, 2] * (s[1, 0] * s[2, 1] - s[1, 1] * s[2, 0]))
            cof[r, c] = f32(det3) if (r + c) % 2 == 0 else f32(-det3)
    det = f32(a[0, 0] * cof[0, 0] + a[0, 1] * cof[0, 1] + a[0, 2] * cof[0, 2] + a[0, 3] * cof[0, 3])
    if det == 0 or not np.isfinite(det):
        return None
    inv = (cof.T * (f32(1.0) / det)).astype(f32)
    if a[3, 0] == 0 and a[3, 1] == 0 and a[3, 2] == 0 and a[3, 3] == 1:
        # affine in, affine out: the device keeps 3x4 matrices (w is exactly 1)
        inv[3, :] = (0, 0, 0, 1)
    return inv


def colmajor(m: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(m, dtype=f32).T.reshape(16))
