"""include/rtb/umath.h pinned INDEPENDENTLY of its two consumers.

The CUDA kernels and the CPU oracle both include umath.h, so GPU-vs-oracle parity cannot see an error they share
(a misreading of Unity.Mathematics, SURVEY.md §8c assumption A1).  Here every vector / quaternion function the sample
path calls is compared with a float64 numpy restatement written from the published Unity.Mathematics formulas —
not from umath.h — on random inputs, to a few float32 ulps of the result's scale; plus the algebraic identities the
formulas imply (rotation preserves length, inverse undoes transform, reflect is an involution) and the NaN rules of
math.min / max / saturate that `samplesToAccumulate` depends on (SampleBatchJob.cs:118-126)."""
import numpy as np

RNG = np.random.default_rng(20261017)
N = 4096
EPS = float(np.finfo(np.float32).eps)


def unit_quats(n):
    q = RNG.normal(size=(n, 4))
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


def vecs(n, scale=3.0):
    return (RNG.normal(size=(n, 3)) * scale).astype(np.float32)


def close(got, want64, scale, ulps):
    """|got - want| <= ulps * eps * scale, per record (scale: magnitude of the intermediate terms)."""
    err = np.abs(got.astype(np.float64) - want64)
    tol = ulps * EPS * np.maximum(scale, 1e-30)
    assert (err <= tol).all(), float((err / tol).max())


# ---- float64 restatements of the Unity.Mathematics definitions (math.cs / quaternion.cs / RigidTransform.cs) ----
def rotate64(q, v):            # t = 2 * cross(q.xyz, v); v + q.w * t + cross(q.xyz, t)
    t = 2.0 * np.cross(q[:, :3], v)
    return v + q[:, 3:4] * t + np.cross(q[:, :3], t)


def quat_to_matrix(q):         # the rotation matrix of a unit quaternion (textbook form, independent of rotate())
    x, y, z, w = (q[:, i] for i in range(4))
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)


def test_reflect(oracle):
    i, n = vecs(N), vecs(N, 1.0)
    n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    got = oracle.umath_vec(0, np.hstack([i, n]))
    i64, n64 = i.astype(np.float64), n.astype(np.float64)
    want = i64 - 2.0 * n64 * np.sum(i64 * n64, axis=1, keepdims=True)       # reflect(i, n) = i - 2 n dot(i, n)
    scale = np.abs(i64).sum(axis=1, keepdims=True) * 3
    close(got, want, scale, 4)
    # reflecting twice about a unit normal is the identity
    back = oracle.umath_vec(0, np.hstack([got, n]))
    close(back, i64, scale, 16)


def test_rotate_matches_the_rotation_matrix_and_preserves_length(oracle):
    q, v = unit_quats(N), vecs(N)
    got = oracle.umath_vec(1, np.hstack([q, v]))
    q64, v64 = q.astype(np.float64), v.astype(np.float64)
    scale = np.linalg.norm(v64, axis=1, keepdims=True) * 4
    close(got, rotate64(q64, v64), scale, 6)
    close(got, np.einsum("nij,nj->ni", quat_to_matrix(q64), v64), scale, 6)   # same rotation as the matrix of q
    assert np.abs(np.linalg.norm(got.astype(np.float64), axis=1) - np.linalg.norm(v64, axis=1)).max() < 3e-6 * 10
    # identity quaternion: exactly v (what the sphere fast path relies on)
    ident = np.tile(np.array([0, 0, 0, 1], np.float32), (N, 1))
    assert np.array_equal(oracle.umath_vec(1, np.hstack([ident, v])), v)


def test_rigid_inverse_and_transform(oracle):
    q, p, x = unit_quats(N), vecs(N, 10.0), vecs(N, 10.0)
    inv = oracle.umath_vec(2, np.hstack([q, p]))
    q64, p64, x64 = q.astype(np.float64), p.astype(np.float64), x.astype(np.float64)
    # inverse(RigidTransform): rot = conj(q) / dot(q, q); pos = rotate(invRot, -pos)
    inv_rot = np.hstack([-q64[:, :3], q64[:, 3:4]]) / np.sum(q64 * q64, axis=1, keepdims=True)
    close(inv[:, :4], inv_rot, np.ones((N, 1)), 4)
    pscale = np.linalg.norm(p64, axis=1, keepdims=True) * 4
    close(inv[:, 4:], rotate64(inv_rot, -p64), pscale, 8)
    # transform(rt, x) = rotate(rt.rot, x) + rt.pos, and inverse undoes it
    fwd = oracle.umath_vec(3, np.hstack([q, p, x]))
    scale = (np.linalg.norm(x64, axis=1, keepdims=True) + np.linalg.norm(p64, axis=1, keepdims=True)) * 4
    close(fwd, rotate64(q64, x64) + p64, scale, 8)
    back = oracle.umath_vec(3, np.hstack([inv, fwd]))
    close(back, x64, scale, 40)
    # a non-unit quaternion: the inverse divides by |q|^2 (Entity.cs:52 takes whatever rotation the host passes)
    q2 = (q * 1.7).astype(np.float32)
    inv2 = oracle.umath_vec(2, np.hstack([q2, p]))
    q264 = q2.astype(np.float64)
    close(inv2[:, :4], np.hstack([-q264[:, :3], q264[:, 3:4]]) / np.sum(q264 * q264, axis=1, keepdims=True), np.ones((N, 1)), 4)


def test_lerp_normalize_cross_dot_mul(oracle):
    a, b, s = vecs(N), vecs(N), RNG.random((N, 1)).astype(np.float32)
    a64, b64, s64 = a.astype(np.float64), b.astype(np.float64), s.astype(np.float64)
    close(oracle.umath_vec(4, np.hstack([a, b, s])), a64 + s64 * (b64 - a64), np.abs(a64) + np.abs(b64), 4)     # lerp = a + s (b - a)
    nrm = oracle.umath_vec(5, a)
    close(nrm, a64 / np.linalg.norm(a64, axis=1, keepdims=True), np.ones((N, 1)), 4)                            # rsqrt(dot(v, v)) * v
    assert np.abs(np.linalg.norm(nrm.astype(np.float64), axis=1) - 1).max() < 4 * EPS
    sc = (np.abs(a64).sum(axis=1, keepdims=True)) * (np.abs(b64).sum(axis=1, keepdims=True))
    close(oracle.umath_vec(6, np.hstack([a, b])), np.cross(a64, b64), sc, 3)
    close(oracle.umath_vec(8, np.hstack([a, b])), np.sum(a64 * b64, axis=1, keepdims=True), sc, 3)
    c0, c1, c2, v = vecs(N), vecs(N), vecs(N), vecs(N)
    want = c0.astype(np.float64) * v[:, 0:1] + c1.astype(np.float64) * v[:, 1:2] + c2.astype(np.float64) * v[:, 2:3]   # column-constructed float3x3
    msc = (np.abs(c0) + np.abs(c1) + np.abs(c2)).astype(np.float64) * np.abs(v).astype(np.float64).max(axis=1, keepdims=True)
    close(oracle.umath_vec(7, np.hstack([c0, c1, c2, v])), want, msc, 4)


def test_scalar_rules_and_nan_semantics(oracle):
    x = np.concatenate([RNG.normal(size=N - 6) * 5, [0.5, 1.5, 2.5, -0.5, -1.5, 1e-3]]).astype(np.float32)
    y = (RNG.normal(size=N) * 5).astype(np.float32)
    z = (RNG.normal(size=N) * 5).astype(np.float32)
    out = oracle.umath_vec(9, np.stack([x, y, z], 1))
    x64, y64, z64 = x.astype(np.float64), y.astype(np.float64), z.astype(np.float64)
    assert np.array_equal(out[:, 0], np.minimum(x, y)) and np.array_equal(out[:, 1], np.maximum(x, y))
    assert np.array_equal(out[:, 2], np.clip(x, 0, 1))
    assert np.array_equal(out[:, 3], np.rint(x))                                        # MathF.Round: half to even
    close(out[:, 4:5], ((z64 - x64) / (y64 - x64))[:, None], (np.abs((z64 - x64) / (y64 - x64)))[:, None] + 1e-30, 4)   # unlerp = (x - a) / (b - a)
    close(out[:, 5:6], (x64 + z64 * (y64 - x64))[:, None], (np.abs(x64) + np.abs(z64) * (np.abs(y64) + np.abs(x64)))[:, None], 4)
    close(out[:, 6:7], (1.0 / x64)[:, None], np.abs(1.0 / x64)[:, None], 1)                # rcp = 1 / x, correctly rounded
    close(out[:, 7:8], (1.0 / np.sqrt(np.abs(x64)))[:, None], (1.0 / np.sqrt(np.abs(x64)))[:, None], 2)
    # math.min(x, y) = isnan(y) || x < y ? x : y: the non-NaN operand when exactly one is NaN; saturate(NaN) = 1
    nan = np.float32("nan")
    r = oracle.umath_vec(9, np.array([[nan, 2.0, 0.0], [2.0, nan, 0.0], [nan, nan, 0.0]], np.float32))
    assert r[0, 0] == 2.0 and r[0, 1] == 2.0 and r[1, 0] == 2.0 and r[1, 1] == 2.0
    assert np.isnan(r[2, 0]) and np.isnan(r[2, 1])
    assert r[0, 2] == 1.0 and r[2, 2] == 1.0                                            # saturate(NaN) == 1
