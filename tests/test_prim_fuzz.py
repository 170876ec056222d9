"""CPU tier: the kernels' leaf routine prim_test (csrc/ray_core.cuh, compiled for the host by
tests/emu) against the oracle's orc_ray_geom, bit for bit, on random AND adversarial (geom, ray)
cases.  The leaf routine takes shortcuts the oracle does not (sign-only bounding-sphere pre-test
without the square root and divisions, lazy second root, lazy pose fetch); every shortcut must be
invisible in the result."""
import numpy as np
import pytest

TYPES = np.array([0, 2, 3, 4, 5, 6], np.int32)  # plane sphere capsule ellipsoid cylinder box


def _rot(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1)


def _bound_radius(t, size):
    """radius of the pre-test sphere of capsule / cylinder / box (engine_ray.c), else a size scale"""
    r = np.linalg.norm(size, axis=1)
    r = np.where(t == 3, size[:, 0] + size[:, 1], r)
    r = np.where(t == 5, np.hypot(size[:, 0], size[:, 1]), r)
    return r


def _cases(seed, n):
    rng = np.random.default_rng(seed)
    t = TYPES[rng.integers(0, len(TYPES), n)]
    pos = rng.uniform(-20, 20, (n, 3))
    mat = _rot(rng, n)
    # a share of axis-aligned frames: exact zeros in lvec exercise the |lvec| > mjMINVAL guards
    ident = rng.random(n) < 0.15
    mat[ident] = np.eye(3).reshape(9)
    size = rng.uniform(0.05, 1.5, (n, 3))
    size[(t == 0) & (rng.random(n) < 0.5)] = 0.0  # unbounded planes
    pnt = rng.uniform(-25, 25, (n, 3))
    R = _bound_radius(t, size)
    c = pos - pnt
    dist = np.linalg.norm(c, axis=1)
    kind = rng.integers(0, 8, n)
    vec = rng.normal(size=(n, 3))
    # 1: aimed at the centre (+ noise); 2-3: tangent to the pre-test sphere up to a relative
    # perturbation of 1e-17 .. 1e-6; 4: aimed at a random point of the geom's box (edges, corners,
    # faces); 5: origin inside the bounding sphere; 6: pointing away; 7: unnormalised long / short
    aim = c + rng.normal(size=(n, 3)) * (0.3 * R)[:, None]
    vec = np.where((kind == 1)[:, None], aim, vec)
    u = np.cross(c, rng.normal(size=(n, 3)))
    u /= np.linalg.norm(u, axis=1, keepdims=True) + 1e-300
    eps = 10.0 ** rng.uniform(-17, -6, n) * rng.choice([-1.0, 0.0, 1.0], n)
    tang = c + u * (R * (1 + eps) * dist / np.sqrt(np.maximum(dist * dist - R * R, 1e-12)))[:, None]
    vec = np.where(((kind == 2) | (kind == 3))[:, None], tang, vec)
    corner = rng.choice([-1.0, 0.0, 1.0, 0.5], (n, 3)) * size
    m3 = mat.reshape(n, 3, 3)
    tgt = pos + np.einsum("nij,nj->ni", m3, corner)
    vec = np.where((kind == 4)[:, None], tgt - pnt, vec)
    inside = pos + rng.normal(size=(n, 3)) * (0.3 * R)[:, None]
    pnt = np.where((kind == 5)[:, None], inside, pnt)
    vec = np.where((kind == 6)[:, None], -c + rng.normal(size=(n, 3)), vec)
    scale = np.where(kind == 7, 10.0 ** rng.uniform(-6, 6, n), 1.0)
    vec = vec * scale[:, None]
    bad = np.einsum("ni,ni->n", vec, vec) < 1e-20
    vec[bad] = [1.0, 0.0, 0.0]
    return pos, mat, size, pnt, vec, t


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_prim_test_matches_oracle_bitwise(oracle, emu, seed):
    args = _cases(seed, 400_000)
    want = oracle.ray_geom_batch(*args)
    got = emu.prim_batch(*args)
    bad = np.nonzero(want.view(np.int64) != got.view(np.int64))[0]
    assert bad.size == 0, (bad[:5], want[bad[:5]], got[bad[:5]], args[5][bad[:5]])
    # the case mix must exercise both outcomes of every type
    for t in TYPES:
        sel = args[5] == t
        assert (want[sel] >= 0).any() and (want[sel] < 0).any()


def test_pretest_boundary_cases(oracle, emu):
    """rays whose pre-test discriminant sits within a few ulp of b*b (origin on the bounding
    sphere), and exactly tangent axis-aligned rays: the sign shortcut must fall back to the
    upstream arithmetic there."""
    rng = np.random.default_rng(7)
    n = 200_000
    t = np.full(n, 6, np.int32)
    size = rng.uniform(0.1, 1.0, (n, 3))
    pos = rng.uniform(-3, 3, (n, 3))
    mat = np.tile(np.eye(3).reshape(9), (n, 1))
    R = np.linalg.norm(size, axis=1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # origin on the bounding sphere up to a few ulp: c = |dif|^2 - R^2 ~ 0  =>  det ~ b*b
    pnt = pos + d * (R * (1 + rng.integers(-4, 5, n) * 2.0 ** -52))[:, None]
    vec = rng.normal(size=(n, 3))
    want = oracle.ray_geom_batch(pos, mat, size, pnt, vec, t)
    got = emu.prim_batch(pos, mat, size, pnt, vec, t)
    assert np.array_equal(want.view(np.int64), got.view(np.int64))
    # exactly tangent: origin at distance R off the axis, ray parallel to the axis
    pnt2 = pos + np.stack([R, np.zeros(n), -5 * np.ones(n)], axis=1)
    vec2 = np.tile([0.0, 0.0, 1.0], (n, 1))
    want = oracle.ray_geom_batch(pos, mat, size, pnt2, vec2, t)
    got = emu.prim_batch(pos, mat, size, pnt2, vec2, t)
    assert np.array_equal(want.view(np.int64), got.view(np.int64))


def test_sphere_miss_is_the_sign_of_ray_sphere(oracle, emu):
    """sphere_miss(pos, r*r, ...) must equal (mju_rayGeom of a SPHERE of radius r) < 0 in every case,
    in particular where sqrt(det) and b differ by an ulp or not at all: origins on the sphere up to
    a few ulp (det ~ b*b), tangent rays (det ~ 0), rays pointing away, huge and tiny |vec|."""
    rng = np.random.default_rng(11)
    n = 600_000
    pos = rng.uniform(-5, 5, (n, 3))
    r = rng.uniform(0.05, 3.0, n)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    kind = rng.integers(0, 6, n)
    k_ulp = rng.integers(-6, 7, n)
    on = pos + d * (r * (1 + k_ulp * 2.0 ** -52))[:, None]
    far = pos + d * (r * rng.uniform(1.0, 30.0, n))[:, None]
    inside = pos + d * (r * rng.uniform(0.0, 1.0, n))[:, None]
    pnt = np.where((kind <= 1)[:, None], on, np.where((kind == 2)[:, None], inside, far))
    vec = rng.normal(size=(n, 3))
    # kind 1: origin on the sphere, ray along +-d (b = +-|vec| r exactly representable-ish)
    sgn = rng.choice([-1.0, 1.0], n)
    vec = np.where((kind == 1)[:, None], d * sgn[:, None] * rng.uniform(0.1, 10, n)[:, None], vec)
    # kind 3: tangent from outside, perturbed by 0 .. 1e-6 relative
    c = pos - pnt
    dist = np.linalg.norm(c, axis=1)
    u = np.cross(c, rng.normal(size=(n, 3)))
    u /= np.linalg.norm(u, axis=1, keepdims=True) + 1e-300
    eps = 10.0 ** rng.uniform(-17, -6, n) * rng.choice([-1.0, 0.0, 1.0], n)
    tang = c + u * (r * (1 + eps) * dist / np.sqrt(np.maximum(dist * dist - r * r, 1e-12)))[:, None]
    vec = np.where((kind == 3)[:, None], tang, vec)
    # kind 5: extreme lengths
    vec = np.where((kind == 5)[:, None], vec * (10.0 ** rng.uniform(-14, 14, n))[:, None], vec)
    size = np.stack([r, r, r], axis=1)
    mat = np.tile(np.eye(3).reshape(9), (n, 1))
    want = oracle.ray_geom_batch(pos, mat, size, pnt, vec, np.full(n, 2, np.int32)) < 0
    got = emu.sphere_miss_batch(pos, r, pnt, vec)
    bad = np.nonzero(want != got)[0]
    assert bad.size == 0, (bad[:5], kind[bad[:5]])
    assert want.any() and (~want).any()
    # the shortcut's decision band must actually be hit: origins within a few ulp of the sphere
    sel = kind <= 1
    assert want[sel].any() and (~want[sel]).any()

