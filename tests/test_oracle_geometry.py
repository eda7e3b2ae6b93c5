"""Geometric sanity of the oracle's collision routines beyond the reference's 13 known-answer vectors: on seeded random
configurations the restated parry2d contact / Ball::collision_check_with_rectangle (mechanics.rs:318-443) must describe
the actual geometry — computed here independently in float64 — whatever the implementation details are."""
import numpy as np
import pytest


def _dist_point_box(p, bmin, bmax):
    d = np.maximum(np.maximum(bmin - p, 0.0), p - bmax)
    return float(np.hypot(d[0], d[1]))


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        bmin = np.float32(rng.uniform(50, 500, size=2))
        size = np.float32(rng.choice([25.0, 60.0, 10.0], size=2))
        bmax = np.float32(bmin + size)
        # ball centre near the box surface, outside by 10..12 (touching range), random side / corner
        ang = rng.uniform(0, 2 * np.pi)
        centre = (bmin + bmax) / 2 + np.array([np.cos(ang), np.sin(ang)]) * (np.max(size) / 2 + rng.uniform(9.0, 16.0))
        c = np.float32(centre)
        if _dist_point_box(c.astype(np.float64), bmin.astype(np.float64), bmax.astype(np.float64)) < 10.0:
            continue                                          # the reference never starts a move penetrating
        mv_ang = rng.uniform(0, 2 * np.pi)
        mv = np.float32(np.array([np.cos(mv_ang), np.sin(mv_ang)]) * rng.uniform(0.5, 4.0))
        out.append((c, mv, bmin, bmax))
    return out


def test_rectangle_sweep_describes_the_geometry(O):
    n_hit = n_none = 0
    for c, mv, bmin, bmax in _cases(6000, 5):
        some, way, approx, nx, ny, err = O.collision_rect((float(c[0]), float(c[1])), 10.0, (float(mv[0]), float(mv[1])),
                                                          (float(bmin[0]), float(bmin[1])), (float(bmax[0]), float(bmax[1])))
        c64, mv64, lo, hi = c.astype(np.float64), mv.astype(np.float64), bmin.astype(np.float64), bmax.astype(np.float64)
        L = float(np.hypot(*mv64))
        d_end = _dist_point_box(c64 + mv64, lo, hi) - 10.0
        if some:
            n_hit += 1
            assert way <= L + 1e-4          # (the closed-form back-off may answer a position slightly BEHIND the start at grazing
                                            #  angles - negative way, mechanics.rs:401-402 - the contact position still has to be right)
            assert -1e-6 <= approx <= 0.8 + 1e-6                                  # ContactCandidates::consider's assert (:511)
            assert abs(np.hypot(nx, ny) - 1.0) < 1e-5
            p = c64 + mv64 / L * way
            assert abs((_dist_point_box(p, lo, hi) - 10.0) - approx) < 2e-3       # approximation = gap left at the contact position
            assert nx * mv64[0] + ny * mv64[1] < 0.0                              # only surfaces facing the motion (:327)
            # the surface normal points from the box to the ball
            q = np.clip(p, lo, hi)
            v = p - q
            assert (v[0] * nx + v[1] * ny) > 0.0
        else:
            n_none += 1
            # None means: out of prediction range at the end of the move, or the nearest surface faces away from the motion
            if d_end <= 0.8 - 1e-3:
                pe = c64 + mv64
                v = pe - np.clip(pe, lo, hi)
                if np.hypot(*v) > 1e-9:
                    assert v[0] * mv64[0] + v[1] * mv64[1] > -1e-3 * L, "a facing surface within range was not reported"
    assert n_hit > 500 and n_none > 500


def test_contact_distance_and_normal_via_zero_length_probe(O):
    """A tiny move towards the box turns the sweep into a plain contact query at the end point: way ~ |mv|, approximation =
    distance ball surface <-> box, normal = direction box -> ball."""
    rng = np.random.default_rng(9)
    checked = 0
    for _ in range(4000):
        bmin = np.float32(rng.uniform(100, 400, size=2)); bmax = np.float32(bmin + np.float32(25.0))
        ang = rng.uniform(0, 2 * np.pi)
        gap = rng.uniform(0.05, 0.75)
        centre0 = (bmin + bmax).astype(np.float64) / 2 + np.array([np.cos(ang), np.sin(ang)]) * 40.0
        q = np.clip(centre0, bmin, bmax)
        u = (centre0 - q) / np.hypot(*(centre0 - q))
        # not exactly anti-parallel to the normal: there the reference's acos(dot) can see dot < -1 after rounding -> NaN -> None
        rot = rng.uniform(0.15, 0.6) * rng.choice([-1.0, 1.0])
        w = np.array([np.cos(rot) * -u[0] - np.sin(rot) * -u[1], np.sin(rot) * -u[0] + np.cos(rot) * -u[1]]) * 0.01
        mv = np.float32(w)
        centre = np.float32(q + u * (10.0 + gap) - mv.astype(np.float64))
        some, way, approx, nx, ny, err = O.collision_rect((float(centre[0]), float(centre[1])), 10.0, (float(mv[0]), float(mv[1])),
                                                          (float(bmin[0]), float(bmin[1])), (float(bmax[0]), float(bmax[1])))
        assert some and err == 0
        end = centre.astype(np.float64) + mv.astype(np.float64)
        true_gap = _dist_point_box(end, bmin.astype(np.float64), bmax.astype(np.float64)) - 10.0
        assert abs(approx - true_gap) < 2e-4
        assert abs(way - np.hypot(*mv.astype(np.float64))) < 1e-6
        qe = np.clip(end, bmin, bmax); ve = (end - qe) / np.hypot(*(end - qe))
        assert abs(nx - ve[0]) < 1e-4 and abs(ny - ve[1]) < 1e-4
        checked += 1
    assert checked == 4000
