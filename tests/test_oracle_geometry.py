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


def test_trajectory_invariants(O):
    """What every Breakout trajectory of the reference satisfies, whatever the arithmetic details: the ball stays between the
    walls and under the ceiling (up to the contact prediction), moves 4.0 per step along straight legs, bricks only ever
    disappear and each one scores 1 (mechanics.rs:150-162), the paddle stays on the board at |speed| <= 160 in 0.001 steps
    (:553-587,612-649), an episode ends exactly when the ball passes the paddle line or no brick is left (:131-135)."""
    n, steps, seed = 48, 2500, 21
    env = O.VecEnv(n, seed=seed)
    acts = O.synthetic_actions(seed, 0, n, 0, steps)
    # half of the envs track the ball so that episodes last and bricks get hit
    prev = env.state()
    total_hits = 0
    for t in range(steps):
        a = acts[t].copy()
        track = np.arange(n) < n // 2
        centre = (prev["pad_min_x"] + prev["pad_max_x"]) / 2
        a[track] = np.where(prev["ball_cx"][track] < centre[track] - 5, 1, np.where(prev["ball_cx"][track] > centre[track] + 5, 2, 0))
        reward, done = env.step(a)
        cur = env.state()
        fresh = done != 0                                       # auto-reset: these envs show a new episode now
        live = ~fresh & (cur["err"] == 0) & (prev["err"] == 0)
        # bricks only disappear, one point each; reward is the score difference
        gone = prev["bricks"] & ~cur["bricks"]
        assert not np.any((cur["bricks"] & ~prev["bricks"])[live])
        hits = np.array([bin(int(x)).count("1") for x in gone])
        assert np.array_equal(hits[live], reward[live].astype(np.int64))
        assert np.array_equal((cur["score"] - prev["score"])[live], hits[live])
        total_hits += int(hits[live].sum())
        # ball inside the walls / under the ceiling (0.8 contact prediction + rounding), paddle on the board
        assert np.all(cur["ball_cx"][live] >= 10.0 - 1e-3) and np.all(cur["ball_cx"][live] <= 590.0 + 1e-3)
        assert np.all(cur["ball_cy"][live] >= 10.0 - 1e-3)
        assert np.all(cur["pad_min_x"] >= 0.0) and np.all(cur["pad_max_x"] <= 600.0)
        assert np.allclose((cur["pad_max_x"] - cur["pad_min_x"]), 60.0, atol=1e-3)
        sp = cur["pad_speed"]
        assert np.all(np.abs(sp) <= 160.0) and np.allclose(sp * 1000.0, np.round(sp * 1000.0), atol=1e-2)
        # a leg without a bounce moves the ball by exactly |mv| = 200 * 0.02 = 4 in its direction
        same_dir = live & (cur["ball_dx"] == prev["ball_dx"]) & (cur["ball_dy"] == prev["ball_dy"])
        moved = np.hypot(cur["ball_cx"] - prev["ball_cx"], cur["ball_cy"] - prev["ball_cy"])
        assert np.allclose(moved[same_dir], 4.0, atol=1e-3)
        # with bounces the path is folded, never longer - once the direction is a unit vector: until the first bounce of an episode it
        # is (dx, -1), |.| = 1.01..1.06, and the reference advances the centre by direction * way (mechanics.rs:166), i.e. a bit further
        unit = np.abs(np.hypot(prev["ball_dx"], prev["ball_dy"]) - 1.0) < 1e-5
        assert np.all(moved[live & unit] <= 4.0 + 1e-3)
        assert np.all(moved[live & ~unit] <= 4.0 * 1.07)
        # an unfinished env has its ball above the paddle line and bricks left
        assert np.all(cur["ball_cy"][~fresh & (cur["finished"] == 0)] < 575.0) and np.all(cur["bricks"][~fresh & (cur["finished"] == 0)] != 0)
        prev = cur
    st = env.stats()
    assert st["episodes"] > n and total_hits > 200                # random play loses quickly, tracking play hits bricks
    env.close()


def _raster_numpy(cx, cy, pmin, pmax, bricks):
    """The raster spec of DESIGN.md section 5, restated with whole-frame numpy float32 arrays (no loops over shapes' bounding
    boxes, no conservative bounds): bricks luma 96, ball ring luma 236 on top, paddle luma 255 on top."""
    f = np.float32
    px = (np.arange(84, dtype=np.float32) + f(0.5))[None, :]
    py = (np.arange(84, dtype=np.float32) + f(0.5))[:, None]
    sc = lambda v: f(f(v) * f(84.0)) / f(600.0)
    img = np.zeros((84, 84), dtype=np.uint8)
    for b in range(60):
        if (int(bricks) >> b) & 1:
            r, k = divmod(b, 20)
            x0, x1 = sc(30.0 + 27.0 * k), sc(55.0 + 27.0 * k)
            y0, y1 = sc(35.0 + 27.0 * r), sc(60.0 + 27.0 * r)
            img[((px >= x0) & (px < x1)) & ((py >= y0) & (py < y1))] = 96
    bx, by, rs = sc(cx), sc(cy), sc(10.0)
    dx, dy = px - bx, py - by
    d2 = dx * dx + dy * dy
    out2, in2 = f(rs + f(1.0)) * f(rs + f(1.0)), f(rs - f(1.0)) * f(rs - f(1.0))
    img[(d2 <= out2) & (d2 >= in2)] = 236
    x0, x1, y0, y1 = sc(pmin), sc(pmax), sc(565.0), sc(575.0)
    img[((px >= x0) & (px < x1)) & ((py >= y0) & (py < y1))] = 255
    return img


def test_rendered_frames_follow_the_raster_spec(O):
    """Every newest frame of 30 envs over 220 steps (tracking play: bricks disappear, the ball visits the brick band, the
    walls and the paddle) equals the raster spec restated independently in numpy; the frame lands in ring slot (k-1) mod 4
    and older slots keep older frames (frame_ring_buffer.rs:53-63)."""
    n, steps, seed = 30, 220, 8
    env = O.VecEnv(n, seed=seed)
    acts = O.synthetic_actions(seed, 0, n, 0, steps)
    prev = env.state()
    last = {}
    checked = 0
    for t in range(steps):
        a = acts[t].copy()
        centre = (prev["pad_min_x"] + prev["pad_max_x"]) / 2
        a[: n // 2] = np.where(prev["ball_cx"][: n // 2] < centre[: n // 2] - 5, 1, np.where(prev["ball_cx"][: n // 2] > centre[: n // 2] + 5, 2, 0))
        _, done = env.step(a)
        cur = env.state()
        obs = env.obs_u8()
        for e in range(n):
            k = int(cur["episode_step"][e])
            if done[e] or k == 0:
                last.pop(e, None)
                continue                                        # restarted on this step: the stack was cleared
            want = _raster_numpy(cur["ball_cx"][e], cur["ball_cy"][e], cur["pad_min_x"][e], cur["pad_max_x"][e], cur["bricks"][e])
            slot = (k - 1) % 4
            assert np.array_equal(obs[e, slot], want), "env %d step %d" % (e, t)
            if e in last and k >= 2:
                assert np.array_equal(obs[e, (k - 2) % 4], last[e])          # the previous frame is still in its slot
            if k < 4:
                assert not obs[e, k:].any()                                  # slots not written yet in this episode are zero
            last[e] = want
            checked += 1
        prev = cur
    assert checked > 0.9 * n * steps
    env.close()
