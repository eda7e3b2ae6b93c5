"""Generates tests/golden/contact_inputs_v1.txt — ~1,200 ball / box configurations (f32 bit patterns) for pinning the restated
parry2d `query::contact` (oracle/breakout_oracle.c: orc_contact_test_circle_aabb; reference call site algebra_2d.rs:62-75) at BIT
level wherever a Rust toolchain exists:

    cargo run -p trace-dumper --release -- --contacts contact_inputs_v1.txt > rust_contacts.txt      (reference workspace)
    python tests/compare_rust_trace.py --contacts rust_contacts.txt tests/golden/contact_inputs_v1.txt (this repository)

and tests/golden/contact_oracle_v1.txt — what the oracle answers today (guards the oracle against regressions on the CPU).
Line = cx cy radius min_x min_y max_x max_y  (hex f32 bits). Families: the game's own geometry (bricks 25 x 25, paddle 60 x 10,
radius 10) with the ball outside within / beyond the 0.8 prediction, touching, penetrating, centre inside the box, centre on
the boundary (parry's degenerate branch), corners, exact-axis cases, the 7 rectangle cases of mechanics.rs:708-722 at their
start / end / mid positions."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def f32hex(x):
    return "%08x" % int(np.float32(x).view(np.uint32))


def cases():
    rng = np.random.default_rng(20261018)
    out = []
    boxes = [((42.5, 47.5), (12.5, 12.5)), ((300.0, 570.0), (30.0, 5.0)), ((150.0, 74.5), (12.5, 12.5)), ((555.5, 101.5), (12.5, 12.5))]
    for _ in range(900):
        (bx, by), (hx, hy) = boxes[rng.integers(len(boxes))]
        bx += float(rng.uniform(-3, 3)) * (rng.random() < 0.3)
        r = 10.0
        kind = rng.integers(6)
        ang = rng.uniform(0, 2 * np.pi)
        # a point on / around the box outline in direction `ang`, at signed gap g from the surface
        g = [rng.uniform(0.0, 0.8), rng.uniform(0.8, 3.0), 0.0, -rng.uniform(0.0, 9.9), rng.uniform(-0.001, 0.001), rng.uniform(0.79, 0.81)][kind]
        d = np.array([np.cos(ang), np.sin(ang)])
        t = min(hx / max(abs(d[0]), 1e-12), hy / max(abs(d[1]), 1e-12))            # ray / box outline
        p = np.array([bx, by]) + d * t + d * (r + g)
        if rng.random() < 0.2:                                                       # exact-axis: face contacts
            p[rng.integers(2)] = [bx, by][0] if rng.random() < 0.5 else p[0]
        out.append((p[0], p[1], r, bx - hx, by - hy, bx + hx, by + hy))
    for _ in range(120):                                                             # ball centre inside the box
        (bx, by), (hx, hy) = boxes[rng.integers(len(boxes))]
        p = np.array([bx + rng.uniform(-hx, hx), by + rng.uniform(-hy, hy)])
        out.append((p[0], p[1], 10.0, bx - hx, by - hy, bx + hx, by + hy))
    for _ in range(60):                                                              # centre exactly on the boundary / corners
        (bx, by), (hx, hy) = boxes[rng.integers(len(boxes))]
        sx, sy = rng.choice([-1, 1]), rng.choice([-1, 1])
        p = [(bx + sx * hx, by + rng.uniform(-hy, hy)), (bx + rng.uniform(-hx, hx), by + sy * hy), (bx + sx * hx, by + sy * hy)][rng.integers(3)]
        out.append((p[0], p[1], 10.0, bx - hx, by - hy, bx + hx, by + hy))
    rst = [((10.0, 0.0), (150.0, 90.0), (170.0, 110.0)), ((5.0, 0.0), (110.0, 90.0), (130.0, 110.0)), ((3.0, -3.0), (100.0, 70.0), (120.0, 93.0)),
           ((-8.0, -8.0), (70.0, 80.0), (90.0, 100.0)), ((-1.46, -1.46), (80.0, 80.0), (95.0, 95.0)), ((-5.0, -5.0), (80.0, 80.0), (95.0, 95.0)),
           ((-4.2, -4.2), (80.0, 80.0), (90.0, 90.0))]
    for mv, lo, hi in rst:                                                           # mechanics.rs:708-722
        for f in (0.0, 0.25, 0.5, 0.75, 1.0):
            cx = np.float32(100.0) + np.float32(mv[0]) * np.float32(f)
            cy = np.float32(100.0) + np.float32(mv[1]) * np.float32(f)
            out.append((cx, cy, 5.0, lo[0], lo[1], hi[0], hi[1]))
    return [tuple(np.float32(v) for v in c) for c in out]


def oracle_lines(cs):
    from oracle import oracle as O
    O.build()
    import ctypes as C
    L = O.lib()

    class V2(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float)]

    class Circle(C.Structure):
        _fields_ = [("center", V2), ("radius", C.c_float)]

    class Aabb(C.Structure):
        _fields_ = [("min", V2), ("max", V2)]

    class Contact(C.Structure):
        _fields_ = [("some", C.c_int), ("dist", C.c_float), ("normal1", V2), ("normal2", V2)]

    L.orc_contact_test_circle_aabb.restype = Contact
    L.orc_contact_test_circle_aabb.argtypes = [C.POINTER(Circle), C.POINTER(Aabb), C.POINTER(C.c_uint32)]
    lines = []
    for cx, cy, r, x0, y0, x1, y1 in cs:
        err = C.c_uint32(0)
        c = L.orc_contact_test_circle_aabb(C.byref(Circle(V2(cx, cy), r)), C.byref(Aabb(V2(x0, y0), V2(x1, y1))), C.byref(err))
        if c.some:
            lines.append("1 " + " ".join(f32hex(v) for v in (c.dist, c.normal1.x, c.normal1.y, c.normal2.x, c.normal2.y)))
        else:
            lines.append("0")
    return lines


def read_inputs(path):
    out = []
    for line in open(path):
        if line.strip():
            out.append(tuple(np.uint32(int(t, 16)).view(np.float32) for t in line.split()))
    return out


if __name__ == "__main__":
    cs = cases()
    with open(os.path.join(HERE, "contact_inputs_v1.txt"), "w") as f:
        for c in cs:
            f.write(" ".join(f32hex(v) for v in c) + "\n")
    with open(os.path.join(HERE, "contact_oracle_v1.txt"), "w") as f:
        f.write("\n".join(oracle_lines(cs)) + "\n")
    print(len(cs), "cases")
