"""Compares a trace printed by bindings/rust/trace-dumper (the REAL reference mechanics, run wherever cargo exists) with
the CPU oracle on the same explicit inputs. Reports the first differing step, or that the oracle is pinned on this trace.
Usage: python tools/compare_rust_trace.py rust_trace.txt <dir_x> actions.txt
Also: python tools/compare_rust_trace.py --emit <dir_x> actions.txt   prints the oracle's trace in the same format."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402


def oracle_trace(dir_x, actions):
    env = O.VecEnv(1, seed=0)
    env.reset_env(0, float(np.float32(dir_x)))
    lines = []
    for step, a in enumerate(actions):
        _, d = env.step(np.array([a], dtype=np.uint8))
        # the vectorised driver restarts finished episodes; the last pre-reset state is not observable afterwards, so
        # stop at the step before `finished` (the dumper stops there as well)
        if d[0]:
            lines.append("%d finished" % step)
            break
        s = env.state()
        bits = [int(s[k].view(np.uint32)[0]) for k in ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed")]
        lines.append("%d %s %d %d %d" % (step, " ".join("%08x" % b for b in bits), bin(int(s["bricks"][0])).count("1"), int(s["score"][0]), 0))
    return lines


def main():
    if sys.argv[1] == "--emit":
        acts = [int(t) for t in open(sys.argv[3]).read().split()]
        print("\n".join(oracle_trace(float(sys.argv[2]), acts)))
        return
    rust = [l.strip() for l in open(sys.argv[1]) if l.strip()]
    acts = [int(t) for t in open(sys.argv[3]).read().split()]
    mine = oracle_trace(float(sys.argv[2]), acts)
    for i, (r, m) in enumerate(zip(rust, mine)):
        if m.endswith("finished"):
            ok = r.split()[-1] == "1"
            print("both finished at step %d" % i if ok else "oracle finished at step %d, reference did not" % i)
            return
        if r != m:
            print("first difference at step %d:\n  reference: %s\n  oracle:    %s" % (i, r, m))
            sys.exit(1)
    print("oracle == reference on all %d compared steps" % min(len(rust), len(mine)))


if __name__ == "__main__":
    main()
