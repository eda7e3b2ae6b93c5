"""Compares a trace printed by bindings/rust/trace-dumper (the REAL reference mechanics, run wherever cargo exists) with
the CPU oracle on the same explicit inputs. Reports the first differing step, or that the oracle is pinned on this trace.
Usage: python tests/compare_rust_trace.py rust_trace.txt <dir_x> actions.txt
Also: python tests/compare_rust_trace.py --emit <dir_x> actions.txt   prints the oracle's trace in the same format.
      python tests/compare_rust_trace.py --contacts rust_contacts.txt tests/golden/contact_inputs_v1.txt
          compares parry2d's ball / cuboid contact query (trace-dumper --contacts) with the oracle's restatement, bit for bit
      python tests/compare_rust_trace.py --emit-contacts tests/golden/contact_inputs_v1.txt   prints the oracle's answers."""
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


def contact_lines(inputs_path):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_contact_inputs as M
    return M.oracle_lines(M.read_inputs(inputs_path))


def main():
    if sys.argv[1] == "--emit-contacts":
        print("\n".join(contact_lines(sys.argv[2])))
        return
    if sys.argv[1] == "--contacts":
        rust = [l.strip() for l in open(sys.argv[2]) if l.strip()]
        mine = contact_lines(sys.argv[3])
        bad = [i for i, (r, m) in enumerate(zip(rust, mine)) if r != m]
        if len(rust) != len(mine):
            print("line counts differ: reference %d, oracle %d" % (len(rust), len(mine)))
            sys.exit(1)
        if bad:
            i = bad[0]
            print("%d of %d contact queries differ; first at input line %d:\n  reference: %s\n  oracle:    %s" % (len(bad), len(mine), i + 1, rust[i], mine[i]))
            sys.exit(1)
        print("oracle == parry2d on all %d contact queries (bit for bit)" % len(mine))
        return
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
