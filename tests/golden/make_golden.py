"""Generates tests/golden/oracle_trace_v1.json from the CPU oracle (BASELINE configs[0]: single Breakout env, fixed seed,
random policy, 10k steps, parity trace) plus a small multi-env trace with replay samples.

The Rust reference cannot be run here (no cargo/rustc; its renderer is unimplemented!()), so these vectors pin the
ORACLE (and, through the GPU tests, the CUDA path) against regressions; they are not outputs of the reference.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

STATE_KEYS = ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed", "bricks", "score", "episode_step")


def trace(n_envs, seed, steps, replay_capacity=0, checkpoints=()):
    env = O.VecEnv(n_envs, seed=seed, replay_capacity=replay_capacity)
    acts = O.synthetic_actions(seed, 0, n_envs, 0, steps)
    h_state, h_frames, h_rd = hashlib.sha256(), hashlib.sha256(), hashlib.sha256()
    out = {"n_envs": n_envs, "seed": seed, "steps": steps, "replay_capacity": replay_capacity, "checkpoints": {}}
    for t in range(steps):
        r, d = env.step(acts[t])
        h_rd.update(r.tobytes()); h_rd.update(d.tobytes())
        st = env.state()
        for k in STATE_KEYS:
            h_state.update(st[k].tobytes())
        h_frames.update(env.obs_u8().tobytes())
        if t + 1 in checkpoints:
            out["checkpoints"][str(t + 1)] = {k: [int(x) for x in st[k].view(np.uint32 if st[k].dtype == np.float32 else st[k].dtype)] for k in STATE_KEYS}
    out["sha256_state_bits_every_step"] = h_state.hexdigest()
    out["sha256_obs_u8_every_step"] = h_frames.hexdigest()
    out["sha256_reward_done_every_step"] = h_rd.hexdigest()
    out["stats"] = {k: float(v) for k, v in env.stats().items()}
    out["err_or"] = int(np.bitwise_or.reduce(env.state()["err"]))
    if replay_capacity:
        ln = env.replay_len()
        idx = O.sample_distinct(seed, 0, ln, 32)
        g = env.get_many(idx, "u8")
        out["replay"] = {"len": int(ln), "indices_call0_batch32": [int(x) for x in idx],
                         "sha256_state_u8": hashlib.sha256(g["state"].tobytes()).hexdigest(),
                         "sha256_next_u8": hashlib.sha256(g["state_next"].tobytes()).hexdigest(),
                         "reward": [float(x) for x in g["reward"]], "action": [int(x) for x in g["action"]], "done": [int(x) for x in g["done"]],
                         "sha256_state_f32": hashlib.sha256(env.get_many(idx, "f32")["state"].tobytes()).hexdigest()}
    env.close()
    return out


def tracking_actions(state, n_envs, t):
    """Deterministic paddle-follows-ball policy with an env/time dependent aim offset (keeps episodes alive, so that
    paddle bounces, wall bounces, multi-brick contacts and the bisection paths are exercised)."""
    centre = (state["pad_min_x"] + state["pad_max_x"]) / 2
    off = (((np.arange(n_envs) * 37 + t * 11) % 51) - 25).astype(np.float32)
    target = state["ball_cx"] + off
    return np.where(target < centre - 4, 1, np.where(target > centre + 4, 2, 0)).astype(np.uint8)


def tracking_trace(n_envs, seed, steps):
    env = O.VecEnv(n_envs, seed=seed)
    h_state, h_act, h_frames = hashlib.sha256(), hashlib.sha256(), hashlib.sha256()
    for t in range(steps):
        a = tracking_actions(env.state(), n_envs, t)
        h_act.update(a.tobytes())
        env.step(a)
        st = env.state()
        for k in STATE_KEYS:
            h_state.update(st[k].tobytes())
        if t % 16 == 15:
            h_frames.update(env.obs_u8().tobytes())
    st = env.state()
    out = {"n_envs": n_envs, "seed": seed, "steps": steps, "sha256_actions": h_act.hexdigest(), "sha256_state_bits_every_step": h_state.hexdigest(),
           "sha256_obs_u8_every_16_steps": h_frames.hexdigest(), "stats": {k: float(v) for k, v in env.stats().items()},
           "final_score": [int(x) for x in st["score"]], "final_bricks": [int(x) for x in st["bricks"]], "err_or": int(np.bitwise_or.reduce(st["err"]))}
    env.close()
    return out


def main():
    O.build(force=True)
    gold = {
        "what": "CPU-oracle traces (not reference outputs): see make_golden.py",
        "single_env_10k": trace(1, 20261018, 10000, checkpoints=(1, 100, 1000, 10000)),
        "multi_env_replay": trace(24, 7, 400, replay_capacity=24 * 64, checkpoints=(400,)),
        "tracking_policy": tracking_trace(16, 5, 4000),
        "dir_x": {"bits_0": float(O.lib().orc_dir_x_from_bits(0)), "env3_ep5_seed9": float(O.lib().orc_reset_dir_x(9, 3, 5))},
        "sample_distinct": {"seed1_call0_len100_b50": [int(x) for x in O.sample_distinct(1, 0, 100, 50)],
                            "seed1_call5_len33_b32": [int(x) for x in O.sample_distinct(1, 5, 33, 32)]},
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_trace_v1.json")
    with open(path, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
