#!/usr/bin/env python
"""The reference's speed test (gymnax_exchange/jaxen/Speed_test.py:165-224) on this package's MARLEnv: NUM_ENVS
environments, a rollout of NUM_STEPS steps with uniformly sampled actions drawn ON the device every step, the rollout
run once untimed ("compile") and once timed; steps/s = NUM_ENVS * NUM_STEPS / wall time of the second rollout.
    python tools/speed_test.py [--config 2_player_fq_fqc] [--envs 4000] [--steps 50]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import jaxmarl_hft_b200 as lob
from jaxmarl_hft_b200 import env as E


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="2_player_fq_fqc")
    ap.add_argument("--envs", type=int, default=4000)       # Speed_test.py:70
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    mac = lob.load_config_from_file(os.path.join(ROOT, "jaxmarl-hft_b200", "configs", a.config + ".json"))
    t0 = time.time()
    env = E.MARLEnv(None, mac, num_envs=a.envs, device="cuda:0", seed=0,
                    synth={"cache_dir": os.path.join(ROOT, ".cache")})
    params = env.default_params
    t1 = time.time()
    obs, state = env.reset(None, params)
    torch.cuda.synchronize()
    reset_time = time.time() - t1
    torch.manual_seed(0)     # the default CUDA generator is graph-safe (philox offsets advance per replay)
    n_per = mac.number_of_agents_per_type

    def policy(k, obs):          # space.sample() for every agent of every env, on the device
        return [torch.randint(0, sp.n, (a.envs, n), device="cuda:0", dtype=torch.int32)
                for sp, n in zip(env.action_spaces, n_per)]

    graph, traj = env.capture_rollout(state, policy, a.steps, params)     # "compile": capture (+ one warm-up step)
    graph.replay()
    torch.cuda.synchronize()
    start = time.time()
    graph.replay()
    torch.cuda.synchronize()
    dt = time.time() - start
    total = a.steps * a.envs
    print(f"Total Envs:           {a.envs}\nEnv construction:     {t1 - t0:.4f} seconds\nReset time:           {reset_time:.4f} seconds\n"
          f"Rollout time:         {dt:.4f} seconds ({a.steps} steps)\nAvg time per step:    {dt / total:.3e} seconds\n"
          f"Avg steps per second: {total / dt:.4e}\nMessages per second:  {total * env.num_msgs_per_step / dt:.4e}")


if __name__ == "__main__":
    main()
