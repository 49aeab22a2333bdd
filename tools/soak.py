#!/usr/bin/env python
"""Randomised parity soak (development aid): CUDA env.step / replay vs the CPU oracle on bigger batches, longer
rollouts and more seeds than the test-suite uses.   python tools/soak.py [--seeds 4] [--envs 384] [--steps 150]"""
import argparse
import dataclasses
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import helpers as H
from jaxmarl_hft_b200 import config as C
from oracle import lob_oracle


def rollout(oracle, mac, day, B, steps, seed, threads):
    ld = H.load_for(mac, day)
    ref = H.OracleEnv(oracle, mac, ld, B)
    gpu = H.CudaEnv(mac, ld, B, ref.params)
    rng = np.random.default_rng(seed)
    H.draw_prng(rng, ref.cfg, ref.arrays)
    gpu.set_inputs(ref.arrays)
    ref.reset(); gpu.reset()
    H.assert_arrays_match(ref.arrays, gpu.numpy(), ref.cfg, float_exact=True)
    for s in range(steps):
        H.draw_prng(rng, ref.cfg, ref.arrays)
        H.draw_actions(rng, ref.cfg, ref.arrays)
        if s % 5 == 2:
            for t in range(ref.cfg.n_agent_types):
                a = ref.arrays[f"actions{t}"]
                a[::3] = rng.integers(-4, 50, size=a[::3].shape)
        gpu.set_inputs(ref.arrays)
        ref.step(n_threads=threads); gpu.step()
        try:
            H.assert_arrays_match(ref.arrays, gpu.numpy(), ref.cfg, float_exact=True)   # bit patterns, floats included
        except AssertionError as e:
            raise AssertionError(f"seed {seed} step {s}: {e}") from None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--steps", type=int, default=140)
    a = ap.parse_args()
    oracle = lob_oracle.load()
    threads = oracle.max_threads()
    cases = []
    base = H.load_mac("2_player_fq_fqc")
    cases.append(("2_player", base, dict(n_events=60000)))
    cases.append(("2_player stress small book", H.load_mac("2_player_fq_fqc", nOrders=40, nTrades=24), dict(seed=9, n_events=60000, stress=True)))
    cases.append(("hetero deep", H.load_mac("hetero_deep_book"), dict(n_events=60000)))
    cases.append(("2_player 200-row book", H.load_mac("2_player_fq_fqc", nOrders=200, nTrades=150), dict(n_events=60000)))
    # deep books on the capacity-stress day: the books outgrow the 128-row shared-memory window -> both passes of the step
    cases.append(("hetero deep, stress day (2 passes)", H.load_mac("hetero_deep_book"), dict(seed=9, n_events=60000, stress=True)))
    cases.append(("2_player 300-row book, stress day", H.load_mac("2_player_fq_fqc", nOrders=300, nTrades=64), dict(seed=9, n_events=60000, stress=True)))
    cases.append(("cancel mode 3 + MKT", H.load_mac("2_player_fq_fqc", nOrders=48, nTrades=20, cancel_mode=3, type_4_interpretation=2),
                  dict(seed=9, n_events=60000, stress=True)))
    cases.append(("fixed_time", H.load_mac("2_player_fq_fqc", ep_type="fixed_time", episode_time=900, start_resolution=300), dict(n_events=60000)))
    ag = dict(base.dict_of_agents_configs)
    cases.append(("10+10 agents", H.with_agents(base, ag, [10, 10]), dict(n_events=60000)))
    for name, mac, dk in cases:
        day = H.small_day(**dk)
        for seed in range(a.seeds):
            t0 = time.time()
            B = a.envs if "10+10" not in name and "hetero" not in name else max(32, a.envs // 4)
            rollout(oracle, mac, day, B, a.steps, 1000 + seed, threads)
            print(f"ok  {name:32s} seed {seed}  {B} envs x {a.steps} steps  {time.time() - t0:.1f}s", flush=True)
    # replay: long adversarial + random streams on several shapes
    for no, nt, t4, cm in ((100, 100, 0, 1), (64, 16, 2, 3), (200, 64, 1, 2), (512, 256, 0, 1), (7, 3, 0, 3)):
        bc = C.book_config(C.World_EnvironmentConfig(nOrders=no, nTrades=nt, type_4_interpretation=t4, cancel_mode=cm))
        for seed in range(a.seeds):
            rng = np.random.default_rng(7000 + seed)
            Bk, T = 256, 3000
            msgs = np.where((rng.random(Bk * T) < 0.3)[:, None], H.adversarial_messages(rng, Bk * T, bc),
                            H.random_messages(rng, Bk * T, bc, price_lo=99_000, price_hi=100_800)).astype(np.int32)
            cu = rng.integers(0, 2 ** 23, size=(Bk, T, 2)).astype(np.float32) / np.float32(2 ** 23) if cm >= 2 else None
            start = np.arange(Bk, dtype=np.int64) * T
            a0 = np.full((Bk, no, 6), -1, np.int32); b0 = a0.copy(); t0_ = np.full((Bk, nt, 8), -1, np.int32)
            ra, rb, rt = a0.copy(), b0.copy(), t0_.copy()
            oracle.replay(bc, ra, rb, rt, msgs, start, T, n_threads=threads, cancel_u=cu)
            ga, gb, gt = H.cuda_replay(bc, a0, b0, t0_, msgs, start, T, cancel_u=cu)
            assert (ga == ra).all() and (gb == rb).all() and (gt == rt).all(), (no, nt, t4, cm, seed)
            print(f"ok  replay no={no} nt={nt} t4={t4} cancel_mode={cm} seed {seed}", flush=True)
    print("soak ok")


if __name__ == "__main__":
    main()
