#!/usr/bin/env python
"""Quick kernel-only timing of the replay and step kernels (development aid; bench.py is the contract).
    LOB_SO=/path/to/variant.so python tools/kbench.py [--books 16384] [--iters 5]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import bench as Bn
import helpers as H  # noqa: E402  (development aid: shares the tests' config helpers)
from jaxmarl_hft_b200 import _lib, config as C, env as E, states


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--books", type=int, default=16384)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--config", default="2_player_fq_fqc")
    ap.add_argument("--skip-replay", action="store_true")
    ap.add_argument("--agents", default="", help="agents per type override, e.g. 10,10")
    ap.add_argument("--norders", type=int, default=0, help="book rows per side override")
    ap.add_argument("--replay-msgs", type=int, default=6400, help="messages per book per replay launch (bench.py: 38400)")
    ap.add_argument("--grouped", action="store_true", help="replay through the 4-books-per-warp measurement variant")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    L = _lib.lib()
    mac = H.load_mac(a.config, **({"nOrders": a.norders} if a.norders else {}))
    if a.agents:
        mac = H.with_agents(mac, dict(mac.dict_of_agents_configs), [int(x) for x in a.agents.split(",")])
    ld = Bn._load_day(mac)
    bc = C.book_config(mac.world_config)
    base_env = E.BaseLOBEnv(mac.world_config, loaded=ld, device=dev)
    P = base_env._params_np
    M, W, B = ld.msgs.shape[0], ld.starts.shape[0], a.books
    widx = np.arange(B) % W
    out = {"so": os.path.basename(_lib.SO_PATH)}
    if not a.skip_replay:
        msgs_d = torch.from_numpy(ld.msgs).to(dev)
        asks = torch.from_numpy(P["init_asks"][widx]).to(dev); bids = torch.from_numpy(P["init_bids"][widx]).to(dev)
        trades = torch.from_numpy(P["init_trades"][widx]).to(dev)
        base = ld.starts[widx].astype(np.int64) + (np.arange(B) // W) % 100
        TW = a.replay_msgs
        starts = torch.from_numpy(np.stack([(base + k * TW) % (M - TW) for k in range(a.iters + 2)])).to(dev)
        for k in range(2):
            E.replay_books(bc, asks, bids, trades, msgs_d, starts[k], TW, grouped=a.grouped)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
        for k in range(a.iters):
            ev[k][0].record(); E.replay_books(bc, asks, bids, trades, msgs_d, starts[2 + k], TW, grouped=a.grouped); ev[k][1].record()
        torch.cuda.synchronize()
        ms = float(np.mean([x.elapsed_time(y) for x, y in ev]))
        out["replay_ms"] = round(ms, 3); out["replay_msgs_per_s"] = f"{B * TW / ms * 1e3:.3e}"
    env = E.MARLEnv(None, mac, num_envs=B, loaded=ld, device=dev, seed=1)
    envp = env.default_params
    obs, state = env.reset(None, envp)
    T = env.cfg.n_agent_types
    g = torch.Generator(device=dev); g.manual_seed(5)
    acts = [torch.randint(0, env.action_spaces[t].n, (B, env.cfg.agent[t].n_agents), generator=g, device=dev, dtype=torch.int32) for t in range(T)]
    for k in range(3):
        env.step(None, state, acts, envp)
    bufs = states.pack_buffers(env.cfg, state.arrays, env.base_env.device_params())
    st = _lib.current_stream_ptr()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for k in range(a.iters):
        ev[k][0].record(); _lib.check(L.lob_step_launch(ctypes.byref(env.cfg), ctypes.byref(bufs), B, st), "step"); ev[k][1].record()
    torch.cuda.synchronize()
    ms = float(np.mean([x.elapsed_time(y) for x, y in ev]))
    out["step_ms"] = round(ms, 4); out["env_steps_per_s"] = f"{B / ms * 1e3:.3e}"
    print(out)


if __name__ == "__main__":
    main()
