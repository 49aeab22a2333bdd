#!/usr/bin/env python
"""Benchmark of the LOB hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Headline workload = BASELINE.json configs[1]: pure book replay (BaseLOBEnv.step_env / job.scan_through_entire_array)
of a synthetic LOBSTER day, nOrders = nTrades = 100, 16384 independent books per GPU, each step = ONE launch that
scans a 6400-message window (64 base-env steps of 100 messages) for every book.  The same run also times the
multi-agent ``env.step`` (configs[3] shapes: 2_player_fq_fqc, 16384 envs per GPU) and reports it under "env_step".

One JSON line on stdout (rank 0).  ``--impl reference`` times the CPU restatement of the reference (oracle/, OpenMP
over books; the reference itself is JAX and cannot be installed offline) on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WINDOW = 6400          # messages per book per replay step
ND = 100               # data messages per base-env step
DAY_EVENTS = 400_000
DAY_SEED = 20220103
METRIC = "lob_messages_per_sec"
UNIT = "msgs/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(kernel):
    """DRAM bytes per launch of ``kernel`` from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f).get(kernel)
        if d:
            return d["bytes"]
    return None


def _load_day(mac):
    from jaxmarl_hft_b200 import lobster
    cache = os.path.join(ROOT, ".cache")
    return lobster.load_or_generate(mac.world_config, seed=DAY_SEED, n_events=DAY_EVENTS, cache_dir=cache)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _dist():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return ws, rank, local


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's algorithm on the host cores: oracle/lob_oracle.c (a restatement -- JAX is not installable
    offline), OpenMP over books, all host threads, on a bounded sample of the replay workload."""
    ws, rank, _ = _dist()
    if rank != 0:
        return
    from jaxmarl_hft_b200 import config as C, env as E
    from oracle import harness as H, lob_oracle
    oracle = lob_oracle.load()
    mac = C.load_named_config("2_player_fq_fqc")
    ld = _load_day(mac)
    bc = C.book_config(mac.world_config)
    params = E.build_reset_params(ld, mac.world_config, H.oracle_replay_fn(oracle, bc))
    threads = oracle.max_threads()
    B = args.ref_books
    W = ld.starts.shape[0]
    widx = np.arange(B) % W
    asks, bids, trades = params["init_asks"][widx].copy(), params["init_bids"][widx].copy(), params["init_trades"][widx].copy()
    M = ld.msgs.shape[0]
    base = ld.starts[widx].astype(np.int64)
    times = []
    for k in range(args.warmup + args.steps):
        start = (base + k * WINDOW) % (M - WINDOW)
        t0 = time.perf_counter()
        oracle.replay(bc, asks, bids, trades, ld.msgs, start, WINDOW, n_threads=threads)
        dt = time.perf_counter() - t0
        if k >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = B * WINDOW * len(times) / total
    sample = f"{B} books x {WINDOW} msgs per step, {len(times)} steps, OpenMP over books"
    # the same for MARLEnv.step (2_player_fq_fqc): the oracle's step over 1024 environments, all host threads
    nenv = 1024
    oenv = H.OracleEnv(oracle, mac, ld, nenv)
    rng = np.random.default_rng(0)
    H.draw_prng(rng, oenv.cfg, oenv.arrays)
    oenv.reset()
    st = []
    for k in range(args.warmup + args.steps):
        H.draw_prng(rng, oenv.cfg, oenv.arrays)
        H.draw_actions(rng, oenv.cfg, oenv.arrays)
        t0 = time.perf_counter()
        oenv.step(n_threads=threads)
        if k >= args.warmup:
            st.append(time.perf_counter() - t0)
    env_value = nenv * len(st) / sum(st)
    return json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "pure book replay, synthetic LOBSTER day, nOrders=100, nTrades=100 (BASELINE configs[1]), "
                               "bounded sample", "books": B, "msgs_per_book_per_step": WINDOW},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "env_step": {"metric": "env_steps_per_sec", "value": env_value, "unit": "env-steps/s",
                     "sample": f"{nenv} envs x {len(st)} steps of MARLEnv.step 2_player_fq_fqc, oracle, OpenMP over envs"},
    })


# --------------------------------------------------------------------------------------------------- native arm
def _step_bytes(env, No, Nt, N, Nd):
    """Algorithmic bytes per env-step: books in+out, data slice in, trades out, per-message bests out,
    scalars / agent state / actions / perm in+out, obs / reward / done / info out."""
    cfg = env.cfg
    T = cfg.n_agent_types
    n_mm = sum(cfg.agent[t].n_agents for t in range(T) if cfg.agent[t].kind == 0)
    n_ex = sum(cfg.agent[t].n_agents for t in range(T) if cfg.agent[t].kind == 1)
    n_ag = n_mm + n_ex
    S_r = 64 + 20 * n_mm + 52 * n_ex + 4 * n_ag + 4 * env.num_action_msgs_per_step_by_all_agents
    d_obs = sum(cfg.agent[t].n_agents * env.observation_spaces[t].shape[0] for t in range(T))
    S_w = 64 + 20 * n_mm + 52 * n_ex + 4 * d_obs + 4 * n_ag + n_ag + 1 + 4 * (15 + 24 * n_mm + 9 * n_ex)
    return (2 * No * 24 + Nd * 32 + S_r) + (2 * No * 24 + Nt * 32 + 2 * N * 8 + S_w)


def run_native(args):
    import torch
    import torch.distributed as dist
    from jaxmarl_hft_b200 import _lib, abi, config as C, dist as D, env as E, states

    ws, rank, local = _dist()
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if ws > 1:   # (NCCL_DEBUG is left as the caller set it: main() keeps fd 1 clean for the one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()          # raises when csrc/liblobstep.so is missing: there is no fallback
    hbm_peak, peak_src = _peaks()

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if ws == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, n):
        """n calls of fn bracketed by barrier + synchronize on both sides; device time, max over ranks (ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n):
            fn(k)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    mac = C.load_named_config("2_player_fq_fqc")
    ld = _load_day(mac)
    bc = C.book_config(mac.world_config)
    No, Nt = bc.n_orders, bc.n_trades
    base_env = E.BaseLOBEnv(mac.world_config, loaded=ld, device=dev)   # reset states through the CUDA replay kernel
    params_np = base_env._params_np
    M, W = ld.msgs.shape[0], ld.starts.shape[0]
    B = args.books
    K, Wm = args.steps, args.warmup
    TW = WINDOW * args.windows_per_step     # messages per book per step (ONE launch)
    # every rank replays its own shard of books: book g = rank * B + i starts in window g % W (no collective)
    gidx = rank * B + np.arange(B)
    widx = gidx % W
    msgs_d = torch.from_numpy(ld.msgs).to(dev)
    base = ld.starts[widx].astype(np.int64) + (gidx // W) % ND
    starts_np = np.stack([(base + k * TW) % (M - TW) for k in range(Wm + K)])

    def fresh_books():
        return (torch.from_numpy(params_np["init_asks"][widx]).to(dev), torch.from_numpy(params_np["init_bids"][widx]).to(dev),
                torch.from_numpy(params_np["init_trades"][widx]).to(dev))

    # ---- (1) replay, inputs resident in HBM: K steps, one launch each ----
    asks, bids, trades = fresh_books()
    starts_d = torch.from_numpy(starts_np).to(dev)
    for k in range(Wm):
        E.replay_books(bc, asks, bids, trades, msgs_d, starts_d[k], TW)
    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    L.lob_launch_count_reset()
    barrier()
    if rank == 0:
        sampler.start()

    def replay_step(k):
        ev[k][0].record()
        E.replay_books(bc, asks, bids, trades, msgs_d, starts_d[Wm + k], TW)
        ev[k][1].record()

    total_ms = timed(replay_step, K)
    launches = int(L.lob_launch_count())
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    value = ws * B * TW * K / (total_ms * 1e-3)
    bytes_per_launch = B * (2 * (2 * No * 24) + 2 * Nt * 32 + 32 * TW + 8)
    achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9
    del asks, bids, trades

    # ---- (2) replay END TO END with HOST buffers through the C ABI (lob_host_replay_run): every step copies the books,
    #          trade logs and start offsets from pinned host memory, scans, and copies books + trade logs back ----
    h_asks = torch.from_numpy(params_np["init_asks"][widx]).pin_memory()
    h_bids = torch.from_numpy(params_np["init_bids"][widx]).pin_memory()
    h_trades = torch.from_numpy(params_np["init_trades"][widx]).pin_memory()
    h_starts = torch.from_numpy(starts_np).pin_memory()
    msgs_host = np.ascontiguousarray(ld.msgs, np.int32)
    handle = L.lob_host_replay_create(ctypes.byref(bc), B, M, local)
    if not handle:
        raise _lib.LobError(f"lob_host_replay_create: {L.lob_last_error().decode()}")
    _lib.check(L.lob_host_replay_set_messages(handle, msgs_host.ctypes.data_as(abi.p_i32), M), "lob_host_replay_set_messages")
    h2d, d2h = ctypes.c_int64(0), ctypes.c_int64(0)
    pi32 = lambda t: ctypes.cast(t.data_ptr(), abi.p_i32)

    def host_step(k):   # blocking: H2D + kernel + D2H inside the call
        _lib.check(L.lob_host_replay_run(handle, pi32(h_asks), pi32(h_bids), pi32(h_trades),
                                         ctypes.cast(h_starts[k].data_ptr(), abi.p_i64), TW, B,
                                         ctypes.byref(h2d), ctypes.byref(d2h)), "lob_host_replay_run")

    for k in range(Wm):
        host_step(k)
    e2e_ms = timed(lambda k: host_step(Wm + k), K)
    e2e_value = ws * B * TW * K / (e2e_ms * 1e-3)
    L.lob_host_replay_destroy(handle)
    # the same with the state device-resident as in the reference (only start offsets in, best bid / ask per book out)
    asks, bids, trades = fresh_books()
    start_dev = torch.empty(B, dtype=torch.int64, device=dev)
    best_dev = torch.empty((B, 4), dtype=torch.int32, device=dev)
    best_host = torch.empty((B, 4), dtype=torch.int32).pin_memory()

    def resident_step(k):
        start_dev.copy_(h_starts[k], non_blocking=True)
        base_env.replay(asks, bids, trades, start_dev, TW, best_out=best_dev)
        best_host.copy_(best_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for k in range(Wm):
        resident_step(k)
    resident_ms = timed(lambda k: resident_step(Wm + k), K)
    resident_value = ws * B * TW * K / (resident_ms * 1e-3)
    del asks, bids, trades, h_asks, h_bids, h_trades

    # ---- (3) multi-agent env.step (2_player_fq_fqc), weak scaling: args.envs per GPU; one bench step = `inner` launches ----
    def step_leg(mac_x, ld_x, n_envs, inner, seed):
        """Kernel-only env.step: K x inner back-to-back lob_step_launch on n_envs environments (PRNG products and actions
        of the last draw stay resident).  Returns (env, state, total ms, mean ms per launch, launches)."""
        env = E.MARLEnv(None, mac_x, num_envs=n_envs, loaded=ld_x, device=dev, seed=seed + rank)
        envp = env.default_params
        _, state = env.reset(None, envp)
        Tn = env.cfg.n_agent_types
        g = torch.Generator(device=dev); g.manual_seed(99 + seed + rank)
        acts = [torch.randint(0, env.action_spaces[t].n, (n_envs, env.cfg.agent[t].n_agents), generator=g, device=dev,
                              dtype=torch.int32) for t in range(Tn)]
        for k in range(Wm):
            env.step(None, state, acts, envp)
        bufs = states.pack_buffers(env.cfg, state.arrays, env.base_env.device_params())
        stream = _lib.current_stream_ptr(dev)
        L.lob_launch_count_reset()

        def one(k):
            for _ in range(inner):
                _lib.check(L.lob_step_launch(ctypes.byref(env.cfg), ctypes.byref(bufs), n_envs, stream), "lob_step_launch")

        ms = timed(one, K)
        return env, state, acts, ms, ms / (K * inner), int(L.lob_launch_count())

    inner = args.env_inner
    env, state, acts_d0, step_ms, step_kern_ms, step_launches = step_leg(mac, ld, args.envs, inner, 1234)
    clocks = sampler.stop() if rank == 0 else None      # sampled over the replay, host-replay and env.step regions
    envp = env.default_params
    T = env.cfg.n_agent_types
    N = env.num_msgs_per_step
    step_value = ws * args.envs * K * inner / (step_ms * 1e-3)
    step_bytes = _step_bytes(env, No, Nt, N, ND)
    step_achieved = step_bytes * args.envs / (step_kern_ms * 1e-3) / 1e9
    n_act_space = [env.action_spaces[t].n for t in range(T)]
    n_i = [env.cfg.agent[t].n_agents for t in range(T)]
    g = torch.Generator(device=dev); g.manual_seed(99 + rank)
    acts_d = [[torch.randint(0, n_act_space[t], (args.envs, n_i[t]), generator=g, device=dev, dtype=torch.int32)
               for t in range(T)] for _ in range(4)]
    # env.step end to end through MARLEnv: actions from pinned host memory, PRNG draws, obs / reward / done to the host
    acts_h = [[a.cpu().pin_memory() for a in row] for row in acts_d]
    acts_in = [torch.empty_like(a) for a in acts_d[0]]
    outs_h = None

    def env_e2e(k):
        nonlocal outs_h
        for t in range(T):
            acts_in[t].copy_(acts_h[k % 4][t], non_blocking=True)
        o, st, r, d, info = env.step(None, state, acts_in, envp)
        outs = list(o) + list(r) + [state.arrays["done_all"]]
        if outs_h is None:
            outs_h = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in outs]
        for h, x in zip(outs_h, outs):
            h.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for k in range(Wm):
        env_e2e(k)
    # the same step as ONE CUDA graph launch (H2D actions from pinned memory, draw, step, D2H results)
    acts_pin = [torch.empty_like(a).pin_memory() for a in acts_h[0]]

    def pre():
        for t in range(T):
            acts_in[t].copy_(acts_pin[t], non_blocking=True)

    def post(out):
        o, st, r, d, info = out
        for h, x in zip(outs_h, list(o) + list(r) + [state.arrays["done_all"]]):
            h.copy_(x, non_blocking=True)

    graph, _ = env.capture_step(state, acts_in, envp, pre=pre, post=post)

    def env_e2e_graph(k):
        for t in range(T):
            acts_pin[t].copy_(acts_h[k % 4][t])     # this step's actions arrive in pinned host memory
        graph.replay()
        torch.cuda.current_stream().synchronize()

    for k in range(Wm):
        env_e2e_graph(k)
    n_e2e = K * max(1, inner // 8)
    step_e2e_value = ws * args.envs * n_e2e / (timed(env_e2e_graph, n_e2e) * 1e-3)
    step_e2e_eager_value = ws * args.envs * n_e2e / (timed(env_e2e, n_e2e) * 1e-3)
    step_h2d = sum(a.numel() * 4 for a in acts_in)
    step_d2h = sum(h.numel() * h.element_size() for h in outs_h)
    # a 64-step rollout as ONE CUDA graph (MARLEnv.capture_rollout: the trainer's jit(scan(vmap(env.step))), state resident
    # for the whole rollout; pre-sampled actions as the policy, Speed_test.py:165-214), trajectory read back per rollout,
    # then the episode statistics of the rollout reduced over the ranks (the path's ONE collective: NCCL all-reduce of
    # O(10) floats, ippo_rnn_JAXMARL_pmap.py:566-567 does the same for gradients)
    RT = 64
    pol = [torch.randint(0, n_act_space[t], (RT, args.envs, n_i[t]), generator=g, device=dev, dtype=torch.int32)
           for t in range(T)]
    rgraph, traj = env.capture_rollout(state, lambda k, obs: [pol[t][k] for t in range(T)], RT, envp)
    traj_dev = list(traj["obs"]) + list(traj["reward"]) + [traj["done"]]
    traj_host = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in traj_dev]
    n_roll = max(2, K // 4)
    red_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_roll + 1)]
    stats_out = {}

    def rollout_once(k):
        rgraph.replay()
        for h, x in zip(traj_host, traj_dev):
            h.copy_(x, non_blocking=True)
        local_stats = torch.stack([D.local_episode_stats(traj["reward"][t]) for t in range(T)])   # [T, 5] on the device
        red_ev[k][0].record()
        stats_out.update(D.reduce_episode_stats(local_stats))
        red_ev[k][1].record()
        torch.cuda.current_stream().synchronize()

    rollout_once(n_roll)
    roll_ms = timed(rollout_once, n_roll)
    rollout_value = ws * args.envs * RT * n_roll / (roll_ms * 1e-3)
    rollout_d2h = sum(h.numel() * h.element_size() for h in traj_host)
    reduce_us = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in red_ev[:n_roll]]))) * 1e3
    stats_count = float(stats_out["count"].sum().item())
    # the same 64-step rollout as ONE kernel launch (MARLEnv.rollout -> lob_rollout_launch: the books stay in shared memory
    # for the whole rollout), PRNG draws included, trajectory read back
    def rollout_kernel_once(k):
        tk, _ = env.rollout(state, pol, RT, envp)
        for h, x in zip(traj_host, list(tk["obs"]) + list(tk["reward"]) + [tk["done"]]):
            h.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    rollout_kernel_once(0)
    rollout_kernel_value = ws * args.envs * RT * n_roll / (timed(rollout_kernel_once, n_roll) * 1e-3)
    del env, state, graph, rgraph, traj, traj_dev, traj_host, pol

    # ---- (3a) STRONG scaling of BASELINE configs[3]: args.total_envs environments over the N GPUs (contiguous blocks,
    #           dist.shard_range == reshape_pytree_leading_dim of the pmap trainer), beside the per-GPU rate at the full batch --
    shard = D.shard_range(args.total_envs, rank, ws)
    s_inner = max(1, inner // 4)
    _, _, _, s_ms, s_kern, _ = step_leg(mac, ld, shard.count, s_inner, 4321)
    strong_value = args.total_envs * K * s_inner / (s_ms * 1e-3)
    if ws > 1:
        _, _, _, f_ms, f_kern, _ = step_leg(mac, ld, args.total_envs, max(1, s_inner // 2), 4321)
        full_rate = args.total_envs * K * max(1, s_inner // 2) / (f_ms * 1e-3)     # env-steps/s of ONE GPU on all envs
    else:
        full_rate, f_kern = strong_value, s_kern
    warps_per_pass = 148 * 23      # lob_step_scan_kernel<4>: one persistent CTA of 23 warps (= environments) per SM
    strong = {"total_envs": args.total_envs, "envs_per_gpu": shard.count, "value": strong_value, "unit": "env-steps/s",
              "ms_per_step": s_kern, "one_gpu_all_envs_value": full_rate,
              "efficiency": strong_value / (ws * full_rate),
              "grid_passes": {"exact": shard.count / warps_per_pass, "run": -(-shard.count // warps_per_pass),
                              "quantisation_efficiency": (shard.count / warps_per_pass) / (-(-shard.count // warps_per_pass))},
              "note": "per-GPU batch shrinks with N: the last pass of the scan kernel's persistent grid (148 SMs x 23 "
                      "environments) runs partly empty -- e.g. 8192 envs = 2.41 passes run as 3"}

    # ---- (3b) the other env.step configurations of BASELINE.json, kernel-only (same timing rules) ----
    others = []
    for cfg_name, n_envs, o_inner, label in (
            ("exec_longrun_fixed_quants_complex", 8192, inner, "BASELINE configs[2]: single execution agent"),
            ("hetero_deep_book", 16384, max(1, inner // 8), "BASELINE configs[4] shapes: 3 MM + 2 EXE + 2 directional, "
                                                           "512-row sides, 256-row trade log (131072 envs / 8 GPUs)")):
        mac2 = C.load_named_config(cfg_name)
        ld2 = ld if mac2.world_config.n_data_msg_per_step == ND and mac2.world_config.episode_time == mac.world_config.episode_time else _load_day(mac2)
        env2, st2, _, ms2, kern2, _ = step_leg(mac2, ld2, n_envs, o_inner, 77)
        No2, Nt2, N2 = env2.cfg.book.n_orders, env2.cfg.book.n_trades, env2.num_msgs_per_step
        by2 = _step_bytes(env2, No2, Nt2, N2, env2.cfg.n_data_msg_per_step)
        v2 = ws * n_envs * K * o_inner / (ms2 * 1e-3)
        others.append({"workload": f"MARLEnv.step {cfg_name} ({label})", "envs_per_gpu": n_envs, "msgs_per_env_step": N2,
                       "value": v2, "unit": "env-steps/s", "msgs_per_sec": v2 * N2, "ms_per_step": kern2,
                       "algorithmic_bytes_per_env_step": by2,
                       "roofline_frac": by2 * n_envs / (kern2 * 1e-3) / 1e9 / hbm_peak})
        del env2, st2

    # ---- (4) CPU baseline beside it (rank 0, N=1): the oracle on a bounded sample of the replay workload ----
    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu:
        from oracle import lob_oracle
        oracle = lob_oracle.load()
        threads = oracle.max_threads()
        cb = args.ref_books
        cw = np.arange(cb) % W
        ca, cbid, ct = params_np["init_asks"][cw].copy(), params_np["init_bids"][cw].copy(), params_np["init_trades"][cw].copy()
        cstart = ld.starts[cw].astype(np.int64)
        oracle.replay(bc, ca.copy(), cbid.copy(), ct.copy(), ld.msgs, cstart, 200, n_threads=threads)  # warm the threads
        t0 = time.perf_counter()
        reps = 0
        while True:
            oracle.replay(bc, ca, cbid, ct, ld.msgs, (cstart + reps * WINDOW) % (M - WINDOW), WINDOW, n_threads=threads)
            reps += 1
            if time.perf_counter() - t0 > args.cpu_seconds or reps >= 400:
                break
        dt = time.perf_counter() - t0
        cpu = {"value": cb * WINDOW * reps / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{cb} books x {WINDOW} msgs x {reps} passes ({dt:.1f} s), oracle/lob_oracle.c, OpenMP over books"}

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": K, "warmup": Wm,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": "pure book replay (job.scan_through_entire_array), synthetic LOBSTER day, nOrders=100, "
                                   "nTrades=100, BASELINE configs[1]",
                       "books_per_gpu": B, "msgs_per_book_per_step": TW, "day_msgs": int(M),
                       "step": f"ONE launch scanning {args.windows_per_step} consecutive 6400-message windows per book",
                       "l2_policy": f"inputs larger than L2: {B * (2 * No * 24 + Nt * 32) / 1e6:.0f} MB of book state per GPU"},
            "base_env_steps_per_sec": value / ND,
            "roofline": {"bound": "hbm", "kernel": "lob_replay_kernel", "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": _traffic("lob_replay_kernel"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms": kern_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d.value), "d2h_bytes_per_step": int(d2h.value),
                    "api": "lob_host_replay_run (C ABI, HOST buffers): books + trade logs + start offsets from pinned host "
                           "memory, scan, books + trade logs back to the host -- every step",
                    "resident_value": resident_value,
                    "resident_api": "BaseLOBEnv.replay: book state device-resident as in the reference; start offsets "
                                    f"({B * 8} B) in, best bid / ask per book ({B * 16} B) out"},
            "gpu_launches": launches,
            "clocks": clocks,
            "env_step": {"metric": "env_steps_per_sec", "value": step_value, "unit": "env-steps/s",
                         "msgs_per_sec": step_value * N, "ms_per_step": step_kern_ms,
                         "config": {"workload": "MARLEnv.step 2_player_fq_fqc (MM fixed_quants + EXE fixed_quants_complex), "
                                                "BASELINE configs[3] shapes", "envs_per_gpu": args.envs, "msgs_per_env_step": N,
                                    "launches_per_bench_step": inner},
                         "roofline": {"bound": "hbm", "kernel": "lob_step_scan_kernel (dominant: ~75 % of the step) + "
                                                                   "lob_step_prep_kernel + lob_agents_finish_kernel + "
                                                                   "lob_step_reset_done_kernel: the four launches of ONE "
                                                                   "lob_step_launch call, timed together",
                                      "achieved": step_achieved,
                                      "peak": hbm_peak, "unit": "GB/s", "frac": step_achieved / hbm_peak, "traffic": _traffic("lob_step_piped"),
                                      "algorithmic_bytes_per_env_step": step_bytes, "kernel_ms": step_kern_ms},
                         "e2e": {"value": step_e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": step_h2d,
                                 "d2h_bytes_per_step": step_d2h,
                                 "api": "MARLEnv.capture_step: actions from pinned host memory, PRNG draw, the piped step (4 launches), "
                                        "obs / rewards / done read back -- one CUDA graph launch per step",
                                 "eager_value": step_e2e_eager_value,
                                 "rollout_value": rollout_value,
                                 "rollout": f"MARLEnv.capture_rollout: {RT} steps + pre-sampled policy as one CUDA graph, "
                                            f"{rollout_d2h} B of trajectory (obs, rewards, done) read back per rollout, then "
                                            "dist.reduce_episode_stats over the ranks",
                                 "rollout_kernel_value": rollout_kernel_value,
                                 "rollout_kernel": f"MARLEnv.rollout: the same {RT} steps as ONE launch of lob_rollout_launch "
                                                   "(books resident in shared memory across the steps), draws and "
                                                   "trajectory read-back included",
                                 "episode_stat_reduce_us": reduce_us if ws > 1 else None,
                                 "episode_stat_samples": stats_count},
                         "strong_scaling": strong,
                         "gpu_launches": step_launches},
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        out["other_configs"] = others
    if ws > 1:
        dist.destroy_process_group()
    return json.dumps(out) if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--books", type=int, default=16384, help="books per GPU (replay workload)")
    ap.add_argument("--envs", type=int, default=16384, help="environments per GPU (env.step workload, weak scaling)")
    ap.add_argument("--total-envs", type=int, default=65536, help="environments over ALL GPUs (strong-scaling leg, BASELINE configs[3])")
    ap.add_argument("--windows-per-step", type=int, default=6, help="6400-message windows one replay step (launch) scans per book")
    ap.add_argument("--env-inner", type=int, default=192, help="env.step launches per bench step (3 episodes of 64 steps)")
    ap.add_argument("--ref-books", type=int, default=256, help="books in the CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    # stdout carries exactly ONE line (the JSON): anything a library prints on fd 1 meanwhile (NCCL's version banner)
    # is sent to stderr, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_native(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if line is not None:
        print(line, flush=True)


if __name__ == "__main__":
    main()
