"""The public Python API (env.MARLEnv / BaseLOBEnv, the reference's names) on a B200: shapes, dict keys, the device-side
PRNG products, and agreement with the oracle when it is fed the same draws."""
import numpy as np
import pytest

import helpers as H
from jaxmarl_hft_b200 import abi, config as C, env as E, lobster

pytestmark = pytest.mark.gpu


def test_draw_kernel_products_are_valid_and_deterministic():
    import torch
    mac = H.load_mac("hetero_deep_book")
    ld = H.load_for(mac, H.small_day(n_events=30000))
    env = E.MARLEnv(None, mac, num_envs=4096, loaded=ld, device="cuda:0", seed=7)
    obs, state = env.reset(None, env.default_params)
    a = {k: state.arrays[k].clone() for k in ("perm", "reset_window", "reset_is_sell")}
    n_act = env.num_action_msgs_per_step_by_all_agents
    srt = torch.sort(a["perm"], dim=1).values.cpu().numpy()
    np.testing.assert_array_equal(srt, np.tile(np.arange(n_act), (4096, 1)))       # every row is a permutation
    assert int(a["reset_window"].min()) >= 0 and int(a["reset_window"].max()) < env.cfg.n_windows
    assert set(np.unique(a["reset_is_sell"].cpu().numpy())) <= {0, 1}
    assert len(np.unique(a["reset_window"].cpu().numpy())) == env.cfg.n_windows    # all windows get drawn
    first = a["perm"][:, 0].float().mean().item()
    assert abs(first - (n_act - 1) / 2) < 0.5                                       # roughly uniform
    env2 = E.MARLEnv(None, mac, num_envs=4096, loaded=ld, device="cuda:0", seed=7)
    _, st2 = env2.reset(None, env2.default_params)
    for k in a:
        assert torch.equal(a[k], st2.arrays[k])                                     # deterministic in (seed, counter)


def test_marl_env_api_matches_reference_surface_and_oracle(oracle):
    import torch
    mac = H.load_mac("2_player_fq_fqc")
    ld = H.load_for(mac, H.small_day(n_events=30000))
    B = 64
    env = E.MARLEnv(None, mac, num_envs=B, loaded=ld, device="cuda:0", seed=3)
    assert env.type_names == ["MM", "EXE"] and env.num_agents == 2
    assert [s.n for s in env.action_spaces] == [10, 13]
    assert [s.shape for s in env.observation_spaces] == [(2,), (12,)]
    params = env.default_params
    assert params.loaded_params.message_data.shape == (ld.msgs.shape[0], 8)
    obs, state = env.reset(None, params)
    assert [tuple(o.shape) for o in obs] == [(B, 1, 2), (B, 1, 12)]
    assert tuple(state.world_state.ask_raw_orders.shape) == (B, 100, 6)
    ref = H.OracleEnv(oracle, mac, ld, B)
    ref.arrays.update({k: v.copy() for k, v in H.to_numpy(state.arrays).items()})   # same reset state + draws
    rng = np.random.default_rng(0)
    for s in range(66):
        acts = [torch.from_numpy(rng.integers(0, sp.n, size=(B, 1)).astype(np.int32)).cuda() for sp in env.action_spaces]
        obs, state, rewards, dones, info = env.step(None, state, acts, params)
        got = H.to_numpy(state.arrays)
        H.copy_inputs(got, ref.arrays)              # the oracle gets the draws the device made
        ref.step()
        H.assert_arrays_match(ref.arrays, got, ref.cfg)
        assert dones["__all__"].dtype == torch.bool and tuple(dones["agents"][1].shape) == (B, 1)
    assert set(info["world"]) >= {"window_index", "end_mid_price", "time", "average_best_ask", "abort_episode", "spread"}
    assert set(info["agents"][0]) == set(abi.MMINFO_I32 + abi.MMINFO_F32)
    assert set(info["agents"][1]) == set(abi.EXEINFO_I32 + abi.EXEINFO_F32)
    with pytest.raises(ValueError):
        env.reset(None, None)


def test_base_env_replay_api(oracle):
    import torch
    mac = H.load_mac("2_player_fq_fqc")
    ld = H.load_for(mac, H.small_day(n_events=30000))
    be = E.BaseLOBEnv(mac.world_config, loaded=ld, device="cuda:0")
    p = be.default_params
    W = be.n_windows
    a = p.init_states_array["init_asks"].clone(); b = p.init_states_array["init_bids"].clone()
    t = p.init_states_array["init_trades"].clone()
    start = torch.from_numpy(ld.starts.astype(np.int64)).cuda()
    best = torch.zeros((W, 4), dtype=torch.int32, device="cuda")
    be.replay(a, b, t, start, 100, best_out=best)      # one base-env step (base_env.py:189) for every window
    ra, rb, rt = (x.cpu().numpy().copy() for x in (p.init_states_array["init_asks"], p.init_states_array["init_bids"],
                                                    p.init_states_array["init_trades"]))
    rbest = np.zeros((W, 4), np.int32)
    oracle.replay(be.book_cfg, ra, rb, rt, ld.msgs, ld.starts.astype(np.int64), 100, best_out=rbest)
    np.testing.assert_array_equal(a.cpu().numpy(), ra); np.testing.assert_array_equal(b.cpu().numpy(), rb)
    np.testing.assert_array_equal(t.cpu().numpy(), rt); np.testing.assert_array_equal(best.cpu().numpy(), rbest)


def test_capture_step_graph_replays_like_eager_steps(oracle):
    """MARLEnv.capture_step: the graph (actions copy + device draw + step) replayed k times == k eager steps of a twin
    env with the same seed (the draw counter lives in device memory, so every replay draws fresh values)."""
    import torch
    mac = H.load_mac("2_player_fq_fqc")
    ld = H.load_for(mac, H.small_day(n_events=30000))
    B = 128
    envs = [E.MARLEnv(None, mac, num_envs=B, loaded=ld, device="cuda:0", seed=11) for _ in range(2)]
    params = [e.default_params for e in envs]
    states_ = [e.reset(None, p)[1] for e, p in zip(envs, params)]
    rng = np.random.default_rng(1)
    acts = [torch.zeros((B, 1), dtype=torch.int32, device="cuda") for _ in range(2)]
    before = H.to_numpy(states_[0].arrays)
    graph, out = envs[0].capture_step(states_[0], acts, params[0])     # the warm-up step is undone: capturing leaves no trace
    after = H.to_numpy(states_[0].arrays)
    for k in before:
        np.testing.assert_array_equal(before[k], after[k], err_msg=k)
    for k in range(5):
        a_np = [rng.integers(0, sp.n, size=(B, 1)).astype(np.int32) for sp in envs[0].action_spaces]
        for t in range(2):
            acts[t].copy_(torch.from_numpy(a_np[t]))
        graph.replay()
        envs[1].step(None, states_[1], acts, params[1])
    torch.cuda.synchronize()
    a, b = H.to_numpy(states_[0].arrays), H.to_numpy(states_[1].arrays)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert int(a["step_counter"].max()) == 5


def test_marl_env_fixed_time_episodes():
    """ep_type="fixed_time" through the public API: time-grid windows from the loader, 2- / 15-field observations, episodes
    of different lengths all ending and auto-resetting (marl_env.py:717-718 counts steps in either mode)."""
    import torch
    mac = H.load_mac("2_player_fq_fqc", ep_type="fixed_time", episode_time=1800, start_resolution=900)
    ld = H.load_for(mac, H.small_day(n_events=30000))
    B = 256
    env = E.MARLEnv(None, mac, num_envs=B, loaded=ld, device="cuda:0", seed=3)
    assert [sp.shape for sp in env.observation_spaces] == [(2,), (15,)]
    params = env.default_params
    obs, state = env.reset(None, params)
    assert obs[1].shape == (B, 1, 15)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    n_done = torch.zeros(B, dtype=torch.int64, device="cuda")
    for _ in range(40):
        acts = [torch.randint(0, sp.n, (B, 1), generator=g, device="cuda", dtype=torch.int32) for sp in env.action_spaces]
        obs, state, rewards, dones, info = env.step(None, state, acts, params)
        n_done += dones["__all__"].to(torch.int64)
        assert torch.isfinite(obs[1]).all()
    assert int(n_done.min()) >= 1                                  # every env finished at least one episode
    it = state.world_state.init_time.cpu().numpy()
    assert (it[:, 1] == 0).all() and ((it[:, 0] - 34200) % 900 == 0).all()   # base_env.py:288-290: on the time grid


def test_marl_env_random_cancel_mode_draws():
    """cancel_mode 3 through the public API: the device generator fills cancel_u (multiples of 2^-23 in [0,1), fresh every
    step) and the step consumes it."""
    import torch
    mac = H.load_mac("2_player_fq_fqc", nOrders=40, nTrades=24, cancel_mode=3)
    ld = H.load_for(mac, H.small_day(seed=9, n_events=30000, stress=True))
    B = 64
    env = E.MARLEnv(None, mac, num_envs=B, loaded=ld, device="cuda:0", seed=5)
    params = env.default_params
    obs, state = env.reset(None, params)
    acts = [torch.zeros((B, 1), dtype=torch.int32, device="cuda") for _ in env.action_spaces]
    seen = []
    for _ in range(3):
        obs, state, rewards, dones, info = env.step(None, state, acts, params)
        u = state.arrays["cancel_u"].clone()
        assert u.shape == (B, env.num_msgs_per_step, 2)
        assert float(u.min()) >= 0.0 and float(u.max()) < 1.0
        assert torch.equal(u * 2 ** 23, torch.round(u * 2 ** 23))
        seen.append(u)
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])
    assert 0.45 < float(seen[0].mean()) < 0.55


def test_capture_rollout_graph_equals_eager_steps():
    """MARLEnv.capture_rollout: T steps + policy as one graph == T eager steps of a twin env (same seed, same policy)."""
    import torch
    mac = H.load_mac("2_player_fq_fqc")
    ld = H.load_for(mac, H.small_day(n_events=30000))
    B, T = 96, 8
    envs = [E.MARLEnv(None, mac, num_envs=B, loaded=ld, device="cuda:0", seed=4) for _ in range(2)]
    params = [e.default_params for e in envs]
    states_ = [e.reset(None, p)[1] for e, p in zip(envs, params)]
    n = [sp.n for sp in envs[0].action_spaces]

    def policy(k, obs):    # a deterministic function of the observation: capturable device work only
        return [((obs[t].abs().sum(-1) * 1000.0).to(torch.int64) % n[t]).to(torch.int32) for t in range(2)]

    graph, traj = envs[0].capture_rollout(states_[0], policy, T, params[0])          # (the warm-up step is undone)
    obs1 = [states_[1].arrays[f"obs{t}"] for t in range(2)]
    graph.replay()
    ref = {"obs": [[], []], "reward": [[], []], "done": []}
    for k in range(T):
        o, _, r, d, _ = envs[1].step(None, states_[1], policy(k, obs1), params[1])
        for t in range(2):
            ref["obs"][t].append(o[t].clone()); ref["reward"][t].append(r[t].clone())
        ref["done"].append(d["__all__"].clone())
    torch.cuda.synchronize()
    for t in range(2):
        assert torch.equal(traj["obs"][t], torch.stack(ref["obs"][t]))
        assert torch.equal(traj["reward"][t], torch.stack(ref["reward"][t]))
    assert torch.equal(traj["done"].bool(), torch.stack(ref["done"]))
    a, b = H.to_numpy(states_[0].arrays), H.to_numpy(states_[1].arrays)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert int(a["step_counter"].max()) == T


def test_graft_entry_smoke_runs():
    """The driver's smoke check on cuda:0."""
    import importlib
    importlib.import_module("__graft_entry__").smoke()


@pytest.mark.parametrize("ep_type", ["fixed_steps", "fixed_time"])
def test_base_env_reset_step_pair(oracle, ep_type):
    """BaseLOBEnv.reset_env / step_env (base_env.py:189-234), batched, against the oracle's replay of the same slices."""
    import torch
    over = dict(ep_type="fixed_time", episode_time=1800, start_resolution=900) if ep_type == "fixed_time" else {}
    mac = H.load_mac("2_player_fq_fqc", **over)
    ld = H.load_for(mac, H.small_day(n_events=30000))
    w = mac.world_config
    be = E.BaseLOBEnv(w, loaded=ld, device="cuda:0")
    B = 40
    _, st = be.reset_env(None, be.default_params, num_envs=B, seed=3)
    widx = st.window_index.cpu().numpy()
    assert len(set(widx.tolist())) > 3
    P = be._params_np
    ra, rb, rt = P["init_asks"][widx].copy(), P["init_bids"][widx].copy(), P["init_trades"][widx].copy()
    np.testing.assert_array_equal(st.ask_raw_orders.cpu().numpy(), ra)
    init_t = P["init_init_time"][widx]
    Nd = w.n_data_msg_per_step
    for k in range(5):
        obs, st, rew, done, info = be.step_env(None, st, None, be.default_params)
        off = np.clip(P["init_start_index"][widx].astype(np.int64) + Nd * k, 0, ld.msgs.shape[0] - Nd)
        if ep_type == "fixed_time":
            sl = ld.msgs[off[:, None] + np.arange(Nd)[None, :]].copy()
            late = sl[:, :, 6] >= (init_t[:, 0] + w.episode_time)[:, None]
            sl[late, :6] = 0
            oracle.replay(be.book_cfg, ra, rb, rt, sl.reshape(-1, 8), np.arange(B, dtype=np.int64) * Nd, Nd)
        else:
            oracle.replay(be.book_cfg, ra, rb, rt, ld.msgs, off, Nd)
        np.testing.assert_array_equal(st.ask_raw_orders.cpu().numpy(), ra); np.testing.assert_array_equal(st.bid_raw_orders.cpu().numpy(), rb)
        np.testing.assert_array_equal(st.trades.cpu().numpy(), rt)
        np.testing.assert_array_equal(done.cpu().numpy(), (ld.msgs[off + Nd - 1, 6] - init_t[:, 0]) >= w.episode_time)
        assert obs == 0 and rew == 0 and info == {"info": 0} and int(st.step_counter[0]) == k + 1


def test_host_buffer_replay_matches_oracle(oracle):
    """lob_host_replay_* (the C ABI's HOST-buffer path: H2D + replay + D2H inside one call, bench.py's e2e leg) against the
    oracle; bad offsets are an error, not a silently unprocessed book."""
    import ctypes
    from jaxmarl_hft_b200 import _lib, abi
    L = _lib.lib()
    rng = np.random.default_rng(12)
    bc = C.book_config(C.World_EnvironmentConfig(nOrders=100, nTrades=100))
    B, T, M = 200, 333, 200 * 333 + 50
    msgs = H.random_messages(rng, M, bc)
    start = rng.integers(0, M - T + 1, size=B).astype(np.int64)
    a = np.full((B, 100, 6), -1, np.int32); b = a.copy(); t = np.full((B, 100, 8), -1, np.int32)
    ra, rb, rt = a.copy(), b.copy(), t.copy()
    oracle.replay(bc, ra, rb, rt, msgs, start, T)
    h = L.lob_host_replay_create(ctypes.byref(bc), B, M, 0)
    assert h, L.lob_last_error()
    p = lambda x: x.ctypes.data_as(abi.p_i32)
    try:
        _lib.check(L.lob_host_replay_set_messages(h, p(msgs), M), "set_messages")
        h2d, d2h = ctypes.c_int64(0), ctypes.c_int64(0)
        for _ in range(2):      # two calls continue the same books: replay twice on the host side too
            _lib.check(L.lob_host_replay_run(h, p(a), p(b), p(t), start.ctypes.data_as(abi.p_i64), T, B,
                                             ctypes.byref(h2d), ctypes.byref(d2h)), "run")
        oracle.replay(bc, ra, rb, rt, msgs, start, T)
        np.testing.assert_array_equal(a, ra); np.testing.assert_array_equal(b, rb); np.testing.assert_array_equal(t, rt)
        assert h2d.value == 2 * a.nbytes + t.nbytes + start.nbytes and d2h.value == 2 * a.nbytes + t.nbytes
        bad = start.copy(); bad[7] = M - T + 1
        rc = L.lob_host_replay_run(h, p(a), p(b), p(t), bad.ctypes.data_as(abi.p_i64), T, B, None, None)
        assert rc == abi.LOB_E_INVALID and b"start[7]" in L.lob_last_error()
    finally:
        L.lob_host_replay_destroy(h)


def _day_with_everything(seed):
    """A synthetic day plus what a real LOBSTER file also holds: rows of types 5 / 6 / 7 (dropped), rows outside the trading
    window, bursts of same-time-stamp executions in BOTH directions interleaved with other rows (merge_market_orders)."""
    rng = np.random.default_rng(seed)
    day = H.small_day(seed=seed, n_events=20000)
    m, ob = day.messages.copy(), day.orderbook.copy()
    n = m.shape[0]
    for i in rng.choice(np.arange(50, n - 50), size=300, replace=False):    # same-time-stamp bursts of executions
        k = int(rng.integers(2, 6))
        m[i:i + k, 0] = m[i, 0]
        m[i:i + k, 1] = rng.choice([4, 4, 4, 1, 5], size=k)
        m[i:i + k, 5] = rng.choice([-1, 1], size=k)
    odd = rng.choice(n, size=200, replace=False)
    m[odd, 1] = rng.choice([5, 6, 7], size=200)
    m[:40, 0] = 34100.0 + np.arange(40) * 0.5            # before day_start
    m[-30:, 0] = 57601.0 + np.arange(30) * 0.25          # after day_end
    assert (np.diff(m[:, 0]) >= 0).all()
    return lobster.RawDay(messages=m, orderbook=ob, levels=day.levels)


@pytest.mark.parametrize("seed", [3, 4])
def test_loader_preprocessing_on_the_device_matches_the_host_loader(seed):
    """csrc/lob_loader.cu (lobster_loader.py:891-945, :1073-1132 on the device) == lobster.preprocess_day, which
    tests/test_loader_vs_reference.py pins against the reference's own pandas loader."""
    import torch
    from jaxmarl_hft_b200 import lobster
    day = _day_with_everything(seed)
    m_ref, ob_ref, t_ref = lobster.preprocess_day(day, return_time=True)
    msgs, rows, tm = lobster.preprocess_day_cuda(day, device="cuda:0")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(msgs.cpu().numpy().astype(np.int64), m_ref)
    np.testing.assert_array_equal(day.orderbook[rows.cpu().numpy()], ob_ref)
    np.testing.assert_array_equal(tm.cpu().numpy(), t_ref)
    assert (m_ref[:, 0] == 4).sum() > 100
    for kind, kw in (("fixed_steps", {}), ("fixed_time", dict(window_length=1800, window_resolution=900))):
        a = lobster.load_days([day], kw.get("window_length", 64), 100, kw.get("window_resolution", 64), window_type=kind)
        b = lobster.load_days([day], kw.get("window_length", 64), 100, kw.get("window_resolution", 64), window_type=kind,
                              device="cuda:0")
        for f in ("msgs", "starts", "ends", "books", "max_msgs"):
            np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f"{kind} {f}")
        assert b.msgs_device is not None and b.msgs_device.is_cuda
    bad = lobster.RawDay(messages=day.messages[::-1].copy(), orderbook=day.orderbook, levels=day.levels)
    with pytest.raises(ValueError, match="time-sorted"):
        lobster.preprocess_day_cuda(bad, device="cuda:0")


@pytest.mark.parametrize("config,kw,fused", [("2_player_fq_fqc", {}, False), ("2_player_fq_fqc", {}, True),
                                             ("2_player_fq_fqc", {"cancel_mode": 3, "nOrders": 48, "nTrades": 20}, False),
                                             ("2_player_fq_fqc", {"cancel_mode": 3, "nOrders": 48, "nTrades": 20}, True),
                                             ("hetero_deep_book", {}, False)])
def test_rollout_kernel_equals_eager_steps(config, kw, fused):
    """MARLEnv.rollout (lob_rollout_launch: T steps per environment in one launch, books resident in shared memory) == T
    calls of MARLEnv.step on a twin env with the same seed: trajectory outputs row by row, every state / output / info
    leaf at the end.  70 steps cross the auto-reset."""
    import torch
    mac = H.load_mac(config, **kw)
    ld = H.load_for(mac, H.small_day(n_events=30000, **({"seed": 9, "stress": True} if kw else {})))
    B, T = 80, 70
    envs = [E.MARLEnv(None, mac, num_envs=B, loaded=ld, device="cuda:0", seed=21) for _ in range(2)]
    params = [e.default_params for e in envs]
    states_ = [e.reset(None, p)[1] for e, p in zip(envs, params)]
    nt = envs[0].cfg.n_agent_types
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    acts = [torch.randint(0, envs[0].action_spaces[t].n, (T,) + tuple(states_[0].arrays[f"actions{t}"].shape), generator=g,
                          device="cuda", dtype=torch.int32) for t in range(nt)]
    if fused:   # without the split workspace lob_rollout_launch is ONE launch of the fused kernel (the books stay in shared
        states_[0].arrays.pop("work_split")   # memory from the first to the last step); with it, T piped steps (lob_pipe.cuh)
        envs[0]._cache = {}                   # (the packed buffer struct of these arrays was cached with the workspace pointer)
    traj, _ = envs[0].rollout(states_[0], acts, T, params[0])
    ref = {"obs": [[] for _ in range(nt)], "reward": [[] for _ in range(nt)], "done": []}
    for k in range(T):
        o, _, r, d, _ = envs[1].step(None, states_[1], [a[k] for a in acts], params[1])
        for t in range(nt):
            ref["obs"][t].append(o[t].clone()); ref["reward"][t].append(r[t].clone())
        ref["done"].append(d["__all__"].clone())
    torch.cuda.synchronize()
    for t in range(nt):
        assert torch.equal(traj["obs"][t].view(torch.int32), torch.stack(ref["obs"][t]).view(torch.int32))
        assert torch.equal(traj["reward"][t].view(torch.int32), torch.stack(ref["reward"][t]).view(torch.int32))
    assert torch.equal(traj["done"].bool(), torch.stack(ref["done"]))
    assert int(traj["done"].sum()) == B
    a, b = H.to_numpy(states_[0].arrays), H.to_numpy(states_[1].arrays)
    inputs = ("perm", "reset_window", "reset_is_sell", "cancel_u")     # the rollout reads its own [T, ...] copies of these
    for k in a:
        if not (k.startswith("work_") or k.startswith("actions") or k in inputs):
            np.testing.assert_array_equal(a[k].view(np.uint8), b[k].view(np.uint8), err_msg=k)
