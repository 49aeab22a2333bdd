"""Parity of the CUDA path (through the C ABI) against the CPU oracle, on a B200.

Integer / byte leaves: bit-exact.  float32 leaves: rel 1e-5 (the tolerance BASELINE.json's north_star states); the
kernel and the oracle both sum float reductions left-to-right, so in practice they agree far tighter."""
import ctypes

import numpy as np
import pytest

import helpers as H
from jaxmarl_hft_b200 import abi, config as C, env as E, lobster, states

pytestmark = pytest.mark.gpu


def _book_cfg(no=100, nt=100, **kw):
    w = C.World_EnvironmentConfig(nOrders=no, nTrades=nt, **kw)
    return C.book_config(w)


@pytest.mark.parametrize("no,nt,t4,fill", [(100, 100, 0, True), (32, 16, 0, True), (20, 8, 1, True), (100, 50, 2, True),
                                           (64, 100, 0, False), (200, 128, 0, True), (512, 256, 0, True), (7, 3, 0, True)])
def test_replay_random_streams_bit_exact(oracle, no, nt, t4, fill):
    """job.scan_through_entire_array on adversarial random streams, including books that overflow (eviction, the
    last-row quirks Q1/Q2/Q5) and trade logs that overflow (Q3)."""
    rng = np.random.default_rng(no * 1000 + nt)
    bc = _book_cfg(no, nt, type_4_interpretation=t4, check_book_fill=fill)
    B, T = 64, 700
    msgs = H.random_messages(rng, B * T, bc, price_lo=99_000, price_hi=100_600 if no < 64 else 101_500)
    start = (np.arange(B, dtype=np.int64) * T)
    a0 = np.full((B, no, 6), -1, np.int32); b0 = a0.copy(); t0 = np.full((B, nt, 8), -1, np.int32)
    ra, rb, rt = a0.copy(), b0.copy(), t0.copy()
    rbest = np.zeros((B, 4), np.int32)
    oracle.replay(bc, ra, rb, rt, msgs, start, T, best_out=rbest)
    ga, gb, gt, gbest = H.cuda_replay(bc, a0, b0, t0, msgs, start, T, want_best=True)
    np.testing.assert_array_equal(ga, ra); np.testing.assert_array_equal(gb, rb)
    np.testing.assert_array_equal(gt, rt); np.testing.assert_array_equal(gbest, rbest)
    assert (rt[:, :, 0] >= 0).any() and (ra[:, :, 0] >= 0).any()


def test_replay_ragged_and_empty(oracle):
    """Zero messages, one message, a non-multiple of the staging chunk, windows ending at the array end."""
    rng = np.random.default_rng(3)
    bc = _book_cfg()
    msgs = H.random_messages(rng, 5000, bc)
    for T in (0, 1, 63, 64, 65, 129):
        B = 9
        start = rng.integers(0, 5000 - T + 1, size=B).astype(np.int64)
        start[-1] = 5000 - T
        a0 = np.full((B, 100, 6), -1, np.int32); b0 = a0.copy(); t0 = np.full((B, 100, 8), -1, np.int32)
        ra, rb, rt = a0.copy(), b0.copy(), t0.copy()
        oracle.replay(bc, ra, rb, rt, msgs, start, T)
        ga, gb, gt = H.cuda_replay(bc, a0, b0, t0, msgs, start, T)
        np.testing.assert_array_equal(ga, ra); np.testing.assert_array_equal(gb, rb); np.testing.assert_array_equal(gt, rt)


def test_replay_synthetic_day_and_l2(oracle):
    """Config #2 at test size: windows of the synthetic LOBSTER day replayed from their reset states; then L2."""
    mac = H.load_mac("2_player_fq_fqc")
    day = H.small_day(n_events=30000)
    ld = H.load_for(mac, day)
    bc = C.book_config(mac.world_config)
    params = E.build_reset_params(ld, mac.world_config, H.oracle_replay_fn(oracle, bc))
    W = ld.starts.shape[0]
    B = 48
    widx = np.arange(B) % W
    a0, b0, t0 = params["init_asks"][widx].copy(), params["init_bids"][widx].copy(), params["init_trades"][widx].copy()
    start = ld.starts[widx].astype(np.int64) + (np.arange(B) // W) * 100
    T = 6400
    ra, rb, rt = a0.copy(), b0.copy(), t0.copy()
    oracle.replay(bc, ra, rb, rt, ld.msgs, start, T)
    ga, gb, gt = H.cuda_replay(bc, a0, b0, t0, ld.msgs, start, T)
    np.testing.assert_array_equal(ga, ra); np.testing.assert_array_equal(gb, rb); np.testing.assert_array_equal(gt, rt)
    import torch
    l2 = E.l2_state(bc, torch.from_numpy(ga).cuda(), torch.from_numpy(gb).cuda(), 10).cpu().numpy()
    np.testing.assert_array_equal(l2, oracle.l2(bc, ra, rb, 10))
    # reset-state precompute through the CUDA replay == through the oracle (base_env.py:245-296)
    p2 = E.build_reset_params(ld, mac.world_config, E._cuda_replay_fn(bc, "cuda:0"))
    for k in params:
        np.testing.assert_array_equal(p2[k], params[k], err_msg=k)


def test_l2_on_sparse_and_empty_books(oracle):
    rng = np.random.default_rng(8)
    bc = _book_cfg(100, 100)
    B = 32
    asks = np.full((B, 100, 6), -1, np.int32); bids = asks.copy()
    for b in range(B):
        for side in (asks, bids):
            n = int(rng.integers(0, 12)) if b % 3 else 0
            rows = rng.choice(100, size=n, replace=False)
            side[b, rows, 0] = rng.integers(990, 1010, size=n) * 100
            side[b, rows, 1] = rng.integers(1, 500, size=n)
            side[b, rows, 2:] = 5
    import torch
    for n_levels in (1, 5, 10, 20):
        got = E.l2_state(bc, torch.from_numpy(asks).cuda(), torch.from_numpy(bids).cuda(), n_levels).cpu().numpy()
        np.testing.assert_array_equal(got, oracle.l2(bc, asks, bids, n_levels))


def _rollout_parity(oracle, mac, day, B, steps, seed, stress_actions=False, fused=False):
    ld = H.load_for(mac, day)
    ref = H.OracleEnv(oracle, mac, ld, B)
    gpu = H.CudaEnv(mac, ld, B, ref.params)
    if fused:   # no split workspace -> lob_step_launch runs the ONE fused kernel instead of the piped step (lob_pipe.cuh)
        gpu.arrays.pop("work_split")
    rng = np.random.default_rng(seed)
    H.draw_prng(rng, ref.cfg, ref.arrays)
    gpu.set_inputs(ref.arrays)
    ref.reset(); gpu.reset()
    H.assert_arrays_match(ref.arrays, gpu.numpy(), ref.cfg)
    n_done = 0
    for s in range(steps):
        H.draw_prng(rng, ref.cfg, ref.arrays)
        H.draw_actions(rng, ref.cfg, ref.arrays)
        if stress_actions and s % 7 == 3:   # out-of-range actions: jnp gather wraps once, then clamps
            for t in range(ref.cfg.n_agent_types):
                ref.arrays[f"actions{t}"][::5] = rng.integers(-3, 40, size=ref.arrays[f"actions{t}"][::5].shape)
        gpu.set_inputs(ref.arrays)
        ref.step(n_threads=8); gpu.step()
        try:
            H.assert_arrays_match(ref.arrays, gpu.numpy(), ref.cfg)
        except AssertionError as e:
            raise AssertionError(f"step {s}: {e}") from None
        n_done += int(ref.arrays["done_all"].sum())
    return ref, n_done


def test_step_2_player_rollout_with_auto_reset(oracle):
    """BASELINE config #1 shapes (2_player_fq_fqc: MM fixed_quants + EXE fixed_quants_complex), 70 steps so every env
    crosses an episode boundary (done on the 64th step, quirk Q14) and the fused auto-reset is exercised."""
    mac = H.load_mac("2_player_fq_fqc")
    ref, n_done = _rollout_parity(oracle, mac, H.small_day(n_events=30000), B=96, steps=70, seed=1, stress_actions=True)
    assert n_done == 96
    assert (ref.arrays["info_i32_1"][..., 0] >= 0).all()


@pytest.mark.parametrize("config,B,steps", [("2_player_fq_fqc", 96, 70), ("exec_longrun_fixed_quants_complex", 64, 70)])
def test_step_fused_kernel_without_workspace(oracle, config, B, steps):
    """The same rollouts through the FUSED step kernel (a caller that passes no workspace): the piped step -- four launches,
    what every other test of this file runs -- and the fused kernel are two schedules of the same device functions."""
    mac = H.load_mac(config)
    _, n_done = _rollout_parity(oracle, mac, H.small_day(n_events=30000), B=B, steps=steps, seed=11, stress_actions=True, fused=True)
    assert n_done == B


def test_step_exec_only(oracle):
    """BASELINE config #3 shapes: single execution agent, fixed_quants_complex."""
    mac = H.load_mac("exec_longrun_fixed_quants_complex")
    _rollout_parity(oracle, mac, H.small_day(n_events=30000), B=64, steps=40, seed=2)


def test_step_hetero_deep_book(oracle):
    """BASELINE config #5 shapes: 3 MM + 2 EXE + 2 directional, 512-row book sides, 256-row trade log."""
    mac = H.load_mac("hetero_deep_book")
    _rollout_parity(oracle, mac, H.small_day(n_events=30000), B=24, steps=30, seed=3)


def test_step_stress_day_capacity(oracle):
    """A day whose flow overfills a small book: eviction, last-row cancels and trade-log overflow inside env.step."""
    mac = H.load_mac("2_player_fq_fqc", nOrders=40, nTrades=24)
    _rollout_parity(oracle, mac, H.small_day(seed=9, n_events=30000, stress=True), B=48, steps=66, seed=4)


@pytest.mark.parametrize("reward", ["portfolio_value", "buy_sell_pnl", "complex", "zero_inv", "spooner",
                                    "spooner_damped", "spooner_asym_damped", "spooner_scaled",
                                    "delta_portfolio_value"])
def test_step_mm_reward_variants(oracle, reward):
    import dataclasses
    mac = H.load_mac("2_player_fq_fqc")
    agents = dict(mac.dict_of_agents_configs)
    mm = agents["MarketMaking"]
    agents["MarketMaking"] = dataclasses.replace(
        mm, reward_function=reward, observation_space="engineered", inv_penalty="quadratic",
        reference_price="far_touch" if reward in ("spooner", "portfolio_value") else "mid_avg",
        unwind_price="far_touch" if reward == "complex" else "mid_avg", clip_reward=True, exclude_extreme_spreads=True,
        volume_traded_bonus="market_share")
    exe = agents["Execution"]
    agents["Execution"] = dataclasses.replace(exe, action_space="fixed_quants", observation_space="basic",
                                              reference_price="far_touch", task="sell", reward_lambda=0.5)
    mac2 = H.with_agents(mac, agents, [2, 2])
    _rollout_parity(oracle, mac2, H.small_day(n_events=30000), B=32, steps=66, seed=5)


def test_full_size_properties():
    """BASELINE config #2 at full size (16384 books): size-independent properties instead of an oracle run --
    (i) determinism, (ii) replay(T1) then replay(T2) == replay(T1+T2) (scan composition), (iii) book invariants."""
    import torch
    mac = H.load_mac("2_player_fq_fqc")
    ld = H.load_for(mac, H.small_day(n_events=30000))
    bc = C.book_config(mac.world_config)
    params = E.build_reset_params(ld, mac.world_config, E._cuda_replay_fn(bc, "cuda:0"))
    B, W = 16384, ld.starts.shape[0]
    widx = np.arange(B) % W
    dev = torch.device("cuda:0")
    msgs = torch.from_numpy(ld.msgs).to(dev)
    start = torch.from_numpy(ld.starts[widx].astype(np.int64) + (np.arange(B) // W) % 3000).to(dev)

    def run(splits):
        a = torch.from_numpy(params["init_asks"][widx]).to(dev)
        b = torch.from_numpy(params["init_bids"][widx]).to(dev)
        t = torch.from_numpy(params["init_trades"][widx]).to(dev)
        off = 0
        for n in splits:
            E.replay_books(bc, a, b, t, msgs, start + off, n)
            off += n
        torch.cuda.synchronize()
        return a, b, t

    a1, b1, t1 = run([2000])
    a2, b2, t2 = run([2000])
    a3, b3, t3 = run([700, 1237, 63])
    for x, y in ((a1, a2), (b1, b2), (t1, t2), (a1, a3), (b1, b3), (t1, t3)):
        assert torch.equal(x, y)
    # invariants of reference-made books: a row is either all -1 or has qty > 0; the book is never crossed
    for side in (a1, b1):
        blank = (side == -1).all(dim=2)
        assert bool((blank | (side[:, :, 1] > 0)).all())
    best_ask = torch.where(a1[:, :, 0] == -1, torch.full_like(a1[:, :, 0], 2**31 - 1), a1[:, :, 0]).min(dim=1).values
    best_bid = b1[:, :, 0].max(dim=1).values
    assert bool((best_bid < best_ask).all())


@pytest.mark.parametrize("mm_space,exe_space", [("bobRL", "twap"), ("bobStrategy", "fixed_quants_1msg"),
                                                ("bobRL", "simplest_case"), ("simple", "twap"),
                                                ("spread_skew", "fixed_quants_complex"), ("AvSt", "fixed_quants_complex")])
def test_step_more_action_spaces(oracle, mm_space, exe_space):
    """First "next" row of SURVEY 8(f): MM bobRL / bobStrategy (mm_env.py:1474 / :1400) and EXE twap /
    fixed_quants_1msg / simplest_case (exec_env.py:1126 / :732 / :935)."""
    import dataclasses
    mac = H.load_mac("2_player_fq_fqc")
    agents = dict(mac.dict_of_agents_configs)
    agents["MarketMaking"] = dataclasses.replace(agents["MarketMaking"], action_space=mm_space, bob_v0=5,
                                                 observation_space="engineered", fixed_quant_value=2,
                                                 **({"n_actions": 4} if mm_space == "simple" else {}))
    agents["Execution"] = dataclasses.replace(agents["Execution"], action_space=exe_space, task_size=150,
                                              reward_function="simplest_case" if exe_space == "simplest_case" else "normal")
    _rollout_parity(oracle, H.with_agents(mac, agents, [2, 2]), H.small_day(n_events=30000), B=32, steps=66, seed=11,
                    stress_actions=True)


@pytest.mark.parametrize("episode_time,resolution", [(1800, 900), (900, 300)])
def test_step_fixed_time_episodes(oracle, episode_time, resolution):
    """ep_type="fixed_time" (base_env.py:288-291, :358-368; mm_env.py:3032, :1291; exec_env.py:1943): time-grid windows
    of unequal length, init_time on the grid (the last windows wrap to day_start so their data is masked), the
    10- / 15-field engineered observations and the AvSt horizon in seconds."""
    import dataclasses
    mac = H.load_mac("2_player_fq_fqc", ep_type="fixed_time", episode_time=episode_time, start_resolution=resolution)
    agents = dict(mac.dict_of_agents_configs)
    mm, ex = agents["MarketMaking"], agents["Execution"]
    agents = {"MarketMaking": dataclasses.replace(mm, observation_space="engineered"),
              "AvSt": dataclasses.replace(mm, short_name="AV", action_space="AvSt", observation_space="engineered",
                                          normalize=False, fixed_quant_value=4),
              "Execution": ex,
              "Exec2": dataclasses.replace(ex, short_name="EXE2", observation_space="simplest_case", task="buy", task_size=70)}
    ref, n_done = _rollout_parity(oracle, H.with_agents(mac, agents, [1, 2, 2, 1]), H.small_day(n_events=30000), B=64,
                                  steps=40, seed=17, stress_actions=True)
    assert n_done >= 64
    assert len(set(ref.params["init_max_steps"].tolist())) > 1   # windows really differ in length


def test_step_ten_plus_ten_agents(oracle):
    """The reference's largest shipped trainer config steps 10 market makers + 10 execution agents per environment
    (config/rl_configs/PMAP_ippo_rnn_JAXMARL_2player.yaml): 20 agents, 120 agent messages + 100 data messages per step,
    a 60-row permutation."""
    mac = H.load_mac("2_player_fq_fqc")
    agents = dict(mac.dict_of_agents_configs)
    ref, n_done = _rollout_parity(oracle, H.with_agents(mac, agents, [10, 10]), H.small_day(n_events=30000), B=24, steps=66,
                                  seed=23, stress_actions=True)
    assert ref.cfg.n_agent_types == 2 and C.num_msgs_per_step(ref.cfg) == 100 + 10 * 4 + 10 * 8
    assert n_done == 24


def test_step_six_agent_types(oracle):
    """MultiAgentConfig takes any number of agent types (marl_env.py:71-79); six different ones in one environment."""
    import dataclasses
    mac = H.load_mac("2_player_fq_fqc")
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    agents = {"MarketMaking": mm,
              "Skew": dataclasses.replace(mm, short_name="SK", action_space="spread_skew", fixed_quant_value=3),
              "Directional": dataclasses.replace(mm, short_name="DIR", action_space="directional_trading", fixed_quant_value=7),
              "Execution": ex,
              "Twap": dataclasses.replace(ex, short_name="TW", action_space="twap", task_size=150),
              "Prices": dataclasses.replace(ex, short_name="FP", action_space="fixed_prices", n_actions=3, fixed_quant_value=5,
                                            task="sell", task_size=90)}
    _rollout_parity(oracle, H.with_agents(mac, agents, [1, 2, 1, 1, 2, 1]), H.small_day(n_events=30000), B=32, steps=66,
                    seed=37)


def test_step_mm_only_single_data_message(oracle):
    """The shipped MM-only configs step ONE data message at a time (n_data_msg_per_step = 1): a 32-byte data slice, N = 5."""
    import dataclasses
    mac = H.load_mac("2_player_fq_fqc", n_data_msg_per_step=1, episode_time=50, start_resolution=50)   # as mm_bobRL.json, shorter
    agents = {"MarketMaking": dataclasses.replace(mac.dict_of_agents_configs["MarketMaking"], action_space="bobRL", bob_v0=5)}
    ref, n_done = _rollout_parity(oracle, H.with_agents(mac, agents, [1]), H.small_day(n_events=6000), B=64, steps=80,
                                  seed=29, stress_actions=True)
    assert C.num_msgs_per_step(ref.cfg) == 5 and n_done >= 64


def test_step_sell_buy_all_option(oracle):
    """sell_buy_all_option=True (mm_env.py:1018-1024 in fixed_quants, :1144-1172 in simple): inventory-sized orders, the
    9-entry offset tables with negative offsets, out-of-range actions."""
    import dataclasses
    mac = H.load_mac("2_player_fq_fqc")
    agents = dict(mac.dict_of_agents_configs)
    mm, ex = agents["MarketMaking"], agents["Execution"]
    agents = {"MarketMaking": dataclasses.replace(mm, sell_buy_all_option=True, fixed_quant_value=3),
              "Simple": dataclasses.replace(mm, short_name="SI", action_space="simple", n_actions=4, sell_buy_all_option=True,
                                            fixed_quant_value=5, reward_function="spooner"),
              "Execution": ex}
    _rollout_parity(oracle, H.with_agents(mac, agents, [2, 2, 1]), H.small_day(n_events=30000), B=48, steps=66, seed=19,
                    stress_actions=True)


def test_step_fixed_prices_vector_actions(oracle):
    """EXE fixed_prices (exec_env.py:1001): the action is a vector of quantities per price level ([B,n_i,n_actions])."""
    import dataclasses
    mac = H.load_mac("2_player_fq_fqc")
    agents = dict(mac.dict_of_agents_configs)
    ex = agents["Execution"]
    agents["Execution"] = dataclasses.replace(ex, action_space="fixed_prices", n_actions=4, fixed_quant_value=11, task_size=120)
    agents["Exec2"] = dataclasses.replace(ex, short_name="EXE2", action_space="fixed_prices", n_actions=1, fixed_quant_value=6,
                                          observation_space="simplest_case", task="sell", task_size=90)
    _rollout_parity(oracle, H.with_agents(mac, agents, [1, 2, 2]), H.small_day(n_events=30000), B=32, steps=66, seed=13)


@pytest.mark.parametrize("no,nt,t4,fill", [(100, 100, 0, True), (24, 12, 0, True), (33, 7, 1, False), (64, 32, 2, True),
                                           (200, 64, 0, True)])
def test_replay_adversarial_streams_bit_exact(oracle, no, nt, t4, fill):
    """Inputs no market produces (-1 in any field, non-positive quantities / prices, INT32 extremes): the fast paths must
    bail out to the literal generic path exactly where the reference's array semantics differ."""
    rng = np.random.default_rng(no * 7 + nt)
    bc = _book_cfg(no, nt, type_4_interpretation=t4, check_book_fill=fill)
    B, T = 96, 600
    msgs = H.adversarial_messages(rng, B * T, bc)
    start = np.arange(B, dtype=np.int64) * T
    a0 = np.full((B, no, 6), -1, np.int32); b0 = a0.copy(); t0 = np.full((B, nt, 8), -1, np.int32)
    ra, rb, rt = a0.copy(), b0.copy(), t0.copy()
    rbest = np.zeros((B, 4), np.int32)
    oracle.replay(bc, ra, rb, rt, msgs, start, T, best_out=rbest)
    ga, gb, gt, gbest = H.cuda_replay(bc, a0, b0, t0, msgs, start, T, want_best=True)
    np.testing.assert_array_equal(ga, ra); np.testing.assert_array_equal(gb, rb)
    np.testing.assert_array_equal(gt, rt); np.testing.assert_array_equal(gbest, rbest)
    # and continuing from those (odd) states, in two more legs, with an ordinary stream
    msgs2 = H.random_messages(rng, B * 300, bc, price_lo=99_500, price_hi=100_500)
    start2 = np.arange(B, dtype=np.int64) * 300
    oracle.replay(bc, ra, rb, rt, msgs2, start2, 300)
    ga, gb, gt = H.cuda_replay(bc, ga, gb, gt, msgs2, start2, 300)
    np.testing.assert_array_equal(ga, ra); np.testing.assert_array_equal(gb, rb); np.testing.assert_array_equal(gt, rt)


@pytest.mark.parametrize("mode,no,nt,adversarial", [(2, 100, 100, False), (3, 100, 100, False), (3, 24, 12, False),
                                                    (2, 40, 16, True), (3, 200, 64, True)])
def test_replay_random_cancel_modes_bit_exact(oracle, mode, no, nt, adversarial):
    """cancel_mode 2 / 3 (job:142-164): a cancel that matches neither an order id nor initial liquidity falls on a uniformly
    chosen order at its price (holding at least its quantity; mode 3: then any at its price).  The uniform draws are inputs,
    so the choice -- cumsum / searchsorted as jax.random.choice -- is bit-exact."""
    rng = np.random.default_rng(mode * 100 + no)
    bc = _book_cfg(no, nt, cancel_mode=mode)
    B, T = 96, 600
    msgs = (H.adversarial_messages(rng, B * T, bc) if adversarial else
            H.random_messages(rng, B * T, bc, price_lo=99_500, price_hi=100_300, id_pool=4000))   # most cancels miss their id
    cu = (rng.integers(0, 2 ** 23, size=(B, T, 2)).astype(np.float32) / np.float32(2 ** 23))
    cu[::7, ::5] = 0.0                                      # the smallest draw picks the LAST candidate (r = total)
    start = np.arange(B, dtype=np.int64) * T
    a0 = np.full((B, no, 6), -1, np.int32); b0 = a0.copy(); t0 = np.full((B, nt, 8), -1, np.int32)
    ra, rb, rt = a0.copy(), b0.copy(), t0.copy()
    rbest = np.zeros((B, 4), np.int32)
    oracle.replay(bc, ra, rb, rt, msgs, start, T, best_out=rbest, cancel_u=cu, n_threads=8)
    ga, gb, gt, gbest = H.cuda_replay(bc, a0, b0, t0, msgs, start, T, want_best=True, cancel_u=cu)
    np.testing.assert_array_equal(ga, ra); np.testing.assert_array_equal(gb, rb)
    np.testing.assert_array_equal(gt, rt); np.testing.assert_array_equal(gbest, rbest)
    # the draws matter: the same stream under cancel_mode 1 ends in a different book
    bc1 = _book_cfg(no, nt, cancel_mode=1)
    xa, xb, xt = a0.copy(), b0.copy(), t0.copy()
    oracle.replay(bc1, xa, xb, xt, msgs, start, T, n_threads=8)
    assert not (np.array_equal(xa, ra) and np.array_equal(xb, rb))


@pytest.mark.parametrize("mode", [2, 3])
def test_step_random_cancel_modes(oracle, mode):
    """env.step under cancel_mode 2 / 3 on the capacity-stress day (many data cancels miss their order id)."""
    mac = H.load_mac("2_player_fq_fqc", nOrders=40, nTrades=24, cancel_mode=mode)
    _rollout_parity(oracle, mac, H.small_day(seed=9, n_events=30000, stress=True), B=48, steps=66, seed=30 + mode)


def test_step_on_adversarial_day(oracle):
    """env.step with a day tensor made of adversarial messages: the per-message best bid/ask rows, rewards and
    observations go through the same bail-out logic as the replay."""
    import dataclasses
    mac = H.load_mac("2_player_fq_fqc", nOrders=48, nTrades=20)
    ld = H.load_for(mac, H.small_day(n_events=30000))
    rng = np.random.default_rng(21)
    bc = C.book_config(mac.world_config)
    adv = H.adversarial_messages(rng, ld.msgs.shape[0], bc, tick=100)
    adv[:, 3] = np.where(adv[:, 3] > 90_000, adv[:, 3] + 1_400_000, adv[:, 3])      # around the day's price level
    keep = rng.random(ld.msgs.shape[0]) < 0.5                                         # half real flow, half adversarial
    msgs = np.where(keep[:, None], ld.msgs, adv).astype(np.int32)
    ld2 = dataclasses.replace(ld, msgs=np.ascontiguousarray(msgs))
    B = 64
    ref = H.OracleEnv(oracle, mac, ld2, B)
    gpu = H.CudaEnv(mac, ld2, B, ref.params)
    H.draw_prng(rng, ref.cfg, ref.arrays); gpu.set_inputs(ref.arrays)
    ref.reset(); gpu.reset()
    for s in range(66):
        H.draw_prng(rng, ref.cfg, ref.arrays); H.draw_actions(rng, ref.cfg, ref.arrays)
        gpu.set_inputs(ref.arrays)
        ref.step(n_threads=8); gpu.step()
        try:
            H.assert_arrays_match(ref.arrays, gpu.numpy(), ref.cfg)
        except AssertionError as e:
            raise AssertionError(f"step {s}: {e}") from None


# ---- BASELINE.json sizes, CUDA vs the oracle (OpenMP over books / envs on the box's host cores) ----------------------
def _bench_day(mac):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return lobster.load_or_generate(mac.world_config, seed=20220103, n_events=400_000, cache_dir=os.path.join(root, ".cache"))


def _full_size_rollout(oracle, mac, B, steps, seed):
    """env.reset + ``steps`` x env.step on B environments of the bench's synthetic day: every state, output and info leaf of
    every environment after EVERY step (ints bit-exact, floats rel 1e-5)."""
    ld = _bench_day(mac)
    ref = H.OracleEnv(oracle, mac, ld, B)
    gpu = H.CudaEnv(mac, ld, B, ref.params)
    threads = oracle.max_threads()
    rng = np.random.default_rng(seed)
    H.draw_prng(rng, ref.cfg, ref.arrays)
    gpu.set_inputs(ref.arrays)
    ref.reset(); gpu.reset()
    H.assert_arrays_match(ref.arrays, gpu.numpy(), ref.cfg)
    n_done = 0
    for s in range(steps):
        H.draw_prng(rng, ref.cfg, ref.arrays)
        H.draw_actions(rng, ref.cfg, ref.arrays)
        gpu.set_inputs(ref.arrays)
        gpu.step(); ref.step(n_threads=threads)      # (the launch is asynchronous: both sides work at the same time)
        try:
            H.assert_arrays_match(ref.arrays, gpu.numpy(), ref.cfg)
        except AssertionError as e:
            raise AssertionError(f"step {s}: {e}") from None
        n_done += int(ref.arrays["done_all"].sum())
    return ref, n_done


def test_baseline_config2_exec_8192_envs(oracle):
    """BASELINE configs[2] AS WRITTEN: single execution agent (exec_longrun_fixed_quants_complex), NUM_ENVS = 8192,
    a whole 64-step episode plus the first step after the auto-reset."""
    mac = H.load_mac("exec_longrun_fixed_quants_complex")
    ref, n_done = _full_size_rollout(oracle, mac, 8192, 65, seed=21)
    assert n_done == 8192


def test_baseline_config3_2player_16384_envs(oracle):
    """BASELINE configs[3] shapes at the bench's per-GPU batch: 2_player_fq_fqc, 16384 envs, 66 steps (crosses the
    auto-reset: done on the 64th step, quirk Q14)."""
    mac = H.load_mac("2_player_fq_fqc")
    ref, n_done = _full_size_rollout(oracle, mac, 16384, 66, seed=22)
    assert n_done == 16384


def test_baseline_config1_replay_16384_books(oracle):
    """BASELINE configs[1] AS WRITTEN: 16384 books x 6400 messages of the bench's day, from the windows' reset states:
    books, trade logs and best prices against the oracle's replay (105 M messages on the host cores)."""
    mac = H.load_mac("2_player_fq_fqc")
    ld = _bench_day(mac)
    bc = C.book_config(mac.world_config)
    params = E.build_reset_params(ld, mac.world_config, H.oracle_replay_fn(oracle, bc))
    B, W, M, T = 16384, ld.starts.shape[0], ld.msgs.shape[0], 6400
    widx = np.arange(B) % W
    a0, b0, t0 = params["init_asks"][widx].copy(), params["init_bids"][widx].copy(), params["init_trades"][widx].copy()
    start = ((ld.starts[widx].astype(np.int64) + (np.arange(B) // W) % 100) % (M - T)).astype(np.int64)   # bench.py's offsets
    ra, rb, rt = a0.copy(), b0.copy(), t0.copy()
    rbest = np.zeros((B, 4), np.int32)
    oracle.replay(bc, ra, rb, rt, ld.msgs, start, T, best_out=rbest, n_threads=oracle.max_threads())
    ga, gb, gt, gbest = H.cuda_replay(bc, a0, b0, t0, ld.msgs, start, T, want_best=True)
    np.testing.assert_array_equal(ga, ra); np.testing.assert_array_equal(gb, rb)
    np.testing.assert_array_equal(gt, rt); np.testing.assert_array_equal(gbest, rbest)


def test_grouped_replay_variant_bit_exact(oracle):
    """The measurement variant lob_replay_launch_grouped (4 books per warp) gives the same books as the oracle."""
    import torch
    rng = np.random.default_rng(77)
    bc = _book_cfg(100, 100)
    B, T = 70, 900
    msgs = H.random_messages(rng, B * T, bc)
    start = np.arange(B, dtype=np.int64) * T
    a0 = np.full((B, 100, 6), -1, np.int32); b0 = a0.copy(); t0 = np.full((B, 100, 8), -1, np.int32)
    ra, rb, rt = a0.copy(), b0.copy(), t0.copy()
    oracle.replay(bc, ra, rb, rt, msgs, start, T)
    dev = torch.device("cuda:0")
    ta, tb, tt = (torch.from_numpy(x).to(dev) for x in (a0, b0, t0))
    E.replay_books(bc, ta, tb, tt, torch.from_numpy(msgs).to(dev), torch.from_numpy(start).to(dev), T, grouped=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ta.cpu().numpy(), ra); np.testing.assert_array_equal(tb.cpu().numpy(), rb)
    np.testing.assert_array_equal(tt.cpu().numpy(), rt)


def test_launch_follows_the_envs_device(oracle):
    """an env on cuda:1 while the current device is 0 must launch on GPU 1 (needs two GPUs), and the C entry
    point refuses buffers of another device instead of faulting."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from jaxmarl_hft_b200 import _lib
    mac = H.load_mac("2_player_fq_fqc")
    ld = H.load_for(mac, H.small_day(n_events=30000))
    torch.cuda.set_device(0)
    env = E.MARLEnv(None, mac, num_envs=64, loaded=ld, device="cuda:1", seed=5)
    obs, state = env.reset(None, env.default_params)
    acts = [torch.zeros((64, 1), dtype=torch.int32, device="cuda:1") for _ in range(2)]
    env.step(None, state, acts, env.default_params)
    torch.cuda.synchronize(1)
    assert state.arrays["asks"].device.index == 1 and int(state.arrays["step_counter"].max()) == 1
    bufs = states.pack_buffers(env.cfg, state.arrays, env.base_env.device_params())
    rc = _lib.lib().lob_step_launch(ctypes.byref(env.cfg), ctypes.byref(bufs), 64, _lib.current_stream_ptr())   # current device 0
    assert rc == abi.LOB_E_INVALID and b"device" in _lib.lib().lob_last_error()


def test_step_deep_book_window_and_second_pass(oracle):
    """Deep books run on a 128-row shared-memory window; a book that outgrows it is redone at full capacity by the second
    pass of the same lob_step_launch.  A 200-row book on the capacity-stress day: both passes are exercised (some
    environments fit the window, some do not), every leaf equals the oracle's after every step."""
    mac = H.load_mac("2_player_fq_fqc", nOrders=200, nTrades=128)
    ld = H.load_for(mac, H.small_day(seed=9, n_events=30000, stress=True))
    B = 96
    ref = H.OracleEnv(oracle, mac, ld, B)
    gpu = H.CudaEnv(mac, ld, B, ref.params)
    assert "work_redo_list" in gpu.arrays
    rng = np.random.default_rng(6)
    H.draw_prng(rng, ref.cfg, ref.arrays)
    gpu.set_inputs(ref.arrays)
    ref.reset(); gpu.reset()
    redone = []
    for s in range(70):
        H.draw_prng(rng, ref.cfg, ref.arrays)
        H.draw_actions(rng, ref.cfg, ref.arrays)
        gpu.set_inputs(ref.arrays)
        ref.step(n_threads=8); gpu.step()
        got = gpu.numpy()
        H.assert_arrays_match(ref.arrays, got, ref.cfg)
        redone.append(int(got["work_redo_count"][0]))
    live = np.maximum((ref.arrays["asks"] != -1).any(axis=2).sum(axis=1), (ref.arrays["bids"] != -1).any(axis=2).sum(axis=1))
    assert 0 < max(redone) and min(redone) < B, (redone, live.max())
    assert int(got["work_redo_count"][1]) == sum(redone)
