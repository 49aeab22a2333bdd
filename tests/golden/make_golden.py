#!/usr/bin/env python
"""Generate golden vectors by executing the UNMODIFIED reference sources (/root/reference/gymnax_exchange) under the
NumPy-backed JAX emulation in tests/golden/jaxshim (jax / jaxlib are not installable offline).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

For each case it builds the reference's own ``MARLEnv`` from the reference's own env JSON (paths / stock pointed at a
synthetic LOBSTER CSV pair written by our generator and read by the reference's own pandas loader), then runs
``vmap(env.reset)`` and ``vmap(env.step)`` exactly as the trainer does (ippo_rnn_JAXMARL.py:571,616) with recorded
random actions, and dumps per step: every state leaf, obs, rewards, dones, the info dicts, and the PRNG products the
reference drew inside the step (action-message permutation, reset window, is_sell_task) so that the oracle / CUDA path
receive the same draws as inputs.  Also dumps job.scan_through_entire_array cases (pure replay) and get_L2_state.

This script needs /root/reference; the tests that consume the .npz files do not.
"""
import dataclasses
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, REFERENCE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402

assert getattr(jax, "SHIM", False), "this script must import the NumPy-backed jax shim"

import jaxmarl_hft_b200  # noqa: E402,F401
from jaxmarl_hft_b200 import lobster  # noqa: E402


def _np(x):
    return np.asarray(x.view(np.ndarray) if isinstance(x, np.ndarray) else x)


def write_day(tmp, seed, n_events, stress):
    day = lobster.generate_day(seed=seed, n_events=n_events, stress=stress)
    lobster.write_lobster_csv(day, os.path.join(tmp, "data", "rawLOBSTER", "GOOG", "2022"))
    return day


def reference_config(json_name, tmp, **world_overrides):
    from gymnax_exchange.jaxob.config_io import load_config_from_file
    from gymnax_exchange.jaxob.jaxob_config import MultiAgentConfig
    mac = load_config_from_file(os.path.join(REFERENCE, "config", "env_configs", json_name))
    world = dataclasses.replace(mac.world_config, dataPath=os.path.join(tmp, "data"), alphatradePath=os.path.join(tmp, "at"),
                                stock="GOOG", timePeriod="2022", use_pickles_for_init=False, **world_overrides)
    return MultiAgentConfig(world_config=world, dict_of_agents_configs=mac.dict_of_agents_configs,
                            number_of_agents_per_type=mac.number_of_agents_per_type)


def flatten_state(state, n_types, kinds):
    """MultiAgentState (batched) -> dict of numpy arrays with the leaf names of jaxmarl_hft_b200.states."""
    w = state.world_state
    out = {
        "asks": _np(w.ask_raw_orders), "bids": _np(w.bid_raw_orders), "trades": _np(w.trades),
        "init_time": _np(w.init_time), "window_index": _np(w.window_index), "max_steps": _np(w.max_steps_in_episode),
        "start_index": _np(w.start_index), "step_counter": _np(w.step_counter), "best_bids": _np(w.best_bids),
        "best_asks": _np(w.best_asks), "time": _np(w.time), "order_id_counter": _np(w.order_id_counter),
        "mid_price": _np(w.mid_price), "delta_time": _np(w.delta_time),
    }
    for t in range(n_types):
        a = state.agent_states[t]
        for f in dataclasses.fields(a):
            out[f"a{t}_{f.name}"] = _np(getattr(a, f.name))
    return out


def collect_prng(trace, B, n_act, n_types):
    """Split the recorded draws of ONE vmapped call into per-env arrays.  Under the shim a vmapped call runs env by env,
    so the trace is B consecutive groups."""
    perms, windows, sells = [], [], []
    for fn, caller, key, args, res in trace:
        if fn == "permutation":
            perms.append(np.asarray(res, np.int32))
        elif fn == "randint" and caller.startswith("reset_env") and args[2] != 2:
            windows.append(int(res))
        elif fn == "randint":
            sells.append(int(res))
    return perms, windows, sells


def hetero_agents(mac):
    """3 agent types for the reference: MM fixed_quants (engineered obs, spooner reward, far-touch reference),
    EXE fixed_quants_complex (far_touch doom price, lambda 0.5), directional (MM class, directional_trading)."""
    from gymnax_exchange.jaxob.jaxob_config import MultiAgentConfig
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    agents = {
        "MarketMaking": dataclasses.replace(mm, observation_space="engineered", reward_function="spooner",
                                            reference_price="far_touch", unwind_price="far_touch", inv_penalty="quadratic",
                                            fixed_quant_value=5, clip_reward=True, exclude_extreme_spreads=True,
                                            volume_traded_bonus="market_share", unwind_price_penalty=3),
        "Execution": dataclasses.replace(ex, reference_price="far_touch", reward_lambda=0.5, task_size=300, task="random"),
        "Directional": dataclasses.replace(mm, short_name="DIR", action_space="directional_trading", observation_space="basic",
                                           reward_function="delta_portfolio_value", reference_price="mid_avg",
                                           unwind_price="mid_avg", fixed_quant_value=7),
    }
    return MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents, number_of_agents_per_type=[2, 2, 1])


def mm_complex_agents(mac):
    from gymnax_exchange.jaxob.jaxob_config import MultiAgentConfig
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    agents = {
        "MarketMaking": dataclasses.replace(mm, reward_function="complex", reference_price="near_touch", unwind_price="mid",
                                            inv_penalty="threshold", inv_penalty_threshold=2.0, auto_liquidate_threshold=3,
                                            observation_space="engineered", normalize=False, fixed_quant_value=4),
        "Execution": dataclasses.replace(ex, observation_space="basic", reward_function="finish_fast", task="sell",
                                         normalize=False, task_size=80),
    }
    return MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents, number_of_agents_per_type=[2, 1])


def bob_twap_agents(mac):
    from gymnax_exchange.jaxob.jaxob_config import MultiAgentConfig
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    agents = {
        "MarketMaking": dataclasses.replace(mm, action_space="bobRL", bob_v0=2, fixed_quant_value=3),
        "Execution": dataclasses.replace(ex, action_space="twap", task_size=200),
    }
    return MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents, number_of_agents_per_type=[2, 1])


def bobstrat_1msg_agents(mac):
    from gymnax_exchange.jaxob.jaxob_config import MultiAgentConfig
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    agents = {
        "MarketMaking": dataclasses.replace(mm, action_space="bobStrategy", bob_v0=5, observation_space="engineered"),
        "Execution": dataclasses.replace(ex, action_space="fixed_quants_1msg", task_size=60, fixed_quant_value=7),
        "Exec2": dataclasses.replace(ex, short_name="EXE2", action_space="simplest_case", reward_function="simplest_case",
                                     observation_space="simplest_case", task="buy", task_size=40, fixed_quant_value=9),
    }
    return MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents, number_of_agents_per_type=[1, 2, 1])


def simple_skew_avst_agents(mac):
    from gymnax_exchange.jaxob.jaxob_config import MultiAgentConfig
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    agents = {
        "MarketMaking": dataclasses.replace(mm, action_space="simple", n_actions=4, fixed_quant_value=2),
        "Skew": dataclasses.replace(mm, short_name="SK", action_space="spread_skew", multiplier_type="spread", fixed_quant_value=3,
                                    reward_function="spooner"),
        "AvSt": dataclasses.replace(mm, short_name="AV", action_space="AvSt", observation_space="engineered", fixed_quant_value=4),
        "Execution": ex,
    }
    return MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents, number_of_agents_per_type=[1, 1, 2, 1])


def fixed_prices_agents(mac):
    from gymnax_exchange.jaxob.jaxob_config import MultiAgentConfig
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    agents = {
        "MarketMaking": mm,
        "Execution": dataclasses.replace(ex, action_space="fixed_prices", n_actions=4, fixed_quant_value=11, task_size=120),
        "Exec2": dataclasses.replace(ex, short_name="EXE2", action_space="fixed_prices", n_actions=2, fixed_quant_value=6,
                                     task="sell", task_size=90),
    }
    return MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents, number_of_agents_per_type=[1, 2, 1])


def fixed_time_agents(mac):
    """fixed_time episodes: both engineered observation variants (10 / 15 fields), AvSt (time-based horizon) and the
    simplest_case EXE observation."""
    from gymnax_exchange.jaxob.jaxob_config import MultiAgentConfig
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    agents = {
        "MarketMaking": dataclasses.replace(mm, observation_space="engineered"),
        "AvSt": dataclasses.replace(mm, short_name="AV", action_space="AvSt", observation_space="engineered", normalize=False,
                                    fixed_quant_value=4),
        "Execution": ex,
        "Exec2": dataclasses.replace(ex, short_name="EXE2", observation_space="simplest_case", normalize=False, task="buy",
                                     task_size=70),
    }
    return MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents,
                            number_of_agents_per_type=[1, 1, 2, 1])


def sell_buy_all_agents(mac):
    """sell_buy_all_option=True in the two action spaces that read it (mm_env.py:1018-1024, :1144-1172)."""
    from gymnax_exchange.jaxob.jaxob_config import MultiAgentConfig
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    agents = {
        "MarketMaking": dataclasses.replace(mm, sell_buy_all_option=True, fixed_quant_value=3, observation_space="engineered"),
        "Simple": dataclasses.replace(mm, short_name="SI", action_space="simple", n_actions=4, sell_buy_all_option=True,
                                      fixed_quant_value=2),
        "Simple3": dataclasses.replace(mm, short_name="S3", action_space="simple", n_actions=3, simple_nothing_action=False,
                                       sell_buy_all_option=True, fixed_quant_value=4),
        "Execution": ex,
    }
    return MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents, number_of_agents_per_type=[2, 1, 1, 1])


MUTATORS = {"sell_buy_all": sell_buy_all_agents, "fixed_time": fixed_time_agents, "fixed_prices": fixed_prices_agents, "simple_skew_avst": simple_skew_avst_agents, "hetero": hetero_agents, "mm_complex": mm_complex_agents, "bob_twap": bob_twap_agents,
            "bobstrat_1msg": bobstrat_1msg_agents}


def run_env_case(name, json_name, seed, B, steps, n_events=30000, stress=False, out_dir=HERE, mutate=None, **world_overrides):
    from gymnax_exchange.jaxen.marl_env import MARLEnv
    import jax.random as jr
    with tempfile.TemporaryDirectory() as tmp:
        write_day(tmp, seed=5 if not stress else 9, n_events=n_events, stress=stress)
        mac = reference_config(json_name, tmp, **world_overrides)
        if mutate:
            mac = MUTATORS[mutate](mac)
        env = MARLEnv(jr.PRNGKey(seed), mac)
        params = env.default_params
        n_types = len(env.instance_list)
        n_per = list(mac.number_of_agents_per_type)
        rng = np.random.default_rng(seed)
        key = jr.PRNGKey(seed + 1)
        key, k_reset = jr.split(key)
        reset_keys = jr.split(k_reset, B)
        jr.TRACE.clear()
        obs, state = jax.vmap(env.reset, in_axes=(0, None))(reset_keys, params)
        reset_trace = list(jr.TRACE)
        record = {"B": np.int64(B), "steps": np.int64(steps), "json": np.array(json_name),
                  "day_seed": np.int64(5 if not stress else 9), "n_events": np.int64(n_events), "stress": np.int64(stress),
                  "world_overrides": np.array(repr(sorted(world_overrides.items()))), "mutate": np.array(mutate or "")}
        record["reset_trace"] = np.array([f"{fn}|{caller}|{args}|{np.asarray(res).tolist()}" for fn, caller, k, args, res in reset_trace])
        st0 = flatten_state(state, n_types, None)
        for k2, v in st0.items():
            record[f"reset/state/{k2}"] = v
        for t in range(n_types):
            record[f"reset/obs{t}"] = _np(obs[t])
        for s in range(steps):
            key, k_step, k_act = jr.split(key, 3)
            step_keys = jr.split(k_step, B)
            actions = []
            for t in range(n_types):
                sp = env.action_spaces[t]
                if np.ndim(sp.n) == 0:
                    n_a = int(sp.n)
                    a = rng.integers(0, n_a, size=(B, n_per[t])).astype(np.int32)
                    if s % 9 == 4:
                        a[::3] = rng.integers(-2, n_a + 3, size=a[::3].shape)   # out-of-range actions (gather wrap + clamp)
                else:   # MultiDiscrete (EXE fixed_prices): a vector of quantities per agent
                    cats = np.asarray(sp.num_categories)
                    a = rng.integers(0, cats, size=(B, n_per[t], len(cats))).astype(np.int32)
                actions.append(a)
            # the trainer squeezes single-agent types (marl_env.py:265-266 handles both)
            jax_actions = [jnp.asarray(a[:, 0] if a.shape[1] == 1 else a) for a in actions]
            jr.TRACE.clear()
            obs, state, rewards, dones, info = jax.vmap(env.step, in_axes=(0, 0, 0, None))(step_keys, state, jax_actions, params)
            trace = list(jr.TRACE)
            record[f"step{s}/trace"] = np.array([f"{fn}|{caller}|{args}|{np.asarray(res).tolist()}" for fn, caller, k, args, res in trace
                                                 if fn != "choice"])
            if int(getattr(mac.world_config, "cancel_mode", 1)) >= 2:
                record[f"step{s}/cancel_u"] = choice_draws(trace, (B, env.num_msgs_per_step))
            for t in range(n_types):
                record[f"step{s}/actions{t}"] = actions[t]
                record[f"step{s}/obs{t}"] = _np(obs[t])
                record[f"step{s}/reward{t}"] = _np(rewards[t])
                record[f"step{s}/done_agents{t}"] = _np(dones["agents"][t])
                for k2, v in info["agents"][t].items():
                    record[f"step{s}/info{t}/{k2}"] = _np(v)
            record[f"step{s}/done_all"] = _np(dones["__all__"])
            for k2, v in info["world"].items():
                record[f"step{s}/winfo/{k2}"] = _np(v)
            for k2, v in flatten_state(state, n_types, None).items():
                record[f"step{s}/state/{k2}"] = v
        record["n_windows"] = np.int64(env.base_env.n_windows)
        record["type_names"] = np.array(env.type_names)
        path = os.path.join(out_dir, f"{name}.npz")
        np.savez_compressed(path, **record)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def choice_draws(trace, shape):
    """The uniform draws of the random-cancel fallbacks (job:146, :161) recorded by the shim's jax.random.choice ->
    float32 ``shape + (2,)`` indexed by (outermost vmap index if shape has two dims, message index)."""
    cu = np.zeros(tuple(shape) + (2,), np.float32)
    for fn, caller, key, pos, u in trace:
        if fn != "choice":
            continue
        stage = 1 if caller.startswith("get_random_large_id_match") else 0
        assert caller.startswith("get_random"), caller
        idx = (pos[1],) if len(shape) == 1 else (pos[0], pos[1])
        cu[idx + (stage,)] = u
    return cu


def run_replay_case(name, seed, B, T, no, nt, t4, fill, out_dir=HERE, adversarial=False, cancel_mode=1):
    """job.scan_through_entire_array_save_bidask on adversarial random streams + job.get_L2_state of the result."""
    import helpers as H
    from gymnax_exchange.jaxob import JaxOrderBookArrays as job
    from gymnax_exchange.jaxob.jaxob_config import JAXLOB_Configuration
    from jaxmarl_hft_b200 import config as C
    import jax.random as jr
    cfg = JAXLOB_Configuration(nOrders=no, nTrades=nt, type_4_interpretation=t4, check_book_fill=fill, cancel_mode=cancel_mode)
    bc = C.book_config(C.World_EnvironmentConfig(nOrders=no, nTrades=nt, type_4_interpretation=t4, check_book_fill=fill,
                                                 cancel_mode=cancel_mode))
    rng = np.random.default_rng(seed)
    msgs = (H.adversarial_messages(rng, B * T, bc) if adversarial else
            H.random_messages(rng, B * T, bc, price_lo=99_000, price_hi=100_600 if no < 64 else 101_500))
    rec = {"msgs": msgs, "B": np.int64(B), "T": np.int64(T), "no": np.int64(no), "nt": np.int64(nt), "t4": np.int64(t4),
           "fill": np.int64(fill), "cancel_mode": np.int64(cancel_mode)}
    A, Bd, Tr, BA, BB, L2, CU = [], [], [], [], [], [], []
    for b in range(B):
        jr.TRACE.clear()
        asks = job.init_orderside(no)
        bids = job.init_orderside(no)
        trades = (jnp.ones((nt, 8)) * -1).astype(jnp.int32)
        (asks, bids, trades), (ba, bb) = job.scan_through_entire_array_save_bidask(
            cfg, jr.PRNGKey(0), jnp.asarray(msgs[b * T:(b + 1) * T]), (asks, bids, trades), T)
        A.append(_np(asks)); Bd.append(_np(bids)); Tr.append(_np(trades)); BA.append(_np(ba)); BB.append(_np(bb))
        L2.append(_np(job.get_L2_state(asks, bids, 10, cfg)))
        CU.append(choice_draws(list(jr.TRACE), (T,)))
    rec.update(asks=np.stack(A), bids=np.stack(Bd), trades=np.stack(Tr), best_asks=np.stack(BA), best_bids=np.stack(BB),
               l2=np.stack(L2))
    if cancel_mode >= 2:
        rec["cancel_u"] = np.stack(CU)
        print("  random-cancel draws used:", int((rec["cancel_u"] != 0).sum()))
    path = os.path.join(out_dir, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def run_orderbook_case(name="orderbook_api", out_dir=HERE):
    """The reference's OrderBook object (jaxob/jorderbook.py): its own __main__ scenario (jorderbook.py:288-318) and the
    query methods on a book driven by a random stream."""
    import helpers as H
    from gymnax_exchange.jaxob.jorderbook import OrderBook
    from gymnax_exchange.jaxob.jaxob_config import JAXLOB_Configuration
    from jaxmarl_hft_b200 import config as C
    ob = OrderBook(JAXLOB_Configuration(maxint=2147483647, nOrders=64, nTrades=32))
    l2init = np.array([354200, 452, 350100, 89, 361200, 100, 344000, 400, 362900, 100, 343100, 100, 364000, 400, 338700, 100,
                       371900, 1100, 337100, 1000, 372200, 100, 336400, 1000, 372300, 200, 336000, 300, 372800, 1000, 333600,
                       1000, 374600, 1000, 332500, 100, 376700, 100, 331600, 100], np.int32)
    rec = {"l2init": l2init, "no": np.int64(64), "nt": np.int64(32)}

    def put(prefix, st):
        rec[prefix + "asks"], rec[prefix + "bids"], rec[prefix + "trades"] = _np(st.asks), _np(st.bids), _np(st.trades)

    state = ob.reset(jnp.asarray(l2init))
    put("reset/", state)
    quote = {"type": "limit", "side": "bid", "quantity": 99, "price": 346000, "trade_id": 8888, "order_id": 8888,
             "timestamp": "3400.005000000"}
    put("dict/", ob.process_order(state, quote))
    put("market/", ob.process_order(state, dict(quote, type="market", price=355000, quantity=500)))
    put("cancel/", ob.process_order(state, dict(quote, type="cancel", price=344000, quantity=150, order_id=-5)))
    msgs2 = np.array([[1, 1, 99, 346000, 8888, 8888, 3400, 5000000], [1, -1, 2, 346000, 8777, 8777, 3401, 5060000]], np.int32)
    put("one/", ob.process_order_array(state, jnp.asarray(msgs2[0])))
    st2, l2s = ob.process_orders_array_l2(state, jnp.asarray(msgs2), 10)
    put("two/", st2)
    rec["two/l2"] = _np(l2s)
    rec["q/vol_init"] = _np(ob.get_volume_at_price(state, 1, 344000, True))
    rec["q/vol"] = _np(ob.get_volume_at_price(st2, 1, 346000, False))
    rec["q/next_bid"], rec["q/next_ask"] = _np(ob.get_next_executable_order(state, 1)), _np(ob.get_next_executable_order(state, 0))
    rec["q/best_bid"], rec["q/best_ask"] = _np(ob.get_best_price(state, 1)), _np(ob.get_best_price(state, 0))
    ba, bb = ob.get_best_bid_and_ask_inclQuants(state)
    rec["q/best_ask_q"], rec["q/best_bid_q"] = _np(ba), _np(bb)
    # a livelier book: a random stream on top
    bc = C.book_config(C.World_EnvironmentConfig(nOrders=64, nTrades=32))
    stream = H.random_messages(np.random.default_rng(31), 500, bc, price_lo=340_000, price_hi=365_000, tick=100)
    rec["stream"] = stream
    st3 = ob.process_orders_array(state, jnp.asarray(stream))
    put("stream/", st3)
    rec["stream/l2"] = _np(ob.get_L2_state(st3, 7))
    for side in (0, 1):
        arr = _np(st3.bids if side == 1 else st3.asks)
        rec[f"stream/ids{side}"] = _np(ob.get_side_ids(st3, side))
        live = arr[arr[:, 0] != -1]
        probe = [int(live[0, 2]), int(live[-1, 2]), 123456789]                 # two present ids, one absent
        rec[f"stream/probe_ids{side}"] = np.array(probe, np.int64)
        rec[f"stream/order{side}"] = np.stack([_np(ob.get_order(st3, side, i)) for i in probe])
        rec[f"stream/order_p{side}"] = np.stack([_np(ob.get_order(st3, side, i, int(live[0, 0]))) for i in probe])
        times = [(int(live[0, 4]), int(live[0, 5])), (int(live[-1, 4]), int(live[-1, 5])), (1, 2)]
        rec[f"stream/probe_times{side}"] = np.array(times, np.int64)
        rec[f"stream/at_time{side}"] = np.stack([_np(ob.get_order_at_time(st3, side, a, b)) for a, b in times])
        rec[f"stream/at_time_p{side}"] = np.stack([_np(ob.get_order_at_time(st3, side, a, b, int(live[-1, 0]))) for a, b in times])
        rec[f"stream/next{side}"] = _np(ob.get_next_executable_order(st3, side))
        rec[f"stream/vol{side}"] = np.array([int(_np(ob.get_volume_at_price(st3, side, int(p), False))) for p in live[:5, 0]], np.int64)
        rec[f"stream/vol_prices{side}"] = live[:5, 0].astype(np.int64)
    path = os.path.join(out_dir, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["replay", "env"]
    if "replay" in which:
        run_replay_case("replay_100", seed=11, B=6, T=500, no=100, nt=100, t4=0, fill=True)
        run_replay_case("replay_small_full", seed=12, B=6, T=400, no=20, nt=8, t4=0, fill=True)
        run_replay_case("replay_t4lim_nofill", seed=13, B=4, T=300, no=24, nt=16, t4=1, fill=False)
        run_replay_case("replay_mkt", seed=14, B=4, T=300, no=32, nt=16, t4=2, fill=True)
    if "replay_adv" in which:
        run_replay_case("replay_adversarial", seed=15, B=6, T=400, no=24, nt=12, t4=0, fill=True, adversarial=True)
        run_replay_case("replay_adversarial_100", seed=16, B=4, T=500, no=100, nt=100, t4=1, fill=True, adversarial=True)
    if "orderbook" in which:
        run_orderbook_case()
    if "replay_cnl" in which:
        run_replay_case("replay_cancel_uniform", seed=17, B=6, T=400, no=24, nt=12, t4=0, fill=True, cancel_mode=2)
        run_replay_case("replay_cancel_uniform_large", seed=18, B=6, T=500, no=100, nt=100, t4=0, fill=True, cancel_mode=3)
        run_replay_case("replay_cancel_large_adversarial", seed=19, B=4, T=400, no=33, nt=16, t4=1, fill=True,
                        adversarial=True, cancel_mode=3)
    if "env" in which:
        run_env_case("env_2player", "2_player_fq_fqc.json", seed=3, B=16, steps=66)
        run_env_case("env_exec", "exec_longrun_fixed_quants_complex.json", seed=4, B=16, steps=20)
    if "env2" in which:
        run_env_case("env_hetero_smallbook", "2_player_fq_fqc.json", seed=6, B=2, steps=66, stress=True, mutate="hetero",
                     nOrders=40, nTrades=24)
        run_env_case("env_mm_complex", "2_player_fq_fqc.json", seed=7, B=2, steps=66, mutate="mm_complex")
    if "env_ft" in which:
        run_env_case("env_fixed_time", "2_player_fq_fqc.json", seed=21, B=4, steps=60, mutate="fixed_time",
                     ep_type="fixed_time", episode_time=1800, start_resolution=900)
        # the last window's nominal start wraps to day_start (base:288-290), so every data message is past the end
        run_env_case("env_fixed_time_masked", "2_player_fq_fqc.json", seed=22, B=2, steps=24, mutate="fixed_time",
                     ep_type="fixed_time", episode_time=1800, start_resolution=900, window_selector=25)
    if "env_sba" in which:
        run_env_case("env_sell_buy_all", "2_player_fq_fqc.json", seed=24, B=2, steps=66, mutate="sell_buy_all")
    if "env_cnl" in which:
        run_env_case("env_cancel_uniform_large", "2_player_fq_fqc.json", seed=23, B=3, steps=40, stress=True, mutate="hetero",
                     nOrders=40, nTrades=24, cancel_mode=3)
    if "env3" in which:
        run_env_case("env_bob_twap", "2_player_fq_fqc.json", seed=8, B=2, steps=66, mutate="bob_twap")
        run_env_case("env_simple_skew_avst", "2_player_fq_fqc.json", seed=10, B=2, steps=66, mutate="simple_skew_avst")
        run_env_case("env_bobstrat_1msg", "2_player_fq_fqc.json", seed=9, B=2, steps=66, mutate="bobstrat_1msg")
