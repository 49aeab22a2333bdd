"""The CPU oracle (and, on a GPU box, the CUDA path) against golden vectors produced by executing the UNMODIFIED reference
sources under the NumPy JAX emulation (tests/golden/make_golden.py, tests/golden/jaxshim).

What is pinned: job.scan_through_entire_array_save_bidask + get_L2_state on adversarial streams (full books, eviction,
the -1 wrap, trade-log overflow, IOC / LIM / MKT type-4 interpretation), and MARLEnv.reset / MARLEnv.step rollouts through
the reference's own loader, reset-state precompute, agent message construction, rewards, observations, info dicts and
auto-reset.  Integer leaves are compared bit for bit.  Float leaves: rel 1e-5 (north_star) on EVERY leaf, with no exception
list: the oracle and the CUDA path sum the trade log left to right in row order, the order the golden vectors were produced
with, so even the EXE quantities the reference computes by cancelling two ~1e7-sized float32 products (advantage / drift /
reward and their running means, exe:1627-1665) are reproduced (see DESIGN.md "Float parity").
"""
import ast
import glob
import os

import numpy as np
import pytest

import helpers as H
from jaxmarl_hft_b200 import abi, config as C, env as E, lobster, states

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REPLAY_CASES = sorted(glob.glob(os.path.join(GOLDEN, "replay_*.npz")))
ENV_CASES = sorted(glob.glob(os.path.join(GOLDEN, "env_*.npz")))

def _book_cfg(z):
    return C.book_config(C.World_EnvironmentConfig(nOrders=int(z["no"]), nTrades=int(z["nt"]),
                                                   type_4_interpretation=int(z["t4"]), check_book_fill=bool(z["fill"]),
                                                   cancel_mode=int(z["cancel_mode"]) if "cancel_mode" in z.files else 1))


@pytest.mark.parametrize("path", REPLAY_CASES, ids=[os.path.basename(p) for p in REPLAY_CASES])
def test_oracle_replay_matches_reference(oracle, path):
    z = np.load(path)
    B, T, no, nt = int(z["B"]), int(z["T"]), int(z["no"]), int(z["nt"])
    bc = _book_cfg(z)
    for b in range(B):
        a = np.full((no, 6), -1, np.int32); d = a.copy(); t = np.full((nt, 8), -1, np.int32)
        cu = np.ascontiguousarray(z["cancel_u"][b]) if "cancel_u" in z.files else None
        ba, bb = oracle.scan_save_bidask(bc, a, d, t, z["msgs"][b * T:(b + 1) * T], cancel_u=cu)
        np.testing.assert_array_equal(a, z["asks"][b]); np.testing.assert_array_equal(d, z["bids"][b])
        np.testing.assert_array_equal(t, z["trades"][b])
        np.testing.assert_array_equal(ba, z["best_asks"][b]); np.testing.assert_array_equal(bb, z["best_bids"][b])
    np.testing.assert_array_equal(oracle.l2(bc, z["asks"], z["bids"], 10), z["l2"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", REPLAY_CASES, ids=[os.path.basename(p) for p in REPLAY_CASES])
def test_cuda_replay_matches_reference(path):
    import torch
    z = np.load(path)
    B, T, no, nt = int(z["B"]), int(z["T"]), int(z["no"]), int(z["nt"])
    bc = _book_cfg(z)
    a = np.full((B, no, 6), -1, np.int32); d = a.copy(); t = np.full((B, nt, 8), -1, np.int32)
    ga, gb, gt, best = H.cuda_replay(bc, a, d, t, z["msgs"], np.arange(B, dtype=np.int64) * T, T, want_best=True,
                                     cancel_u=z["cancel_u"] if "cancel_u" in z.files else None)
    np.testing.assert_array_equal(ga, z["asks"]); np.testing.assert_array_equal(gb, z["bids"])
    np.testing.assert_array_equal(gt, z["trades"])
    # the reference's last per-message best pair is the unfilled get_best_bid_and_ask_inclQuants of the final book
    np.testing.assert_array_equal(best[:, 0:2], z["best_asks"][:, -1]); np.testing.assert_array_equal(best[:, 2:4], z["best_bids"][:, -1])
    l2 = E.l2_state(bc, torch.from_numpy(ga).cuda(), torch.from_numpy(gb).cuda(), 10).cpu().numpy()
    np.testing.assert_array_equal(l2, z["l2"])


# ---- env rollouts -------------------------------------------------------------------------------------------------
def _parse_trace(lines):
    out = []
    for s in lines:
        fn, caller, args, res = str(s).split("|")
        out.append((fn, caller, ast.literal_eval(args), ast.literal_eval(res)))
    return out


def _set_draws(arrays, trace, B, n_act, T, kinds, n_agents=None, random_task=None, window_selector=-1):
    """One vmapped call under the shim runs env by env: split its trace into B groups."""
    n_agents = n_agents or [1] * T
    random_task = random_task or [kinds[t] == abi.AGENT_EXE for t in range(T)]  # exec_env.py:220-223: only task="random" draws
    n_exe = sum(n_agents[t] for t in range(T) if random_task[t])   # one draw per agent, same key per type (Q9)
    per = len(trace) // B
    assert per * B == len(trace)
    for e in range(B):
        grp = trace[e * per:(e + 1) * per]
        perms = [r for fn, c, a, r in grp if fn == "permutation"]
        win = [r for fn, c, a, r in grp if fn == "randint" and c.startswith("reset_env:224")]
        sell = [r for fn, c, a, r in grp if fn == "randint" and not c.startswith("reset_env:224")]
        if perms:
            arrays["perm"][e, :n_act] = perms[0]
        assert len(win) == 1 and len(sell) == n_exe, (grp,)
        arrays["reset_window"][e] = win[0] if window_selector == -1 else window_selector   # base:222-225
        it = iter(sell)
        for t in range(T):
            if random_task[t]:
                draws = [next(it) for _ in range(n_agents[t])]
                assert len(set(draws)) == 1    # marl_env.py:187: all agents of a type share the reset key
                arrays["reset_is_sell"][e, t] = draws[0]
            else:
                arrays["reset_is_sell"][e, t] = 0


_STATE_ALIAS = {}


def _compare(z, prefix, arrays, cfg, what):
    """Every golden leaf under ``prefix`` against our buffer table: ints bit for bit, floats rel 1e-5 (atol 1e-6)."""
    T = cfg.n_agent_types
    errs = []

    def chk(name, got, ref):
        ref = np.asarray(ref)
        got = np.asarray(got).reshape(ref.shape) if np.asarray(got).size == ref.size else np.asarray(got)
        if ref.dtype.kind in "iub":
            if not np.array_equal(got.astype(np.int64), ref.astype(np.int64)):
                errs.append(f"{what} {name}: int mismatch at {np.argwhere(got.astype(np.int64) != ref.astype(np.int64))[:4].tolist()} "
                            f"got {got.ravel()[:6]} ref {ref.ravel()[:6]}")
        else:
            ok = np.isclose(got, ref, rtol=1e-5, atol=1e-6, equal_nan=True)
            if not ok.all():
                bad = np.argwhere(~ok)[0]
                errs.append(f"{what} {name}: float mismatch at {bad.tolist()} got {got[tuple(bad)]!r} ref {ref[tuple(bad)]!r}")

    for k in z.files:
        if not k.startswith(prefix):
            continue
        leaf = k[len(prefix):]
        if leaf.startswith("state/"):
            name = leaf[6:]
            if name in arrays:
                chk(name, arrays[name], z[k])
        elif leaf.startswith("obs"):
            chk(leaf, arrays[leaf], z[k])
        elif leaf.startswith("reward") or leaf.startswith("done_agents"):
            chk(leaf, arrays[leaf], z[k])
        elif leaf == "done_all":
            chk(leaf, arrays["done_all"], z[k])
    return errs


def _compare_info(z, prefix, arrays, cfg):
    errs = []
    wi, wf = arrays["info_world_i32"], arrays["info_world_f32"]
    world = {k: wi[:, j] for j, k in enumerate(abi.WINFO_I32)}
    world.update({k: wf[:, j] for j, k in enumerate(abi.WINFO_F32)})
    for key in z.files:
        if key.startswith(prefix + "winfo/"):
            name = key.split("/")[-1]
            ref = z[key]
            got = np.stack([world["time_s"], world["time_ns"]], 1) if name == "time" else world[name]
            if ref.dtype.kind in "iub":
                if not np.array_equal(got.astype(np.int64), ref.astype(np.int64)):
                    errs.append(f"winfo {name}: got {got} ref {ref}")
            elif not np.allclose(got, ref, rtol=1e-5, atol=1e-6):
                errs.append(f"winfo {name}: got {got} ref {ref}")
    for t in range(cfg.n_agent_types):
        ki, kf = abi.info_cols(cfg.agent[t].kind)
        d = {k: arrays[f"info_i32_{t}"][..., j] for j, k in enumerate(ki)}
        d.update({k: arrays[f"info_f32_{t}"][..., j] for j, k in enumerate(kf)})
        for key in z.files:
            if key.startswith(f"{prefix}info{t}/"):
                name = key.split("/")[-1]
                ref = z[key]
                got = d[name].reshape(ref.shape)
                if ref.dtype.kind in "iub":
                    if not np.array_equal(got.astype(np.int64), ref.astype(np.int64)):
                        errs.append(f"info{t} {name}: got {got.ravel()} ref {ref.ravel()}")
                else:
                    if not np.isclose(got, ref, rtol=1e-5, atol=1e-6, equal_nan=True).all():
                        errs.append(f"info{t} {name}: got {got.ravel()} ref {ref.ravel()}")
    return errs


def _mutate(mac, how):
    """The same agent sets tests/golden/make_golden.py builds for the reference (MUTATORS there), in our config classes."""
    import dataclasses
    d = dict(mac.dict_of_agents_configs)
    mm, ex = d["MarketMaking"], d["Execution"]
    if how == "hetero":
        agents = {
            "MarketMaking": dataclasses.replace(mm, observation_space="engineered", reward_function="spooner",
                                                reference_price="far_touch", unwind_price="far_touch", inv_penalty="quadratic",
                                                fixed_quant_value=5, clip_reward=True, exclude_extreme_spreads=True,
                                                volume_traded_bonus="market_share", unwind_price_penalty=3),
            "Execution": dataclasses.replace(ex, reference_price="far_touch", reward_lambda=0.5, task_size=300, task="random"),
            "Directional": dataclasses.replace(mm, short_name="DIR", action_space="directional_trading",
                                               observation_space="basic", reward_function="delta_portfolio_value",
                                               reference_price="mid_avg", unwind_price="mid_avg", fixed_quant_value=7),
        }
        return H.with_agents(mac, agents, [2, 2, 1])
    if how == "mm_complex":
        agents = {
            "MarketMaking": dataclasses.replace(mm, reward_function="complex", reference_price="near_touch", unwind_price="mid",
                                                inv_penalty="threshold", inv_penalty_threshold=2.0, auto_liquidate_threshold=3,
                                                observation_space="engineered", normalize=False, fixed_quant_value=4),
            "Execution": dataclasses.replace(ex, observation_space="basic", reward_function="finish_fast", task="sell",
                                             normalize=False, task_size=80),
        }
        return H.with_agents(mac, agents, [2, 1])
    if how == "simple_skew_avst":
        agents = {
            "MarketMaking": dataclasses.replace(mm, action_space="simple", n_actions=4, fixed_quant_value=2),
            "Skew": dataclasses.replace(mm, short_name="SK", action_space="spread_skew", multiplier_type="spread",
                                        fixed_quant_value=3, reward_function="spooner"),
            "AvSt": dataclasses.replace(mm, short_name="AV", action_space="AvSt", observation_space="engineered",
                                        fixed_quant_value=4),
            "Execution": ex,
        }
        return H.with_agents(mac, agents, [1, 1, 2, 1])
    if how == "fixed_prices":
        agents = {
            "MarketMaking": mm,
            "Execution": dataclasses.replace(ex, action_space="fixed_prices", n_actions=4, fixed_quant_value=11, task_size=120),
            "Exec2": dataclasses.replace(ex, short_name="EXE2", action_space="fixed_prices", n_actions=2, fixed_quant_value=6,
                                         task="sell", task_size=90),
        }
        return H.with_agents(mac, agents, [1, 2, 1])
    if how == "bob_twap":
        agents = {"MarketMaking": dataclasses.replace(mm, action_space="bobRL", bob_v0=2, fixed_quant_value=3),
                  "Execution": dataclasses.replace(ex, action_space="twap", task_size=200)}
        return H.with_agents(mac, agents, [2, 1])
    if how == "bobstrat_1msg":
        agents = {
            "MarketMaking": dataclasses.replace(mm, action_space="bobStrategy", bob_v0=5, observation_space="engineered"),
            "Execution": dataclasses.replace(ex, action_space="fixed_quants_1msg", task_size=60, fixed_quant_value=7),
            "Exec2": dataclasses.replace(ex, short_name="EXE2", action_space="simplest_case", reward_function="simplest_case",
                                         observation_space="simplest_case", task="buy", task_size=40, fixed_quant_value=9),
        }
        return H.with_agents(mac, agents, [1, 2, 1])
    if how == "sell_buy_all":
        agents = {
            "MarketMaking": dataclasses.replace(mm, sell_buy_all_option=True, fixed_quant_value=3,
                                                observation_space="engineered"),
            "Simple": dataclasses.replace(mm, short_name="SI", action_space="simple", n_actions=4, sell_buy_all_option=True,
                                          fixed_quant_value=2),
            "Simple3": dataclasses.replace(mm, short_name="S3", action_space="simple", n_actions=3,
                                           simple_nothing_action=False, sell_buy_all_option=True, fixed_quant_value=4),
            "Execution": ex,
        }
        return H.with_agents(mac, agents, [2, 1, 1, 1])
    if how == "fixed_time":
        agents = {
            "MarketMaking": dataclasses.replace(mm, observation_space="engineered"),
            "AvSt": dataclasses.replace(mm, short_name="AV", action_space="AvSt", observation_space="engineered",
                                        normalize=False, fixed_quant_value=4),
            "Execution": ex,
            "Exec2": dataclasses.replace(ex, short_name="EXE2", observation_space="simplest_case", normalize=False,
                                         task="buy", task_size=70),
        }
        return H.with_agents(mac, agents, [1, 1, 2, 1])
    raise KeyError(how)


def _setup(z):
    name = str(z["json"]).replace(".json", "")
    overrides = dict(ast.literal_eval(str(z["world_overrides"]))) if "world_overrides" in z.files else {}
    mac = H.load_mac(name, **overrides)
    if "mutate" in z.files and str(z["mutate"]):
        mac = _mutate(mac, str(z["mutate"]))
    day = lobster.generate_day(seed=int(z["day_seed"]), n_events=int(z["n_events"]), stress=bool(int(z["stress"])))
    ld = H.load_for(mac, day)
    return mac, ld


def _rollout(z, env, step_fn, reset_fn, get_arrays, set_inputs):
    cfg = env.cfg
    overrides = dict(ast.literal_eval(str(z["world_overrides"]))) if "world_overrides" in z.files else {}
    wsel = int(overrides.get("window_selector", -1))
    B, steps, T = int(z["B"]), int(z["steps"]), cfg.n_agent_types
    kinds = [cfg.agent[t].kind for t in range(T)]
    n_act = C.num_action_msgs(cfg)
    inp = states.alloc_numpy(cfg, B)
    n_agents = [cfg.agent[t].n_agents for t in range(T)]
    rnd = [cfg.agent[t].kind == abi.AGENT_EXE and cfg.agent[t].task == abi.EXE_TASKS["random"] for t in range(T)]
    _set_draws(inp, _parse_trace(z["reset_trace"]), B, n_act, T, kinds, n_agents, rnd, wsel)
    set_inputs(inp)
    reset_fn()
    errs = _compare(z, "reset/", get_arrays(), cfg, "reset")
    assert not errs, "\n".join(errs[:10])
    for s in range(steps):
        _set_draws(inp, _parse_trace(z[f"step{s}/trace"]), B, n_act, T, kinds, n_agents, rnd, wsel)
        for t in range(T):
            inp[f"actions{t}"][...] = z[f"step{s}/actions{t}"]
        if "cancel_u" in inp:
            inp["cancel_u"][...] = z[f"step{s}/cancel_u"]
        set_inputs(inp)
        step_fn()
        arr = get_arrays()
        errs = _compare(z, f"step{s}/", arr, cfg, f"step {s}") + _compare_info(z, f"step{s}/", arr, cfg)
        assert not errs, f"step {s}:\n" + "\n".join(errs[:12])


@pytest.mark.parametrize("path", ENV_CASES, ids=[os.path.basename(p) for p in ENV_CASES])
def test_oracle_env_matches_reference(oracle, path):
    z = np.load(path)
    mac, ld = _setup(z)
    env = H.OracleEnv(oracle, mac, ld, int(z["B"]))
    assert env.cfg.n_windows == int(z["n_windows"])
    _rollout(z, env, env.step, env.reset, lambda: env.arrays, lambda inp: H.copy_inputs(inp, env.arrays))


@pytest.mark.gpu
@pytest.mark.parametrize("path", ENV_CASES, ids=[os.path.basename(p) for p in ENV_CASES])
def test_cuda_env_matches_reference(oracle, path):
    z = np.load(path)
    mac, ld = _setup(z)
    bc = C.book_config(mac.world_config)
    params = E.build_reset_params(ld, mac.world_config, E._cuda_replay_fn(bc, "cuda:0"))
    gpu = H.CudaEnv(mac, ld, int(z["B"]), params)
    _rollout(z, gpu, gpu.step, gpu.reset, gpu.numpy, gpu.set_inputs)
