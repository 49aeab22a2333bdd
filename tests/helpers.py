"""Shared test scaffolding: configs, a numpy env driven by the CPU oracle, seeded PRNG products."""
import dataclasses
import os

import numpy as np

from jaxmarl_hft_b200 import abi, config as C, env as E, lobster, states

CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "jaxmarl-hft_b200", "configs")


def load_mac(name, **world_overrides) -> C.MultiAgentConfig:
    mac = C.load_config_from_file(os.path.join(CONFIG_DIR, name + ".json"))
    if world_overrides:
        mac = C.MultiAgentConfig(world_config=dataclasses.replace(mac.world_config, **world_overrides),
                                 dict_of_agents_configs=mac.dict_of_agents_configs,
                                 number_of_agents_per_type=mac.number_of_agents_per_type)
    return mac


def with_agents(mac, agents: dict, n_per_type) -> C.MultiAgentConfig:
    return C.MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents,
                              number_of_agents_per_type=list(n_per_type))


_DAY_CACHE = {}


def small_day(seed=5, n_events=30000, stress=False, levels=10):
    key = (seed, n_events, stress, levels)
    if key not in _DAY_CACHE:
        _DAY_CACHE[key] = lobster.generate_day(seed=seed, n_events=n_events, stress=stress, levels=levels)
    return _DAY_CACHE[key]


def load_for(mac, day):
    w = mac.world_config
    return lobster.load_days([day], w.episode_time, w.n_data_msg_per_step, w.start_resolution, w.day_start, w.day_end)


def oracle_replay_fn(oracle, book_cfg):
    def fn(asks, bids, trades, msgs, start, n_msgs):
        oracle.replay(book_cfg, asks, bids, trades, msgs, start, n_msgs)
        return asks, bids, trades
    return fn


class OracleEnv:
    """MARLEnv over numpy arrays, computed by the oracle.  Mirrors jaxmarl_hft_b200.env.MARLEnv's buffer handling."""

    def __init__(self, oracle, mac, loaded, num_envs):
        self.oracle, self.mac, self.loaded, self.B = oracle, mac, loaded, num_envs
        w = mac.world_config
        self.book_cfg = C.book_config(w)
        self.params = E.build_reset_params(loaded, w, oracle_replay_fn(oracle, self.book_cfg))
        self.cfg = C.to_step_config(mac, loaded.starts.shape[0], loaded.msgs.shape[0])
        self.arrays = states.alloc_numpy(self.cfg, num_envs)

    def reset(self):
        self.oracle.reset(self.cfg, self.arrays, self.params)

    def step(self, n_threads=1):
        self.oracle.step(self.cfg, self.arrays, self.params, n_threads)


def draw_prng(rng, cfg, arrays):
    """Seeded stand-ins for the jax.random products (perm, reset window, is_sell)."""
    B = arrays["asks"].shape[0]
    arrays["reset_window"][:] = rng.integers(0, cfg.n_windows, size=B)
    arrays["reset_is_sell"][:] = rng.integers(0, 2, size=arrays["reset_is_sell"].shape)
    n_act = C.num_action_msgs(cfg)
    if n_act:
        arrays["perm"][:] = np.argsort(rng.random((B, n_act)), axis=1)


def draw_actions(rng, cfg, arrays):
    for t in range(cfg.n_agent_types):
        arrays[f"actions{t}"][:] = rng.integers(0, cfg.agent[t].n_actions, size=arrays[f"actions{t}"].shape)


def copy_inputs(src, dst):
    for k in src:
        if k.startswith("actions") or k in ("perm", "reset_window", "reset_is_sell"):
            dst[k][...] = src[k]


def assert_arrays_match(ref: dict, got: dict, cfg, rtol=1e-5, atol=1e-6, float_exact=False, skip=()):
    """ints / bytes bit-exact; floats bit-exact where ``float_exact`` else rel 1e-5 (north_star's tolerance).
    NaN == NaN (market_share is 0/0 when nothing traded, mm:2408)."""
    for k, r in ref.items():
        if k in skip:
            continue
        g = got[k]
        if r.dtype.kind in "iu":
            np.testing.assert_array_equal(g, r, err_msg=k)
        elif float_exact:
            np.testing.assert_array_equal(g.view(np.uint32) if False else g, r, err_msg=k)
        else:
            np.testing.assert_allclose(g, r, rtol=rtol, atol=atol, equal_nan=True, err_msg=k)


def to_numpy(arrays):
    return {k: v.detach().cpu().numpy() for k, v in arrays.items()}


def empty_book(n_orders=100, n_trades=100):
    return (np.full((n_orders, 6), -1, np.int32), np.full((n_orders, 6), -1, np.int32),
            np.full((n_trades, 8), -1, np.int32))
