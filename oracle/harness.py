"""The environment API over numpy arrays, computed by the CPU oracle (TEST INFRASTRUCTURE -- only tests/,
__graft_entry__.smoke() and bench.py's CPU legs import this; the product package never does).

``OracleEnv`` mirrors jaxmarl_hft_b200.env.MARLEnv's buffer handling; ``draw_prng`` / ``draw_actions`` are the seeded
stand-ins of the jax.random products the reference draws inside a step (marl_env.py:294-295, base_env.py:222-225,
exec_env.py:221) and of a random policy."""
import numpy as np

from jaxmarl_hft_b200 import abi, config as C, env as E, states


def oracle_replay_fn(oracle, book_cfg):
    """The reset-state precompute (base_env.py:245-296) through the oracle's replay."""
    book_cfg = E.limit_only_book_config(book_cfg)

    def fn(asks, bids, trades, msgs, start, n_msgs):
        oracle.replay(book_cfg, asks, bids, trades, msgs, start, n_msgs)
        return asks, bids, trades
    return fn


class OracleEnv:
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
    """Seeded stand-ins for the jax.random products (perm, reset window, is_sell, the random-cancel uniforms)."""
    B = arrays["asks"].shape[0]
    arrays["reset_window"][:] = rng.integers(0, cfg.n_windows, size=B)
    arrays["reset_is_sell"][:] = rng.integers(0, 2, size=arrays["reset_is_sell"].shape)
    n_act = C.num_action_msgs(cfg)
    if n_act:
        arrays["perm"][:] = np.argsort(rng.random((B, n_act)), axis=1)
    if "cancel_u" in arrays:   # jax.random.uniform: multiples of 2^-23 in [0, 1)
        arrays["cancel_u"][:] = rng.integers(0, 2 ** 23, size=arrays["cancel_u"].shape).astype(np.float32) / np.float32(2 ** 23)


def draw_actions(rng, cfg, arrays):
    for t in range(cfg.n_agent_types):
        a = cfg.agent[t]
        hi = a.fixed_quant_value if abi.action_width(a) > 1 else a.n_actions   # fixed_prices: a vector of quantities
        arrays[f"actions{t}"][:] = rng.integers(0, hi, size=arrays[f"actions{t}"].shape)
