"""ctypes mirror of ``include/lobstep.h`` (the C-ABI of the LOB step).

Field order and types follow the header exactly; ``check_sizes`` compares ``ctypes.sizeof`` with the
``lob_sizeof_*`` exports of a loaded library so a drift between the two is caught at load time.
"""
import ctypes as C

LOB_ABI_VERSION = 11
LOB_MAX_AGENT_TYPES = 8
LOB_MAX_AGENT_I32 = 4
LOB_MAX_AGENT_F32 = 10

LOB_OK, LOB_E_INVALID, LOB_E_UNSUPPORTED, LOB_E_CUDA = 0, -1, -2, -3

# enums (values are part of the ABI)
AGENT_MM, AGENT_EXE = 0, 1
MM_ACTION_SPACES = {"fixed_quants": 0, "directional_trading": 1, "bobRL": 2, "bobStrategy": 3, "simple": 4, "spread_skew": 5,
                    "AvSt": 6}
EXE_ACTION_SPACES = {"fixed_quants": 0, "fixed_quants_complex": 1, "fixed_quants_1msg": 2, "simplest_case": 3, "twap": 4,
                     "fixed_prices": 5}
OBS_SPACES = {"engineered": 0, "basic": 1, "simplest_case": 2}
MM_REWARDS = {"portfolio_value": 0, "buy_sell_pnl": 1, "complex": 2, "zero_inv": 3, "spooner": 4,
              "spooner_damped": 5, "spooner_asym_damped": 6, "spooner_asym_damped2": 7, "spooner_scaled": 8,
              "delta_portfolio_value": 9}
EXE_REWARDS = {"normal": 0, "finish_fast": 1, "simplest_case": 2}
REF_PRICES = {"mid": 0, "mid_avg": 1, "far_touch": 2, "near_touch": 3}
INV_PENALTIES = {"none": 0, "linear": 1, "quadratic": 2, "exp4": 3, "threshold": 4}
EXE_TASKS = {"random": 0, "buy": 1, "sell": 2}

WINFO_I32 = ("window_index", "step_counter", "time_s", "time_ns", "order_id_counter", "best_asks", "best_bids",
             "current_step", "ep_done_time", "abort_episode", "spread")
WINFO_F32 = ("end_mid_price", "average_best_ask", "average_best_bid", "delta_time")
MMINFO_I32 = ("done", "inventory", "forced_unwind", "posted_bid_price", "posted_ask_price",
              "bid_distance_from_best", "ask_distance_from_best", "ask_quant", "bid_quant")
MMINFO_F32 = ("reward", "reward_portfolio_value", "reward_spooner", "end_of_ep_pv", "reward_spooner_damped",
              "reward_spooner_asym_damped", "reward_spooner_asym_damped2", "reward_delta_pv", "total_PnL",
              "delta_mid_price", "market_share", "buyPnL", "invPnL", "sellPnL", "inventoryValue")
EXEINFO_I32 = ("quant_left", "done", "doom_quant", "is_sell_task")
EXEINFO_F32 = ("revenue_direction_normalised", "vwap_rm", "drift", "advantage", "reward")

MM_STATE_I32 = ("posted_distance_bid", "posted_distance_ask", "inventory")
MM_STATE_F32 = ("total_PnL", "cash_balance")
EXE_STATE_I32 = ("task_to_execute", "quant_executed", "is_sell_task")
EXE_STATE_F32 = ("init_price", "p_vwap", "total_revenue", "drift_return", "advantage_return", "slippage_rm",
                 "price_adv_rm", "price_drift_rm", "vwap_rm", "trade_duration")

i32, i64, f64 = C.c_int32, C.c_int64, C.c_double
p_i32, p_f32, p_u8, p_i64 = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)


class LobBookConfig(C.Structure):
    _fields_ = [(n, i32) for n in ("n_orders", "n_trades", "maxint", "init_id", "book_depth", "cancel_mode",
                                   "type_4_interpretation", "check_book_fill")]


class LobAgentTypeConfig(C.Structure):
    _fields_ = [(n, i32) for n in (
        "kind", "n_agents", "trader_id_start", "action_space", "observation_space", "reward_function", "n_actions",
        "num_messages_by_agent", "num_action_messages_by_agent", "normalize", "time_delay_obs_act",
        "fixed_quant_value", "n_ticks_offset", "tenth_action_market_order", "sell_buy_all_option",
        "fixed_action_setting", "fixed_action", "auto_liquidate_threshold", "unwind_price_penalty", "inv_penalty",
        "volume_traded_bonus_market_share", "reference_price", "unwind_price", "clip_reward",
        "exclude_extreme_spreads", "task", "task_size", "n_ticks_in_book", "larger_far_touch_quant",
        "doom_price_penalty", "bob_v0", "simple_nothing_action", "multiplier_type_spread")] + [(n, f64) for n in (
            "auto_liquidate_alpha", "inv_penalty_lambda", "inv_penalty_quadratic_factor", "inv_penalty_threshold",
            "reward_scaling_quo", "inventoryPnL_eta", "inventoryPnL_gamma", "rebate_bps", "unrealizedPnL_lambda",
            "reward_lambda", "spread_multiplier", "skew_multiplier", "avst_k_parameter", "avst_var_parameter")]


class LobStepConfig(C.Structure):
    _fields_ = [("book", LobBookConfig)] + [(n, i32) for n in (
        "n_data_msg_per_step", "tick_size", "ep_type_fixed_time", "episode_time", "order_id_counter_start",
        "placeholder_order_id", "artificial_trader_id_end_episode", "artificial_order_id_end_episode",
        "shuffle_action_messages", "n_agent_types", "n_windows", "_pad0")] + [
            ("n_messages", i64), ("agent", LobAgentTypeConfig * LOB_MAX_AGENT_TYPES)]


class LobStepBuffers(C.Structure):
    _fields_ = [(n, p_i32) for n in ("asks", "bids", "trades", "init_time", "window_index", "max_steps",
                                     "start_index", "step_counter", "best_bids", "best_asks", "time",
                                     "order_id_counter")] + [
        ("mid_price", p_f32), ("delta_time", p_f32),
        ("agent_i32", (p_i32 * LOB_MAX_AGENT_I32) * LOB_MAX_AGENT_TYPES),
        ("agent_f32", (p_f32 * LOB_MAX_AGENT_F32) * LOB_MAX_AGENT_TYPES),
        ("actions", p_i32 * LOB_MAX_AGENT_TYPES),
        ("perm", p_i32), ("reset_window", p_i32), ("reset_is_sell", p_i32), ("cancel_u", p_f32),
        ("message_data", p_i32), ("init_asks", p_i32), ("init_bids", p_i32), ("init_trades", p_i32),
        ("init_init_time", p_i32), ("init_max_steps", p_i32), ("init_start_index", p_i32),
        ("obs", p_f32 * LOB_MAX_AGENT_TYPES), ("reward", p_f32 * LOB_MAX_AGENT_TYPES),
        ("done_all", p_u8), ("done_agents", p_u8 * LOB_MAX_AGENT_TYPES),
        ("info_world_i32", p_i32), ("info_world_f32", p_f32),
        ("info_agent_i32", p_i32 * LOB_MAX_AGENT_TYPES), ("info_agent_f32", p_f32 * LOB_MAX_AGENT_TYPES),
        ("work_redo_list", p_i32), ("work_redo_count", p_i32), ("work_split", p_i32)]


class LobRolloutBuffers(C.Structure):
    _fields_ = [("n_steps", i32), ("_pad0", i32), ("batch", i64),
                ("actions", p_i32 * LOB_MAX_AGENT_TYPES), ("perm", p_i32), ("reset_window", p_i32), ("reset_is_sell", p_i32),
                ("cancel_u", p_f32), ("obs", p_f32 * LOB_MAX_AGENT_TYPES), ("reward", p_f32 * LOB_MAX_AGENT_TYPES),
                ("done_agents", p_u8 * LOB_MAX_AGENT_TYPES), ("done_all", p_u8)]


class LobReplayBuffers(C.Structure):
    _fields_ = [("asks", p_i32), ("bids", p_i32), ("trades", p_i32), ("msgs", p_i32), ("start", p_i64),
                ("n_msgs_total", i64), ("n_msgs", i32), ("_pad0", i32), ("best_out", p_i32),
                ("cancel_u", p_f32)]


_SIZEOF = {"lob_sizeof_book_config": LobBookConfig, "lob_sizeof_agent_type_config": LobAgentTypeConfig,
           "lob_sizeof_step_config": LobStepConfig, "lob_sizeof_step_buffers": LobStepBuffers,
           "lob_sizeof_replay_buffers": LobReplayBuffers, "lob_sizeof_rollout_buffers": LobRolloutBuffers}


# the sentinel fields of lob_abi_offsets (include/lobstep.h), in its order
_SENTINELS = [(LobBookConfig, "cancel_mode"), (LobBookConfig, "check_book_fill"),
              (LobAgentTypeConfig, "fixed_quant_value"), (LobAgentTypeConfig, "task_size"),
              (LobAgentTypeConfig, "doom_price_penalty"), (LobAgentTypeConfig, "reward_scaling_quo"),
              (LobAgentTypeConfig, "reward_lambda"),
              (LobStepConfig, "tick_size"), (LobStepConfig, "episode_time"), (LobStepConfig, "n_agent_types"),
              (LobStepConfig, "n_messages"), (LobStepConfig, "agent"),
              (LobStepBuffers, "best_asks"), (LobStepBuffers, "mid_price"), (LobStepBuffers, "agent_f32"),
              (LobStepBuffers, "perm"), (LobStepBuffers, "message_data"), (LobStepBuffers, "obs"),
              (LobStepBuffers, "done_all"), (LobStepBuffers, "info_agent_f32"), (LobStepBuffers, "work_redo_count"),
              (LobReplayBuffers, "start"), (LobReplayBuffers, "n_msgs"), (LobReplayBuffers, "best_out"),
              (LobReplayBuffers, "cancel_u")]


def check_sizes(lib):
    """Raise if the loaded library's structs differ from this mirror: in size, or (the CUDA library) in the byte offset of
    the sentinel fields ``lob_abi_offsets`` reports -- two same-sized fields swapped on one side would pass the size check."""
    for fn, cls in _SIZEOF.items():
        f = getattr(lib, fn)
        f.restype = C.c_int64
        got = int(f())
        if got != C.sizeof(cls):
            raise RuntimeError(f"ABI mismatch: {fn}() = {got}, ctypes mirror = {C.sizeof(cls)}")
    if hasattr(lib, "lob_abi_offsets"):
        n = len(_SENTINELS)
        out = (C.c_int64 * n)()
        lib.lob_abi_offsets.argtypes = [C.POINTER(C.c_int64), C.c_int32]
        lib.lob_abi_offsets.restype = C.c_int32
        if int(lib.lob_abi_offsets(out, n)) != n:
            raise RuntimeError("ABI mismatch: lob_abi_offsets reports another number of sentinel fields")
        for (cls, name), got in zip(_SENTINELS, out):
            if getattr(cls, name).offset != int(got):
                raise RuntimeError(f"ABI mismatch: offsetof({cls.__name__}, {name}) = {int(got)} in the library, "
                                   f"{getattr(cls, name).offset} in the ctypes mirror")


SPLIT_ENV_WORDS, SPLIT_AGENT_WORDS = 24, 0     # csrc/lob_kernels.cuh kSplitEnvWords / kSplitAgentWords


def split_workspace_words(cfg, batch: int) -> int:
    """== lob_split_workspace_words: 32-bit words of ``LobStepBuffers.work_split``."""
    n = sum(cfg.agent[t].n_agents for t in range(cfg.n_agent_types))
    n_am = sum(cfg.agent[t].n_agents * cfg.agent[t].num_messages_by_agent for t in range(cfg.n_agent_types))
    # env records, then (the piped step, csrc/lob_pipe.cuh) the agents' cancel + action messages of every environment
    return int(batch) * (SPLIT_ENV_WORDS + SPLIT_AGENT_WORDS * n) + int(batch) * n_am * 8


def action_width(a) -> int:
    """Ints per agent in the actions buffer: n_actions for the EXE fixed_prices space (a vector of quantities), else 1."""
    return int(a.n_actions) if (a.kind == AGENT_EXE and a.action_space == EXE_ACTION_SPACES["fixed_prices"]) else 1


def obs_dim(kind, observation_space, fixed_time=False):
    """mm_env.py:3195-3223 ; exec_env.py:2188-2202 (the engineered spaces grow time fields under fixed_time)."""
    if kind == AGENT_MM:
        return 2 if observation_space == OBS_SPACES["basic"] else (10 if fixed_time else 8)
    return (15 if fixed_time else 12) if observation_space == OBS_SPACES["engineered"] else 3


def info_cols(kind):
    return (MMINFO_I32, MMINFO_F32) if kind == AGENT_MM else (EXEINFO_I32, EXEINFO_F32)


def state_leaves(kind):
    return (MM_STATE_I32, MM_STATE_F32) if kind == AGENT_MM else (EXE_STATE_I32, EXE_STATE_F32)
