"""Configuration objects of the LOB step -- the same *API surface* as the reference's frozen dataclasses
(gymnax_exchange/jaxob/jaxob_config.py:12-253): same class names, field names and defaults, so a
``MultiAgentConfig`` written for the reference constructs here unchanged.  The derived fields
(``n_actions``, ``num_messages_by_agent``, ``num_action_messages_by_agent``) are computed from the action
space exactly as the reference's ``__post_init__`` does (jaxob_config.py:98-141,175-200) -- values given in
JSON files for those three are overridden.

``to_step_config`` lowers a ``MultiAgentConfig`` to the POD ``LobStepConfig`` of include/lobstep.h.
"""
import dataclasses
import json
import os
from dataclasses import dataclass, field

from . import abi

MAXINT_32 = 2_147_483_647  # cst.MaxInt._64_Bit_Signed (sic) jaxob_constants.py:3-5
INITID = -2                # jaxob_constants.py:8
EMPTY_SLOT = -1            # jaxob_constants.py:11


def _set(obj, **kw):
    for k, v in kw.items():
        object.__setattr__(obj, k, v)


@dataclass(frozen=True)
class JAXLOB_Configuration:
    """jaxob_config.py:12-30"""
    maxint: int = MAXINT_32
    init_id: int = INITID
    book_depth: int = 10
    cancel_mode: int = 1             # CancelMode.INCLUDE_INITS
    type_4_interpretation: int = 0   # Type4Interpretation.IOC
    seed: int = 42
    nTrades: int = 100
    nOrders: int = 100
    simulator_mode: int = 0          # SimulatorMode.GENERAL_EXCHANGE
    empty_slot_val: int = EMPTY_SLOT
    debug_mode: bool = False
    check_book_fill: bool = True
    start_resolution: int = 6400
    alphatradePath: str = os.path.expanduser("~")
    dataPath: str = os.path.expanduser("~") + "/data"
    stock: str = "AMZN"
    timePeriod: str = "2024_Dec"


# (n_actions, num_messages_by_agent, num_action_messages_by_agent) per MM action space; None = keep the given value
_MM_DERIVED = {
    "spread_skew": (6, 4, 2), "bobStrategy": (5, 4, 2), "directional_trading": (3, 4, 2), "AvSt": (8, 4, 2),
}
_BOB_RL_ACTIONS = {1: 3, 2: 5, 5: 11, 10: 21}


@dataclass(frozen=True)
class MarketMaking_EnvironmentConfig:
    """jaxob_config.py:33-141"""
    debug_mode: bool = False
    short_name: str = "MM"
    normalize: bool = True
    clip_reward: bool = False
    exclude_extreme_spreads: bool = False
    fixed_action_setting: bool = False
    fixed_action: int = 0
    simple_nothing_action: bool = True
    sell_buy_all_option: bool = False
    based_on_mid_price_of_action: bool = True
    tenth_action: str = "MarketOrder"
    bob_v0: int = 1
    action_space: str = "bobRL"
    observation_space: str = "engineered"
    reward_function: str = "spooner_asym_damped2"
    spread_multiplier: float = 3.0
    skew_multiplier: float = 5.0
    n_ticks_offset: int = 1
    fixed_quant_value: int = 10
    auto_liquidate_threshold: int = 10000
    auto_liquidate_alpha: float = 1.0
    unwind_price_penalty: int = 5
    inv_penalty: str = "none"
    volume_traded_bonus: str = "none"
    reference_price: str = "mid"
    unwind_price: str = "mid"
    inv_penalty_lambda: float = 1.0
    inv_penalty_quadratic_factor: float = 50.0
    inv_penalty_threshold: float = 10.0
    multiplier_type: str = "tick"
    reward_scaling_quo: float = 1.0
    inventoryPnL_eta: float = 0.6
    inventoryPnL_gamma: float = 0.5
    rebate_bps: float = 10.0
    unrealizedPnL_lambda: float = 0.1
    avst_k_parameter: float = 0.4
    avst_var_parameter: float = 1e-8
    time_delay_obs_act: int = 0
    n_actions: int = 10
    num_messages_by_agent: int = 4
    num_action_messages_by_agent: int = 2

    def __post_init__(self):
        sp = self.action_space
        if sp == "fixed_quants":
            if self.tenth_action not in ("NA", "MarketOrder"):
                raise ValueError(f"Invalid tenth_action {self.tenth_action} for fixed_quants action space")
            _set(self, n_actions=9 if self.tenth_action == "NA" else 10, num_messages_by_agent=4,
                 num_action_messages_by_agent=2)
        elif sp == "bobRL":
            if self.bob_v0 not in _BOB_RL_ACTIONS:
                raise ValueError(f"Invalid bob_v0 {self.bob_v0} for bobRL action space")
            _set(self, n_actions=_BOB_RL_ACTIONS[self.bob_v0], num_messages_by_agent=4,
                 num_action_messages_by_agent=2)
        elif sp in _MM_DERIVED:
            n, m, a = _MM_DERIVED[sp]
            _set(self, n_actions=n, num_messages_by_agent=m, num_action_messages_by_agent=a)
        elif sp == "fixed_prices":
            _set(self, num_messages_by_agent=self.n_actions * 2, num_action_messages_by_agent=self.n_actions)


_EXE_DERIVED = {
    "fixed_quants": (5, 8, 4), "fixed_quants_complex": (13, 8, 4), "simplest_case": (3, 4, 2),
    "fixed_quants_1msg": (5, 2, 1), "twap": (1, 4, 2),
}


@dataclass(frozen=True)
class Execution_EnvironmentConfig:
    """jaxob_config.py:144-200"""
    debug_mode: bool = False
    larger_far_touch_quant: bool = False
    normalize: bool = True
    short_name: str = "EXE"
    action_type: str = "pure"
    task: str = "random"
    action_space: str = "fixed_quants_complex"
    observation_space: str = "engineered"
    reward_function: str = "normal"
    task_size: int = 600
    n_ticks_in_book: int = 1
    fixed_quant_value: int = 10
    reward_lambda: float = 0.0
    reward_scaling_quo: float = 1.0
    doom_price_penalty: int = 5
    reference_price: str = "mid"
    time_delay_obs_act: int = 0
    n_actions: int = 5
    num_messages_by_agent: int = 8
    num_action_messages_by_agent: int = 4

    def __post_init__(self):
        sp = self.action_space
        if sp in _EXE_DERIVED:
            n, m, a = _EXE_DERIVED[sp]
            _set(self, n_actions=n, num_messages_by_agent=m, num_action_messages_by_agent=a)
        elif sp == "fixed_prices":
            _set(self, num_messages_by_agent=self.n_actions * 2, num_action_messages_by_agent=self.n_actions)


@dataclass(frozen=True)
class World_EnvironmentConfig(JAXLOB_Configuration):
    """jaxob_config.py:205-225"""
    n_data_msg_per_step: int = 1
    window_selector: int = -1
    ep_type: str = "fixed_steps"
    episode_time: int = 6400
    day_start: int = 34200
    day_end: int = 57600
    tick_size: int = 100
    trader_id_range_start: int = -100
    placeholder_order_id: int = -198
    artificial_trader_id_end_episode: int = -199
    artificial_order_id_end_episode: int = -199
    any_message_obs_space: bool = False
    order_id_counter_start_when_resetting: int = -200
    shuffle_action_messages: bool = True
    use_pickles_for_init: bool = True
    save_raw_observations: bool = False


def _default_agents():
    return {"MarketMaking": MarketMaking_EnvironmentConfig(), "Execution": Execution_EnvironmentConfig()}


@dataclass(frozen=True)
class MultiAgentConfig:
    """jaxob_config.py:228-250"""
    world_config: World_EnvironmentConfig = World_EnvironmentConfig()
    dict_of_agents_configs: dict = field(default_factory=_default_agents)
    number_of_agents_per_type: list = field(default_factory=lambda: [1, 1])

    def __post_init__(self):
        for cfg in self.dict_of_agents_configs.values():
            if "message" in cfg.observation_space:
                _set(self.world_config, any_message_obs_space=True)


CONFIG_OBJECT_DICT = {"MarketMaking": MarketMaking_EnvironmentConfig, "Execution": Execution_EnvironmentConfig}


def _from_dict(cls, d):
    names = {f.name for f in dataclasses.fields(cls)}
    return cls(**{k: v for k, v in d.items() if k in names})


def load_config_from_file(path):
    """Read a reference-format env JSON (config/env_configs/*.json; jaxob/config_io.py:43-162).  Agent blocks
    with an unknown key are classified by their fields, as config_io.py:144-162 does."""
    with open(path) as f:
        raw = json.load(f)
    world = _from_dict(World_EnvironmentConfig, raw.get("world_config", {}))
    agents = {}
    for name, block in raw.get("dict_of_agents_configs", {}).items():
        cls = CONFIG_OBJECT_DICT.get(name)
        if cls is None:
            cls = Execution_EnvironmentConfig if "task_size" in block else MarketMaking_EnvironmentConfig
        agents[name] = _from_dict(cls, block)
    return MultiAgentConfig(world_config=world, dict_of_agents_configs=agents,
                            number_of_agents_per_type=list(raw.get("number_of_agents_per_type", [1] * len(agents))))


# ----------------------------------------------------------------------------------------------------------
# lowering to the POD struct of include/lobstep.h
# ----------------------------------------------------------------------------------------------------------

def _enum(table, value, what):
    if value not in table:
        raise ValueError(f"Invalid {what} specified: {value!r} (built here: {sorted(table)})")
    return table[value]


def book_config(world: JAXLOB_Configuration) -> abi.LobBookConfig:
    if world.simulator_mode != 0:
        raise ValueError("The simulator mode does not match an expected value.")  # job:620-622
    if world.cancel_mode not in (0, 1, 2, 3):
        raise ValueError(f"cancel_mode={world.cancel_mode}: not a cst.CancelMode value")
    return abi.LobBookConfig(n_orders=world.nOrders, n_trades=world.nTrades, maxint=world.maxint,
                             init_id=world.init_id, book_depth=world.book_depth, cancel_mode=world.cancel_mode,
                             type_4_interpretation=world.type_4_interpretation,
                             check_book_fill=int(world.check_book_fill))


def agent_type_config(cfg, n_agents: int, trader_id_start: int) -> abi.LobAgentTypeConfig:
    a = abi.LobAgentTypeConfig()
    a.n_agents = int(n_agents)
    a.trader_id_start = int(trader_id_start)
    a.n_actions = cfg.n_actions
    a.num_messages_by_agent = cfg.num_messages_by_agent
    a.num_action_messages_by_agent = cfg.num_action_messages_by_agent
    a.normalize = int(cfg.normalize)
    a.time_delay_obs_act = cfg.time_delay_obs_act
    a.fixed_quant_value = cfg.fixed_quant_value
    a.reward_scaling_quo = cfg.reward_scaling_quo
    a.observation_space = _enum(abi.OBS_SPACES, cfg.observation_space, "observation_space")
    if isinstance(cfg, MarketMaking_EnvironmentConfig):
        if cfg.observation_space == "simplest_case":
            raise ValueError("Invalid observation_space specified.")  # mm_env.py:2787
        a.kind = abi.AGENT_MM
        a.action_space = _enum(abi.MM_ACTION_SPACES, cfg.action_space, "action_space")
        a.reward_function = _enum(abi.MM_REWARDS, cfg.reward_function, "reward_space")
        a.n_ticks_offset = cfg.n_ticks_offset
        a.bob_v0 = cfg.bob_v0
        a.simple_nothing_action = int(cfg.simple_nothing_action)
        if cfg.multiplier_type not in ("tick", "spread"):
            raise ValueError(f"Invalid multiplier_type {cfg.multiplier_type!r}")
        a.multiplier_type_spread = int(cfg.multiplier_type == "spread")
        a.spread_multiplier = cfg.spread_multiplier
        a.skew_multiplier = cfg.skew_multiplier
        a.avst_k_parameter = cfg.avst_k_parameter
        a.avst_var_parameter = cfg.avst_var_parameter
        if cfg.action_space in ("bobRL", "bobStrategy") and cfg.bob_v0 not in _BOB_RL_ACTIONS and cfg.action_space == "bobRL":
            raise ValueError("cfg.bob_v0 must be one of [1,2,5,10]")  # mm:1522
        a.tenth_action_market_order = int(cfg.tenth_action == "MarketOrder")
        a.sell_buy_all_option = int(cfg.sell_buy_all_option)
        a.fixed_action_setting = int(cfg.fixed_action_setting)
        a.fixed_action = cfg.fixed_action
        a.auto_liquidate_threshold = cfg.auto_liquidate_threshold
        a.auto_liquidate_alpha = cfg.auto_liquidate_alpha
        a.unwind_price_penalty = cfg.unwind_price_penalty
        a.inv_penalty = _enum(abi.INV_PENALTIES, cfg.inv_penalty, "inventory penalty")
        a.volume_traded_bonus_market_share = int(cfg.volume_traded_bonus == "market_share")
        a.reference_price = _enum(abi.REF_PRICES, cfg.reference_price, "reference price type")
        if cfg.unwind_price not in ("mid", "mid_avg", "far_touch"):
            raise ValueError("Invalid unwind price type.")  # mm:2302-2303
        a.unwind_price = abi.REF_PRICES[cfg.unwind_price]
        a.clip_reward = int(cfg.clip_reward)
        a.exclude_extreme_spreads = int(cfg.exclude_extreme_spreads)
        a.inv_penalty_lambda = cfg.inv_penalty_lambda
        a.inv_penalty_quadratic_factor = cfg.inv_penalty_quadratic_factor
        a.inv_penalty_threshold = cfg.inv_penalty_threshold
        a.inventoryPnL_eta = cfg.inventoryPnL_eta
        a.inventoryPnL_gamma = cfg.inventoryPnL_gamma
        a.rebate_bps = cfg.rebate_bps
        a.unrealizedPnL_lambda = cfg.unrealizedPnL_lambda
    elif isinstance(cfg, Execution_EnvironmentConfig):
        a.kind = abi.AGENT_EXE
        a.action_space = _enum(abi.EXE_ACTION_SPACES, cfg.action_space, "action_space")
        a.reward_function = _enum(abi.EXE_REWARDS, cfg.reward_function, "reward_function")
        a.task = _enum(abi.EXE_TASKS, cfg.task, "task")
        a.task_size = cfg.task_size
        a.n_ticks_in_book = cfg.n_ticks_in_book
        a.larger_far_touch_quant = int(cfg.larger_far_touch_quant)
        if cfg.action_space == "fixed_quants_1msg" and cfg.larger_far_touch_quant:
            # exec_env.py:790 branches in Python on a traced action: the reference cannot trace this combination
            raise NotImplementedError("fixed_quants_1msg with larger_far_touch_quant=True does not trace in the reference")
        if not isinstance(cfg.doom_price_penalty, int):
            raise NotImplementedError("a non-integer doom_price_penalty turns exec_env.py:1564-1575 into float32 arithmetic: not built")
        a.doom_price_penalty = cfg.doom_price_penalty
        if cfg.reference_price not in ("mid", "far_touch"):
            raise ValueError("Invalid reference price type.")  # exe:1576-1580
        a.reference_price = abi.REF_PRICES[cfg.reference_price]
        a.reward_lambda = cfg.reward_lambda
        if cfg.action_space == "fixed_prices" and cfg.action_type not in ("delta", "pure"):
            raise ValueError("Invalid action_type specified.")  # exe:2168-2173
        if cfg.action_space == "fixed_prices" and not (1 <= cfg.n_actions <= 4):
            raise ValueError("fixed_prices supports 1 to 4 price levels (exec_env.py:1047-1054)")
    else:
        raise ValueError(f"Invalid agent type: {type(cfg).__name__}")  # marl:79
    return a


def to_step_config(mac: MultiAgentConfig, n_windows: int, n_messages: int) -> abi.LobStepConfig:
    w = mac.world_config
    if w.ep_type not in ("fixed_steps", "fixed_time"):
        raise NotImplementedError('Use either "fixed_time" or "fixed_steps"')      # ldr:998
    if w.any_message_obs_space or w.debug_mode:
        raise NotImplementedError("message-based observation spaces / debug_mode logging are not built")
    if w.save_raw_observations:
        raise NotImplementedError("save_raw_observations=True (info['agents'][i]['obs_raw'], marl_env.py:673-674) is not built")
    types = list(mac.dict_of_agents_configs.values())
    if len(types) != len(mac.number_of_agents_per_type):
        raise ValueError("number_of_agents_per_type must have one entry per agent config")
    if len(types) > abi.LOB_MAX_AGENT_TYPES:
        raise ValueError(f"at most {abi.LOB_MAX_AGENT_TYPES} agent types")
    if w.episode_time < 1:
        raise ValueError(f"episode_time={w.episode_time}: must be >= 1")
    if w.window_selector < -1:
        raise ValueError(f"window_selector={w.window_selector}: -1 (random) or a window index")   # base:222-225
    c = abi.LobStepConfig()
    c.book = book_config(w)
    c.n_data_msg_per_step = w.n_data_msg_per_step
    c.tick_size = w.tick_size
    c.ep_type_fixed_time = int(w.ep_type == "fixed_time")
    c.episode_time = w.episode_time
    c.order_id_counter_start = w.order_id_counter_start_when_resetting
    c.placeholder_order_id = w.placeholder_order_id
    c.artificial_trader_id_end_episode = w.artificial_trader_id_end_episode
    c.artificial_order_id_end_episode = w.artificial_order_id_end_episode
    c.shuffle_action_messages = int(w.shuffle_action_messages)
    c.n_agent_types = len(types)
    c.n_windows = int(n_windows)
    c.n_messages = int(n_messages)
    tid = w.trader_id_range_start  # marl:103-115: ids count down across types
    for i, (cfg, n) in enumerate(zip(types, mac.number_of_agents_per_type)):
        c.agent[i] = agent_type_config(cfg, n, tid)
        # the values the kernels divide by (mirrors check_step_cfg in csrc/lobstep.cu)
        if c.agent[i].reward_scaling_quo == 0:
            raise ValueError(f"{cfg.short_name}: reward_scaling_quo must not be 0")
        if c.agent[i].kind == abi.AGENT_MM and c.agent[i].sell_buy_all_option and c.agent[i].fixed_quant_value < 1:
            raise ValueError(f"{cfg.short_name}: sell_buy_all_option needs fixed_quant_value >= 1")   # mm:1018
        if c.agent[i].kind == abi.AGENT_EXE and c.agent[i].task_size < 1:
            raise ValueError(f"{cfg.short_name}: task_size must be >= 1")
        if (c.ep_type_fixed_time and c.agent[i].kind == abi.AGENT_EXE
                and c.agent[i].action_space == abi.EXE_ACTION_SPACES["twap"]):
            raise NotImplementedError("TWAP not implemented for fixed time episodes")   # exe:1141-1142
        tid -= n
    return c


CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


def load_named_config(name: str, **world_overrides) -> MultiAgentConfig:
    """One of the env JSONs shipped under ``configs/`` (the reference's config/env_configs/*.json), with optional
    overrides of world-config fields."""
    mac = load_config_from_file(os.path.join(CONFIG_DIR, name + ".json"))
    if world_overrides:
        mac = MultiAgentConfig(world_config=dataclasses.replace(mac.world_config, **world_overrides),
                               dict_of_agents_configs=mac.dict_of_agents_configs,
                               number_of_agents_per_type=mac.number_of_agents_per_type)
    return mac


def with_agents(mac: MultiAgentConfig, agents: dict, n_per_type) -> MultiAgentConfig:
    """The same world with another set of agent types / counts."""
    return MultiAgentConfig(world_config=mac.world_config, dict_of_agents_configs=agents,
                            number_of_agents_per_type=list(n_per_type))


def num_action_msgs(c: abi.LobStepConfig) -> int:
    return sum(c.agent[i].n_agents * c.agent[i].num_action_messages_by_agent for i in range(c.n_agent_types))


def num_cancel_msgs(c: abi.LobStepConfig) -> int:
    return sum(c.agent[i].n_agents * (c.agent[i].num_messages_by_agent - c.agent[i].num_action_messages_by_agent)
               for i in range(c.n_agent_types))


def num_msgs_per_step(c: abi.LobStepConfig) -> int:
    """marl_env.py:85-94"""
    return c.n_data_msg_per_step + num_action_msgs(c) + num_cancel_msgs(c)
