"""lobstep-b200: B200-native batched limit-order-book step (one hot path of JaxMARL-HFT)."""
from . import abi, config, jorderbook, lobster  # noqa: F401
from .jorderbook import LobState, OrderBook  # noqa: F401
from .config import (  # noqa: F401
    JAXLOB_Configuration, World_EnvironmentConfig, MarketMaking_EnvironmentConfig, Execution_EnvironmentConfig,
    MultiAgentConfig, CONFIG_OBJECT_DICT, load_config_from_file)

__version__ = "0.1.0"
