"""gymnax.environments.environment stand-in: the base class the reference's BaseLOBEnv inherits from."""
from flax import struct


@struct.dataclass
class EnvState:
    time: int = 0


@struct.dataclass
class EnvParams:
    max_steps_in_episode: int = 1


class Environment:
    def __init__(self):
        pass

    @property
    def default_params(self):
        return EnvParams()
