from . import environment, spaces  # noqa: F401
