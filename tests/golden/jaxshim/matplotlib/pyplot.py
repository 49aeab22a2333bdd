def __getattr__(name):
    raise AttributeError(f"matplotlib.pyplot.{name} is not available in the shim")
