class Discrete:
    def __init__(self, num_categories):
        self.n = num_categories
        self.shape = ()


class Box:
    def __init__(self, low, high, shape, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class Dict:
    def __init__(self, spaces):
        self.spaces = spaces


class Tuple:
    def __init__(self, spaces):
        self.spaces = spaces
