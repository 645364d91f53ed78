"""PyG's signature inspector is pickled with every MessagePassing module; the shim's propagate() reads the
message signature itself, so this only has to exist and take its pickled state."""


class Inspector(object):
    def __init__(self, base_class=None):
        self.base_class = base_class
        self.params = {}
