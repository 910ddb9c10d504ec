"""CameraSensor is part of the vision stages (rendering path) and out of scope for the
post-physics hot path.  The symbol exists so reference-style imports
resolve; using it raises."""
from .base import Sensor


class CameraSensor(Sensor):
    def __init__(self, cfg):
        raise NotImplementedError("CameraSensor (rendering path) is outside the shifu_b200 hot path")
