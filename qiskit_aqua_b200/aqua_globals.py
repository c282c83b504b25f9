"""Process-wide seed / RNG / `massive` switch, the three globals the hot path's callers consult
(reference qiskit/aqua/aqua_globals.py:36-58 random_seed, :89-96 random, :98-108 massive)."""
import numpy as np


class _AquaGlobals:
    def __init__(self):
        self._random_seed = None
        self._random = None
        self.massive = False

    @property
    def random_seed(self):
        return self._random_seed

    @random_seed.setter
    def random_seed(self, seed):
        self._random_seed = seed
        self._random = None  # the generator is rebuilt lazily from the new seed

    @property
    def random(self):
        if self._random is None:
            self._random = np.random.default_rng(self._random_seed)
        return self._random


aqua_globals = _AquaGlobals()
