"""jax.random stub: only what is evaluated at import time of the reference
(default-argument `PRNGKey(42)`).  The hot path draws no random numbers.
TEST INFRASTRUCTURE ONLY."""


class PRNGKey:
    def __init__(self, seed=0):
        self.seed = seed


def split(key, num=2):
    return [PRNGKey((key.seed, i)) for i in range(num)]


def permutation(*_a, **_k):
    raise NotImplementedError("jax.random is not emulated; the hot path does not use it")


choice = permutation
