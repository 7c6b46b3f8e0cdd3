"""Seed handling that reproduces the reference's RNG streams bit-for-bit.

The reference seeds every stream as ``np.random.RandomState(u32 words of sha512(str(seed))[:8])``
(/root/reference/pgdrive/utils/random_utils.py:14-50,89-100) and samples block / vehicle parameters by
seeding EVERY Box of a parameter space with the same integer and drawing one legacy ``uniform`` per Box
(utils/space.py:109-113,423-457; base_class/base_runnable.py:81-88).  The host side of the reset path
(map search, episode templates) is Python, like the reference; only the sampling order is restated here.
"""
import hashlib
import struct

import numpy as np

MAX_RAND_INT = 65536  # base_class/randomizable.py:10


def hash_seed(seed: int) -> int:
    digest = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    lo, hi = struct.unpack("<2I", digest)
    return lo + (hi << 32)


def seeded(seed: int) -> np.random.RandomState:
    """``get_np_random(seed)`` of the reference: legacy MT19937 seeded with the hashed seed's u32 words."""
    if not (isinstance(seed, (int, np.integer)) and seed >= 0):
        raise ValueError("Seed must be a non-negative integer, not {}".format(seed))
    big = hash_seed(int(seed))
    words = []
    while big > 0:
        big, mod = divmod(big, 2**32)
        words.append(mod)
    rs = np.random.RandomState()
    rs.seed(words or [0])
    return rs


def draw_seed(rs: np.random.RandomState) -> int:
    """``Randomizable.generate_seed`` (randomizable.py:20-21)."""
    return int(rs.randint(0, MAX_RAND_INT))


def box_f32(rs: np.random.RandomState, low: float, high: float) -> float:
    """One float Box sample: uniform in double between the float32-rounded bounds, rounded to float32."""
    lo, hi = np.float32(low), np.float32(high)
    u = rs.uniform(low=np.array([lo]), high=np.array([hi]), size=(1, ))
    return float(np.float32(u[0]))


def box_int(rs: np.random.RandomState, low: int, high: int) -> int:
    """One integer Box sample: floor(uniform(low, high + 1)) (space.py:437-451)."""
    u = rs.uniform(low=np.array([np.int64(low)]), high=np.array([np.int64(high) + 1]), size=(1, ))
    return int(np.floor(u[0]))


def sample_space(space: dict, rs_parent: np.random.RandomState) -> dict:
    """``BaseRunnable.sample_parameters``: one ``randint(0, 1e6)`` from the owner's stream, then every
    entry of ``space`` draws from its OWN fresh stream seeded with that integer.

    ``space`` maps name -> ("f", low, high) | ("i", low, high) | ("c", value).  Note the reference's
    vehicle spaces are written ``BoxSpace(750, 850)`` against ``namedtuple("BoxSpace", "max min")``, i.e.
    low=850 / high=750 (space.py:14,222-223); callers pass the bounds in that literal order.
    """
    q = int(rs_parent.randint(low=0, high=int(1e6)))
    out = {}
    for name in sorted(space):
        spec = space[name]
        if spec[0] == "c":
            # ConstantSpace -> Box(low=v, high=v): still a bounded float Box, sampled and rounded to f32
            out[name] = box_f32(seeded(q), spec[1], spec[1])
        elif spec[0] == "f":
            out[name] = box_f32(seeded(q), spec[1], spec[2])
        elif spec[0] == "i":
            out[name] = box_int(seeded(q), spec[1], spec[2])
        else:
            raise ValueError(spec)
    return out
