"""B200-native batched driving simulator behind the PGDriveEnv gym surface."""
from .env import ENVIRONMENTS, PGDriveEnv, SafePGDriveEnv, VecPGDriveEnv, make  # noqa: F401

from .register import get_env_list, register_all  # noqa: F401

register_all()  # gym / gymnasium ids, when one of them is importable (register.py)

__all__ = ["PGDriveEnv", "SafePGDriveEnv", "VecPGDriveEnv", "make", "ENVIRONMENTS", "get_env_list", "register_all"]
