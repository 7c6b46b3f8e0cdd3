"""B200-native batched driving simulator behind the PGDriveEnv gym surface."""
from .env import ENVIRONMENTS, PGDriveEnv, VecPGDriveEnv, make  # noqa: F401

__all__ = ["PGDriveEnv", "VecPGDriveEnv", "make", "ENVIRONMENTS"]
