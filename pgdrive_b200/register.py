"""gym registration of the eight PGDrive ids (/root/reference/pgdrive/register.py:7-43): same ids, same
``kwargs=dict(config=dict(start_seed=..., environment_num=...))``, entry point = this package's ``PGDriveEnv``.

The reference registers at ``import pgdrive``; here it happens at ``import pgdrive_b200`` when ``gym`` (the API the
reference was written against) or ``gymnasium`` is importable, and is skipped silently otherwise -- ``pgdrive_b200.make``
works either way."""
from .env import ENVIRONMENTS

registered_with = []


def get_env_list():
    return list(ENVIRONMENTS.keys())


def register_all():
    for name in ("gym", "gymnasium"):
        try:
            mod = __import__(name + ".envs.registration", fromlist=["register", "registry"])
        except Exception:  # not installed (or a broken install): nothing to register with
            continue
        if name in registered_with:
            continue
        registry = getattr(mod, "registry", None)
        for env_id, cfg in ENVIRONMENTS.items():
            try:
                known = env_id in registry if registry is not None and not hasattr(registry, "env_specs") else \
                    env_id in registry.env_specs
            except Exception:
                known = False
            if known:
                continue
            mod.register(id=env_id, entry_point="pgdrive_b200.env:PGDriveEnv", kwargs=dict(config=dict(cfg)))
        registered_with.append(name)
    return list(registered_with)
