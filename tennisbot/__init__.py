"""Drop-in for the reference's `tennisbot` package: `import tennisbot; gym.make("SwingRacket-v0")` keeps working.

Same two ids and entry points as the reference's registration module, served by the CUDA envs of
tennisbot_rl_b200.  Registration is skipped when neither gym nor gymnasium is installed (use
`tennisbot_rl_b200.make(id)` then).
"""
ENTRY_POINTS = {
    "Tennisbot-v0": "tennisbot.envs:TennisbotEnv",
    "SwingRacket-v0": "tennisbot.envs:SwingRacketEnv",
}


def _register_all():
    for mod in ("gym.envs.registration", "gymnasium.envs.registration"):
        try:
            reg = __import__(mod, fromlist=["register"])
        except Exception:
            continue
        for env_id, entry in ENTRY_POINTS.items():
            try:
                reg.register(id=env_id, entry_point=entry)
            except Exception:  # already registered
                pass


_register_all()
