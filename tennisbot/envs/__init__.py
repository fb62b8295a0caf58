from tennisbot_rl_b200.envs import SwingRacketEnv, TennisbotEnv  # noqa: F401
