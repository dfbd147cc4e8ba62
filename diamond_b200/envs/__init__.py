from .world_model_env import WorldModelEnv, WorldModelEnvConfig
