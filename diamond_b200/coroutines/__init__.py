from .env_loop import ImaginationLoop, make_env_loop
