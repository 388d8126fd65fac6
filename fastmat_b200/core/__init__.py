from .types import *        # noqa: F401,F403
