from .engine import TrainEngine  # noqa: F401
