from . import random  # noqa: F401
