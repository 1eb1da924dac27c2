"""Minimal torch_geometric stand-in (only what the reference's hot-path callers touch)."""
from . import data, nn, utils  # noqa: F401
