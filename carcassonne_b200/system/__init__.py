"""Lattice systems of the center-site path (reference carcassonne/system/)."""
from ._2d import System, sideFromCorner  # noqa: F401
from .base import BaseSystem  # noqa: F401
