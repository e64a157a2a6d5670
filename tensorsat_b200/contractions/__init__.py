"""Contraction strategies (mirror of ``tensorsat.contractions``)."""

from .abc import Contraction
from .simple import SimpleContraction
from .tree import TreeContraction

__all__ = ("Contraction", "SimpleContraction", "TreeContraction")
