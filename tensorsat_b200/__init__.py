"""
tensorsat_b200 — B200-native contraction executor behind TensorSat's box / autoray API.

Scope: the contraction hot path only: pairwise boolean / counting
contractions, rewires, slicing hooks and whole-tree execution on sm_100a, exposed as a
C-ABI shared library (include/tsb.h) with this package as its host-side mirror of the
reference's plug-in interface.
"""

__version__ = "0.1.0"
