"""Base class of contraction strategies (counterpart of ``tensorsat/contractions/abc.py:32-97``)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any

from ..diagrams import Box, Diagram, Wiring


class Contraction(ABC):
    """Binds a box class to a wiring; contracts any diagram that has exactly that wiring."""

    def __init__(self, box_class: type, wiring: Wiring) -> None:
        if not isinstance(box_class, type) or not issubclass(box_class, Box):
            raise TypeError("box_class must be a subclass of Box.")
        if not isinstance(wiring, Wiring):
            raise TypeError("wiring must be a Wiring.")
        if not box_class.can_be_contracted():
            raise ValueError("Given box class cannot be contracted.")
        self._box_class, self._wiring = box_class, wiring

    box_class = property(lambda self: self._box_class, doc="Box class the results belong to.")
    wiring = property(lambda self: self._wiring, doc="The wiring this contraction was planned for.")

    def _accepts(self, cls: type) -> bool:
        return issubclass(cls, self._box_class)

    def validate(self, diagram: Diagram) -> None:
        """``ValueError`` unless the diagram is flat, has no open slots, has this contraction's wiring and
        only boxes of (a subclass of) its box class — the four conditions of ``abc.py:64-89``."""
        if not isinstance(diagram, Diagram):
            raise TypeError("Expected a Diagram.")
        foreign = [cls for cls in diagram.box_classes if not self._accepts(cls)]
        if foreign:
            raise ValueError(
                f"Cannot contract diagram: diagram box classes {foreign} are"
                f" not subclasses of contraction box class {self._box_class}"
            )
        if diagram.wiring != self._wiring:
            raise ValueError("Diagram's wiring must match contraction wiring.")
        if diagram.num_open_slots:
            raise ValueError("Diagram cannot have open slots.")
        if not diagram.is_flat:
            raise ValueError("Diagram must be flat.")

    def can_contract(self, diagram: Diagram) -> bool:
        try:
            self.validate(diagram)
        except ValueError:
            return False
        return True

    @abstractmethod
    def contract(self, diagram: Diagram, *args: Any, **kwargs: Any) -> Box:
        """Validates and contracts the given diagram."""
