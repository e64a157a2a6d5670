"""Abstract base class of contractions (tensorsat/contractions/abc.py:32-97)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any

from ..diagrams import Box, Diagram, Wiring


class Contraction(ABC):
    """A way of contracting every diagram that has a given wiring."""

    def __init__(self, box_class: type, wiring: Wiring) -> None:
        if not (isinstance(box_class, type) and issubclass(box_class, Box)):
            raise TypeError("box_class must be a subclass of Box.")
        if not isinstance(wiring, Wiring):
            raise TypeError("wiring must be a Wiring.")
        if not box_class.can_be_contracted():
            raise ValueError("Given box class cannot be contracted.")
        self._box_class = box_class
        self._wiring = wiring

    @property
    def box_class(self) -> type:
        return self._box_class

    @property
    def wiring(self) -> Wiring:
        return self._wiring

    def can_contract(self, diagram: Diagram) -> bool:
        try:
            self.validate(diagram)
            return True
        except ValueError:
            return False

    def _accepts(self, cls: type) -> bool:
        return issubclass(cls, self._box_class)

    def validate(self, diagram: Diagram) -> None:
        """Raises ``ValueError`` unless the diagram is flat, closed, has this wiring and
        boxes of the contraction's box class (abc.py:64-89)."""
        if not isinstance(diagram, Diagram):
            raise TypeError("Expected a Diagram.")
        if not all(self._accepts(cls) for cls in diagram.box_classes):
            raise ValueError(
                f"Cannot contract diagram: diagram box classes {set(diagram.box_classes)} are"
                f" not subclasses of contraction box class {self._box_class}"
            )
        if diagram.wiring != self.wiring:
            raise ValueError("Diagram's wiring must match contraction wiring.")
        if diagram.num_open_slots > 0:
            raise ValueError("Diagram cannot have open slots.")
        if not diagram.is_flat:
            raise ValueError("Diagram must be flat.")

    @abstractmethod
    def contract(self, diagram: Diagram, *args: Any, **kwargs: Any) -> Box:
        """Validates and contracts the given diagram."""
