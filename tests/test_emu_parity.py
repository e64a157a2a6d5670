"""CPU suite: the parity cases against the C-ABI code compiled for the host emulation."""

import pytest

from tests.backend import bind_emulation
from tests.parity_cases import *  # noqa: F401,F403


@pytest.fixture(autouse=True, scope="module")
def _bind():
    bind_emulation()
