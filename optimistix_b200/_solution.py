"""RESULTS and Solution - mirrors optimistix/_solution.py:11-43 of the reference."""
from __future__ import annotations

import dataclasses
from typing import Any

_MESSAGES = {
    # lineax members (lx.RESULTS), then optimistix/_solution.py:12-20
    "successful": "",
    "max_steps_reached": "The maximum number of solver steps was reached. Try increasing `max_steps`.",
    "singular": "The linear solver returned non-finite (NaN or inf) output. This usually means that an "
    "operator was not well-posed, and that its solver does not support this.",
    "breakdown": "A form of iterative breakdown has occured in a linear solve.",
    "stagnation": "A stagnation in an iterative linear solve has occurred.",
    "nonlinear_max_steps_reached": "The maximum number of steps was reached in the nonlinear solver. "
    "The problem may not be solveable (e.g., a root-find on a function that has no "
    "roots), or you may need to increase `max_steps`.",
    "nonlinear_divergence": "Nonlinear solve diverged.",
    "nonfinite": "Nonfinite (inf or nan) values detected during solve.",
}


class _ResultsMeta(type):
    def __getitem__(cls, code):
        """`RESULTS[result]` -> human-readable message (reference: `optimistix.RESULTS[result]`)."""
        return _MESSAGES[cls.names[int(code)]]

    def __iter__(cls):
        return iter(range(len(cls.names)))


class RESULTS(metaclass=_ResultsMeta):
    """Integer result codes; member names are those of `optimistix.RESULTS`. Compare by name across
    implementations (the integer values of the reference depend on the installed lineax)."""

    names = tuple(_MESSAGES)
    successful = 0
    max_steps_reached = 1
    singular = 2
    breakdown = 3
    stagnation = 4
    nonlinear_max_steps_reached = 5
    nonlinear_divergence = 6
    nonfinite = 7

    @classmethod
    def name(cls, code) -> str:
        return cls.names[int(code)]


@dataclasses.dataclass
class Solution:
    """optimistix/_solution.py:22-43.  For a batched solve every field carries a leading batch axis."""

    value: Any
    result: Any
    aux: Any
    stats: dict
    state: Any
