"""Pydantic construction surfaced as a Result (reference: validation.py:17-29)."""

from __future__ import annotations

from typing import TypeVar

from pydantic import BaseModel, ValidationError

from spectralmc_b200.result import Failure, Result, Success

TModel = TypeVar("TModel", bound=BaseModel)


def validate_model(model_cls: type[TModel], **data: object) -> Result[TModel, ValidationError]:
    try:
        return Success(model_cls(**data))
    except ValidationError as exc:
        return Failure(exc)
