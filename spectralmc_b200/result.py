"""Result ADT used on every expected-failure path of the reference API.

Mirrors the interface of /root/reference/src/spectralmc/result.py:39-231 (``Success`` /
``Failure`` with ``unwrap``/``map``/``and_then``, ``collect_results``, ``fold_results``,
``expect``) so callers and tests written against the reference read the same here.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Generic, Iterable, NoReturn, TypeVar, Union

T = TypeVar("T")
E = TypeVar("E")
U = TypeVar("U")
F = TypeVar("F")


@dataclass(frozen=True)
class Success(Generic[T]):
    value: T

    def is_success(self) -> bool:
        return True

    def is_failure(self) -> bool:
        return False

    def unwrap(self) -> T:
        return self.value

    def unwrap_or(self, default: T) -> T:
        return self.value

    def unwrap_or_else(self, f: Callable[[object], T]) -> T:
        return self.value

    def map(self, f: Callable[[T], U]) -> "Success[U]":
        return Success(f(self.value))

    def map_error(self, f: Callable[[object], object]) -> "Success[T]":
        return self

    def flat_map(self, f: Callable[[T], "Result[U, E]"]) -> "Result[U, E]":
        return f(self.value)

    and_then = flat_map


@dataclass(frozen=True)
class Failure(Generic[E]):
    error: E

    def is_success(self) -> bool:
        return False

    def is_failure(self) -> bool:
        return True

    def unwrap(self) -> NoReturn:
        raise RuntimeError(f"Called unwrap() on Failure: {self.error}")

    def unwrap_or(self, default: T) -> T:
        return default

    def unwrap_or_else(self, f: Callable[[E], T]) -> T:
        return f(self.error)

    def map(self, f: Callable[[object], object]) -> "Failure[E]":
        return self

    def map_error(self, f: Callable[[E], F]) -> "Failure[F]":
        return Failure(f(self.error))

    def flat_map(self, f: Callable[[object], object]) -> "Failure[E]":
        return self

    and_then = flat_map


Result = Union[Success[T], Failure[E]]


def expect(result: "Result[T, E]") -> T:
    """Unwrap or raise AssertionError (test ergonomics; result.py:133-144 of the reference)."""
    if isinstance(result, Success):
        return result.value
    raise AssertionError(f"Unexpected failure: {result.error}")


def collect_results(results: Iterable["Result[T, E]"]) -> "Result[list[T], E]":
    """All successes -> Success(list); otherwise the first Failure."""
    values: list[T] = []
    for r in results:
        if isinstance(r, Failure):
            return r
        values.append(r.value)
    return Success(values)


def partition_results(results: Iterable["Result[T, E]"]) -> tuple[list[T], list[E]]:
    ok: list[T] = []
    bad: list[E] = []
    for r in results:
        (ok if isinstance(r, Success) else bad).append(r.value if isinstance(r, Success) else r.error)
    return ok, bad


def fold_results(items: Iterable[T], f: Callable[[U, T], "Result[U, E]"], initial: U) -> "Result[U, E]":
    """Left fold that stops at the first Failure."""
    acc = initial
    for item in items:
        step = f(acc, item)
        if isinstance(step, Failure):
            return step
        acc = step.value
    return Success(acc)
