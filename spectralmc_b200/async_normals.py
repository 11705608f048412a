"""Normal-draw supply — drop-in for the reference's ``spectralmc.async_normals``.

Same public surface as /root/reference/src/spectralmc/async_normals.py
(``BufferConfig.create`` :112, ``ConcurrentNormGeneratorConfig.create`` :156,
``_NormGenerator.create/enqueue/get_matrix/is_ready`` :187-249,
``ConcurrentNormGenerator.create/get_matrix/snapshot/get_time_spent_synchronizing/get_idle_time``
:299-459) with the CuPy XORWOW generator replaced by the counter-based Philox4x32-10 kernel
``smc_philox_normals`` (csrc/smc_normals.cu):

* matrix ``k`` (``k`` = matrices ever served, the reference's ``skips``) is a pure function of
  ``(seed, k)`` — Philox is keyed by ``seed`` and ``k`` is a counter word, so restoring from a
  snapshot needs no fast-forward of a NumPy seed stream (reference :319-326) and the buffer size
  cannot change the stream (tests/test_async_normals.py:94-123 of the reference);
* matrices are ``torch`` CUDA tensors (DLPack-native); ``dtype`` properties return ``torch.dtype``;
* hand-off is stream-ordered (the consumer's stream waits on the worker's event) instead of a
  host ``stream.synchronize()`` (reference :230); the two host-side diagnostics are kept.
"""

from __future__ import annotations

from dataclasses import dataclass
from time import time

import torch

from spectralmc_b200 import _cabi
from spectralmc_b200.errors import InvalidDType, InvalidShape, QueueBusy, QueueEmpty, SeedOutOfRange
from spectralmc_b200.numerical import Precision
from spectralmc_b200.result import Failure, Result, Success, collect_results

__all__ = ["BufferConfig", "ConcurrentNormGeneratorConfig", "ConcurrentNormGenerator"]

# Philox takes any 64-bit key; the reference's limit of 1e9 (async_normals.py:84) was a CuPy
# seed restriction.  Seeds must still be positive, as in the reference (:161,:204).
_SEED_LIMIT: int = 1 << 63


def _validate_dtype(dtype: torch.dtype) -> Result[torch.dtype, InvalidDType]:
    return Success(dtype) if dtype in (torch.float32, torch.float64) else Failure(InvalidDType(requested=str(dtype)))


@dataclass(frozen=True)
class BufferConfig:
    """Number of matrices kept in flight (reference :105-125)."""

    size: int

    @classmethod
    def create(cls, size: int, matrix_rows: int, matrix_cols: int) -> Result["BufferConfig", InvalidShape]:
        if size > matrix_rows * matrix_cols or min(matrix_rows, matrix_cols) <= 0 or size <= 0:
            return Failure(InvalidShape(rows=matrix_rows, cols=matrix_cols))
        return Success(cls(size=size))


@dataclass(frozen=True)
class ConcurrentNormGeneratorConfig:
    """Serialisable generator state: ``skips`` matrices have been served (reference :128-165)."""

    rows: int
    cols: int
    seed: int
    dtype: Precision
    skips: int = 0
    stream_version: int = 0  # 0: Philox4x32-10 (default); 1: the opt-in Philox4x32-7 stream (include/spectralmc_b200.h)

    @classmethod
    def create(
        cls, *, rows: int, cols: int, seed: int, dtype: Precision, skips: int = 0, stream_version: int = 0
    ) -> Result["ConcurrentNormGeneratorConfig", InvalidShape | SeedOutOfRange]:
        if rows <= 0 or cols <= 0:
            return Failure(InvalidShape(rows=rows, cols=cols))
        if seed <= 0 or seed >= _SEED_LIMIT:
            return Failure(SeedOutOfRange(seed=seed))
        if skips < 0:
            return Failure(SeedOutOfRange(seed=skips))
        if stream_version not in (0, 1):
            raise ValueError(f"stream_version must be 0 (Philox4x32-10) or 1 (Philox4x32-7); got {stream_version}")
        return Success(cls(rows=rows, cols=cols, seed=seed, dtype=dtype, skips=skips, stream_version=stream_version))


class _NormGenerator:
    """One worker: fills matrices on its own CUDA stream (reference :173-256)."""

    def __init__(self, rows: int, cols: int, *, dtype: torch.dtype, stream_version: int = 0) -> None:
        self._rows, self._cols, self._dtype, self._stream_version = rows, cols, dtype, stream_version
        self._stream = torch.cuda.Stream()
        self._generated: torch.Tensor | None = None
        self._event: torch.cuda.Event | None = None
        self._sync_time = 0.0
        self.matrix_index: int | None = None  # which matrix of the stream is in flight

    @classmethod
    def create(cls, rows: int, cols: int, *, dtype: torch.dtype, stream_version: int = 0) -> Result["_NormGenerator", InvalidShape | InvalidDType]:
        if min(rows, cols) <= 0:
            return Failure(InvalidShape(rows=rows, cols=cols))
        checked = _validate_dtype(dtype)
        if isinstance(checked, Failure):
            return checked
        return Success(cls(rows, cols, dtype=dtype, stream_version=stream_version))

    def enqueue(self, seed: int, matrix_index: int = 0) -> Result[None, QueueBusy | SeedOutOfRange]:
        """Launch the fill kernel for matrix ``matrix_index`` of stream ``seed`` (non-blocking)."""
        if self._generated is not None:
            return Failure(QueueBusy())
        if seed <= 0 or seed >= _SEED_LIMIT:
            return Failure(SeedOutOfRange(seed=seed))
        self._event = torch.cuda.Event()
        with torch.cuda.stream(self._stream):
            out = torch.empty((self._rows, self._cols), dtype=self._dtype, device="cuda")
            _cabi.philox_normals(out, seed, matrix_index, self._stream_version)
            self._event.record()
        self._generated = out
        self.matrix_index = matrix_index
        return Success(None)

    def discard(self) -> None:
        """Drop the matrix in flight (used when the consumer skipped past its index)."""
        self._generated = None
        self.matrix_index = None

    def get_matrix(
        self, next_seed: int, next_matrix_index: int = 0
    ) -> Result[torch.Tensor, QueueEmpty | QueueBusy | SeedOutOfRange]:
        """Order the caller's stream after the fill, hand the matrix out, queue the next one."""
        if self._generated is None or self._event is None:
            return Failure(QueueEmpty())
        t0 = time()
        consumer = torch.cuda.current_stream()
        consumer.wait_event(self._event)
        ready = self._generated
        ready.record_stream(consumer)
        self._sync_time += time() - t0
        self._generated = None
        queued = self.enqueue(next_seed, next_matrix_index)
        return Success(ready) if isinstance(queued, Success) else queued

    def get_time_spent_synchronizing(self) -> float:
        return self._sync_time

    def is_ready(self) -> bool:
        return self._event is not None and self._event.query()

    @property
    def dtype(self) -> torch.dtype:
        return self._dtype


class ConcurrentNormGenerator:
    """Round-robin pool of workers with a deterministic ``(seed, skips)`` checkpoint."""

    def __init__(self, *, pool: list[_NormGenerator], rows: int, cols: int, dtype: torch.dtype, base_seed: int, served: int,
                 stream_version: int = 0) -> None:
        self._rows, self._cols, self._dtype, self._stream_version = rows, cols, dtype, stream_version
        self._base_seed = base_seed
        self._served = served
        self._pool = pool
        self._idle_accum = 0.0
        self._idle_start: float | None = None
        self._update_idle_state()

    @classmethod
    def create(
        cls, buffer_result: Result[BufferConfig, InvalidShape], config: ConcurrentNormGeneratorConfig
    ) -> Result["ConcurrentNormGenerator", InvalidShape | InvalidDType | QueueBusy | SeedOutOfRange]:
        if isinstance(buffer_result, Failure):
            return buffer_result
        buffer = buffer_result.value
        dtype = config.dtype.to_torch()

        def _make(slot: int) -> Result[_NormGenerator, InvalidShape | InvalidDType | QueueBusy | SeedOutOfRange]:
            made = _NormGenerator.create(config.rows, config.cols, dtype=dtype, stream_version=config.stream_version)
            if isinstance(made, Failure):
                return made
            queued = made.value.enqueue(config.seed, config.skips + slot)
            return made if isinstance(queued, Success) else queued

        made = collect_results([_make(i) for i in range(buffer.size)])
        if isinstance(made, Failure):
            return made
        # slot (k mod size) always holds matrix k
        pool = sorted(made.value, key=lambda g: (g.matrix_index or 0) % buffer.size)
        return Success(cls(pool=pool, rows=config.rows, cols=config.cols, dtype=dtype, base_seed=config.seed, served=config.skips,
                           stream_version=config.stream_version))

    def _update_idle_state(self) -> None:
        """Accumulate wall time during which every worker's matrix is ready (reference :361-382)."""
        all_ready = all(g.is_ready() for g in self._pool)
        now = time()
        if all_ready and self._idle_start is None:
            self._idle_start = now
        elif not all_ready and self._idle_start is not None:
            self._idle_accum += now - self._idle_start
            self._idle_start = None

    def get_matrix(self) -> Result[torch.Tensor, QueueEmpty | QueueBusy | SeedOutOfRange]:
        """Matrix number ``served`` of the stream; its slot is refilled with ``served + pool``."""
        k = self._served
        gen = self._pool[k % len(self._pool)]
        if gen.matrix_index != k:  # the consumer skipped ahead (fused path): refill this slot
            gen.discard()
            queued = gen.enqueue(self._base_seed, k)
            if isinstance(queued, Failure):
                return queued
        got = gen.get_matrix(self._base_seed, k + len(self._pool))
        if isinstance(got, Success):
            self._served += 1
            self._update_idle_state()
        return got

    def skip(self, count: int) -> None:
        """Mark ``count`` matrices as consumed without materialising them (fused path)."""
        self._served += count

    def snapshot(self) -> ConcurrentNormGeneratorConfig:
        made = ConcurrentNormGeneratorConfig.create(
            rows=self._rows, cols=self._cols, seed=self._base_seed, dtype=Precision.from_torch(self._dtype), skips=self._served,
            stream_version=self._stream_version,
        )
        if isinstance(made, Failure):
            raise AssertionError(f"Invalid ConcurrentNormGeneratorConfig snapshot: {made.error}")
        return made.value

    def get_time_spent_synchronizing(self) -> float:
        return sum(g.get_time_spent_synchronizing() for g in self._pool)

    def get_idle_time(self) -> float:
        self._update_idle_state()
        return self._idle_accum + (time() - self._idle_start) if self._idle_start is not None else self._idle_accum

    @property
    def dtype(self) -> torch.dtype:
        return self._dtype
