"""CPU oracle for the SpectralMC batch-generation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``spectralmc_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the CPU baseline — never as the thing shipped.

What it restates (reference = /root/reference/src/spectralmc, cited per function):

* ``oracle.gbm``      – the Numba path kernel (gbm.py:241-257), the engine numerics
                        (gbm.py:411,428-440), the payoff (gbm.py:464-474), the host
                        reduction (gbm.py:491-513) and the CF estimate
                        (gbm_trainer.py:814-817).
* ``oracle.philox``   – the NEW counter-based normal stream (Philox4x32-10 +
                        Box–Muller) that replaces the CuPy XORWOW generator
                        (async_normals.py:214-215).  The reference's RNG bits are
                        a third-party detail that no reference test pins, so this
                        is a specification of the new stream, pinned by the
                        Random123 known-answer vectors.
* ``oracle.black76``  – closed-form Black-76, standing in for QuantLib
                        ``ql.blackFormula`` (quantlib.py:21-29; QuantLib is not
                        installed in this image).
* ``oracle.sobol``    – the Sobol contract batch (sobol_sampler.py:187-203,238-239).
* ``oracle/gbm_oracle.c`` – plain-C restatement of the same path (OpenMP), used as
                        the multi-core CPU baseline and as a second, independent
                        checker of the Python restatement.

Parity pinning: the reference holds no golden vectors for this path (SURVEY.md
§8c).  The restatement is therefore pinned against OUTPUTS OF THE REFERENCE
ITSELF, produced in the build container by ``tests/golden/make_golden.py`` (the
reference's own ``BlackScholes._simulate/price`` and kernel body executed under
``NUMBA_ENABLE_CUDASIM=1`` with a NumPy stand-in for CuPy) and committed as
``tests/golden/*.npz``.  The RNG stream and the FFT library call are "parity
unpinned" in the reference (third-party CuPy/cuFFT, never value-tested there).
"""
