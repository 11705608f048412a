"""Import-level drop-in: make ``from spectralmc.gbm import BlackScholes`` (and the other modules of the hot
path) resolve to this package, so code written against the reference runs unchanged.

    import spectralmc_b200.compat as compat
    compat.install_as_spectralmc()
    from spectralmc.gbm import BlackScholes, SimulateBlackScholes      # -> spectralmc_b200.gbm
    from spectralmc.effects.montecarlo import PathScheme               # -> spectralmc_b200.effects

Only the modules of the path are aliased (reference file -> module here); the rest of the reference (storage,
serialization, the effect system as a system, ...) is out of scope and stays unresolved, loudly.  A real
``spectralmc`` installation is never shadowed unless ``force=True``.
"""

from __future__ import annotations

import importlib
import importlib.util
import sys
import types

# reference module (under /root/reference/src/) -> module of this package
ALIASES = {
    "spectralmc.gbm": "spectralmc_b200.gbm",
    "spectralmc.async_normals": "spectralmc_b200.async_normals",
    "spectralmc.sobol_sampler": "spectralmc_b200.sobol_sampler",
    "spectralmc.gbm_trainer": "spectralmc_b200.gbm_trainer",
    "spectralmc.cvnn": "spectralmc_b200.cvnn",
    "spectralmc.cvnn_factory": "spectralmc_b200.cvnn_factory",
    "spectralmc.result": "spectralmc_b200.result",
    "spectralmc.validation": "spectralmc_b200.validation",
    "spectralmc.errors": "spectralmc_b200.errors",
    "spectralmc.errors.gbm": "spectralmc_b200.errors",
    "spectralmc.errors.async_normals": "spectralmc_b200.errors",
    "spectralmc.models.numerical": "spectralmc_b200.numerical",
    "spectralmc.effects.montecarlo": "spectralmc_b200.effects",
    "spectralmc.effects.interpreter": "spectralmc_b200.interpreter",
    "spectralmc.quantlib": "spectralmc_b200.analytic",
}
_PACKAGES = ("spectralmc", "spectralmc.models", "spectralmc.effects")


def install_as_spectralmc(*, force: bool = False) -> list[str]:
    """Register the aliases in ``sys.modules``; returns the names installed.  Raises if a real ``spectralmc`` is
    importable (or already imported) and ``force`` is not set."""
    existing = sys.modules.get("spectralmc")
    if not force and (getattr(existing, "__spectralmc_b200_alias__", False) is False) and (
        existing is not None or importlib.util.find_spec("spectralmc") is not None
    ):
        raise ImportError("a real `spectralmc` package is importable; pass force=True to shadow it with spectralmc_b200")
    for name in _PACKAGES:  # namespace shells for the packages the aliases hang from
        shell = types.ModuleType(name)
        shell.__path__ = []  # type: ignore[attr-defined]  (a package with nothing else inside)
        shell.__spectralmc_b200_alias__ = True  # type: ignore[attr-defined]
        sys.modules[name] = shell
    installed = []
    for ref_name, here in ALIASES.items():
        module = importlib.import_module(here)
        sys.modules[ref_name] = module
        parent, _, leaf = ref_name.rpartition(".")
        setattr(sys.modules[parent], leaf, module)
        installed.append(ref_name)
    return installed


def uninstall() -> None:
    for name in list(sys.modules):
        if name == "spectralmc" or name.startswith("spectralmc."):
            if name in ALIASES or getattr(sys.modules[name], "__spectralmc_b200_alias__", False):
                del sys.modules[name]
