"""pyapes_b200 — B200-native finite-difference hot path behind the pyapes API.

Module paths mirror the reference (`pyapes.geometry`, `.mesh`, `.variables`, `.solver.fdm`,
`.solver.fdc`, `.solver.ops`, `.solver.linalg`, `.testing.poisson`).  `install_as_pyapes()`
registers the package under the name `pyapes` so that existing scripts run unchanged.
"""
__version__ = "0.1.0"


def install_as_pyapes() -> None:
    """Alias this package as `pyapes` in sys.modules (drop-in for `import pyapes...`)."""
    import importlib
    import sys

    names = ["", ".backend", ".geometry", ".geometry.basis", ".geometry.box", ".geometry.cylinder",
             ".mesh", ".mesh._mesh", ".mesh.tools", ".variables", ".variables.bcs", ".variables.fields", ".variables.container",
             ".solver", ".solver.tools", ".solver.types", ".solver.fdc", ".solver.fdm", ".solver.linalg",
             ".solver.ops", ".testing", ".testing.poisson", ".testing.burgers", ".runner"]
    for n in names:
        sys.modules["pyapes" + n] = importlib.import_module("pyapes_b200" + n)
