"""pytest plugin (`-p pyapes_b200.pytest_alias`): run a test suite written for the reference package
`pyapes` against this package without touching the test files -- every `import pyapes...` resolves to
`pyapes_b200...` (install_as_pyapes).  The terminal summary names the module the tests really used."""
import sys

import pyapes_b200

pyapes_b200.install_as_pyapes()


def pytest_terminal_summary(terminalreporter):
    mod = sys.modules.get("pyapes.solver.ops")
    terminalreporter.write_line(f"pyapes_alias: pyapes.solver.ops -> {getattr(mod, '__name__', None)}")
