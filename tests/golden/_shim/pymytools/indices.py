"""Stand-in for the un-vendored third-party `pymytools.indices` (pinned 0.1.17 in the
reference's poetry.lock:1458).  Only `tensor_idx` is imported by the reference
(`pyapes/solver/fdc.py:12`), and only `hessian` uses it.  Used solely by
tests/golden/make_golden.py when importing the real reference in the build container."""


def tensor_idx(dim: int) -> list[tuple[int, int]]:
    return [(i, j) for i in range(dim) for j in range(i, dim)]
