from ._mesh import Mesh

__all__ = ["Mesh"]
