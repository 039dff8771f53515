from .sparse_lion import SparseLion  # noqa: F401
