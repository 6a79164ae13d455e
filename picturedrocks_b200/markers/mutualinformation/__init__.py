from . import infoset, iterative  # noqa: F401
