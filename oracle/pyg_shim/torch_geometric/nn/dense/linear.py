from .. import Linear  # noqa: F401
