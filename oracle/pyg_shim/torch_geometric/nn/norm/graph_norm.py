"""Import path PyG pickles GraphNorm under (checkpoints/cns.pth)."""
from . import GraphNorm  # noqa: F401
