from .grid import GridSpec, HashEncoder, hash_encode  # noqa: F401
