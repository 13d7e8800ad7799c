from .sh import SHEncoder, sh_encode  # noqa: F401
