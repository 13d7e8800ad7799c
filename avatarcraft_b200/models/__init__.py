from . import instant_nsr  # noqa: F401
