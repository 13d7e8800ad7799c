"""Reference package name `encoder` (encoder/__init__.py:4 get_encoder): the factory of avatarcraft_b200.encoder."""
from avatarcraft_b200.encoder import get_encoder  # noqa: F401
