"""Reference import path `encoder.shencoder` -> `avatarcraft_b200.encoder.shencoder` (drop-in: the reference's entry points import
this name; the implementation lives in the avatarcraft_b200 package).  The module object itself is aliased, so every public
name -- and isinstance / pickling by module path -- behaves as if the implementation had been imported directly."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("avatarcraft_b200.encoder.shencoder")
