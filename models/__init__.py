"""Reference package name `models` (models/instant_nsr.py, models/neus.py, models/diffusion.py, models/smpl.py): aliases of avatarcraft_b200.models.*."""
