"""Reference package name `utils` (utils/render_utils.py, utils/ray_utils.py, utils/constant.py): aliases of avatarcraft_b200.utils.*."""
