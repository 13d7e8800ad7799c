"""avatarcraft_b200 -- B200-native (sm_100a) implementation of AvatarCraft's render hot path
behind the reference's own Python operator / model API.

    avatarcraft_b200.encoder      get_encoder, HashEncoder, SHEncoder     (reference: encoder/)
    avatarcraft_b200.models       instant_nsr.NeRFNetwork / NeRFRenderer  (reference: models/)
    avatarcraft_b200.utils        render_utils.render_instantnsr_naive    (reference: utils/)

All arithmetic on the hot path runs in libavatarcraft_b200.so (hand-written CUDA, C ABI in
include/avatarcraft_b200.h).  PyTorch is used for device memory, streams and autograd glue.
"""
__version__ = "0.1.0"
