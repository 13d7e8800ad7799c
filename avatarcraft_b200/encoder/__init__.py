"""Encoder factory with the reference's signature (encoder/__init__.py:4-32)."""
from . import freq_encoder
from .hashencoder import HashEncoder


def get_encoder(encoder_type: str, encoder_configs: dict):
    """Construct and return (encoder module, output dimension)."""
    if encoder_type == "frequency":
        return freq_encoder.get_freq_embedder(encoder_configs["freq_multires"], encoder_configs["in_dim"])
    if encoder_type in ("hash", "hashgrid"):
        c = encoder_configs
        enc = HashEncoder(c["in_dim"], c["hash_num_levels"], c["hash_level_dim"], c["hash_per_level_scale"],
                          c["hash_base_resolution"], c["hash_log2_hashmap_size"], c["hash_desired_resolution"])
        return enc, enc.output_dim
    if encoder_type in ("sh", "sphere_harmonics"):
        from .shencoder import SHEncoder
        enc = SHEncoder(encoder_configs["in_dim"])
        return enc, enc.output_dim
    raise NotImplementedError("Encoder type {} not implemented".format(encoder_type))
