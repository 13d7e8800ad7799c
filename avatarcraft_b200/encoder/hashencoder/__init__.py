from .hashgrid import HashEncoder
