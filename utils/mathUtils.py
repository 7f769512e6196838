"""Drop-in for the reference's `utils.mathUtils`: re-exports the host helpers."""
from admmnet_b200.mathutils import awgn, kr, pskdemod, pskmod, vander_vec  # noqa: F401
