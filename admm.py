"""Drop-in for the reference's top-level `admm` module (same import path): re-exports the B200 mirror."""
from admmnet_b200.admm import admm_for_us, admm_for_us_batched  # noqa: F401
