"""Drop-in for the reference's top-level `admm_net` module (same import path): re-exports the B200 mirror."""
from admmnet_b200.admm_net import (ADMMNet, GLayer, HLayer, PeakSearchLayer, PhiEstADMMNet, PhiLayer,  # noqa: F401
                                    ZLayer)
