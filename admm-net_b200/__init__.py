"""admmnet_b200 — B200 (sm_100a) implementation of the ADMM-Net forward / classical ADMM / peak-search
hot path of E-J408/admm-net behind the reference's Python API.  See DESIGN.md."""
from . import _capi  # noqa: F401
from .admm import admm_for_us, admm_for_us_batched  # noqa: F401
from .admm_net import ADMMNet, GLayer, HLayer, PeakSearchLayer, PhiEstADMMNet, PhiLayer, ZLayer  # noqa: F401
from .peaksearch import alt_peak_search, alt_peak_search_batched, peak_search, peak_search_func  # noqa: F401
from .generate import generate_signals  # noqa: F401
from .autograd import BasicANMLoss, PhiAlignmentLoss  # noqa: F401
