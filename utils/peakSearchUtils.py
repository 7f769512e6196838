"""Drop-in for the reference's `utils.peakSearchUtils` hot-path functions: re-exports the B200 mirror."""
from admmnet_b200.peaksearch import alt_peak_search, peak_search, peak_search_func  # noqa: F401
