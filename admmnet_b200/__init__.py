"""Import alias: the package lives in the directory `admm-net_b200/` (not a valid Python identifier);
`import admmnet_b200` resolves to it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "admm-net_b200")]
exec(open(_os.path.join(__path__[0], "__init__.py")).read())
