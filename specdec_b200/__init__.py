"""Import alias: the product package lives in the directory `speculative-decoding_b200/`
(the name the project layout prescribes, which is not a valid Python identifier).
`import specdec_b200` resolves every submodule from that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "speculative-decoding_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
