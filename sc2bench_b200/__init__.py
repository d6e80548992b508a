"""Import alias: `import sc2bench_b200` resolves to the code in `sc2-benchmark_b200/` (a directory name Python
cannot import directly).  This package only redirects its search path; all code lives next door."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'sc2-benchmark_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f
