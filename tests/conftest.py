import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, 'oracle')
for p in (ROOT, ORACLE):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def oracle_compressai():
    """The CPU restatement of CompressAI (oracle/shim), imported under a private name so that it never
    shadows a real `compressai` and is never confused with the product."""
    import importlib.util
    shim = os.path.join(ORACLE, 'shim', 'compressai', '__init__.py')
    if 'compressai' in sys.modules and getattr(sys.modules['compressai'], '__file__', '') == shim:
        return sys.modules['compressai']
    spec = importlib.util.spec_from_file_location('compressai', shim, submodule_search_locations=[os.path.dirname(shim)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules['compressai'] = mod
    spec.loader.exec_module(mod)
    import compressai.entropy_models  # noqa: F401
    import compressai.layers  # noqa: F401
    import compressai.models  # noqa: F401
    import compressai.zoo  # noqa: F401
    return mod
