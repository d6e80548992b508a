"""Import-level stand-in for timm (sc2bench/models/registry.py:1, backbone.py:5).  TEST INFRASTRUCTURE ONLY."""
from . import models  # noqa: F401
