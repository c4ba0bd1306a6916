import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
  """The product has no fallback: make sure libsntc.so exists (built in-tree by __graft_entry__.build)."""
  import __graft_entry__ as g
  g.build()


@pytest.fixture(scope="session")
def gpu_ctx():
  from shallow_ntc_b200 import Context
  return Context(0)
