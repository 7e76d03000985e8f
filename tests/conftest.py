import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
  """The CPU restatement (oracle/oracle.py + fastremap_oracle.c) -- test infrastructure only."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as orc
  orc.build()
  return orc


@pytest.fixture(scope="session")
def reference():
  """The compiled, unmodified reference (oracle/_ref) if it was built (it travels to the GPU box)."""
  ref_dir = os.path.join(ROOT, "oracle", "_ref")
  if not os.path.isdir(os.path.join(ref_dir, "fastremap")):
    pytest.skip("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
  import importlib.util
  spec = importlib.util.spec_from_file_location(
    "fastremap", os.path.join(ref_dir, "fastremap", "__init__.py"),
    submodule_search_locations=[os.path.join(ref_dir, "fastremap")])
  mod = importlib.util.module_from_spec(spec)
  saved = sys.modules.get("fastremap")
  sys.modules["fastremap"] = mod
  try:
    spec.loader.exec_module(mod)
  except Exception as e:  # e.g. a different Python ABI on another box
    if saved is not None:
      sys.modules["fastremap"] = saved
    else:
      sys.modules.pop("fastremap", None)
    pytest.skip("oracle/_ref present but not importable here: %r" % (e,))
  return mod


@pytest.fixture(scope="session")
def frb():
  """The product package, on a GPU."""
  import torch
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  import fastremap_b200
  return fastremap_b200
