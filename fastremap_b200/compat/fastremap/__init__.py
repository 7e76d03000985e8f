"""`import fastremap` shim: put this directory's parent (fastremap_b200/compat) on sys.path / PYTHONPATH in front of an installed
seung-lab/fastremap and existing code that says `import fastremap` runs on the B200 path unchanged."""
from fastremap_b200 import *  # noqa: F401,F403
from fastremap_b200 import FastremapB200Error, __all__, __version__  # noqa: F401
