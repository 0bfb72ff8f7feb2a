import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
OUT = os.path.join(ROOT, "examples", "out")
os.makedirs(OUT, exist_ok=True)
