"""`import gvom` shim: put this directory on PYTHONPATH and the reference's
scripts/gvom_ros.py (`import gvom; gvom.Gvom(...)`, gvom_ros.py:4,44) runs
against the B200 implementation unchanged."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gvom_b200.gvom import Gvom  # noqa: E402,F401
