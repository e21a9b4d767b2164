#!/usr/bin/env python
"""Drop-in for /root/reference/run_semantic_super.py:8-24 (SemanticSuPerOptions defaults: superv2, seg inputs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from super_b200.options import SemanticSuPerOptions
from run_super import main

if __name__ == "__main__":
    main(options=SemanticSuPerOptions)
