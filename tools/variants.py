#!/usr/bin/env python3
"""Tuning builds of libe1b200.so: python tools/variants.py 0 1 2 ... -> galileo-sdr-sim_b200/lib/libe1b200_v<N>.so
(compiled with -DE1_VARIANT=<N>; select one at run time with E1B200_LIB=<path>)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
import build as B  # noqa: E402

for v in sys.argv[1:]:
    print(B.build_lib(force=True, defines=[f"E1_VARIANT={v}"], out=B.PKG / "lib" / f"libe1b200_v{v}.so"))
