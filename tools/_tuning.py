"""The sweep tools A/B restart / launch-plan constants through BLP_* environment variables, which only a
library built with -DBLP_TUNING reads (the production libblp.so compiles the defaults in). Importing this
module BEFORE simple_mip_solver_b200.engine builds such a library under gpurun_out/ and selects it."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simple_mip_solver_b200 import _build  # noqa: E402

_out = ROOT / 'gpurun_out' / 'libblp_tuning.so'
_out.parent.mkdir(exist_ok=True)
if not _out.exists() or any(p.stat().st_mtime > _out.stat().st_mtime for p in _build.SOURCES + _build.DEPS):
    _build.build_extension(tuning=True, out=_out)
os.environ['BLP_LIB'] = str(_out)
