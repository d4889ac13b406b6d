"""LTE constants shared by the Python helpers: the 188 turbo code block lengths of TS 36.212 Table 5.1.3-3, read from the
same table the library is compiled with (csrc/qpp_table.inc: {K, f1, f2} rows)."""
from __future__ import annotations

import os
import re

_INC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "qpp_table.inc")
QPP_ROWS = [tuple(int(x) for x in m.groups()) for m in re.finditer(r"\{\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\}", open(_INC).read())]
assert len(QPP_ROWS) == 188
CB_SIZES = [r[0] for r in QPP_ROWS]
