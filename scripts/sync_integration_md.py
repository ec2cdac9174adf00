#!/usr/bin/env python3
"""Refresh the code block of INTEGRATION.md §1 from include/small_gicp/registration/reduction_cuda.hpp (the file tests/host_ref compiles);
tests/test_host_ref_dropin.py::test_integration_md_shows_the_compiled_header fails when the two drift apart."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hdr = open(os.path.join(ROOT, "include", "small_gicp", "registration", "reduction_cuda.hpp")).read().strip()
path = os.path.join(ROOT, "INTEGRATION.md")
doc = open(path).read()
block = "<!-- BEGIN reduction_cuda.hpp -->\n```cpp\n" + hdr + "\n```\n<!-- END reduction_cuda.hpp -->"
new, n = re.subn(r"<!-- BEGIN reduction_cuda.hpp -->.*?<!-- END reduction_cuda.hpp -->", lambda m: block, doc, flags=re.S)
assert n == 1, "markers not found"
open(path, "w").write(new)
print("INTEGRATION.md refreshed" if new != doc else "INTEGRATION.md up to date")
