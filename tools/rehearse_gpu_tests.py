#!/usr/bin/env python
"""Dress rehearsal of the `-m gpu` tests without a GPU: runs them against the SIMT-interpreter build of the product library
(tests/_emu.py), which executes the same CUDA sources on the CPU.  It proves the test bodies themselves (arguments, shapes, oracle
calls, comparisons) before GPU minutes are spent; it says nothing about speed, and the largest cases are too slow for it.

usage: python tools/rehearse_gpu_tests.py [pytest args]      e.g.  tests/test_wavefront.py -k "not gate_render"
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "tests"), str(ROOT)]

import pytest  # noqa: E402

import _emu  # noqa: E402

with _emu.emulated_backend():
    sys.exit(pytest.main(["-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", *sys.argv[1:]]))
