"""DEVELOPMENT TOOL (see include/cuda_runtime.h): runs a script of the emulation copy with the torch device shim loaded first.
  python emu_exec.py <script> [arguments]"""
import runpy
import sys

import emu_plugin  # noqa: F401

sys.argv = sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
