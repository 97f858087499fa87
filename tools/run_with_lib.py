#!/usr/bin/env python3
"""Run a Python script with pollen_b200 bound to another build of libflatgfa.so (e.g. the ASAN/UBSAN
one from `make asan`).   usage: run_with_lib.py <lib.so> <script.py> [args...]"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pollen_b200.binding as b  # noqa: E402

b._LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
