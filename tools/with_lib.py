"""Run a script of this repo against another build of the library: python tools/with_lib.py <lib.so> <script.py> [args...]
(A/B timing of two builds on the same GPU box.)"""
import os, runpy, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtucker_b200._lib as m
m.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
