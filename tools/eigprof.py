"""Cycle breakdown of the block-Jacobi eigensolver (one 400x400 problem) from its in-kernel counters.

Build the instrumented copy first (`make -C r-tucker_b200/csrc eigprof`), then run this on the GPU box: the kernel
prints, for CTA 5, phase A split into load / visit / store, the two grid syncs, phase B, and per inner round of a
visit the cycles of the three roles (warp 0: next rotations, warp 1: fp64 renormalisation, warps 2-7: apply).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtucker_b200._lib as m
m.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_prof", "librt_prof.so")
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "eigdist.py")).read())
