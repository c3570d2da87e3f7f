#!/usr/bin/env python
"""Smallest run that exercises every kernel of the large-N path (blocked reduction, multishift
iteration: chase / DMMA updates with fused scan / shifts / final blocks) - for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import psd_b200, psd_rng, psd_checks as K
n, p = int(sys.argv[1]) if len(sys.argv) > 1 else 192, int(sys.argv[2]) if len(sys.argv) > 2 else 3
A = psd_rng.gen_uniform(5, n, p, 1)
T, Z, lam, info = psd_b200.pschur_batched(A, "R")
print("info", info, psd_b200.default_handle().large_stats())
print(K.pschur_check(A[0], T[0], Z[0], lam[0], tol=64, check_lambda=False))
