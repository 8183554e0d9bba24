#!/bin/sh
# compute-sanitizer memcheck over small end-to-end runs (graph k=25/31/55, kmer-set k=33, multi-batch).
# Run on the GPU box: gpurun -- sh tools/sanitize.sh
exec compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py
